/* Force-included (-include) when compiling the reference's svo.cu: CUDA 12 removed the legacy
 * texture<>/surface<> reference templates and svo.cu:19-20 still declares two (dead) ones.
 * Supplying empty templates lets the file compile UNMODIFIED from where it lies. */
#pragma once
template <class T, int N> struct texture {};
template <class T, int N> struct surface {};
