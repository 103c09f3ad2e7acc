/* Typedef-only stand-in for <GL/gl.h>, pulled in by cuda_gl_interop.h (src/sensor/image_kernels.cu:5).
 * No GL function is called on the hot path. */
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
