// osl_integrate.cu -- depth map / point cloud / voxel grid  ->  sparse voxel octree, one frame per call.
//
// Replaces the reference's svoFromPointCloud / svoFromVoxelGrid (svo.cu:584-696) together with the per-frame
// image kernels that feed it (image_kernels.cu:24-53,206-215).  Same results (node indices, node words), very
// different structure -- see DESIGN.md section 3:
//
//   k_emit        back-project + pose + Morton key per input; tile-local de-duplication in a shared-memory hash
//   k_sort_bucket barrier-free sort of the (small) de-duplicated key list: splitter ranges, one CTA per range
//   k_sort        cooperative multi-CTA LSD radix sort (8-bit digits) for large lists (voxel grids, first frames)
//   k_structure   cooperative: per sorted key the common-prefix length with its predecessor (=> which tree levels it
//                 heads) and the frontier depth from ONE walk of the pre-frame tree; per-CTA counter vectors behind
//                 flags; allocation plan in the reference's order (pass = depth - frontier depth, then numeric key);
//                 dense per-level node lists with deterministic node / child-tile indices; child pointers and the
//                 value words of new tiles written here -- after this kernel the tree's STRUCTURE (word0) is final
//   k_levels      cooperative, bottom-up, VALUES only (word1): leaf blends, integer mean / max per touched node
// A frame is 4 kernel launches, asynchronous; in pipelined mode the four stages of consecutive frames overlap on four
// streams (osl_run_integrate).  The host never waits for the device unless the pool has to grow, the caller asks for
// sizes / counters, or it runs more than 3 frames ahead.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include "osl_internal.cuh"

namespace cg = cooperative_groups;

#define FULL 0xFFFFFFFFu

// phase checkpoints (SM clock of CTA 0 / thread 0) for tools/phase_profile.py; one predicated store each
__device__ unsigned long long g_osl_prof[128];
__device__ unsigned long long g_osl_ctaprof[4][1024];  // per CTA of the big-input structure stage: SM clock at the start / end of A, start / end of C
// (`bid` = the CTA's index inside its role: the bodies below run as kernels of their own and as roles of k_frame)
#define PROF(i) do { if (bid == 0 && threadIdx.x == 0) g_osl_prof[i] = (unsigned long long)clock64(); } while (0)

// ------------------------------------------------------------------------------------------------ k_emit
// One CTA per 64x32-pixel tile (mode 0) or per 2048 consecutive inputs (modes 1, 2); 4 inputs per thread (16 warps per
// CTA hide the latency of the dependent float descent better than 8 inputs on 8 warps).
// Back-projection + pose + Morton key in registers, then the tile's keys are DE-DUPLICATED in a shared-memory hash
// table (64-bit CAS on the key, atomicMin on the input index): a 1 cm leaf is seen by ~20 neighbouring pixels of a
// 640x480 frame, so ~2048 pixels collapse to ~150 (key, lowest pixel) entries before anything is sorted.  Tiles
// append their entries to the key list with one atomicAdd; the order of the list is irrelevant because the sort is by
// key and k_structure takes the MINIMUM payload of every run of equal keys (canonical Q7: lowest pixel wins).
// Mode 2 (voxel grid, Q11: colour j goes to the j-th smallest key) keeps its duplicates: k_emit_grid below.
#define EMIT_THREADS 512
#define EMIT_PPT 4
#define EMIT_TILE (EMIT_THREADS * EMIT_PPT)
#define EMIT_TW 64
#define EMIT_TH 32
#define EMIT_SLOTS 4096
#define EMIT_EMPTY 0xFFFFFFFFFFFFFFFFull
#define EMIT_SMEM (EMIT_SLOTS * 8 + EMIT_SLOTS * 4 + EMIT_TILE * 2 + 16 + EMIT_TILE * 2 + OSL_BUCKETS * 16)

// `split` != NULL: the tile's entries are ALSO filed by splitter range (the OSL_BUCKETS key ranges of k_sort_bucket)
// into bkeys / bpay[range][OSL_BUCKET_CAP], so that the sort reads its range directly instead of scanning the list.
__device__ __forceinline__ void emit_body(const EmitParams& p, const TreeParams& tp, int vec_ok, u64* __restrict__ keys,
                                          u32* __restrict__ pay, FrameState* fs,
                                          int parity, int bid, unsigned char* s_raw, const u64* __restrict__ split,
                                          u64* __restrict__ bkeys, u32* __restrict__ bpay) {
  u64* s_key = reinterpret_cast<u64*>(s_raw);
  u32* s_pay = reinterpret_cast<u32*>(s_raw + EMIT_SLOTS * 8);
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_raw + EMIT_SLOTS * 12);
  u32* s_misc = reinterpret_cast<u32*>(s_raw + EMIT_SLOTS * 12 + EMIT_TILE * 2);
  u32 &s_count = s_misc[0], &s_valid = s_misc[1], &s_base = s_misc[2];
  const int tid = threadIdx.x, lane = tid & 31;
  PROF(0);

  {
    uint4* k4 = reinterpret_cast<uint4*>(s_key);
    for (int i = tid; i < EMIT_SLOTS / 2; i += EMIT_THREADS) k4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    uint4* p4 = reinterpret_cast<uint4*>(s_pay);
    for (int i = tid; i < EMIT_SLOTS / 4; i += EMIT_THREADS) p4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  if (tid == 0) { s_count = 0; s_valid = 0; }

  u64 k[EMIT_PPT];
  u32 vmask = 0;
  int first;  // input index of this thread's first element (its 8 elements are consecutive)
  if (p.mode == 0) {
    const int tx = bid % p.tiles_x, ty = bid / p.tiles_x;
    const int y = ty * EMIT_TH + tid / (EMIT_TW / EMIT_PPT), x0 = tx * EMIT_TW + (tid % (EMIT_TW / EMIT_PPT)) * EMIT_PPT;
    first = y * p.w + x0;
    int dv[EMIT_PPT];
    if (y < p.h && vec_ok && x0 + EMIT_PPT <= p.w) {  // one 64-bit load of 4 depth pixels (a warp reads 256 B)
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(p.depth + first));
      dv[0] = q.x & 0xFFFF; dv[1] = q.x >> 16; dv[2] = q.y & 0xFFFF; dv[3] = q.y >> 16;
    } else {
#pragma unroll
      for (int i = 0; i < EMIT_PPT; i++) dv[i] = (y < p.h && x0 + i < p.w) ? (int)__ldg(p.depth + first + i) : 0;
    }
    float M[16];  // the pose: a launch parameter, or (tracked frames) device memory written earlier on the stream
    if (p.M_dev) {
#pragma unroll
      for (int i = 0; i < 16; i++) M[i] = p.M_dev[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) M[i] = p.M[i];
    }
#pragma unroll
    for (int i = 0; i < EMIT_PPT; i++) {
      float X, Y, Z;
      osl_vertex(dv[i], x0 + i, y, p.w, p.h, p.w, p.h, p.fx, p.fy, X, Y, Z);
      osl_transform(M, X, Y, Z);
      const bool ok = osl_key(X, Y, Z, tp, k[i]) && (y < p.h && x0 + i < p.w);
      vmask |= (u32)ok << i;
    }
  } else {
    first = bid * EMIT_TILE + tid * EMIT_PPT;  // (mode 1: point clouds; voxel grids have their own kernel, k_emit_grid)
#pragma unroll
    for (int i = 0; i < EMIT_PPT; i++) {
      const int idx = first + i;
      bool ok = false;
      k[i] = 0;
      if (idx < p.n) {
        const float* q = p.pts + (size_t)p.stride * idx;
        ok = osl_key(__ldg(q), __ldg(q + 1), __ldg(q + 2), tp, k[i]);
      }
      vmask |= (u32)ok << i;
    }
  }
  __syncthreads();
  PROF(1);

  {  // valid inputs of the tile (statistics; mode 2: the number of invalid keys that sort to the front)
    u32 c = __popc(vmask);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0 && c) atomicAdd(&s_valid, c);
  }
#pragma unroll
  for (int i = 0; i < EMIT_PPT; i++) {
    if (!((vmask >> i) & 1u)) continue;
    const u32 src = (u32)(first + i);
    if (i > 0 && ((vmask >> (i - 1)) & 1u) && k[i] == k[i - 1]) continue;  // the earlier element already won
    u32 slot = (u32)((k[i] * 0x9E3779B97F4A7C15ull) >> 52);
    for (;;) {
      const u64 prev = atomicCAS(reinterpret_cast<unsigned long long*>(&s_key[slot]), EMIT_EMPTY, k[i]);
      if (prev == EMIT_EMPTY) s_list[atomicAdd(&s_count, 1u)] = (unsigned short)slot;
      if (prev == EMIT_EMPTY || prev == k[i]) { atomicMin(&s_pay[slot], src); break; }
      slot = (slot + 1) & (EMIT_SLOTS - 1);
    }
  }
  __syncthreads();
  PROF(2);
  const u32 cnt = s_count;
  if (tid == 0) {
    s_base = cnt ? atomicAdd(reinterpret_cast<u32*>(&fs->acc_emit[parity]), cnt) : 0u;
    if (s_valid) atomicAdd(reinterpret_cast<u32*>(&fs->acc_valid[parity]), s_valid);
  }
  __syncthreads();
  const u32 base = s_base;
  for (u32 i = tid; i < cnt; i += EMIT_THREADS) {
    const u32 slot = (u32)s_list[i];
    keys[base + i] = s_key[slot];
    pay[base + i] = s_pay[slot];
  }
  if (split) {
    // file the entries by splitter range: range of a key = number of splitters <= key (k_sort_bucket's lo <= k < hi)
    unsigned short* s_pos = reinterpret_cast<unsigned short*>(s_misc + 4);             // [EMIT_TILE] rank inside (tile, range)
    u32* s_bcnt = reinterpret_cast<u32*>(s_pos + EMIT_TILE);                          // [OSL_BUCKETS]
    u32* s_bbase = s_bcnt + OSL_BUCKETS;                                              // [OSL_BUCKETS]
    u64* s_split = reinterpret_cast<u64*>(s_bbase + OSL_BUCKETS);                     // [OSL_BUCKETS - 1] (+1 pad)
    if (tid < OSL_BUCKETS) { s_bcnt[tid] = 0; s_split[tid] = tid < OSL_BUCKETS - 1 ? __ldg(&split[tid]) : ~0ull; }
    __syncthreads();
    for (u32 i = tid; i < cnt; i += EMIT_THREADS) {
      const u64 k = s_key[s_list[i]];
      int b = 0;
#pragma unroll
      for (int step = OSL_BUCKETS / 2; step > 0; step >>= 1)
        if (b + step - 1 < OSL_BUCKETS - 1 && s_split[b + step - 1] <= k) b += step;
      s_pos[i] = (unsigned short)atomicAdd(&s_bcnt[b], 1u);
    }
    __syncthreads();
    if (tid < OSL_BUCKETS && s_bcnt[tid])
      s_bbase[tid] = atomicAdd(reinterpret_cast<u32*>(&fs->acc_bucket[parity][tid]), s_bcnt[tid]);
    __syncthreads();
    for (u32 i = tid; i < cnt; i += EMIT_THREADS) {
      const u32 slot = (u32)s_list[i];
      const u64 k = s_key[slot];
      int b = 0;
#pragma unroll
      for (int step = OSL_BUCKETS / 2; step > 0; step >>= 1)
        if (b + step - 1 < OSL_BUCKETS - 1 && s_split[b + step - 1] <= k) b += step;
      const u32 at = s_bbase[b] + (u32)s_pos[i];
      if (at < OSL_BUCKET_CAP) {  // (a range that overflows is counted in full and sorted the slow way, from the list)
        bkeys[(size_t)b * OSL_BUCKET_CAP + at] = k;
        bpay[(size_t)b * OSL_BUCKET_CAP + at] = s_pay[slot];
      }
    }
  }
  PROF(3);
}

__global__ void __launch_bounds__(EMIT_THREADS)
k_emit(EmitParams p, TreeParams tp, int vec_ok, u64* __restrict__ keys, u32* __restrict__ pay,
       FrameState* fs, int parity, const u64* __restrict__ split, u64* __restrict__ bkeys,
       u32* __restrict__ bpay) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  emit_body(p, tp, vec_ok, keys, pay, fs, parity, (int)blockIdx.x, s_raw, split, bkeys, bpay);
}

// ------------------------------------------------------------------------------------------------ k_sort
// ------------------------------------------------------------------------------------------------ voxel grids (mode 2)
// svoFromVoxelGrid's inputs keep their duplicates (Q11: colour j goes to the j-th smallest key), so there is nothing to
// de-duplicate: the keys go straight to the key list IN INPUT ORDER, 8 bytes per voxel, and the kernel only notes whether
// the list needs sorting at all (the voxelisers and the extraction emit Morton order; svo.cu:602 is then the identity).
// An invalid voxel (outside the cube, non-finite) leaves a gap marker; the rare grid that has one is compacted by
// k_grid_compact / k_grid_copy_back before the sort.  (The first version staged every tile through the shared-memory
// append of the point-cloud path and wrote a dense copy, the list and a payload nobody reads: 20 bytes per voxel and
// 0.8 ms for the 57.5 M voxels of cfg2.)
#define GRID_THREADS 256
#define GRID_PPT 4
#define GRID_GAP 0xFFFFFFFFFFFFFFFFull

// Closed form of computeKey (svo.cu:33-66) with a certificate.  The reference's key is D rounds of "compare with the
// centre, step the centre by +-e" in FP32; every centre it compares against is lo + (integer) * cell, computed with at
// most D roundings of half an ulp(|centre| + half) each.  So when a coordinate is farther than that from every cell
// boundary, all D comparisons of the float descent agree with exact arithmetic and the cell index is
// floor((p - lo) / cell).  That quotient is evaluated in FP32 fixed point here (one subtract, one multiply, one
// conversion per axis); its own rounding errors and the rounding of lo widen the band (make_grid_fast).  Coordinates
// inside the band (a few per cent of a surface's voxels at depth 12; all of them when the tree is too deep for FP32) are
// not decided here: their voxel goes on a list and k_grid_fix runs the reference's descent for it.  [The descent costs
// ~18 instructions per level and voxel and made k_emit_grid issue-bound: 0.75 ms for 57.5 M voxels.]
struct GridFast {
  float lox, loy, loz;  // fl(centre - half)
  float scale;          // 2^sh / cell: one multiply gives the cell index and sh fraction bits in fixed point
  u32 tol;              // the band around every cell boundary, in 2^-sh cells
  int sh;               // 31 - D
  int G;                // cells per axis
  int on;
};
__device__ __forceinline__ bool grid_fast_axis(float p, float lo, const GridFast& gf, int& idx) {
  const float s = (p - lo) * gf.scale;
  idx = 0;
  if (!(s >= 0.0f)) return true;  // below the cube, or NaN (y only, Q1): every comparison of the descent is false
  if (s >= 2147483648.0f) { idx = gf.G - 1; return true; }  // above the cube (or +INF): every comparison true
  const u32 u = __float2uint_rz(s);
  const u32 fr = u & ((1u << gf.sh) - 1u);
  idx = (int)(u >> gf.sh);
  return fr > gf.tol && fr < (1u << gf.sh) - gf.tol;
}
__device__ __forceinline__ u32 grid_spread10(u32 v) {
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
// digit = x + 2y + 4z, most significant level first; sp = the 1024-entry table of grid_spread10 in shared memory (six
// look-ups instead of six 10-instruction bit spreads per voxel)
__device__ __forceinline__ u64 grid_morton(int ix, int iy, int iz, const u32* sp) {
  const u32 lo30 = sp[ix & 1023] | (sp[iy & 1023] << 1) | (sp[iz & 1023] << 2);
  const u32 hi30 = sp[((u32)ix >> 10) & 1023] | (sp[((u32)iy >> 10) & 1023] << 1) | (sp[((u32)iz >> 10) & 1023] << 2);
  return ((u64)hi30 << 30) | (u64)lo30;
}
// key of one voxel: true = decided (valid or a gap), false = valid but inside the uncertainty band (key provisional)
__device__ __forceinline__ bool grid_key(float x, float y, float z, const TreeParams& tp, const GridFast& gf, u64& key,
                                         bool& ok, const u32* sp) {
  if (!gf.on) { ok = osl_key(x, y, z, tp, key); return true; }
  ok = isfinite(x) && isfinite(z);
  int ix, iy, iz;
  const bool cx = grid_fast_axis(x, gf.lox, gf, ix);
  const bool cy = grid_fast_axis(y, gf.loy, gf, iy);
  const bool cz = grid_fast_axis(z, gf.loz, gf, iz);
  key = grid_morton(ix, iy, iz, sp);
  return !ok || (cx && cy && cz);
}

__global__ void __launch_bounds__(GRID_THREADS)
k_emit_grid(const float* __restrict__ pts, int stride, int n, TreeParams tp, GridFast gf, u64* __restrict__ keys,
            u32* __restrict__ fixlist, FrameState* fs, int parity) {
  // a warp takes 128 consecutive voxels, lane l the voxels w0 + 32 i + l: loads and stores are contiguous over the warp,
  // the predecessor of a voxel is the neighbouring lane's (one shuffle), and only the warp's very first voxel needs the
  // key of a voxel the warp does not own
  __shared__ u32 s_sp[1024];
  for (int v = threadIdx.x; v < 1024; v += GRID_THREADS) s_sp[v] = grid_spread10((u32)v);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long w0 = ((long long)blockIdx.x * (GRID_THREADS / 32) + (threadIdx.x >> 5)) * (32 * GRID_PPT);
  const bool vec = stride == 4 && (reinterpret_cast<uintptr_t>(pts) & 15) == 0;
  bool bad = false;
  int valid = 0;
  u64 pk = 0;           // key / state of the voxel before this lane's current one
  bool pok = false, pdec = false;
  if (lane == 0 && w0 > 0 && w0 < n) {
    const float* q = pts + (size_t)stride * (size_t)(w0 - 1);
    pdec = grid_key(__ldg(q), __ldg(q + 1), __ldg(q + 2), tp, gf, pk, pok, s_sp);
  }
  u32 und_bal[GRID_PPT];
#pragma unroll
  for (int i = 0; i < GRID_PPT; i++) {
    const long long idx = w0 + 32 * i + lane;
    u64 kk = GRID_GAP;
    bool ok = false, dec = true;
    if (idx < n) {
      float x, y, z;
      if (vec) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pts) + idx);
        x = q.x; y = q.y; z = q.z;
      } else {
        const float* q = pts + (size_t)stride * (size_t)idx;
        x = __ldg(q); y = __ldg(q + 1); z = __ldg(q + 2);
      }
      dec = grid_key(x, y, z, tp, gf, kk, ok, s_sp);
      if (!ok) { kk = GRID_GAP; bad = true; }  // an invalid voxel: the list has a gap, it is compacted and sorted
      keys[idx] = kk;
    }
    // the predecessor: the lane below; lane 0 takes lane 31's previous voxel (or the one before the warp's range)
    const u64 nk = __shfl_up_sync(FULL, kk, 1);
    const bool nok = __shfl_up_sync(FULL, (int)ok, 1) != 0, ndec = __shfl_up_sync(FULL, (int)dec, 1) != 0;
    if (lane > 0) { pk = nk; pok = nok; pdec = ndec; }
    // (order is checked here between decided neighbours only; k_grid_fix_check looks at the others)
    if (idx < n && ok && dec && pok && pdec && kk < pk) bad = true;
    valid += ok ? 1 : 0;
    und_bal[i] = __ballot_sync(FULL, idx < n && ok && !dec);
    // lane 0's predecessor for the next round: this round's lane 31
    const u64 lk = __shfl_sync(FULL, kk, 31);
    const bool lok = __shfl_sync(FULL, (int)ok, 31) != 0, ldec = __shfl_sync(FULL, (int)dec, 31) != 0;
    if (lane == 0) { pk = lk; pok = lok; pdec = ldec; }
  }
  {  // undecided voxels -> the list (one atomicAdd per warp)
    int tot = 0;
#pragma unroll
    for (int i = 0; i < GRID_PPT; i++) tot += __popc(und_bal[i]);
    if (tot) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&fs->acc_bucket[parity][0], tot);
      base = __shfl_sync(FULL, base, 0);
#pragma unroll
      for (int i = 0; i < GRID_PPT; i++) {
        if ((und_bal[i] >> lane) & 1u) fixlist[base + __popc(und_bal[i] & lanemask_lt())] = (u32)(w0 + 32 * i + lane);
        base += __popc(und_bal[i]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(FULL, valid, o);
  __shared__ int s_valid;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  if (lane == 0 && valid) atomicAdd(&s_valid, valid);
  if (__any_sync(FULL, bad) && lane == 0) atomicOr(reinterpret_cast<u32*>(&fs->acc_unsorted[parity]), 1u);
  __syncthreads();
  if (threadIdx.x == 0 && s_valid) {
    atomicAdd(&fs->acc_valid[parity], s_valid);
    atomicAdd(&fs->acc_emit[parity], s_valid);
  }
}

// the voxels k_emit_grid could not decide: the reference's float descent, then the order against both neighbours
__global__ void __launch_bounds__(GRID_THREADS)
k_grid_fix(const float* __restrict__ pts, int stride, TreeParams tp, u64* __restrict__ keys, const u32* __restrict__ fixlist,
           const FrameState* fs, int parity) {
  const int m = fs->acc_bucket[parity][0];
  for (int i = blockIdx.x * GRID_THREADS + threadIdx.x; i < m; i += gridDim.x * GRID_THREADS) {
    const u32 j = fixlist[i];
    const float* q = pts + (size_t)stride * (size_t)j;
    u64 kk;
    if (osl_key(__ldg(q), __ldg(q + 1), __ldg(q + 2), tp, kk)) keys[j] = kk;
  }
}
__global__ void __launch_bounds__(GRID_THREADS)
k_grid_fix_check(const u64* __restrict__ keys, int n, const u32* __restrict__ fixlist, FrameState* fs, int parity) {
  const int m = fs->acc_bucket[parity][0];
  bool bad = false;
  for (int i = blockIdx.x * GRID_THREADS + threadIdx.x; i < m; i += gridDim.x * GRID_THREADS) {
    const u32 j = fixlist[i];
    const u64 k = keys[j];
    if (j > 0) { const u64 kp = keys[j - 1]; bad |= kp != GRID_GAP && k < kp; }
    if ((int)j + 1 < n) { const u64 kn = keys[j + 1]; bad |= kn != GRID_GAP && kn < k; }
  }
  if (bad) atomicOr(reinterpret_cast<u32*>(&fs->acc_unsorted[parity]), 1u);
}

// the rare grid with invalid voxels: drop the gap markers (the order of the list does not matter, it is sorted next)
__global__ void __launch_bounds__(GRID_THREADS)
k_grid_compact(const u64* __restrict__ keys, u64* __restrict__ tmp, int n, FrameState* fs, int parity) {
  if (fs->acc_valid[parity] == n) return;
  const int lane = threadIdx.x & 31;
  for (long long j = (long long)blockIdx.x * GRID_THREADS + threadIdx.x; j < (long long)((n + 31) / 32) * 32;
       j += (long long)gridDim.x * GRID_THREADS) {
    const u64 k = j < n ? keys[j] : GRID_GAP;
    const u32 bal = __ballot_sync(FULL, k != GRID_GAP);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(&fs->acc_tiles[parity], __popc(bal));
    base = __shfl_sync(FULL, base, 0);
    if (k != GRID_GAP) tmp[base + __popc(bal & lanemask_lt())] = k;
  }
}
__global__ void __launch_bounds__(GRID_THREADS)
k_grid_copy_back(u64* __restrict__ keys, const u64* __restrict__ tmp, int n, const FrameState* fs, int parity) {
  const int v = fs->acc_valid[parity];
  if (v == n) return;
  for (long long j = (long long)blockIdx.x * GRID_THREADS + threadIdx.x; j < v; j += (long long)gridDim.x * GRID_THREADS)
    keys[j] = tmp[j];
}

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_WARPS (SORT_THREADS / 32)

// One launch sorts (key, payload) by the low 8*passes key bits.  Stable LSD radix sort.  Each CTA owns a CONTIGUOUS
// range of tiles, so a pass needs one grid barrier between "count" and "scatter" and the cross-CTA prefix is just a
// sum over <= gridDim.x per-CTA histograms (no per-tile look-back chain, which serialises when all tiles are
// co-resident as they are for one 640x480 frame).  When a CTA owns a single tile (the per-frame case) the tile lives
// in registers for the whole pass and the warp-level ranking is done BEFORE the barrier, so only the prefix and the
// scatter sit behind it.
__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32* s_wsum, int lane, int warp);

struct SortTile {
  u64 key[SORT_ITEMS];
  u32 val[SORT_ITEMS], rank[SORT_ITEMS];
};

// warp-striped load keeps (warp, item, lane) order == global index order (stability)
__device__ __forceinline__ void sort_load(SortTile& t, const u64* kin, const u32* pin, int base, int lane, int n) {
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const int idx = base + i * 32 + lane;
    const bool ok = idx < n;
    t.key[i] = ok ? kin[idx] : ~0ull;
    t.val[i] = ok ? pin[idx] : 0u;
  }
}

// rank of every item among the items of the same digit in this warp (match-any), warp digit counts in whist
__device__ __forceinline__ void sort_rank(SortTile& t, u32* whist, int base, int lane, int n, int shift, u32 lt) {
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const bool ok = (base + i * 32 + lane) < n;
    const u32 digit = ok ? ((u32)(t.key[i] >> shift) & 0xFFu) : 256u;
    const u32 peers = __match_any_sync(FULL, digit);
    const int leader = __ffs(peers) - 1;
    u32 old = 0;
    if (lane == leader && ok) {
      old = whist[digit];
      whist[digit] = old + __popc(peers);
    }
    old = __shfl_sync(FULL, old, leader);
    t.rank[i] = old + __popc(peers & lt);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(SORT_THREADS, 3)
k_sort(u64* kA, u32* pA, u64* kB, u32* pB, u32* cta_hist, const FrameState* fs, int passes, int parity, int mode) {
  cg::grid_group grid = cg::this_grid();
  if (mode == 2 && fs->acc_unsorted[parity] == 0) return;  // the key list is already sorted (uniform exit)
  __shared__ u32 s_hist[256];
  __shared__ u32 s_whist[SORT_WARPS][256];
  __shared__ u32 s_run[256];
  __shared__ u32 s_base[256];
  __shared__ u32 s_goff[256];
  __shared__ u32 s_wsum[SORT_WARPS];
  __shared__ u64 s_skey[SORT_TILE];  // multi-tile path: the tile staged in digit order
  __shared__ u32 s_sval[SORT_TILE];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = fs->acc_emit[parity];
  const int tiles = (n + SORT_TILE - 1) / SORT_TILE;
  const int G = gridDim.x;
  const int per = (tiles + G - 1) / G;
  const int t0 = min(tiles, (int)blockIdx.x * per), t1 = min(tiles, t0 + per);
  const u32 lt = lanemask_lt();
  const bool single = (per <= 1);
  const int Gused = per > 0 ? (tiles + per - 1) / per : 0;  // CTAs that own tiles
  SortTile T;

  for (int pass = 0; pass < passes; pass++) {
    const int shift = 8 * pass;
    const u64* kin = (pass & 1) ? kB : kA;
    const u32* pin = (pass & 1) ? pB : pA;
    u64* kout = (pass & 1) ? kA : kB;
    u32* pout = (pass & 1) ? pA : pB;
    const int base1 = t0 * SORT_TILE + warp * (32 * SORT_ITEMS);

    // phase 1: digit histogram of this CTA's tiles
    if (single) {
#pragma unroll
      for (int w = 0; w < SORT_WARPS; w++) s_whist[w][tid] = 0;
      __syncthreads();
      if (t0 < t1) {
        sort_load(T, kin, pin, base1, lane, n);
        sort_rank(T, s_whist[warp], base1, lane, n, shift, lt);
      }
      __syncthreads();
      u32 sum = 0;
#pragma unroll
      for (int w = 0; w < SORT_WARPS; w++) {
        const u32 v = s_whist[w][tid];
        s_whist[w][tid] = sum;
        sum += v;
      }
      if ((int)blockIdx.x < Gused) cta_hist[blockIdx.x * 256 + tid] = sum;
    } else {
      s_hist[tid] = 0;
      __syncthreads();
      for (int tile = t0; tile < t1; tile++) {
        const int base = tile * SORT_TILE;
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
          const int idx = base + i * SORT_THREADS + tid;
          if (idx < n) atomicAdd(&s_hist[(u32)(kin[idx] >> shift) & 0xFFu], 1u);
        }
      }
      __syncthreads();
      if ((int)blockIdx.x < Gused) cta_hist[blockIdx.x * 256 + tid] = s_hist[tid];
    }
    grid.sync();

    // phase 2: global base of digit `tid` for this CTA = (all smaller digits) + (same digit in earlier CTAs)
    u32 before = 0, tot = 0;
    {
      int c = 0;
      for (; c + 16 <= Gused; c += 16) {  // 16 independent loads in flight
        u32 v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = __ldcg(&cta_hist[(c + k) * 256 + tid]);
#pragma unroll
        for (int k = 0; k < 16; k++) {
          tot += v[k];
          if (c + k < (int)blockIdx.x) before += v[k];
        }
      }
      for (; c < Gused; c++) {
        const u32 v = __ldcg(&cta_hist[c * 256 + tid]);
        tot += v;
        if (c < (int)blockIdx.x) before += v;
      }
    }
    u32 incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
      if (w < warp) woff += s_wsum[w];
    s_run[tid] = woff + (incl - tot) + before;
    __syncthreads();

    if (single) {
      if (t0 < t1) {
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
          if ((base1 + i * 32 + lane) < n) {
            const u32 digit = (u32)(T.key[i] >> shift) & 0xFFu;
            const u32 pos = s_run[digit] + s_whist[warp][digit] + T.rank[i];
            kout[pos] = T.key[i];
            pout[pos] = T.val[i];
          }
        }
      }
    } else {
      for (int tile = t0; tile < t1; tile++) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) s_whist[w][tid] = 0;
        __syncthreads();
        const int base = tile * SORT_TILE + warp * (32 * SORT_ITEMS);
        sort_load(T, kin, pin, base, lane, n);
        sort_rank(T, s_whist[warp], base, lane, n, shift, lt);
        __syncthreads();
        {  // digit `tid`: exclusive scan over the warps, advance the running base; tile-local order of the digits
          u32 sum = 0;
#pragma unroll
          for (int w = 0; w < SORT_WARPS; w++) {
            const u32 v = s_whist[w][tid];
            s_whist[w][tid] = sum;
            sum += v;
          }
          const u32 b = s_run[tid];
          s_run[tid] = b + sum;
          const u32 lex = block_excl_scan_256(sum, s_wsum, lane, warp);  // (contains a block barrier)
          s_base[tid] = lex;      // first tile-local slot of the digit
          s_goff[tid] = b - lex;  // global position = s_goff[digit] + tile-local slot
        }
        __syncthreads();
        // stage the tile in digit order in shared memory, then write it out linearly: runs of equal digits go to
        // consecutive global addresses (coalesced) instead of one 32-byte sector per 8-byte key
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
          if ((base + i * 32 + lane) < n) {
            const u32 digit = (u32)(T.key[i] >> shift) & 0xFFu;
            const u32 lp = s_base[digit] + s_whist[warp][digit] + T.rank[i];
            s_skey[lp] = T.key[i];
            s_sval[lp] = T.val[i];
          }
        }
        __syncthreads();
        const int cnt = min(SORT_TILE, n - tile * SORT_TILE);
        for (int l = tid; l < cnt; l += SORT_THREADS) {
          const u64 k = s_skey[l];
          const u32 pos = s_goff[(u32)(k >> shift) & 0xFFu] + (u32)l;
          kout[pos] = k;
          pout[pos] = s_sval[l];
        }
        __syncthreads();
      }
    }
    grid.sync();
  }
}

// ------------------------------------------------------------------------------------------------ k_sort_bucket
// The per-frame sort when the key list is small (after de-duplication a 640x480 frame is ~20-30 k entries): no grid
// barrier at all.  The key space is cut into BK_BUCKETS contiguous ranges by splitters taken from the sorted keys of
// an earlier frame (k_structure writes them; consecutive frames see almost the same surface, so the ranges stay
// balanced).  CTA b scans the whole list (L2-resident), keeps the entries of its range in shared memory, counts the
// entries of lower ranges (= its output offset), sorts its <= 2048 entries with a shared-memory LSD radix sort that
// skips the digits on which all its keys agree (a contiguous key range shares its high digits), and writes them out.
// A bucket that does not fit (scene cut, first frames) is sorted by the same CTA through global memory: slow but
// correct, and the next frame gets fresh splitters.
#define BK_BUCKETS OSL_BUCKETS
#define BK_THREADS SORT_THREADS
#define BK_CAP SORT_TILE
static_assert(BK_CAP == OSL_BUCKET_CAP, "bucket capacity");
#define BK_SMEM (2 * BK_CAP * 8 + 2 * BK_CAP * 4 + (int)sizeof(BucketShared))

struct BucketShared {
  u32 whist[SORT_WARPS][256];
  u32 run[256];
  u32 base[256];
  u32 wsum[SORT_WARPS];
  u32 cnt, below, or_lo, or_hi, and_lo, and_hi;
};

__device__ __forceinline__ void sort_load_s(SortTile& t, const u64* kin, const u32* pin, int base, int lane, int n) {
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const int idx = base + i * 32 + lane;
    const bool ok = idx < n;
    t.key[i] = ok ? kin[idx] : ~0ull;
    t.val[i] = ok ? pin[idx] : 0u;
  }
}

// exclusive scan of the 256 per-digit totals held one per thread; returns this digit's exclusive prefix
__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32* s_wsum, int lane, int warp) {
  u32 incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  u32 woff = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++)
    if (w < warp) woff += s_wsum[w];
  return woff + incl - v;
}

// single-CTA stable LSD radix sort of c entries through global memory (k0/p0 <-> k1/p1); result ends in k0/p0
__device__ void cta_sort_global(u64* k0, u32* p0, u64* k1, u32* p1, int c, int passes, BucketShared& S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = lanemask_lt();
  SortTile T;
  int cur = 0;
  for (int pass = 0; pass < passes; pass++) {
    const int shift = 8 * pass;
    u64* kin = cur ? k1 : k0; u32* pin = cur ? p1 : p0;
    u64* kout = cur ? k0 : k1; u32* pout = cur ? p0 : p1;
    S.run[tid] = 0;
    __syncthreads();
    for (int idx = tid; idx < c; idx += BK_THREADS) atomicAdd(&S.run[(u32)(kin[idx] >> shift) & 0xFFu], 1u);
    __syncthreads();
    const u32 tot = S.run[tid];
    const u32 ex = block_excl_scan_256(tot, S.wsum, lane, warp);
    __syncthreads();
    S.run[tid] = ex;
    __syncthreads();
    for (int tile0 = 0; tile0 < c; tile0 += SORT_TILE) {
#pragma unroll
      for (int w = 0; w < SORT_WARPS; w++) S.whist[w][tid] = 0;
      __syncthreads();
      const int base = tile0 + warp * (32 * SORT_ITEMS);
      sort_load_s(T, kin, pin, base, lane, c);
      sort_rank(T, S.whist[warp], base, lane, c, shift, lt);
      __syncthreads();
      {
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
          const u32 v = S.whist[w][tid];
          S.whist[w][tid] = sum;
          sum += v;
        }
        const u32 b = S.run[tid];
        S.base[tid] = b;
        S.run[tid] = b + sum;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < SORT_ITEMS; i++) {
        if ((base + i * 32 + lane) < c) {
          const u32 digit = (u32)(T.key[i] >> shift) & 0xFFu;
          const u32 pos = S.base[digit] + S.whist[warp][digit] + T.rank[i];
          kout[pos] = T.key[i];
          pout[pos] = T.val[i];
        }
      }
      __syncthreads();
    }
    cur ^= 1;
  }
  if (cur) {
    for (int idx = tid; idx < c; idx += BK_THREADS) { k0[idx] = k1[idx]; p0[idx] = p1[idx]; }
    __syncthreads();
  }
}

// (runs on the first BK_THREADS threads of its CTA)
// bkeys != NULL: k_emit filed the list by range -- a range that fits is read directly, the scan is skipped.
__device__ __forceinline__ void bucket_body(const u64* kin, const u32* pin, u64* kout, u32* pout, u64* kscr, u32* pscr,
                                            const FrameState* fs, const u64* __restrict__ split, int passes, int parity,
                                            int bid, unsigned char* s_raw, const u64* __restrict__ bkeys,
                                            const u32* __restrict__ bpay) {
  BucketShared& S = *reinterpret_cast<BucketShared*>(s_raw + 2 * BK_CAP * 8 + 2 * BK_CAP * 4);
  u64* const s_key0 = reinterpret_cast<u64*>(s_raw);                    // [2][BK_CAP]
  u32* const s_pay0 = reinterpret_cast<u32*>(s_raw + 2 * BK_CAP * 8);   // [2][BK_CAP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = lanemask_lt();
  const int n = __ldcg(&fs->acc_emit[parity]);
  const int b = bid;
  const u64 lo = (b == 0) ? 0ull : __ldg(&split[b - 1]);
  const u64 hi = (b == BK_BUCKETS - 1) ? ~0ull : __ldg(&split[b]);
  if (tid == 0) { S.cnt = 0; S.below = 0; S.or_lo = S.or_hi = 0u; S.and_lo = S.and_hi = ~0u; }
  __syncthreads();
  PROF(8);

  bool direct = false;
  if (bkeys) {
    // counts per range are final (k_emit ran in an earlier launch / kernel): own count, and the ranges below = offset
    const int mine = __ldcg(&fs->acc_bucket[parity][b]);
    if (tid < b) {
      const u32 v = (u32)__ldcg(&fs->acc_bucket[parity][tid]);
      if (v) atomicAdd(&S.below, v);
    }
    if (mine <= BK_CAP) {
      direct = true;
      for (int e = tid; e < mine; e += BK_THREADS) {
        s_key0[e] = __ldcg(&bkeys[(size_t)b * BK_CAP + e]);
        s_pay0[e] = __ldcg(&bpay[(size_t)b * BK_CAP + e]);
      }
      if (tid == 0) S.cnt = (u32)mine;
    } else if (tid == 0) {
      S.below = 0;  // (the scan below recounts)
    }
    __syncthreads();
    if (!direct) { if (tid == 0) S.below = 0; __syncthreads(); }
  }
  // scan the whole list: entries of lower buckets are counted, entries of this bucket are kept
  u32 below = 0;
  // (every CTA reads the same L2-resident lines: each starts at a different rotation of the list so that the 64 CTAs
  // spread over the L2 slices instead of queueing on the same line at the same time.  A TMA bulk-copy ring
  // (cp.async.bulk + mbarrier, 3 x 16 KB stages) was measured here too: 8.6 us vs 7.0 us for these register-staged
  // loads -- the limit is that contention, not load issue)
  const int iters = direct ? 0 : (n + BK_THREADS * 16 - 1) / (BK_THREADS * 16);
  const int rot = iters > 0 ? (int)(((long long)b * iters) / BK_BUCKETS) : 0;
  for (int it = 0; it < iters; it++) {  // 16 independent loads in flight per thread
    const int i0 = ((it + rot) % iters) * (BK_THREADS * 16);
    u64 k[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int i = i0 + u * BK_THREADS + tid;
      k[u] = (i < n) ? kin[i] : ~0ull;
    }
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int i = i0 + u * BK_THREADS + tid;
      if (i >= n) continue;
      if (k[u] < lo) {
        below++;
      } else if (k[u] < hi || b == BK_BUCKETS - 1) {
        const u32 pos = atomicAdd(&S.cnt, 1u);
        if (pos < BK_CAP) { s_key0[pos] = k[u]; s_pay0[pos] = (u32)i; }  // the payload is fetched after the scan
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(FULL, below, o);
  if (lane == 0 && below) atomicAdd(&S.below, below);
  __syncthreads();
  const int c = (int)S.cnt;
  const u32 offset = S.below;
  // payloads of the kept entries: one parallel round trip (a load inside the scan loop would stall it every time)
  if (c <= BK_CAP && !direct)
    for (int e = tid; e < c; e += BK_THREADS) s_pay0[e] = pin[s_pay0[e]];
  PROF(9);
  if (c == 0) return;

  if (c > BK_CAP) {
    // slow path: gather this bucket into its own output range, sort it there through global scratch
    __syncthreads();
    if (tid == 0) S.cnt = 0;
    __syncthreads();
    for (int i = tid; i < n; i += BK_THREADS) {
      const u64 k = kin[i];
      if (k >= lo && (k < hi || b == BK_BUCKETS - 1)) {
        const u32 pos = atomicAdd(&S.cnt, 1u);
        kout[offset + pos] = k;
        pout[offset + pos] = pin[i];
      }
    }
    __syncthreads();
    cta_sort_global(kout + offset, pout + offset, kscr + offset, pscr + offset, c, passes, S);
    return;
  }

  // digits on which all keys of the bucket agree need no pass
  {
    u32 olo = 0, ohi = 0, alo = ~0u, ahi = ~0u;
    for (int i = tid; i < c; i += BK_THREADS) {
      const u64 k = s_key0[i];
      olo |= (u32)k; ohi |= (u32)(k >> 32); alo &= (u32)k; ahi &= (u32)(k >> 32);
    }
    olo = __reduce_or_sync(FULL, olo); ohi = __reduce_or_sync(FULL, ohi);
    alo = __reduce_and_sync(FULL, alo); ahi = __reduce_and_sync(FULL, ahi);
    if (lane == 0) {
      atomicOr(&S.or_lo, olo); atomicOr(&S.or_hi, ohi);
      atomicAnd(&S.and_lo, alo); atomicAnd(&S.and_hi, ahi);
    }
  }
  __syncthreads();
  const u64 diff = ((u64)(S.or_hi ^ S.and_hi) << 32) | (u64)(S.or_lo ^ S.and_lo);

  SortTile T;
  int cur = 0;
  const int base = warp * (32 * SORT_ITEMS);
  for (int pass = 0; pass < passes; pass++) {
    const int shift = 8 * pass;
    if (((diff >> shift) & 0xFFull) == 0) continue;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) S.whist[w][tid] = 0;
    __syncthreads();
    sort_load_s(T, s_key0 + cur * BK_CAP, s_pay0 + cur * BK_CAP, base, lane, c);
    sort_rank(T, S.whist[warp], base, lane, c, shift, lt);
    __syncthreads();
    u32 sum = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
      const u32 v = S.whist[w][tid];
      S.whist[w][tid] = sum;
      sum += v;
    }
    const u32 ex = block_excl_scan_256(sum, S.wsum, lane, warp);
    S.run[tid] = ex;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
      if ((base + i * 32 + lane) < c) {
        const u32 digit = (u32)(T.key[i] >> shift) & 0xFFu;
        const u32 pos = S.run[digit] + S.whist[warp][digit] + T.rank[i];
        s_key0[(cur ^ 1) * BK_CAP + pos] = T.key[i];
        s_pay0[(cur ^ 1) * BK_CAP + pos] = T.val[i];
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  PROF(10);
  for (int i = tid; i < c; i += BK_THREADS) {
    kout[offset + i] = s_key0[cur * BK_CAP + i];
    pout[offset + i] = s_pay0[cur * BK_CAP + i];
  }
  PROF(11);
}

__global__ void __launch_bounds__(BK_THREADS)
k_sort_bucket(const u64* kin, const u32* pin, u64* kout, u32* pout, u64* kscr, u32* pscr, const FrameState* fs,
              const u64* __restrict__ split, int passes, int parity, const u64* __restrict__ bkeys,
              const u32* __restrict__ bpay) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  bucket_body(kin, pin, kout, pout, kscr, pscr, fs, split, passes, parity, (int)blockIdx.x, s_raw, bkeys, bpay);
}

// ------------------------------------------------------------------------------------------------ k_analyze
#define AN_THREADS 512
#define AN_WARPS (AN_THREADS / 32)
#define NC_MAX OSL_NCOUNT(OSL_MAXD)

__device__ __forceinline__ int key_digit(u64 key, int D, int d) { return (int)((key >> (3 * (D - d))) & 7ull); }

// Walk the PRE-FRAME tree along `key` (splitKeys, svo.cu:108-142).  Returns the frontier depth s: the depth of the
// first node on the path without the has-children flag, OSL_NONE if the path exists down to the leaf's parent.
// Q3 (svo.cu:123 `>= 15`): when the last digit is 7 the reference also tests the LEAF and reports it for splitting.
// Walk cache: a direct-mapped table (prefix of WC_H digits -> node at depth WC_H), one 64-bit word per entry
// ((prefix << 30) | node: written and read atomically).  A node, once allocated, keeps its index and its ancestors
// keep their has-children flags for the life of the tree, so an entry can never go stale (the host clears the table
// when the pool is replaced: reset / upload / expand).  A hit replaces WC_H dependent loads from the root by one.
#define WC_SLOTS 4096
#define PATH_KEEP 8
#define SHALLOW_SLOTS 32  // keys per CTA whose shallow path (levels above the last PATH_KEEP) is kept as well
__device__ __forceinline__ int walk_cache_depth(int D) { const int h = D - 5 < 11 ? D - 5 : 11; return h >= 3 ? h : 0; }
__device__ __forceinline__ u32 walk_cache_slot(u64 prefix) { return (u32)((prefix * 0x9E3779B97F4A7C15ull) >> 52); }

// Hint table: (prefix of t digits, t) -> the node index an earlier walk found there; hashed, direct-mapped, NOT tagged.
// A walk from the root is a chain of D-1 dependent steps (node -> child tile -> node ...) executed by ONE lane -- 15 L2
// round trips plus their address arithmetic, 5 us for a depth-16 tree -- and every frame has a few dozen keys that need
// it: the first key, the first key of every far-away sub-tree (they head levels above the walk cache), keys below a prefix
// the walk cache does not hold yet.  The slowest key sets the length of phase A.  With the hints the WARP walks such a
// key: lane t-1 takes level t, loads the hinted node of its level and that node's pool word (two round trips for the
// whole path), and the links are checked against each other with one shuffle -- level t+1's hint must EQUAL the child the
// verified level t points to.  A hint is therefore only ever a guess: a wrong, stale or torn one ends the verified part
// of the path there, the owner walks the rest the ordinary way and refreshes the hints; no result depends on the table
// (it needs no clearing, no tags, no ordering between writers).
#define WH_BITS 18
#define WH_SLOTS (1u << WH_BITS)
#define WALK_COOP_MAX 4
__device__ __forceinline__ u32 walk_hint_slot(u64 key, int D, int t) {
  const u64 p = (key >> (3 * (D - t))) + (u64)t * 0xD6E8FEB86659FD93ull;
  return (u32)((p * 0x9E3779B97F4A7C15ull) >> (64 - WH_BITS));
}

// Called by ALL lanes of a warp (active = this lane has a key to walk).  limit: nodes of the pre-frame pool (a hint at or
// beyond it is not followed).
__device__ __forceinline__ int walk_frontier(const u32* __restrict__ pool, u64 key, int D, int quirks, int m, bool active,
                                             u32& start, u64* wcache, u32* s_path, u32* s_shallow, int slot, u32 limit) {
  const int lane = threadIdx.x & 31;
  u32 node = (u32)key_digit(key, D, 1);
  int t = 1;
  const int h = wcache ? walk_cache_depth(D) : 0;
  u32* hints = reinterpret_cast<u32*>(wcache + WC_SLOTS);  // (used when h > 0 only)
  u64 prefix = 0;
  bool fill = false;
  if (active && h && m + 1 >= h) {  // (the node at depth m+1, which phase C needs, lies at or below the cached depth)
    prefix = key >> (3 * (D - h));
    const u64 e = __ldcg(&wcache[walk_cache_slot(prefix)]);
    if ((e >> 30) == prefix) { node = (u32)e & OSL_MASK; t = h; }
    else fill = true;
  }
  // Q3: when the last digit is 7 the LEAF is tested as well
  const bool q3 = quirks && key_digit(key, D, D) == 7;
  int result = OSL_NONE;
  bool done = !active;
  const bool from_root = active && h && t == 1;
  u32 todo = __ballot_sync(FULL, from_root);
  // (a prefix that is new to the walk cache is missed by all of its keys at once: when many lanes start at the root they
  // walk side by side as before -- and leave the hints behind -- rather than queue for the warp)
  if (__popc(todo) > WALK_COOP_MAX) todo = 0;
  while (todo) {  // ---- the warp walks the keys that start at the root, one at a time
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const u64 ks = __shfl_sync(FULL, key, src);
    const int ms = __shfl_sync(FULL, m, src);
    const int ls = __shfl_sync(FULL, q3 ? 1 : 0, src) ? D : D - 1;  // deepest level tested
    const int slot_s = __shfl_sync(FULL, slot, src);
    const int fill_s = __shfl_sync(FULL, fill ? 1 : 0, src);
    const int tl = lane + 1;  // this lane's level
    const bool valid = tl <= ls;
    u32 nd = 0xFFFFFFFFu;
    if (valid) nd = (tl == 1) ? (u32)key_digit(ks, D, 1) : __ldcg(&hints[walk_hint_slot(ks, D, tl)]);
    const bool okh = valid && nd < limit;
    const u32 w = okh ? pool[2 * (size_t)nd] : 0u;
    const u32 nxt = __shfl_down_sync(FULL, nd, 1);
    const u32 child = (w & OSL_MASK) + (u32)key_digit(ks, D, tl < D ? tl + 1 : D);
    const bool link = okh && tl < ls && (w & OSL_FLAG) && child == nxt;
    const u32 linked = __ballot_sync(FULL, link);
    const int c = __ffs(~linked) - 1;  // lanes 0..c-1 are linked: the nodes of levels 1..c+1 are the real ones
    const int tv = c + 1;              // deepest verified level (its pool word is valid: a real node lies below `limit`)
    const u32 wc = __shfl_sync(FULL, w, c);
    const u32 child_c = __shfl_sync(FULL, child, c);
    const bool flag_c = (wc & OSL_FLAG) != 0u;
    // what the sequential walk does at the levels 1..tv
    if (tl <= tv && tl <= D - 1 && (tl < tv || flag_c)) {
      if (D - 1 - tl < PATH_KEEP) s_path[(D - 1 - tl) * AN_THREADS + (threadIdx.x & ~31) + src] = w & OSL_MASK;
      else if (slot_s >= 0) s_shallow[tl * SHALLOW_SLOTS + slot_s] = w & OSL_MASK;
    }
    if (fill_s && tl == h && tl <= tv) {
      const u64 ph = ks >> (3 * (D - h));
      wcache[walk_cache_slot(ph)] = (ph << 30) | (u64)nd;
    }
    const u32 st_s = __shfl_sync(FULL, nd, ms < 31 ? ms : 31);  // node of level ms+1 (when ms+1 <= tv)
    if (lane == src) {
      if (ms + 1 <= tv) start = st_s;
      if (!flag_c) { result = tv; done = true; }          // frontier (for tv == D: the leaf of Q3 without children)
      else if (tv == ls) {                                 // the whole path exists
        if (ms + 1 == D && tv == D - 1) start = child_c;   // (the leaf itself)
        done = true;
      } else {                                             // no (or a wrong) hint below level tv: go on from there
        node = child_c;
        t = tv + 1;
      }
    }
  }
  if (!done) {
    for (; t <= D - 1; t++) {
      if (from_root && t >= 2) hints[walk_hint_slot(key, D, t)] = node;
      if (fill && t == h) wcache[walk_cache_slot(prefix)] = (prefix << 30) | (u64)node;
      if (t == m + 1) start = node;  // the first node this key heads: where phase C resumes the walk
      const u32 w0 = pool[2 * (size_t)node];
      if (!(w0 & OSL_FLAG)) return t;
      // the child tiles along the last PATH_KEEP levels stay in shared memory: phase C walks the same nodes again
      if (D - 1 - t < PATH_KEEP) s_path[(D - 1 - t) * AN_THREADS + threadIdx.x] = w0 & OSL_MASK;
      else if (slot >= 0) s_shallow[t * SHALLOW_SLOTS + slot] = w0 & OSL_MASK;  // (the few keys that head shallow levels)
      node = (w0 & OSL_MASK) + (u32)key_digit(key, D, t + 1);
    }
    if (m + 1 == D) start = node;  // the leaf itself (its whole path exists)
    if (q3) {
      if (from_root) hints[walk_hint_slot(key, D, D)] = node;
      if (!(pool[2 * (size_t)node] & OSL_FLAG)) return D;
    }
  }
  return result;
}

// ------------------------------------------------------------------------------------------------ k_structure
// ONE cooperative kernel builds the frame's structure plan:
//   phase A  (per virtual block of 512 sorted keys) common-prefix length m with the predecessor, lowest payload of
//            each run of equal keys, frontier depth s from a walk of the pre-frame tree, per-block counts of level
//            heads and (s, depth) split buckets
//   exchange every CTA owns a contiguous range of blocks, publishes the counter vector of its range behind an epoch
//            flag and sums all CTAs' vectors itself (totals + its own exclusive prefix) -- one wait, no grid barrier
//            and no per-block counters in global memory
//   phase B2 every CTA derives the allocation plan from the totals (bucket bases in the reference's order:
//            pass = depth - s, then numeric key); CTA 0 publishes the FrameState; overflow => nothing is written
//   phase C  dense per-level node lists with deterministic child-tile indices; tiles allocated this frame are
//            initialised here (svo.cu:272-275) -- they lie beyond the pre-frame pool, which is all phase C reads

__device__ __forceinline__ void analyze_block(int vb, int n, const u64* __restrict__ keys, u32* pay, int mode,
                                              const u32* pool, const TreeParams& tp, uint8_t* __restrict__ m8,
                                              uint8_t* __restrict__ s8, u32* __restrict__ start,
                                              u32* s_ctot, u32* s_cnt, u64* wcache, u32* s_path, u32* s_shallow,
                                              u64& k_out, int& m_out, int& s_out, u32& st_out, int& slot_out,
                                              int has_prev, u64 prev_key, u32 limit) {
  const int D = tp.D, NC = OSL_NCOUNT(D);
  k_out = 0; m_out = D; s_out = OSL_NONE; st_out = 0; slot_out = -1;
  for (int c = threadIdx.x; c < NC; c += AN_THREADS) s_cnt[c] = 0;
  if (threadIdx.x == 0) s_shallow[OSL_MAXD * SHALLOW_SLOTS] = 0u;  // slots handed out
  __syncthreads();
  const int j = vb * AN_THREADS + threadIdx.x;
  if (vb == 0 && threadIdx.x == 0) g_osl_prof[24] = (unsigned long long)clock64();
  u64 k = 0;
  int m = D, slot = -1;
  if (j < n) {
    k = keys[j];
    m = 0;
    if (j > 0 || has_prev) {
      const u64 x = k ^ (j > 0 ? keys[j - 1] : prev_key);
      m = x ? (D - 1 - (63 - __clzll((long long)x)) / 3) : D;
    }
    if (m < D) {
      if (mode != 2) {  // canonical Q7: the lowest input index of the run of equal keys wins (runs are short:
        u32 pm = pay[j];  // one entry per 64x32-pixel tile that saw the leaf)
        for (int jj = j + 1; jj < n && keys[jj] == k; jj++) pm = min(pm, pay[jj]);
        pay[j] = pm;
      }
      // a key that heads levels above the last PATH_KEEP (the first key of a frame, of a far-away sub-tree) takes one
      // of the CTA's few shallow-path slots: phase C then finds every child tile of its path in shared memory
      if (m + 1 < D - PATH_KEEP) {
        slot = (int)atomicAdd(&s_shallow[OSL_MAXD * SHALLOW_SLOTS], 1u);
        if (slot >= SHALLOW_SLOTS) slot = -1;
      }
    }
  }
  const bool uniq = m < D;  // (j < n)
  if (vb == 0 && threadIdx.x == 0) g_osl_prof[25] = (unsigned long long)clock64();
  u32 st = 0;
  const int s = walk_frontier(pool, k, D, tp.quirks, m, uniq, st, wcache, s_path, s_shallow, slot, limit);  // (whole warps)
  if (vb == 0 && threadIdx.x == 0) g_osl_prof[26] = (unsigned long long)clock64();
  if (uniq) {
    slot_out = slot;
    start[j] = st;
    st_out = st;
    atomicAdd(&s_cnt[OSL_CLVL(D, m + 1)], 1u);  // heads every level d > m
    if (s != OSL_NONE) {
      const int lo = (s == D) ? D : max(m + 1, s);
      if (s == D || lo <= D - 1) atomicAdd(&s_cnt[OSL_CBKT(D, s, lo)], 1u);
    }
  }
  if (j < n) {
    m8[j] = (uint8_t)m;
    s8[j] = (uint8_t)s;
    k_out = k; m_out = m; s_out = s;
  }
  __syncthreads();
  // prefix over depth turns "first level headed" histograms into per-level / per-bucket counts
  if (threadIdx.x == 0) {
    u32 run = 0;
    for (int d = 1; d <= D; d++) { run += s_cnt[OSL_CLVL(D, d)]; s_cnt[OSL_CLVL(D, d)] = run; }
  } else if ((int)threadIdx.x <= D) {
    const int s = threadIdx.x;
    u32 run = 0;
    for (int d = s; d <= D; d++) {
      run += s_cnt[OSL_CBKT(D, s, d)];
      s_cnt[OSL_CBKT(D, s, d)] = (d == D && s != D) ? 0u : run;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NC; c += AN_THREADS) s_ctot[c] += s_cnt[c];  // this CTA's running total
  __syncthreads();
}

// Walk the pre-frame tree along `key` down to the leaf's PARENT (depth D-1).  Returns the frontier depth (1..D-1), or
// OSL_NONE when the path exists that far; then ptile is the parent's child tile.  st = the node at depth m+1 when the
// walk gets there (m+1 <= D-1).  The leaf test of Q3 is the caller's.
__device__ __forceinline__ int walk_parent(const u32* __restrict__ pool, u64 key, int D, int m, u64* wcache, u32& st,
                                           u32& ptile) {
  u32 node = (u32)key_digit(key, D, 1);
  int t = 1;
  const int h = wcache ? walk_cache_depth(D) : 0;
  u64 prefix = 0;
  bool fill = false;
  if (h && m + 1 >= h) {
    prefix = key >> (3 * (D - h));
    const u64 e = __ldcg(&wcache[walk_cache_slot(prefix)]);
    if ((e >> 30) == prefix) { node = (u32)e & OSL_MASK; t = h; }
    else fill = true;
  }
  for (; t <= D - 1; t++) {
    if (fill && t == h) wcache[walk_cache_slot(prefix)] = (prefix << 30) | (u64)node;
    if (t == m + 1) st = node;
    const u32 w0 = pool[2 * (size_t)node];
    if (!(w0 & OSL_FLAG)) return t;
    if (t == D - 1) { ptile = w0 & OSL_MASK; break; }
    node = (w0 & OSL_MASK) + (u32)key_digit(key, D, t + 1);
  }
  return OSL_NONE;
}

// Phase A for a CTA that owns MANY blocks (voxel grids of millions of keys): the per-block counts only matter as the
// CTA's total -- phase C recomputes every block's prefix on the way -- so the range is streamed without a single block
// barrier.  A warp takes AN_R * 32 CONSECUTIVE sorted keys, and only the keys that open a new parent (common prefix
// with the predecessor shorter than D-1 digits; the batch's first key always) walk the tree: they are compacted into a
// per-warp list (a surface fills about four of a parent's eight cells, so 128 keys are ~35 walks instead of 128), a lane
// walks one list entry from the walk cache down to the parent, and every key then takes frontier and parent tile from
// the entry of the last opener at or before it.  What is left per key is the leaf test of Q3 (last digit 7 only).
// Openers that miss the walk cache share one walk from the root per prefix; the few that head levels above the cached
// depth take the scalar walk.
#define AN_R 4
#define AN_HEAD_BYTES (AN_WARPS * AN_R * 32 * 9)  // per warp: AN_R*32 entries of 8 bytes (key -> st:ptile) + 1 byte (m -> s)
__device__ __forceinline__ void analyze_stream(int j0, int j1, int n, const u64* __restrict__ keys, u32* pay, int mode,
                                               const u32* pool, const TreeParams& tp, uint8_t* __restrict__ m8,
                                               uint8_t* __restrict__ s8, u32* __restrict__ start, u32* s_ctot,
                                               u32* s_cnt, unsigned char* s_heads, u64* wcache, int has_prev,
                                               u64 prev_key) {
  const int D = tp.D, NC = OSL_NCOUNT(D);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = wcache ? walk_cache_depth(D) : 0;
  const u32 lt = lanemask_lt(), le = lt | (1u << lane);
  u64* s_hk = reinterpret_cast<u64*>(s_heads) + warp * (AN_R * 32);
  uint8_t* s_hm = s_heads + AN_WARPS * AN_R * 32 * 8 + warp * (AN_R * 32);
  for (int c = tid; c < NC; c += AN_THREADS) s_cnt[c] = 0;
  __syncthreads();
  for (int base = j0; base < j1; base += AN_THREADS * AN_R) {
    const int wb = base + warp * (AN_R * 32);  // this warp's keys: [wb, wb + 128)
    u64 k[AN_R];
    int m[AN_R], s[AN_R], hidx[AN_R];
    u32 st[AN_R];
    int nh = 0;  // openers of this batch
    u64 before = 0;  // the key in front of the batch
    bool has_before = false;
    if (lane == 0 && wb < j1) {
      if (wb > 0) { before = keys[wb - 1]; has_before = true; }
      else if (has_prev) { before = prev_key; has_before = true; }
    }
#pragma unroll
    for (int r = 0; r < AN_R; r++) {
      const int j = wb + r * 32 + lane;
      k[r] = j < j1 ? keys[j] : 0ull;
    }
#pragma unroll
    for (int r = 0; r < AN_R; r++) {
      const int j = wb + r * 32 + lane;
      u64 pk = __shfl_up_sync(FULL, k[r], 1);
      bool hp = true;
      if (lane == 0) {
        if (r == 0) { pk = before; hp = has_before; }
      }
      if (r > 0) {
        const u64 last = __shfl_sync(FULL, k[r - 1], 31);
        if (lane == 0) pk = last;
      }
      m[r] = D; s[r] = OSL_NONE; st[r] = 0;
      if (j < j1) {
        const u64 x = k[r] ^ pk;
        m[r] = !hp ? 0 : (x ? (D - 1 - (63 - __clzll((long long)x)) / 3) : D);
      }
      const bool opener = j < j1 && (m[r] < D - 1 || (r == 0 && lane == 0));
      const u32 bal = __ballot_sync(FULL, opener);
      hidx[r] = nh + __popc(bal & le) - 1;
      if (opener) { s_hk[hidx[r]] = k[r]; s_hm[hidx[r]] = (uint8_t)m[r]; }
      nh += __popc(bal);
    }
    __syncwarp();
    if (mode != 2) {  // canonical Q7: the lowest input index of the run of equal keys wins
#pragma unroll
      for (int r = 0; r < AN_R; r++) {
        const int j = wb + r * 32 + lane;
        if (m[r] < D) {
          u32 pm = pay[j];
          for (int jj = j + 1; jj < n && keys[jj] == k[r]; jj++) pm = min(pm, pay[jj]);
          pay[j] = pm;
        }
      }
    }
    // ---- the openers' walks, 32 at a time
    for (int q0 = 0; q0 < nh; q0 += 32) {
      const int idx = q0 + lane;
      const bool active = idx < nh;
      const u64 hk = active ? s_hk[idx] : 0ull;
      const int hm = active ? (int)s_hm[idx] : D;
      const u64 prefix = h ? hk >> (3 * (D - h)) : 0ull;
      const bool cand = active && h && hm + 1 >= h;
      bool slow = active && !cand;
      u64 e = ~0ull;
      if (cand) e = __ldcg(&wcache[walk_cache_slot(prefix)]);
      const bool hit = cand && (e >> 30) == prefix;
      u32 node = hit ? ((u32)e & OSL_MASK) : 0u;
      bool act = hit;
      // A prefix that is not in the table yet is missed by EVERY opener below it in this row (they are neighbours in key
      // order).  One of the lanes that miss the same prefix walks from the root to depth h, enters it, and hands the
      // node to the others.
      const bool miss = cand && !hit;
      const u32 missers = __ballot_sync(FULL, miss);
      if (missers) {
        int leader = lane;
        u32 ndh = 0xFFFFFFFFu;
        if (miss) {
          leader = __ffs(__match_any_sync(missers, prefix)) - 1;
          if (lane == leader) {
            u32 nd = (u32)key_digit(hk, D, 1);
            int t = 1;
            for (; t < h; t++) {
              const u32 w0 = pool[2 * (size_t)nd];
              if (!(w0 & OSL_FLAG)) break;
              nd = (w0 & OSL_MASK) + (u32)key_digit(hk, D, t + 1);
            }
            if (t == h) {
              ndh = nd;
              wcache[walk_cache_slot(prefix)] = (prefix << 30) | (u64)nd;
            }
          }
        }
        ndh = __shfl_sync(FULL, ndh, leader);
        if (miss) {
          if (ndh != 0xFFFFFFFFu) { node = ndh; act = true; }
          else slow = true;  // (the tree ends above depth h on this path)
        }
      }
      int fs = OSL_NONE;
      u32 hst = 0, ptile = 0;
      for (int t = h; t <= D - 1 && act; t++) {
        if (t == hm + 1) hst = node;
        const u32 w0 = pool[2 * (size_t)node];
        if (!(w0 & OSL_FLAG)) { fs = t; act = false; }
        else if (t == D - 1) { ptile = w0 & OSL_MASK; act = false; }
        else node = (w0 & OSL_MASK) + (u32)key_digit(hk, D, t + 1);
      }
      if (slow) fs = walk_parent(pool, hk, D, hm, wcache, hst, ptile);
      if (active) { s_hk[idx] = ((u64)hst << 32) | (u64)ptile; s_hm[idx] = (uint8_t)fs; }
    }
    __syncwarp();
    // ---- every key: frontier and parent tile of its opener; the leaf test of Q3
    u32 leaf[AN_R], lw[AN_R];
#pragma unroll
    for (int r = 0; r < AN_R; r++) {
      leaf[r] = 0; lw[r] = OSL_FLAG;
      if (m[r] < D) {
        const u64 res = s_hk[hidx[r]];
        const int fs = (int)s_hm[hidx[r]];
        const bool own = m[r] < D - 1;  // this key's own entry: st is the node at its depth m+1
        if (fs != (int)OSL_NONE) {
          s[r] = fs;
          st[r] = own ? (u32)(res >> 32) : 0u;
        } else {
          leaf[r] = ((u32)res & OSL_MASK) + (u32)key_digit(k[r], D, D);
          st[r] = own ? (u32)(res >> 32) : leaf[r];
          if (tp.quirks && key_digit(k[r], D, D) == 7) lw[r] = pool[2 * (size_t)leaf[r]];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < AN_R; r++)
      if (!(lw[r] & OSL_FLAG)) s[r] = D;
    __syncwarp();  // (the list is rewritten by the next batch)
#pragma unroll
    for (int r = 0; r < AN_R; r++) {
      const int j = wb + r * 32 + lane;
      const bool uniq = m[r] < D;
      if (j < j1) {
        m8[j] = (uint8_t)m[r];
        s8[j] = (uint8_t)s[r];
        if (uniq) start[j] = st[r];
      }
      // counters: warp-aggregated for the three deepest first-headed levels (nearly every key of a surface), shared
      // atomics for the rest
#pragma unroll
      for (int q = 1; q <= 3; q++) {
        const u32 bal = __ballot_sync(FULL, uniq && m[r] == D - q);
        if (lane == 0 && bal && D - q >= 0) atomicAdd(&s_cnt[OSL_CLVL(D, D - q + 1)], (u32)__popc(bal));
      }
      if (uniq && m[r] < D - 3) atomicAdd(&s_cnt[OSL_CLVL(D, m[r] + 1)], 1u);
      const bool sp = uniq && s[r] != OSL_NONE;
      if (__any_sync(FULL, sp)) {
        const int lo = (s[r] == D) ? D : max(m[r] + 1, s[r]);
        const bool cnt = sp && (s[r] == D || lo <= D - 1);
        const u32 peers = __match_any_sync(FULL, cnt ? (s[r] << 8) | lo : 0xFFFF);
        if (cnt && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[OSL_CBKT(D, s[r], lo)], (u32)__popc(peers));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 run = 0;
    for (int d = 1; d <= D; d++) { run += s_cnt[OSL_CLVL(D, d)]; s_cnt[OSL_CLVL(D, d)] = run; }
  } else if ((int)threadIdx.x <= D) {
    const int sd = threadIdx.x;
    u32 run = 0;
    for (int d = sd; d <= D; d++) {
      run += s_cnt[OSL_CBKT(D, sd, d)];
      s_cnt[OSL_CBKT(D, sd, d)] = (d == D && sd != D) ? 0u : run;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NC; c += AN_THREADS) s_ctot[c] += s_cnt[c];
  __syncthreads();
}

// s_base: this block's exclusive prefix of every counter (shared memory)
__device__ __forceinline__ void assign_block(int vb, int n, const u64* __restrict__ keys, const u32* pay, u32* pool,
                                             const TreeParams& tp, const uint8_t* __restrict__ m8,
                                             const uint8_t* __restrict__ s8, const u32* __restrict__ start,
                                             u32* s_base, const LevelArrays& lv, int mode, u32 size0,
                                             int n_invalid_front, const u32* s_plan, u32 (*s_w)[NC_MAX],
                                             bool carried, u64 k_in, int m_in, int s_in, u32 st_in,
                                             const u32* s_path, const u32* s_shallow, int slot, bool shard,
                                             bool preloaded = false, u32 pay_in = 0u) {
  const int D = tp.D, NC = OSL_NCOUNT(D);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = lanemask_lt();
  const int j = vb * AN_THREADS + tid;
  u64 k = 0;
  int m = D, s = OSL_NONE;
  u32 node = 0;
  if (carried || preloaded) { k = k_in; m = m_in; s = s_in; node = st_in; }  // (a CTA that owns ONE block kept phase A's
  else if (j < n) { k = keys[j]; m = m8[j]; s = s8[j]; node = start[j]; }    // registers; big inputs: loaded one block ahead)
  const bool unique = m < D;
  // the winning input of the leaf (needed at the last level only): in flight while the levels above are laid out
  const u32 paymin = preloaded ? pay_in : (unique && mode != 2 && j < n) ? __ldcg(&pay[j]) : 0u;

  // No lane of this warp heads a level <= the warp's smallest m: the per-level collectives start one level above it
  // (one level early so that par_idx / path_tile of the first headed level come out of the loop itself).
  const int d0 = max(1, (int)__reduce_min_sync(FULL, (unsigned)m));
  // (steady state: the whole path of every key exists -- no lane of the warp splits anything and the per-bucket
  // collectives, three match_any per level, are skipped; when no key of the BLOCK splits, only the D level counters
  // are live and the bucket counters, (D+1)^2 of them, are not even touched)
  const bool any_split = __any_sync(FULL, unique && s != OSL_NONE);
  const int NCu = __syncthreads_or(any_split ? 1 : 0) ? NC : D;
  // pass 1: per-warp totals of every counter this block can touch
  for (int c = lane; c < NCu; c += 32) s_w[warp][c] = 0;
  __syncwarp();
  for (int d = d0; d <= D; d++) {
    const bool f = unique && m < d;
    const u32 bal = __ballot_sync(FULL, f);
    if (lane == 0) s_w[warp][OSL_CLVL(D, d)] = __popc(bal);
    if (any_split) {
      const bool sp = f && s != OSL_NONE && s <= d && (d <= D - 1 || s == D);
      const u32 peers = __match_any_sync(FULL, sp ? s : 0);
      if (sp && lane == __ffs(peers) - 1) s_w[warp][OSL_CBKT(D, s, d)] = __popc(peers);
    }
  }
  __syncthreads();
  if (vb == 0 && tid == 0) g_osl_prof[27] = (unsigned long long)clock64();
  // exclusive scan over the warps + the block's global base + the bucket's global rank base
  for (int c = tid; c < NCu; c += AN_THREADS) {
    u32 run = s_base[c] + s_plan[c];
#pragma unroll
    for (int w = 0; w < AN_WARPS; w++) {
      const u32 v = s_w[w][c];
      s_w[w][c] = run;
      run += v;
    }
    s_base[c] = run - s_plan[c];  // exclusive prefix of the CTA's NEXT block
  }
  __syncthreads();
  // pass 2: down the levels this key heads (d > m), resuming the walk of phase A at depth m+1; existing child tiles
  // come from the pre-frame pool, new ones from their rank.  Every head also knows ITS OWN node index:
  //   first headed level (d = m+1): the node phase A reached (start) when it pre-exists, else its slot in the tile
  //     its parent received this frame -- that tile's rank follows from the same bucket counters, because the
  //     parent's head (an earlier key with the same first m digits) has the same frontier depth s;
  //   deeper levels: slot in the child tile found / allocated one iteration earlier.
  // With it the head writes the child pointer of a node split this frame (svo.cu:269) right here, and k_levels
  // needs no dependent look-ups.  New tiles only get their value words initialised (svo.cu:272-275): the pool
  // beyond the live nodes is kept zero, so their word0 is already 0 and cannot race with the pointer writes.
  if (vb == 0 && tid == 0) g_osl_prof[28] = (unsigned long long)clock64();
  const int s_eff = (s == OSL_NONE) ? D : s;
  const u32 le = lt | (1u << lane);
  // index (in level d-1) of the node on this key's path; below d0 the warp has no heads, so it is the last node
  // headed before this warp
  u32 par_idx = (d0 > 1) ? s_w[warp][OSL_CLVL(D, d0 - 1)] - 1u : 0u;
  u32 path_tile = 0; // tile holding the level-(d) nodes below this key's level-(d-1) node (root: tile 0)
  size_t o_prev = 0;  // where this lane's level-(d-1) entry went (valid when it heads that level)
  // Solo prefix.  The warp's collectives start at its smallest m, but the levels between the smallest and the
  // second-smallest m are headed by ONE lane (typically: the first key of a frame or of a far sub-tree heads a dozen
  // levels on its own, the other 31 lanes only the last three or four).  That lane lays those levels out by itself,
  // rank 0 of the warp at each of them, without a single vote; the warp-wide loop starts where company begins.
  int d_main = d0;
  if (!any_split) {
    const unsigned um = unique ? (unsigned)m : (unsigned)D;
    const unsigned m_min = __reduce_min_sync(FULL, um);
    const u32 who = __ballot_sync(FULL, um == m_min);
    if (__popc(who) == 1) {
      const int solo = __ffs(who) - 1;
      const int m_2nd = (int)__reduce_min_sync(FULL, lane == solo ? (unsigned)D : um);
      // With the path's child tiles in shared memory (a CTA that owns one block) the levels do not depend on each other:
      // lane i lays out level d0 + i (the first key of a frame heads a dozen levels: 2 us for one lane)
      const int slot_s = __shfl_sync(FULL, slot, solo);
      const bool spread = carried && (slot_s >= 0 || D - 1 - d0 < PATH_KEEP);
      if (m_2nd > d0 && spread) {
        const u64 ks = __shfl_sync(FULL, k, solo);
        const u32 node_s = __shfl_sync(FULL, node, solo);
        const u32 par0 = __shfl_sync(FULL, par_idx, solo);
        const int ms = (int)m_min, tid_s = (tid & ~31) + solo;
        auto tile_of = [&](int dd) -> u32 {
          return (D - 1 - dd < PATH_KEEP) ? s_path[(D - 1 - dd) * AN_THREADS + tid_s] : s_shallow[dd * SHALLOW_SLOTS + slot_s];
        };
        const int d = d0 + lane;
        if (d < m_2nd && ms < d) {
          const u32 lbase = s_w[warp][OSL_CLVL(D, d)];
          const u32 ct = tile_of(d);
          const u32 self = (d == ms + 1) ? node_s : tile_of(d - 1) + (u32)key_digit(ks, D, d);
          u32 par = par0;
          if (d > d0) {
            const u32 lb1 = s_w[warp][OSL_CLVL(D, d - 1)];
            par = (ms < d - 1) ? lb1 : lb1 - 1u;
            if (ms < d - 1) lv.fc[lv.off[d - 1] + lb1] = lbase;
          }
          const size_t o = lv.off[d] + lbase;
          lv.ctile[o] = ct;
          lv.digit[o] = (uint8_t)key_digit(ks, D, d);
          lv.par[o] = par;
          lv.self[o] = self;
        }
        if (lane == solo) {  // where the warp-wide loop finds this key
          const int last = m_2nd - 1;
          const u32 lbl = s_w[warp][OSL_CLVL(D, last)];
          if (ms < last) {
            path_tile = tile_of(last);
            node = path_tile + (u32)key_digit(k, D, last + 1);
            o_prev = lv.off[last] + lbl;
            par_idx = lbl;
          } else {
            par_idx = lbl - 1u;
          }
        }
        d_main = m_2nd;
      } else if (m_2nd > d0) {
        if (lane == solo) {
          for (int d = d0; d < m_2nd; d++) {
            const u32 lbase = s_w[warp][OSL_CLVL(D, d)];
            if (m < d) {
              const u32 self = (d == m + 1) ? node : path_tile + (u32)key_digit(k, D, d);
              const u32 ct = (carried && D - 1 - d < PATH_KEEP) ? s_path[(D - 1 - d) * AN_THREADS + tid]
                             : (carried && slot >= 0)         ? s_shallow[d * SHALLOW_SLOTS + slot]
                                                               : (pool[2 * (size_t)node] & OSL_MASK);
              node = ct + (u32)key_digit(k, D, d + 1);
              path_tile = ct;
              const size_t o = lv.off[d] + lbase;
              if (d >= 2 && m < d - 1) lv.fc[o_prev] = lbase;
              o_prev = o;
              lv.ctile[o] = ct;
              lv.digit[o] = (uint8_t)key_digit(k, D, d);
              lv.par[o] = par_idx;
              lv.self[o] = self;
              par_idx = lbase;
            } else {
              par_idx = lbase - 1u;
            }
          }
        }
        d_main = m_2nd;
      }
    }
  }
  if (vb == 0 && tid == 0) { g_osl_prof[96] = (unsigned long long)clock64(); g_osl_prof[97] = (unsigned long long)d_main; }
  for (int d = d_main; d <= D; d++) {
    const bool f = unique && m < d;
    const u32 bal = __ballot_sync(FULL, f);
    bool sp = false;
    u32 peers = 0, same_s = 0, spmask = 0;
    if (any_split) {
      sp = f && s != OSL_NONE && s <= d && (d <= D - 1 || s == D);
      peers = __match_any_sync(FULL, sp ? s : 0);
      // same frontier depth among ALL unique keys of the warp (heads or not), for the parent-tile rank below
      same_s = __match_any_sync(FULL, unique ? s : 0x100 + lane);
      spmask = __ballot_sync(FULL, sp);
    }
    u32 self = 0;
    if (f) self = (d == m + 1 && m < s_eff) ? node : path_tile + (u32)key_digit(k, D, d);
    u32 ct = 0xFFFFFFFFu;
    if (f) {
      if (d < D && d < s_eff) {
        // (a CTA that owns one block finds the child tile phase A read on the same path in shared memory)
        ct = (carried && D - 1 - d < PATH_KEEP) ? s_path[(D - 1 - d) * AN_THREADS + tid]
             : (carried && slot >= 0)         ? s_shallow[d * SHALLOW_SLOTS + slot]
                                               : (pool[2 * (size_t)node] & OSL_MASK);
        node = ct + (u32)key_digit(k, D, d + 1);
      } else if (sp) {
        const u32 rank = s_w[warp][OSL_CBKT(D, s, d)] + __popc(peers & lt);
        const u32 tile = size0 + 8u * rank;
        ct = tile | OSL_NEWBIT;
        u32* tw = pool + 2 * (size_t)tile;
#pragma unroll
        for (int i = 0; i < 8; i++) tw[2 * i + 1] = OSL_EMPTY;
        pool[2 * (size_t)self] = OSL_FLAG | tile;
      }
      path_tile = ct & OSL_MASK;
    } else if (unique && d == m && s != OSL_NONE && s <= d && d <= D - 1) {
      // not a head here, but a head one level down whose parent was split this frame by an earlier key: the
      // parent is the last level-d head with frontier s at or before this lane (or before this warp)
      const u32 rank = s_w[warp][OSL_CBKT(D, s, d)] + __popc(same_s & spmask & le) - 1u;
      path_tile = size0 + 8u * rank;
    }
    const u32 lbase = s_w[warp][OSL_CLVL(D, d)];
    if (f) {
      const size_t o = lv.off[d] + lbase + __popc(bal & lt);
      // the key that heads a node also heads that node's first (lowest-key) touched child: its own entry one level down
      if (d >= 2 && m < d - 1) lv.fc[o_prev] = lbase + __popc(bal & lt);
      o_prev = o;
      lv.ctile[o] = ct;
      lv.digit[o] = (uint8_t)key_digit(k, D, d);
      lv.par[o] = par_idx;
      lv.self[o] = self;
      // sharded build: a node of this rank may live in a tile the LOWER rank allocates (the slice's first key continues
      // below nodes that rank splits); its value word must be "empty" before this rank's value fold blends into it
      if (shard && self >= size0) pool[2 * (size_t)self + 1] = OSL_EMPTY;
      if (d == D) lv.src[lbase + __popc(bal & lt)] = (mode == 2) ? (u32)(n_invalid_front + j) : paymin;
    }
    // the level-d node on this key's path: its own if it heads it, else the last one headed before it
    par_idx = lbase + __popc(bal & le) - 1u;
  }
  if (vb == 0 && lane == 0) {  // (profiling: when each warp of CTA 0's first block left the loop, and its first level)
    g_osl_prof[64 + warp] = (unsigned long long)clock64();
    g_osl_prof[80 + warp] = (unsigned long long)d0 | ((unsigned long long)any_split << 8);
  }
  __syncthreads();
}

__device__ __forceinline__ u32 ld_vol(const u32* p) { return *(const volatile u32*)p; }

struct StructArgs {
  const u64* keys_sorted; const u64* keys_dense; u32* pay; u32* pool; TreeParams tp;
  FrameState* fs; FrameState* fr; FrameState* hr; uint8_t* m8; uint8_t* s8; u32* start; u32* ctatot;
  u32* flags; u32 epoch; LevelArrays lv; int mode; int capacity; int n_in; int parity; u64* split_out;
  u64* wcache;  // walk cache (NULL = off)
  u64* lvltag;  // [CTA][OSL_MAXD] per-level counters of the frame-sized exchange, each word tagged with the epoch
  // one map built by several GPUs (osl_shard_*): this rank's keys are a contiguous slice of the globally sorted list
  int shard;          // 0 normal; 1 analyse only: publish the counters, write this rank's totals, stop; 2 assign only
  int has_prev; u64 prev_key;  // the key before the slice (last key of the lower rank): the first key's predecessor
  const u32* ext;     // shard 2: [2][NC] bucket counters -- exclusive prefix over the lower ranks, totals over all ranks
  u32* rank_tot;      // shard 1: [NC] this rank's totals
  u32 src_base;       // voxel grids: global sorted position of the slice's first key (colour index, Q11)
};
#define STRUCT_MAXG 1024  // most CTAs a structure grid / role may have (s_has)
#define STRUCT_SMEM ((AN_WARPS + 4) * NC_MAX * 4 + AN_WARPS * 4 + 2 * STRUCT_MAXG + 8 * 512 * 4 + (OSL_MAXD * 32 + 4) * 4)

template <bool BIG>
__device__ __forceinline__ void structure_body(const StructArgs& A, int bid, int G, unsigned char* s_raw) {
  const u64* __restrict__ keys_sorted = A.keys_sorted; const u64* __restrict__ keys_dense = A.keys_dense;
  u32* pay = A.pay; u32* pool = A.pool; const TreeParams tp = A.tp;
  FrameState* fs = A.fs; FrameState* fr = A.fr; FrameState* hr = A.hr;
  uint8_t* m8 = A.m8; uint8_t* s8 = A.s8; u32* start = A.start; u32* ctatot = A.ctatot; u32* flags = A.flags;
  const u32 epoch = A.epoch; const LevelArrays& lv = A.lv;
  const int mode = A.mode, capacity = A.capacity, n_in = A.n_in, parity = A.parity;
  u64* split_out = A.split_out;
  u32 (*s_w)[NC_MAX] = reinterpret_cast<u32 (*)[NC_MAX]>(s_raw);
  u32* s_plan = reinterpret_cast<u32*>(s_raw) + AN_WARPS * NC_MAX;
  u32* s_tot = s_plan + NC_MAX;
  u32* s_base = s_tot + NC_MAX;
  u32* s_ctot = s_base + NC_MAX;
  u32* s_scan = s_ctot + NC_MAX;
  u32* s_path = reinterpret_cast<u32*>(s_raw + (AN_WARPS + 4) * NC_MAX * 4 + AN_WARPS * 4 + 2 * STRUCT_MAXG);  // [PATH_KEEP][512]
  u32* s_shallow = s_path + PATH_KEEP * AN_THREADS;  // [OSL_MAXD][SHALLOW_SLOTS] + 1 counter
  const int D = tp.D, NC = OSL_NCOUNT(D);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = __ldcg(&fs->acc_emit[parity]);
  const int n_valid = __ldcg(&fs->acc_valid[parity]);
  const int n_invalid_front = n_in - n_valid + (int)A.src_base;
  const int cur = __ldcg(&fs->cur_size);
  // voxel grids that arrived sorted and gap-free were not sorted again: the key list as k_emit_grid wrote it
  const u64* __restrict__ keys = (mode == 2 && __ldcg(&fs->acc_unsorted[parity]) == 0) ? keys_dense : keys_sorted;
  const u32 size0 = (u32)(cur > 8 ? cur : 8);
  const int nvb = (n + AN_THREADS - 1) / AN_THREADS;
  // every CTA owns a CONTIGUOUS range of virtual blocks: the exclusive prefix of its first block is the sum of the
  // lower CTAs' totals, and the prefix of each further block follows by adding the previous block's counts
  const int per = (nvb + G - 1) / G;
  const int vb0 = min(nvb, bid * per), vb1 = min(nvb, vb0 + per);

  PROF(16);
  if (threadIdx.x == 0 && bid < 1024) g_osl_ctaprof[0][bid] = (unsigned long long)clock64();
  // splitters for k_sort_bucket of a later frame: BK_BUCKETS-quantiles of this frame's sorted keys
  if (bid == G - 1 && tid < BK_BUCKETS - 1 && n >= BK_BUCKETS)
    split_out[tid] = keys[(size_t)(((long long)(tid + 1) * n) / BK_BUCKETS)];
  for (int c = tid; c < NC; c += AN_THREADS) s_ctot[c] = 0;
  __syncthreads();
  // (Sharing walks between the blocks of a CTA -- thread 0 walks the block's first key, the others resume where they
  // leave its path -- was measured at 50 M keys: 3.5 ms against 3.2 ms for these independent walks; with 4 CTAs per
  // SM the walk latency is already hidden and the extra serial step only adds two block barriers per block.)
  u64 ck = 0; int cm = D, cs = OSL_NONE, cslot = -1; u32 cst = 0;  // this thread's key state when the CTA owns a single block
  if (BIG && vb1 - vb0 >= 2 && A.shard != 2) {
    // (phase A's counters take the first row of s_w; the openers' list the rows behind it)
    static_assert(((NC_MAX * 4 + 7) & ~7) + AN_HEAD_BYTES <= AN_WARPS * NC_MAX * 4, "the openers' list must fit s_w");
    analyze_stream(vb0 * AN_THREADS, min(n, vb1 * AN_THREADS), n, keys, pay, mode, pool, tp, m8, s8, start, s_ctot,
                   &s_w[0][0], s_raw + ((NC_MAX * 4 + 7) & ~7), A.wcache, A.shard ? A.has_prev : 0, A.prev_key);
  } else {
    for (int vb = vb0; vb < vb1; vb++)
      if (A.shard != 2)
        analyze_block(vb, n, keys, pay, mode, pool, tp, m8, s8, start, s_ctot, &s_w[0][0], A.wcache, s_path, s_shallow,
                      ck, cm, cs, cst, cslot, A.shard ? A.has_prev : 0, A.prev_key, size0);
  }
  const bool carried = (vb1 - vb0 == 1) && A.shard != 2;
  // publish this CTA's counter vector behind an epoch flag (every CTA publishes, also one without blocks).  Compact
  // form: the D per-level counters always; the (D+1)^2 bucket counters only when a key of this CTA splits a node --
  // bit 31 of the flag says so -- which in steady state (the map already holds the surface) is no CTA at all.
  u32 mine_split = 0;
  for (int c = D + tid; c < NC; c += AN_THREADS) mine_split |= s_ctot[c];
  const int has_split = __syncthreads_or(mine_split != 0u);
  // Frame-sized inputs publish the D per-level counters as SELF-TAGGED 64-bit words, (epoch << 32) | count, bit 31 of
  // word 0 = "this CTA also published bucket counters": a word is valid the moment it carries this frame's epoch, so the
  // steady state needs no fence, no flag and no second round trip -- the readers' loads of the counters ARE the wait.
  // (The flag protocol cost ~2.7 us per frame: the writers' fence, the flag's way to L2, a poll, then the loads of the
  // counters; tools/frame_timeline.py prints the exchange time of every CTA.)  Counts of this path stay below 2^31.
  const bool tagged = !BIG && A.shard == 0 && A.lvltag != nullptr;
  if (tagged) {
    if (has_split) {  // (rare: the bucket counters go first, fenced)
      for (int c = D + tid; c < NC; c += AN_THREADS) ctatot[(size_t)bid * NC + c] = s_ctot[c];
      __syncthreads();
    }
    if (tid < D) {
      if (has_split) __threadfence();
      const u64 wv = ((u64)epoch << 32) | (u64)s_ctot[tid] | ((tid == 0 && has_split) ? 0x80000000ull : 0ull);
      *(volatile u64*)&A.lvltag[(size_t)bid * OSL_MAXD + tid] = wv;
    }
  } else if (A.shard != 2) {  // (an assign-only launch finds the vectors and flags its analyse-only launch published)
    for (int c = tid; c < (has_split ? NC : D); c += AN_THREADS) ctatot[(size_t)bid * NC + c] = s_ctot[c];
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      *(volatile u32*)&flags[bid] = epoch | (has_split ? 0x80000000u : 0u);
    }
  }
  PROF(17);
  if (threadIdx.x == 0 && bid < 1024) g_osl_ctaprof[1][bid] = (unsigned long long)clock64();

  // wait for every CTA's vector (all CTAs are co-resident: cooperative launch / first roles of k_frame), then sum them:
  // totals for the plan, exclusive prefix for the own range -- one wait, no grid barrier
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_scan + AN_WARPS);  // CTAs that published bucket counters
  if (tagged) {
    if (tid == 0) s_scan[0] = 0u;
  } else if (warp == 0) {
    u32 n_list = 0;
    for (int b0 = 0; b0 < G; b0 += 32) {
      const int b = b0 + lane;
      u32 v = 0;
      if (b < G) {
        long long spin = 0;  // bounded: a vector that never arrives traps (error to the host) instead of hanging the GPU
        while (((v = ld_vol(&flags[b])) & 0x7FFFFFFFu) != epoch)
          if (++spin > (1ll << 31)) __trap();
      }
      const u32 has = __ballot_sync(FULL, (v >> 31) != 0u);
      if (v >> 31) s_list[n_list + __popc(has & lanemask_lt())] = (unsigned short)b;
      n_list += __popc(has);
    }
    if (lane == 0) s_scan[0] = n_list;
    __threadfence();
  }
  __syncthreads();
  PROF(18);
  {
    // per-level counters (always published): 16 groups of CTAs x 32 counter lanes, one batch of independent loads per
    // thread, then a fold over the groups in shared memory -- one L2 round trip instead of one per 16 CTAs
    u32* s_red = &s_w[0][0];  // [16][64]: totals, exclusive prefixes (phase A's scratch is free, phase C re-initialises it)
    const int c = tid & 31, g = tid >> 5;
    u32 tot = 0, pre = 0;
    if (c < D) {
      for (int b0 = g; b0 < G; b0 += 4 * AN_WARPS) {
        u32 v[4];
        if (tagged) {
          bool saw_split = false;
          u64 tv[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int b = b0 + k * AN_WARPS;
            tv[k] = (b < G) ? *(const volatile u64*)&A.lvltag[(size_t)b * OSL_MAXD + c] : ((u64)epoch << 32);
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int b = b0 + k * AN_WARPS;
            long long spin = 0;  // bounded: a vector that never arrives traps (error to the host) instead of hanging the GPU
            while ((u32)(tv[k] >> 32) != epoch) {
              tv[k] = *(const volatile u64*)&A.lvltag[(size_t)b * OSL_MAXD + c];
              if (++spin > (1ll << 31)) __trap();
            }
            v[k] = (u32)tv[k];
            if (c == 0 && (v[k] & 0x80000000u)) {  // this CTA published bucket counters as well
              v[k] &= 0x7FFFFFFFu;
              s_list[atomicAdd(&s_scan[0], 1u)] = (unsigned short)b;
              saw_split = true;
            }
          }
          // acquire side of the bucket counters: the thread that observed the mark fences BEFORE the block barrier
          // behind which the other threads read them (the flag protocol's order: poll, fence, barrier, loads)
          if (saw_split) __threadfence();
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int b = b0 + k * AN_WARPS;
            v[k] = (b < G) ? __ldcg(&ctatot[(size_t)b * NC + c]) : 0u;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          tot += v[k];
          if (b0 + k * AN_WARPS < bid) pre += v[k];
        }
      }
    }
    s_red[g * 64 + c] = tot;
    s_red[g * 64 + 32 + c] = pre;
    __syncthreads();
    const int n_list = (int)s_scan[0];
    if (tid < D) {
      u32 t2 = 0, p2 = 0;
#pragma unroll
      for (int q = 0; q < AN_WARPS; q++) { t2 += s_red[q * 64 + tid]; p2 += s_red[q * 64 + 32 + tid]; }
      s_tot[tid] = t2;
      s_base[tid] = p2;
    }
    // bucket counters: only the (few) CTAs that split something published them
    for (int cc = D + tid; cc < NC; cc += AN_THREADS) {
      u32 t2 = 0, p2 = 0;
      for (int l0 = 0; l0 < n_list; l0 += 4) {
        u32 v[4]; int bb[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          bb[k] = (l0 + k < n_list) ? (int)s_list[l0 + k] : -1;
          v[k] = bb[k] >= 0 ? __ldcg(&ctatot[(size_t)bb[k] * NC + cc]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          t2 += v[k];
          if (bb[k] >= 0 && bb[k] < bid) p2 += v[k];
        }
      }
      s_tot[cc] = t2;
      s_base[cc] = p2;
    }
  }
  __syncthreads();
  PROF(19);
  PROF(20);
  if (A.shard == 1) {  // sharded build, first half: this rank's totals go to the host, which sums them over the ranks
    if (bid == G - 1)
      for (int c = tid; c < NC; c += AN_THREADS) A.rank_tot[c] = s_tot[c];
    return;
  }
  if (A.shard == 2) {
    // second half: the bucket counters become global (totals over all ranks for the plan, the lower ranks' counts in
    // front of this rank's for the tile ranks); the per-level counters stay local -- the level lists are this rank's own
    for (int c = D + tid; c < NC; c += AN_THREADS) {
      s_base[c] += __ldg(&A.ext[c]);
      s_tot[c] = __ldg(&A.ext[NC + c]);
    }
    __syncthreads();
  }

  // phase B2: the allocation plan.  Entry e = i*D + (d-1) in the reference's order (pass i, then depth d):
  // bucket (s = d - i, d).  Exclusive scan over the <= D*D entries gives each bucket's first global rank.
  for (int c = tid; c < NC; c += AN_THREADS) s_plan[c] = 0;
  u32 val = 0;
  int bs = 0, bd = 0;
  if (tid < D * D) {
    const int i = tid / D;
    bd = tid % D + 1;
    bs = bd - i;
    if (bs >= 1 && !(bd == D && bs != D)) val = s_tot[OSL_CBKT(D, bs, bd)];
    else bs = 0;
  }
  u32 incl = val;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  u32 woff = 0, n_split = 0;
#pragma unroll
  for (int w = 0; w < AN_WARPS; w++) {
    const u32 t = s_scan[w];
    if (w < warp) woff += t;
    n_split += t;
  }
  if (bs >= 1) s_plan[OSL_CBKT(D, bs, bd)] = woff + incl - val;
  const long long after = (long long)size0 + 8ll * n_split;
  const bool overflow = after > (long long)capacity;
  PROF(21);
  if (threadIdx.x == 0 && bid < 1024) g_osl_ctaprof[2][bid] = (unsigned long long)clock64();
  if (!overflow) {
    if (BIG && !carried) {
      // many blocks per CTA: the next block's inputs are loaded while this one is laid out (each block otherwise starts
      // with a full DRAM round trip before its first vote)
      u64 nk = 0; int nm = D, ns = OSL_NONE; u32 nst = 0, npay = 0;
      {
        const int j = vb0 * AN_THREADS + tid;
        if (vb0 < vb1 && j < n) {
          nk = keys[j]; nm = m8[j]; ns = s8[j]; nst = start[j];
          if (mode != 2) npay = __ldcg(&pay[j]);
        }
      }
      for (int vb = vb0; vb < vb1; vb++) {
        const u64 k0 = nk; const int m0 = nm, s0 = ns; const u32 st0 = nst, pay0 = npay;
        nk = 0; nm = D; ns = OSL_NONE; nst = 0; npay = 0;
        const int j = (vb + 1) * AN_THREADS + tid;
        if (vb + 1 < vb1 && j < n) {
          nk = keys[j]; nm = m8[j]; ns = s8[j]; nst = start[j];
          if (mode != 2) npay = __ldcg(&pay[j]);
        }
        assign_block(vb, n, keys, pay, pool, tp, m8, s8, start, s_base, lv, mode, size0, n_invalid_front, s_plan, s_w,
                     false, k0, m0, s0, st0, s_path, s_shallow, -1, A.shard != 0, true, pay0);
      }
    } else {
      for (int vb = vb0; vb < vb1; vb++)
        assign_block(vb, n, keys, pay, pool, tp, m8, s8, start, s_base, lv, mode, size0, n_invalid_front, s_plan, s_w,
                     carried, ck, cm, cs, cst, s_path, s_shallow, cslot, A.shard != 0);
    }
  }
  PROF(22);
  if (threadIdx.x == 0 && bid < 1024) g_osl_ctaprof[3][bid] = (unsigned long long)clock64();

  // Book-keeping by the LAST CTA (the grid is sized with a margin, so it usually has no block of its own and this
  // stays off the critical path): the frame's result block, to the device copy k_levels reads and straight to the
  // pinned host ring (no cudaMemcpyAsync per frame; the host reads it after the event that follows k_levels).
  if (bid == G - 1) {
    if (tid < OSL_BUCKETS) fs->acc_bucket[parity][tid] = 0;
    if (bs >= 1) {
      const int v = (int)(woff + incl - val);
      fr->base[bs * (D + 1) + bd] = v; hr->base[bs * (D + 1) + bd] = v;
    }
    if (tid < D) {  // |codes[i]| of reference pass i
      u32 pc = 0;
      for (int d = 1; d <= D; d++) {
        const int s = d - tid;
        if (s >= 1 && !(d == D && s != D)) pc += s_tot[OSL_CBKT(D, s, d)];
      }
      fr->pass_count[tid] = (int)pc; hr->pass_count[tid] = (int)pc;
    }
    if (tid >= 1 && tid <= D) { const int v = (int)s_tot[OSL_CLVL(D, tid)]; fr->n_level[tid] = v; hr->n_level[tid] = v; }
    if (tid == 0) {
      const int seq = __ldcg(&fs->frame_seq) + 1;
      FrameState* out[2] = {fr, hr};
#pragma unroll
      for (int k = 0; k < 2; k++) {
        FrameState* o = out[k];
        o->n_in = n_in; o->n_valid = n_valid; o->n_emit = n; o->n_invalid_front = n_invalid_front;
        o->n_level[0] = s_tot[OSL_CLVL(D, 1)] > 0 ? 1 : 0;
        o->n_level[D + 1] = 0;
        o->n_split = (int)n_split;
        o->size_before = (int)size0;
        o->capacity = capacity;
        o->size_after = (int)min(after, (long long)0x7FFFFFFF);
        o->overflow = overflow ? 1 : 0;
        o->frame_seq = seq;
        o->cur_size = overflow ? cur : (int)after;
      }
      // every CTA that has work read these before it published / passed the barrier; a CTA that starts later sees
      // 0 entries and idles
      fs->acc_valid[parity] = 0; fs->acc_emit[parity] = 0; fs->acc_unsorted[parity] = 0; fs->acc_tiles[parity] = 0;
      if (!overflow) fs->cur_size = (int)after;
      fs->frame_seq = seq;
    }
  }
}

// (3 CTAs per SM: 40 registers, no spills; at 50 M keys the extra resident walks are worth 20 %)
__global__ void __launch_bounds__(AN_THREADS, 3) k_structure(StructArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  structure_body<false>(A, (int)blockIdx.x, (int)gridDim.x, s_raw);
}

// Voxel grids / clouds of millions of keys: 2 CTAs per SM with 64 registers -- four walks in flight per thread in phase A
// and the next block's inputs prefetched in phase C hide more latency than a third resident CTA does.
__global__ void __launch_bounds__(AN_THREADS, 2) k_structure_big(StructArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  structure_body<true>(A, (int)blockIdx.x, (int)gridDim.x, s_raw);
}

// ------------------------------------------------------------------------------------------------ k_levels
// Bottom-up VALUE update (word1 only; every word0 is final after k_structure, which is what lets the next frame's
// k_structure overlap this kernel): every touched node writes its new value into its own slot; one step per level,
// so a level costs one 64-byte tile load per touched node and one 4-byte store, all independent.
//   phase 1  leaves: blend the winning input's colour into word1 (svo.cu:366-381 / :318-332)
//   phase 2  d = D-1 .. 1: word1 = integer mean / max of the node's 8 children (svo.cu:384-441, Q5)
//   phase 3  Q6: the root average lands in node 0's value word (svo.cu:399-412,439)
// Cooperative.  Wide levels use the whole grid with a grid barrier in between.  As soon as a level fits one CTA
// (n_level is monotone in d) the other CTAs signal "done" and exit, and CTA 0 finishes alone: it fetches the tiles
// of ALL remaining levels at once (independent loads) and folds them bottom-up in shared memory, block barriers
// while a level is wider than a warp, warp barriers above that.
#define LEVEL_THREADS 512
#define LEVEL_NARROW 512
#ifndef LEVEL_STAGE
#define LEVEL_STAGE 1024
#endif
#define LEVEL_SMEM (LEVEL_STAGE * (32 + 4 + 4 + 2 + 2) + 64 + 512)

__device__ __forceinline__ void level_leaf(u32* pool, const LevelArrays& lv, int D, int idx, int mode,
                                           const uint8_t* __restrict__ rgb, const float* __restrict__ colors4,
                                           size_t rgb_bytes) {
  const size_t oD = lv.off[D];
  const u32 src = __ldg(&lv.src[idx]);
  const u32 node = __ldg(&lv.self[oD + idx]);
  u32* w = pool + 2 * (size_t)node;
  const u32 cur = __ldcg(w + 1);
  u32 nv;
  if (mode == 2) {
    const float4 col = __ldg(reinterpret_cast<const float4*>(colors4) + src);
    nv = osl_blend_f4(cur, col.x, col.y, col.z);
  } else {
    // the winner's 3 colour bytes: one or two aligned 32-bit loads (the plane may live in pinned HOST memory, where
    // every load is a PCIe read -- osl_integrate_depth_host)
    const size_t off = 3 * (size_t)src;
    u32 r8, g8, b8;
    if ((reinterpret_cast<uintptr_t>(rgb) & 3) == 0 && (off & ~(size_t)3) + 8 <= rgb_bytes) {
      const u32* wp = reinterpret_cast<const u32*>(rgb + (off & ~(size_t)3));
      const int sh = 8 * (int)(off & 3);
      unsigned long long v = __ldg(wp);
      if (sh > 8) v |= (unsigned long long)__ldg(wp + 1) << 32;
      v >>= sh;
      r8 = (u32)v & 0xFFu; g8 = (u32)(v >> 8) & 0xFFu; b8 = (u32)(v >> 16) & 0xFFu;
    } else {
      const uint8_t* q = rgb + off;
      r8 = __ldg(q); g8 = __ldg(q + 1); b8 = __ldg(q + 2);
    }
    nv = osl_blend_u8(cur, r8, g8, b8);
  }
  w[1] = nv;
}

__device__ __forceinline__ size_t rgb_bytes_of(int n_in) { return 3 * (size_t)n_in; }

__device__ __forceinline__ void level_inner(u32* pool, const LevelArrays& lv, int d, int idx) {
  const size_t od = lv.off[d];
  const u32 ct = __ldg(&lv.ctile[od + idx]);
  const u32 node = __ldg(&lv.self[od + idx]);
  const uint4* tile = reinterpret_cast<const uint4*>(pool + 2 * (size_t)(ct & OSL_MASK));
  u32 v[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint4 q = __ldcg(tile + i);
    v[2 * i] = q.y; v[2 * i + 1] = q.w;
  }
  pool[2 * (size_t)node + 1] = osl_average8(v);
}

struct LevelArgs {
  u32* pool; LevelArrays lv; const FrameState* fr; u32* done; int D; int mode;
  const uint8_t* rgb; const float* colors4;
  int no_root;                   // sharded build: the root average (Q6) is written by osl_shard_fixup, not here
  FrameState* hr; int done_tag;  // pinned result block of the frame; done_tag (frame number + 1) is stored into
                                 // hr->done_flag when every value of the frame has been written
};

// barrier over the G co-resident CTAs of this role (k_levels is launched cooperatively; as a role of k_frame its CTAs
// are the first of the grid): *ctr counts arrivals over the successive barriers of one launch, `phase` = 1, 2, ...
__device__ __forceinline__ void role_barrier(u32* ctr, int G, int phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    long long spin = 0;
    while (*(volatile u32*)ctr < (u32)(G * phase))
      if (++spin > (1ll << 31)) __trap();
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void levels_body(const LevelArgs& A, int bid, int G, unsigned char* s_raw) {
  u32* pool = A.pool; const LevelArrays& lv = A.lv; const FrameState* fr = A.fr; u32* done = A.done;
  const int D = A.D, mode = A.mode;
  const uint8_t* __restrict__ rgb = A.rgb; const float* __restrict__ colors4 = A.colors4;
  int* s_nl = reinterpret_cast<int*>(s_raw + LEVEL_STAGE * (32 + 4 + 4 + 2 + 2) + 64);  // n_level[d]
  int* s_pre = s_nl + (OSL_MAXD + 2);  // s_pre[d] = sum of n_level[1..d-1]
  int& s_overflow = s_pre[OSL_MAXD + 2];
  int& s_nin = s_pre[OSL_MAXD + 3];
  const int tid = threadIdx.x;
  const int gtid = bid * LEVEL_THREADS + tid, gsz = G * LEVEL_THREADS;
  // the frame's level counts: one parallel round trip, then a prefix over <= 20 values
  if (tid <= D && tid >= 1) s_nl[tid] = fr->n_level[tid];
  if (tid == 0) { s_overflow = fr->overflow; s_nin = fr->n_in; }
  __syncthreads();
  if (s_overflow) {  // the frame was dropped by the structure stage: nothing to fold
    if (bid == 0 && tid == 0 && A.hr) { __threadfence_system(); *(volatile int*)&A.hr->done_flag = A.done_tag; }
    return;
  }
  if (tid == 0) {
    int run = 0;
    s_nl[0] = 0; s_pre[0] = 0;
    for (int d = 1; d <= D; d++) { s_pre[d] = run; run += s_nl[d]; }
    s_pre[D + 1] = run;
  }
  __syncthreads();
  PROF(32);

  // Subtrees.  The deepest level that fits the narrow path (or is small enough to be shared out) is the CUT: every CTA
  // takes a contiguous chunk of the cut level's nodes and folds their touched subtrees -- contiguous ranges of every
  // deeper level list (lv.fc) -- bottom-up BY ITSELF, block barriers only.  (First version: one grid barrier per wide
  // level, ~2 us each and three of them for a 640x480 frame.)  Levels above the cut that are still too wide for the
  // narrow path (huge inputs only) take the grid-barrier route as before.
  int cut = D - 1;
  const bool big = !A.no_root && s_nl[D] >= (1 << 20);
  {
    // (huge inputs: a deeper cut -- more, smaller subtrees -- so that the shares below can be balanced closely)
    const int roots_max = max(LEVEL_NARROW, (big ? 256 : 16) * G);
    while (cut >= 1 && s_nl[cut] > roots_max) cut--;
  }
  if (D >= 2 && cut >= 1) {
    int* s_lo = s_pre + (OSL_MAXD + 4);  // [D+2] first / one-past-last entry of this CTA at every level below the cut
    int* s_hi = s_lo + (OSL_MAXD + 2);
    const int n_c = s_nl[cut];
    if (tid < 2) {  // two dependent chains of first-child look-ups, side by side
      // This CTA's share of the cut level.  Frames: equal node counts.  Huge inputs: equal LEAF counts -- the subtrees of
      // a surface differ by an order of magnitude in size -- so the share starts at the cut-level ancestor (parent
      // chain) of leaf bid * n_D / G.
      const int b = bid + tid;
      int a = (int)(((long long)b * n_c) / G);
      if (big) {
        if (b == 0) a = 0;
        else if (b >= G) a = n_c;
        else {
          a = (int)(((long long)b * s_nl[D]) / G);
          for (int l = D; l > cut; l--) a = (int)__ldcg(&lv.par[lv.off[l] + a]);
        }
      }
      int* dst = tid == 0 ? s_lo : s_hi;
      dst[cut] = a;
      for (int l = cut; l < D; l++) {
        a = (a < s_nl[l]) ? (int)__ldcg(&lv.fc[lv.off[l] + a]) : s_nl[l + 1];
        // (sharded build: the first key of a slice may head only levels below the cut -- those entries, at the front
        // of their lists, descend from no node of this rank's cut level; the first CTA takes them along)
        dst[l + 1] = (tid == 0 && bid == 0) ? 0 : a;
      }
    }
    __syncthreads();
    for (int idx = s_lo[D] + tid; idx < s_hi[D]; idx += LEVEL_THREADS) level_leaf(pool, lv, D, idx, mode, rgb, colors4, rgb_bytes_of(s_nin));
    for (int l = D - 1; l >= cut; l--) {
      __syncthreads();  // (this CTA wrote the children; level_inner reads them from L2)
      for (int idx = s_lo[l] + tid; idx < s_hi[l]; idx += LEVEL_THREADS) level_inner(pool, lv, l, idx);
    }
  } else {
    const int n_D = s_nl[D];
    for (int idx = gtid; idx < n_D; idx += gsz) level_leaf(pool, lv, D, idx, mode, rgb, colors4, rgb_bytes_of(s_nin));
    cut = D;
  }
  PROF(35);

  // levels above the cut that are too wide for one CTA: the whole grid, a barrier per level
  int d = cut - 1;
  for (int phase = 1; d >= 1; d--, phase++) {
    const int n_d = s_nl[d];
    if (n_d <= LEVEL_NARROW) break;
    role_barrier(done + 1, G, phase);
    for (int idx = gtid; idx < n_d; idx += gsz) level_inner(pool, lv, d, idx);
  }
  PROF(36);
  // one-sided barrier: everybody signals, only CTA 0 waits
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicAdd(done, 1u);
  }
  if (bid != 0) return;
  if (tid == 0) {
    long long spin = 0;
    while (*(volatile u32*)done != (u32)G)
      if (++spin > (1ll << 31)) __trap();  // bounded wait, see k_structure
    *(volatile u32*)done = 0u;  // the next launch starts from zero (launches of one tree are stream-ordered)
    *(volatile u32*)(done + 1) = 0u;  // (every CTA passed its last role_barrier before it signalled `done`)
    __threadfence();
  }
  __syncthreads();
  PROF(37);

  // narrow levels d..1 + the root: CTA 0 alone
  const int staged = (d >= 1) ? s_pre[d + 1] : 0;  // nodes of levels 1..d
  if (staged + 1 <= LEVEL_STAGE) {
    u32 (*s_w1)[8] = reinterpret_cast<u32 (*)[8]>(s_raw);
    u32* s_node = reinterpret_cast<u32*>(s_raw + LEVEL_STAGE * 32);
    u32* s_ct = s_node + LEVEL_STAGE;
    unsigned short* s_par = reinterpret_cast<unsigned short*>(s_ct + LEVEL_STAGE);
    unsigned short* s_dig = s_par + LEVEL_STAGE;
    // staged index: root = 0, node idx of level l = 1 + s_pre[l] + idx
    for (int e = tid; e <= staged; e += LEVEL_THREADS) {
      u32 ct = 0u, node = 0u, par = 0u, dig = 0u;
      if (e > 0) {
        int l = 1;
        while (l < d && (e - 1) >= s_pre[l + 1]) l++;
        const int idx = e - 1 - s_pre[l];
        const size_t ol = lv.off[l];
        ct = __ldg(&lv.ctile[ol + idx]);
        dig = (u32)__ldg(&lv.digit[ol + idx]);
        node = __ldg(&lv.self[ol + idx]);
        const u32 pr = __ldg(&lv.par[ol + idx]);  // 0xFFFFFFFF: the parent is in another rank's lists (sharded build)
        par = (l == 1) ? 0u : (pr == 0xFFFFFFFFu ? 0xFFFFu : (u32)(1 + s_pre[l - 1]) + pr);
      }
      const uint4* tile = reinterpret_cast<const uint4*>(pool + 2 * (size_t)(ct & OSL_MASK));
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint4 q = __ldcg(tile + i);
        s_w1[e][2 * i] = q.y; s_w1[e][2 * i + 1] = q.w;
      }
      s_node[e] = node; s_ct[e] = ct; s_par[e] = (unsigned short)par; s_dig[e] = (unsigned short)dig;
    }
    __syncthreads();
    PROF(38);
    int l = d;
    for (; l >= 1 && s_nl[l] > 32; l--) {
      const int n_l = s_nl[l], base = 1 + s_pre[l];
      for (int idx = tid; idx < n_l; idx += LEVEL_THREADS) {
        const int e = base + idx;
        const u32 avg = osl_average8(s_w1[e]);
        pool[2 * (size_t)s_node[e] + 1] = avg;
        if (s_par[e] != 0xFFFFu) s_w1[s_par[e]][s_dig[e]] = avg;
      }
      __syncthreads();
    }
    if (tid < 32) {  // levels of at most 32 nodes: one warp, warp barriers
      for (; l >= 1; l--) {
        const int n_l = s_nl[l], base = 1 + s_pre[l];
        if (tid < n_l) {
          const int e = base + tid;
          const u32 avg = osl_average8(s_w1[e]);
          pool[2 * (size_t)s_node[e] + 1] = avg;
          if (s_par[e] != 0xFFFFu) s_w1[s_par[e]][s_dig[e]] = avg;
        }
        __syncwarp();
      }
      if (tid == 0 && s_nl[1] > 0 && !A.no_root) pool[1] = osl_average8(s_w1[0]);  // phase 3 (Q6)
    }
    PROF(39);
    __syncthreads();
    if (tid == 0 && A.hr) { __threadfence_system(); *(volatile int*)&A.hr->done_flag = A.done_tag; }
    return;
  }
  // fallback: the narrow part does not fit the staging area -> one block barrier + one L2 round trip per level
  for (; d >= 1; d--) {
    const int n_d = s_nl[d];
    for (int idx = tid; idx < n_d; idx += LEVEL_THREADS) level_inner(pool, lv, d, idx);
    __syncthreads();
  }
  if (tid == 0 && s_nl[1] > 0 && !A.no_root) {  // phase 3 (Q6)
    const uint4* tile = reinterpret_cast<const uint4*>(pool);
    u32 v[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint4 q = __ldcg(tile + i);
      v[2 * i] = q.y; v[2 * i + 1] = q.w;
    }
    pool[1] = osl_average8(v);
  }
  __syncthreads();
  if (tid == 0 && A.hr) { __threadfence_system(); *(volatile int*)&A.hr->done_flag = A.done_tag; }
}

__global__ void __launch_bounds__(LEVEL_THREADS) k_levels(LevelArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  levels_body(A, (int)blockIdx.x, (int)gridDim.x, s_raw);
}

// ------------------------------------------------------------------------------------------------ k_frame
// ONE launch per frame for pipelined depth frames: the four stages run as ROLES of one grid, each on the frame that
// has reached it --
//     launch f = { emit(f),  sort(f-1),  structure(f-2),  values(f-3) }
// on ONE stream, so every dependency between frames is plain stream order (no events, no host round trips, no flags
// inside a launch): sort(f-1) needs emit(f-1); structure(f-2) needs sort(f-2) and structure(f-3); values(f-3) needs
// structure(f-3) -- all in launch f-1; the key-list slot emit(f) fills was last read by structure(f-3), the level
// lists structure(f-2) fills were last read by values(f-4).  (A first version chained emit(f) -> sort(f) inside one
// launch through a flag: the sort role, sharing its SMs with three other roles, then needs 18 us after the 13 us of
// emit, and that chain -- 31 us -- set the period; profiles/r02_timeline_a.txt.)  Launches are chained with programmatic dependent launch: the next grid's CTAs become
// resident while this one drains and block in griddepcontrol.wait, so the stream never idles through a launch latency.
// CTA ranges: [0, gS) structure, then gV values, then gE emit tiles, then BK_BUCKETS sort buckets.  The roles that
// spin on flags of their own kind (structure: counter exchange; values: level barrier) come FIRST, the roles that
// never wait (emit, sort) last: with at most 2 * num_sms CTAs of spinning roles per device (the host caps gS + gV) a
// spinning CTA can never keep the CTA it waits for off the machine.
struct SortArgs {
  const u64* kin; const u32* pin; u64* kout; u32* pout; u64* kscr; u32* pscr; const FrameState* fs;
  const u64* split; int passes; int parity; const u64* bkeys; const u32* bpay;
};
struct EmitArgs {
  EmitParams p; TreeParams tp; int vec_ok; u64* keys; u32* pay; FrameState* fs; int parity;
  const u64* split; u64* bkeys; u32* bpay;
};
struct FrameArgs {
  StructArgs S; LevelArgs V; EmitArgs E; SortArgs So;
  int gS, gV, gE, gSo;
  int trace;  // >= 0: record the time span of every role of this launch in g_osl_span[trace] (tools/frame_timeline.py)
};

// tracing aid: [launch slot][0 structure, 1 values, 2 emit, 3 sort, 4 CTA arrival (before griddepcontrol.wait)][min
// start, max end] in %globaltimer nanoseconds
#define OSL_SPAN_SLOTS 32
__device__ unsigned long long g_osl_span[OSL_SPAN_SLOTS][5][2];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void span_mark(int slot, int role, bool end) {
  if (slot < 0 || threadIdx.x != 0) return;
  const unsigned long long t = global_ns();
  if (end) atomicMax(&g_osl_span[slot][role][1], t); else atomicMin(&g_osl_span[slot][role][0], t);
}
#define FRAME_THREADS 512
constexpr int osl_max4(int a, int b, int c, int d) { return (a > b ? a : b) > (c > d ? c : d) ? (a > b ? a : b) : (c > d ? c : d); }
#define FRAME_SMEM osl_max4(LEVEL_SMEM, EMIT_SMEM, BK_SMEM, STRUCT_SMEM)
static_assert(FRAME_SMEM >= EMIT_SMEM && FRAME_SMEM >= BK_SMEM && FRAME_SMEM >= STRUCT_SMEM, "k_frame shared memory");
static_assert(AN_THREADS == FRAME_THREADS && LEVEL_THREADS == FRAME_THREADS && EMIT_THREADS == FRAME_THREADS &&
              BK_THREADS <= FRAME_THREADS, "k_frame role CTAs");

__device__ __forceinline__ void spin_until_eq(const u32* p, u32 want) {
  long long spin = 0;
  while (*(const volatile u32*)p != want)
    if (++spin > (1ll << 31)) __trap();  // bounded: an error to the host instead of a hung GPU
  __threadfence();
}

__global__ void __launch_bounds__(FRAME_THREADS, 2) k_frame(const __grid_constant__ FrameArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  // programmatic dependent launch: let the next launch's CTAs take their places now (they wait below until this grid
  // has completed), then wait for the previous launch ourselves
  asm volatile("griddepcontrol.launch_dependents;");
  span_mark(A.trace, 4, false);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  span_mark(A.trace, 4, true);
  int b = (int)blockIdx.x;
  if (b < A.gS) {
    span_mark(A.trace, 0, false);
    structure_body<false>(A.S, b, A.gS, s_raw);
    span_mark(A.trace, 0, true);
    return;
  }
  b -= A.gS;
  if (b < A.gV) {
    span_mark(A.trace, 1, false);
    levels_body(A.V, b, A.gV, s_raw);
    span_mark(A.trace, 1, true);
    return;
  }
  b -= A.gV;
  if (b < A.gE) {
    span_mark(A.trace, 2, false);
    if (A.E.p.ready) {  // host frame: the staging copies of this frame (copy engine) have landed
      if (threadIdx.x == 0) spin_until_eq(A.E.p.ready, A.E.p.ready_seq);
      __syncthreads();
    }
    emit_body(A.E.p, A.E.tp, A.E.vec_ok, A.E.keys, A.E.pay, A.E.fs, A.E.parity, b, s_raw, A.E.split,
              A.E.bkeys, A.E.bpay);
    span_mark(A.trace, 2, true);
    return;
  }
  b -= A.gE;
  if (threadIdx.x >= BK_THREADS) return;  // whole warps leave: the barriers below count the remaining ones only
  span_mark(A.trace, 3, false);
  bucket_body(A.So.kin, A.So.pin, A.So.kout, A.So.pout, A.So.kscr, A.So.pscr, A.So.fs, A.So.split, A.So.passes,
              A.So.parity, b, s_raw, A.So.bkeys, A.So.bpay);
  span_mark(A.trace, 3, true);
}

// ------------------------------------------------------------------------------------------------ host side
static size_t level_cap(size_t n, int d) {
  // n_d <= min(n, 8^d)
  if (3 * d >= 40) return n;
  const size_t p = (size_t)1 << (3 * d);
  return p < n ? p : n;
}

int osl_sort_occupancy() {
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_sort, SORT_THREADS, 0);
  if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; return 0; }
  return occ;
}

osl_status osl_ensure_workspace(osl_svo* t, size_t n) {
  if (n <= t->ws_cap) return OSL_OK;
  size_t cap = t->ws_cap ? t->ws_cap : (t->ws_want ? t->ws_want : 1);  // ws_want: capacity before osl_drop_workspace
  while (cap < n) cap *= 2;
  if (cap < 4096) cap = 4096;
  for (int f = 0; f < OSL_FRONT; f++) {
    cudaFree(t->d_keysA[f]); cudaFree(t->d_keysB[f]); cudaFree(t->d_payA[f]); cudaFree(t->d_payB[f]);
    t->d_keysA[f] = t->d_keysB[f] = nullptr; t->d_payA[f] = t->d_payB[f] = nullptr;
  }
  cudaFree(t->d_m); cudaFree(t->d_s); cudaFree(t->d_blockcnt);
  for (int b = 0; b < OSL_BACK; b++) { cudaFree(t->d_level_mem[b]); t->d_level_mem[b] = nullptr; }
  cudaFree(t->d_keysC); cudaFree(t->d_payC); cudaFree(t->d_start); cudaFree(t->d_flags);
  t->d_keysC = nullptr; t->d_payC = nullptr; t->d_start = nullptr; t->d_flags = nullptr; t->d_lvltag = nullptr;
  t->d_m = t->d_s = nullptr;
  t->d_blockcnt = nullptr;
  t->ws_cap = 0;
  const int D = t->tp.D;
  for (int f = 0; f < OSL_FRONT; f++) {
    OSL_CUDA(cudaMalloc(&t->d_keysA[f], cap * sizeof(u64)));
    OSL_CUDA(cudaMalloc(&t->d_keysB[f], cap * sizeof(u64)));
    OSL_CUDA(cudaMalloc(&t->d_payA[f], cap * sizeof(u32)));
    OSL_CUDA(cudaMalloc(&t->d_payB[f], cap * sizeof(u32)));
  }
  OSL_CUDA(cudaMalloc(&t->d_keysC, cap * sizeof(u64)));  // k_sort_bucket's slow-path scratch
  OSL_CUDA(cudaMalloc(&t->d_payC, cap * sizeof(u32)));
  OSL_CUDA(cudaMalloc(&t->d_m, cap));
  OSL_CUDA(cudaMalloc(&t->d_s, cap));
  const size_t nctas = (size_t)(t->structure_grid > 0 ? t->structure_grid : 1);
  OSL_CUDA(cudaMalloc(&t->d_blockcnt, nctas * OSL_NCOUNT(D) * sizeof(u32)));  // one counter vector per CTA
  OSL_CUDA(cudaMalloc(&t->d_start, cap * sizeof(u32)));
  // (flags, then -- 8-byte aligned -- the tagged per-level counter words of the frame-sized exchange)
  const size_t flag_bytes = (nctas * sizeof(u32) + 7) & ~(size_t)7;
  OSL_CUDA(cudaMalloc(&t->d_flags, flag_bytes + nctas * OSL_MAXD * sizeof(u64)));
  OSL_CUDA(cudaMemset(t->d_flags, 0, flag_bytes + nctas * OSL_MAXD * sizeof(u64)));  // epoch-tagged (frame number + 1), never reset
  t->d_lvltag = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(t->d_flags) + flag_bytes);
  for (int b = 0; b < OSL_BACK; b++) {
    LevelArrays& lv = t->lv[b];
    size_t total = 0;
    for (int d = 0; d <= D + 1; d++) {
      lv.off[d] = total;
      total += (d >= 1 && d <= D) ? level_cap(cap, d) : 0;
    }
    // ctile(4) + par(4) + self(4) + digit(1) bytes per level entry, src(4) per leaf
    uint8_t* mem;
    OSL_CUDA(cudaMalloc(&mem, total * 17 + cap * 4 + 64));
    t->d_level_mem[b] = mem;
    lv.ctile = reinterpret_cast<u32*>(mem);
    lv.par = lv.ctile + total;
    lv.self = lv.par + total;
    lv.fc = lv.self + total;
    lv.src = lv.fc + total;
    lv.digit = reinterpret_cast<uint8_t*>(lv.src + cap);
  }
  t->ws_cap = cap;
  return OSL_OK;
}

osl_status osl_integrate_init(osl_svo* t) {
  OSL_CUDA(cudaFuncSetAttribute((const void*)k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, EMIT_SMEM));
  OSL_CUDA(cudaFuncSetAttribute((const void*)k_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, LEVEL_SMEM));
  OSL_CUDA(cudaFuncSetAttribute((const void*)k_structure, cudaFuncAttributeMaxDynamicSharedMemorySize, STRUCT_SMEM));
  OSL_CUDA(cudaFuncSetAttribute((const void*)k_structure_big, cudaFuncAttributeMaxDynamicSharedMemorySize, STRUCT_SMEM));
  OSL_CUDA(cudaFuncSetAttribute((const void*)k_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, FRAME_SMEM));
  OSL_CUDA(cudaFuncSetAttribute((const void*)k_sort_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
  OSL_CUDA(cudaMalloc(&t->d_split, OSL_FRONT * BK_BUCKETS * sizeof(u64)));
  OSL_CUDA(cudaMalloc(&t->d_blockcnt_tot, 3 * NC_MAX * sizeof(u32)));
  for (int f = 0; f < OSL_FRONT; f++) {
    OSL_CUDA(cudaMalloc(&t->d_bkeys[f], (size_t)OSL_BUCKETS * OSL_BUCKET_CAP * sizeof(u64)));
    OSL_CUDA(cudaMalloc(&t->d_bpay[f], (size_t)OSL_BUCKETS * OSL_BUCKET_CAP * sizeof(u32)));
  }
  if (!getenv("OSL_NO_WALK_CACHE")) {
    // (the walk cache, and behind it the hint table of walk_frontier)
    OSL_CUDA(cudaMalloc(&t->d_wcache, WC_SLOTS * sizeof(u64) + WH_SLOTS * sizeof(u32)));
    OSL_CUDA(cudaMemset(t->d_wcache, 0xFF, WC_SLOTS * sizeof(u64) + WH_SLOTS * sizeof(u32)));
  }
  return osl_reset_splitters(t);
}

// default splitters: an even partition of the key space (valid, merely unbalanced) until a frame has written some
osl_status osl_reset_splitters(osl_svo* t) {
  u64 h_split[OSL_FRONT * BK_BUCKETS];
  for (int i = 0; i < OSL_FRONT * BK_BUCKETS; i++)
    h_split[i] = (u64)((i % BK_BUCKETS) + 1) * ((1ull << (3 * t->tp.D)) / BK_BUCKETS);
  OSL_CUDA(cudaMemcpy(t->d_split, h_split, sizeof(h_split), cudaMemcpyHostToDevice));
  return OSL_OK;
}

// The workspace layout depends on max_depth (counter vectors, level arrays): drop it so that the next frame
// re-allocates it.  The pipeline must be idle.
void osl_drop_workspace(osl_svo* t) {
  t->ws_want = t->ws_cap > t->ws_want ? t->ws_cap : t->ws_want;
  t->ws_cap = 0;
}

// Invariant: every word of the pool beyond the live nodes is ZERO (a tile allocated by a frame then already has
// word0 = 0, so k_structure only writes its value words and the child pointers written by other threads cannot race
// with an initialisation).
osl_status osl_grow_pool(osl_svo* t, size_t want_nodes, cudaStream_t st) {
  if (want_nodes > ((size_t)1 << 30)) return OSL_ERR_POOL_OVERFLOW;
  if (want_nodes <= t->cap_nodes) return OSL_OK;
  size_t cap = t->cap_nodes ? t->cap_nodes : 8;
  while (cap < want_nodes) cap *= 2;
  if (cap > ((size_t)1 << 30)) cap = (size_t)1 << 30;
  u32* np;
  OSL_CUDA(cudaMalloc(&np, cap * 8));
  const size_t live = t->d_pool ? (size_t)(t->size > 8 ? t->size : 8) : 0;
  OSL_CUDA(cudaMemsetAsync(np + 2 * live, 0, (cap - live) * 8, st));
  if (t->d_pool) {
    OSL_CUDA(cudaMemcpyAsync(np, t->d_pool, live * 8, cudaMemcpyDeviceToDevice, st));
    OSL_CUDA(cudaStreamSynchronize(st));
    cudaFree(t->d_pool);
  }
  t->d_pool = np;
  t->cap_nodes = cap;
  return OSL_OK;
}

int osl_structure_occupancy() {
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_structure, AN_THREADS, STRUCT_SMEM);
  if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; return 0; }
  return occ;
}
int osl_structure_big_occupancy() {
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_structure_big, AN_THREADS, STRUCT_SMEM);
  return e == cudaSuccess ? occ : 0;
}

int osl_levels_occupancy() {
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_levels, LEVEL_THREADS, LEVEL_SMEM);
  if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; return 0; }
  return occ;
}

static inline int mode_of(const osl_svo* t, int slot) { return t->ring_mode[slot]; }

static StructArgs make_struct_args(osl_svo* t, const u64* skeys, u32* spay, FrameState* fs, FrameState* fr,
                                   unsigned long long f, const LevelArrays& lv, int mode, int n, int fslot) {
  StructArgs A;
  memset(&A, 0, sizeof(A));
  A.keys_sorted = skeys; A.keys_dense = t->d_keysA[fslot]; A.pay = spay; A.pool = t->d_pool; A.tp = t->tp;  // (dense: voxel grids in Morton order, the key list as emitted)
  A.fs = fs; A.fr = fr; A.hr = &t->h_ring[f % OSL_RING];  // pinned host memory, device-accessible (UVA)
  A.m8 = t->d_m; A.s8 = t->d_s; A.start = t->d_start; A.ctatot = t->d_blockcnt; A.flags = t->d_flags;
  A.epoch = (u32)(f + 1); A.lv = lv; A.mode = mode; A.capacity = (int)t->cap_nodes; A.n_in = n; A.parity = fslot;
  A.split_out = t->d_split + fslot * BK_BUCKETS;
  A.wcache = t->d_wcache;
  A.lvltag = t->d_lvltag;
  return A;
}

static LevelArgs make_level_args(osl_svo* t, const LevelArrays& lv, const FrameState* fr, unsigned long long f,
                                 int mode, const uint8_t* rgb, const float* colors4) {
  LevelArgs A;
  memset(&A, 0, sizeof(A));
  A.pool = t->d_pool; A.lv = lv; A.fr = fr;
  A.done = t->d_scan_totals + OSL_NCOUNT(OSL_MAXD);  // [0] one-sided barrier, [1] level barrier (zero at rest)
  A.D = t->tp.D; A.mode = mode; A.rgb = rgb; A.colors4 = colors4;
  A.hr = &t->h_ring[f % OSL_RING]; A.done_tag = (int)(f + 1);
  return A;
}

// Consume the result blocks of frames that have completed (non-blocking unless `block`): exact node count, counters.
osl_status osl_poll_results(osl_svo* t, bool block) {
  while (t->ring_tail != t->ring_head) {
    const int slot = (int)(t->ring_tail % OSL_RING);
    if (t->ring_kind[slot] == 1) {  // k_frame path: completion is the tag the value stage stores in the pinned block
      const int tag = (int)(t->ring_tail + 1);
      if (block && *(volatile int*)&t->h_ring[slot].done_flag != tag) {
        osl_status rc = osl_fused_flush(t);
        if (rc) return rc;
        cudaError_t e = cudaStreamSynchronize(t->pipe[2]);
        if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; return OSL_ERR_CUDA; }
      }
      if (*(volatile int*)&t->h_ring[slot].done_flag != tag) {
        if (block) return OSL_ERR_CUDA;  // (cannot happen: the stream has drained)
        break;
      }
      __sync_synchronize();  // the block's other words were written before the tag
    } else {
      cudaError_t e = block ? cudaEventSynchronize(t->ring_ev[slot]) : cudaEventQuery(t->ring_ev[slot]);
      if (e == cudaErrorNotReady) break;
      if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; return OSL_ERR_CUDA; }
    }
    const FrameState& F = t->h_ring[slot];
    t->inflight_headroom -= t->ring_headroom[slot];
    t->ring_tail++;
    if (F.overflow) {
      t->sticky_error = OSL_ERR_POOL_OVERFLOW;  // the frame was dropped on the device (nothing written)
    } else {
      t->size = F.size_after;
    }
    const int D = t->tp.D;
    osl_counters& c = t->counters;
    if (mode_of(t, slot) != 2 && F.n_in > 0) {
      int widest = 0;
      for (int d = 0; d <= D; d++) widest = F.n_level[d] > widest ? F.n_level[d] : widest;
      t->hint_emit = F.n_emit; t->hint_level = widest; t->hint_n_in = F.n_in;
    }
    c.n_points = F.n_in;
    c.n_valid = F.n_valid;
    c.n_unique = F.n_level[D];
    c.n_split = F.n_split;
    int64_t psum = 0;
    for (int i = 0; i <= OSL_MAX_DEPTH; i++) {
      c.pass_sizes[i] = (i < D) ? F.pass_count[i] : 0;
      c.parents[i] = (i < D) ? F.n_level[i] : 0;
      psum += c.parents[i];
    }
    c.n_nodes = t->size;
    const int mode = t->ring_mode[slot];
    const int64_t in_bytes = mode == 0 ? 5ll * F.n_in : (mode == 1 ? 15ll * F.n_in : 32ll * F.n_in);
    c.algorithmic_bytes = in_bytes + 8 * c.n_unique + 68 * c.n_split + 68 * psum;
    c.total_algorithmic_bytes += c.algorithmic_bytes;
    c.frames++;
  }
  return OSL_OK;
}

// Wait for the OLDEST frame in flight only (throttling), fold its result block.
static osl_status wait_oldest(osl_svo* t) {
  if (t->ring_tail == t->ring_head) return OSL_OK;
  const int slot = (int)(t->ring_tail % OSL_RING);
  if (t->ring_kind[slot] == 1) {
    // the oldest frame's last stage rides in the launch three frames later: issue it if it has not been issued
    if (t->ring_head - t->ring_tail <= 3) {
      osl_status rc = osl_fused_flush(t);
      if (rc) return rc;
    }
    const int tag = (int)(t->ring_tail + 1);
    for (long long spin = 0; *(volatile int*)&t->h_ring[slot].done_flag != tag; spin++) {
      if ((spin & 0x3FF) == 0x3FF) {  // a kernel that trapped (or a lost launch) must not hang the host
        cudaError_t e = cudaStreamQuery(t->pipe[2]);
        if (e == cudaSuccess) break;  // drained: poll below sees the tag (or reports the inconsistency)
        if (e != cudaErrorNotReady) { g_osl_last_cuda_error = (int)e; return OSL_ERR_CUDA; }
      }
    }
    if (*(volatile int*)&t->h_ring[slot].done_flag != tag) return osl_poll_results(t, true);
    return osl_poll_results(t, false);
  }
  cudaError_t e = cudaEventSynchronize(t->ring_ev[slot]);
  if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; return OSL_ERR_CUDA; }
  return osl_poll_results(t, false);
}

static int grid_for(long long items, int per_cta, int cap) {
  long long g = (items + per_cta - 1) / per_cta;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static osl_status drain_streams(osl_svo* t, cudaStream_t st) {
  OSL_CUDA(cudaStreamSynchronize(st));
  OSL_CUDA(cudaStreamSynchronize(t->copy_stream));
  for (int i = 0; i < 4; i++) OSL_CUDA(cudaStreamSynchronize(t->pipe[i]));
  return OSL_OK;
}

int g_osl_piped_trees = 0;  // trees of this process that have used pipelined mode and are still alive

// One frame: k_emit -> k_sort(_bucket) -> k_structure -> k_levels (the result block goes straight into the pinned ring).
//
// Strict mode (default for device inputs): everything is enqueued on the caller's stream, in order.
// Pipelined mode (osl_svo_set_pipeline, and always for host frames whose copies the library owns): the four stages
// run on four internal streams, so that consecutive frames overlap --
//     E: k_emit(f+3)   So: k_sort(f+2)   S: k_structure(f+1)   V: k_levels(f)
// k_structure(f+1) only needs the STRUCTURE of the tree after frame f (word0, final after k_structure(f)); the value
// fold of frame f (word1) runs concurrently.  Buffers: 3 key-list slots (E/So/S), 2 level-list + result slots (S/V).
// Other streams are ordered after the pipeline lazily (osl_join: the library's own raycast / extraction / download
// entry points call it; foreign work calls osl_svo_join).  Cooperative grids are capped at num_sms/3 CTAs in this mode (num_sms/(3*T) when T trees of the
// process pipeline): at most three cooperative kernels per tree (grid sort, k_structure, k_levels) run concurrently,
// <= num_sms CTAs in total, so a waiting CTA always finds an empty SM and no grid barrier can deadlock.
static osl_status osl_order_after_readers(osl_svo* t, cudaStream_t st, bool piped, cudaStream_t sS, cudaStream_t sV);

// One k_frame launch: the structure stage of frame fz_s, the value stage of frame fz_v and -- when `nw` is given --
// emit + sort of the new frame.  nw == NULL drains the pipeline by one stage (osl_fused_flush).
static osl_status fused_launch(osl_svo* t, const osl_svo::FzStage* nw, const EmitParams* ep) {
  FrameArgs A;
  memset(&A, 0, sizeof(A));
  FrameState* fs = t->d_fs;
  if (t->fz_s.valid) {
    const osl_svo::FzStage& q = t->fz_s;
    A.S = make_struct_args(t, t->d_keysB[q.fslot], t->d_payB[q.fslot], fs, t->d_fs + 1 + q.bslot, q.f, t->lv[q.bslot], 0,
                           q.n, q.fslot);
    A.gS = q.gS;
  }
  if (t->fz_v.valid) {
    const osl_svo::FzStage& q = t->fz_v;
    A.V = make_level_args(t, t->lv[q.bslot], t->d_fs + 1 + q.bslot, q.f, 0, q.rgb, nullptr);
    A.gV = q.gV;
  }
  if (t->fz_so.valid) {
    const osl_svo::FzStage& q = t->fz_so;
    A.So.kin = t->d_keysA[q.fslot]; A.So.pin = t->d_payA[q.fslot];
    A.So.kout = t->d_keysB[q.fslot]; A.So.pout = t->d_payB[q.fslot];
    A.So.kscr = t->d_keysC; A.So.pscr = t->d_payC; A.So.fs = fs;
    A.So.split = t->d_split + q.fslot * BK_BUCKETS; A.So.passes = (3 * t->tp.D + 7) / 8; A.So.parity = q.fslot;
    A.So.bkeys = t->d_bkeys[q.fslot]; A.So.bpay = t->d_bpay[q.fslot];
    A.gSo = BK_BUCKETS;
  }
  if (nw) {
    A.E.p = *ep;
    A.E.p.tiles_x = (ep->w + EMIT_TW - 1) / EMIT_TW;
    A.E.p.tiles_y = (ep->h + EMIT_TH - 1) / EMIT_TH;
    A.E.tp = t->tp;
    A.E.vec_ok = ((reinterpret_cast<uintptr_t>(ep->depth) & 7) == 0) && (ep->w % 4 == 0);
    A.E.keys = t->d_keysA[nw->fslot]; A.E.pay = t->d_payA[nw->fslot];
    A.E.fs = fs; A.E.parity = nw->fslot;
    A.E.split = t->d_split + nw->fslot * BK_BUCKETS; A.E.bkeys = t->d_bkeys[nw->fslot]; A.E.bpay = t->d_bpay[nw->fslot];
    A.gE = A.E.p.tiles_x * A.E.p.tiles_y;
  }
  const int grid = A.gS + A.gV + A.gE + A.gSo;
  A.trace = t->trace_on ? (int)(t->trace_seq++ % OSL_SPAN_SLOTS) : -1;
  if (grid > 0) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(FRAME_THREADS); cfg.dynamicSmemBytes = FRAME_SMEM;
    cfg.stream = t->pipe[2];
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    OSL_CUDA(cudaLaunchKernelEx(&cfg, k_frame, A));
    OSL_LAUNCHED(1);
  }
  t->fz_v = t->fz_s;
  t->fz_s = t->fz_so;
  if (nw) t->fz_so = *nw; else t->fz_so.valid = 0;
  return OSL_OK;
}

// Issue the launches that carry the remaining stages of the frames enqueued so far (at most three).
osl_status osl_fused_flush(osl_svo* t) {
  while (t->fz_so.valid || t->fz_s.valid || t->fz_v.valid) {
    osl_status rc = fused_launch(t, nullptr, nullptr);
    if (rc) return rc;
  }
  return OSL_OK;
}

// parameters of the certified closed-form key (k_emit_grid); off when FP32 leaves no margin inside a cell
static GridFast make_grid_fast(const TreeParams& tp) {
  GridFast g;
  memset(&g, 0, sizeof(g));
  static const int no_fast = getenv("OSL_NO_FAST_KEYS") ? 1 : 0;
  const double half = (double)tp.half;
  if (no_fast || !(half > 0.0) || tp.D < 1 || tp.D > 20) return g;
  g.G = 1 << tp.D;
  g.sh = 31 - tp.D;
  const double inv_cs = (double)g.G / (2.0 * half);
  g.lox = (float)((double)tp.cx - half); g.loy = (float)((double)tp.cy - half); g.loz = (float)((double)tp.cz - half);
  g.scale = (float)(inv_cs * ldexp(1.0, g.sh));
  const double M = fmax(fmax(fabs((double)tp.cx), fabs((double)tp.cy)), fabs((double)tp.cz)) + half;
  const double ulpM = ldexp(1.0, ilogb(M) + 1 - 23);  // ulp of the binade above M
  // band, in cells: the descent's D <= 20 roundings of half an ulp (16 ulp taken), the rounding of lo (half an ulp), the
  // three roundings of (p - lo) * scale (3 * 2^-24 relative to a quotient of at most G: G * 2^-22 taken)
  const double tol = 16.5 * ulpM * inv_cs + (double)g.G * ldexp(1.0, -22);
  if (!(tol < 0.25)) return g;
  g.tol = (u32)ceil(tol * ldexp(1.0, g.sh)) + 1u;
  g.on = 1;
  return g;
}

osl_status osl_run_integrate(osl_svo* t, EmitParams& ep, const void* colors, cudaStream_t st, const HostFrame* host) {
  const int n = ep.n;
  const int D = t->tp.D;
  const bool inputs_on_front = host != nullptr;
  if (n < 0) return OSL_ERR_INVALID;
  if (t->sticky_error) return t->sticky_error;
  osl_status rc = OSL_OK;
  if ((size_t)(n > 0 ? n : 1) > t->ws_cap) {
    rc = osl_poll_results(t, true);  // the workspace is in use by frames in flight
    if (rc) return rc;
    rc = drain_streams(t, st);
    if (rc) return rc;
    rc = osl_ensure_workspace(t, (size_t)(n > 0 ? n : 1));
    if (rc) return rc;
  }
  rc = osl_poll_results(t, false);
  if (rc) return rc;
  if (t->ring_head - t->ring_tail >= OSL_RING) {  // result ring full: wait for the oldest frame
    rc = wait_oldest(t);
    if (rc) return rc;
  }
  // Pool head-room.  A frame of n inputs splits at most n*D nodes (8*n*D new nodes); a frame in flight is accounted
  // with that bound until its exact result has been read back.  The pool is grown to hold OSL_PIPE_DEPTH such
  // frames beyond the known size, so in steady state the host never waits here; when it runs further ahead it waits
  // for the oldest frame only.  (The device re-checks exactly: a frame that would overflow writes nothing.)
  const size_t headroom = 8ull * (size_t)(n > 0 ? n : 0) * (size_t)D;
  const size_t limit = (size_t)1 << 30;
  for (;;) {
    const size_t base = (size_t)(t->size > 8 ? t->size : 8);
    if (base + t->inflight_headroom + headroom <= t->cap_nodes) break;
    size_t want = base + (size_t)(OSL_PIPE_DEPTH + 1) * headroom;
    if (want > limit) want = limit;
    if (want > t->cap_nodes) {  // grow (rare, geometric): needs the exact size, i.e. an idle pipeline
      rc = osl_poll_results(t, true);
      if (rc) return rc;
      rc = drain_streams(t, st);
      if (rc) return rc;
      rc = osl_grow_pool(t, want, st);
      if (rc) return rc;
      continue;
    }
    if (t->ring_tail == t->ring_head) break;  // at the 2^30-node cap: the device-side check protects the pool
    rc = wait_oldest(t);
    if (rc) return rc;
  }

  const unsigned long long f = t->seq;
  const int fslot = (int)(f % OSL_FRONT), bslot = (int)(f % OSL_BACK);
  const bool piped = t->pipeline || inputs_on_front;
  if (piped && !t->counted_piped) {  // trees that pipeline share the device: the CTA budget is split between them
    t->counted_piped = 1;
    g_osl_piped_trees++;
  }
  const int trees = g_osl_piped_trees > 0 ? g_osl_piped_trees : 1;
  const int passes = (3 * D + 7) / 8;
  bool use_bucket = false;
  // expected number of sorted entries / widest level, from the last completed frame (grid sizing only: every
  // kernel is grid-stride, a wrong guess costs time, not correctness)
  long long exp_emit = n, exp_level = n;
  if (t->hint_emit >= 0 && t->hint_n_in > 0 && ep.mode != 2) {
    const double scale = 1.5 * (double)n / (double)t->hint_n_in;
    exp_emit = (long long)(t->hint_emit * scale) + SORT_TILE;
    exp_level = (long long)(t->hint_level * scale) + LEVEL_THREADS;
    if (exp_emit > n) exp_emit = n;
    if (exp_level > n) exp_level = n;
    // small key list and splitters of frame f - OSL_FRONT in place -> barrier-free bucket sort
    use_bucket = f >= OSL_FRONT && exp_emit <= (long long)BK_BUCKETS * BK_CAP / 2 && !t->force_grid_sort;
  }
  // k_frame path (one launch per frame): pipelined depth frames with a host pose once the bucket sort applies.  Its
  // spinning roles may hold at most 2 CTAs per SM in total over all pipelining trees of the process.
  const int spin_cap = (2 * t->num_sms / trees) / 2;
  const bool fused = piped && t->fused_enabled && ep.mode == 0 && !ep.M_dev && n > 0 && use_bucket && spin_cap >= 8;
  if (f > 0 && fused != (t->last_fused != 0)) {
    // Path switch (the first frames of a stream run as four kernels until splitters and size hints exist; a scene
    // of another mode, a tracked frame): the two paths order their stages and recycle their slots differently, so the
    // pipeline is drained in between.  Rare, and it makes every cross-path hazard vanish.
    rc = osl_poll_results(t, true);
    if (rc) return rc;
    rc = drain_streams(t, st);
    if (rc) return rc;
  }
  cudaStream_t sE = piped ? t->pipe[0] : st, sSo = piped ? t->pipe[1] : st, sS = piped ? t->pipe[2] : st,
               sV = piped ? t->pipe[3] : st;
  if (fused) sE = sSo = sV = sS;
  if (f > 0 && !fused && (piped != (t->last_piped != 0) || (!piped && st != t->last_stream))) {
    // mode (or stream) switch: the previous frame must be complete before anything of this one starts
    cudaEvent_t prev = t->ring_ev[(f - 1) % OSL_RING];
    OSL_CUDA(cudaStreamWaitEvent(sE, prev, 0));
    if (piped) {
      OSL_CUDA(cudaStreamWaitEvent(sSo, prev, 0));
      OSL_CUDA(cudaStreamWaitEvent(sS, prev, 0));
      OSL_CUDA(cudaStreamWaitEvent(sV, prev, 0));
    }
  }
  const int sharers = 3 * trees;
  const int coop_cap = piped ? (t->num_sms / sharers > 0 ? t->num_sms / sharers : 1) : 0x7FFFFFFF;
  u64* skeys = (passes & 1) ? t->d_keysB[fslot] : t->d_keysA[fslot];
  u32* spay = (passes & 1) ? t->d_payB[fslot] : t->d_payA[fslot];
  if (use_bucket) { skeys = t->d_keysB[fslot]; spay = t->d_payB[fslot]; }
  static const int no_big_sort = getenv("OSL_NO_BIG_SORT") ? 1 : 0;
  const bool big_sort = !use_bucket && exp_emit >= (1ll << 19) && !no_big_sort;
  if (big_sort) {
    for (int q = 0; q < OSL_FRONT; q++) {  // (all slots at once: an allocation in a later frame would stall the stream)
      rc = osl_sort_big_reserve(&t->sort_ws[q]);
      if (rc) return rc;
    }
    const int pb = osl_sort_big_passes(3 * D, nullptr);
    skeys = (pb & 1) ? t->d_keysB[fslot] : t->d_keysA[fslot];
    spay = (pb & 1) ? t->d_payB[fslot] : t->d_payA[fslot];
  }
  FrameState* fs = t->d_fs;              // persistent part
  FrameState* fr = t->d_fs + 1 + bslot;  // this frame's result block
  const LevelArrays& lv = t->lv[bslot];

  // Host frames: the planes go through one of OSL_STAGES device slots on the copy stream, so that the transfer of
  // frame f+1 overlaps the kernels of frame f.  Four-kernel path: events in both directions (slot free <- k_levels,
  // k_emit <- copies).  k_frame path: the host recycles a slot once the frame that used it has reported completion in
  // its pinned result block (no event), and the frame stream waits for the copies through one event.
  const int sslot = (int)(t->stage_seq % OSL_STAGES);
  if (host) {
    const size_t np = (size_t)n;
    if (fused) {
      while (t->stage_frame[sslot] && t->ring_tail < t->stage_frame[sslot]) {
        rc = wait_oldest(t);
        if (rc) return rc;
      }
    } else if (t->stage_seq >= OSL_STAGES) {
      OSL_CUDA(cudaStreamWaitEvent(t->copy_stream, t->stage_free[sslot], 0));
    }
    OSL_CUDA(cudaMemcpyAsync(t->d_depth_stage[sslot], host->h_depth, np * 2, cudaMemcpyHostToDevice, t->copy_stream));
    // Colours: the device reads ONE pixel per observed leaf (the lowest pixel index that maps to it, ~5 % of the frame)
    // and it knows which only after the sort.  Opt-in (osl_svo_set_quirks bit 2): a PINNED colour plane is not copied,
    // k_levels gathers the winners' 3 bytes straight from host memory (zero-copy loads under UVA).  That takes 60 % of
    // the frame's payload off the link, but measured on B200 / PCIe 5 it does not pay (profiles/r02_e2e_zero_copy.md):
    // ~15 k scattered 32-byte PCIe reads per frame are bound by outstanding-request latency and stretch k_levels, so
    // the default stages the colour plane with one DMA like the depth plane.
    const uint8_t* rgb_dev = nullptr;
    if (t->zero_copy_rgb) {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, host->h_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
          attr.devicePointer)
        rgb_dev = static_cast<const uint8_t*>(attr.devicePointer);
      else
        cudaGetLastError();  // (pageable memory is reported as an error by older drivers)
    }
    if (!rgb_dev) {
      OSL_CUDA(cudaMemcpyAsync(t->d_rgb_stage[sslot], host->h_rgb, np * 3, cudaMemcpyHostToDevice, t->copy_stream));
      rgb_dev = t->d_rgb_stage[sslot];
    }
    ep.depth = t->d_depth_stage[sslot]; ep.rgb = rgb_dev;
    if (fused) {
      // copies -> emit role: one event edge per frame on the frame stream.  (The alternative, a 4-byte copy queued
      // behind the planes that the emit CTAs spin on -- which would keep the frame stream free of anything but
      // launches -- was measured slower, 47.4 vs 41.6 us per frame: 150 resident CTAs polling a flag for most of a
      // 36 us transfer get in the way of the other three roles.  OSL_FZ_HOST_FLAG=1 selects it.)
      static const int host_flag = getenv("OSL_FZ_HOST_FLAG") ? 1 : 0;
      if (!host_flag) {
        OSL_CUDA(cudaEventRecord(t->stage_copied[sslot], t->copy_stream));
        OSL_CUDA(cudaStreamWaitEvent(sS, t->stage_copied[sslot], 0));
      } else {
        const u32 val = (u32)(t->stage_seq + 1);
        t->h_ready_vals[t->stage_seq % 64] = val;
        OSL_CUDA(cudaMemcpyAsync(t->d_ready + sslot, &t->h_ready_vals[t->stage_seq % 64], sizeof(u32),
                                 cudaMemcpyHostToDevice, t->copy_stream));
        ep.ready = t->d_ready + sslot; ep.ready_seq = val;
      }
      t->stage_frame[sslot] = f + 1;
    } else {
      OSL_CUDA(cudaEventRecord(t->stage_copied[sslot], t->copy_stream));  // k_emit (stream E) waits for it
    }
  }

  if (fused) {
    rc = osl_order_after_readers(t, st, true, sS, sS);
    if (rc) return rc;
    const int cap = spin_cap < t->structure_grid ? spin_cap : t->structure_grid;
    osl_svo::FzStage nw;
    nw.valid = 1; nw.f = f; nw.n = n; nw.fslot = fslot; nw.bslot = bslot; nw.rgb = ep.rgb;
    // (the grids only cost time when they are too small -- every role loops -- so they follow the last frame closely:
    // the four roles of a 640x480 frame then fit the machine in ONE wave of 2 CTAs per SM)
    const long long fit_emit = (long long)(1.15 * t->hint_emit * (double)n / (double)t->hint_n_in) + 64;
    const long long fit_level = (long long)(1.15 * t->hint_level * (double)n / (double)t->hint_n_in) + 64;
    nw.gS = grid_for(fit_emit < exp_emit ? fit_emit : exp_emit, AN_THREADS, cap) + 1;  // (+1: the book-keeping CTA usually has no block of its own)
    if (nw.gS > cap) nw.gS = cap;
    nw.gV = grid_for(fit_level < exp_level ? fit_level : exp_level, LEVEL_THREADS, spin_cap < t->levels_grid ? spin_cap : t->levels_grid);
    rc = fused_launch(t, &nw, &ep);
    if (rc) return rc;
    const int slot = (int)(f % OSL_RING);
    t->ring_kind[slot] = 1;
    t->fz_event_valid = 0;
    t->join_pending = 1;
    t->ring_headroom[slot] = headroom;
    t->ring_mode[slot] = ep.mode;
    t->inflight_headroom += headroom;
    t->ring_head++;
    t->seq++;
    if (host) t->stage_seq++;
    t->last_stream = st;
    t->last_piped = 1;
    t->last_fused = 1;
    return OSL_OK;
  }

  const bool timing = t->stage_timing && !piped && n > 0;
  if (timing) OSL_CUDA(cudaEventRecord(t->stage_ev[0], st));
  if (n > 0) {
    // ---- E: back-projection, keys, tile-local de-duplication
    if (piped && f >= OSL_FRONT)  // the key-list slot was last read by k_structure of frame f - 3
      OSL_CUDA(cudaStreamWaitEvent(sE, t->struct_ev[(f - OSL_FRONT) % OSL_RING], 0));
    if (inputs_on_front) OSL_CUDA(cudaStreamWaitEvent(sE, t->stage_copied[sslot], 0));
    if (piped && ep.mode == 0 && ep.M_dev) {  // the pose is produced by work queued on the caller's stream
      if (!t->pose_ev) OSL_CUDA(cudaEventCreateWithFlags(&t->pose_ev, cudaEventDisableTiming));
      OSL_CUDA(cudaEventRecord(t->pose_ev, st));
      OSL_CUDA(cudaStreamWaitEvent(sE, t->pose_ev, 0));
    }
    int vec_ok = 0;
    int etiles;
    if (ep.mode == 0) {
      vec_ok = ((reinterpret_cast<uintptr_t>(ep.depth) & 7) == 0) && (ep.w % 4 == 0);
      ep.tiles_x = (ep.w + EMIT_TW - 1) / EMIT_TW;
      ep.tiles_y = (ep.h + EMIT_TH - 1) / EMIT_TH;
      etiles = ep.tiles_x * ep.tiles_y;
    } else {
      etiles = (n + EMIT_TILE - 1) / EMIT_TILE;
    }
    const bool file_ranges = use_bucket && ep.mode != 2;
    if (ep.mode == 2) {
      const int per = GRID_THREADS * GRID_PPT;
      const GridFast gf = make_grid_fast(t->tp);
      const int cg2 = 4 * t->num_sms;
      k_emit_grid<<<(n + per - 1) / per, GRID_THREADS, 0, sE>>>(ep.pts, ep.stride, n, t->tp, gf, t->d_keysA[fslot],
                                                                t->d_payA[fslot], fs, fslot);
      if (gf.on) {  // (the undecided voxels; the list lives in the payload buffer, which voxel grids do not use)
        k_grid_fix<<<cg2, GRID_THREADS, 0, sE>>>(ep.pts, ep.stride, t->tp, t->d_keysA[fslot], t->d_payA[fslot], fs, fslot);
        k_grid_fix_check<<<cg2, GRID_THREADS, 0, sE>>>(t->d_keysA[fslot], n, t->d_payA[fslot], fs, fslot);
        OSL_LAUNCHED(2);
      }
      k_grid_compact<<<cg2, GRID_THREADS, 0, sE>>>(t->d_keysA[fslot], t->d_keysB[fslot], n, fs, fslot);
      k_grid_copy_back<<<cg2, GRID_THREADS, 0, sE>>>(t->d_keysA[fslot], t->d_keysB[fslot], n, fs, fslot);
      OSL_LAUNCHED(2);
    } else {
      k_emit<<<etiles, EMIT_THREADS, EMIT_SMEM, sE>>>(ep, t->tp, vec_ok, t->d_keysA[fslot], t->d_payA[fslot], fs, fslot,
                                                     file_ranges ? t->d_split + fslot * BK_BUCKETS : nullptr,
                                                     t->d_bkeys[fslot], t->d_bpay[fslot]);
    }
    OSL_LAUNCHED(1);
    if (piped && ep.mode == 0 && ep.M_dev) {
      // the producer of the pose (a tracker on the caller's stream) rewrites it for the next frame: order the
      // caller's stream after the only kernel that reads it
      if (!t->pose_read_ev) OSL_CUDA(cudaEventCreateWithFlags(&t->pose_read_ev, cudaEventDisableTiming));
      OSL_CUDA(cudaEventRecord(t->pose_read_ev, sE));
      OSL_CUDA(cudaStreamWaitEvent(st, t->pose_read_ev, 0));
    }
    if (timing) OSL_CUDA(cudaEventRecord(t->stage_ev[1], st));
    // ---- So: sort
    if (piped) {
      OSL_CUDA(cudaEventRecord(t->emit_done[fslot], sE));
      OSL_CUDA(cudaStreamWaitEvent(sSo, t->emit_done[fslot], 0));
    }
    if (use_bucket) {
      k_sort_bucket<<<BK_BUCKETS, BK_THREADS, BK_SMEM, sSo>>>(t->d_keysA[fslot], t->d_payA[fslot], t->d_keysB[fslot],
                                                              t->d_payB[fslot], t->d_keysC, t->d_payC, fs,
                                                              t->d_split + fslot * BK_BUCKETS, passes, fslot,
                                                              ep.mode != 2 ? t->d_bkeys[fslot] : nullptr, t->d_bpay[fslot]);
    } else if (big_sort) {
      // big inputs: osl_sort.cu (9-bit digits, prefetched tiles, keys only for voxel grids -- Q11: their colour index
      // is the sorted position)
      rc = osl_sort_big(&t->sort_ws[fslot], t->d_keysA[fslot], t->d_payA[fslot], t->d_keysB[fslot], t->d_payB[fslot],
                        &fs->acc_emit[fslot], ep.mode == 2 ? &fs->acc_unsorted[fslot] : nullptr, exp_emit, 3 * D,
                        ep.mode != 2, piped ? coop_cap : 0, sSo);
      if (rc) return rc;
      OSL_LAUNCHED(-1);  // (counted below)
    } else {
      const int grid = grid_for(exp_emit, SORT_TILE, t->sort_grid < coop_cap ? t->sort_grid : coop_cap);
      u64* kA = t->d_keysA[fslot]; u32* pA = t->d_payA[fslot]; u64* kB = t->d_keysB[fslot]; u32* pB = t->d_payB[fslot];
      u32* ch = t->d_cta_hist[fslot]; const FrameState* fsc = fs; int pp = passes; int parity = fslot;
      int md = ep.mode;
      void* args[] = {&kA, &pA, &kB, &pB, &ch, &fsc, &pp, &parity, &md};
      OSL_CUDA(cudaLaunchCooperativeKernel((void*)k_sort, dim3(grid), dim3(SORT_THREADS), args, 0, sSo));
    }
    OSL_LAUNCHED(1);
    if (timing) OSL_CUDA(cudaEventRecord(t->stage_ev[2], st));
    if (piped) {
      OSL_CUDA(cudaEventRecord(t->sort_done[fslot], sSo));
      OSL_CUDA(cudaStreamWaitEvent(sS, t->sort_done[fslot], 0));
    }
  }
  // ---- S: structure plan + child pointers
  {  // readers of the pool queued before this call (raycasts, foreign work on joined streams) finish first: the
     // structure stage rewrites word0 / new tiles, the value stage rewrites word1
    osl_status rr = osl_order_after_readers(t, st, piped, sS, sV);
    if (rr) return rr;
  }
  if (piped && f >= OSL_BACK)  // level lists + result block of this slot were last read by k_levels of frame f - 2
    OSL_CUDA(cudaStreamWaitEvent(sS, t->ring_ev[(f - OSL_BACK) % OSL_RING], 0));
  {
    static const int no_big = getenv("OSL_NO_BIG") ? 1 : 0;
    const bool big = exp_emit >= (1ll << 20) && t->structure_big_grid > 0 && !no_big;
    const int gmax = big ? t->structure_big_grid : t->structure_grid;
    const int grid = grid_for(exp_emit, AN_THREADS, gmax < coop_cap ? gmax : coop_cap);
    StructArgs A = make_struct_args(t, skeys, spay, fs, fr, f, lv, ep.mode, n, fslot);
    void* args[] = {&A};
    // (a plain launch was measured to be no faster than the cooperative one, which guarantees the co-residency the
    // flag exchange relies on)
    OSL_CUDA(cudaLaunchCooperativeKernel(big ? (void*)k_structure_big : (void*)k_structure, dim3(grid), dim3(AN_THREADS),
                                         args, STRUCT_SMEM, sS));
    OSL_LAUNCHED(1);
    if (timing) OSL_CUDA(cudaEventRecord(t->stage_ev[3], st));
    if (piped) {
      OSL_CUDA(cudaEventRecord(t->struct_ev[f % OSL_RING], sS));
      OSL_CUDA(cudaStreamWaitEvent(sV, t->struct_ev[f % OSL_RING], 0));
    }
  }
  // ---- V: values, bottom-up
  if (n > 0) {
    const int grid = grid_for(exp_level, LEVEL_THREADS, t->levels_grid < coop_cap ? t->levels_grid : coop_cap);
    LevelArgs A = make_level_args(t, lv, fr, f, ep.mode, ep.rgb, (const float*)colors);
    void* args[] = {&A};
    OSL_CUDA(cudaLaunchCooperativeKernel((void*)k_levels, dim3(grid), dim3(LEVEL_THREADS), args, LEVEL_SMEM, sV));
    OSL_LAUNCHED(1);
  }
  if (timing) { OSL_CUDA(cudaEventRecord(t->stage_ev[4], st)); t->stage_valid = 1; }
  // end of frame: k_structure has written the result block into the pinned ring slot; the host reads it lazily
  // after this event (the caller never waits for it unless it asks for sizes / counters)
  const int slot = (int)(f % OSL_RING);
  OSL_CUDA(cudaEventRecord(t->ring_ev[slot], sV));
  if (host) {
    OSL_CUDA(cudaEventRecord(t->stage_free[sslot], sV));  // the colours are last read by k_levels
    t->stage_seq++;
  }
  t->ring_kind[slot] = 0;
  t->join_pending = piped ? 1 : 0;  // other streams are ordered after the pipeline by osl_svo_join (lazily)
  t->ring_headroom[slot] = headroom;
  t->ring_mode[slot] = ep.mode;
  t->inflight_headroom += headroom;
  t->ring_head++;
  t->seq++;
  t->last_stream = st;
  t->last_piped = piped ? 1 : 0;
  t->last_fused = 0;
  return OSL_OK;
}

// The pool's readers.  osl_note_reader: called by the library's own asynchronous readers right after they enqueue
// their kernel on `st`.  osl_note_foreign_reader: `st` was handed to osl_svo_join, so foreign work that reads the pool
// may be queued on it until the next integrate call; the event is recorded then.
osl_status osl_note_reader(osl_svo* t, cudaStream_t st) {
  if (!t->reader_ev) OSL_CUDA(cudaEventCreateWithFlags(&t->reader_ev, cudaEventDisableTiming));
  if (t->reader_pending) OSL_CUDA(cudaStreamWaitEvent(st, t->reader_ev, 0));  // keep earlier readers covered
  OSL_CUDA(cudaEventRecord(t->reader_ev, st));
  t->reader_pending = 1;
  return OSL_OK;
}

osl_status osl_note_foreign_reader(osl_svo* t, cudaStream_t st) {
  for (int i = 0; i < t->foreign_n; i++)
    if (t->foreign_reader[i] == st) return OSL_OK;
  if (t->foreign_n == 8) {  // table full: cover what the oldest stream carries so far
    osl_status rc = osl_note_reader(t, t->foreign_reader[0]);
    if (rc) return rc;
    for (int i = 1; i < 8; i++) t->foreign_reader[i - 1] = t->foreign_reader[i];
    t->foreign_n = 7;
  }
  t->foreign_reader[t->foreign_n++] = st;
  return OSL_OK;
}

// Called when a frame is enqueued: its pool-writing stages (streams sS, sV) wait for every reader noted so far.
// Strict-mode frames on the stream the readers ran on are ordered already.
static osl_status osl_order_after_readers(osl_svo* t, cudaStream_t st, bool piped, cudaStream_t sS, cudaStream_t sV) {
  for (int i = 0; i < t->foreign_n; i++) {
    if (!piped && t->foreign_reader[i] == st) continue;
    osl_status rc = osl_note_reader(t, t->foreign_reader[i]);
    if (rc) return rc;
  }
  t->foreign_n = 0;
  if (t->reader_pending) {
    OSL_CUDA(cudaStreamWaitEvent(sS, t->reader_ev, 0));
    if (sV != sS) OSL_CUDA(cudaStreamWaitEvent(sV, t->reader_ev, 0));
    t->reader_pending = 0;
  }
  return OSL_OK;
}

// Orders `stream` after every frame enqueued so far (one cudaStreamWaitEvent).  Strict-mode frames are already
// stream-ordered; pipelined frames complete on an internal stream.
osl_status osl_join(osl_svo* t, cudaStream_t st) {
  if (t->seq == 0) return OSL_OK;
  if (t->last_fused && !t->fz_event_valid) {  // k_frame path: issue the remaining stages, then mark the end of the frame
    osl_status rc = osl_fused_flush(t);
    if (rc) return rc;
    OSL_CUDA(cudaEventRecord(t->ring_ev[(t->seq - 1) % OSL_RING], t->pipe[2]));
    t->fz_event_valid = 1;
  }
  if (t->join_pending || st != t->last_stream)
    OSL_CUDA(cudaStreamWaitEvent(st, t->ring_ev[(t->seq - 1) % OSL_RING], 0));
  return OSL_OK;
}

// Stand-alone use of the cooperative LSD radix sort (k_sort) for callers outside the frame pipeline (the mesh
// voxeliser): sorts n (key, value) pairs by the low key_bits bits; *in_B tells which buffer holds the result.
osl_status osl_device_sort_pairs(u64* kA, u32* pA, u64* kB, u32* pB, int n, int key_bits, cudaStream_t st, int* in_B) {
  const int passes = (key_bits + 7) / 8;
  *in_B = passes & 1;
  if (n <= 0) return OSL_OK;
  if (n >= (1 << 19) && !getenv("OSL_NO_BIG_SORT")) {  // big inputs: osl_sort.cu
    // (a process-wide workspace per device, kept for the next call; the sort is stream-ordered and the callers --
    // the voxelisers -- run one at a time per device)
    static OslSortWs s_ws[64];
    static int* s_dn[64] = {nullptr};
    int dev = 0;
    OSL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return OSL_ERR_INVALID;
    *in_B = osl_sort_big_passes(key_bits, nullptr) & 1;
    if (!s_dn[dev]) OSL_CUDA(cudaMalloc(&s_dn[dev], sizeof(int)));
    static int s_n[64];
    s_n[dev] = n;
    OSL_CUDA(cudaMemcpyAsync(s_dn[dev], &s_n[dev], sizeof(int), cudaMemcpyHostToDevice, st));
    osl_status rc = osl_sort_big(&s_ws[dev], kA, pA, kB, pB, s_dn[dev], nullptr, n, key_bits, true, 0, st);
    if (rc) return rc;
    OSL_CUDA(cudaStreamSynchronize(st));  // (s_n is pageable: the copy above has been staged by now; callers expect completion)
    return OSL_OK;
  }
  int dev = 0, sms = 0;
  OSL_CUDA(cudaGetDevice(&dev));
  OSL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int occ = osl_sort_occupancy();
  if (occ < 1) return OSL_ERR_CUDA;
  const int cap = (occ > 4 ? 4 : occ) * sms;
  const int grid = grid_for(n, SORT_TILE, cap);
  FrameState* fs = nullptr;
  u32* hist = nullptr;
  OSL_CUDA(cudaMalloc(&fs, sizeof(FrameState)));
  cudaError_t e = cudaMalloc(&hist, (size_t)grid * 256 * sizeof(u32));
  if (e == cudaSuccess) e = cudaMemsetAsync(fs, 0, sizeof(FrameState), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&fs->acc_emit[0], &n, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    const FrameState* fsc = fs; int pp = passes; int parity = 0; int md = 0;
    void* args[] = {&kA, &pA, &kB, &pB, &hist, &fsc, &pp, &parity, &md};
    e = cudaLaunchCooperativeKernel((void*)k_sort, dim3(grid), dim3(SORT_THREADS), args, 0, st);
    OSL_LAUNCHED(1);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(hist);
  cudaFree(fs);
  OSL_CUDA(e);
  return OSL_OK;
}

// Tracing aid (tools/frame_timeline.py): enable = 1 clears the span table and records the role spans of the next
// k_frame launches of `t` (round-robin over 32 slots); out (optional) receives the table, 32 x 5 x 2 nanosecond values.
extern "C" osl_status osl_debug_trace(osl_svo* t, int enable, unsigned long long* out) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaDeviceSynchronize());
  if (out) OSL_CUDA(cudaMemcpyFromSymbol(out, g_osl_span, sizeof(unsigned long long) * OSL_SPAN_SLOTS * 5 * 2));
  if (enable) {
    unsigned long long init[OSL_SPAN_SLOTS][5][2];
    for (int i = 0; i < OSL_SPAN_SLOTS; i++)
      for (int r = 0; r < 5; r++) { init[i][r][0] = ~0ull; init[i][r][1] = 0ull; }
    OSL_CUDA(cudaMemcpyToSymbol(g_osl_span, init, sizeof(init)));
    t->trace_seq = 0;
  }
  t->trace_on = enable ? 1 : 0;
  return OSL_OK;
}

// Test aid: overwrites the hint table of the tree walks (walk_frontier) with pseudo-random words -- node indices inside
// the pool, beyond it, and 0xFFFFFFFF -- to show that no result depends on it.  Synchronizes.
__global__ void k_scramble_hints(u32* hints, u32 n, u32 seed, u32 size) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x = (i + 1u) * 0x9E3779B1u ^ seed;
  x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
  const u32 kind = x & 3u;
  hints[i] = kind == 0 ? 0xFFFFFFFFu : kind == 1 ? x : (x >> 2) % (size ? size : 1u);
}
extern "C" osl_status osl_debug_scramble_hints(osl_svo* t, unsigned seed) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaDeviceSynchronize());
  if (!t->d_wcache) return OSL_OK;
  k_scramble_hints<<<(WH_SLOTS + 255) / 256, 256>>>(reinterpret_cast<u32*>(t->d_wcache + WC_SLOTS), WH_SLOTS, seed, (u32)t->size);
  OSL_CUDA(cudaGetLastError());
  OSL_CUDA(cudaDeviceSynchronize());
  return OSL_OK;
}

// ------------------------------------------------------------------------------------------------ sharded build
// ONE map built by several GPUs ("partition the key space by contiguous Morton ranges ... each GPU
// all-gathers its per-pass split counts -> exclusive prefix over GPUs -> local ranks become global indices").  Input: a
// voxel grid in Morton order without invalid entries (what the voxelisers and the extraction emit), present on every
// rank; rank r takes the contiguous slice [lo, hi).  Three calls per rank, the exchange between them is the caller's
// (shard.integrate_voxels_sharded: one all-gather of NC counters, one of the deltas):
//   osl_shard_analyze   keys of the slice, phase A of k_structure with the lower rank's last key as the predecessor of
//                       the slice's first key -> this rank's counter totals
//   osl_shard_assign    phase B2 + C with the bucket counters made global (totals of all ranks for the plan, the lower
//                       ranks' counts in front of this rank's), then the value fold of this rank's sub-trees
//   osl_shard_fixup     (osl_replica.cu) after the ranks exchanged what they changed: the nodes on the paths of the
//                       slices' first keys -- the only ones with children in two ranks -- are re-averaged bottom-up, and
//                       the root average (Q6) is written
// Node indices, child pointers and values come out bit-identical with a single-GPU osl_integrate_voxels of the whole
// grid (tests/test_gpu_multirank.py).
__global__ void k_key_of(const float* __restrict__ pts, int stride, long long idx, TreeParams tp, u64* out, int* valid) {
  const float* q = pts + (size_t)stride * idx;
  u64 k;
  const bool ok = osl_key(q[0], q[1], q[2], tp, k);
  *out = k;
  *valid = ok ? 1 : 0;
}

extern "C" osl_status osl_shard_analyze(osl_svo* t, const float* d_centers4, int n_total, int lo, int hi,
                                        uint32_t* h_totals, int* n_counters, void* stream) {
  if (!t || !d_centers4 || n_total <= 0 || lo < 0 || hi < lo || hi > n_total || !h_totals) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int n = hi - lo, D = t->tp.D, NC = OSL_NCOUNT(D);
  if (n_counters) *n_counters = NC;
  osl_status rc = osl_poll_results(t, true);  // exact size, idle pipeline: the three calls run in stream order on `st`
  if (rc) return rc;
  rc = drain_streams(t, st);
  if (rc) return rc;
  if (t->sticky_error) return t->sticky_error;
  rc = osl_ensure_workspace(t, (size_t)(n > 0 ? n : 1));
  if (rc) return rc;
  const unsigned long long f = t->seq;
  const int fslot = (int)(f % OSL_FRONT), bslot = (int)(f % OSL_BACK);
  FrameState* fs = t->d_fs;
  // the predecessor of the slice's first key (the grid is on every rank: no exchange needed)
  u64 prev_key = 0;
  int has_prev = 0;
  u64* d_tmp = reinterpret_cast<u64*>(t->d_scan_totals);  // scratch (idle pipeline)
  if (lo > 0) {
    k_key_of<<<1, 1, 0, st>>>(d_centers4, 4, (long long)lo - 1, t->tp, d_tmp, reinterpret_cast<int*>(d_tmp + 1));
    OSL_LAUNCHED(1);
    u64 h[2];
    OSL_CUDA(cudaMemcpyAsync(h, d_tmp, sizeof(h), cudaMemcpyDeviceToHost, st));
    OSL_CUDA(cudaStreamSynchronize(st));
    OSL_CUDA(cudaMemsetAsync(d_tmp, 0, sizeof(h), st));
    if (!(int)h[1]) return OSL_ERR_UNSUPPORTED;  // invalid voxels sort to the front in the reference: not a slice-able grid
    prev_key = h[0];
    has_prev = 1;
  }
  EmitParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.pts = d_centers4 + 4 * (size_t)lo; ep.stride = 4; ep.n = n; ep.mode = 2;
  t->shard_n = n; t->shard_lo = lo; t->shard_f = f;
  if (n > 0) {
    const int per = GRID_THREADS * GRID_PPT;
    GridFast gf0;
    memset(&gf0, 0, sizeof(gf0));  // (slices of a sharded build: the plain descent)
    k_emit_grid<<<(n + per - 1) / per, GRID_THREADS, 0, st>>>(ep.pts, ep.stride, n, t->tp, gf0, t->d_keysA[fslot],
                                                              t->d_payA[fslot], fs, fslot);
    OSL_LAUNCHED(1);
  }
  // phase A + counter exchange among this rank's CTAs; the totals come back to the host
  u32* d_tot = t->d_blockcnt_tot;
  {
    const bool big = n >= (1 << 20) && t->structure_big_grid > 0;  // (osl_shard_assign takes the same kernel and grid)
    const int grid = grid_for(n, AN_THREADS, big ? t->structure_big_grid : t->structure_grid);
    StructArgs A = make_struct_args(t, t->d_keysA[fslot], t->d_payA[fslot], fs, t->d_fs + 1 + bslot, f, t->lv[bslot], 2, n, fslot);
    A.shard = 1; A.has_prev = has_prev; A.prev_key = prev_key; A.rank_tot = d_tot; A.src_base = (u32)lo;
    void* args[] = {&A};
    OSL_CUDA(cudaLaunchCooperativeKernel(big ? (void*)k_structure_big : (void*)k_structure, dim3(grid), dim3(AN_THREADS), args,
                                         STRUCT_SMEM, st));
    OSL_LAUNCHED(1);
    t->shard_grid = grid;
  }
  int h_flags[2] = {0, 0};
  OSL_CUDA(cudaMemcpyAsync(h_totals, d_tot, NC * sizeof(u32), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaMemcpyAsync(&h_flags[0], &fs->acc_unsorted[fslot], sizeof(int), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaMemcpyAsync(&h_flags[1], &fs->acc_valid[fslot], sizeof(int), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaStreamSynchronize(st));
  if (n > 0 && (h_flags[0] != 0 || h_flags[1] != n)) {  // unsorted, or invalid voxels: the slices would not be key ranges
    int zero[3] = {0, 0, 0};
    cudaMemcpy(&fs->acc_unsorted[fslot], &zero[0], sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(&fs->acc_valid[fslot], &zero[0], sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(&fs->acc_emit[fslot], &zero[0], sizeof(int), cudaMemcpyHostToDevice);
    return OSL_ERR_UNSUPPORTED;
  }
  if (has_prev && n > 0) {  // the slices must be key ranges in rank order
    u64 first = 0;
    OSL_CUDA(cudaMemcpy(&first, t->d_keysA[fslot], sizeof(u64), cudaMemcpyDeviceToHost));
    if (first < prev_key) return OSL_ERR_UNSUPPORTED;
  }
  return OSL_OK;
}

extern "C" osl_status osl_shard_assign(osl_svo* t, const float* d_colors4, const uint32_t* h_base, const uint32_t* h_totals,
                                       void* stream) {
  if (!t || !h_base || !h_totals || !d_colors4) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int D = t->tp.D, NC = OSL_NCOUNT(D), n = t->shard_n;
  const unsigned long long f = t->shard_f;
  if (f != t->seq) return OSL_ERR_INVALID;  // no osl_shard_analyze pending
  const int fslot = (int)(f % OSL_FRONT), bslot = (int)(f % OSL_BACK);
  FrameState* fs = t->d_fs;
  // the pool must hold what ALL ranks append: 8 nodes per split, over every bucket
  unsigned long long splits = 0;
  for (int c = D; c < NC; c++) splits += h_totals[c];
  {
    const size_t want = (size_t)(t->size > 8 ? t->size : 8) + 8ull * splits;
    if (want > ((size_t)1 << 30)) return OSL_ERR_POOL_OVERFLOW;
    if (want > t->cap_nodes) {
      osl_status rc = osl_grow_pool(t, want, st);
      if (rc) return rc;
    }
  }
  u32* d_ext = t->d_blockcnt_tot + NC_MAX;  // [2][NC]
  OSL_CUDA(cudaMemcpyAsync(d_ext, h_base, NC * sizeof(u32), cudaMemcpyHostToDevice, st));
  OSL_CUDA(cudaMemcpyAsync(d_ext + NC, h_totals, NC * sizeof(u32), cudaMemcpyHostToDevice, st));
  {
    StructArgs A = make_struct_args(t, t->d_keysA[fslot], t->d_payA[fslot], fs, t->d_fs + 1 + bslot, f, t->lv[bslot], 2, n, fslot);
    A.shard = 2; A.ext = d_ext; A.src_base = (u32)t->shard_lo;
    void* args[] = {&A};
    const bool big = n >= (1 << 20) && t->structure_big_grid > 0;
    OSL_CUDA(cudaLaunchCooperativeKernel(big ? (void*)k_structure_big : (void*)k_structure, dim3(t->shard_grid), dim3(AN_THREADS),
                                         args, STRUCT_SMEM, st));
    OSL_LAUNCHED(1);
  }
  if (n > 0) {
    const int grid = grid_for(n, LEVEL_THREADS, t->levels_grid);
    LevelArgs A = make_level_args(t, t->lv[bslot], t->d_fs + 1 + bslot, f, 2, nullptr, d_colors4);
    A.no_root = 1;
    void* args[] = {&A};
    OSL_CUDA(cudaLaunchCooperativeKernel((void*)k_levels, dim3(grid), dim3(LEVEL_THREADS), args, LEVEL_SMEM, st));
    OSL_LAUNCHED(1);
  }
  const int slot = (int)(f % OSL_RING);
  OSL_CUDA(cudaEventRecord(t->ring_ev[slot], st));
  t->ring_kind[slot] = 0;
  t->join_pending = 0;
  t->ring_headroom[slot] = 0;
  t->ring_mode[slot] = 2;
  t->ring_head++;
  t->seq++;
  t->last_stream = st;
  t->last_piped = 0;
  t->last_fused = 0;
  return osl_poll_results(t, true);
}

// per-CTA phase clocks of the last k_structure_big launch: out[4][1024]
extern "C" osl_status osl_debug_cta_profile(unsigned long long* out) {
  if (!out) return OSL_ERR_INVALID;
  OSL_CUDA(cudaDeviceSynchronize());
  OSL_CUDA(cudaMemcpyFromSymbol(out, g_osl_ctaprof, sizeof(unsigned long long) * 4 * 1024));
  return OSL_OK;
}

extern "C" osl_status osl_debug_profile(unsigned long long* out, int n) {
  if (!out || n < 0 || n > 128) return OSL_ERR_INVALID;
  OSL_CUDA(cudaDeviceSynchronize());
  OSL_CUDA(cudaMemcpyFromSymbol(out, g_osl_prof, sizeof(unsigned long long) * (size_t)n));
  return OSL_OK;
}
