// osl_voxelize.cu -- sparse surface voxelisation of a triangle mesh into a VoxelGrid (cell centres + colours).
//
// Replaces voxelization::meshToVoxelGrid (voxelization.h:19-21, voxelization.cu:238-323,381-405) as the FEEDER of
// svoFromVoxelGrid for BASELINE configs 2 and 5.  The reference rasterises into a DENSE 256^3 float framebuffer with
// the vendored voxelpipe (compile-time GRID_RES 8, voxelization.cu:24) and does not build with CUDA 12 (SURVEY.md
// section 8c), so there is no reference output to match: the contract here is our own, restated on the CPU by
// oracle/osl_oracle.c (orc_voxelize_mesh, brute force) and checked bit-exactly against it:
//   a leaf cell of the depth-D grid over the tree cube is occupied iff its box overlaps a triangle (separating-axis
//   test, touching counts); its colour is the colour of the LOWEST-numbered triangle overlapping it; the grid is
//   returned in Morton-key order, so svoFromVoxelGrid's key sort is the identity and quirk Q11 (colours are not
//   permuted with the keys) leaves every colour on its own voxel.
// Sparse and depth-generic (any D <= 20): work is proportional to the projected area of the triangles, never to 8^D.
//
//   k_vox_setup   per triangle: cell bounding box, dominant axis, number of 1024-column chunks of its projection
//   k_vox_items   work items (triangle, chunk), appended with one atomicAdd per triangle
//   k_vox_raster  one warp per item: for every column of the chunk the cell range the triangle's plane crosses, SAT
//                 on those cells; pass 1 counts the hits of every item, pass 2 appends (Morton key, triangle) pairs to
//                 a list -- one atomicAdd per item reserves its stretch, a shared-memory counter places the hits in it
//   k_sort_big    the pairs by key (osl_sort.cu)
//   k_vox_heads / k_vox_scan / k_vox_unique   one entry per cell: lowest triangle of each run of equal keys,
//                 order-preserving compaction (per-block head counts, their scan, placement)
//   k_vox_finish  centres + colours
// (First version: the hits went into a global hash set -- 64-bit CAS on the cell, atomicMin on the triangle -- which was
// then compacted and sorted.  At 57.5 M voxels the random atomics cost 8.9 ms and the table scan 3.3 ms of 18 ms; a
// locality-preserving slot function made it 40x worse: neighbouring cells then queue on the same L2 atomic units.)
// Compiled with -fmad=false: every product and sum below rounds exactly as in the CPU restatement.
#include "osl_internal.cuh"

#define VX_CHUNK 1024
#define VX_EMPTY 0xFFFFFFFFFFFFFFFFull

struct VoxGrid {
  float lox, loy, loz;  // low corner of the tree cube
  float cs, hs;         // cell size, half cell size
  int G;                // cells per axis = 2^D
  int D;
};

struct VoxTri {  // per-triangle set-up
  int lo[3], hi[3];     // cell bounding box (inclusive)
  int axis;             // dominant axis of the normal
  int ncols;            // columns of the projection on the plane orthogonal to `axis`
};

__device__ __forceinline__ int vx_cell(float p, float lo, float cs, int G) {
  int i = (int)floorf((p - lo) / cs);
  return i < 0 ? 0 : (i > G - 1 ? G - 1 : i);
}

__device__ __forceinline__ void vx_load_tri(const float* __restrict__ V, const int* __restrict__ T, int t, float v[3][3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int vi = __ldg(T + 3 * (size_t)t + k);
#pragma unroll
    for (int c = 0; c < 3; c++) v[k][c] = __ldg(V + 3 * (size_t)vi + c);
  }
}

// separating-axis test of the box (centre c, half size h) against triangle v (Akenine-Moller); touching = overlap
__device__ bool vx_overlap(const float c[3], float h, const float v[3][3]) {
  float a[3][3];
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int d = 0; d < 3; d++) a[k][d] = v[k][d] - c[d];
  // 1. box axes
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const float mn = fminf(a[0][d], fminf(a[1][d], a[2][d])), mx = fmaxf(a[0][d], fmaxf(a[1][d], a[2][d]));
    if (mn > h || mx < -h) return false;
  }
  float e[3][3];
#pragma unroll
  for (int d = 0; d < 3; d++) { e[0][d] = a[1][d] - a[0][d]; e[1][d] = a[2][d] - a[1][d]; e[2][d] = a[0][d] - a[2][d]; }
  // 2. triangle plane
  {
    const float nx = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    const float ny = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    const float nz = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    const float s = (nx * a[0][0] + ny * a[0][1]) + nz * a[0][2];
    const float r = h * ((fabsf(nx) + fabsf(ny)) + fabsf(nz));
    if (s > r || s < -r) return false;
  }
  // 3. the nine cross products unit_k x e_i
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float ex = e[i][0], ey = e[i][1], ez = e[i][2];
    {  // unit_x x e = (0, -ez, ey)
      const float p0 = ey * a[0][2] - ez * a[0][1], p1 = ey * a[1][2] - ez * a[1][1], p2 = ey * a[2][2] - ez * a[2][1];
      const float r = h * (fabsf(ez) + fabsf(ey));
      if (fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r) return false;
    }
    {  // unit_y x e = (ez, 0, -ex)
      const float p0 = ez * a[0][0] - ex * a[0][2], p1 = ez * a[1][0] - ex * a[1][2], p2 = ez * a[2][0] - ex * a[2][2];
      const float r = h * (fabsf(ez) + fabsf(ex));
      if (fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r) return false;
    }
    {  // unit_z x e = (-ey, ex, 0)
      const float p0 = ex * a[0][1] - ey * a[0][0], p1 = ex * a[1][1] - ey * a[1][0], p2 = ex * a[2][1] - ey * a[2][0];
      const float r = h * (fabsf(ey) + fabsf(ex));
      if (fminf(p0, fminf(p1, p2)) > r || fmaxf(p0, fmaxf(p1, p2)) < -r) return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(256)
k_vox_setup(const float* __restrict__ V, const int* __restrict__ T, int nt, VoxGrid g, VoxTri* tri,
            unsigned long long* total_chunks) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  float v[3][3];
  vx_load_tri(V, T, t, v);
  VoxTri o;
  const float lo3[3] = {g.lox, g.loy, g.loz};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    o.lo[d] = vx_cell(fminf(v[0][d], fminf(v[1][d], v[2][d])), lo3[d], g.cs, g.G);
    o.hi[d] = vx_cell(fmaxf(v[0][d], fmaxf(v[1][d], v[2][d])), lo3[d], g.cs, g.G);
  }
  const float e0[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
  const float e1[3] = {v[2][0] - v[1][0], v[2][1] - v[1][1], v[2][2] - v[1][2]};
  const float n[3] = {fabsf(e0[1] * e1[2] - e0[2] * e1[1]), fabsf(e0[2] * e1[0] - e0[0] * e1[2]),
                      fabsf(e0[0] * e1[1] - e0[1] * e1[0])};
  o.axis = (n[0] >= n[1] && n[0] >= n[2]) ? 0 : (n[1] >= n[2] ? 1 : 2);
  const int u = (o.axis + 1) % 3, w = (o.axis + 2) % 3;
  const long long cols = (long long)(o.hi[u] - o.lo[u] + 1) * (long long)(o.hi[w] - o.lo[w] + 1);
  o.ncols = (int)(cols > 0x7FFFFFFFll ? 0x7FFFFFFFll : cols);
  tri[t] = o;
  atomicAdd(total_chunks, (unsigned long long)((o.ncols + VX_CHUNK - 1) / VX_CHUNK));
}

__global__ void __launch_bounds__(256)
k_vox_items(const VoxTri* __restrict__ tri, int nt, uint2* items, unsigned long long* cursor) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const int nch = (tri[t].ncols + VX_CHUNK - 1) / VX_CHUNK;
  const unsigned long long base = atomicAdd(cursor, (unsigned long long)nch);
  for (int c = 0; c < nch; c++) items[base + c] = make_uint2((unsigned)t, (unsigned)c);
}

__device__ __forceinline__ unsigned long long vx_code(int ix, int iy, int iz) {
  return ((unsigned long long)iz << 42) | ((unsigned long long)iy << 21) | (unsigned long long)ix;
}

// leading-1 Morton key of a cell, digit = x + 2y + 4z per level, most significant level first (svo.cu:33-66)
__device__ __forceinline__ u64 vx_morton(int ix, int iy, int iz, int D) {
  u64 k = 1;
  for (int l = D - 1; l >= 0; l--)
    k = (k << 3) | (u64)(((ix >> l) & 1) | (((iy >> l) & 1) << 1) | (((iz >> l) & 1) << 2));
  return k;
}

// INSERT = false: count the cells every item overlaps; true: append them, as (Morton key, triangle) pairs, to the list
template <bool INSERT>
__global__ void __launch_bounds__(256)
k_vox_raster(const float* __restrict__ V, const int* __restrict__ T, const VoxTri* __restrict__ tri,
             const uint2* __restrict__ items, unsigned long long n_items, VoxGrid g, unsigned long long* hits,
             u32* item_hits, u64* out_keys, u32* out_tris) {
  const unsigned long long item = (unsigned long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  unsigned long long local = 0;
  __shared__ unsigned s_cur[8];            // INSERT: hits placed so far in this warp's stretch
  __shared__ unsigned long long s_base[8]; // INSERT: where the stretch starts
  if (INSERT) {
    if (lane == 0) {
      s_cur[threadIdx.x >> 5] = 0;
      const u32 cnt = item < n_items ? item_hits[item] : 0u;
      s_base[threadIdx.x >> 5] = cnt ? atomicAdd(hits, (unsigned long long)cnt) : 0ull;
    }
    __syncwarp();
  }
  if (item < n_items) {
    const uint2 it = items[item];
    const int t = (int)it.x;
    const VoxTri o = tri[t];
    float v[3][3];
    vx_load_tri(V, T, t, v);
    const int A = o.axis, U = (A + 1) % 3, W = (A + 2) % 3;
    const float lo3[3] = {g.lox, g.loy, g.loz};
    // plane n.x = d for the candidate range along the dominant axis
    const float e0[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
    const float e1[3] = {v[2][0] - v[1][0], v[2][1] - v[1][1], v[2][2] - v[1][2]};
    const float n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    const float dpl = (n[0] * v[0][0] + n[1] * v[0][1]) + n[2] * v[0][2];
    const bool flat = fabsf(n[A]) > 0.0f;
    const int nu = o.hi[U] - o.lo[U] + 1;
    const int c0 = (int)it.y * VX_CHUNK;
    const int c1 = min(o.ncols, c0 + VX_CHUNK);
    for (int col = c0 + lane; col < c1; col += 32) {
      const int iu = o.lo[U] + col % nu, iw = o.lo[W] + col / nu;
      int a0 = o.lo[A], a1 = o.hi[A];
      if (flat) {  // range of the plane over the column footprint, +-1 cell of slack; the SAT below decides
        const float u0 = lo3[U] + (float)iu * g.cs, u1 = u0 + g.cs, w0 = lo3[W] + (float)iw * g.cs, w1 = w0 + g.cs;
        float amin = 3.0e38f, amax = -3.0e38f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float uu = (k & 1) ? u1 : u0, ww = (k & 2) ? w1 : w0;
          const float aa = ((dpl - n[U] * uu) - n[W] * ww) / n[A];
          amin = fminf(amin, aa); amax = fmaxf(amax, aa);
        }
        if (amin == amin && amax == amax) {  // not NaN
          const float fa0 = floorf((amin - lo3[A]) / g.cs) - 1.0f, fa1 = floorf((amax - lo3[A]) / g.cs) + 1.0f;
          if (fa0 > (float)a0) a0 = fa0 > (float)a1 ? a1 + 1 : (int)fa0;
          if (fa1 < (float)a1) a1 = fa1 < (float)o.lo[A] ? o.lo[A] - 1 : (int)fa1;
        }
      }
      for (int ia = a0; ia <= a1; ia++) {
        int idx[3];
        idx[A] = ia; idx[U] = iu; idx[W] = iw;
        const float c[3] = {lo3[0] + ((float)idx[0] + 0.5f) * g.cs, lo3[1] + ((float)idx[1] + 0.5f) * g.cs,
                            lo3[2] + ((float)idx[2] + 0.5f) * g.cs};
        if (!vx_overlap(c, g.hs, v)) continue;
        if (!INSERT) {
          local++;
        } else {
          const unsigned long long pos = s_base[threadIdx.x >> 5] + atomicAdd(&s_cur[threadIdx.x >> 5], 1u);
          out_keys[pos] = vx_morton(idx[0], idx[1], idx[2], g.D);
          out_tris[pos] = (u32)t;
        }
      }
    }
  }
  if (!INSERT) {
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o2);
    if (lane == 0 && item < n_items) item_hits[item] = (u32)local;
    if (lane == 0 && local) atomicAdd(hits, local);
  }
}

// One entry per cell from the sorted pair list (runs of equal keys = the triangles that overlap the cell).
#define VX_UBLOCK 2048
__global__ void __launch_bounds__(256)
k_vox_heads(const u64* __restrict__ keys, long long n, u32* blockcnt) {
  const long long b0 = (long long)blockIdx.x * VX_UBLOCK;
  int c = 0;
  for (int i = threadIdx.x; i < VX_UBLOCK; i += 256) {
    const long long j = b0 + i;
    if (j < n && (j == 0 || keys[j] != keys[j - 1])) c++;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  __shared__ int s_c[8];
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; w++) t += s_c[w];
    blockcnt[blockIdx.x] = (u32)t;
  }
}

// exclusive scan of the block counts (one CTA; a few ten thousand values), total -> *count
__global__ void __launch_bounds__(1024)
k_vox_scan(u32* blockcnt, int nb, unsigned long long* count) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const u32 v = b < nb ? blockcnt[b] : 0u;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long woff = 0, tot = 0;
    for (int w = 0; w < 32; w++) {
      if (w < (int)(threadIdx.x >> 5)) woff += s_w[w];
      tot += s_w[w];
    }
    const unsigned long long base = s_run;
    if (b < nb) blockcnt[b] = (u32)(base + woff + incl - v);
    __syncthreads();
    if (threadIdx.x == 0) s_run = base + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_run;
}

__global__ void __launch_bounds__(256)
k_vox_unique(const u64* __restrict__ keys, const u32* __restrict__ tris, long long n, const u32* __restrict__ blockbase,
             u64* okeys, u32* otris) {
  const long long b0 = (long long)blockIdx.x * VX_UBLOCK;
  __shared__ int s_w[8];
  __shared__ int s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < VX_UBLOCK; i0 += 256) {  // (consecutive threads take consecutive entries: the order is kept)
    const long long j = b0 + i0 + threadIdx.x;
    const bool head = j < n && (j == 0 || keys[j] != keys[j - 1]);
    const u32 bal = __ballot_sync(0xFFFFFFFFu, head);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      if (w < warp) woff += s_w[w];
      tot += s_w[w];
    }
    if (head) {
      const u64 k = keys[j];
      u32 tm = tris[j];
      for (long long jj = j + 1; jj < n && keys[jj] == k; jj++) tm = min(tm, tris[jj]);  // lowest triangle of the cell
      const long long pos = (long long)blockbase[blockIdx.x] + s_run + woff + __popc(bal & ((1u << lane) - 1u));
      okeys[pos] = k;
      otris[pos] = tm;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_run += tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_vox_finish(const u64* __restrict__ keys, const u32* __restrict__ tris, long long n, VoxGrid g,
             const float4* __restrict__ tri_colors, float4* centers, float4* colors, long long* keys_out,
             int* tris_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 k = keys[i];
  int ix = 0, iy = 0, iz = 0;
  for (int l = g.D - 1; l >= 0; l--) {
    const int dgt = (int)((k >> (3 * l)) & 7);
    ix = (ix << 1) | (dgt & 1); iy = (iy << 1) | ((dgt >> 1) & 1); iz = (iz << 1) | ((dgt >> 2) & 1);
  }
  centers[i] = make_float4(g.lox + ((float)ix + 0.5f) * g.cs, g.loy + ((float)iy + 0.5f) * g.cs,
                           g.loz + ((float)iz + 0.5f) * g.cs, 1.0f);
  const u32 t = tris[i];
  colors[i] = tri_colors ? __ldg(tri_colors + t) : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
  if (keys_out) keys_out[i] = (long long)k;
  if (tris_out) tris_out[i] = (int)t;
}

extern "C" void osl_free_device(void* p) { cudaFree(p); }

extern "C" osl_status osl_copy_device(void* d_dst, const void* d_src, size_t bytes) {
  if (!d_dst || !d_src) return OSL_ERR_INVALID;
  OSL_CUDA(cudaMemcpy(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice));
  return OSL_OK;
}

extern "C" osl_status osl_voxelize_mesh(const float* d_vertices, int n_vertices, const int* d_triangles, int n_triangles,
                                        const float* d_tri_colors4, const float center[3], float half_edge,
                                        int max_depth, float** d_centers4_out, float** d_colors4_out,
                                        int64_t** d_keys_out, int** d_tris_out, int64_t* n_out, void* stream) {
  if (!n_out || !center || max_depth < 1 || max_depth > OSL_MAX_DEPTH || !(half_edge > 0.0f) || n_triangles < 0 ||
      n_vertices < 0 || (n_triangles > 0 && (!d_vertices || !d_triangles)))
    return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  *n_out = 0;
  {
    static bool pool_ready = false;  // keep the scratch of one call in the pool for the next one
    if (!pool_ready) {
      int dev = 0;
      cudaMemPool_t mp;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_ready = true;
    }
  }
  if (d_centers4_out) *d_centers4_out = nullptr;
  if (d_colors4_out) *d_colors4_out = nullptr;
  if (d_keys_out) *d_keys_out = nullptr;
  if (d_tris_out) *d_tris_out = nullptr;
  if (n_triangles == 0) return OSL_OK;
  VoxGrid g;
  g.D = max_depth; g.G = 1 << max_depth;
  g.lox = center[0] - half_edge; g.loy = center[1] - half_edge; g.loz = center[2] - half_edge;
  g.cs = (2.0f * half_edge) / (float)g.G;
  g.hs = g.cs * 0.5f;

  VoxTri* tri = nullptr;
  unsigned long long* d_ctr = nullptr;  // [0] chunks, [1] cursor, [2] hits, [3] unique
  uint2* items = nullptr;
  u32* item_hits = nullptr;
  u32* blockcnt = nullptr;
  u64 *kA = nullptr, *kB = nullptr;
  u32 *pA = nullptr, *pB = nullptr;
  float4 *centers = nullptr, *colors = nullptr;
  long long* keys_out = nullptr;
  int* tris_out = nullptr;
  osl_status rc = OSL_OK;
  cudaError_t e = cudaSuccess;
  unsigned long long h_ctr[4] = {0, 0, 0, 0};
  long long n = 0;
  unsigned long long H = 0;  // hits = (cell, triangle) pairs
#define VX_CHECK(x) do { e = (x); if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; rc = (e == cudaErrorMemoryAllocation) ? OSL_ERR_OOM : OSL_ERR_CUDA; goto done; } } while (0)
  VX_CHECK(cudaMallocAsync(&tri, sizeof(VoxTri) * (size_t)n_triangles, st));
  VX_CHECK(cudaMallocAsync(&d_ctr, 4 * sizeof(unsigned long long), st));
  VX_CHECK(cudaMemsetAsync(d_ctr, 0, 4 * sizeof(unsigned long long), st));
  k_vox_setup<<<(n_triangles + 255) / 256, 256, 0, st>>>(d_vertices, d_triangles, n_triangles, g, tri, d_ctr);
  OSL_LAUNCHED(1);
  VX_CHECK(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
  VX_CHECK(cudaStreamSynchronize(st));
  if (h_ctr[0] == 0) goto done;
  if (h_ctr[0] > (1ull << 31)) { rc = OSL_ERR_UNSUPPORTED; goto done; }  // > 2^41 columns: not a sparse problem
  VX_CHECK(cudaMallocAsync(&items, sizeof(uint2) * h_ctr[0], st));
  k_vox_items<<<(n_triangles + 255) / 256, 256, 0, st>>>(tri, n_triangles, items, d_ctr + 1);
  OSL_LAUNCHED(1);
  {
    const unsigned long long blocks = (h_ctr[0] + 7) / 8;
    VX_CHECK(cudaMallocAsync(&item_hits, sizeof(u32) * h_ctr[0], st));
    k_vox_raster<false><<<(unsigned)blocks, 256, 0, st>>>(d_vertices, d_triangles, tri, items, h_ctr[0], g, d_ctr + 2,
                                                          item_hits, nullptr, nullptr);
    OSL_LAUNCHED(1);
    VX_CHECK(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
    VX_CHECK(cudaStreamSynchronize(st));
    H = h_ctr[2];
    if (H == 0) goto done;
    if (H >= (1ull << 31)) { rc = OSL_ERR_POOL_OVERFLOW; goto done; }
    VX_CHECK(cudaMallocAsync(&kA, H * 8, st)); VX_CHECK(cudaMallocAsync(&kB, H * 8, st));
    VX_CHECK(cudaMallocAsync(&pA, H * 4, st)); VX_CHECK(cudaMallocAsync(&pB, H * 4, st));
    VX_CHECK(cudaMemsetAsync(d_ctr + 2, 0, sizeof(unsigned long long), st));  // now the append cursor
    k_vox_raster<true><<<(unsigned)blocks, 256, 0, st>>>(d_vertices, d_triangles, tri, items, h_ctr[0], g, d_ctr + 2,
                                                         item_hits, kA, pA);
    OSL_LAUNCHED(1);
  }
  {
    int in_B = 0;
    rc = osl_device_sort_pairs(kA, pA, kB, pB, (int)H, 3 * max_depth, st, &in_B);  // (the leading 1 at bit 3D is common to all keys)
    if (rc) goto done;
    const u64* sk = in_B ? kB : kA;
    const u32* sp = in_B ? pB : pA;
    u64* uk = in_B ? kA : kB;  // the other pair of buffers takes the one-entry-per-cell list
    u32* up = in_B ? pA : pB;
    const int nb = (int)((H + VX_UBLOCK - 1) / VX_UBLOCK);
    VX_CHECK(cudaMallocAsync(&blockcnt, sizeof(u32) * (size_t)nb, st));
    k_vox_heads<<<nb, 256, 0, st>>>(sk, (long long)H, blockcnt);
    k_vox_scan<<<1, 1024, 0, st>>>(blockcnt, nb, d_ctr + 3);
    k_vox_unique<<<nb, 256, 0, st>>>(sk, sp, (long long)H, blockcnt, uk, up);
    OSL_LAUNCHED(3);
    VX_CHECK(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
    VX_CHECK(cudaStreamSynchronize(st));
    n = (long long)h_ctr[3];
    sk = uk; sp = up;
    // (stream-ordered pool allocations as well: a fresh gigabyte from cudaMalloc costs milliseconds of page mapping;
    // the caller releases them with osl_free_device = cudaFree, which hands them back to the pool)
    VX_CHECK(cudaMallocAsync(&centers, sizeof(float4) * (size_t)n, st));
    VX_CHECK(cudaMallocAsync(&colors, sizeof(float4) * (size_t)n, st));
    if (d_keys_out) VX_CHECK(cudaMallocAsync(&keys_out, sizeof(long long) * (size_t)n, st));
    if (d_tris_out) VX_CHECK(cudaMallocAsync(&tris_out, sizeof(int) * (size_t)n, st));
    k_vox_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sk, sp, n, g, reinterpret_cast<const float4*>(d_tri_colors4),
                                                              centers, colors, keys_out, tris_out);
    OSL_LAUNCHED(1);
    VX_CHECK(cudaStreamSynchronize(st));
  }
  *n_out = n;
  if (d_centers4_out) { *d_centers4_out = reinterpret_cast<float*>(centers); centers = nullptr; }
  if (d_colors4_out) { *d_colors4_out = reinterpret_cast<float*>(colors); colors = nullptr; }
  if (d_keys_out) { *d_keys_out = reinterpret_cast<int64_t*>(keys_out); keys_out = nullptr; }
  if (d_tris_out) { *d_tris_out = tris_out; tris_out = nullptr; }
done:
  // scratch comes from the stream-ordered pool (cudaMallocAsync): after the first call these are pool hits, not
  // multi-gigabyte driver allocations; the outputs are plain cudaMalloc (the caller frees them with cudaFree)
  {
    void* scratch[] = {tri, d_ctr, items, item_hits, blockcnt, kA, kB, pA, pB};
    for (void* q : scratch)
      if (q) cudaFreeAsync(q, st);
  }
  cudaFree(centers); cudaFree(colors); cudaFree(keys_out); cudaFree(tris_out);
  return rc;
#undef VX_CHECK
}
