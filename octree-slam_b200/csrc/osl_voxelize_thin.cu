// osl_voxelize_thin.cu -- the REFERENCE's mesh voxelisation rule on the reference's grid.
//
// Replaces voxelization::meshToVoxelGrid (voxelization.cu:238-323,381-405), which hands the mesh to the vendored
// voxelpipe library: THIN_RASTER / NO_BLENDING on a dense 2^8-per-axis grid over the MESH BOUNDING BOX (anisotropic
// cells), 8^3 tiles.  The rule, restated from the library's source (file:line in external/include/voxelpipe):
//   coarse.h:48-103   per triangle: integer bounding box clamped to the grid, dominant axis of the normal
//   coarse.h:646-690  the triangle goes to every tile its box touches; fine.h:1219-1330 per (tile, triangle): clamp the
//                     box to the tile, reject when the triangle's plane misses the tile box
//   fine.h:368-540    per scanline v: the u range from the three 2-D edge functions moved to the pixel corner (2-D
//                     conservative coverage; utils.h:185-231, fine.h:130-152); per column ONE voxel, the one that holds
//                     the plane's depth at the pixel centre (utils.h:236-253), kept when it lies in the tile
// NO_BLENDING is a race between triangles; canon = lowest triangle index (atomicMin).  The colour of a voxel is its
// triangle's flat colour (ColorShader, voxelization.cu:90-139; computed by the caller).
// This file is compiled with -fmad=false and every float operation is written in the source's order, one per
// statement, so that it is bit-identical with the CPU restatement oracle/osl_oracle_thin.c (tests/test_voxelize.py).
// What is NOT pinned: the library was compiled with FMA contraction on and cannot be built any more, so
// a pixel exactly on an edge may have fallen on the other side in the original binary.
//
// One warp per triangle, lanes over the tiles of its box; a dense N^3 table of "lowest triangle" (67 MB at 256^3).
// The occupied cells are compacted, ordered by the Morton key of their centre in the octree cube the grid is meant
// for (ties by cell index: deterministic), and emitted as a VoxelGrid -- in that order svoFromVoxelGrid's key sort is
// the identity and quirk Q11 (colours not permuted with the keys, svo.cu:601-602,629) cannot scramble the colours.
#include "osl_internal.cuh"

struct ThinGrid {
  float b0x, b0y, b0z, b1x, b1y, b1z;
  float dx, dy, dz, ix, iy, iz;  // bbox_delta, inv_bbox_delta (voxelpipe_inline.h:111-118)
  int N, log_N;
};

struct f3 { float x, y, z; };
__device__ __forceinline__ float sel_u(int a, f3 v) { return a == 0 ? v.y : v.x; }
__device__ __forceinline__ float sel_v(int a, f3 v) { return a == 2 ? v.y : v.z; }
__device__ __forceinline__ float sel_w(int a, f3 v) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ int isel_u(int a, const int v[3]) { return a == 0 ? v[1] : v[0]; }
__device__ __forceinline__ int isel_v(int a, const int v[3]) { return a == 2 ? v[1] : v[2]; }
__device__ __forceinline__ int isel_w(int a, const int v[3]) { return a == 0 ? v[0] : (a == 1 ? v[1] : v[2]); }

#define THIN_LOG_T 3
#define THIN_T 8

__global__ void __launch_bounds__(256)
k_thin_raster(const float* __restrict__ verts, const int* __restrict__ tris, int n_tris, ThinGrid g, int* __restrict__ grid) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_tris) return;
  const int N = g.N;
  const float* p0 = verts + 3 * (size_t)tris[3 * t];
  const float* p1 = verts + 3 * (size_t)tris[3 * t + 1];
  const float* p2 = verts + 3 * (size_t)tris[3 * t + 2];
  const f3 v0 = {p0[0], p0[1], p0[2]}, v1 = {p1[0], p1[1], p1[2]}, v2 = {p2[0], p2[1], p2[2]};
  const f3 bbox0 = {g.b0x, g.b0y, g.b0z};
  const f3 delta = {g.dx, g.dy, g.dz}, inv_delta = {g.ix, g.iy, g.iz};
  // setup_triangle (coarse.h:48-103)
  int b0[3], b1[3];
  {
    const float a0[3] = {(v0.x - bbox0.x) * inv_delta.x, (v0.y - bbox0.y) * inv_delta.y, (v0.z - bbox0.z) * inv_delta.z};
    const float a1[3] = {(v1.x - bbox0.x) * inv_delta.x, (v1.y - bbox0.y) * inv_delta.y, (v1.z - bbox0.z) * inv_delta.z};
    const float a2[3] = {(v2.x - bbox0.x) * inv_delta.x, (v2.y - bbox0.y) * inv_delta.y, (v2.z - bbox0.z) * inv_delta.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float lo = fminf(a2[i], fminf(a1[i], a0[i]));
      const float hi = fmaxf(a2[i], fmaxf(a1[i], a0[i]));
      b0[i] = min(max((int)lo, 0), N - 1);
      b1[i] = min(max((int)ceilf(hi), 0), N - 1);
    }
  }
  const f3 e0 = {v1.x - v0.x, v1.y - v0.y, v1.z - v0.z};
  const f3 e1 = {v2.x - v1.x, v2.y - v1.y, v2.z - v1.z};
  const f3 e2 = {v0.x - v2.x, v0.y - v2.y, v0.z - v2.z};
  const f3 n = {e0.z * e2.y - e0.y * e2.z, e0.x * e2.z - e0.z * e2.x, e0.y * e2.x - e0.x * e2.y};  // anti_cross
  const bool byx = fabsf(n.y) > fabsf(n.x), byz = fabsf(n.y) > fabsf(n.z), bzx = fabsf(n.z) > fabsf(n.x);
  const int axis = byx ? (byz ? 1 : 2) : (bzx ? 2 : 0);
  // triangle_setup<AXIS> (utils.h:185-231)
  const float sgn = axis == 0 ? (n.x > 0.0f ? 1.0f : -1.0f) : (axis == 1 ? (n.y < 0.0f ? 1.0f : -1.0f) : (n.z > 0.0f ? 1.0f : -1.0f));
  const f3 edges[3] = {e0, e1, e2}, vs[3] = {v0, v1, v2};
  float a[3], ndu[3], ndv[3], inv_du[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float nx = -sel_v(axis, edges[i]) * sgn, ny = sel_u(axis, edges[i]) * sgn;
    const float t1 = nx * sel_u(axis, vs[i]);
    const float t2 = ny * sel_v(axis, vs[i]);
    float d = -(t1 + t2);
    d = d + fmaxf(0.0f, sel_u(axis, delta) * nx);
    d = d + fmaxf(0.0f, sel_v(axis, delta) * ny);
    const float s1 = nx * sel_u(axis, bbox0);
    const float s2 = ny * sel_v(axis, bbox0);
    a[i] = (s1 + s2) + d;
    ndu[i] = nx * sel_u(axis, delta);
    ndv[i] = ny * sel_v(axis, delta);
    inv_du[i] = __fdiv_rn(1.0f, ndu[i]);
  }
  // plane_setup<AXIS> (utils.h:236-253)
  const float inv_n = __frcp_rn(sel_w(axis, n));
  const float pex = sel_u(axis, n) * inv_n, pey = sel_v(axis, n) * inv_n;
  float pez = pex * sel_u(axis, v0);
  pez = pez + pey * sel_v(axis, v0);
  pez = pez + sel_w(axis, v0);
  pez = pez - sel_w(axis, bbox0);
  pez = pez - pex * sel_u(axis, bbox0);
  pez = pez - pey * sel_v(axis, bbox0);

  const int tx0 = b0[0] >> THIN_LOG_T, ty0 = b0[1] >> THIN_LOG_T, tz0 = b0[2] >> THIN_LOG_T;
  const int ntx = (b1[0] >> THIN_LOG_T) - tx0 + 1, nty = (b1[1] >> THIN_LOG_T) - ty0 + 1, ntz = (b1[2] >> THIN_LOG_T) - tz0 + 1;
  const long long n_tiles = (long long)ntx * nty * ntz;
  for (long long ti = lane; ti < n_tiles; ti += 32) {
    const int tile[3] = {(tx0 + (int)(ti % ntx)) * THIN_T, (ty0 + (int)((ti / ntx) % nty)) * THIN_T,
                         (tz0 + (int)(ti / ((long long)ntx * nty))) * THIN_T};
    int c0[3], c1[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { c0[i] = max(b0[i], tile[i]); c1[i] = min(b1[i], tile[i] + THIN_T - 1); }
    {  // plane / tile-box test (fine.h:1254-1283)
      const float cx = n.x > 0 ? delta.x * THIN_T : 0.0f, cy = n.y > 0 ? delta.y * THIN_T : 0.0f, cz = n.z > 0 ? delta.z * THIN_T : 0.0f;
      float r1 = n.x * (cx - v0.x);
      r1 = r1 + n.y * (cy - v0.y);
      r1 = r1 + n.z * (cz - v0.z);
      float r2 = n.x * (delta.x * THIN_T - cx - v0.x);
      r2 = r2 + n.y * (delta.y * THIN_T - cy - v0.y);
      r2 = r2 + n.z * (delta.z * THIN_T - cz - v0.z);
      float np = n.x * (bbox0.x + tile[0] * delta.x);
      np = np + n.y * (bbox0.y + tile[1] * delta.y);
      np = np + n.z * (bbox0.z + tile[2] * delta.z);
      if ((np + r1) * (np + r2) > 0.0f) continue;
    }
    for (int v = isel_v(axis, c0); v <= isel_v(axis, c1); v++) {  // rasterize<AXIS> (fine.h:368-540)
      const float b[3] = {a[0] + (float)v * ndv[0], a[1] + (float)v * ndv[1], a[2] + (float)v * ndv[2]};
      int min_u = isel_u(axis, c0), max_u = isel_u(axis, c1);
#pragma unroll
      for (int i = 0; i < 3; i++) {  // compute_scanline_bounds (fine.h:130-152)
        if (ndu[i] > 0.0f) min_u = max(min_u, (int)ceilf(-b[i] * inv_du[i]));
        else if (ndu[i] < 0.0f) max_u = min(max_u, (int)(-b[i] * inv_du[i]));
        else if (b[i] < 0.0f) min_u = max_u + 1;
      }
      for (int u = min_u; u <= max_u; u++) {
        const float uf = ((float)u + 0.5f) * sel_u(axis, delta);
        const float vf = ((float)v + 0.5f) * sel_v(axis, delta);
        const float q1 = pex * uf;
        const float q2 = pey * vf;
        const float wf = pez - (q1 + q2);
        const int w = (int)(wf * sel_w(axis, inv_delta));
        if (w >= isel_w(axis, tile) && w < isel_w(axis, tile) + THIN_T) {
          int x, y, z;
          if (axis == 0) { x = w; y = u; z = v; }
          else if (axis == 1) { x = u; y = w; z = v; }
          else { x = u; y = v; z = w; }
          atomicMin(&grid[((size_t)z * N + y) * N + x], t);
        }
      }
    }
  }
}

// getCenterFromIndex (voxelization.cu:58-78): tile and in-tile coordinates, the bbox split into M tiles of T pixels
__device__ __forceinline__ void thin_center(const ThinGrid& g, int x, int y, int z, float& cx, float& cy, float& cz) {
  const int M = g.N >> THIN_LOG_T;
  const float tdx = (g.b1x - g.b0x) / (float)M, tdy = (g.b1y - g.b0y) / (float)M, tdz = (g.b1z - g.b0z) / (float)M;
  const float pdx = tdx / (float)THIN_T, pdy = tdy / (float)THIN_T, pdz = tdz / (float)THIN_T;
  cx = g.b0x + (x >> THIN_LOG_T) * tdx + (x & (THIN_T - 1)) * pdx + pdx / 2.0f;
  cy = g.b0y + (y >> THIN_LOG_T) * tdy + (y & (THIN_T - 1)) * pdy + pdy / 2.0f;
  cz = g.b0z + (z >> THIN_LOG_T) * tdz + (z & (THIN_T - 1)) * pdz + pdz / 2.0f;
}

#define THIN_EMPTY 0x7FFFFFFF

__global__ void __launch_bounds__(256) k_thin_fill(int* __restrict__ grid, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) grid[i] = THIN_EMPTY;
}

// occupied cells -> (sort key, cell); key = Morton key of the centre in the target cube (when given), ties by cell.
// keys == NULL: count only.
__global__ void __launch_bounds__(256)
k_thin_compact(const int* __restrict__ grid, long long total, ThinGrid g, TreeParams tp, int cell_bits, u64* __restrict__ keys,
               u32* __restrict__ cells, unsigned long long* count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool occ = i < total && grid[i] != THIN_EMPTY;
  const u32 bal = __ballot_sync(0xFFFFFFFFu, occ);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(bal));
  if (!keys) return;
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (occ) {
    u64 key = (u64)i;
    if (tp.D > 0) {
      const int N = g.N;
      float cx, cy, cz;
      thin_center(g, (int)(i % N), (int)((i / N) % N), (int)(i / ((long long)N * N)), cx, cy, cz);
      u64 k;
      osl_key(cx, cy, cz, tp, k);
      key = (k << cell_bits) | (u64)i;
    }
    const unsigned long long o = base + __popc(bal & ((1u << lane) - 1u));
    keys[o] = key;
    cells[o] = (u32)i;
  }
}

__global__ void __launch_bounds__(256)
k_thin_finish(const u32* __restrict__ cells, long long n, const int* __restrict__ grid, ThinGrid g,
              const float4* __restrict__ tri_colors, float4* __restrict__ centers, float4* __restrict__ colors,
              int* __restrict__ cells_out, int* __restrict__ tris_out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const u32 i = cells[j];
  const int N = g.N;
  const int x = (int)(i % N), y = (int)((i / N) % N), z = (int)(i / ((u32)N * N));
  const int t = grid[i];
  float cx, cy, cz;
  thin_center(g, x, y, z, cx, cy, cz);
  centers[j] = make_float4(cx, cy, cz, 1.0f);
  colors[j] = tri_colors ? tri_colors[t] : make_float4(1.f, 1.f, 1.f, 0.f);
  if (cells_out) { cells_out[3 * j] = x; cells_out[3 * j + 1] = y; cells_out[3 * j + 2] = z; }
  if (tris_out) tris_out[j] = t;
}

extern "C" osl_status osl_voxelize_thin(const float* d_vertices, int n_vertices, const int* d_triangles, int n_triangles,
                                        const float* d_tri_colors4, const float bbox0[3], const float bbox1[3], int log_n,
                                        const float cube_center[3], float cube_half, int cube_depth,
                                        float** d_centers4_out, float** d_colors4_out, int** d_cells_out, int** d_tris_out,
                                        int64_t* n_out, void* stream) {
  if (!n_out || !bbox0 || !bbox1 || log_n < THIN_LOG_T || log_n > 9 || n_triangles < 0 || n_vertices < 0 ||
      (n_triangles > 0 && (!d_vertices || !d_triangles)) || cube_depth < 0 || cube_depth > 12 ||
      (cube_depth > 0 && (!cube_center || !(cube_half > 0.0f))))
    return OSL_ERR_INVALID;
  for (int i = 0; i < 3; i++)
    if (!(bbox1[i] > bbox0[i])) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  *n_out = 0;
  if (d_centers4_out) *d_centers4_out = nullptr;
  if (d_colors4_out) *d_colors4_out = nullptr;
  if (d_cells_out) *d_cells_out = nullptr;
  if (d_tris_out) *d_tris_out = nullptr;
  if (n_triangles == 0) return OSL_OK;
  ThinGrid g;
  g.N = 1 << log_n; g.log_N = log_n;
  g.b0x = bbox0[0]; g.b0y = bbox0[1]; g.b0z = bbox0[2]; g.b1x = bbox1[0]; g.b1y = bbox1[1]; g.b1z = bbox1[2];
  g.dx = (g.b1x - g.b0x) / (float)g.N; g.dy = (g.b1y - g.b0y) / (float)g.N; g.dz = (g.b1z - g.b0z) / (float)g.N;
  g.ix = (float)g.N / (g.b1x - g.b0x); g.iy = (float)g.N / (g.b1y - g.b0y); g.iz = (float)g.N / (g.b1z - g.b0z);
  TreeParams tp;
  tp.cx = cube_depth ? cube_center[0] : 0.f; tp.cy = cube_depth ? cube_center[1] : 0.f; tp.cz = cube_depth ? cube_center[2] : 0.f;
  tp.half = cube_half; tp.D = cube_depth; tp.quirks = 1;
  const long long total = (long long)g.N * g.N * g.N;
  const int cell_bits = 3 * log_n;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  int* grid = nullptr;
  unsigned long long* d_count = nullptr;
  u64 *kA = nullptr, *kB = nullptr;
  u32 *pA = nullptr, *pB = nullptr;
  float4 *centers = nullptr, *colors = nullptr;
  int *cells_out = nullptr, *tris_out = nullptr;
  osl_status rc = OSL_OK;
  cudaError_t e = cudaSuccess;
  unsigned long long h_count = 0;
  long long n = 0;
  (void)n_vertices;
#define VT_CHECK(x) do { e = (x); if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; rc = (e == cudaErrorMemoryAllocation) ? OSL_ERR_OOM : OSL_ERR_CUDA; goto done; } } while (0)
  VT_CHECK(cudaMalloc(&grid, sizeof(int) * (size_t)total));
  VT_CHECK(cudaMalloc(&d_count, sizeof(unsigned long long)));
  VT_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
  k_thin_fill<<<blocks, 256, 0, st>>>(grid, total);
  k_thin_raster<<<(unsigned)(((long long)n_triangles * 32 + 255) / 256), 256, 0, st>>>(d_vertices, d_triangles, n_triangles, g, grid);
  k_thin_compact<<<blocks, 256, 0, st>>>(grid, total, g, tp, cell_bits, nullptr, nullptr, d_count);
  OSL_LAUNCHED(3);
  VT_CHECK(cudaMemcpyAsync(&h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, st));
  VT_CHECK(cudaStreamSynchronize(st));
  n = (long long)h_count;
  if (n == 0) goto done;
  VT_CHECK(cudaMalloc(&kA, sizeof(u64) * (size_t)n)); VT_CHECK(cudaMalloc(&kB, sizeof(u64) * (size_t)n));
  VT_CHECK(cudaMalloc(&pA, sizeof(u32) * (size_t)n)); VT_CHECK(cudaMalloc(&pB, sizeof(u32) * (size_t)n));
  VT_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
  k_thin_compact<<<blocks, 256, 0, st>>>(grid, total, g, tp, cell_bits, kA, pA, d_count);
  OSL_LAUNCHED(1);
  {
    int in_B = 0;
    rc = osl_device_sort_pairs(kA, pA, kB, pB, (int)n, 3 * cube_depth + cell_bits, st, &in_B);
    if (rc) goto done;
    const u32* sp = in_B ? pB : pA;
    VT_CHECK(cudaMalloc(&centers, sizeof(float4) * (size_t)n));
    VT_CHECK(cudaMalloc(&colors, sizeof(float4) * (size_t)n));
    if (d_cells_out) VT_CHECK(cudaMalloc(&cells_out, sizeof(int) * 3 * (size_t)n));
    if (d_tris_out) VT_CHECK(cudaMalloc(&tris_out, sizeof(int) * (size_t)n));
    k_thin_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sp, n, grid, g, reinterpret_cast<const float4*>(d_tri_colors4),
                                                               centers, colors, cells_out, tris_out);
    OSL_LAUNCHED(1);
    VT_CHECK(cudaStreamSynchronize(st));
  }
  *n_out = n;
  if (d_centers4_out) { *d_centers4_out = reinterpret_cast<float*>(centers); centers = nullptr; }
  if (d_colors4_out) { *d_colors4_out = reinterpret_cast<float*>(colors); colors = nullptr; }
  if (d_cells_out) { *d_cells_out = cells_out; cells_out = nullptr; }
  if (d_tris_out) { *d_tris_out = tris_out; tris_out = nullptr; }
done:
  cudaFree(grid); cudaFree(d_count); cudaFree(kA); cudaFree(kB); cudaFree(pA); cudaFree(pB);
  cudaFree(centers); cudaFree(colors); cudaFree(cells_out); cudaFree(tris_out);
  return rc;
#undef VT_CHECK
}
