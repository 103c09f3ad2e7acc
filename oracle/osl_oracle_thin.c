/* osl_oracle_thin.c -- CPU restatement of the reference's mesh voxelisation RULE.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference voxelises with a third-party library it vendors under external/include/voxelpipe (version unknown;
 * "voxelpipe" by J. Pantaleoni, NVIDIA, 2011 -- the library does not build with CUDA 12):
 *   voxelization.cu:24,281-285   context->fine_raster<Float, FP32S_FORMAT, THIN_RASTER, NO_BLENDING, ColorShader>
 *                                on a dense N^3 grid, N = 2^GRID_RES = 256, tiles of 8^3, over the MESH bounding box
 * What follows restates what that call computes, from the library's own source as vendored in the reference tree:
 *   coarse.h:48-103     setup_triangle: integer bounding box (clamped to the grid) and dominant axis
 *   coarse.h:646-690    a triangle is handed to every tile its integer bounding box touches
 *   fine.h:1219-1330    per (tile, triangle): clamp the box to the tile, plane/tile-box test, dispatch on the axis
 *   fine.h:368-540      rasterize<AXIS>: for every scanline v, the u range from the three 2-D edge functions offset
 *                       towards the pixel corner (2-D CONSERVATIVE coverage; utils.h:185-231 triangle_setup,
 *                       fine.h:130-152 compute_scanline_bounds), then ONE voxel per column: w = int(depth of the
 *                       plane at the pixel CENTRE * inv_delta_w) (utils.h:236-253 plane_setup), kept when it lies in
 *                       the tile
 *   utils.h:115-180     uvw<AXIS>: axis 0 -> (u,v,w) = (y,z,x), 1 -> (x,z,y), 2 -> (x,y,z); ccw signs
 * NO_BLENDING lets the last triangle that reaches a voxel win (a race on the GPU); the canon here, as everywhere in
 * this repository where the reference races, is the LOWEST triangle index.
 * Float operations are written one per line in the source's order and compiled without contraction.  The real library
 * was compiled by nvcc with FMA contraction on, which can move a pixel exactly on an edge to the other side; since
 * the library cannot be built, that last bit is not pinned ("parity pinned to the restated rule").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } f3;

static inline float sel_u(int axis, f3 a) { return axis == 0 ? a.y : a.x; }
static inline float sel_v(int axis, f3 a) { return axis == 2 ? a.y : a.z; }
static inline float sel_w(int axis, f3 a) { return axis == 0 ? a.x : (axis == 1 ? a.y : a.z); }
static inline int isel_u(int axis, const int a[3]) { return axis == 0 ? a[1] : a[0]; }
static inline int isel_v(int axis, const int a[3]) { return axis == 2 ? a[1] : a[2]; }
static inline int isel_w(int axis, const int a[3]) { return axis == 0 ? a[0] : (axis == 1 ? a[1] : a[2]); }
static inline float ccw(int axis, f3 n) {
  if (axis == 0) return n.x > 0.0f ? 1.0f : -1.0f;
  if (axis == 1) return n.y < 0.0f ? 1.0f : -1.0f;
  return n.z > 0.0f ? 1.0f : -1.0f;
}
/* float -> int as the GPU converts (F2I.TRUNC saturates: NaN -> 0, out of range -> INT_MIN / INT_MAX); a host cast is
 * undefined there, and degenerate triangles (zero-area: n = 0, 1/n = inf) do produce such values */
static inline int f2i(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int)f;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* fine.h:130-152 */
static void scanline_bounds(const float b[3], const float ndu[3], const float inv[3], int* min_u, int* max_u) {
  for (int i = 0; i < 3; i++) {
    if (ndu[i] > 0.0f) *min_u = imax(*min_u, f2i(ceilf(-b[i] * inv[i])));
    else if (ndu[i] < 0.0f) *max_u = imin(*max_u, f2i(-b[i] * inv[i]));
    else if (b[i] < 0.0f) *min_u = *max_u + 1;
  }
}

/* Returns the number of occupied voxels; cells (x, y, z per voxel) and tris (lowest triangle index) are filled in
 * ascending order of z*N*N + y*N + x up to `cap` entries. */
long long orc_voxelize_thin(const float* verts, int n_verts, const int* tris_in, int n_tris, const float bbox0_[3],
                            const float bbox1_[3], int log_N, int* cells, int* tris, long long cap) {
  (void)n_verts;
  const int N = 1 << log_N, LOG_T = 3, T = 1 << LOG_T;
  const f3 bbox0 = {bbox0_[0], bbox0_[1], bbox0_[2]}, bbox1 = {bbox1_[0], bbox1_[1], bbox1_[2]};
  /* voxelpipe_inline.h:111-118 */
  const f3 delta = {(bbox1.x - bbox0.x) / (float)N, (bbox1.y - bbox0.y) / (float)N, (bbox1.z - bbox0.z) / (float)N};
  const f3 inv_delta = {(float)N / (bbox1.x - bbox0.x), (float)N / (bbox1.y - bbox0.y), (float)N / (bbox1.z - bbox0.z)};
  const size_t total = (size_t)N * N * N;
  int* grid = (int*)malloc(total * sizeof(int));
  if (!grid) return -1;
  memset(grid, 0xFF, total * sizeof(int)); /* -1 = empty */

  for (int t = 0; t < n_tris; t++) {
    const float* p0 = verts + 3 * (size_t)tris_in[3 * t], *p1 = verts + 3 * (size_t)tris_in[3 * t + 1],
                *p2 = verts + 3 * (size_t)tris_in[3 * t + 2];
    const f3 v0 = {p0[0], p0[1], p0[2]}, v1 = {p1[0], p1[1], p1[2]}, v2 = {p2[0], p2[1], p2[2]};
    /* coarse.h:48-103 setup_triangle */
    float lo[3], hi[3];
    {
      const float a0[3] = {(v0.x - bbox0.x) * inv_delta.x, (v0.y - bbox0.y) * inv_delta.y, (v0.z - bbox0.z) * inv_delta.z};
      const float a1[3] = {(v1.x - bbox0.x) * inv_delta.x, (v1.y - bbox0.y) * inv_delta.y, (v1.z - bbox0.z) * inv_delta.z};
      const float a2[3] = {(v2.x - bbox0.x) * inv_delta.x, (v2.y - bbox0.y) * inv_delta.y, (v2.z - bbox0.z) * inv_delta.z};
      for (int i = 0; i < 3; i++) {
        lo[i] = fminf(a2[i], fminf(a1[i], a0[i]));
        hi[i] = fmaxf(a2[i], fmaxf(a1[i], a0[i]));
      }
    }
    int b0[3], b1[3];
    for (int i = 0; i < 3; i++) {
      b0[i] = imin(imax(f2i(lo[i]), 0), N - 1);
      b1[i] = imin(imax(f2i(ceilf(hi[i])), 0), N - 1);
    }
    const f3 e0 = {v1.x - v0.x, v1.y - v0.y, v1.z - v0.z};
    const f3 e1 = {v2.x - v1.x, v2.y - v1.y, v2.z - v1.z};
    const f3 e2 = {v0.x - v2.x, v0.y - v2.y, v0.z - v2.z};
    /* utils.h:97-103 anti_cross(edge0, edge2) */
    const f3 n = {e0.z * e2.y - e0.y * e2.z, e0.x * e2.z - e0.z * e2.x, e0.y * e2.x - e0.x * e2.y};
    const int byx = fabsf(n.y) > fabsf(n.x), byz = fabsf(n.y) > fabsf(n.z), bzx = fabsf(n.z) > fabsf(n.x);
    const int axis = byx ? (byz ? 1 : 2) : (bzx ? 2 : 0);

    /* utils.h:185-231 triangle_setup<AXIS> */
    const float sgn = ccw(axis, n);
    const f3 edges[3] = {e0, e1, e2}, vs[3] = {v0, v1, v2};
    float a[3], ndu[3], ndv[3], inv_du[3];
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
      inv_du[i] = 1.0f / ndu[i];
    }
    /* utils.h:236-253 plane_setup<AXIS> */
    const float inv_n = 1.0f / sel_w(axis, n);
    const float pex = sel_u(axis, n) * inv_n, pey = sel_v(axis, n) * inv_n;
    float pez = pex * sel_u(axis, v0);
    pez = pez + pey * sel_v(axis, v0);
    pez = pez + sel_w(axis, v0);
    pez = pez - sel_w(axis, bbox0);
    pez = pez - pex * sel_u(axis, bbox0);
    pez = pez - pey * sel_v(axis, bbox0);

    /* coarse.h:646-690: every tile the integer box touches; fine.h:1219-1330 per tile */
    for (int tz = b0[2] >> LOG_T; tz <= b1[2] >> LOG_T; tz++)
      for (int ty = b0[1] >> LOG_T; ty <= b1[1] >> LOG_T; ty++)
        for (int tx = b0[0] >> LOG_T; tx <= b1[0] >> LOG_T; tx++) {
          const int tile[3] = {tx * T, ty * T, tz * T};
          int c0[3], c1[3];
          for (int i = 0; i < 3; i++) { c0[i] = imax(b0[i], tile[i]); c1[i] = imin(b1[i], tile[i] + T - 1); }
          {
            /* plane / tile-box test */
            const float cx = n.x > 0 ? delta.x * T : 0.0f, cy = n.y > 0 ? delta.y * T : 0.0f, cz = n.z > 0 ? delta.z * T : 0.0f;
            float r1 = n.x * (cx - v0.x);
            r1 = r1 + n.y * (cy - v0.y);
            r1 = r1 + n.z * (cz - v0.z);
            float r2 = n.x * (delta.x * T - cx - v0.x);
            r2 = r2 + n.y * (delta.y * T - cy - v0.y);
            r2 = r2 + n.z * (delta.z * T - cz - v0.z);
            float np = n.x * (bbox0.x + tile[0] * delta.x);
            np = np + n.y * (bbox0.y + tile[1] * delta.y);
            np = np + n.z * (bbox0.z + tile[2] * delta.z);
            if ((np + r1) * (np + r2) > 0.0f) continue;
          }
          /* fine.h:368-540 rasterize<AXIS> */
          for (int v = isel_v(axis, c0); v <= isel_v(axis, c1); v++) {
            const float b[3] = {a[0] + (float)v * ndv[0], a[1] + (float)v * ndv[1], a[2] + (float)v * ndv[2]};
            int min_u = isel_u(axis, c0), max_u = isel_u(axis, c1);
            scanline_bounds(b, ndu, inv_du, &min_u, &max_u);
            for (int u = min_u; u <= max_u; u++) {
              const float uf = ((float)u + 0.5f) * sel_u(axis, delta);
              const float vf = ((float)v + 0.5f) * sel_v(axis, delta);
              const float q1 = pex * uf;
              const float q2 = pey * vf;
              const float wf = pez - (q1 + q2);
              const int w = f2i(wf * sel_w(axis, inv_delta));
              if (w >= isel_w(axis, tile) && w < isel_w(axis, tile) + T) {
                int x, y, z;
                if (axis == 0) { x = w; y = u; z = v; }
                else if (axis == 1) { x = u; y = w; z = v; }
                else { x = u; y = v; z = w; }
                int* cell = &grid[((size_t)z * N + y) * N + x];
                if (*cell < 0 || t < *cell) *cell = t;
              }
            }
          }
        }
  }
  long long count = 0;
  for (size_t i = 0; i < total; i++)
    if (grid[i] >= 0) {
      if (count < cap) {
        cells[3 * count] = (int)(i % N); cells[3 * count + 1] = (int)((i / N) % N); cells[3 * count + 2] = (int)(i / ((size_t)N * N));
        tris[count] = grid[i];
      }
      count++;
    }
  free(grid);
  return count;
}
