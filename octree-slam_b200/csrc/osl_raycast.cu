// osl_raycast.cu -- octree raycast ("cone trace") of the SVO, one ray per thread, the whole march in registers.
//
// Replaces rendering::coneTraceSVO (cone_tracing_kernels.cu:157-198), createRays (:29-51) and coneTrace (:53-146).
// The reference launches one kernel + one stream compaction + one host sync PER MARCH STEP and round-trips the ray
// state through global memory; here a thread keeps its ray in registers until it terminates.
//
// Every float operation reproduces the shape nvcc 12.9 emits for the reference (SASS inspected, DESIGN.md
// "Float shapes"), so images are bit-identical to the reference's on the same GPU:
//   len   = sqrt_rn(fma(z,z, fma(x,x, y*y)))
//   lod   = (int)ceil(logf(size / (len*pix_scale)) / 0.693147182f)     (libdevice logf, IEEE divides)
//   step  = size / powf(2, lod);  ray *= (len + step) / len
// mode 0 (ref_exact) reproduces quirk Q8: the reference re-reads the (never updated) zeroed pixel every step, so the
// accumulator restarts from 0 each step and a pixel is the single sample taken at the terminating step.
// mode 1 (fixed_accumulate) keeps the accumulator in registers across steps (what the code was meant to do).
#include <math.h>

#include "osl_internal.cuh"

struct RayParams {
  float ox, oy, oz;        // camera origin
  float xdx, xdy, xdz;     // x_dir
  float ydx, ydy, ydz;     // y_dir
  float crx, cry, crz;     // cross(x_dir, -y_dir)
  float resx, resy;
  float pix_scale;
  float cx, cy, cz, size;  // SVO centre, half edge
  float fx, fy, start_dist, max_range;
  int mode;
  int W, H;
  int row0, rows;          // `rows` image rows are rendered into out[0 .. rows*W): local row r is image row
  int band_h, band_stride; // row0 + (r / band_h) * band_h * band_stride + r % band_h  (interleaved bands of a rank)
};

__device__ __forceinline__ float ray_length(float x, float y, float z) {
  return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y))));
}

// F2I.U32.TRUNC then byte store: NaN/negative -> 0, > 255 wraps mod 256 (Q16)
__device__ __forceinline__ u32 f2u8(float f) { return __float2uint_rz(f) & 0xFFu; }

// lod = (int)ceil(logf(q) / ln2f) with q = size / pix, the reference's expression (cone_tracing_kernels.cu:66-69).
// For a finite positive normal q = 1.m * 2^e whose mantissa is at least 2^-10 away from a power of two, log2(q) lies
// in (e + 1.4e-3, e + 1 - 7e-4) while the float evaluation is off by at most |log2 q| * 2e-7 < 3e-5, so the result
// is e + 1 and neither logf nor its division is needed.  The test is made on a FAST quotient (<= 2 ulp from the IEEE
// one) with a 2^-9 margin, which implies the 2^-10 margin for the exact quotient, so the IEEE division is skipped as
// well.  Everything else (0.4 % of the samples, non-finite / non-positive values) evaluates the reference's
// expression verbatim.
__device__ __forceinline__ int lod_depth(float size, float pix) {
  const float qa = __fdividef(size, pix);
  const u32 b = __float_as_uint(qa);
  const u32 ex = b >> 23;  // sign + exponent
  const u32 man = b & 0x7FFFFFu;
  if (ex >= 2u && ex <= 253u && man > 0x4000u && man < 0x7FC000u) return (int)ex - 126;
  const float q = __fdiv_rn(size, pix);
  return (int)ceilf(__fdiv_rn(logf(q), 0.693147182464599609375f));
}

// size / powf(2, depth) (cone_tracing_kernels.cu:126): powf(2, n) is exactly 2^n, and dividing by 2^n equals
// multiplying by 2^-n (both are the correctly rounded size * 2^-n)
__device__ __forceinline__ float node_step(float size, int depth) {
  if (depth >= -126 && depth <= 126) return __fmul_rn(size, __uint_as_float((u32)(127 - depth) << 23));
  return __fdiv_rn(size, powf(2.0f, (float)depth));
}

#define RAY_THREADS 128

// A cell of the tree a ray's descent went through, with the EXACT half-open bounds the root descent implies for it
// (the centres compared on the way down): a sample inside them takes the same branches from the root, so the descent
// may resume here.  A cached cell never becomes stale (the tree is read-only during a render).
struct RayCell {
  int lvl;             // >= 1; -1 = empty
  u32 child, self;     // first child of the cell's node, the node itself
  float cx, cy, cz, e; // centre and half size
  float lox, hix, loy, hiy, loz, hiz;
};

__device__ __forceinline__ bool cell_has(const RayCell& c, int depth, float tx, float ty, float tz) {
  return c.lvl >= 1 && depth >= c.lvl && tx > c.lox && !(tx > c.hix) && ty > c.loy && !(ty > c.hiy) && tz > c.loz &&
         !(tz > c.hiz);
}

// One ray per thread, the whole march in registers.  Per step the reference descends from the root to the LOD level;
// here a thread keeps TWO ancestor cells of its last sample -- a deep one 3 levels above the sample (hit by ~85 % of
// the next samples: consecutive samples are half a leaf apart) and a shallow one 7 levels above (hit by nearly all
// the others, which matters because one lane restarting at the root stalls its whole warp) -- and resumes the descent
// at the deepest one that contains the new sample.  Identical results by construction.
__global__ void __launch_bounds__(RAY_THREADS)
k_raycast(const u32* __restrict__ pool, RayParams P, uchar4* __restrict__ out, unsigned long long* stats) {
  __shared__ float s_af[256];  // (A - 127) / 127.0f for every alpha byte (Q9: no clamp)
  for (int a = threadIdx.x; a < 256; a += RAY_THREADS) s_af[a] = __fdiv_rn((float)(a - 127), 127.0f);
  __syncthreads();
  // a warp renders an 8x4-pixel patch (not 32 pixels of one row): neighbouring rays visit the same nodes and take
  // similar numbers of steps, so fewer lanes idle while the longest ray of the warp finishes
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (P.W + 7) >> 3;
  const int px = (gwarp % tiles_x) * 8 + (lane & 7);
  const int lr = (gwarp / tiles_x) * 4 + (lane >> 3);
  const int idx = lr * P.W + px;
  unsigned long long steps = 0, visits = 0;
  if (px < P.W && lr < P.rows) {
    const int py = P.row0 + (lr / P.band_h) * P.band_h * P.band_stride + lr % P.band_h;
    // createRays (cone_tracing_kernels.cu:29-51)
    const float magx = __fdiv_rn(__fmaf_rn(P.resx, -0.5f, (float)px), P.fx);
    const float magy = __fdiv_rn(__fmaf_rn(P.resy, -0.5f, (float)py), P.fy);
    const float dx = __fadd_rn(__fmaf_rn(magx, P.xdx, __fmul_rn(magy, P.ydx)), P.crx);
    const float dy = __fadd_rn(__fmaf_rn(magx, P.xdy, __fmul_rn(magy, P.ydy)), P.cry);
    const float dz = __fadd_rn(__fmaf_rn(magx, P.xdz, __fmul_rn(magy, P.ydz)), P.crz);
    const float dot = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const float inv = __frcp_rn(__fsqrt_rn(dot));
    float rx = __fmul_rn(__fmul_rn(dx, inv), P.start_dist);
    float ry = __fmul_rn(__fmul_rn(dy, inv), P.start_dist);
    float rz = __fmul_rn(__fmul_rn(dz, inv), P.start_dist);
    float len = ray_length(rx, ry, rz);

    const float INF = __int_as_float(0x7f800000);
    RayCell A, B;  // deep, shallow
    A.lvl = -1; B.lvl = -1;
    A.child = A.self = B.child = B.self = 0u;
    A.cx = A.cy = A.cz = A.e = B.cx = B.cy = B.cz = B.e = 0.f;
    A.lox = A.loy = A.loz = B.lox = B.loy = B.loz = -INF;
    A.hix = A.hiy = A.hiz = B.hix = B.hiy = B.hiz = INF;
    int last_lvl = 8;

    u32 vx = 0, vy = 0, vz = 0, vw = 0;  // uchar4 accumulator (mod-256 arithmetic)
    u32 result = 0;
    for (;;) {
      steps++;
      const float tx = __fadd_rn(P.ox, rx), ty = __fadd_rn(P.oy, ry), tz = __fadd_rn(P.oz, rz);
      const float pix = __fmul_rn(len, P.pix_scale);
      int depth = lod_depth(P.size, pix);

      // where the descent starts: deepest cached cell containing the sample, else the root
      u32 node = 0, child = 0;
      float cx = P.cx, cy = P.cy, cz = P.cz, e = P.size;
      float blx = -INF, bhx = INF, bly = -INF, bhy = INF, blz = -INF, bhz = INF;  // bounds of the current cell
      int i = 0;
      if (cell_has(A, depth, tx, ty, tz)) {
        i = A.lvl; node = A.self; child = A.child; cx = A.cx; cy = A.cy; cz = A.cz; e = A.e;
        blx = A.lox; bhx = A.hix; bly = A.loy; bhy = A.hiy; blz = A.loz; bhz = A.hiz;
      } else if (cell_has(B, depth, tx, ty, tz)) {
        i = B.lvl; node = B.self; child = B.child; cx = B.cx; cy = B.cy; cz = B.cz; e = B.e;
        blx = B.lox; bhx = B.hix; bly = B.loy; bhy = B.hiy; blz = B.loz; bhz = B.hiz;
      }
      visits += (unsigned long long)i;  // the word0 reads the root descent would have made down to here
      bool open = true;                 // the descent has not met a node without children

#define RAY_STEP_TRACKED()                                                         \
      {                                                                            \
        const bool bx = tx > cx, by = ty > cy, bz = tz > cz;                       \
        node = child + (u32)((int)bx + 2 * (int)by + 4 * (int)bz);                 \
        const u32 w0 = __ldg(pool + 2 * (size_t)node);                             \
        visits++;                                                                  \
        if (!(w0 & OSL_FLAG)) { depth = i + 1; open = false; break; }              \
        child = w0 & OSL_MASK;                                                     \
        if (bx) blx = fmaxf(blx, cx); else bhx = fminf(bhx, cx);                   \
        if (by) bly = fmaxf(bly, cy); else bhy = fminf(bhy, cy);                   \
        if (bz) blz = fmaxf(blz, cz); else bhz = fminf(bhz, cz);                   \
        e = __fmul_rn(e, 0.5f);                                                    \
        cx = __fadd_rn(cx, bx ? e : -e);                                           \
        cy = __fadd_rn(cy, by ? e : -e);                                           \
        cz = __fadd_rn(cz, bz ? e : -e);                                           \
        i++;                                                                       \
      }
#define RAY_SNAPSHOT(C, L)                                                         \
      {                                                                            \
        C.lvl = (L); C.self = node; C.child = child; C.cx = cx; C.cy = cy; C.cz = cz; C.e = e; \
        C.lox = blx; C.hix = bhx; C.loy = bly; C.hiy = bhy; C.loz = blz; C.hiz = bhz;          \
      }

      // tracked segments: down to the shallow target, snapshot, down to the deep target, snapshot
      const int ltB = max(last_lvl - 7, 1), ltA = max(last_lvl - 3, ltB + 1);
      if (i < ltB) {
        for (; i < depth && i < ltB;) RAY_STEP_TRACKED()
        if (open && i == ltB) RAY_SNAPSHOT(B, ltB)
      }
      if (open && i < ltA) {
        for (; i < depth && i < ltA;) RAY_STEP_TRACKED()
        if (open && i == ltA) RAY_SNAPSHOT(A, ltA)
      }
      if (open) {
        for (; i < depth;) {
          const bool bx = tx > cx, by = ty > cy, bz = tz > cz;
          node = child + (u32)((int)bx + 2 * (int)by + 4 * (int)bz);
          const u32 w0 = __ldg(pool + 2 * (size_t)node);
          visits++;
          if (!(w0 & OSL_FLAG)) { depth = i + 1; break; }
          child = w0 & OSL_MASK;
          e = __fmul_rn(e, 0.5f);
          cx = __fadd_rn(cx, bx ? e : -e);
          cy = __fadd_rn(cy, by ? e : -e);
          cz = __fadd_rn(cz, bz ? e : -e);
          i++;
        }
      }
#undef RAY_STEP_TRACKED
#undef RAY_SNAPSHOT
      last_lvl = depth;

      if (P.mode == 0) { vx = vy = vz = vw = 0; }  // Q8
      const u32 ov = __ldg(pool + 2 * (size_t)node + 1);
      const int alpha = (int)(ov >> 24) - 127;  // Q9: the reference's max(0, unsigned) is a no-op
      const float af = s_af[ov >> 24];
      vx = (vx + f2u8(__fmul_rn((float)(ov & 0xFFu), af))) & 0xFFu;
      vy = (vy + f2u8(__fmul_rn((float)((ov >> 8) & 0xFFu), af))) & 0xFFu;
      vz = (vz + f2u8(__fmul_rn((float)((ov >> 16) & 0xFFu), af))) & 0xFFu;
      if ((int)vw + alpha < 127) {
        vw = (vw + (u32)alpha) & 0xFFu;
      } else {
        result = vx | (vy << 8) | (vz << 16) | (255u << 24);
        break;
      }
      const float nd = node_step(P.size, depth);
      const float sc = __fdiv_rn(__fadd_rn(len, nd), len);
      rx = __fmul_rn(rx, sc); ry = __fmul_rn(ry, sc); rz = __fmul_rn(rz, sc);
      len = ray_length(rx, ry, rz);  // also the next step's |ray| (the reference recomputes the same value)
      if (len > P.max_range) {
        const float f = __fdiv_rn(127.0f, (float)vw);
        result = f2u8(__fmul_rn((float)vx, f)) | (f2u8(__fmul_rn((float)vy, f)) << 8) |
                 (f2u8(__fmul_rn((float)vz, f)) << 16) | (255u << 24);
        break;
      }
    }
    out[idx] = make_uchar4(result & 0xFF, (result >> 8) & 0xFF, (result >> 16) & 0xFF, result >> 24);
  }
  if (stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      steps += __shfl_xor_sync(0xFFFFFFFFu, steps, o);
      visits += __shfl_xor_sync(0xFFFFFFFFu, visits, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&stats[0], steps);
      atomicAdd(&stats[1], visits);
    }
  }
}

// glm 0.9.5.4 compute_inverse<tmat4x4> (glm/detail/type_mat4x4.inl:477-529) -- same operation order, host floats
static void mat4_inverse(const float a[16], float out[16]) {
#define M(c, r) a[4 * (c) + (r)]
  const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
  const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
  const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
  const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
  float inv[16];
  for (int k = 0; k < 4; k++) {
    inv[0 + k] = ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]) * sa[k];
    inv[4 + k] = ((v0[k] * f0[k] - v2[k] * f3[k]) + v3[k] * f4[k]) * sb[k];
    inv[8 + k] = ((v0[k] * f1[k] - v1[k] * f3[k]) + v3[k] * f5[k]) * sa[k];
    inv[12 + k] = ((v0[k] * f2[k] - v1[k] * f4[k]) + v2[k] * f5[k]) * sb[k];
  }
  const float dot1 = (M(0, 0) * inv[0] + M(0, 1) * inv[4]) + (M(0, 2) * inv[8] + M(0, 3) * inv[12]);
#undef M
  const float ood = 1.0f / dot1;
  for (int k = 0; k < 16; k++) out[k] = inv[k] * ood;
}

static void mat4_mul_vec4(const float m[16], const float v[4], float o[4]) {
  for (int r = 0; r < 4; r++) {
    const float a = m[0 + r] * v[0], b = m[4 + r] * v[1], c = m[8 + r] * v[2], d = m[12 + r] * v[3];
    o[r] = (a + b) + (c + d);
  }
}

osl_status osl_launch_raycast(const u32* d_pool, const float center[3], float half_edge, uint8_t* d_out, int w, int h,
                              int row0, int rows, int band_h, int band_stride, float fov_deg, const float view[16],
                              const osl_raycast_params* prm, unsigned long long* d_stats, cudaStream_t st) {
  if (!d_pool || !d_out || w <= 0 || h <= 0 || row0 < 0 || rows < 0 || band_h < 1 || band_stride < 1)
    return OSL_ERR_INVALID;
  if (rows == 0) return OSL_OK;
  {  // the last local row must lie inside the image
    const int lr = rows - 1;
    if (row0 + (lr / band_h) * band_h * band_stride + lr % band_h >= h) return OSL_ERR_INVALID;
  }
  osl_raycast_params p = {532.57f, 531.54f, 0.002f, 10.0f, 0};
  if (prm) p = *prm;
  // host part of coneTraceSVO (cone_tracing_kernels.cu:161-171)
  float inv[16], o4[4], xd[4], yd[4];
  mat4_inverse(view, inv);
  const float e_o[4] = {0, 0, 0, 1}, e_x[4] = {-1, 0, 0, 0}, e_y[4] = {0, -1, 0, 0};
  mat4_mul_vec4(inv, e_o, o4);
  mat4_mul_vec4(inv, e_x, xd);
  mat4_mul_vec4(inv, e_y, yd);
  RayParams P;
  P.ox = o4[0]; P.oy = o4[1]; P.oz = o4[2];
  P.xdx = xd[0]; P.xdy = xd[1]; P.xdz = xd[2];
  P.ydx = yd[0]; P.ydy = yd[1]; P.ydz = yd[2];
  // cross(x_dir, -y_dir) in the reference's contracted form: FFMA(b, c, -FMUL(d, e))
  P.crx = fmaf(yd[1], xd[2], -(xd[1] * yd[2]));
  P.cry = fmaf(xd[0], yd[2], -(yd[0] * xd[2]));
  P.crz = fmaf(yd[0], xd[1], -(yd[1] * xd[0]));
  P.resx = (float)w; P.resy = (float)h;
  P.pix_scale = tanf(fov_deg * 3.14159f / 180.0f) / (float)h;
  P.cx = center[0]; P.cy = center[1]; P.cz = center[2]; P.size = half_edge;
  P.fx = p.fx; P.fy = p.fy; P.start_dist = p.start_dist; P.max_range = p.max_range;
  P.mode = p.mode; P.W = w; P.H = h; P.row0 = row0; P.rows = rows; P.band_h = band_h; P.band_stride = band_stride;
  const long long warps = (long long)((w + 7) / 8) * ((rows + 3) / 4);  // one warp per 8x4-pixel patch
  const int wpb = RAY_THREADS / 32;
  k_raycast<<<(unsigned)((warps + wpb - 1) / wpb), RAY_THREADS, 0, st>>>(d_pool, P, reinterpret_cast<uchar4*>(d_out),
                                                                         d_stats);
  OSL_LAUNCHED(1);
  OSL_CUDA(cudaGetLastError());
  return OSL_OK;
}
