// osl_track.cu -- camera tracking: the step before integration in the SLAM loop.
//
// Replaces  sensor::bilateralFilter / subsampleDepth / generateNormalMap / transformNormalMap / colorToIntensity /
//           subsample            image_kernels.cu:104-321
//           sensor::computeICPCost2                     localization_kernels.cu:155-231, 313-330
//           sensor::RGBDCamera::update / solveCholesky   rgbd_camera.cpp:53-224
//
// Design (not the reference's): one frame of tracking is 27 asynchronous launches on one stream and NO host round
// trip.  The reference synchronises after every kernel, copies 42 floats to the host 19 times per frame, solves the
// 6x6 system there and materialises a transformed copy of the vertex / normal maps between iterations.  Here each
// ICP iteration is ONE kernel: it applies the previous iteration's increment to the working maps on the fly,
// accumulates the 27 distinct normal-equation terms per thread, reduces them (shuffles -> shared -> one partial per
// CTA), and the last CTA to finish (ticket) folds the partials in a fixed order, solves the system (Cholesky, the
// reference's float storage / double sums), builds the pose increment and composes it into the running update --
// all in device memory.  The pose is read back once per frame.
#include "osl_internal.cuh"

#define TRK_LEVELS 3
#define TRK_THREADS 256
#define TRK_TERMS 27  // 21 upper-triangle terms of A + 6 of b
#define TRK_MAX_CTAS 1024

struct TrackState {
  float update[16];     // update_trans (rgbd_camera.cpp:100)
  float inc[16];        // this_trans of the last solved iteration
  float world[16];      // exact mode: camera-to-world pose
  float position[3];    // RGBDCamera::position_
  float orientation[9]; // RGBDCamera::orientation_ (column-major mat3)
  float pose[16];       // what main.cpp:40 applies to the vertex map: mat4(orientation) * translate(position)
  float A[36], b[6], x[6];
  int level_lost[TRK_LEVELS];
  int lost;
  int pairs;            // correspondences of the last iteration
  unsigned ticket;
  int frames;
};

struct osl_tracker {
  int device;
  int w, h;
  float fx, fy;
  int flags;  // bit 0: exact Jacobian + consistent increment (not the reference)
  int pass;   // RGBDCamera::pass_ (caps at 2)
  int last;   // which pyramid set holds the last frame
  float *vtx[2][TRK_LEVELS], *nrm[2][TRK_LEVELS];
  float *work_v, *work_n;
  uint16_t *filt, *tmp, *stage;
  float* partials;  // [CTAs][TRK_TERMS]
  TrackState* d_state;
  TrackState* h_state;  // pinned
  cudaEvent_t done;
  int num_sms;
  bool pending;
};

static const int TRK_ITERS[TRK_LEVELS] = {10, 5, 4};  // rgbd_camera.cpp:19

// ------------------------------------------------------------------------------------------------ image kernels

// image_kernels.cu:137-166.  Float shapes from the reference's SASS: t = FMUL(color2, sig_dep);
// t = FFMA(space2, sig_spat, t); arg = FMUL(t, -log2e); below -126 the argument is halved and the result squared
// (__expf's range handling); MUFU.EX2; sum1 = FFMA(e, depth, sum1); sum2 = FADD(e, sum2); F2I.RN of the IEEE quotient.
// The squared differences are converted as UNSIGNED integers (dims is uint2 in the reference).
#define BIL_TX 32
#define BIL_TY 8
#define BIL_R 3
__global__ void __launch_bounds__(BIL_TX* BIL_TY) k_bilateral(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                               int w, int h, float sig_spat, float sig_dep) {
  __shared__ uint16_t tile[BIL_TY + 2 * BIL_R][BIL_TX + 2 * BIL_R + 2];
  const int x0 = blockIdx.x * BIL_TX, y0 = blockIdx.y * BIL_TY;
  for (int i = threadIdx.y * BIL_TX + threadIdx.x; i < (BIL_TY + 2 * BIL_R) * (BIL_TX + 2 * BIL_R);
       i += BIL_TX * BIL_TY) {
    const int ty = i / (BIL_TX + 2 * BIL_R), tx = i % (BIL_TX + 2 * BIL_R);
    const int gx = x0 + tx - BIL_R, gy = y0 + ty - BIL_R;
    tile[ty][tx] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? in[(size_t)gy * w + gx] : (uint16_t)0;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= w || y >= h) return;
  const int value = tile[threadIdx.y + BIL_R][threadIdx.x + BIL_R];
  const int tx_end = min(x - BIL_R + 7, w - 1), ty_end = min(y - BIL_R + 7, h - 1);  // exclusive (reference bounds)
  float sum1 = 0.0f, sum2 = 0.0f;
  for (int cy = max(y - BIL_R, 0); cy < ty_end; cy++) {
    for (int cx = max(x - BIL_R, 0); cx < tx_end; cx++) {
      const int depth = tile[cy - y0 + BIL_R][cx - x0 + BIL_R];
      const float space2 = (float)(u32)((x - cx) * (x - cx) + (y - cy) * (y - cy));
      const u32 dd = (u32)(value - depth);
      const float color2 = (float)(dd * dd);
      float t = __fmul_rn(color2, sig_dep);
      t = __fmaf_rn(space2, sig_spat, t);
      t = __fmul_rn(t, -1.4426950216293334961f);
      const bool small = t < -126.0f;
      if (small) t = __fmul_rn(t, 0.5f);
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
      if (small) e = __fmul_rn(e, e);
      sum1 = __fmaf_rn(e, (float)depth, sum1);
      sum2 = __fadd_rn(e, sum2);
    }
  }
  out[(size_t)y * w + x] = (uint16_t)__float2int_rn(__fdiv_rn(sum1, sum2));
}

// image_kernels.cu:228-260: (width, height) are the OUTPUT dimensions; `in` is (2*width) wide
__global__ void k_subsample_depth(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int width, int height,
                                  float sigma) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width * height) return;
  const int x = idx % width, y = idx / width;
  const float center = (float)in[4 * y * width + 2 * x];
  const int tx = min(2 * x - 2 + 5, 2 * width - 1), ty = min(2 * y - 2 + 5, 2 * height - 1);
  float sum = 0.0f, count = 0.0f;
  for (int cy = max(0, 2 * y - 2); cy < ty; cy++)
    for (int cx = max(0, 2 * x - 2); cx < tx; cx++) {
      const float val = (float)in[2 * cy * width + cx];
      if (fabsf(__fsub_rn(val, center)) < sigma) {
        sum = __fadd_rn(sum, val);
        count = __fadd_rn(count, 1.0f);
      }
    }
  const float r = count == 0.0f ? 0.0f : __fdiv_rn(sum, count);
  out[idx] = (uint16_t)__float2uint_rz(r);
}

// image_kernels.cu:104-129: cross = FFMA(a, b, -FMUL(c, d)) per component, dot = FFMA(cz,cz,FFMA(cx,cx,FMUL(cy,cy))),
// normal = cross * -(1 / sqrt(dot)) with IEEE sqrt and reciprocal
__device__ __forceinline__ void trk_normal(float cX, float cY, float cZ, float rX, float rY, float rZ, float bX,
                                           float bY, float bZ, float& nx, float& ny, float& nz) {
  const float v1x = __fsub_rn(rX, cX), v1y = __fsub_rn(rY, cY), v1z = __fsub_rn(rZ, cZ);
  const float v2x = __fsub_rn(bX, cX), v2y = __fsub_rn(bY, cY), v2z = __fsub_rn(bZ, cZ);
  const float cx = __fmaf_rn(v1y, v2z, -__fmul_rn(v1z, v2y));
  const float cy = __fmaf_rn(v1z, v2x, -__fmul_rn(v1x, v2z));
  const float cz = __fmaf_rn(v1x, v2y, -__fmul_rn(v1y, v2x));
  const float dot = __fmaf_rn(cz, cz, __fmaf_rn(cx, cx, __fmul_rn(cy, cy)));
  const float inv = -__frcp_rn(__fsqrt_rn(dot));
  nx = __fmul_rn(cx, inv); ny = __fmul_rn(cy, inv); nz = __fmul_rn(cz, inv);
}

__global__ void k_normal_map(const float* __restrict__ vtx, float* __restrict__ nrm, int w, int h) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * h) return;
  const int x = idx % w, y = idx / w;
  float nx, ny, nz;
  if (x == w - 1 || y == h - 1) {
    nx = ny = nz = __int_as_float(0x7f800000);
  } else {
    const float *c = vtx + 3 * (size_t)idx, *r = c + 3, *b = c + 3 * (size_t)w;
    trk_normal(c[0], c[1], c[2], r[0], r[1], r[2], b[0], b[1], b[2], nx, ny, nz);
  }
  nrm[3 * (size_t)idx] = nx; nrm[3 * (size_t)idx + 1] = ny; nrm[3 * (size_t)idx + 2] = nz;
}

// generateVertexMap + generateNormalMap of one pyramid level in one pass over the depth image (three vertices are
// rebuilt per pixel from three depth loads instead of re-reading a 12-byte vertex map)
__global__ void k_vertex_normal(const uint16_t* __restrict__ depth, float* __restrict__ vtx, float* __restrict__ nrm,
                                int w, int h, float fx, float fy, int img_w, int img_h) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * h) return;
  const int x = idx % w, y = idx / w;
  float cX, cY, cZ, nx, ny, nz;
  osl_vertex(depth[idx], x, y, w, h, img_w, img_h, fx, fy, cX, cY, cZ);
  if (x == w - 1 || y == h - 1) {
    nx = ny = nz = __int_as_float(0x7f800000);
  } else {
    float rX, rY, rZ, bX, bY, bZ;
    osl_vertex(depth[idx + 1], x + 1, y, w, h, img_w, img_h, fx, fy, rX, rY, rZ);
    osl_vertex(depth[idx + w], x, y + 1, w, h, img_w, img_h, fx, fy, bX, bY, bZ);
    trk_normal(cX, cY, cZ, rX, rY, rZ, bX, bY, bZ, nx, ny, nz);
  }
  vtx[3 * (size_t)idx] = cX; vtx[3 * (size_t)idx + 1] = cY; vtx[3 * (size_t)idx + 2] = cZ;
  nrm[3 * (size_t)idx] = nx; nrm[3 * (size_t)idx + 1] = ny; nrm[3 * (size_t)idx + 2] = nz;
}

// image_kernels.cu:217-226: trans * vec4(n, 0)
__device__ __forceinline__ void trk_rotate(const float* __restrict__ M, float& x, float& y, float& z) {
  float o[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    float t = __fmul_rn(y, M[4 + r]);
    t = __fmaf_rn(x, M[0 + r], t);
    const float u = __fmaf_rn(z, M[8 + r], __fmul_rn(M[12 + r], 0.0f));
    o[r] = __fadd_rn(t, u);
  }
  x = o[0]; y = o[1]; z = o[2];
}

struct Mat16 { float m[16]; };

__global__ void k_transform_normals(float* __restrict__ nrm, Mat16 M, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float x = nrm[3 * (size_t)idx], y = nrm[3 * (size_t)idx + 1], z = nrm[3 * (size_t)idx + 2];
  trk_rotate(M.m, x, y, z);
  nrm[3 * (size_t)idx] = x; nrm[3 * (size_t)idx + 1] = y; nrm[3 * (size_t)idx + 2] = z;
}

// image_kernels.cu:178-186 (r, b, b: the green channel is never read)
__global__ void k_color_to_intensity(const uint8_t* __restrict__ rgb, float* __restrict__ out, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float r = __fdiv_rn((float)rgb[3 * (size_t)idx], 255.0f), b = __fdiv_rn((float)rgb[3 * (size_t)idx + 2], 255.0f);
  out[idx] = __fmaf_rn(b, 0.114f, __fmaf_rn(r, 0.299f, __fmul_rn(b, 0.587f)));  // the reference's SASS shape
}

// image_kernels.cu:285-297: (width, height) are the OUTPUT dimensions
__global__ void k_subsample_f32(const float* __restrict__ in, float* __restrict__ out, int width, int height) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width * height) return;
  const int x = idx % width, y = idx / width;
  out[idx] = in[4 * y * width + 2 * x];
}

// ------------------------------------------------------------------------------------------------ pose algebra

__device__ __forceinline__ void trk_identity(float* m) {
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

// out = a * b, column-major; out may alias.  Unrolled: the matrices live in registers.
__device__ __forceinline__ void trk_mul(const float* a, const float* b, float* out) {
  float r[16];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int k = 0; k < 4; k++)
      r[4 * c + k] = a[k] * b[4 * c] + a[4 + k] * b[4 * c + 1] + a[8 + k] * b[4 * c + 2] + a[12 + k] * b[4 * c + 3];
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = r[i];
}

// glm::rotate(mat4(1), degrees, unit axis) (gtc/matrix_transform.inl:48-86)
__device__ __forceinline__ void trk_rotate_deg(float angle_deg, float ax, float ay, float az, float* out) {
  const float a = angle_deg * 0.01745329251994329576923690768489f;
  const float c = cosf(a), s = sinf(a);
  const float t0 = (1.0f - c) * ax, t1 = (1.0f - c) * ay, t2 = (1.0f - c) * az;
  trk_identity(out);
  out[0] = c + t0 * ax;      out[1] = t0 * ay + s * az; out[2] = t0 * az - s * ay;
  out[4] = t1 * ax - s * az; out[5] = c + t1 * ay;      out[6] = t1 * az + s * ax;
  out[8] = t2 * ax + s * ay; out[9] = t2 * ay - s * ax; out[10] = c + t2 * az;
}

// rgbd_camera.cpp:153-158: Rz(-x2) * Ry(-x1) * Rx(-x0) * T(x3, x4, x5), angles through 180 / 3.14159f.
// exact: T * Rz(x2) * Ry(x1) * Rx(x0) -- the increment the linearised residual n . (v + w x v + t - v1) solves for.
__device__ __forceinline__ void trk_increment(const float* x, bool exact, float* out) {
  float rz[16], ry[16], rx[16], t[16], m[16];
  const float sg = exact ? 1.0f : -1.0f;
  trk_rotate_deg(sg * x[2] * 180.0f / 3.14159f, 0.0f, 0.0f, 1.0f, rz);
  trk_rotate_deg(sg * x[1] * 180.0f / 3.14159f, 0.0f, 1.0f, 0.0f, ry);
  trk_rotate_deg(sg * x[0] * 180.0f / 3.14159f, 1.0f, 0.0f, 0.0f, rx);
  trk_identity(t);
  t[12] = x[3]; t[13] = x[4]; t[14] = x[5];
  trk_mul(rz, ry, m);
  trk_mul(m, rx, m);
  if (exact) trk_mul(t, m, out);
  else trk_mul(m, t, out);
}

// rgbd_camera.cpp:193-224: float storage, double sums.  Every loop has constant bounds and is unrolled so that the
// factor stays in registers (this runs on ONE thread at the end of every ICP iteration: its latency is serial).
__device__ __forceinline__ void trk_cholesky(const float* A, const float* b, float* x) {
  float LU[36], yv[6];
#pragma unroll
  for (int i = 0; i < 36; i++) LU[i] = 0.0f;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double sum = 0.0;
#pragma unroll
    for (int p = 0; p < k; p++) sum += LU[k * 6 + p] * LU[k * 6 + p];
    LU[k * 6 + k] = (float)sqrt(A[k * 6 + k] - sum);
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      double s2 = 0.0;
#pragma unroll
      for (int p = 0; p < k; p++) s2 += LU[i * 6 + p] * LU[k * 6 + p];
      LU[i * 6 + k] = (float)((A[i * 6 + k] - s2) / LU[k * 6 + k]);
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < i; k++) sum += LU[i * 6 + k] * yv[k];
    yv[i] = (float)((b[i] - sum) / LU[i * 6 + i]);
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double sum = 0.0;
#pragma unroll
    for (int k = i + 1; k < 6; k++) sum += LU[k * 6 + i] * x[k];
    x[i] = (float)((yv[i] - sum) / LU[i * 6 + i]);
  }
}

// ------------------------------------------------------------------------------------------------ ICP iteration

// One Gauss-Newton iteration (rgbd_camera.cpp:123-168 with computeICPCost2, localization_kernels.cu:155-231).
//   apply = 0: the working maps are the sources as they are
//   apply = 1: source maps transformed by state->update  (first iteration of a finer level, rgbd_camera.cpp:116-120)
//   apply = 2: source maps transformed by state->inc     (rgbd_camera.cpp:163-167 of the previous iteration)
// The transformed maps are written to dst (may alias src: every thread rewrites only its own points).
// solve = 0: only A, b and the pair count are produced (osl_icp_cost).
__global__ void __launch_bounds__(TRK_THREADS) k_icp_step(const float* __restrict__ last_v, const float* __restrict__ last_n,
                                                         const float* src_v, const float* src_n, float* dst_v,
                                                         float* dst_n, int n, int apply, int level, int solve,
                                                         int exact, TrackState* st, float* __restrict__ partials) {
  if (solve && st->level_lost[level]) return;  // `break` of an earlier iteration of this level (uniform)
  __shared__ float s_M[16];
  __shared__ float s_red[TRK_THREADS / 32][TRK_TERMS + 1];
  __shared__ bool s_last;
  if (threadIdx.x < 16 && apply) s_M[threadIdx.x] = apply == 1 ? st->update[threadIdx.x] : st->inc[threadIdx.x];
  __syncthreads();

  float acc[TRK_TERMS];
#pragma unroll
  for (int i = 0; i < TRK_TERMS; i++) acc[i] = 0.0f;
  int pairs = 0;
  // software pipeline: the 12 values of the thread's next point are in flight while the current one is processed
  // (a thread only ever rewrites its own points, so reading ahead of the stores is safe)
  const int stride = gridDim.x * TRK_THREADS;
  float nx[12];
  {
    const int i0 = blockIdx.x * TRK_THREADS + threadIdx.x;
    if (i0 < n) {
      const size_t o = 3 * (size_t)i0;
#pragma unroll
      for (int k = 0; k < 3; k++) { nx[k] = src_v[o + k]; nx[3 + k] = src_n[o + k]; nx[6 + k] = last_v[o + k]; nx[9 + k] = last_n[o + k]; }
    }
  }
  for (int i = blockIdx.x * TRK_THREADS + threadIdx.x; i < n; i += stride) {
    const size_t o = 3 * (size_t)i;
    float v2x = nx[0], v2y = nx[1], v2z = nx[2];
    float n2x = nx[3], n2y = nx[4], n2z = nx[5];
    const float v1x = nx[6], v1y = nx[7], v1z = nx[8];
    const float n1x = nx[9], n1y = nx[10], n1z = nx[11];
    if (i + stride < n) {
      const size_t o2 = 3 * (size_t)(i + stride);
#pragma unroll
      for (int k = 0; k < 3; k++) { nx[k] = src_v[o2 + k]; nx[3 + k] = src_n[o2 + k]; nx[6 + k] = last_v[o2 + k]; nx[9 + k] = last_n[o2 + k]; }
    }
    if (apply) {
      osl_transform(s_M, v2x, v2y, v2z);
      trk_rotate(s_M, n2x, n2y, n2z);
    }
    if (apply || dst_v != src_v) {
      dst_v[o] = v2x; dst_v[o + 1] = v2y; dst_v[o + 2] = v2z;
      dst_n[o] = n2x; dst_n[o + 1] = n2y; dst_n[o + 2] = n2z;
    }
    bool ok = isfinite(v2x) && isfinite(v2y) && isfinite(v2z) && isfinite(v1x) && isfinite(v1y) && isfinite(v1z) &&
              !(v1z < 0.1f) && !(v2z < 0.1f) && !(v1z > 10.0f) && !(v2z > 10.0f);
    ok = ok && isfinite(n2x) && isfinite(n2y) && isfinite(n2z) && isfinite(n1x) && isfinite(n1y) && isfinite(n1z);
    if (!ok) continue;
    const float dx = v2x - v1x, dy = v2y - v1y, dz = v2z - v1z;
    if (sqrtf(dx * dx + dy * dy + dz * dz) > 0.1f) continue;   // DIST_THRESH
    if (n2x * n1x + n2y * n1y + n2z * n1z < 0.87f) continue;    // NORM_THRESH
    float at[6];
    if (exact) {  // v2 x n1
      at[0] = v2y * n1z - v2z * n1y;
      at[1] = v2z * n1x - v2x * n1z;
      at[2] = v2x * n1y - v2y * n1x;
    } else {      // the reference's G^T (quirk Q17)
      at[0] = -v2x * n1y - v2y * n1z;
      at[1] = -v2z * n1x + v2x * n1z;
      at[2] = v2y * n1x + v2z * n1y;
    }
    at[3] = n1x; at[4] = n1y; at[5] = n1z;
    const float bb = n1x * (v1x - v2x) + n1y * (v1y - v2y) + n1z * (v1z - v2z);
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = r; c < 6; c++) acc[k++] += at[r] * at[c];
#pragma unroll
    for (int r = 0; r < 6; r++) acc[21 + r] += bb * at[r];
    pairs++;
  }

  // CTA reduction: shuffles, then one row per warp in shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < TRK_TERMS; i++) {
    float v = acc[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) s_red[warp][i] = v;
  }
  pairs = __reduce_add_sync(0xffffffffu, pairs);
  if (lane == 0) s_red[warp][TRK_TERMS] = __int_as_float(pairs);
  __syncthreads();
  if (threadIdx.x <= TRK_TERMS) {
    float* p = partials + (size_t)blockIdx.x * (TRK_TERMS + 1) + threadIdx.x;
    if (threadIdx.x < TRK_TERMS) {
      float v = 0.0f;
      for (int wv = 0; wv < TRK_THREADS / 32; wv++) v += s_red[wv][threadIdx.x];
      *p = v;
    } else {
      int c = 0;
      for (int wv = 0; wv < TRK_THREADS / 32; wv++) c += __float_as_int(s_red[wv][TRK_TERMS]);
      *p = __int_as_float(c);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;

  // last CTA: fold the partials deterministically, in double: 8 row groups x 28 columns of threads, group g sums the
  // CTAs g, g + 8, ... in order, then column c adds the 8 groups in order
  __threadfence();
  __shared__ double s_fold[8][TRK_TERMS + 1];
  __shared__ float s_sum[TRK_TERMS];
  {
    const int g = threadIdx.x >> 5, c = threadIdx.x & 31;
    if (c <= TRK_TERMS) {
      double v = 0.0;
      int cnt = 0;
      for (u32 k = g; k < gridDim.x; k += 8) {
        const float p = __ldcg(partials + (size_t)k * (TRK_TERMS + 1) + c);
        v += (double)p;
        cnt += __float_as_int(p);
      }
      s_fold[g][c] = c < TRK_TERMS ? v : (double)cnt;
    }
  }
  __syncthreads();
  if (threadIdx.x <= TRK_TERMS) {
    double v = 0.0;
#pragma unroll
    for (int g = 0; g < 8; g++) v += s_fold[g][threadIdx.x];
    if (threadIdx.x < TRK_TERMS) s_sum[threadIdx.x] = (float)v;
    else st->pairs = (int)v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  st->ticket = 0;
  float A[36], b[6], x[6];
  {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = r; c < 6; c++) { A[6 * r + c] = s_sum[k]; A[6 * c + r] = s_sum[k]; k++; }
  }
#pragma unroll
  for (int r = 0; r < 6; r++) b[r] = s_sum[21 + r];
#pragma unroll
  for (int i = 0; i < 36; i++) st->A[i] = A[i];
#pragma unroll
  for (int i = 0; i < 6; i++) st->b[i] = b[i];
  if (!solve) return;
  trk_cholesky(A, b, x);
#pragma unroll
  for (int i = 0; i < 6; i++) st->x[i] = x[i];
  if (isnan(x[0]) || isnan(x[1]) || isnan(x[2]) || isnan(x[3]) || isnan(x[4]) || isnan(x[5])) {
    st->level_lost[level] = 1;  // "Camera tracking is lost": the remaining iterations of this level are skipped
    st->lost = 1;
    return;
  }
  float inc[16], upd[16];
  trk_increment(x, exact != 0, inc);
#pragma unroll
  for (int i = 0; i < 16; i++) upd[i] = st->update[i];
  trk_mul(inc, upd, upd);
#pragma unroll
  for (int i = 0; i < 16; i++) { st->inc[i] = inc[i]; st->update[i] = upd[i]; }
}

__global__ void k_track_begin(TrackState* st) {
  if (threadIdx.x == 0) {
    trk_identity(st->update);
    trk_identity(st->inc);
    for (int i = 0; i < TRK_LEVELS; i++) st->level_lost[i] = 0;
    st->lost = 0;
    st->pairs = 0;
    st->ticket = 0;
  }
}

// rgbd_camera.cpp:171-173: position_ = vec3(vec4(position_, 1) * update_trans) (row vector times matrix: quirk Q18,
// an affine update leaves a zero position at zero); orientation_ = mat3(mat4(orientation_) * update_trans).
// exact: world = world * update, position / orientation read from it.
__device__ void trk_publish_pose(TrackState* st, int exact) {
  if (exact) {
    for (int i = 0; i < 16; i++) st->pose[i] = st->world[i];
    return;
  }
  // mat4(orientation) * translate(position): columns 0-2 = orientation, column 3 = orientation * position
  for (int i = 0; i < 16; i++) st->pose[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) st->pose[4 * c + k] = st->orientation[3 * c + k];
  for (int k = 0; k < 3; k++)
    st->pose[12 + k] = st->orientation[k] * st->position[0] + st->orientation[3 + k] * st->position[1] +
                       st->orientation[6 + k] * st->position[2] + 0.0f;
}

__global__ void k_track_finish(TrackState* st, int tracked, int exact) {
  if (threadIdx.x != 0) return;
  st->frames++;
  if (!tracked) {
    trk_publish_pose(st, exact);
    return;
  }
  if (exact) {
    trk_mul(st->world, st->update, st->world);
    for (int c = 0; c < 3; c++)
      for (int k = 0; k < 3; k++) st->orientation[3 * c + k] = st->world[4 * c + k];
    for (int k = 0; k < 3; k++) st->position[k] = st->world[12 + k];
    trk_publish_pose(st, exact);
    return;
  }
  const float p[4] = {st->position[0], st->position[1], st->position[2], 1.0f};
  float np[3];
  for (int c = 0; c < 3; c++)
    np[c] = p[0] * st->update[4 * c] + p[1] * st->update[4 * c + 1] + p[2] * st->update[4 * c + 2] +
            p[3] * st->update[4 * c + 3];
  for (int k = 0; k < 3; k++) st->position[k] = np[k];
  float o[16], r[16];
  trk_identity(o);
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) o[4 * c + k] = st->orientation[3 * c + k];
  trk_mul(o, st->update, r);
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) st->orientation[3 * c + k] = r[4 * c + k];
  trk_publish_pose(st, exact);
}

// ------------------------------------------------------------------------------------------------ host side

static int icp_grid(int n, int num_sms) {
  // about 4 points per thread (the 27-term reduction of a CTA costs as much as a few points per thread), at most
  // two CTAs per SM: the last CTA folds one partial row per CTA
  int g = (n + 4 * TRK_THREADS - 1) / (4 * TRK_THREADS);
  const int cap = num_sms * 2 < TRK_MAX_CTAS ? num_sms * 2 : TRK_MAX_CTAS;
  return g < cap ? (g > 0 ? g : 1) : cap;
}

static osl_status init_state(osl_tracker* t) {
  TrackState s;
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < 16; i++) s.update[i] = s.inc[i] = s.world[i] = s.pose[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  s.orientation[0] = s.orientation[4] = s.orientation[8] = 1.0f;  // glm's default constructors (rgbd_camera.cpp:22)
  *t->h_state = s;
  OSL_CUDA(cudaMemcpy(t->d_state, &s, sizeof(s), cudaMemcpyHostToDevice));
  t->pass = 0;
  t->last = 0;
  t->pending = false;
  return OSL_OK;
}

extern "C" {

static osl_status tracker_alloc(osl_tracker* t) {
  cudaDeviceProp prop;
  OSL_CUDA(cudaGetDeviceProperties(&prop, t->device));
  t->num_sms = prop.multiProcessorCount;
  const size_t n0 = (size_t)t->w * t->h;
  for (int s = 0; s < 2; s++)
    for (int i = 0; i < TRK_LEVELS; i++) {
      const size_t n = (size_t)(t->w >> i) * (size_t)(t->h >> i);
      OSL_CUDA(cudaMalloc(&t->vtx[s][i], 12 * n));
      OSL_CUDA(cudaMalloc(&t->nrm[s][i], 12 * n));
    }
  OSL_CUDA(cudaMalloc(&t->work_v, 12 * n0));
  OSL_CUDA(cudaMalloc(&t->work_n, 12 * n0));
  OSL_CUDA(cudaMalloc(&t->filt, 2 * n0));
  OSL_CUDA(cudaMalloc(&t->tmp, 2 * n0));
  OSL_CUDA(cudaMalloc(&t->stage, 2 * n0));
  OSL_CUDA(cudaMalloc(&t->partials, (size_t)TRK_MAX_CTAS * (TRK_TERMS + 1) * sizeof(float)));
  OSL_CUDA(cudaMalloc(&t->d_state, sizeof(TrackState)));
  OSL_CUDA(cudaMallocHost(&t->h_state, sizeof(TrackState)));
  OSL_CUDA(cudaEventCreateWithFlags(&t->done, cudaEventDisableTiming));
  return init_state(t);
}

osl_status osl_tracker_create(osl_tracker** out, int width, int height, float fx, float fy, int flags, int device) {
  if (!out || width < 8 || height < 8 || (width % 4) || (height % 4) || !(fx > 0.0f) || !(fy > 0.0f))
    return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(device));
  osl_tracker* t = new osl_tracker();
  memset(t, 0, sizeof(*t));
  t->device = device; t->w = width; t->h = height; t->fx = fx; t->fy = fy; t->flags = flags;
  const osl_status rc = tracker_alloc(t);
  if (rc) {  // release whatever was allocated before the failure (every pointer starts out null)
    osl_tracker_destroy(t);
    return rc;
  }
  *out = t;
  return OSL_OK;
}

void osl_tracker_destroy(osl_tracker* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  for (int s = 0; s < 2; s++)
    for (int i = 0; i < TRK_LEVELS; i++) { cudaFree(t->vtx[s][i]); cudaFree(t->nrm[s][i]); }
  cudaFree(t->work_v); cudaFree(t->work_n); cudaFree(t->filt); cudaFree(t->tmp); cudaFree(t->stage);
  cudaFree(t->partials); cudaFree(t->d_state);
  if (t->h_state) cudaFreeHost(t->h_state);
  if (t->done) cudaEventDestroy(t->done);
  delete t;
}

osl_status osl_tracker_reset(osl_tracker* t) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  OSL_CUDA(cudaDeviceSynchronize());
  return init_state(t);
}

// RGBDCamera::update (rgbd_camera.cpp:53-191) for a depth image in device memory.  Asynchronous on `stream`.
osl_status osl_tracker_update(osl_tracker* t, const uint16_t* d_depth, void* stream) {
  if (!t || !d_depth) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int cur = t->last ^ 1, last = t->last;
  const bool exact = (t->flags & 1) != 0;
  const float sig_spat = 0.5f / (4.5f * 4.5f);
  const float sig_dep = (float)(0.5 / (double)(40.0f * 40.0f));
  int launches = 0;
  k_track_begin<<<1, 32, 0, st>>>(t->d_state);
  k_bilateral<<<dim3((t->w + BIL_TX - 1) / BIL_TX, (t->h + BIL_TY - 1) / BIL_TY), dim3(BIL_TX, BIL_TY), 0, st>>>(
      d_depth, t->filt, t->w, t->h, sig_spat, sig_dep);
  launches += 2;
  uint16_t *src = t->filt, *dst = t->tmp;
  for (int i = 0; i < TRK_LEVELS; i++) {
    const int wi = t->w >> i, hi = t->h >> i, n = wi * hi;
    k_vertex_normal<<<(n + 255) / 256, 256, 0, st>>>(src, t->vtx[cur][i], t->nrm[cur][i], wi, hi, t->fx, t->fy, t->w,
                                                     t->h);
    launches++;
    if (i != TRK_LEVELS - 1) {
      const int n2 = (wi / 2) * (hi / 2);
      k_subsample_depth<<<(n2 + 255) / 256, 256, 0, st>>>(src, dst, wi / 2, hi / 2, 40.0f * 3.0f);
      launches++;
      uint16_t* sw = src; src = dst; dst = sw;
    }
  }
  const bool tracked = t->pass >= 1;
  if (tracked) {
    for (int i = TRK_LEVELS - 1; i >= 0; i--) {
      const int n = (t->w >> i) * (t->h >> i);
      const int grid = icp_grid(n, t->num_sms);
      for (int j = 0; j < TRK_ITERS[i]; j++) {
        const bool first = j == 0;
        const float* sv = first ? t->vtx[cur][i] : t->work_v;
        const float* sn = first ? t->nrm[cur][i] : t->work_n;
        const int apply = first ? (i < TRK_LEVELS - 1 ? 1 : 0) : 2;
        k_icp_step<<<grid, TRK_THREADS, 0, st>>>(t->vtx[last][i], t->nrm[last][i], sv, sn, t->work_v, t->work_n, n,
                                                apply, i, 1, exact ? 1 : 0, t->d_state, t->partials);
        launches++;
      }
    }
  }
  k_track_finish<<<1, 32, 0, st>>>(t->d_state, tracked ? 1 : 0, exact ? 1 : 0);
  launches++;
  OSL_CUDA(cudaGetLastError());
  OSL_CUDA(cudaMemcpyAsync(t->h_state, t->d_state, sizeof(TrackState), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaEventRecord(t->done, st));
  OSL_LAUNCHED(launches);
  t->pending = true;
  if (t->pass < 2) t->pass++;
  t->last = cur;
  return OSL_OK;
}

// Same with the depth image in host memory (OpenNIDevice::readFrame's H2D copy, openni_device.cpp:122, folded in).
osl_status osl_tracker_update_host(osl_tracker* t, const uint16_t* h_depth, void* stream) {
  if (!t || !h_depth) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  OSL_CUDA(cudaMemcpyAsync(t->stage, h_depth, 2 * (size_t)t->w * t->h, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return osl_tracker_update(t, t->stage, stream);
}

// Waits for the last update.  pose = the matrix main.cpp:40 applies to the vertex map,
// mat4(orientation_) * translate(mat4(1), position_) (exact mode: the camera-to-world pose).
osl_status osl_tracker_get_pose(osl_tracker* t, float pose[16], float position[3], float orientation[9], int* lost,
                                int* pairs) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  if (t->pending) {
    OSL_CUDA(cudaEventSynchronize(t->done));
    t->pending = false;
  }
  const TrackState& s = *t->h_state;
  if (position) memcpy(position, s.position, 12);
  if (orientation) memcpy(orientation, s.orientation, 36);
  if (lost) *lost = s.lost;
  if (pairs) *pairs = s.pairs;
  if (pose) memcpy(pose, s.pose, 64);
  return OSL_OK;
}

// Device address of the 16 floats osl_tracker_get_pose reports as `pose` (column-major), valid for the tracker's
// lifetime and rewritten by every update: hand it to osl_integrate_depth_posed on the same stream to fuse the frame
// the tracker has just localised without waiting for the pose on the host.
osl_status osl_tracker_pose_device(osl_tracker* t, const float** d_pose) {
  if (!t || !d_pose) return OSL_ERR_INVALID;
  *d_pose = t->d_state->pose;
  return OSL_OK;
}

// The pyramid of the last processed frame (device pointers, valid until the next update): for tests and for callers
// that want the filtered vertex / normal maps.
osl_status osl_tracker_view(osl_tracker* t, int level, const float** d_vertex, const float** d_normal, int* width,
                            int* height) {
  if (!t || level < 0 || level >= TRK_LEVELS) return OSL_ERR_INVALID;
  if (d_vertex) *d_vertex = t->vtx[t->last][level];
  if (d_normal) *d_normal = t->nrm[t->last][level];
  if (width) *width = t->w >> level;
  if (height) *height = t->h >> level;
  return OSL_OK;
}

// ---- the reference's free functions on device buffers ------------------------------------------------------------

osl_status osl_bilateral_filter(const uint16_t* d_in, uint16_t* d_out, int width, int height, void* stream) {
  if (!d_in || !d_out || width <= 0 || height <= 0) return OSL_ERR_INVALID;
  k_bilateral<<<dim3((width + BIL_TX - 1) / BIL_TX, (height + BIL_TY - 1) / BIL_TY), dim3(BIL_TX, BIL_TY), 0,
                (cudaStream_t)stream>>>(d_in, d_out, width, height, 0.5f / (4.5f * 4.5f),
                                        (float)(0.5 / (double)(40.0f * 40.0f)));
  OSL_CUDA(cudaGetLastError());
  OSL_LAUNCHED(1);
  return OSL_OK;
}

// (width, height) are the dimensions of d_in; d_out receives (width/2) x (height/2) and must not alias d_in
osl_status osl_subsample_depth(const uint16_t* d_in, uint16_t* d_out, int width, int height, void* stream) {
  if (!d_in || !d_out || width < 2 || height < 2 || d_in == d_out) return OSL_ERR_INVALID;
  const int n = (width / 2) * (height / 2);
  k_subsample_depth<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, width / 2, height / 2, 120.0f);
  OSL_CUDA(cudaGetLastError());
  OSL_LAUNCHED(1);
  return OSL_OK;
}

osl_status osl_subsample_f32(const float* d_in, float* d_out, int width, int height, void* stream) {
  if (!d_in || !d_out || width < 2 || height < 2 || d_in == d_out) return OSL_ERR_INVALID;
  const int n = (width / 2) * (height / 2);
  k_subsample_f32<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, width / 2, height / 2);
  OSL_CUDA(cudaGetLastError());
  OSL_LAUNCHED(1);
  return OSL_OK;
}

osl_status osl_generate_normal_map(const float* d_vertex, float* d_normal, int width, int height, void* stream) {
  if (!d_vertex || !d_normal || width <= 0 || height <= 0) return OSL_ERR_INVALID;
  const int n = width * height;
  k_normal_map<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_vertex, d_normal, width, height);
  OSL_CUDA(cudaGetLastError());
  OSL_LAUNCHED(1);
  return OSL_OK;
}

osl_status osl_transform_normal_map(float* d_normal, const float trans[16], int n, void* stream) {
  if (!d_normal || !trans || n < 0) return OSL_ERR_INVALID;
  if (n == 0) return OSL_OK;
  Mat16 M;
  memcpy(M.m, trans, 64);
  k_transform_normals<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_normal, M, n);
  OSL_CUDA(cudaGetLastError());
  OSL_LAUNCHED(1);
  return OSL_OK;
}

osl_status osl_color_to_intensity(const uint8_t* d_rgb, float* d_out, int n, void* stream) {
  if (!d_rgb || !d_out || n < 0) return OSL_ERR_INVALID;
  if (n == 0) return OSL_OK;
  k_color_to_intensity<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_rgb, d_out, n);
  OSL_CUDA(cudaGetLastError());
  OSL_LAUNCHED(1);
  return OSL_OK;
}

// computeICPCost2 (localization_kernels.cu:313-330): A[36] (row-major, symmetric) and b[6] in HOST memory.
// Synchronises the stream.  flags bit 0: exact Jacobian.
osl_status osl_icp_cost(const float* d_last_vertex, const float* d_last_normal, const float* d_this_vertex,
                        const float* d_this_normal, int n, int flags, float A[36], float b[6], int* pairs,
                        void* stream) {
  if (!d_last_vertex || !d_last_normal || !d_this_vertex || !d_this_normal || n <= 0 || !A || !b) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  OSL_CUDA(cudaGetDevice(&dev));
  OSL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TrackState* d_state = nullptr;
  float* d_part = nullptr;
  TrackState h;
  cudaError_t e = cudaMallocAsync(&d_state, sizeof(TrackState), st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_part, (size_t)TRK_MAX_CTAS * (TRK_TERMS + 1) * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_state, 0, sizeof(TrackState), st);
  if (e == cudaSuccess) {
    float* nv = const_cast<float*>(d_this_vertex);  // never written: apply = 0 and dst == src
    float* nn = const_cast<float*>(d_this_normal);
    k_icp_step<<<icp_grid(n, sms), TRK_THREADS, 0, st>>>(d_last_vertex, d_last_normal, d_this_vertex, d_this_normal, nv,
                                                        nn, n, 0, 0, 0, flags & 1, d_state, d_part);
    e = cudaGetLastError();
    OSL_LAUNCHED(1);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d_state, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (d_state) cudaFreeAsync(d_state, st);  // released on every path
  if (d_part) cudaFreeAsync(d_part, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    g_osl_last_cuda_error = (int)e;
    return e == cudaErrorMemoryAllocation ? OSL_ERR_OOM : OSL_ERR_CUDA;
  }
  memcpy(A, h.A, sizeof(h.A));
  memcpy(b, h.b, sizeof(h.b));
  if (pairs) *pairs = h.pairs;
  return OSL_OK;
}

}  // extern "C"
