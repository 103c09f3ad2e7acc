/* osl_oracle_track.c -- CPU restatement of the reference's camera tracking.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as osl_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Restates, function by function:
 *   image_kernels.cu:137-176   bilateralKernel / bilateralFilter
 *   image_kernels.cu:228-283   subsampleDepthKernel / subsampleDepth<uint16_t>
 *   image_kernels.cu:104-135   generateNormalMapKernel
 *   image_kernels.cu:217-230   transformNormalMapKernel
 *   image_kernels.cu:178-192   colorToIntensityKernel, :285-321 subsample<float>
 *   localization_kernels.cu:155-231, 313-330  computeICPCostsUncorrespondedKernel / computeICPCost2
 *   rgbd_camera.cpp:53-191     RGBDCamera::update,  :193-224 solveCholesky
 *
 * Parity pinning: the float shapes of the per-pixel kernels follow the SASS of the reference built with nvcc 12.9
 * for sm_100a (oracle/Makefile `ref`); tests/test_gpu_vs_reference.py runs the reference's own kernels on the GPU
 * box against the CUDA path.  What CANNOT be restated bit-exactly on a CPU: `__expf` (MUFU.EX2) in the bilateral
 * filter -- exp2f is used here, so filtered depths may differ by 1 mm where the quotient sits on a rounding
 * boundary -- and the order of the float reductions (thrust::reduce is unspecified): the sums are taken in double
 * here and compared with a tolerance (north_star: 1e-4 on accumulated values). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ image kernels */

/* image_kernels.cu:137-166.  dims is uint2 in the reference: the squared differences are converted as UNSIGNED
 * (I2FP.F32.U32); shapes: t = color2*sig_dep; t = fma(space2, sig_spat, t); e = ex2(t * -log2e) with the
 * two-step scaling below -126; sum1 = fma(e, depth, sum1); sum2 += e; out = rint(sum1 / sum2), NaN -> 0. */
void orc_bilateral(const uint16_t *in, uint16_t *out, int w, int h) {
  const int ks = 7;
  const float sig_spat = 0.5f / (4.5f * 4.5f);
  const float sig_dep = (float)(0.5 / (double)(40.0f * 40.0f));
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int value = in[y * w + x];
      int tx = x - ks / 2 + ks; if (tx > w - 1) tx = w - 1;
      int ty = y - ks / 2 + ks; if (ty > h - 1) ty = h - 1;
      float sum1 = 0.0f, sum2 = 0.0f;
      for (int cy = (y - ks / 2 > 0 ? y - ks / 2 : 0); cy < ty; cy++)
        for (int cx = (x - ks / 2 > 0 ? x - ks / 2 : 0); cx < tx; cx++) {
          const int depth = in[cy * w + cx];
          const float space2 = (float)(uint32_t)((x - cx) * (x - cx) + (y - cy) * (y - cy));
          const uint32_t dd = (uint32_t)(value - depth);
          const float color2 = (float)(uint32_t)(dd * dd);
          float t = color2 * sig_dep;
          t = fmaf(space2, sig_spat, t);
          t = t * -1.4426950216293334961f;
          float e;
          if (t < -126.0f) { e = exp2f(t * 0.5f); e = e * e; } else e = exp2f(t);
          sum1 = fmaf(e, (float)depth, sum1);
          sum2 = e + sum2;
        }
      const float q = sum1 / sum2;
      int r;
      if (isnan(q)) r = 0;
      else if (q >= 2147483648.0f) r = 2147483647;
      else r = (int)rintf(q);
      out[y * w + x] = (uint16_t)r;
    }
}

/* image_kernels.cu:228-260 + :262-277: (w, h) are the dimensions of `in`; out is (w/2) x (h/2).  The sums are
 * exact integers in FP32; the quotient is an IEEE divide, stored through a truncating float -> u16 conversion. */
void orc_subsample_depth(const uint16_t *in, uint16_t *out, int w, int h) {
  const int width = w / 2, height = h / 2, D = 5;
  const float sigma = 40.0f * 3.0f;
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const float center = (float)in[4 * y * width + 2 * x];
      int tx = 2 * x - D / 2 + D; if (tx > 2 * width - 1) tx = 2 * width - 1;
      int ty = 2 * y - D / 2 + D; if (ty > 2 * height - 1) ty = 2 * height - 1;
      float sum = 0.0f, count = 0.0f;
      for (int cy = (2 * y - D / 2 > 0 ? 2 * y - D / 2 : 0); cy < ty; cy++)
        for (int cx = (2 * x - D / 2 > 0 ? 2 * x - D / 2 : 0); cx < tx; cx++) {
          const float val = (float)in[2 * cy * width + cx];
          if (fabsf(val - center) < sigma) { sum += val; count += 1.0f; }
        }
      const float r = (count == 0.0f) ? 0.0f : sum / count;
      out[y * width + x] = (uint16_t)(uint32_t)r;
    }
}

/* image_kernels.cu:104-129.  cross = v1 x v2 with v1 = right - centre, v2 = below - centre, each component
 * fma(a, b, -(c * d)); dot = fma(cz, cz, fma(cx, cx, cy * cy)); out = c * -(1 / sqrt(dot)) (IEEE sqrt and
 * reciprocal: glm's inversesqrt is 1/sqrt).  The last column and the last row are +inf. */
void orc_normal_map(const float *vtx, float *nrm, int w, int h) {
  for (int idx = 0; idx < w * h; idx++) {
    const int x = idx % w, y = idx / w;
    float *o = nrm + 3 * (size_t)idx;
    if (x == w - 1 || y == h - 1) { o[0] = o[1] = o[2] = INFINITY; continue; }
    const float *c = vtx + 3 * (size_t)idx, *r = c + 3, *b = c + 3 * (size_t)w;
    const float v1x = r[0] - c[0], v1y = r[1] - c[1], v1z = r[2] - c[2];
    const float v2x = b[0] - c[0], v2y = b[1] - c[1], v2z = b[2] - c[2];
    const float cx = fmaf(v1y, v2z, -(v1z * v2y));
    const float cy = fmaf(v1z, v2x, -(v1x * v2z));
    const float cz = fmaf(v1x, v2y, -(v1y * v2x));
    const float dot = fmaf(cz, cz, fmaf(cx, cx, cy * cy));
    const float inv = 1.0f / sqrtf(dot);
    o[0] = cx * -inv; o[1] = cy * -inv; o[2] = cz * -inv;
  }
}

/* image_kernels.cu:217-226: trans * vec4(n, 0) */
void orc_transform_normals(float *nrm, int n, const float M[16]) {
  for (int i = 0; i < n; i++) {
    float *p = nrm + 3 * (size_t)i;
    const float x = p[0], y = p[1], z = p[2];
    float o[3];
    for (int r = 0; r < 3; r++) {
      float t = y * M[4 + r];
      t = fmaf(x, M[0 + r], t);
      const float u = fmaf(z, M[8 + r], M[12 + r] * 0.0f);
      o[r] = t + u;
    }
    p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
  }
}

/* same shape as osl_oracle.c orc_transform (image_kernels.cu:206-215) */
static void transform_points(float *xyz, int n, const float M[16]) {
  for (int i = 0; i < n; i++) {
    float *p = xyz + 3 * (size_t)i;
    const float x = p[0], y = p[1], z = p[2];
    float o[3];
    for (int r = 0; r < 3; r++) {
      float t = y * M[4 + r];
      t = fmaf(x, M[0 + r], t);
      const float u = fmaf(z, M[8 + r], M[12 + r]);
      o[r] = t + u;
    }
    p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
  }
}

/* image_kernels.cu:178-186 (the green channel is never read: r, b, b) */
void orc_color_to_intensity(const uint8_t *rgb, float *out, int n) {
  for (int i = 0; i < n; i++) {
    const float r = (float)rgb[3 * i] / 255.0f, b = (float)rgb[3 * i + 2] / 255.0f;
    out[i] = fmaf(b, 0.114f, fmaf(r, 0.299f, b * 0.587f)); /* SASS: FMUL(b,y); FFMA(r,x,.); FFMA(b,z,.) */
  }
}

/* image_kernels.cu:285-313 subsample<float>: (w, h) are the dimensions of `in` */
void orc_subsample_f32(const float *in, float *out, int w, int h) {
  const int width = w / 2, height = h / 2;
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) out[y * width + x] = in[4 * y * width + 2 * x];
}

/* the vertex map of osl_oracle.c (image_kernels.cu:24-53) */
void orc_vertex_map(const uint16_t *depth_px, float *xyz, int width, int height, float fx, float fy, int img_w,
                    int img_h);

/* ------------------------------------------------------------------------------------------ ICP cost */

/* localization_kernels.cu:155-231 + 313-330 (computeICPCost2): every pixel pairs with the same pixel of the last
 * frame; a pair counts when both points and normals are finite, both depths lie in [0.1, 10] m, the points are
 * within 10 cm and the normals within ~30 degrees.  A_T = (G^T n1, n1) with the reference's G^T (rows
 * (0,-x,-y), (-z,0,x), (y,z,0) of v2 -- quirk Q17: not the skew matrix of v2; `exact_jacobian` != 0 uses v2 x n1
 * instead); b = n1 . (v1 - v2).  Out: A[36] row-major (full, symmetric), b[6].  Sums in double. */
int64_t orc_icp_cost(const float *last_v, const float *last_n, const float *cur_v, const float *cur_n, int n,
                     int exact_jacobian, float A[36], float b[6]) {
  double acc[42];
  for (int i = 0; i < 42; i++) acc[i] = 0.0;
  int64_t pairs = 0;
  for (int i = 0; i < n; i++) {
    const float *v2 = cur_v + 3 * (size_t)i, *n2 = cur_n + 3 * (size_t)i;
    const float *v1 = last_v + 3 * (size_t)i, *n1 = last_n + 3 * (size_t)i;
    if (!isfinite(v2[0]) || !isfinite(v2[1]) || !isfinite(v2[2]) || !isfinite(v1[0]) || !isfinite(v1[1]) ||
        !isfinite(v1[2]) || v1[2] < 0.1f || v2[2] < 0.1f || v1[2] > 10.0f || v2[2] > 10.0f)
      continue;
    if (!isfinite(n2[0]) || !isfinite(n2[1]) || !isfinite(n2[2]) || !isfinite(n1[0]) || !isfinite(n1[1]) ||
        !isfinite(n1[2]))
      continue;
    const float dx = v2[0] - v1[0], dy = v2[1] - v1[1], dz = v2[2] - v1[2];
    if (sqrtf(dx * dx + dy * dy + dz * dz) > 0.1f) continue;
    if (n2[0] * n1[0] + n2[1] * n1[1] + n2[2] * n1[2] < 0.87f) continue;
    float at[6];
    if (exact_jacobian) {
      at[0] = v2[1] * n1[2] - v2[2] * n1[1];
      at[1] = v2[2] * n1[0] - v2[0] * n1[2];
      at[2] = v2[0] * n1[1] - v2[1] * n1[0];
    } else {
      at[0] = 0.0f * n1[0] + -v2[0] * n1[1] + -v2[1] * n1[2];
      at[1] = -v2[2] * n1[0] + 0.0f * n1[1] + v2[0] * n1[2];
      at[2] = v2[1] * n1[0] + v2[2] * n1[1] + 0.0f * n1[2];
    }
    at[3] = n1[0]; at[4] = n1[1]; at[5] = n1[2];
    const float bb = n1[0] * (v1[0] - v2[0]) + n1[1] * (v1[1] - v2[1]) + n1[2] * (v1[2] - v2[2]);
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) acc[6 * r + c] += (double)(at[r] * at[c]);
      acc[36 + r] += (double)(bb * at[r]);
    }
    pairs++;
  }
  for (int i = 0; i < 36; i++) A[i] = (float)acc[i];
  for (int i = 0; i < 6; i++) b[i] = (float)acc[36 + i];
  return pairs;
}

/* rgbd_camera.cpp:193-224 (float storage, double sums) */
void orc_solve_cholesky(int dim, const float *A, const float *b, float *x) {
  float LU[36], yv[6];
  memset(LU, 0, sizeof(LU));
  for (int k = 0; k < dim; k++) {
    double sum = 0.0;
    for (int p = 0; p < k; p++) sum += LU[k * dim + p] * LU[k * dim + p];
    LU[k * dim + k] = (float)sqrt(A[k * dim + k] - sum);
    for (int i = k + 1; i < dim; i++) {
      double s2 = 0.0;
      for (int p = 0; p < k; p++) s2 += LU[i * dim + p] * LU[k * dim + p];
      LU[i * dim + k] = (float)((A[i * dim + k] - s2) / LU[k * dim + k]);
    }
  }
  for (int i = 0; i < dim; i++) {
    double sum = 0.0;
    for (int k = 0; k < i; k++) sum += LU[i * dim + k] * yv[k];
    yv[i] = (float)((b[i] - sum) / LU[i * dim + i]);
  }
  for (int i = dim - 1; i >= 0; i--) {
    double sum = 0.0;
    for (int k = i + 1; k < dim; k++) sum += LU[k * dim + i] * x[k];
    x[i] = (float)((yv[i] - sum) / LU[i * dim + i]);
  }
}

/* ------------------------------------------------------------------------------------------ glm 0.9.5.4 pieces */

static void mat4_identity(float m[16]) { memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.0f; }

/* out = a * b, column-major (glm/detail/type_mat4x4.inl operator*) */
static void mat4_mul(const float a[16], const float b[16], float out[16]) {
  float r[16];
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      r[4 * c + k] = a[k] * b[4 * c] + a[4 + k] * b[4 * c + 1] + a[8 + k] * b[4 * c + 2] + a[12 + k] * b[4 * c + 3];
  memcpy(out, r, 64);
}

/* glm::rotate(mat4(1), angle_deg, axis) for a unit axis (gtc/matrix_transform.inl:48-86, degrees) */
static void mat4_rotate_deg(float angle_deg, const float axis[3], float out[16]) {
  const float a = angle_deg * 0.01745329251994329576923690768489f;
  const float c = cosf(a), s = sinf(a);
  const float t[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
  mat4_identity(out);
  out[0] = c + t[0] * axis[0];
  out[1] = 0 + t[0] * axis[1] + s * axis[2];
  out[2] = 0 + t[0] * axis[2] - s * axis[1];
  out[4] = 0 + t[1] * axis[0] - s * axis[2];
  out[5] = c + t[1] * axis[1];
  out[6] = 0 + t[1] * axis[2] + s * axis[0];
  out[8] = 0 + t[2] * axis[0] + s * axis[1];
  out[9] = 0 + t[2] * axis[1] - s * axis[0];
  out[10] = c + t[2] * axis[2];
}

/* rgbd_camera.cpp:153-158: Rz(-x2) * Ry(-x1) * Rx(-x0) * T(x3, x4, x5), angles converted with 180 / 3.14159f.
 * exact_jacobian (not the reference): T(x3, x4, x5) * Rz(x2) * Ry(x1) * Rx(x0), the increment the linearised
 * point-to-plane residual n . (v + w x v + t - v1) actually solves for. */
void orc_pose_increment(const float x[6], int exact_jacobian, float out[16]) {
  static const float ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
  float rz[16], ry[16], rx[16], t[16], m[16];
  const float sg = exact_jacobian ? 1.0f : -1.0f;
  mat4_rotate_deg(sg * x[2] * 180.0f / 3.14159f, az, rz);
  mat4_rotate_deg(sg * x[1] * 180.0f / 3.14159f, ay, ry);
  mat4_rotate_deg(sg * x[0] * 180.0f / 3.14159f, ax, rx);
  mat4_identity(t);
  t[12] = x[3]; t[13] = x[4]; t[14] = x[5];
  mat4_mul(rz, ry, m);
  mat4_mul(m, rx, m);
  if (exact_jacobian) mat4_mul(t, m, out); /* the linearisation T v = v + w x v + t: rotate, then translate */
  else mat4_mul(m, t, out);
}

/* ------------------------------------------------------------------------------------------ RGBDCamera */

#define ORC_PYR 3
static const int ORC_ITERS[ORC_PYR] = {10, 5, 4}; /* rgbd_camera.cpp:19 */

typedef struct orc_tracker {
  int w, h;
  float fx, fy;
  int exact_jacobian;
  int pass;
  float position[3];
  float orientation[9]; /* column-major mat3 */
  float world[16];      /* exact_jacobian only: camera-to-world pose, world = world * update per frame */
  float *vtx[2][ORC_PYR], *nrm[2][ORC_PYR]; /* [0] = last, [1] = current (swapped per frame) */
  int last, lost;
  int64_t pairs_last_iter;
} orc_tracker;

orc_tracker *orc_tracker_create(int w, int h, float fx, float fy, int exact_jacobian) {
  orc_tracker *t = (orc_tracker *)calloc(1, sizeof(orc_tracker));
  t->w = w; t->h = h; t->fx = fx; t->fy = fy; t->exact_jacobian = exact_jacobian;
  t->orientation[0] = t->orientation[4] = t->orientation[8] = 1.0f; /* glm default constructors */
  mat4_identity(t->world);
  for (int s = 0; s < 2; s++)
    for (int i = 0; i < ORC_PYR; i++) {
      const size_t n = (size_t)(w >> i) * (size_t)(h >> i);
      t->vtx[s][i] = (float *)malloc(12 * n);
      t->nrm[s][i] = (float *)malloc(12 * n);
    }
  return t;
}

void orc_tracker_destroy(orc_tracker *t) {
  if (!t) return;
  for (int s = 0; s < 2; s++)
    for (int i = 0; i < ORC_PYR; i++) { free(t->vtx[s][i]); free(t->nrm[s][i]); }
  free(t);
}

/* main.cpp:40: mat4(orientation) * translate(mat4(1), position) -- the matrix applied to the vertex map */
void orc_tracker_pose(const orc_tracker *t, float pose[16], float position[3], float orientation[9]) {
  if (position) memcpy(position, t->position, 12);
  if (orientation) memcpy(orientation, t->orientation, 36);
  if (pose && t->exact_jacobian) {
    memcpy(pose, t->world, 64);
  } else if (pose) {
    float o[16], tr[16];
    mat4_identity(o);
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) o[4 * c + r] = t->orientation[3 * c + r];
    mat4_identity(tr);
    tr[12] = t->position[0]; tr[13] = t->position[1]; tr[14] = t->position[2];
    mat4_mul(o, tr, pose);
  }
}

int orc_tracker_lost(const orc_tracker *t) { return t->lost; }
int64_t orc_tracker_pairs(const orc_tracker *t) { return t->pairs_last_iter; }

/* rgbd_camera.cpp:53-191.  The intensity pyramid is computed by the reference but never read (computeRGBDCost is
 * empty and its call commented out), so it is not restated here. */
void orc_tracker_update(orc_tracker *t, const uint16_t *depth) {
  const int cur = t->last ^ 1, last = t->last;
  const size_t n0 = (size_t)t->w * (size_t)t->h;
  uint16_t *filt = (uint16_t *)malloc(2 * n0), *tmp = (uint16_t *)malloc(2 * n0);
  orc_bilateral(depth, filt, t->w, t->h);
  for (int i = 0; i < ORC_PYR; i++) {
    const int wi = t->w >> i, hi = t->h >> i;
    orc_vertex_map(filt, t->vtx[cur][i], wi, hi, t->fx, t->fy, t->w, t->h);
    orc_normal_map(t->vtx[cur][i], t->nrm[cur][i], wi, hi);
    if (i != ORC_PYR - 1) {
      orc_subsample_depth(filt, tmp, wi, hi);
      memcpy(filt, tmp, 2 * (size_t)(wi / 2) * (size_t)(hi / 2));
    }
  }
  free(filt); free(tmp);
  t->lost = 0;
  if (t->pass >= 1) {
    float update[16];
    mat4_identity(update);
    for (int i = ORC_PYR - 1; i >= 0; i--) {
      const int n = (t->w >> i) * (t->h >> i);
      float *v = (float *)malloc(12 * (size_t)n), *nn = (float *)malloc(12 * (size_t)n);
      memcpy(v, t->vtx[cur][i], 12 * (size_t)n);
      memcpy(nn, t->nrm[cur][i], 12 * (size_t)n);
      if (i < ORC_PYR - 1) { transform_points(v, n, update); orc_transform_normals(nn, n, update); }
      for (int j = 0; j < ORC_ITERS[i]; j++) {
        float A[36], b[6], x[6], inc[16];
        t->pairs_last_iter = orc_icp_cost(t->vtx[last][i], t->nrm[last][i], v, nn, n, t->exact_jacobian, A, b);
        orc_solve_cholesky(6, A, b, x);
        if (isnan(x[0]) || isnan(x[1]) || isnan(x[2]) || isnan(x[3]) || isnan(x[4]) || isnan(x[5])) {
          t->lost = 1;
          break;
        }
        orc_pose_increment(x, t->exact_jacobian, inc);
        mat4_mul(inc, update, update);
        if (j < ORC_ITERS[i] - 1) { transform_points(v, n, inc); orc_transform_normals(nn, n, inc); }
      }
      free(v); free(nn);
    }
    /* position_ = vec3(vec4(position_, 1) * update_trans)  (row vector times matrix: with an affine update the
     * bottom row is (0,0,0,1), so a position that starts at 0 stays 0 -- quirk Q18);
     * orientation_ = mat3(mat4(orientation_) * update_trans) */
    if (t->exact_jacobian) {
      mat4_mul(t->world, update, t->world);
      for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) t->orientation[3 * c + k] = t->world[4 * c + k];
      memcpy(t->position, t->world + 12, 12);
    } else {
    float p[4] = {t->position[0], t->position[1], t->position[2], 1.0f}, np[3];
    for (int c = 0; c < 3; c++)
      np[c] = p[0] * update[4 * c] + p[1] * update[4 * c + 1] + p[2] * update[4 * c + 2] + p[3] * update[4 * c + 3];
    memcpy(t->position, np, 12);
    float o[16], r[16];
    mat4_identity(o);
    for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) o[4 * c + k] = t->orientation[3 * c + k];
    mat4_mul(o, update, r);
    for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) t->orientation[3 * c + k] = r[4 * c + k];
    }
  }
  if (t->pass < 2) t->pass++;
  t->last = cur;
}
