// osl_image.cu -- stand-alone versions of the three per-frame image kernels main.cpp calls (main.cpp:39-43) and of
// computeKeys, for callers that use the reference's un-fused API.  osl_integrate_depth fuses all of them.
//   generateVertexMap             image_kernels.cu:24-58
//   transformVertexMap            image_kernels.cu:206-219
//   computePointCloudBoundingBox  image_kernels.cu:60-102
//   computeKeys<T>                svo.cu:93-106
#include "osl_internal.cuh"

__global__ void __launch_bounds__(256)
k_vertex_map(const uint16_t* __restrict__ depth, float* __restrict__ xyz, int width, int height, float fx, float fy,
             int img_w, int img_h) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width * height) return;
  float X, Y, Z;
  osl_vertex((int)__ldg(depth + idx), idx % width, idx / width, width, height, img_w, img_h, fx, fy, X, Y, Z);
  xyz[3 * (size_t)idx] = X; xyz[3 * (size_t)idx + 1] = Y; xyz[3 * (size_t)idx + 2] = Z;
}

struct Mat16 { float m[16]; };

__global__ void __launch_bounds__(256) k_transform(float* __restrict__ xyz, Mat16 M, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float x = xyz[3 * (size_t)idx], y = xyz[3 * (size_t)idx + 1], z = xyz[3 * (size_t)idx + 2];
  osl_transform(M.m, x, y, z);
  xyz[3 * (size_t)idx] = x; xyz[3 * (size_t)idx + 1] = y; xyz[3 * (size_t)idx + 2] = z;
}

__global__ void __launch_bounds__(256)
k_compute_keys(const float* __restrict__ pts, int stride, int n, TreeParams tp, long long* __restrict__ keys) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float* q = pts + (size_t)stride * idx;
  u64 k;
  const bool ok = osl_key(__ldg(q), __ldg(q + 1), __ldg(q + 2), tp, k);
  keys[idx] = ok ? (long long)(k | (1ull << (3 * tp.D))) : 1ll;
}

// Bounding box.  The reference folds with NON-associative functors (min_vec3/max_vec3): a (0,0,0) accumulator is
// replaced by the next element (valid or not), elements with non-finite x or z are skipped, otherwise component-wise
// fmin/fmax.  Canonical order = sequential left fold (oracle: orc_bbox).  Parallel evaluation of that fold:
//   acc0 = init; if acc0 == 0: acc = p[0] (whatever it is) and the fold continues from element 1;
//   then acc = fmin/fmax over the valid elements.  (The zero rule re-triggering mid-fold -- the running box being
//   exactly (0,0,0) -- cannot happen for depth data, z > 0; it is not reproduced.)
__global__ void __launch_bounds__(256)
k_bbox_partial(const float* __restrict__ xyz, int start, int n, float* __restrict__ partial) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = start + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = __ldg(xyz + 3 * (size_t)i), y = __ldg(xyz + 3 * (size_t)i + 1), z = __ldg(xyz + 3 * (size_t)i + 2);
    if (isfinite(x) && isfinite(z)) {
      lo[0] = fminf(x, lo[0]); lo[1] = fminf(y, lo[1]); lo[2] = fminf(z, lo[2]);
      hi[0] = fmaxf(x, hi[0]); hi[1] = fmaxf(y, hi[1]); hi[2] = fmaxf(z, hi[2]);
    }
  }
  __shared__ float s[6][256 / 32];
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < 3; k++) { s[k][warp] = lo[k]; s[3 + k][warp] = hi[k]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = s[threadIdx.x][0];
    for (int w = 1; w < 256 / 32; w++) v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][w]) : fmaxf(v, s[threadIdx.x][w]);
    partial[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

extern "C" {

osl_status osl_generate_vertex_map(const uint16_t* d_depth, float* d_xyz, int width, int height, float fx, float fy,
                                   int img_w, int img_h, void* stream) {
  if (!d_depth || !d_xyz || width <= 0 || height <= 0) return OSL_ERR_INVALID;
  const int n = width * height;
  k_vertex_map<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_depth, d_xyz, width, height, fx, fy, img_w, img_h);
  OSL_LAUNCHED(1);
  OSL_CUDA(cudaGetLastError());
  return OSL_OK;
}

osl_status osl_transform_vertex_map(float* d_xyz, const float trans[16], int n, void* stream) {
  if (!d_xyz || !trans || n < 0) return OSL_ERR_INVALID;
  if (n == 0) return OSL_OK;
  Mat16 M;
  for (int i = 0; i < 16; i++) M.m[i] = trans[i];
  k_transform<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_xyz, M, n);
  OSL_LAUNCHED(1);
  OSL_CUDA(cudaGetLastError());
  return OSL_OK;
}

osl_status osl_compute_keys(const float* d_pts, int stride, int n, const float center[3], float half_edge,
                            int max_depth, int64_t* d_keys, void* stream) {
  if (!d_pts || !d_keys || n < 0 || (stride != 3 && stride != 4) || max_depth < 1 || max_depth > OSL_MAX_DEPTH)
    return OSL_ERR_INVALID;
  if (n == 0) return OSL_OK;
  TreeParams tp = {center[0], center[1], center[2], half_edge, max_depth, 1};
  k_compute_keys<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_pts, stride, n, tp, (long long*)d_keys);
  OSL_LAUNCHED(1);
  OSL_CUDA(cudaGetLastError());
  return OSL_OK;
}

osl_status osl_point_cloud_bbox(const float* d_xyz, int n, float bbox[6], void* stream) {
  if (!d_xyz || !bbox || n < 0) return OSL_ERR_INVALID;
  if (n == 0) return OSL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float lo[3] = {bbox[0], bbox[1], bbox[2]}, hi[3] = {bbox[3], bbox[4], bbox[5]};
  const bool lo_zero = lo[0] == 0.0f && lo[1] == 0.0f && lo[2] == 0.0f;
  const bool hi_zero = hi[0] == 0.0f && hi[1] == 0.0f && hi[2] == 0.0f;
  float first[3];
  OSL_CUDA(cudaMemcpyAsync(first, d_xyz, 12, cudaMemcpyDeviceToHost, st));
  const int blocks = 296;
  float* d_partial;
  OSL_CUDA(cudaMalloc(&d_partial, blocks * 6 * sizeof(float)));
  // the fold over elements [1, n) (element 0 is folded on the host because of the zero rule)
  k_bbox_partial<<<blocks, 256, 0, st>>>(d_xyz, 1, n, d_partial);
  OSL_LAUNCHED(1);
  float h_partial[296 * 6];
  cudaError_t e = cudaMemcpyAsync(h_partial, d_partial, sizeof(h_partial), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_partial);
  OSL_CUDA(e);
  const bool first_bad = !isfinite(first[0]) || !isfinite(first[2]);
  for (int k = 0; k < 3; k++) {
    if (lo_zero) lo[k] = first[k]; else if (!first_bad) lo[k] = fminf(first[k], lo[k]);
    if (hi_zero) hi[k] = first[k]; else if (!first_bad) hi[k] = fmaxf(first[k], hi[k]);
  }
  for (int b = 0; b < blocks; b++)
    for (int k = 0; k < 3; k++) {
      lo[k] = fminf(h_partial[b * 6 + k], lo[k]);
      hi[k] = fmaxf(h_partial[b * 6 + 3 + k], hi[k]);
    }
  for (int k = 0; k < 3; k++) { bbox[k] = lo[k]; bbox[3 + k] = hi[k]; }
  return OSL_OK;
}

}  // extern "C"
