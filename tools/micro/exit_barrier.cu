// Micro-test: __syncthreads() in a 512-thread CTA after warps 8..15 have exited (the k_frame sort role runs on 256
// of the CTA's 512 threads).  Prints "ok" when the surviving 256 threads pass 1000 barriers with correct data.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out) {
  __shared__ int s[256];
  if (threadIdx.x >= 256) return;
  int acc = 0;
  for (int it = 0; it < 1000; it++) {
    s[threadIdx.x] = it + threadIdx.x;
    __syncthreads();
    acc += s[(threadIdx.x + 37) & 255];
    __syncthreads();
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
int main() {
  int* d;
  cudaMalloc(&d, 64 * 256 * 4);
  k<<<64, 512>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  int h[256];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long want = 0;
  for (int it = 0; it < 1000; it++) want += it + ((0 + 37) & 255);
  printf("%s err=%d got=%d want=%ld\n", (e == cudaSuccess && h[0] == want) ? "ok" : "FAIL", (int)e, h[0], want);
  return 0;
}
