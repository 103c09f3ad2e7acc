// Micro-benchmark: the gap between two dependent kernels of ~20 us each on one stream, as a function of how the second
// one is launched (plain / cooperative), of event records / waits in between, and of programmatic dependent launch.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/launch_gap.cu -o gpurun_out/launch_gap
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

__device__ unsigned long long g_t[2][512];
__device__ unsigned g_c;

__global__ void __launch_bounds__(512, 3) k_work(int spin_us, int pdl) {
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned idx = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { idx = g_c++ & 511; g_t[0][idx] = t0; }
  unsigned long long t;
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < (unsigned long long)spin_us * 1000ull);
  if (blockIdx.x == 0 && threadIdx.x == 0) g_t[1][idx] = t;
}

static void report(const char* name, int n) {
  cudaDeviceSynchronize();
  unsigned long long h[2][512];
  cudaMemcpyFromSymbol(h, g_t, sizeof(h));
  double gap = 0, per = 0;
  int cnt = 0;
  for (int i = n / 2; i < n - 1; i++) { gap += (double)(h[0][i + 1] - h[1][i]); per += (double)(h[0][i + 1] - h[0][i]); cnt++; }
  printf("%-64s period %6.2f us   end->start gap %5.2f us\n", name, per / cnt / 1e3, gap / cnt / 1e3);
  unsigned z = 0;
  cudaMemcpyToSymbol(g_c, &z, sizeof(z));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaStream_t s, s2;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t ev[8], done;
  for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
  cudaEventRecord(done, s2);
  const int N = 200;
  int spin = 20, pdl = 0;
  for (int grid : {16, 49, 148, 444}) {
    char name[128];
    void* args[] = {&spin, &pdl};
    for (int i = 0; i < N; i++) k_work<<<grid, 512, 0, s>>>(spin, 0);
    snprintf(name, sizeof(name), "plain, grid %d", grid); report(name, N);
    for (int i = 0; i < N; i++) cudaLaunchCooperativeKernel((void*)k_work, dim3(grid), dim3(512), args, 0, s);
    snprintf(name, sizeof(name), "cooperative, grid %d", grid); report(name, N);
    for (int i = 0; i < N; i++) {
      cudaStreamWaitEvent(s, done, 0);
      cudaStreamWaitEvent(s, done, 0);
      cudaLaunchCooperativeKernel((void*)k_work, dim3(grid), dim3(512), args, 0, s);
      cudaEventRecord(ev[i % 8], s);
    }
    snprintf(name, sizeof(name), "cooperative + 2 waits (complete events) + record, grid %d", grid); report(name, N);
    for (int i = 0; i < N; i++) {
      k_work<<<grid, 512, 0, s>>>(spin, 0);
      cudaEventRecord(ev[i % 8], s);
      cudaStreamWaitEvent(s2, ev[i % 8], 0);   // a consumer on another stream, like k_levels
      k_work<<<grid, 512, 0, s2>>>(5, 0);
    }
    snprintf(name, sizeof(name), "plain + record + dependent kernel on 2nd stream, grid %d", grid); report(name, 2 * N);
    // programmatic dependent launch: the next kernel's CTAs are scheduled while the previous one drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int i = 0; i < N; i++) cudaLaunchKernelEx(&cfg, k_work, spin, 1);
    snprintf(name, sizeof(name), "plain + programmatic dependent launch, grid %d", grid); report(name, N);
  }
  return 0;
}
