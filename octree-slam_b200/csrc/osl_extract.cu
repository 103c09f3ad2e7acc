// osl_extract.cu -- SVO -> list of occupied voxels.
// Replaces extractVoxelGridFromSVO (svo.cu:699-745), getOccupiedChildren (:498-536), voxelGridFromKeys (:538-582).
// The reference re-walks the tree from the root for every key at every level, allocates 8x the frontier and
// stream-compacts with Thrust per level.  Here a frontier entry carries its node index, each level is ONE kernel that
// counts, orders (single-word decoupled look-back) and writes the next frontier directly; output order is identical
// (children in octant order under parents in list order = numeric key order).
#include "osl_internal.cuh"

#define EX_THREADS 256
#define FULL 0xFFFFFFFFu

__device__ __forceinline__ u32 ex_ld(const u32* p) { return *(const volatile u32*)p; }
__device__ __forceinline__ void ex_st(u32* p, u32 v) { *(volatile u32*)p = v; }

// frontier entry: leading-1 key + node index (0xFFFFFFFF for the implicit root)
__global__ void __launch_bounds__(EX_THREADS)
k_extract_level(const u32* __restrict__ pool, const long long* __restrict__ keys_in, const u32* __restrict__ nodes_in,
                int n_in, long long* __restrict__ keys_out, u32* __restrict__ nodes_out, u32* status, int* n_out) {
  __shared__ u32 s_warp[EX_THREADS / 32];
  __shared__ u32 s_base;
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = tile * EX_THREADS + tid;
  u32 mask = 0, ptr = 0;
  long long key = 0;
  if (i < n_in) {
    key = keys_in[i];
    const u32 node = nodes_in[i];
    bool has = true;
    if (node != 0xFFFFFFFFu) {
      const u32 w0 = __ldg(pool + 2 * (size_t)node);
      has = (w0 & OSL_FLAG) != 0;
      ptr = w0 & OSL_MASK;
    }
    if (has) {
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const u32 v = __ldg(pool + 2 * (size_t)(ptr + c) + 1);
        if ((v >> 24) > 127u) mask |= 1u << c;  // svo.cu:528
      }
    }
  }
  const u32 cnt = __popc(mask);
  u32 incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  u32 woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < EX_THREADS / 32; w++) {
    const u32 v = s_warp[w];
    if (w < warp) woff += v;
    total += v;
  }
  if (warp == 0) {
    u32 excl = 0;
    if (tile == 0) {
      if (lane == 0) ex_st(&status[0], (2u << 30) | total);
    } else {
      if (lane == 0) ex_st(&status[tile], (1u << 30) | total);
      int look = tile - 1;
      for (;;) {
        const int idx = look - lane;
        u32 v = (idx >= 0) ? ex_ld(&status[idx]) : (2u << 30);
        while (__any_sync(FULL, (v >> 30) == 0)) {
          if ((v >> 30) == 0) v = ex_ld(&status[idx]);
        }
        const u32 inc_mask = __ballot_sync(FULL, (v >> 30) == 2);
        const int stop = inc_mask ? (__ffs(inc_mask) - 1) : 31;
        u32 c = (lane <= stop) ? (v & OSL_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        excl += c;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) ex_st(&status[tile], (2u << 30) | (excl + total));
    }
    if (lane == 0) {
      s_base = excl;
      if (tile == gridDim.x - 1) *n_out = (int)(excl + total);
    }
  }
  __syncthreads();
  u32 pos = s_base + woff + (incl - cnt);
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if ((mask >> c) & 1u) {
      keys_out[pos] = (key << 3) + c;
      nodes_out[pos] = ptr + c;
      pos++;
    }
  }
}

// voxelGridFromKeys (svo.cu:538-582): centre by the same float descent as computeKey, colour = bytes / 255.0f
__global__ void __launch_bounds__(256)
k_extract_finish(const u32* __restrict__ pool, const long long* __restrict__ keys, const u32* __restrict__ nodes, int n,
                 TreeParams tp, int depth, float4* __restrict__ centers, float4* __restrict__ colors,
                 long long* __restrict__ keys_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long key = keys[i];
  float cx = tp.cx, cy = tp.cy, cz = tp.cz, e = tp.half;
  for (int l = depth - 1; l >= 0; l--) {
    const int pos = (int)((key >> (3 * l)) & 7);
    e = __fmul_rn(e, 0.5f);
    cx = __fadd_rn(cx, (pos & 1) ? e : -e);
    cy = __fadd_rn(cy, (pos & 2) ? e : -e);
    cz = __fadd_rn(cz, (pos & 4) ? e : -e);
  }
  const u32 node = nodes[i];
  const u32 v = __ldg(pool + 2 * (size_t)(node == 0xFFFFFFFFu ? 0u : node) + 1);  // root: node_idx stays 0 (svo.cu:551,573)
  if (centers) centers[i] = make_float4(cx, cy, cz, 1.0f);
  if (colors)
    colors[i] = make_float4(__fdiv_rn((float)(v & 0xFFu), 255.0f), __fdiv_rn((float)((v >> 8) & 0xFFu), 255.0f),
                            __fdiv_rn((float)((v >> 16) & 0xFFu), 255.0f), __fdiv_rn((float)(v >> 24), 255.0f));
  if (keys_out) keys_out[i] = key;
}

extern "C" osl_status osl_extract_voxels(const osl_svo* t, int max_depth, float* d_centers4, float* d_colors4,
                                         int64_t* d_keys, int64_t cap, int64_t* n_out, void* stream) {
  if (!t || !n_out || max_depth < 0 || max_depth > OSL_MAX_DEPTH) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  *n_out = 0;
  {
    osl_status prc = osl_poll_results(const_cast<osl_svo*>(t), true);  // frames in flight define the node count
    if (prc) return prc;
    osl_status jr = osl_join(const_cast<osl_svo*>(t), st);
    if (jr) return jr;
  }
  if (t->size == 0) return OSL_OK;
  const size_t maxn = (size_t)t->size + 8;
  long long *kA = nullptr, *kB = nullptr;
  u32 *nA = nullptr, *nB = nullptr, *status = nullptr;
  int* d_n = nullptr;
  osl_status rc = OSL_OK;
  int n = 1;
  cudaError_t e;
#define EX_CHECK(x) do { e = (x); if (e != cudaSuccess) { g_osl_last_cuda_error = (int)e; rc = OSL_ERR_CUDA; goto done; } } while (0)
  EX_CHECK(cudaMalloc(&kA, maxn * 8)); EX_CHECK(cudaMalloc(&kB, maxn * 8));
  EX_CHECK(cudaMalloc(&nA, maxn * 4)); EX_CHECK(cudaMalloc(&nB, maxn * 4));
  EX_CHECK(cudaMalloc(&status, ((maxn + EX_THREADS - 1) / EX_THREADS) * 4));
  EX_CHECK(cudaMalloc(&d_n, 4));
  {
    const long long one = 1; const u32 root = 0xFFFFFFFFu;
    EX_CHECK(cudaMemcpyAsync(kA, &one, 8, cudaMemcpyHostToDevice, st));
    EX_CHECK(cudaMemcpyAsync(nA, &root, 4, cudaMemcpyHostToDevice, st));
    EX_CHECK(cudaStreamSynchronize(st));
  }
  for (int lvl = 0; lvl < max_depth && n > 0; lvl++) {
    const int tiles = (n + EX_THREADS - 1) / EX_THREADS;
    EX_CHECK(cudaMemsetAsync(status, 0, (size_t)tiles * 4, st));
    k_extract_level<<<tiles, EX_THREADS, 0, st>>>(t->d_pool, kA, nA, n, kB, nB, status, d_n);
    OSL_LAUNCHED(1);
    EX_CHECK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, st));
    EX_CHECK(cudaStreamSynchronize(st));
    long long* tk = kA; kA = kB; kB = tk;
    u32* tn = nA; nA = nB; nB = tn;
  }
  *n_out = n;
  if (n > 0 && n <= cap && (d_centers4 || d_colors4 || d_keys)) {
    k_extract_finish<<<(n + 255) / 256, 256, 0, st>>>(t->d_pool, kA, nA, n, t->tp, max_depth,
                                                     reinterpret_cast<float4*>(d_centers4),
                                                     reinterpret_cast<float4*>(d_colors4), (long long*)d_keys);
    OSL_LAUNCHED(1);
    EX_CHECK(cudaStreamSynchronize(st));
  }
done:
  cudaFree(kA); cudaFree(kB); cudaFree(nA); cudaFree(nB); cudaFree(status); cudaFree(d_n);
  return rc;
#undef EX_CHECK
}
