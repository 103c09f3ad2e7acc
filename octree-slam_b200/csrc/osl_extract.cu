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
// One launch per level, NO host round trip between levels: the level's input count is read from cnt[lvl] on the
// device and its output count written to cnt[lvl + 1]; a fixed grid of co-resident CTAs walks the tiles in order
// (CTA b takes tiles b, b + G, ...), so the single-word look-back never waits for a tile that is not running.
// Status words carry an epoch (call, level) and are never reset.
__global__ void __launch_bounds__(EX_THREADS)
k_extract_level(const u32* __restrict__ pool, const long long* __restrict__ keys_in, const u32* __restrict__ nodes_in,
                long long* __restrict__ keys_out, u32* __restrict__ nodes_out, unsigned long long* status, int* cnt,
                int lvl, u32 epoch) {
  __shared__ u32 s_warp[EX_THREADS / 32];
  __shared__ u32 s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_in = cnt[lvl];
  const int tiles = (n_in + EX_THREADS - 1) / EX_THREADS;
  if (tiles == 0) {
    if (blockIdx.x == 0 && tid == 0) cnt[lvl + 1] = 0;
    return;
  }
  const unsigned long long tag = (unsigned long long)epoch << 32;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
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
    const u32 c_ = __popc(mask);
    u32 incl = c_;
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
        if (lane == 0) *(volatile unsigned long long*)&status[0] = tag | (2u << 30) | total;
      } else {
        if (lane == 0) *(volatile unsigned long long*)&status[tile] = tag | (1u << 30) | total;
        int look = tile - 1;
        for (;;) {
          const int idx = look - lane;
          u32 v = (2u << 30);
          if (idx >= 0) {
            unsigned long long w = *(volatile unsigned long long*)&status[idx];
            v = ((w >> 32) == epoch) ? (u32)w : 0u;
          }
          while (__any_sync(FULL, (v >> 30) == 0)) {
            if ((v >> 30) == 0) {
              const unsigned long long w = *(volatile unsigned long long*)&status[idx];
              v = ((w >> 32) == epoch) ? (u32)w : 0u;
            }
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
        if (lane == 0) *(volatile unsigned long long*)&status[tile] = tag | (2u << 30) | (excl + total);
      }
      if (lane == 0) {
        s_base = excl;
        if (tile == tiles - 1) cnt[lvl + 1] = (int)(excl + total);
      }
    }
    __syncthreads();
    u32 pos = s_base + woff + (incl - c_);
#pragma unroll
    for (int c = 0; c < 8; c++) {
      if ((mask >> c) & 1u) {
        keys_out[pos] = (key << 3) + c;
        nodes_out[pos] = ptr + c;
        pos++;
      }
    }
    __syncthreads();  // s_warp / s_base are reused by the next tile
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

// scratch of the extraction, kept in the tree between calls (the reference mallocs 8x the frontier per level)
static osl_status ensure_extract_scratch(osl_svo* t, size_t maxn) {
  if (maxn <= t->ex_cap) return OSL_OK;
  cudaFree(t->ex_kA); cudaFree(t->ex_kB); cudaFree(t->ex_nA); cudaFree(t->ex_nB); cudaFree(t->ex_status);
  t->ex_kA = t->ex_kB = nullptr; t->ex_nA = t->ex_nB = nullptr; t->ex_status = nullptr;
  t->ex_cap = 0;
  size_t cap = 4096;
  while (cap < maxn) cap *= 2;
  OSL_CUDA(cudaMalloc(&t->ex_kA, cap * 8)); OSL_CUDA(cudaMalloc(&t->ex_kB, cap * 8));
  OSL_CUDA(cudaMalloc(&t->ex_nA, cap * 4)); OSL_CUDA(cudaMalloc(&t->ex_nB, cap * 4));
  const size_t tiles = (cap + EX_THREADS - 1) / EX_THREADS;
  OSL_CUDA(cudaMalloc(&t->ex_status, tiles * 8));
  OSL_CUDA(cudaMemset(t->ex_status, 0, tiles * 8));
  if (!t->ex_cnt) {
    OSL_CUDA(cudaMalloc(&t->ex_cnt, (OSL_MAX_DEPTH + 2) * sizeof(int)));
  }
  t->ex_cap = cap;
  t->ex_valid = 0;
  return OSL_OK;
}

extern "C" osl_status osl_extract_voxels(const osl_svo* tc, int max_depth, float* d_centers4, float* d_colors4,
                                         int64_t* d_keys, int64_t cap, int64_t* n_out, void* stream) {
  osl_svo* t = const_cast<osl_svo*>(tc);
  if (!t || !n_out || max_depth < 0 || max_depth > OSL_MAX_DEPTH) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  *n_out = 0;
  OSL_CUDA(cudaSetDevice(t->device));
  {
    osl_status prc = osl_poll_results(t, true);  // frames in flight define the node count
    if (prc) return prc;
    osl_status jr = osl_join(t, st);
    if (jr) return jr;
  }
  if (t->size == 0) return OSL_OK;
  osl_status rc = ensure_extract_scratch(t, (size_t)t->size + 8);
  if (rc) return rc;
  // the usual call pattern is "count, allocate, fill": the second call finds the frontier of the first one
  const bool cached = t->ex_valid && t->ex_seq == t->seq && t->ex_depth == max_depth && t->ex_size == t->size &&
                      t->ex_uploads == t->upload_count;
  int n = t->ex_n;
  if (!cached) {
    const long long one = 1; const u32 root = 0xFFFFFFFFu; const int first = 1;
    OSL_CUDA(cudaMemcpyAsync(t->ex_kA, &one, 8, cudaMemcpyHostToDevice, st));
    OSL_CUDA(cudaMemcpyAsync(t->ex_nA, &root, 4, cudaMemcpyHostToDevice, st));
    OSL_CUDA(cudaMemcpyAsync(t->ex_cnt, &first, 4, cudaMemcpyHostToDevice, st));
    long long *kin = t->ex_kA, *kout = t->ex_kB;
    u32 *nin = t->ex_nA, *nout = t->ex_nB;
    const int grid = t->num_sms > 0 ? t->num_sms : 1;  // one CTA per SM: all co-resident
    for (int lvl = 0; lvl < max_depth; lvl++) {
      const u32 epoch = (u32)(++t->ex_epoch);
      k_extract_level<<<grid, EX_THREADS, 0, st>>>(t->d_pool, kin, nin, kout, nout, t->ex_status, t->ex_cnt, lvl, epoch);
      OSL_LAUNCHED(1);
      long long* tk = kin; kin = kout; kout = tk;
      u32* tn = nin; nin = nout; nout = tn;
    }
    OSL_CUDA(cudaMemcpyAsync(&n, t->ex_cnt + max_depth, 4, cudaMemcpyDeviceToHost, st));
    OSL_CUDA(cudaStreamSynchronize(st));
    t->ex_res_k = kin; t->ex_res_n = nin;
    t->ex_n = n; t->ex_seq = t->seq; t->ex_depth = max_depth; t->ex_size = t->size; t->ex_uploads = t->upload_count;
    t->ex_valid = 1;
  }
  *n_out = n;
  if (n > 0 && n <= cap && (d_centers4 || d_colors4 || d_keys)) {
    k_extract_finish<<<(n + 255) / 256, 256, 0, st>>>(t->d_pool, t->ex_res_k, t->ex_res_n, n, t->tp, max_depth,
                                                     reinterpret_cast<float4*>(d_centers4),
                                                     reinterpret_cast<float4*>(d_colors4), (long long*)d_keys);
    OSL_LAUNCHED(1);
    OSL_CUDA(cudaStreamSynchronize(st));
  }
  return OSL_OK;
}
