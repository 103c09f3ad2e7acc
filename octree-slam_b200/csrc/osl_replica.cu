// osl_replica.cu -- keeping replicas of ONE map on several GPUs (the raycast shards by image
// rows and every rank needs the tree).  Everything stays in device memory: a collective library (NCCL through
// torch.distributed, or a peer copy) moves bytes between device buffers, this file defines what those bytes are.
//
//   full copy   the flat 2*n uint32 pool itself (the reference's wire format, octree.cpp:113-169): the receiver
//               reserves room (osl_svo_reserve), exposes its pool (osl_svo_pool_device), the collective writes into it,
//               osl_svo_adopt validates and publishes it.
//   delta       what ONE integrate call changed: the nodes it appended (a contiguous tail, new tiles are only ever
//               allocated at the end) and the (index, word0, word1) of every pre-existing node on a touched path -- the
//               frame's level lists.  osl_svo_delta_pack / osl_svo_delta_apply.  ~0.3 MB per 640x480 frame against
//               tens of MB for the pool.
#include <string.h>

#include "osl_internal.cuh"

struct DeltaHeader {
  unsigned int magic;        // 'OSLD'
  int max_depth;
  int size_before, size_after;
  int n_touched;             // triples that follow the header
  int reserved[3];
};

__device__ u32 g_osl_node0 = 0u;  // node 0 is always part of a delta: its value word takes the root average (quirk Q6)

// gather: one thread per level-list entry (levels 1..D concatenated) -> (node index, word0, word1) after the frame
__global__ void k_delta_pack(const u32* __restrict__ pool, const u32* __restrict__ self, int n, uint3* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 node = self[i];
  const uint2 w = reinterpret_cast<const uint2*>(pool)[node];
  out[i] = make_uint3(node, w.x, w.y);
}

__global__ void k_delta_apply(u32* __restrict__ pool, const uint3* __restrict__ in, int n, int limit, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint3 e = in[i];
  if ((int)e.x < 0 || (int)e.x >= limit) { atomicAdd(bad, 1); return; }
  reinterpret_cast<uint2*>(pool)[e.x] = make_uint2(e.y, e.z);
}

extern "C" {

osl_status osl_svo_reserve(osl_svo* t, size_t n_nodes) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  if (n_nodes <= t->cap_nodes) return OSL_OK;
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  OSL_CUDA(cudaDeviceSynchronize());
  return osl_grow_pool(t, n_nodes, 0);
}

osl_status osl_svo_pool_device(osl_svo* t, uint32_t** d_pool, size_t* cap_nodes) {
  if (!t || !d_pool) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);  // nothing of ours may be writing while the caller does
  if (rc) return rc;
  OSL_CUDA(cudaDeviceSynchronize());
  *d_pool = t->d_pool;
  if (cap_nodes) *cap_nodes = t->cap_nodes;
  return OSL_OK;
}

osl_status osl_svo_adopt(osl_svo* t, int n_nodes, int max_depth, const float center[3], float half_edge) {
  if (!t || n_nodes < 0 || (n_nodes > 0 && (n_nodes < 8 || (n_nodes & 7))) || (size_t)n_nodes > t->cap_nodes || !center)
    return OSL_ERR_INVALID;
  // a replica must agree with its source on what a node index means
  if (max_depth != t->tp.D || center[0] != t->tp.cx || center[1] != t->tp.cy || center[2] != t->tp.cz ||
      half_edge != t->tp.half)
    return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  OSL_CUDA(cudaDeviceSynchronize());
  const size_t old = (size_t)(t->size > 8 ? t->size : 8);
  if ((size_t)n_nodes < old)  // invariant: every word beyond the live nodes is zero
    OSL_CUDA(cudaMemset(t->d_pool + 2 * (size_t)n_nodes, 0, (old - (size_t)n_nodes) * 8));
  rc = osl_validate_pool(t, n_nodes);
  if (rc) return rc;
  t->sticky_error = OSL_OK;
  t->upload_count++;  // invalidates the cached extraction frontier
  return osl_set_device_size(t, n_nodes);
}

size_t osl_svo_delta_bytes(osl_svo* t) {
  if (!t) return 0;
  cudaSetDevice(t->device);
  if (osl_poll_results(t, true) != OSL_OK || t->seq == 0) return 0;
  const FrameState& F = t->h_ring[(t->seq - 1) % OSL_RING];
  size_t touched = 0;
  for (int d = 1; d <= t->tp.D; d++) touched += (size_t)F.n_level[d];
  touched += 1;  // node 0
  const size_t appended = (size_t)(F.size_after - F.size_before);
  return sizeof(DeltaHeader) + touched * sizeof(uint3) + (F.overflow ? 0 : appended * 8);
}

osl_status osl_svo_delta_pack(osl_svo* t, void* d_buf, size_t cap, size_t* bytes, void* stream) {
  if (!t || !d_buf || !bytes) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  if (t->seq == 0) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  rc = osl_join(t, st);
  if (rc) return rc;
  const unsigned long long f = t->seq - 1;
  const FrameState& F = t->h_ring[f % OSL_RING];
  const LevelArrays& lv = t->lv[f % OSL_BACK];
  const int D = t->tp.D;
  DeltaHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = 0x444C534Fu; h.max_depth = D;
  h.size_before = F.overflow ? t->size : F.size_before;
  h.size_after = F.overflow ? t->size : F.size_after;
  if (t->size == 0) { h.size_before = h.size_after = 0; }
  size_t touched = 0;
  if (!F.overflow)
    for (int d = 1; d <= D; d++) touched += (size_t)F.n_level[d];
  if (h.size_after >= 8) touched += 1;  // node 0
  h.n_touched = (int)touched;
  const size_t appended = (size_t)(h.size_after - h.size_before);
  const size_t need = sizeof(h) + touched * sizeof(uint3) + appended * 8;
  *bytes = need;
  if (need > cap) return OSL_ERR_INVALID;
  unsigned char* p = static_cast<unsigned char*>(d_buf);
  OSL_CUDA(cudaMemcpyAsync(p, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  size_t done = 0;
  for (int d = 1; d <= D && !F.overflow; d++) {  // the level lists are dense per level, at lv.off[d]
    const int n = F.n_level[d];
    if (n <= 0) continue;
    k_delta_pack<<<(n + 255) / 256, 256, 0, st>>>(t->d_pool, lv.self + lv.off[d], n,
                                                  reinterpret_cast<uint3*>(p + sizeof(h)) + done);
    OSL_LAUNCHED(1);
    done += (size_t)n;
  }
  if (h.size_after >= 8) {
    u32* zero = nullptr;
    OSL_CUDA(cudaGetSymbolAddress((void**)&zero, g_osl_node0));
    k_delta_pack<<<1, 32, 0, st>>>(t->d_pool, zero, 1, reinterpret_cast<uint3*>(p + sizeof(h)) + done);
    OSL_LAUNCHED(1);
    done += 1;
  }
  if (appended)
    OSL_CUDA(cudaMemcpyAsync(p + sizeof(h) + touched * sizeof(uint3), t->d_pool + 2 * (size_t)h.size_before,
                             appended * 8, cudaMemcpyDeviceToDevice, st));
  OSL_CUDA(cudaStreamSynchronize(st));  // (the header came from the stack)
  return OSL_OK;
}

osl_status osl_svo_delta_apply(osl_svo* t, const void* d_buf, size_t bytes, void* stream) {
  if (!t || !d_buf || bytes < sizeof(DeltaHeader)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DeltaHeader h;
  OSL_CUDA(cudaMemcpyAsync(&h, d_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaStreamSynchronize(st));
  const int mine = t->size;
  if (h.magic != 0x444C534Fu || h.max_depth != t->tp.D || h.size_before != mine || h.size_after < h.size_before ||
      h.n_touched < 0 || (h.size_after & 7))
    return OSL_ERR_INVALID;  // not the successor state of this replica
  const size_t appended = (size_t)(h.size_after - h.size_before);
  if (bytes != sizeof(h) + (size_t)h.n_touched * sizeof(uint3) + appended * 8) return OSL_ERR_INVALID;
  if ((size_t)h.size_after > t->cap_nodes) {
    OSL_CUDA(cudaDeviceSynchronize());
    rc = osl_grow_pool(t, (size_t)h.size_after, 0);
    if (rc) return rc;
  }
  const unsigned char* p = static_cast<const unsigned char*>(d_buf);
  if (appended)
    OSL_CUDA(cudaMemcpyAsync(t->d_pool + 2 * (size_t)h.size_before, p + sizeof(h) + (size_t)h.n_touched * sizeof(uint3),
                             appended * 8, cudaMemcpyDeviceToDevice, st));
  int* d_bad = reinterpret_cast<int*>(t->d_scan_totals + OSL_NCOUNT(OSL_MAXD) + 2);  // scratch word, zero at rest
  if (h.n_touched) {
    k_delta_apply<<<(h.n_touched + 255) / 256, 256, 0, st>>>(t->d_pool, reinterpret_cast<const uint3*>(p + sizeof(h)),
                                                            h.n_touched, h.size_after, d_bad);
    OSL_LAUNCHED(1);
  }
  int bad = 0;
  OSL_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaStreamSynchronize(st));
  if (bad) {
    OSL_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    return OSL_ERR_INVALID;
  }
  t->upload_count++;
  return osl_set_device_size(t, h.size_after);
}

}  // extern "C"

// ---- sharded build (osl_shard_analyze / osl_shard_assign in osl_integrate.cu) ------------------------------------------
// What a rank changed in a sharded build: the nodes of its level lists only.  The tiles it allocated are scattered
// through the appended range (every pass interleaves the ranks), so the receiver re-creates them from the child
// pointers: a triple whose word0 points at or beyond `size_before` names a tile that did not exist -- its 8 value words
// become "empty" (svo.cu:272-275) before the triples of the touched children are written.
__global__ void k_delta_init_tiles(u32* __restrict__ pool, const uint3* __restrict__ in, int n, u32 size_before, int limit, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint3 e = in[i];
  if (!(e.y & OSL_FLAG)) return;
  const u32 tile = e.y & OSL_MASK;
  if (tile < size_before) return;
  if ((tile & 7u) || (long long)tile + 8 > (long long)limit) { atomicAdd(bad, 1); return; }
  // (only words that are still zero: this rank may already have written a child of its own into the other rank's
  // tile -- the first key of a slice continues below nodes the lower rank split -- and no initialised value is 0)
#pragma unroll
  for (int c = 0; c < 8; c++)
    if (pool[2 * (size_t)(tile + c) + 1] == 0u) pool[2 * (size_t)(tile + c) + 1] = OSL_EMPTY;
}

// The nodes on the paths of the slices' first keys are the only ones with touched children in two ranks.  One CTA: thread b
// walks key b down the FINAL tree (every rank has applied every delta), then the paths are re-averaged level by level,
// deepest first (averageChildren, svo.cu:384-441; shared ancestors get the same value from every thread that owns them);
// thread 0 finally writes the root average into node 0's value word (quirk Q6) exactly as k_levels does on one GPU.
#define FIX_MAX 64
__global__ void k_shard_fixup(u32* __restrict__ pool, const u64* __restrict__ keys, int n_keys, int D, int any_touched) {
  __shared__ u32 s_path[FIX_MAX][OSL_MAXD + 1];
  const int b = threadIdx.x;
  if (b < n_keys) {
    const u64 key = keys[b];
    u32 node = (u32)((key >> (3 * (D - 1))) & 7ull);
    for (int d = 1; d <= D; d++) {
      s_path[b][d] = node;
      const u32 w0 = pool[2 * (size_t)node];
      if (d == D || !(w0 & OSL_FLAG)) {
        for (int q = d + 1; q <= D; q++) s_path[b][q] = 0xFFFFFFFFu;
        break;
      }
      node = (w0 & OSL_MASK) + (u32)((key >> (3 * (D - d - 1))) & 7ull);
    }
  }
  __syncthreads();
  for (int d = D - 1; d >= 1; d--) {
    if (b < n_keys) {
      const u32 node = s_path[b][d];
      if (node != 0xFFFFFFFFu) {
        const u32 w0 = pool[2 * (size_t)node];
        if (w0 & OSL_FLAG) {
          const u32* tile = pool + 2 * (size_t)(w0 & OSL_MASK);
          u32 v[8];
#pragma unroll
          for (int c = 0; c < 8; c++) v[c] = ((volatile const u32*)tile)[2 * c + 1];
          ((volatile u32*)pool)[2 * (size_t)node + 1] = osl_average8(v);
        }
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (b == 0 && any_touched) {
    u32 v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = ((volatile const u32*)pool)[2 * c + 1];
    ((volatile u32*)pool)[1] = osl_average8(v);
  }
}

__global__ void k_keys_of(const float* __restrict__ pts, const int* __restrict__ idx, int n, TreeParams tp, u64* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = pts + 4 * (size_t)idx[i];
  u64 k;
  osl_key(q[0], q[1], q[2], tp, k);
  out[i] = k;
}

extern "C" {

size_t osl_shard_delta_bytes(osl_svo* t) {
  if (!t) return 0;
  cudaSetDevice(t->device);
  if (osl_poll_results(t, true) != OSL_OK || t->seq == 0) return 0;
  const FrameState& F = t->h_ring[(t->seq - 1) % OSL_RING];
  size_t touched = 0;
  for (int d = 1; d <= t->tp.D; d++) touched += (size_t)F.n_level[d];
  return sizeof(DeltaHeader) + touched * sizeof(uint3);
}

// this rank's changes of the last osl_shard_assign: header + (index, word0, word1) of its level-list nodes
osl_status osl_shard_delta_pack(osl_svo* t, void* d_buf, size_t cap, size_t* bytes, void* stream) {
  if (!t || !d_buf || !bytes) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  if (t->seq == 0) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long f = t->seq - 1;
  const FrameState& F = t->h_ring[f % OSL_RING];
  const LevelArrays& lv = t->lv[f % OSL_BACK];
  const int D = t->tp.D;
  if (F.overflow) return OSL_ERR_POOL_OVERFLOW;
  DeltaHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = 0x534C534Fu;  // 'OSLS': sparse
  h.max_depth = D; h.size_before = F.size_before; h.size_after = F.size_after;
  size_t touched = 0;
  for (int d = 1; d <= D; d++) touched += (size_t)F.n_level[d];
  h.n_touched = (int)touched;
  const size_t need = sizeof(h) + touched * sizeof(uint3);
  *bytes = need;
  if (need > cap) return OSL_ERR_INVALID;
  unsigned char* p = static_cast<unsigned char*>(d_buf);
  OSL_CUDA(cudaMemcpyAsync(p, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  size_t done = 0;
  for (int d = 1; d <= D; d++) {
    const int n = F.n_level[d];
    if (n <= 0) continue;
    k_delta_pack<<<(n + 255) / 256, 256, 0, st>>>(t->d_pool, lv.self + lv.off[d], n,
                                                  reinterpret_cast<uint3*>(p + sizeof(h)) + done);
    OSL_LAUNCHED(1);
    done += (size_t)n;
  }
  OSL_CUDA(cudaStreamSynchronize(st));
  return OSL_OK;
}

// another rank's changes of the same sharded build (this rank has run its own osl_shard_assign: same sizes)
osl_status osl_shard_delta_apply(osl_svo* t, const void* d_buf, size_t bytes, void* stream) {
  if (!t || !d_buf || bytes < sizeof(DeltaHeader)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DeltaHeader h;
  OSL_CUDA(cudaMemcpyAsync(&h, d_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaStreamSynchronize(st));
  if (h.magic != 0x534C534Fu || h.max_depth != t->tp.D || h.size_after != t->size || h.n_touched < 0 ||
      bytes != sizeof(h) + (size_t)h.n_touched * sizeof(uint3))
    return OSL_ERR_INVALID;
  if (h.n_touched == 0) return OSL_OK;
  const uint3* tri = reinterpret_cast<const uint3*>(static_cast<const unsigned char*>(d_buf) + sizeof(h));
  int* d_bad = reinterpret_cast<int*>(t->d_scan_totals + OSL_NCOUNT(OSL_MAXD) + 2);  // scratch word, zero at rest
  const int blocks = (h.n_touched + 255) / 256;
  k_delta_init_tiles<<<blocks, 256, 0, st>>>(t->d_pool, tri, h.n_touched, (u32)(h.size_before > 8 ? h.size_before : 8),
                                             h.size_after, d_bad);
  k_delta_apply<<<blocks, 256, 0, st>>>(t->d_pool, tri, h.n_touched, h.size_after, d_bad);
  OSL_LAUNCHED(2);
  int bad = 0;
  OSL_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  OSL_CUDA(cudaStreamSynchronize(st));
  if (bad) {
    OSL_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    return OSL_ERR_INVALID;
  }
  t->upload_count++;
  return OSL_OK;
}

// h_starts: grid index of the first voxel of ranks 1 .. n_ranks-1 (the slices' boundaries); the grid is on every rank
osl_status osl_shard_fixup(osl_svo* t, const float* d_centers4, int n_total, const int* h_starts, int n_bounds, void* stream) {
  if (!t || !d_centers4 || n_total < 0 || n_bounds < 0 || n_bounds > FIX_MAX || (n_bounds > 0 && !h_starts)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int* d_idx = nullptr;
  u64* d_keys = nullptr;
  int starts[FIX_MAX];
  int nb = 0;
  for (int i = 0; i < n_bounds; i++)
    if (h_starts[i] > 0 && h_starts[i] < n_total) starts[nb++] = h_starts[i];  // (an empty slice has no first key)
  OSL_CUDA(cudaMalloc(&d_idx, sizeof(int) * FIX_MAX));
  cudaError_t e = cudaMalloc(&d_keys, sizeof(u64) * FIX_MAX);
  if (e == cudaSuccess && nb) e = cudaMemcpyAsync(d_idx, starts, sizeof(int) * nb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && nb) {
    k_keys_of<<<1, FIX_MAX, 0, st>>>(d_centers4, d_idx, nb, t->tp, d_keys);
    OSL_LAUNCHED(1);
  }
  if (e == cudaSuccess) {
    k_shard_fixup<<<1, FIX_MAX, 0, st>>>(t->d_pool, d_keys, nb, t->tp.D, n_total > 0 ? 1 : 0);
    OSL_LAUNCHED(1);
    e = cudaStreamSynchronize(st);
  }
  cudaFree(d_idx);
  cudaFree(d_keys);
  OSL_CUDA(e);
  t->upload_count++;
  return OSL_OK;
}

}  // extern "C"
