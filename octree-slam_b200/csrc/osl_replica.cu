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
