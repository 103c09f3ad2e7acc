// osl_capi.cu -- the C ABI of libosl_b200.so (include/osl_b200.h): object lifetime, memory, entry points.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "osl_internal.cuh"

int g_osl_last_cuda_error = 0;
long long g_osl_launches = 0;

// wait for every frame in flight and fold its result block into the host-side state
static osl_status drain(osl_svo* t) {
  osl_status rc = osl_poll_results(t, true);
  if (rc) return rc;
  return t->sticky_error;
}

osl_status osl_set_device_size(osl_svo* t, int size) {
  // FrameState::cur_size lives on the device (frames are planned there without a host round trip)
  OSL_CUDA(cudaMemcpy(&t->d_fs->cur_size, &size, sizeof(int), cudaMemcpyHostToDevice));
  if (t->d_wcache) OSL_CUDA(cudaMemset(t->d_wcache, 0xFF, 4096 * sizeof(u64)));  // node indices / depths change meaning
  t->size = size;
  return OSL_OK;
}

// One layer of map growth (osl_svo_expand): the root tile (nodes 0-7) is pushed one level down.  Thread 8*i+k owns
// child k of the new tile of root child i; the tile is nodes [size + 8*i, size + 8*i + 8).  Single CTA of 64 threads.
__global__ void k_expand_root(u32* __restrict__ pool, int size) {
  __shared__ u32 s_val[64];
  const int i = threadIdx.x >> 3, k = threadIdx.x & 7;
  const uint2 old = reinterpret_cast<const uint2*>(pool)[i];
  __syncthreads();  // every old root child is in registers before any is overwritten
  const bool moved = k == 7 - i;  // oppositeNode(i): the octant of new child i that touches the centre
  const uint2 w = moved ? old : make_uint2(0u, 127u << 24);
  reinterpret_cast<uint2*>(pool)[size + threadIdx.x] = w;
  s_val[threadIdx.x] = w.y;
  __syncthreads();
  if (k == 0) {  // averageChildren of the new tile (all eight counted, Q5; alpha = max)
    u32 r = 0, g = 0, b = 0, a = 0;
    for (int c = 0; c < 8; c++) {
      const u32 v = s_val[8 * i + c];
      r += v & 0xFF; g += (v >> 8) & 0xFF; b += (v >> 16) & 0xFF;
      a = max(a, v >> 24);
    }
    reinterpret_cast<uint2*>(pool)[i] =
        make_uint2((1u << 30) | (u32)(size + 8 * i), (r >> 3) | ((g >> 3) << 8) | ((b >> 3) << 16) | (a << 24));
  }
}

// Structural check of an uploaded pool (osl_svo_upload / osl_svo_load): every child pointer must name an 8-aligned
// tile inside the pool and beyond the root tile, or raycast / extraction / k_structure would read out of bounds.
__global__ void k_validate_pool(const u32* __restrict__ pool, int n, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 w0 = pool[2 * (size_t)i];
  if (w0 & OSL_FLAG) {
    const u32 c = w0 & OSL_MASK;
    if ((c & 7u) || c < 8u || (unsigned long long)c + 8ull > (unsigned long long)n) atomicAdd(bad, 1);
  }
}

// Runs k_validate_pool over the first n_nodes nodes; corrupt child pointers leave an EMPTY, valid tree behind.
osl_status osl_validate_pool(osl_svo* t, int n_nodes) {
  if (n_nodes <= 0) return OSL_OK;
  int* d_bad = reinterpret_cast<int*>(t->d_scan_totals + OSL_NCOUNT(OSL_MAXD) + 2);  // scratch word, zero at rest
  int bad = 0;
  k_validate_pool<<<(n_nodes + 255) / 256, 256>>>(t->d_pool, n_nodes, d_bad);
  g_osl_launches++;
  OSL_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  if (bad) {
    OSL_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    OSL_CUDA(cudaMemset(t->d_pool, 0, (size_t)n_nodes * 8));
    t->upload_count++;
    osl_set_device_size(t, 0);
    return OSL_ERR_INVALID;
  }
  return OSL_OK;
}

extern "C" {

const char* osl_version(void) { return "osl_b200 0.6 (sm_100a)"; }
int osl_frame_result_bytes(void) { return (int)sizeof(FrameState); }
int osl_last_cuda_error(void) { return g_osl_last_cuda_error; }
int64_t osl_launch_count(void) { return g_osl_launches; }

const char* osl_status_string(osl_status s) {
  switch (s) {
    case OSL_OK: return "ok";
    case OSL_ERR_INVALID: return "invalid argument";
    case OSL_ERR_CUDA: return "CUDA error";
    case OSL_ERR_OOM: return "out of device memory";
    case OSL_ERR_POOL_OVERFLOW: return "node pool overflow (2^30 nodes)";
    case OSL_ERR_UNSUPPORTED: return "unsupported";
  }
  return "unknown";
}

osl_status osl_svo_create(osl_svo** out, const float center[3], float half_edge, int max_depth, size_t reserve_nodes,
                          int device) {
  if (!out || !center || max_depth < 1 || max_depth > OSL_MAX_DEPTH || !(half_edge > 0.0f)) return OSL_ERR_INVALID;
  *out = nullptr;
  OSL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  OSL_CUDA(cudaGetDeviceProperties(&prop, device));
  if (!prop.cooperativeLaunch) return OSL_ERR_UNSUPPORTED;
  osl_svo* t = new osl_svo();
  memset(t, 0, sizeof(*t));
  t->device = device;
  t->tp.cx = center[0]; t->tp.cy = center[1]; t->tp.cz = center[2];
  t->tp.half = half_edge;
  t->tp.D = max_depth;
  t->tp.quirks = 1;
  t->num_sms = prop.multiProcessorCount;
  if (osl_integrate_init(t) != OSL_OK) { delete t; return OSL_ERR_CUDA; }  // dynamic shared-memory opt-ins first
  const int occ_sort = osl_sort_occupancy(), occ_str = osl_structure_occupancy(), occ_lvl = osl_levels_occupancy();
  if (occ_sort < 1 || occ_str < 1 || occ_lvl < 1) { delete t; return OSL_ERR_CUDA; }
  t->sort_grid = (occ_sort > 4 ? 4 : occ_sort) * t->num_sms;
  t->structure_grid = (occ_str > 4 ? 4 : occ_str) * t->num_sms;
  t->levels_grid = (occ_lvl > 4 ? 4 : occ_lvl) * t->num_sms;
  { const int ob = osl_structure_big_occupancy(); t->structure_big_grid = (ob > 2 ? 2 : ob) * t->num_sms; }
  osl_status rc = OSL_OK;
  t->hint_emit = t->hint_level = -1;
  do {
    bool okh = true;
    for (int f = 0; f < OSL_FRONT && okh; f++)
      okh = cudaMalloc(&t->d_cta_hist[f], (size_t)t->sort_grid * 256 * sizeof(u32)) == cudaSuccess;
    if (!okh) { rc = OSL_ERR_OOM; break; }
    if (cudaMalloc(&t->d_fs, sizeof(FrameState) * (1 + OSL_BACK)) != cudaSuccess) { rc = OSL_ERR_OOM; break; }
    if (cudaMemset(t->d_fs, 0, sizeof(FrameState) * (1 + OSL_BACK)) != cudaSuccess) { rc = OSL_ERR_CUDA; break; }
    if (cudaMalloc(&t->d_scan_totals, (OSL_NCOUNT(OSL_MAXD) + 8) * sizeof(u32)) != cudaSuccess) { rc = OSL_ERR_OOM; break; }
    if (cudaMemset(t->d_scan_totals, 0, (OSL_NCOUNT(OSL_MAXD) + 8) * sizeof(u32)) != cudaSuccess) { rc = OSL_ERR_CUDA; break; }
    if (cudaMallocHost(&t->h_ring, sizeof(FrameState) * OSL_RING) != cudaSuccess) { rc = OSL_ERR_OOM; break; }
    memset(t->h_ring, 0, sizeof(FrameState) * OSL_RING);
    if (cudaMalloc(&t->d_ready, OSL_STAGES * sizeof(u32)) != cudaSuccess) { rc = OSL_ERR_OOM; break; }
    if (cudaMemset(t->d_ready, 0, OSL_STAGES * sizeof(u32)) != cudaSuccess) { rc = OSL_ERR_CUDA; break; }
    if (cudaMallocHost(&t->h_ready_vals, 64 * sizeof(u32)) != cudaSuccess) { rc = OSL_ERR_OOM; break; }
    t->fused_enabled = getenv("OSL_NO_FUSED") ? 0 : 1;
    bool ok = true;
    for (int i = 0; i < OSL_RING && ok; i++)
      ok = cudaEventCreateWithFlags(&t->ring_ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < OSL_STAGES && ok; i++)
      ok = cudaEventCreateWithFlags(&t->stage_copied[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&t->stage_free[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < OSL_FRONT && ok; i++)
      ok = cudaEventCreateWithFlags(&t->emit_done[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&t->sort_done[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < OSL_RING && ok; i++)
      ok = cudaEventCreateWithFlags(&t->struct_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 4 && ok; i++) ok = cudaStreamCreateWithFlags(&t->pipe[i], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { rc = OSL_ERR_CUDA; break; }
    rc = osl_grow_pool(t, reserve_nodes ? reserve_nodes : ((size_t)1 << 20), 0);
    if (rc) break;
    if (cudaMemset(t->d_pool, 0, 64) != cudaSuccess) { rc = OSL_ERR_CUDA; break; }
  } while (0);
  if (rc) { osl_svo_destroy(t); return rc; }
  *out = t;
  return OSL_OK;
}

void osl_svo_destroy(osl_svo* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  if (t->counted_piped && g_osl_piped_trees > 0) g_osl_piped_trees--;
  cudaFree(t->d_pool);
  for (int f = 0; f < OSL_FRONT; f++) {
    cudaFree(t->d_keysA[f]); cudaFree(t->d_keysB[f]); cudaFree(t->d_payA[f]); cudaFree(t->d_payB[f]);
    cudaFree(t->d_cta_hist[f]); cudaFree(t->d_bkeys[f]); cudaFree(t->d_bpay[f]);
    if (t->emit_done[f]) cudaEventDestroy(t->emit_done[f]);
    if (t->sort_done[f]) cudaEventDestroy(t->sort_done[f]);
  }
  for (int b = 0; b < OSL_BACK; b++) cudaFree(t->d_level_mem[b]);
  for (int i = 0; i < OSL_RING; i++)
    if (t->struct_ev[i]) cudaEventDestroy(t->struct_ev[i]);
  for (int i = 0; i < 4; i++)
    if (t->pipe[i]) cudaStreamDestroy(t->pipe[i]);
  cudaFree(t->d_m); cudaFree(t->d_s); cudaFree(t->d_blockcnt);
  cudaFree(t->d_keysC); cudaFree(t->d_payC); cudaFree(t->d_split); cudaFree(t->d_wcache); cudaFree(t->d_blockcnt_tot); cudaFree(t->d_start); cudaFree(t->d_flags);
  cudaFree(t->d_scan_totals); cudaFree(t->d_fs);
  for (int f = 0; f < OSL_FRONT; f++) osl_sort_big_free(&t->sort_ws[f]);
  cudaFree(t->ex_kA); cudaFree(t->ex_kB); cudaFree(t->ex_nA); cudaFree(t->ex_nB); cudaFree(t->ex_status); cudaFree(t->ex_cnt);
  for (int i = 0; i < OSL_STAGES; i++) {
    cudaFree(t->d_depth_stage[i]); cudaFree(t->d_rgb_stage[i]);
    if (t->stage_copied[i]) cudaEventDestroy(t->stage_copied[i]);
    if (t->stage_free[i]) cudaEventDestroy(t->stage_free[i]);
  }
  for (int i = 0; i < OSL_RING; i++)
    if (t->ring_ev[i]) cudaEventDestroy(t->ring_ev[i]);
  for (int i = 0; i < 5; i++)
    if (t->stage_ev[i]) cudaEventDestroy(t->stage_ev[i]);
  if (t->pose_ev) cudaEventDestroy(t->pose_ev);
  if (t->pose_read_ev) cudaEventDestroy(t->pose_read_ev);
  if (t->reader_ev) cudaEventDestroy(t->reader_ev);
  if (t->copy_stream) cudaStreamDestroy(t->copy_stream);
  if (t->h_ring) cudaFreeHost(t->h_ring);
  if (t->h_ready_vals) cudaFreeHost(t->h_ready_vals);
  cudaFree(t->d_ready);
  delete t;
}

osl_status osl_svo_reset(osl_svo* t) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_poll_results(t, true);
  OSL_CUDA(cudaDeviceSynchronize());
  OSL_CUDA(cudaMemset(t->d_pool, 0, (size_t)(t->size > 8 ? t->size : 8) * 8));  // pool beyond the live nodes stays zero
  t->sticky_error = OSL_OK;
  t->upload_count++;  // invalidates the cached extraction frontier
  memset(&t->counters, 0, sizeof(t->counters));
  return osl_set_device_size(t, 0);
}

osl_status osl_svo_expand(osl_svo* t, int layers) {
  if (!t || layers < 1) return OSL_ERR_INVALID;
  if (t->tp.D + layers > OSL_MAX_DEPTH) return OSL_ERR_UNSUPPORTED;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = drain(t);
  if (rc) return rc;
  OSL_CUDA(cudaDeviceSynchronize());
  for (int l = 0; l < layers; l++) {
    if (t->size > 0) {  // an empty tree only changes its geometry
      if ((size_t)t->size + 64 > ((size_t)1 << 30)) return OSL_ERR_POOL_OVERFLOW;
      rc = osl_grow_pool(t, (size_t)t->size + 64, 0);
      if (rc) return rc;
      k_expand_root<<<1, 64>>>(t->d_pool, t->size);
      OSL_CUDA(cudaGetLastError());
      g_osl_launches++;
      rc = osl_set_device_size(t, t->size + 64);
      if (rc) return rc;
    }
    t->tp.half *= 2.0f;
    t->tp.D += 1;
  }
  OSL_CUDA(cudaDeviceSynchronize());
  t->upload_count++;        // invalidates the cached extraction frontier
  t->hint_emit = t->hint_level = -1;
  osl_drop_workspace(t);    // counter vectors and level arrays are laid out per max_depth
  return osl_reset_splitters(t);
}

int osl_svo_max_depth(const osl_svo* t) { return t ? t->tp.D : 0; }

// Per-kernel timing of the integrate pipeline (k_emit, k_sort, k_structure, k_levels) with CUDA events on the
// caller's stream; only non-pipelined frames are timed.  osl_get_stage_times waits for the last timed frame.
osl_status osl_svo_set_stage_timing(osl_svo* t, int enabled) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  if (enabled && !t->stage_ev[0])
    for (int i = 0; i < 5; i++) OSL_CUDA(cudaEventCreate(&t->stage_ev[i]));
  t->stage_timing = enabled ? 1 : 0;
  t->stage_valid = 0;
  return OSL_OK;
}

osl_status osl_get_stage_times(osl_svo* t, float ms[4]) {
  if (!t || !ms || !t->stage_valid) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  OSL_CUDA(cudaEventSynchronize(t->stage_ev[4]));
  for (int i = 0; i < 4; i++) OSL_CUDA(cudaEventElapsedTime(&ms[i], t->stage_ev[i], t->stage_ev[i + 1]));
  return OSL_OK;
}

osl_status osl_svo_set_pipeline(osl_svo* t, int enabled) {
  if (!t) return OSL_ERR_INVALID;
  t->pipeline = enabled ? 1 : 0;
  return OSL_OK;
}

osl_status osl_svo_set_quirks(osl_svo* t, int ref_quirks) {
  if (!t) return OSL_ERR_INVALID;
  t->tp.quirks = (ref_quirks & 1) ? 1 : 0;
  t->force_grid_sort = (ref_quirks & 2) ? 1 : 0;  // bit 1 (testing): never use the bucket sort
  t->zero_copy_rgb = (ref_quirks & 4) ? 1 : 0;    // bit 2 (measurement): read pinned colour planes in place
  if (ref_quirks & 8) t->fused_enabled = 0;       // bit 3 (testing / measurement): pipelined frames as four kernels
  return OSL_OK;
}

osl_status osl_integrate_depth(osl_svo* t, const uint16_t* d_depth, const uint8_t* d_rgb, int w, int h, float fx,
                               float fy, const float pose[16], void* stream) {
  if (!t || !d_depth || !d_rgb || w <= 0 || h <= 0 || !pose || (long long)w * h >= (1ll << 30)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  EmitParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.depth = d_depth; ep.rgb = d_rgb; ep.w = w; ep.h = h; ep.fx = fx; ep.fy = fy;
  memcpy(ep.M, pose, sizeof(ep.M));
  ep.n = w * h; ep.mode = 0;
  return osl_run_integrate(t, ep, nullptr, (cudaStream_t)stream, nullptr);
}

// The pose lives in device memory and is read when k_emit runs: lets a tracker's result drive the integration of
// the same frame without a host round trip (osl_tracker_pose_device).  Ordered after the work already queued on
// `stream` (in pipelined mode the library's emit stream waits for it).
osl_status osl_integrate_depth_posed(osl_svo* t, const uint16_t* d_depth, const uint8_t* d_rgb, int w, int h, float fx,
                                     float fy, const float* d_pose, void* stream) {
  if (!t || !d_depth || !d_rgb || w <= 0 || h <= 0 || !d_pose || (long long)w * h >= (1ll << 30)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  EmitParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.depth = d_depth; ep.rgb = d_rgb; ep.w = w; ep.h = h; ep.fx = fx; ep.fy = fy;
  ep.M_dev = d_pose;
  ep.n = w * h; ep.mode = 0;
  return osl_run_integrate(t, ep, nullptr, (cudaStream_t)stream, nullptr);
}

static osl_status ensure_stage(osl_svo* t, size_t n) {
  if (n <= t->stage_cap) return OSL_OK;
  {  // frames in flight read the slots (the k_frame path also has stages that are not even launched yet)
    osl_status prc = osl_poll_results(t, true);
    if (prc) return prc;
  }
  OSL_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < OSL_STAGES; i++) {
    cudaFree(t->d_depth_stage[i]); cudaFree(t->d_rgb_stage[i]);
    t->d_depth_stage[i] = nullptr; t->d_rgb_stage[i] = nullptr;
  }
  t->stage_cap = 0;
  for (int i = 0; i < OSL_STAGES; i++) {
    OSL_CUDA(cudaMalloc(&t->d_depth_stage[i], n * 2));
    OSL_CUDA(cudaMalloc(&t->d_rgb_stage[i], n * 3));
  }
  t->stage_cap = n;
  t->stage_seq = 0;
  memset(t->stage_frame, 0, sizeof(t->stage_frame));
  return OSL_OK;
}

// Host frames: the H2D copies run on an internal copy stream into one of OSL_STAGES device slots, so the transfer of
// frame f+1 overlaps the kernels of frame f; the integrate itself is stream-ordered on `stream`.  Returns without
// waiting for the device (pinned source buffers make the copies truly asynchronous; pageable ones are staged by the
// driver).  Pinned source buffers must stay untouched until the frame completes (osl_svo_sync / osl_svo_size /
// osl_get_counters / osl_svo_view synchronize).
osl_status osl_integrate_depth_host(osl_svo* t, const uint16_t* h_depth, const uint8_t* h_rgb, int w, int h, float fx,
                                    float fy, const float pose[16], void* stream) {
  if (!t || !h_depth || !h_rgb || w <= 0 || h <= 0 || !pose || (long long)w * h >= (1ll << 30)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  const size_t n = (size_t)w * h;
  osl_status rc = ensure_stage(t, n);
  if (rc) return rc;
  EmitParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.w = w; ep.h = h; ep.fx = fx; ep.fy = fy;  // (depth / rgb: the staging slot osl_run_integrate picks)
  memcpy(ep.M, pose, sizeof(ep.M));
  ep.n = w * h; ep.mode = 0;
  const HostFrame hf = {h_depth, h_rgb};
  return osl_run_integrate(t, ep, nullptr, (cudaStream_t)stream, &hf);  // always pipelined: the library owns the copies
}

osl_status osl_integrate_points(osl_svo* t, const float* d_xyz, const uint8_t* d_rgb, int n, void* stream) {
  if (!t || n < 0 || (n > 0 && (!d_xyz || !d_rgb))) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  EmitParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.pts = d_xyz; ep.stride = 3; ep.rgb = d_rgb; ep.n = n; ep.mode = 1;
  return osl_run_integrate(t, ep, nullptr, (cudaStream_t)stream, nullptr);
}

osl_status osl_integrate_voxels(osl_svo* t, const float* d_centers4, const float* d_colors4, int n, void* stream) {
  if (!t || n < 0 || (n > 0 && (!d_centers4 || !d_colors4))) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  EmitParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.pts = d_centers4; ep.stride = 4; ep.n = n; ep.mode = 2;
  return osl_run_integrate(t, ep, d_colors4, (cudaStream_t)stream, nullptr);
}

osl_status osl_svo_join(osl_svo* t, void* stream) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_join(t, (cudaStream_t)stream);
  if (rc) return rc;
  return osl_note_foreign_reader(t, (cudaStream_t)stream);  // foreign readers on `stream` finish before later frames write
}

osl_status osl_svo_sync(osl_svo* t) {
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  return drain(t);
}

osl_status osl_svo_view(const osl_svo* tc, const uint32_t** d_pool, int* n_nodes, float center[3], float* half_edge) {
  osl_svo* t = const_cast<osl_svo*>(tc);
  if (!t) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = drain(t);
  if (d_pool) *d_pool = t->d_pool;
  if (n_nodes) *n_nodes = t->size;
  if (center) { center[0] = t->tp.cx; center[1] = t->tp.cy; center[2] = t->tp.cz; }
  if (half_edge) *half_edge = t->tp.half;
  return rc;
}

int osl_svo_size(const osl_svo* tc) {
  osl_svo* t = const_cast<osl_svo*>(tc);
  if (!t) return 0;
  cudaSetDevice(t->device);
  osl_poll_results(t, true);
  return t->size;
}

osl_status osl_svo_download(const osl_svo* tc, uint32_t* h_pool, int cap_nodes) {
  osl_svo* t = const_cast<osl_svo*>(tc);
  if (!t || !h_pool) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = drain(t);
  if (rc) return rc;
  if (cap_nodes < t->size) return OSL_ERR_INVALID;
  OSL_CUDA(cudaDeviceSynchronize());
  if (t->size > 0) OSL_CUDA(cudaMemcpy(h_pool, t->d_pool, (size_t)t->size * 8, cudaMemcpyDeviceToHost));
  return OSL_OK;
}

osl_status osl_svo_upload(osl_svo* t, const uint32_t* h_pool, int n_nodes) {
  if (!t || n_nodes < 0 || (n_nodes > 0 && !h_pool)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_poll_results(t, true);
  OSL_CUDA(cudaDeviceSynchronize());
  const size_t old = (size_t)(t->size > 8 ? t->size : 8);
  OSL_CUDA(cudaMemset(t->d_pool, 0, old * 8));  // keep the pool beyond the live nodes zero
  t->size = 0;
  osl_status rc = osl_grow_pool(t, (size_t)(n_nodes > 8 ? n_nodes : 8), 0);
  if (rc) return rc;
  if (n_nodes > 0) {
    if (n_nodes < 8 || (n_nodes & 7)) return OSL_ERR_INVALID;  // whole 8-node tiles, the root's first
    OSL_CUDA(cudaMemcpy(t->d_pool, h_pool, (size_t)n_nodes * 8, cudaMemcpyHostToDevice));
    osl_status vrc = osl_validate_pool(t, n_nodes);
    if (vrc) return vrc;
  }
  t->sticky_error = OSL_OK;
  t->upload_count++;  // invalidates the cached extraction frontier
  return osl_set_device_size(t, n_nodes);
}

// Checkpoint / resume (the reference has none; its de-facto wire format is the flat 2*n uint
// array OctreeNode::pullToCPU/pushToGPU exchange, octree.cpp:41-169).  File = 32-byte header + that array.
struct osl_file_header {
  char magic[8];      // "OSLSVO1\0"
  int32_t max_depth;
  int32_t n_nodes;
  float center[3];
  float half_edge;
};

osl_status osl_svo_save(const osl_svo* tc, const char* path) {
  osl_svo* t = const_cast<osl_svo*>(tc);
  if (!t || !path) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = drain(t);
  if (rc) return rc;
  OSL_CUDA(cudaDeviceSynchronize());
  osl_file_header h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "OSLSVO1", 8);
  h.max_depth = t->tp.D; h.n_nodes = t->size;
  h.center[0] = t->tp.cx; h.center[1] = t->tp.cy; h.center[2] = t->tp.cz; h.half_edge = t->tp.half;
  FILE* f = fopen(path, "wb");
  if (!f) return OSL_ERR_INVALID;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  const size_t chunk_nodes = (size_t)1 << 22;  // 32 MB staging
  uint32_t* buf = (uint32_t*)malloc(chunk_nodes * 8);
  if (!buf) { fclose(f); return OSL_ERR_OOM; }
  for (size_t off = 0; ok && off < (size_t)t->size; off += chunk_nodes) {
    const size_t cnt = ((size_t)t->size - off < chunk_nodes) ? (size_t)t->size - off : chunk_nodes;
    if (cudaMemcpy(buf, t->d_pool + 2 * off, cnt * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
    ok = fwrite(buf, 8, cnt, f) == cnt;
  }
  free(buf);
  ok = (fclose(f) == 0) && ok;
  return ok ? OSL_OK : OSL_ERR_INVALID;
}

// Loads a checkpoint into an existing tree; max_depth, centre and half edge must match the tree's (they define the
// meaning of every node index).  `t` may hold anything before: it is replaced.
osl_status osl_svo_load(osl_svo* t, const char* path) {
  if (!t || !path) return OSL_ERR_INVALID;
  FILE* f = fopen(path, "rb");
  if (!f) return OSL_ERR_INVALID;
  osl_file_header h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "OSLSVO1", 8) != 0 || h.n_nodes < 0 || h.max_depth != t->tp.D ||
      h.center[0] != t->tp.cx || h.center[1] != t->tp.cy || h.center[2] != t->tp.cz || h.half_edge != t->tp.half) {
    fclose(f);
    return OSL_ERR_INVALID;
  }
  {  // the header must agree with the file before anything is allocated from it
    long here = ftell(f), end = -1;
    if (here >= 0 && fseek(f, 0, SEEK_END) == 0) end = ftell(f);
    if (here < 0 || end < 0 || fseek(f, here, SEEK_SET) != 0 || (long long)(end - here) != 8ll * h.n_nodes) {
      fclose(f);
      return OSL_ERR_INVALID;
    }
  }
  uint32_t* buf = (uint32_t*)malloc((size_t)(h.n_nodes > 0 ? h.n_nodes : 1) * 8);
  if (!buf) { fclose(f); return OSL_ERR_OOM; }
  const bool ok = fread(buf, 8, (size_t)h.n_nodes, f) == (size_t)h.n_nodes;
  fclose(f);
  osl_status rc = ok ? osl_svo_upload(t, buf, h.n_nodes) : OSL_ERR_INVALID;
  free(buf);
  return rc;
}

osl_status osl_get_counters(const osl_svo* tc, osl_counters* out) {
  osl_svo* t = const_cast<osl_svo*>(tc);
  if (!t || !out) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  osl_status rc = osl_poll_results(t, true);
  *out = t->counters;
  return rc;
}

static osl_status raycast_rows_pool(const uint32_t* d_pool, const float center[3], float half_edge,
                                    uint8_t* d_out_rgba, int w, int h, int row0, int rows, int band_h,
                                    int band_stride, float fov_deg,
                                    const float view[16], const osl_raycast_params* prm, osl_raycast_stats* h_stats,
                                    void* stream) {
  if (!d_pool || !center || !d_out_rgba || !view) return OSL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* d_stats = nullptr;
  if (h_stats) {
    OSL_CUDA(cudaMalloc(&d_stats, 16));
    OSL_CUDA(cudaMemsetAsync(d_stats, 0, 16, st));
  }
  osl_status rc = osl_launch_raycast(d_pool, center, half_edge, d_out_rgba, w, h, row0, rows, band_h, band_stride,
                                     fov_deg, view, prm, d_stats, st);
  if (h_stats) {
    unsigned long long s[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(s, d_stats, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_stats);
    if (rc == OSL_OK) OSL_CUDA(e);
    h_stats->rays = (int64_t)w * rows; h_stats->steps = (int64_t)s[0]; h_stats->visits = (int64_t)s[1];
  }
  return rc;
}

osl_status osl_raycast_pool(const uint32_t* d_pool, const float center[3], float half_edge, uint8_t* d_out_rgba, int w,
                            int h, float fov_deg, const float view[16], const osl_raycast_params* prm,
                            osl_raycast_stats* h_stats, void* stream) {
  return raycast_rows_pool(d_pool, center, half_edge, d_out_rgba, w, h, 0, h, h, 1, fov_deg, view, prm, h_stats,
                           stream);
}

// Image rows [row0, row0 + rows) only, into d_out_rgba[0 .. rows*w*4): the multi-GPU decomposition of the raycast
// (rays are independent; every GPU holds the tree).  Stream-ordered after the integrates on the same stream.
osl_status osl_raycast_rows(const osl_svo* t, uint8_t* d_out_rgba, int w, int h, int row0, int rows, float fov_deg,
                            const float view[16], const osl_raycast_params* prm, osl_raycast_stats* h_stats,
                            void* stream) {
  if (!t || (t->size == 0 && t->ring_head == 0)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  const float c[3] = {t->tp.cx, t->tp.cy, t->tp.cz};
  osl_status jr = osl_join(const_cast<osl_svo*>(t), (cudaStream_t)stream);
  if (jr) return jr;
  jr = raycast_rows_pool(t->d_pool, c, t->tp.half, d_out_rgba, w, h, row0, rows, rows > 0 ? rows : 1, 1, fov_deg,
                         view, prm, h_stats, stream);
  if (jr) return jr;
  return osl_note_reader(const_cast<osl_svo*>(t), (cudaStream_t)stream);
}

// All rows one rank owns under the interleaved-band decomposition, in ONE launch: bands of band_h rows are dealt
// round-robin to n_ranks ranks (band k belongs to rank k % n_ranks); the rank's rows are written compactly, in image
// order, to d_out_rgba.  *rows_out (optional) receives the number of rows rendered.
osl_status osl_raycast_bands(const osl_svo* t, uint8_t* d_out_rgba, int w, int h, int band_h, int n_ranks, int rank,
                             float fov_deg, const float view[16], const osl_raycast_params* prm, int* rows_out,
                             void* stream) {
  if (!t || (t->size == 0 && t->ring_head == 0) || band_h < 1 || n_ranks < 1 || rank < 0 || rank >= n_ranks || h <= 0)
    return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  int rows = 0;
  for (int row0 = rank * band_h; row0 < h; row0 += band_h * n_ranks) rows += (h - row0 < band_h) ? (h - row0) : band_h;
  if (rows_out) *rows_out = rows;
  const float c[3] = {t->tp.cx, t->tp.cy, t->tp.cz};
  osl_status jr = osl_join(const_cast<osl_svo*>(t), (cudaStream_t)stream);
  if (jr) return jr;
  jr = raycast_rows_pool(t->d_pool, c, t->tp.half, d_out_rgba, w, h, rank * band_h, rows, band_h, n_ranks, fov_deg,
                         view, prm, nullptr, stream);
  if (jr) return jr;
  return osl_note_reader(const_cast<osl_svo*>(t), (cudaStream_t)stream);
}

// Stream-ordered after the integrates enqueued on the same stream (no host synchronisation needed).
osl_status osl_raycast(const osl_svo* t, uint8_t* d_out_rgba, int w, int h, float fov_deg, const float view[16],
                       const osl_raycast_params* prm, void* stream) {
  if (!t || (t->size == 0 && t->ring_head == 0)) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  const float c[3] = {t->tp.cx, t->tp.cy, t->tp.cz};
  osl_status jr = osl_join(const_cast<osl_svo*>(t), (cudaStream_t)stream);
  if (jr) return jr;
  jr = osl_raycast_pool(t->d_pool, c, t->tp.half, d_out_rgba, w, h, fov_deg, view, prm, nullptr, stream);
  if (jr) return jr;
  return osl_note_reader(const_cast<osl_svo*>(t), (cudaStream_t)stream);
}

osl_status osl_raycast_host(const osl_svo* t, uint8_t* h_out_rgba, int w, int h, float fov_deg, const float view[16],
                            const osl_raycast_params* prm, osl_raycast_stats* stats, void* stream) {
  if (!t || (t->size == 0 && t->ring_head == 0) || !h_out_rgba || w <= 0 || h <= 0) return OSL_ERR_INVALID;
  OSL_CUDA(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  {
    osl_status jr = osl_join(const_cast<osl_svo*>(t), st);
    if (jr) return jr;
  }
  uint8_t* d_out;
  OSL_CUDA(cudaMalloc(&d_out, (size_t)w * h * 4));
  const float c[3] = {t->tp.cx, t->tp.cy, t->tp.cz};
  osl_status rc = osl_raycast_pool(t->d_pool, c, t->tp.half, d_out, w, h, fov_deg, view, prm, stats, stream);
  cudaError_t e = cudaSuccess;
  if (rc == OSL_OK) {
    e = cudaMemcpyAsync(h_out_rgba, d_out, (size_t)w * h * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(d_out);
  if (rc == OSL_OK) OSL_CUDA(e);
  return rc;
}

}  // extern "C"
