// osl_internal.cuh -- shared definitions of libosl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/osl_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

#define OSL_FLAG 0x40000000u   // word0 bit 30: node has children            (svo.cu:130,269)
#define OSL_MASK 0x3FFFFFFFu   // word0 bits 0-29: first-child node index      (svo.cu:136)
#define OSL_EMPTY 0x7F000000u  // word1 of a freshly split child: alpha 127    (svo.cu:274)
#define OSL_NEWBIT 0x80000000u // (internal) level-array ctile flag: tile allocated this frame
#define OSL_NONE 0xFFu         // (internal) "no frontier": the whole path of a key already exists

#define OSL_MAXD OSL_MAX_DEPTH
#define OSL_RING 8    // per-frame result blocks in flight
#define OSL_STAGES 6  // device staging slots of the *_host entry points (a frame's colours live until its k_levels)
#define OSL_PIPE_DEPTH 3  // frames in flight the pool head-room is sized for
#define OSL_FRONT 3   // key-list slots (k_emit / k_sort / k_structure of three consecutive frames overlap)
#define OSL_BACK 2    // level-list + result-block slots (k_structure of frame f+1 overlaps k_levels of frame f)
#define OSL_BUCKETS 64  // key ranges of the bucket sort (k_sort_bucket)
#define OSL_BUCKET_CAP 2048  // entries a bucket holds in shared memory
#define OSL_NCOUNT(D) ((D) + ((D) + 1) * ((D) + 1))
#define OSL_CLVL(D, d) ((d)-1)
#define OSL_CBKT(D, s, d) ((D) + (s) * ((D) + 1) + (d))

// Per-frame device-side state; copied to pinned host memory after the structure phase.
struct FrameState {
  int acc_valid[OSL_FRONT];  // [key-list slot] accumulated by k_emit (atomicAdd), consumed and zeroed by k_structure
  int acc_emit[OSL_FRONT];   // [key-list slot] entries k_emit appended to the key list
  int acc_unsorted[OSL_FRONT];  // [key-list slot] voxel path: the inputs are NOT (sorted and all valid)
  int acc_tiles[OSL_FRONT];  // (unused)
  int acc_bucket[OSL_FRONT][OSL_BUCKETS];  // [key-list slot] entries k_emit filed under each splitter range
  int n_in;         // inputs
  int n_valid;      // V  (inputs with a valid key)
  int n_emit;       // entries sorted (modes 0/1: after the tile-local de-duplication; mode 2: == n_valid)
  int n_invalid_front;  // voxel path: invalid keys sort to the front in the reference (key 1)
  int n_split;      // S
  int size_before;  // nodes before this frame (>= 8)
  int size_after;
  int overflow;     // 1: size_after exceeds the pool capacity -> nothing was written
  int capacity;     // nodes
  int cur_size;     // nodes in the pool (persistent across frames; 0 = fresh tree)
  int frame_seq;    // frames processed
  int done_flag;    // (pinned host copy only) frame number + 1, stored by the value stage when the frame is complete
  int n_level[OSL_MAXD + 2];                  // n_level[d] = distinct touched nodes at depth d (d = 1..D)
  int pass_count[OSL_MAXD + 1];               // |codes[i]| of reference pass i
  int base[(OSL_MAXD + 1) * (OSL_MAXD + 1)];  // base[s*(D+1)+d]: first global split rank of bucket (frontier s, depth d)
};

struct LevelArrays {   // dense per-level lists of the nodes a frame touches, level d at offset off[d]
  u32* ctile;          // child tile index | OSL_NEWBIT ; 0xFFFFFFFF = none (unsplit leaf)
  u32* par;            // index (in level d-1) of the parent node
  u32* self;           // the node's own index in the pool
  u32* fc;             // index (in level d+1) of the node's FIRST touched child: a node's touched children, and with them
                       // its whole touched subtree, are contiguous ranges of the deeper level lists
  u32* src;            // leaves only: winning input (pixel index / sorted position), indexed by the level-D index
  uint8_t* digit;      // octant of the node inside its parent's tile
  size_t off[OSL_MAXD + 2];
};

struct TreeParams {
  float cx, cy, cz, half;
  int D;
  int quirks;
};

// ---- exact float shapes of the reference (SASS of the reference built with nvcc 12.9) ----------

// image_kernels.cu:24-53: I2F of the integer bracket, FMUL (float)depth, IEEE divide, FMUL 0.001f
__device__ __forceinline__ void osl_vertex(int d, int x, int y, int width, int height, int img_w, int img_h, float fx,
                                           float fy, float& X, float& Y, float& Z) {
  if (d == 0 || d > 15000) {
    X = Y = Z = __int_as_float(0x7f800000);
    return;
  }
  const float fd = (float)d;
  X = __fmul_rn(__fdiv_rn(__fmul_rn((float)((img_w / width) * x - img_w / 2), fd), fx), 0.001f);
  Y = __fmul_rn(__fdiv_rn(__fmul_rn((float)(img_h / 2 - (img_h / height) * y), fd), fy), 0.001f);
  Z = __fmul_rn(fd, 0.001f);
}

// image_kernels.cu:206-215 + glm mat4*vec4: t = FMUL(y,m1); t = FFMA(x,m0,t); u = FFMA(z,m2,m3); FADD(t,u)
__device__ __forceinline__ void osl_transform(const float* __restrict__ M, float& x, float& y, float& z) {
  float o[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    float t = __fmul_rn(y, M[4 + r]);
    t = __fmaf_rn(x, M[0 + r], t);
    float u = __fmaf_rn(z, M[8 + r], M[12 + r]);
    o[r] = __fadd_rn(t, u);
  }
  x = o[0]; y = o[1]; z = o[2];
}

// svo.cu:33-66 computeKey.  Returns false for invalid points (reference key 1; Q1: only x and z are tested).
// key = Morton digits WITHOUT the leading 1 (3*D bits).
__device__ __forceinline__ bool osl_key(float px, float py, float pz, const TreeParams& tp, u64& key) {
  const bool valid = isfinite(px) && isfinite(pz);
  float cx = tp.cx, cy = tp.cy, cz = tp.cz, e = tp.half;
  u64 k = 0;
  for (int i = 0; i < tp.D; i++) {
    const bool bx = px > cx, by = py > cy, bz = pz > cz;
    k = (k << 3) | (u64)((int)bx + 2 * (int)by + 4 * (int)bz);
    e = __fmul_rn(e, 0.5f);
    cx = __fadd_rn(cx, bx ? e : -e);
    cy = __fadd_rn(cy, by ? e : -e);
    cz = __fadd_rn(cz, bz ? e : -e);
  }
  key = k;
  return valid;
}

// svo.cu:366-381 leaf blend; every product is exact in FP32 => integer arithmetic
__device__ __forceinline__ u32 osl_blend_u8(u32 cur, u32 r8, u32 g8, u32 b8) {
  const u32 a = cur >> 24;
  const u32 r = ((256u - a) * r8 + a * (cur & 0xFFu)) >> 8;
  const u32 g = ((256u - a) * g8 + a * ((cur >> 8) & 0xFFu)) >> 8;
  const u32 b = ((256u - a) * b8 + a * ((cur >> 16) & 0xFFu)) >> 8;
  const u32 na = min(255u, a + 2u);
  return r + (g << 8) + (b << 16) + (na << 24);
}

// svo.cu:318-332 leaf blend for float colours (mesh path, Q15): trunc_s32(FFMA(cur, f2, (c*256)*f1)), fields ADDED
__device__ __forceinline__ u32 osl_blend_f4(u32 cur, float cr, float cg, float cb) {
  const int a = (int)(cur >> 24);
  const float f2 = __fdiv_rn((float)a, 256.0f);
  const float f1 = __fsub_rn(1.0f, f2);
  const float r = __fmaf_rn((float)(cur & 0xFFu), f2, __fmul_rn(__fmul_rn(cr, 256.0f), f1));
  const float g = __fmaf_rn((float)((cur >> 8) & 0xFFu), f2, __fmul_rn(__fmul_rn(cg, 256.0f), f1));
  const float b = __fmaf_rn((float)((cur >> 16) & 0xFFu), f2, __fmul_rn(__fmul_rn(cb, 256.0f), f1));
  const int na = min(255, a + 2);
  return (u32)__float2int_rz(r) + ((u32)__float2int_rz(g) << 8) + ((u32)__float2int_rz(b) << 16) + ((u32)na << 24);
}

// svo.cu:384-441 averageChildren with Q5 (all 8 children counted): integer mean of RGB, max of alpha
__device__ __forceinline__ u32 osl_average8(const u32 v[8]) {
  u32 r = 0, g = 0, b = 0, a = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r += v[i] & 0xFFu; g += (v[i] >> 8) & 0xFFu; b += (v[i] >> 16) & 0xFFu;
    a = max(a, v[i] >> 24);
  }
  return (r >> 3) + ((g >> 3) << 8) + ((b >> 3) << 16) + (a << 24);
}

__device__ __forceinline__ u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---- host-side internals ------------------------------------------------------------------------------------
// workspace of the big-input radix sort (osl_sort.cu)
struct OslSortWs {
  u32* hist = nullptr;  // [grid + 1][512] per-CTA digit counts / totals
  int grid = 0;
};
int osl_sort_big_passes(int key_bits, int* bits_out);
osl_status osl_sort_big(OslSortWs* ws, u64* kA, u32* pA, u64* kB, u32* pB, const int* d_n, const int* d_run_flag,
                        long long n_upper, int key_bits, bool with_pay, int grid_cap, cudaStream_t st);
void osl_sort_big_free(OslSortWs* ws);
osl_status osl_sort_big_reserve(OslSortWs* ws);

struct osl_svo {
  int device;
  TreeParams tp;
  u32* d_pool;
  size_t cap_nodes;
  int size;           // nodes (host copy, valid after sync)
  // workspace (sized for ws_cap inputs)
  size_t ws_cap;
  size_t ws_want;     // capacity to restore after osl_drop_workspace
  cudaEvent_t pose_ev;  // pipelined frames whose pose is produced on the caller's stream (lazily created)
  cudaEvent_t pose_read_ev;  // ... and the caller's stream is ordered after the kernel that reads it
  u64 *d_keysA[OSL_FRONT], *d_keysB[OSL_FRONT];  // sort ping/pong per front buffer
  u32 *d_payA[OSL_FRONT], *d_payB[OSL_FRONT];
  u64* d_keysC; u32* d_payC;   // k_sort_bucket slow-path scratch
  u64* d_bkeys[OSL_FRONT]; u32* d_bpay[OSL_FRONT];  // [OSL_BUCKETS][OSL_BUCKET_CAP]: the key list as k_emit files it by splitter range
  u64* d_split;                // [OSL_FRONT][BK_BUCKETS] splitters written by k_structure of frame f (set f % OSL_FRONT)
  u64* d_wcache;               // k_structure's walk cache (prefix -> node at a fixed depth); cleared when the pool is replaced
  int force_grid_sort;         // testing: always use the cooperative grid sort
  int zero_copy_rgb;           // measurement: osl_integrate_depth_host reads pinned colour planes in place
  uint8_t *d_m, *d_s;
  u32* d_start;       // per sorted key: node at the first depth it heads (k_structure phase A -> C)
  u32* d_flags;       // per virtual block: epoch of the frame whose count vector is published
  u64* d_lvltag;      // inside the d_flags allocation: [CTA][OSL_MAXD] epoch-tagged per-level counters (frame-sized exchange)
  u32* d_blockcnt;    // [k_structure CTAs][NC] per-CTA counter vectors
  u32* d_blockcnt_tot;  // sharded build: [3][NC_MAX] this rank's totals, then the external base / totals
  int shard_n, shard_lo, shard_grid; unsigned long long shard_f;  // between osl_shard_analyze and osl_shard_assign
  u32* d_cta_hist[OSL_FRONT];  // sort: [grid][256]
  u32* d_scan_totals; // [NC_MAX + 8] scratch words; word NC_MAX = arrival counter of k_levels' one-sided barrier
  LevelArrays lv[OSL_BACK];
  void* d_level_mem[OSL_BACK];
  FrameState* d_fs;                      // [0] persistent part (slot counters, cur_size), [1 + b] result block of slot b
  FrameState* h_ring;                    // pinned ring of per-frame result blocks
  cudaEvent_t ring_ev[OSL_RING];
  size_t ring_headroom[OSL_RING];
  int ring_mode[OSL_RING];
  unsigned long long ring_head, ring_tail;  // frames enqueued / consumed
  size_t inflight_headroom;              // worst-case node growth of frames not yet read back
  osl_status sticky_error;
  cudaStream_t last_stream;
  cudaStream_t copy_stream;              // H2D of host frames
  cudaStream_t pipe[4];                  // pipelined mode: E (k_emit), So (k_sort), S (k_structure + k_link), V (k_levels)
  cudaEvent_t emit_done[OSL_FRONT], sort_done[OSL_FRONT];
  cudaEvent_t struct_ev[OSL_RING];       // k_structure of frame f done (slot f % OSL_RING); ring_ev = k_levels done
  int last_piped;
  int counted_piped;                     // this tree is counted in g_osl_piped_trees
  int join_pending;                      // pipelined frames are in flight that other streams have not been ordered after
  // readers -> writers: work that READS the pool (the library's raycasts; foreign work on a stream handed to
  // osl_svo_join) must finish before a later frame rewrites it.  reader_ev fires after every reader noted so far
  // (each new reader's stream first waits for the previous state of the event, then re-records it); streams joined
  // by osl_svo_join are recorded lazily, when the next integrate is enqueued (osl_order_after_readers).
  cudaEvent_t reader_ev; int reader_pending;
  // k_frame pipeline (one launch per frame on pipe[2]): the frames whose structure / value stage the next launch
  // carries; completion is read from the pinned result block (FrameState::done_flag), not from events
  struct FzStage { int valid; unsigned long long f; int n, fslot, bslot, gS, gV; const uint8_t* rgb; } fz_so, fz_s, fz_v;
  int last_fused, fused_enabled;
  int trace_on; unsigned long long trace_seq;  // osl_debug_trace
  int ring_kind[OSL_RING];               // 0: completion = ring_ev, 1: completion = done_flag of the pinned block
  int fz_event_valid;                    // ring_ev of the last frame has been recorded behind the flushed pipeline
  u32* d_ready; u32* h_ready_vals;       // host frames: per staging slot, the sequence number its copies carry
  unsigned long long stage_frame[OSL_STAGES];  // frame number + 1 that last used the staging slot (k_frame path)
  cudaStream_t foreign_reader[8]; int foreign_n;
  int stage_timing, stage_valid;         // per-kernel CUDA-event timing of non-pipelined frames (bench / profiling)
  cudaEvent_t stage_ev[5];
  int pipeline;                          // 1: emit+sort run on the front stream (inputs are ready at call time)
  unsigned long long seq;                // frames enqueued (front buffer = seq % OSL_FRONT)
  int hint_emit, hint_level;             // last known n_emit / widest level (grid sizing); -1 = unknown
  int hint_n_in;
  cudaEvent_t stage_copied[OSL_STAGES], stage_free[OSL_STAGES];
  uint16_t* d_depth_stage[OSL_STAGES]; uint8_t* d_rgb_stage[OSL_STAGES]; size_t stage_cap; unsigned long long stage_seq;
  int structure_grid, levels_grid, structure_big_grid;
  OslSortWs sort_ws[OSL_FRONT];  // one per key-list slot: the sorts of consecutive pipelined frames may overlap
  osl_counters counters;
  // extraction scratch + the frontier of the last call (count / fill call pairs)
  long long *ex_kA, *ex_kB, *ex_res_k; u32 *ex_nA, *ex_nB, *ex_res_n; unsigned long long* ex_status; int* ex_cnt;
  size_t ex_cap; unsigned long long ex_epoch, ex_seq, ex_uploads, upload_count; int ex_valid, ex_depth, ex_size, ex_n;
  int sort_grid;      // co-resident CTAs for the cooperative sort
  int num_sms;
};

extern int g_osl_last_cuda_error;
extern long long g_osl_launches;
extern int g_osl_piped_trees;
#define OSL_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) {                             \
      g_osl_last_cuda_error = (int)e_;                   \
      return e_ == cudaErrorMemoryAllocation ? OSL_ERR_OOM : OSL_ERR_CUDA; \
    }                                                    \
  } while (0)
#define OSL_LAUNCHED(n) (g_osl_launches += (n))

// integrate pipeline (osl_integrate.cu)
struct EmitParams {
  const uint16_t* depth; const uint8_t* rgb; int w, h; float fx, fy; float M[16];  // mode 0
  const float* M_dev;  // mode 0: when set, the pose is read from device memory at kernel time (osl_integrate_depth_posed)
  int tiles_x, tiles_y;                                                             // mode 0: 64x32-pixel tiles
  const float* pts; int stride;                                                      // mode 1 (vec3) / 2 (vec4)
  int n; int mode;
  const u32* ready; u32 ready_seq;  // k_frame, host frames: *ready == ready_seq once the staging copies have landed
};
struct HostFrame { const uint16_t* h_depth; const uint8_t* h_rgb; };  // osl_integrate_depth_host: planes to stage
osl_status osl_run_integrate(osl_svo* t, EmitParams& ep, const void* colors, cudaStream_t st, const HostFrame* host);
osl_status osl_fused_flush(osl_svo* t);
osl_status osl_integrate_init(osl_svo* t);
osl_status osl_ensure_workspace(osl_svo* t, size_t n);
int osl_sort_occupancy();  // co-resident k_sort CTAs per SM
int osl_structure_occupancy();
int osl_structure_big_occupancy();
int osl_levels_occupancy();
osl_status osl_poll_results(osl_svo* t, bool block);
osl_status osl_grow_pool(osl_svo* t, size_t want_nodes, cudaStream_t st);
osl_status osl_reset_splitters(osl_svo* t);
void osl_drop_workspace(osl_svo* t);
osl_status osl_join(osl_svo* t, cudaStream_t st);
osl_status osl_set_device_size(osl_svo* t, int size);  // FrameState::cur_size + host copy; clears the walk cache
osl_status osl_validate_pool(osl_svo* t, int n_nodes);
osl_status osl_note_reader(osl_svo* t, cudaStream_t st);      // after enqueuing work on `st` that reads the pool
osl_status osl_note_foreign_reader(osl_svo* t, cudaStream_t st);  // `st` may carry foreign readers until the next integrate
osl_status osl_device_sort_pairs(u64* kA, u32* pA, u64* kB, u32* pB, int n, int key_bits, cudaStream_t st, int* in_B);

// raycast / extraction / image kernels
osl_status osl_launch_raycast(const u32* d_pool, const float center[3], float half_edge, uint8_t* d_out, int w, int h,
                              int row0, int rows, int band_h, int band_stride, float fov_deg, const float view[16], const osl_raycast_params* prm,
                              unsigned long long* d_stats, cudaStream_t st);
