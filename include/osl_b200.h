/*
 * osl_b200.h -- C ABI of libosl_b200.so: the B200-native (sm_100a) replacement for the per-frame hot path of
 * dkotfis/Octree-SLAM (depth map -> sparse voxel octree integration, octree raycast, voxel extraction).
 *
 * Plain C, POD arguments, explicit stream, int status, never exit().  Each entry point names the reference
 * interface it replaces (file:line in the reference tree).  The reference's own "extern \"C\"" functions take C++
 * references and glm types by value and are therefore not a C ABI; the C++ shim under
 * include/octree_slam/ re-creates those verbatim signatures on top of this header.
 *
 * Node pool layout (kept from the reference, common_types.h:75-79, svo.cu:130-136,269-275):
 *   unsigned int pool[2 * n_nodes]; node i = { word0 = pool[2i], word1 = pool[2i+1] }
 *   word0: bit 30 = has-children, bits 0-29 = index (in nodes) of the first of 8 contiguous children
 *   word1: R | G<<8 | B<<16 | A<<24  (A = observation counter: 127 empty, +2 per observation, saturates at 255)
 *   the root is implicit; its 8 children are nodes 0..7.
 *
 * All matrices are 16 floats, column-major (glm::mat4 memory layout).
 * "d_" pointers are device pointers on the tree's device, "h_" pointers are host pointers.
 * stream arguments are cudaStream_t passed as void* (NULL = the legacy default stream).
 */
#ifndef OSL_B200_H_
#define OSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int osl_status;
enum {
  OSL_OK = 0,
  OSL_ERR_INVALID = -1,       /* bad argument */
  OSL_ERR_CUDA = -2,          /* a CUDA runtime call failed (osl_last_cuda_error() has the code) */
  OSL_ERR_OOM = -3,           /* device allocation failed */
  OSL_ERR_POOL_OVERFLOW = -4, /* > 2^30 nodes: the 30-bit child index of the node layout is exhausted */
  OSL_ERR_UNSUPPORTED = -5
};

#define OSL_MAX_DEPTH 20 /* 1 + 3*20 = 61 key bits */

typedef struct osl_svo osl_svo; /* one GPU-resident octree (the reference's OctreeNode::gpu_data_/gpu_size_) */

/* Per-frame counters of the last integrate call + running totals. */
typedef struct osl_counters {
  int64_t n_points;                     /* N  inputs of the last call */
  int64_t n_valid;                      /* inputs with a valid key */
  int64_t n_unique;                     /* U  distinct leaf keys */
  int64_t n_split;                      /* S  nodes split == 8-node tiles allocated */
  int64_t pass_sizes[OSL_MAX_DEPTH + 1];/* |codes[i]| of the reference's pass i (svo.cu:220) */
  int64_t parents[OSL_MAX_DEPTH + 1];   /* P_l distinct touched nodes at depth l (l = 0 is the root) */
  int64_t n_nodes;                      /* octree_size after the call (nodes) */
  int64_t algorithmic_bytes;            /* 5N (or bytes of the point inputs) + 8U + 68S + 68*sum(P_l) */
  int64_t total_algorithmic_bytes;      /* running sum over all frames */
  int64_t frames;                       /* integrate calls completed so far */
} osl_counters;

typedef struct osl_raycast_params {
  float fx, fy;     /* ray-direction focal lengths; reference hard-codes 532.57 / 531.54 (cone_tracing_kernels.cu:45-46) */
  float start_dist; /* 0.002  (cone_tracing_kernels.cu:27) */
  float max_range;  /* 10.0   (cone_tracing_kernels.cu:24) */
  int mode;         /* 0 = ref_exact (accumulator reset every step, reference quirk Q8); 1 = fixed_accumulate */
} osl_raycast_params;

typedef struct osl_raycast_stats {
  int64_t rays, steps, visits; /* visits = word0 reads in descents; bytes = 4*rays + 4*visits + 4*steps */
} osl_raycast_stats;

/* ---- lifetime -------------------------------------------------------------------------------------------- */

/* Replaces: Octree::Octree + OctreeNode::pushToGPU + svo.cu:24-31 initOctree.
 * half_edge is Octree::size_ (HALF the cube edge, octree.h:118).  reserve_nodes pre-sizes the pool (0 = default);
 * the pool grows geometrically when needed (replacing the reference's per-frame realloc+copy, svo.cu:664-668). */
osl_status osl_svo_create(osl_svo** out, const float center[3], float half_edge, int max_depth,
                          size_t reserve_nodes, int device);
void osl_svo_destroy(osl_svo* t);
/* Drop all nodes (octree_size = 0), keep the allocations. */
osl_status osl_svo_reset(osl_svo* t);
/* Map growth.  Replaces Octree::expandBySize + OctreeNode::expand (octree.cpp:362-378, 183-206), which the reference
 * cannot run on a GPU-backed tree (quirk Q10: `expand()` refuses and `size_` is re-scaled anyway).  Each layer doubles
 * the half edge about the SAME centre and deepens the tree by one level, so the resolution is kept: the eight root
 * children move one level down, old child i becoming child 7-i (`oppositeNode(i)`) of a new node i whose other
 * seven children are empty (word0 = 0, value 127<<24 as splitNodes initialises them, svo.cu:271-275) and whose value
 * is averageChildren of that tile (svo.cu:384-441).  64 nodes are appended per layer.  Waits for the pipeline.
 * OSL_ERR_UNSUPPORTED when max_depth + layers > 20. */
osl_status osl_svo_expand(osl_svo* t, int layers);
int osl_svo_max_depth(const osl_svo* t);
/* bit 0: 1 (default) reproduces reference quirk Q3 (svo.cu:123 `while (r_key >= 15)`); 0 = leaves are never split.
 * bit 1 (testing aid): always sort with the cooperative grid radix sort, never with the splitter-based bucket sort.
 * bit 2 (measurement aid): osl_integrate_depth_host reads PINNED colour planes in place (zero-copy gather of the one
 * colour per observed leaf) instead of staging them; measured slower than the DMA on B200, hence off by default.
 * bit 3 (testing / measurement aid): pipelined depth frames always run as four kernels on four streams, never as one
 * k_frame launch per frame (the environment variable OSL_NO_FUSED does the same for every tree). */
osl_status osl_svo_set_quirks(osl_svo* t, int ref_quirks);

/* Pipelined mode for DEVICE-resident inputs (default 0).  With 1 the caller promises that the input buffers of every
 * osl_integrate_* call are complete when the call is made (not merely stream-ordered before it); the library then runs
 * back-projection + key sort of frame f+1 on an internal stream, overlapped with the tree update of frame f on the
 * caller's stream.  Results are identical.  osl_integrate_depth_host always pipelines (it owns the copies).
 * Readers and writers of the pool are ordered in BOTH directions: the library's raycast / extraction / download / view
 * entry points wait for the frames enqueued before them (osl_svo_join), and every later frame's pool-writing stages
 * wait for the raycasts enqueued before it -- integrate(f); raycast(stream R); integrate(f+1) renders the map after
 * frame f, whole.  Foreign readers: see osl_svo_join. */
osl_status osl_svo_set_pipeline(osl_svo* t, int enabled);

/* Profiling aid: time k_emit / k_sort / k_structure / k_levels of every NON-pipelined integrate call with CUDA events
 * on the caller's stream; osl_get_stage_times returns the 4 durations (ms) of the last timed frame (waits for it). */
osl_status osl_svo_set_stage_timing(osl_svo* t, int enabled);
osl_status osl_get_stage_times(osl_svo* t, float ms[4]);

/* ---- integrate (svo.h:14-16) ------------------------------------------------------------------------------- */

/* Fused main.cpp:39-44: generateVertexMap (image_kernels.cu:24-53) -> transformVertexMap (:206-215) ->
 * svoFromPointCloud (svo.cu:642-696).  d_depth: w*h uint16 millimetres, d_rgb: w*h*3 bytes (Color256).
 * Asynchronous on `stream` (4 kernel launches, no host synchronisation) except when the pool or the workspace has to
 * grow, or when the caller runs more than 3 frames ahead of the device (it then waits for the oldest frame).
 * One stream per tree: consecutive calls on different streams are not ordered against each other. */
osl_status osl_integrate_depth(osl_svo* t, const uint16_t* d_depth, const uint8_t* d_rgb, int w, int h, float fx,
                               float fy, const float pose[16], void* stream);
/* Same with the pose in DEVICE memory (16 floats, column-major), read when the first kernel of the frame runs: the
 * pose may be the result of work queued earlier on `stream` (osl_tracker_update + osl_tracker_pose_device), and
 * work queued on `stream` afterwards may overwrite it (in pipelined mode `stream` is ordered after that kernel). */
osl_status osl_integrate_depth_posed(osl_svo* t, const uint16_t* d_depth, const uint8_t* d_rgb, int w, int h, float fx,
                                     float fy, const float* d_pose_colmajor, void* stream);
/* Same from host buffers (what OpenNIDevice::readFrame + mainLoop do, openni_device.cpp:122,144).  The H2D copies run
 * on an internal copy stream into rotating device slots so the transfer of frame f+1 overlaps the kernels of frame f;
 * the call returns without waiting for the device.  Pinned source buffers must stay untouched until the frame has
 * completed (osl_svo_sync). */
osl_status osl_integrate_depth_host(osl_svo* t, const uint16_t* h_depth, const uint8_t* h_rgb, int w, int h, float fx,
                                    float fy, const float pose[16], void* stream);
/* Replaces svoFromPointCloud (svo.h:16, svo.cu:642): d_xyz = n glm::vec3 (12-byte stride), d_rgb = n Color256. */
osl_status osl_integrate_points(osl_svo* t, const float* d_xyz, const uint8_t* d_rgb, int n, void* stream);
/* Replaces svoFromVoxelGrid (svo.h:14, svo.cu:584): n glm::vec4 centres and n glm::vec4 colours (0..1 floats). */
osl_status osl_integrate_voxels(osl_svo* t, const float* d_centers4, const float* d_colors4, int n, void* stream);

/* Order `stream` after every integrate enqueued so far (device-side wait, the host does not block).  Strict-mode
 * frames are stream-ordered anyway; pipelined frames finish on an internal stream, and this library's own entry
 * points (raycast, extraction, download, view) join automatically -- call this before FOREIGN work on `stream` that
 * reads the pool, or before recording a timing event.  The reverse edge is taken care of as well: whatever is queued on
 * `stream` up to the NEXT osl_integrate_* call on this tree is treated as a reader of the pool, and that frame's
 * pool-writing stages wait for it (`stream` must still exist then).  Foreign readers on other streams, or queued later,
 * are the caller's to order (join again). */
osl_status osl_svo_join(osl_svo* t, void* stream);

/* Wait for every integrate enqueued so far; returns a deferred error (e.g. OSL_ERR_POOL_OVERFLOW) if one occurred. */
osl_status osl_svo_sync(osl_svo* t);

/* ---- views of the tree ------------------------------------------------------------------------------------- */

/* Replaces Octree::extractSVO (octree.cpp:339-360): aliases the pool, no ownership.  Synchronizes. */
osl_status osl_svo_view(const osl_svo* t, const uint32_t** d_pool, int* n_nodes, float center[3], float* half_edge);
int osl_svo_size(const osl_svo* t); /* nodes; synchronizes */
osl_status osl_svo_download(const osl_svo* t, uint32_t* h_pool, int cap_nodes);
osl_status osl_svo_upload(osl_svo* t, const uint32_t* h_pool, int n_nodes);
osl_status osl_get_counters(const osl_svo* t, osl_counters* out);
/* Checkpoint / resume: 32-byte header (magic, max_depth, n_nodes, centre, half edge) + the flat 2*n uint32 pool (the
 * array OctreeNode::pullToCPU / pushToGPU exchange, octree.cpp:41-169).  osl_svo_load requires a tree created with
 * the same max_depth / centre / half edge (after osl_svo_expand: the expanded values, osl_svo_max_depth / osl_svo_view). */
osl_status osl_svo_save(const osl_svo* t, const char* path);
osl_status osl_svo_load(osl_svo* t, const char* path);

/* ---- replicas of one map on several GPUs (image rows shard, every rank holds the tree) ------------ */

/* Everything stays in DEVICE memory; a collective library (NCCL) or a peer copy moves the bytes.
 * Full copy: the receiver reserves room, exposes its pool, the collective writes the sender's flat 2*n uint32 array
 * (the reference's wire format, octree.cpp:113-169) into it, osl_svo_adopt validates the child pointers and publishes
 * it (OSL_ERR_INVALID, and an empty tree, when max_depth / centre / half edge differ from the source's or the array is
 * corrupt).  osl_svo_pool_device waits for the pipeline; the pointer is valid until the pool next grows. */
osl_status osl_svo_reserve(osl_svo* t, size_t n_nodes);
osl_status osl_svo_pool_device(osl_svo* t, uint32_t** d_pool, size_t* cap_nodes);
osl_status osl_svo_adopt(osl_svo* t, int n_nodes, int max_depth, const float center[3], float half_edge);
/* Delta: what the LAST integrate call changed -- a 32-byte header, the (index, word0, word1) of every pre-existing or
 * new node on a touched path (the frame's level lists) and the nodes the call appended (new tiles only ever go to the
 * end of the pool).  ~0.3 MB for a 640x480 frame.  _bytes: size of the packed delta (waits for the frame); _pack writes
 * it to d_buf (device memory, `cap` bytes); _apply on a replica that is in the state BEFORE that call brings it to the
 * state after it (OSL_ERR_INVALID otherwise). */
size_t osl_svo_delta_bytes(osl_svo* t);
osl_status osl_svo_delta_pack(osl_svo* t, void* d_buf, size_t cap, size_t* bytes, void* stream);
osl_status osl_svo_delta_apply(osl_svo* t, const void* d_buf, size_t bytes, void* stream);

/* ---- ONE map built by several GPUs (Morton-range shards + the per-pass split-count prefix) ------------ */

/* Input: a voxel grid in Morton order without invalid entries (what osl_voxelize_* and osl_extract_voxels emit), present
 * on every rank, all ranks' trees in the same state.  Rank r takes the contiguous slice [lo, hi) of the grid.  Between
 * the calls the CALLER exchanges two things with the other ranks (shard.integrate_voxels_sharded: NCCL all-gathers):
 *   1. osl_shard_analyze  -> h_totals[*n_counters]: this rank's counters (levels, then (frontier, depth) buckets)
 *      exchange: h_base[c] = sum of the LOWER ranks' totals, h_totals_all[c] = sum over ALL ranks
 *   2. osl_shard_assign   allocates tiles at the global ranks (pass = depth - frontier depth, then numeric key, over all
 *      ranks: svo.cu:220,284), writes this rank's nodes, folds the values of this rank's sub-trees
 *      exchange: osl_shard_delta_pack on every rank, all-gather, osl_shard_delta_apply of every OTHER rank's delta
 *   3. osl_shard_fixup    re-averages the nodes on the paths of the slices' first keys (children in two ranks) and writes
 *      the root average (Q6)
 * The pool is then bit-identical on every rank with a single-GPU osl_integrate_voxels of the whole grid.
 * OSL_ERR_UNSUPPORTED: the slice is not sorted / contains invalid voxels. */
osl_status osl_shard_analyze(osl_svo* t, const float* d_centers4, int n_total, int lo, int hi, uint32_t* h_totals,
                             int* n_counters, void* stream);
osl_status osl_shard_assign(osl_svo* t, const float* d_colors4, const uint32_t* h_base, const uint32_t* h_totals_all,
                            void* stream);
size_t osl_shard_delta_bytes(osl_svo* t);
osl_status osl_shard_delta_pack(osl_svo* t, void* d_buf, size_t cap, size_t* bytes, void* stream);
osl_status osl_shard_delta_apply(osl_svo* t, const void* d_buf, size_t bytes, void* stream);
osl_status osl_shard_fixup(osl_svo* t, const float* d_centers4, int n_total, const int* h_starts, int n_bounds, void* stream);

/* ---- raycast (cone_tracing_kernels.h:16) ------------------------------------------------------------------- */

/* Replaces rendering::coneTraceSVO (cone_tracing_kernels.cu:157-198).  d_out: w*h uchar4 {R,G,B,A}.
 * prm == NULL uses the reference constants in ref_exact mode.  stats may be NULL. */
osl_status osl_raycast(const osl_svo* t, uint8_t* d_out_rgba, int w, int h, float fov_deg, const float view[16],
                       const osl_raycast_params* prm, void* stream);
osl_status osl_raycast_host(const osl_svo* t, uint8_t* h_out_rgba, int w, int h, float fov_deg, const float view[16],
                            const osl_raycast_params* prm, osl_raycast_stats* stats, void* stream);
/* Rows [row0, row0 + rows) of the w x h image only, written to d_out_rgba[0 .. rows*w*4): the multi-GPU
 * decomposition of coneTraceSVO (rays are independent; every GPU holds the tree).  h_stats may be NULL. */
osl_status osl_raycast_rows(const osl_svo* t, uint8_t* d_out_rgba, int w, int h, int row0, int rows, float fov_deg,
                            const float view[16], const osl_raycast_params* prm, osl_raycast_stats* h_stats,
                            void* stream);
/* All rows one rank owns under the interleaved-band decomposition, in one launch: bands of band_h rows are dealt
 * round-robin to n_ranks ranks (band k -> rank k % n_ranks); the rank's rows land compactly, in image order, in
 * d_out_rgba.  *rows_out (may be NULL) = number of rows rendered. */
osl_status osl_raycast_bands(const osl_svo* t, uint8_t* d_out_rgba, int w, int h, int band_h, int n_ranks, int rank,
                             float fov_deg, const float view[16], const osl_raycast_params* prm, int* rows_out,
                             void* stream);
/* Raycast an arbitrary pool (SVO struct by value in the reference). */
osl_status osl_raycast_pool(const uint32_t* d_pool, const float center[3], float half_edge, uint8_t* d_out_rgba, int w,
                            int h, float fov_deg, const float view[16], const osl_raycast_params* prm,
                            osl_raycast_stats* h_stats, void* stream);

/* ---- extraction (svo.h:18) --------------------------------------------------------------------------------- */

/* Replaces extractVoxelGridFromSVO (svo.cu:699-745).  Returns the voxel count in *n_out; when the d_ buffers are
 * non-NULL and cap >= count they receive glm::vec4 centres / colours (and optionally int64 leading-1 keys). */
osl_status osl_extract_voxels(const osl_svo* t, int max_depth, float* d_centers4, float* d_colors4, int64_t* d_keys,
                              int64_t cap, int64_t* n_out, void* stream);

/* ---- mesh voxelisation (voxelization.h:19-21) ---------------------------------------------------------------- */

/* Sparse stand-in for voxelization::meshToVoxelGrid (voxelization.cu:238-323,381-405; dense 256^3 via voxelpipe, which
 * no longer builds): the leaf cells of the depth-max_depth grid over the tree cube that overlap a triangle, as a
 * VoxelGrid in Morton-key order (so svoFromVoxelGrid's key sort is the identity and quirk Q11 leaves every colour on
 * its voxel).  d_vertices: n_vertices x 3 floats, d_triangles: n_triangles x 3 ints, d_tri_colors4: one glm::vec4 per
 * triangle (NULL = white).  The outputs are cudaMalloc'd here (as extractVoxelGridFromSVO does for VoxelGrid,
 * svo.cu:732-733) and released with osl_free_device / cudaFree; d_keys_out / d_tris_out may be NULL. */
osl_status osl_voxelize_mesh(const float* d_vertices, int n_vertices, const int* d_triangles, int n_triangles,
                             const float* d_tri_colors4, const float center[3], float half_edge, int max_depth,
                             float** d_centers4_out, float** d_colors4_out, int64_t** d_keys_out, int** d_tris_out,
                             int64_t* n_out, void* stream);
/* The REFERENCE's voxelisation rule on the reference's grid: voxelization::meshToVoxelGrid (voxelization.cu:24,281-285)
 * = voxelpipe THIN_RASTER / NO_BLENDING on the dense 2^log_n-per-axis grid over [bbox0, bbox1] (the mesh bounding box:
 * anisotropic cells), restated from the vendored library's source (octree-slam_b200/csrc/osl_voxelize_thin.cu lists
 * file:line) -- per column of the dominant-axis projection that the triangle's 2-D footprint touches, the one voxel
 * that holds the plane's depth at the column centre; lowest triangle index where triangles collide.  Centres follow
 * getCenterFromIndex (voxelization.cu:58-78).  cube_depth > 0: the voxels come out in the order of the Morton keys of
 * their centres in that octree cube (what svoFromVoxelGrid sorts by), so quirk Q11 leaves the colours in place;
 * cube_depth = 0: grid order (z, y, x).  d_cells_out: 3 ints (x, y, z) per voxel.  log_n in [3, 9]. */
osl_status osl_voxelize_thin(const float* d_vertices, int n_vertices, const int* d_triangles, int n_triangles,
                             const float* d_tri_colors4, const float bbox0[3], const float bbox1[3], int log_n,
                             const float cube_center[3], float cube_half, int cube_depth, float** d_centers4_out,
                             float** d_colors4_out, int** d_cells_out, int** d_tris_out, int64_t* n_out, void* stream);
void osl_free_device(void* p);
/* device-to-device copy (for bindings that must move a library-allocated result into a buffer they own) */
osl_status osl_copy_device(void* d_dst, const void* d_src, size_t bytes);

/* ---- per-frame image kernels (image_kernels.h:21,24,52) ---------------------------------------------------- */

osl_status osl_generate_vertex_map(const uint16_t* d_depth, float* d_xyz, int width, int height, float fx, float fy,
                                   int img_w, int img_h, void* stream);
osl_status osl_transform_vertex_map(float* d_xyz, const float trans[16], int n, void* stream);
/* bbox[6] = {min xyz, max xyz}, in/out exactly like BoundingBox& in the reference (left fold seeded with the
 * incoming box; points with non-finite x or z are skipped; a (0,0,0) accumulator is replaced). */
osl_status osl_point_cloud_bbox(const float* d_xyz, int n, float bbox[6], void* stream);
/* computeKeys<vec3|vec4> (svo.cu:93-106): leading-1 Morton keys, 1 for invalid points.  stride = 3 or 4 floats. */
osl_status osl_compute_keys(const float* d_pts, int stride, int n, const float center[3], float half_edge,
                            int max_depth, int64_t* d_keys, void* stream);

/* ---- camera tracking (the step main.cpp:35 has commented out) ------------------- */

/* Free functions of image_kernels.h on device buffers; asynchronous on `stream`.
 * bilateralFilter (image_kernels.cu:168-176; 7x7, sigma 4.5 px / 40 mm). */
osl_status osl_bilateral_filter(const uint16_t* d_in, uint16_t* d_out, int width, int height, void* stream);
/* subsampleDepth<uint16_t> / subsample<float> (image_kernels.cu:262-321).  (width, height) are the dimensions of
 * d_in; d_out receives (width/2) x (height/2) and must not alias d_in (the reference copies back in place). */
osl_status osl_subsample_depth(const uint16_t* d_in, uint16_t* d_out, int width, int height, void* stream);
osl_status osl_subsample_f32(const float* d_in, float* d_out, int width, int height, void* stream);
/* generateNormalMap (image_kernels.cu:131-135), transformNormalMap (:228-230), colorToIntensity (:188-192). */
osl_status osl_generate_normal_map(const float* d_vertex, float* d_normal, int width, int height, void* stream);
osl_status osl_transform_normal_map(float* d_normal, const float trans_colmajor[16], int n, void* stream);
osl_status osl_color_to_intensity(const uint8_t* d_rgb, float* d_out, int n, void* stream);
/* computeICPCost2 (localization_kernels.h:40, localization_kernels.cu:313-330): A[36] (row-major) and b[6] are HOST
 * arrays; the call synchronises `stream`.  flags bit 0: use the exact point-to-plane Jacobian v2 x n1 instead of
 * the reference's mis-indexed G^T (quirk Q17). */
osl_status osl_icp_cost(const float* d_last_vertex, const float* d_last_normal, const float* d_this_vertex,
                        const float* d_this_normal, int n, int flags, float A[36], float b[6], int* pairs,
                        void* stream);

/* sensor::RGBDCamera (rgbd_camera.h:17-82, rgbd_camera.cpp:21-191).  flags bit 0 = 0 reproduces the reference
 * (quirks Q17, Q18: the Jacobian above, angles negated, position_ never leaves 0); flags bit 0 = 1 is the
 * corrected tracker (exact Jacobian, increment T*R, camera-to-world pose accumulated as world * update). */
/* width and height must be multiples of 4 (two pyramid halvings) and at least 8. */
typedef struct osl_tracker osl_tracker;
osl_status osl_tracker_create(osl_tracker** out, int width, int height, float fx, float fy, int flags, int device);
void osl_tracker_destroy(osl_tracker* t);
osl_status osl_tracker_reset(osl_tracker* t);
/* RGBDCamera::update: bilateral filter, 3-level pyramid of vertex + normal maps, 4 + 5 + 10 Gauss-Newton
 * iterations coarse to fine against the previous frame, pose update.  27 launches, no host round trip;
 * asynchronous on `stream`.  _host: the depth image is in host memory (pinned for a truly asynchronous copy). */
osl_status osl_tracker_update(osl_tracker* t, const uint16_t* d_depth, void* stream);
osl_status osl_tracker_update_host(osl_tracker* t, const uint16_t* h_depth, void* stream);
/* Waits for the last update.  pose (column-major) = the matrix main.cpp:40 applies to the vertex map,
 * mat4(orientation_) * translate(mat4(1), position_); any output may be NULL. */
osl_status osl_tracker_get_pose(osl_tracker* t, float pose_colmajor[16], float position[3],
                                float orientation_colmajor[9], int* lost, int* pairs);
/* Device address of the pose matrix (16 floats, column-major; what osl_tracker_get_pose returns as `pose`), rewritten
 * by every update.  With osl_integrate_depth_posed on the same stream the SLAM loop of main.cpp:33-44 (tracking
 * live) runs without any host round trip: 27 + 4 launches per frame. */
osl_status osl_tracker_pose_device(osl_tracker* t, const float** d_pose);
/* The pyramid level of the last processed frame (device pointers, valid until the next update). */
osl_status osl_tracker_view(osl_tracker* t, int level, const float** d_vertex, const float** d_normal, int* width,
                            int* height);

/* ---- misc -------------------------------------------------------------------------------------------------- */
const char* osl_status_string(osl_status s);
int osl_last_cuda_error(void);
const char* osl_version(void);
/* bytes of the per-frame result block the device writes into pinned host memory (sizes, counters) */
int osl_frame_result_bytes(void);
/* Profiling aid: SM-clock checkpoints written by CTA 0 of the last k_emit / k_sort_bucket / k_structure / k_levels
 * launches (indices documented in tools/phase_profile.py).  Synchronizes the device. */
osl_status osl_debug_profile(unsigned long long* out, int n);
/* Tracing aid (tools/large_profile.py): SM clock of every CTA of the last big-input structure launch at the start / end
 * of its analysis phase and the start / end of its assignment phase; out[4][1024]. */
osl_status osl_debug_cta_profile(unsigned long long* out);
/* Profiling aid: enable = 1 clears the table and records, for the next k_frame launches of `t` (round-robin over 32
 * slots), the [min start, max end] %globaltimer nanoseconds of each role: 0 structure, 1 values, 2 emit, 3 sort,
 * 4 CTA arrival before the grid dependency wait.  out (may be NULL) receives 32 x 5 x 2 values.  Synchronizes. */
osl_status osl_debug_trace(osl_svo* t, int enable, unsigned long long* out);
/* Test aid: fills the hint table of the structure stage's tree walks (node-index guesses that are verified against the
 * pool before use) with pseudo-random words; results must not change.  Synchronizes. */
osl_status osl_debug_scramble_hints(osl_svo* t, unsigned seed);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches) */
int64_t osl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OSL_B200_H_ */
