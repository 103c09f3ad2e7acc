// ref_shim.cu -- headless C harness around the REFERENCE's own CUDA sources.  TEST INFRASTRUCTURE ONLY.
//
// The reference (dkotfis/Octree-SLAM) has no CPU path and its main() needs a GLFW window and an OpenNI
// camera (src/main.cpp:113-137), so this file (our code, not reference code) replays exactly the calls of
// mainLoop() (src/main.cpp:38-44) and CUDARenderer::coneTraceSVO (src/rendering/cuda_renderer.cpp:163)
// against the reference's svo.cu / cone_tracing_kernels.cu / image_kernels.cu / common_types.cu objects,
// which oracle/Makefile compiles UNMODIFIED from /root/reference.  It is linked into
// oracle/_ref/libosl_ref.so (and the "ref + 64-bit patch" variant libosl_ref64.so).
// Used (a) to pin the CPU oracle and generate tests/golden, (b) as the parity checker of the -m gpu tests,
// (c) as the timed reference arm of bench.py (--impl reference).  Never used by the product path.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include <octree_slam/common_types.h>
#include <octree_slam/rendering/cone_tracing_kernels.h>
#include <octree_slam/sensor/image_kernels.h>
#include <octree_slam/sensor/localization_kernels.h>
#include <octree_slam/sensor/rgbd_camera.h>
#include <octree_slam/world/svo/svo.h>

namespace os = octree_slam;

struct ref_tree {
  unsigned int* d_pool;  // OctreeNode::gpu_data_ (octree.h:74)
  int size;              // OctreeNode::gpu_size_ (nodes)
  glm::vec3 center;      // Octree::center_
  float half_edge;       // Octree::size_
  int max_depth;
  glm::vec3* d_points;   // main.cpp `points_`
  int points_cap;
  uint16_t* d_depth;     // RawFrame::depth
  Color256* d_color;     // RawFrame::color
  int frame_cap;
};

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

extern "C" {

// Octree::addCloud's depth derivation (octree.cpp:283-284) for node_depth = 0
int ref_max_depth_from_resolution(float half_edge, float resolution) {
  float edge_length = half_edge / pow(2.0f, (float)0);
  return (int)ceil(log((float)(edge_length / resolution)) / log(2.0f));
}

ref_tree* ref_create(const float center[3], float half_edge, int max_depth) {
  ref_tree* t = new ref_tree();
  memset(t, 0, sizeof(*t));
  t->center = glm::vec3(center[0], center[1], center[2]);
  t->half_edge = half_edge;
  t->max_depth = max_depth;
  return t;
}

void ref_destroy(ref_tree* t) {
  if (!t) return;
  if (t->d_pool) cudaFree(t->d_pool);
  if (t->d_points) cudaFree(t->d_points);
  if (t->d_depth) cudaFree(t->d_depth);
  if (t->d_color) cudaFree(t->d_color);
  delete t;
}

int ref_size(const ref_tree* t) { return t->size; }

int ref_download_pool(const ref_tree* t, unsigned int* h_out) {
  if (t->size == 0) return 0;
  return (int)cudaMemcpy(h_out, t->d_pool, sizeof(unsigned int) * 2 * (size_t)t->size, cudaMemcpyDeviceToHost);
}

int ref_upload_pool(ref_tree* t, const unsigned int* h_pool, int size) {
  if (t->d_pool) cudaFree(t->d_pool);
  t->d_pool = nullptr;
  t->size = size;
  if (size == 0) return 0;
  cudaMalloc((void**)&t->d_pool, sizeof(unsigned int) * 2 * (size_t)size);
  return (int)cudaMemcpy(t->d_pool, h_pool, sizeof(unsigned int) * 2 * (size_t)size, cudaMemcpyHostToDevice);
}

static void ensure_frame(ref_tree* t, int n) {
  if (n > t->points_cap) {
    if (t->d_points) cudaFree(t->d_points);
    cudaMalloc((void**)&t->d_points, sizeof(glm::vec3) * (size_t)n);
    t->points_cap = n;
  }
  if (n > t->frame_cap) {
    if (t->d_depth) cudaFree(t->d_depth);
    if (t->d_color) cudaFree(t->d_color);
    cudaMalloc((void**)&t->d_depth, sizeof(uint16_t) * (size_t)n);
    cudaMalloc((void**)&t->d_color, sizeof(Color256) * (size_t)n);
    t->frame_cap = n;
  }
}

static glm::mat4 mat_from(const float m[16]) {
  glm::mat4 r;
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++) r[c][k] = m[4 * c + k];
  return r;
}

// main.cpp:38-44 on device-resident frame buffers; the timed region is the reference path itself
// (its own cudaMalloc/cudaFree, D2D pool copy and syncs included), wall clock + final sync.
int ref_integrate_depth_dev(ref_tree* t, const uint16_t* d_depth, const uint8_t* d_rgb, int w, int h, float fx,
                            float fy, const float pose[16], double* ms_out) {
  const int n = w * h;
  ensure_frame(t, n);
  glm::mat4 M = mat_from(pose);
  cudaDeviceSynchronize();
  double t0 = now_ms();
  os::sensor::generateVertexMap(d_depth, t->d_points, w, h, glm::vec2(fx, fy), make_int2(w, h));
  os::sensor::transformVertexMap(t->d_points, M, n);
  cudaDeviceSynchronize();
  BoundingBox cloud_bbox;
  os::sensor::computePointCloudBoundingBox(t->d_points, n, cloud_bbox);
  os::svo::svoFromPointCloud(t->d_points, (const Color256*)d_rgb, n, t->max_depth, t->d_pool, t->size, t->center,
                             t->half_edge);
  cudaDeviceSynchronize();
  if (ms_out) *ms_out = now_ms() - t0;
  return (int)cudaGetLastError();
}

// Same, from HOST buffers (the H2D copies are what OpenNIDevice::readFrame does, openni_device.cpp:122,144)
int ref_integrate_depth(ref_tree* t, const uint16_t* h_depth, const uint8_t* h_rgb, int w, int h, float fx, float fy,
                        const float pose[16], double* ms_out, double* ms_e2e_out) {
  const int n = w * h;
  ensure_frame(t, n);
  cudaDeviceSynchronize();
  double t0 = now_ms();
  cudaMemcpy(t->d_depth, h_depth, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice);
  cudaMemcpy(t->d_color, h_rgb, 3 * (size_t)n, cudaMemcpyHostToDevice);
  int rc = ref_integrate_depth_dev(t, t->d_depth, (const uint8_t*)t->d_color, w, h, fx, fy, pose, ms_out);
  if (ms_e2e_out) *ms_e2e_out = now_ms() - t0;
  return rc;
}

// generateVertexMap + transformVertexMap only; downloads the vertex map (pins a-1 / a-2)
int ref_vertex_map(const uint16_t* h_depth, int w, int h, float fx, float fy, const float pose[16], float* h_xyz) {
  const int n = w * h;
  uint16_t* d_depth; glm::vec3* d_pts;
  cudaMalloc((void**)&d_depth, sizeof(uint16_t) * (size_t)n);
  cudaMalloc((void**)&d_pts, sizeof(glm::vec3) * (size_t)n);
  cudaMemcpy(d_depth, h_depth, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice);
  os::sensor::generateVertexMap(d_depth, d_pts, w, h, glm::vec2(fx, fy), make_int2(w, h));
  if (pose) os::sensor::transformVertexMap(d_pts, mat_from(pose), n);
  cudaDeviceSynchronize();
  cudaMemcpy(h_xyz, d_pts, sizeof(glm::vec3) * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d_depth); cudaFree(d_pts);
  return (int)cudaGetLastError();
}

int ref_bbox(const float* h_xyz, int n, float bbox[6]) {
  glm::vec3* d_pts;
  cudaMalloc((void**)&d_pts, sizeof(glm::vec3) * (size_t)n);
  cudaMemcpy(d_pts, h_xyz, sizeof(glm::vec3) * (size_t)n, cudaMemcpyHostToDevice);
  BoundingBox b;
  b.bbox0 = glm::vec3(bbox[0], bbox[1], bbox[2]);
  b.bbox1 = glm::vec3(bbox[3], bbox[4], bbox[5]);
  os::sensor::computePointCloudBoundingBox(d_pts, n, b);
  bbox[0] = b.bbox0.x; bbox[1] = b.bbox0.y; bbox[2] = b.bbox0.z;
  bbox[3] = b.bbox1.x; bbox[4] = b.bbox1.y; bbox[5] = b.bbox1.z;
  cudaFree(d_pts);
  return (int)cudaGetLastError();
}

// svoFromPointCloud (svo.cu:642) from host arrays
int ref_integrate_points(ref_tree* t, const float* h_xyz, const uint8_t* h_rgb, int n) {
  ensure_frame(t, n);
  cudaMemcpy(t->d_points, h_xyz, sizeof(glm::vec3) * (size_t)n, cudaMemcpyHostToDevice);
  cudaMemcpy(t->d_color, h_rgb, 3 * (size_t)n, cudaMemcpyHostToDevice);
  os::svo::svoFromPointCloud(t->d_points, t->d_color, n, t->max_depth, t->d_pool, t->size, t->center, t->half_edge);
  cudaDeviceSynchronize();
  return (int)cudaGetLastError();
}

// svoFromVoxelGrid (svo.cu:584) from host arrays
int ref_integrate_voxels(ref_tree* t, const float* h_centers4, const float* h_colors4, int n) {
  VoxelGrid grid;
  cudaMalloc((void**)&grid.centers, sizeof(glm::vec4) * (size_t)n);
  cudaMalloc((void**)&grid.colors, sizeof(glm::vec4) * (size_t)n);
  cudaMemcpy(grid.centers, h_centers4, sizeof(glm::vec4) * (size_t)n, cudaMemcpyHostToDevice);
  cudaMemcpy(grid.colors, h_colors4, sizeof(glm::vec4) * (size_t)n, cudaMemcpyHostToDevice);
  grid.size = n;
  os::svo::svoFromVoxelGrid(grid, t->max_depth, t->d_pool, t->size, t->center, t->half_edge);
  cudaDeviceSynchronize();
  return (int)cudaGetLastError();  // ~VoxelGrid frees centers/colors (common_types.cu:47-52)
}

// extractVoxelGridFromSVO (svo.cu:699).  Returns the voxel count; copies up to cap voxels out.
long long ref_extract_voxels(ref_tree* t, int max_depth, float* h_centers4, float* h_colors4, long long cap) {
  if (t->size == 0) return 0;
  VoxelGrid grid;
  os::svo::extractVoxelGridFromSVO(t->d_pool, t->size, max_depth, t->center, t->half_edge, grid);
  long long n = grid.size;
  if (h_centers4 && h_colors4 && n <= cap && n > 0) {
    cudaMemcpy(h_centers4, grid.centers, sizeof(glm::vec4) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_colors4, grid.colors, sizeof(glm::vec4) * (size_t)n, cudaMemcpyDeviceToHost);
  }
  if (n == 0) { cudaFree(grid.centers); cudaFree(grid.colors); }  // dtor only frees when size > 0
  return n;
}

// rendering::coneTraceSVO (cone_tracing_kernels.cu:157) into a cudaMalloc'd uchar4 buffer, then D2H.
int ref_raycast(const ref_tree* t, uint8_t* h_out, int w, int h, float fov_deg, const float view[16], double* ms_out) {
  uchar4* d_out;
  cudaMalloc((void**)&d_out, sizeof(uchar4) * (size_t)w * h);
  SVO svo;
  svo.data = t->d_pool;
  svo.center = t->center;
  svo.size = t->half_edge;  // Octree::extractSVO: size_/2^node_depth, node_depth = 0 (octree.cpp:357)
  cudaDeviceSynchronize();
  double t0 = now_ms();
  os::rendering::coneTraceSVO(d_out, glm::vec2((float)w, (float)h), fov_deg, mat_from(view), svo);
  cudaDeviceSynchronize();
  if (ms_out) *ms_out = now_ms() - t0;
  if (h_out) cudaMemcpy(h_out, d_out, sizeof(uchar4) * (size_t)w * h, cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  return (int)cudaGetLastError();
}

// ---- camera tracking (image_kernels.cu:104-321, localization_kernels.cu, rgbd_camera.cpp) ----------------------

int ref_bilateral(const uint16_t* h_in, int w, int h, uint16_t* h_out) {
  const size_t n = (size_t)w * h;
  uint16_t *d_in, *d_out;
  cudaMalloc((void**)&d_in, 2 * n); cudaMalloc((void**)&d_out, 2 * n);
  cudaMemcpy(d_in, h_in, 2 * n, cudaMemcpyHostToDevice);
  os::sensor::bilateralFilter(d_in, d_out, w, h);
  cudaMemcpy(h_out, d_out, 2 * n, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out);
  return (int)cudaGetLastError();
}

// subsampleDepth<uint16_t> works in place: the first (w/2)*(h/2) elements of the buffer are the result
int ref_subsample_depth(const uint16_t* h_in, int w, int h, uint16_t* h_out) {
  const size_t n = (size_t)w * h;
  uint16_t* d;
  cudaMalloc((void**)&d, 2 * n);
  cudaMemcpy(d, h_in, 2 * n, cudaMemcpyHostToDevice);
  os::sensor::subsampleDepth<uint16_t>(d, w, h);
  cudaMemcpy(h_out, d, 2 * (n / 4), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return (int)cudaGetLastError();
}

int ref_subsample_f32(const float* h_in, int w, int h, float* h_out) {
  const size_t n = (size_t)w * h;
  float* d;
  cudaMalloc((void**)&d, 4 * n);
  cudaMemcpy(d, h_in, 4 * n, cudaMemcpyHostToDevice);
  os::sensor::subsample<float>(d, w, h);
  cudaMemcpy(h_out, d, 4 * (n / 4), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return (int)cudaGetLastError();
}

int ref_normal_map(const float* h_vtx, int w, int h, float* h_nrm) {
  const size_t n = (size_t)w * h;
  glm::vec3 *d_v, *d_n;
  cudaMalloc((void**)&d_v, 12 * n); cudaMalloc((void**)&d_n, 12 * n);
  cudaMemcpy(d_v, h_vtx, 12 * n, cudaMemcpyHostToDevice);
  os::sensor::generateNormalMap(d_v, d_n, w, h);
  cudaMemcpy(h_nrm, d_n, 12 * n, cudaMemcpyDeviceToHost);
  cudaFree(d_v); cudaFree(d_n);
  return (int)cudaGetLastError();
}

int ref_transform_normals(float* h_nrm, int n, const float m[16]) {
  glm::vec3* d;
  cudaMalloc((void**)&d, 12 * (size_t)n);
  cudaMemcpy(d, h_nrm, 12 * (size_t)n, cudaMemcpyHostToDevice);
  os::sensor::transformNormalMap(d, mat_from(m), n);
  cudaDeviceSynchronize();
  cudaMemcpy(h_nrm, d, 12 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return (int)cudaGetLastError();
}

int ref_color_to_intensity(const uint8_t* h_rgb, int n, float* h_out) {
  Color256* d_c; float* d_o;
  cudaMalloc((void**)&d_c, 3 * (size_t)n); cudaMalloc((void**)&d_o, 4 * (size_t)n);
  cudaMemcpy(d_c, h_rgb, 3 * (size_t)n, cudaMemcpyHostToDevice);
  os::sensor::colorToIntensity(d_c, d_o, n);
  cudaMemcpy(h_out, d_o, 4 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d_c); cudaFree(d_o);
  return (int)cudaGetLastError();
}

// computeICPCost2 on host copies of the four maps
int ref_icp_cost2(const float* last_v, const float* last_n, const float* this_v, const float* this_n, int w, int h,
                  float A[36], float b[6]) {
  const size_t bytes = 12 * (size_t)w * h;
  os::sensor::ICPFrame last(w, h), cur(w, h);
  cudaMemcpy(last.vertex, last_v, bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(last.normal, last_n, bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(cur.vertex, this_v, bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(cur.normal, this_n, bytes, cudaMemcpyHostToDevice);
  os::sensor::computeICPCost2(&last, cur, A, b);
  return (int)cudaGetLastError();
}

// RGBDCamera (rgbd_camera.cpp), driven with host depth frames
struct ref_tracker {
  os::sensor::RGBDCamera* cam;
  RawFrame* frame;
  long long stamp;
};

ref_tracker* ref_tracker_create(int w, int h, float fx, float fy) {
  ref_tracker* t = new ref_tracker();
  // the constructor leaves latest_stamp_ uninitialised (rgbd_camera.cpp:22-24) and update() drops frames whose
  // timestamp is not newer: construct into zeroed storage so that the timestamps 1, 2, ... are always accepted
  void* mem = calloc(1, sizeof(os::sensor::RGBDCamera));
  t->cam = new (mem) os::sensor::RGBDCamera(w, h, glm::vec2(fx, fy));
  t->frame = new RawFrame(w, h);
  t->stamp = 0;
  return t;
}

void ref_tracker_destroy(ref_tracker* t) {
  if (!t) return;
  t->cam->~RGBDCamera();
  free(t->cam);
  delete t->frame;
  delete t;
}

double ref_tracker_update(ref_tracker* t, const uint16_t* h_depth) {
  cudaMemcpy(t->frame->depth, h_depth, 2 * (size_t)t->frame->width * t->frame->height, cudaMemcpyHostToDevice);
  cudaMemset(t->frame->color, 0, 3 * (size_t)t->frame->width * t->frame->height);
  t->frame->timestamp = ++t->stamp;
  cudaDeviceSynchronize();
  const double t0 = now_ms();
  t->cam->update(t->frame);
  cudaDeviceSynchronize();
  return now_ms() - t0;
}

void ref_tracker_pose(const ref_tracker* t, float position[3], float orientation[9]) {
  const glm::vec3 p = t->cam->position();
  const glm::mat3 o = t->cam->orientation();
  position[0] = p.x; position[1] = p.y; position[2] = p.z;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) orientation[3 * c + r] = o[c][r];
}

}  // extern "C"
