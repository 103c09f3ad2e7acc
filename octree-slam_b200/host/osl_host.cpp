// osl_host.cpp -- C++ host side above the C ABI: the reference's world / rendering / sensor interface for the hot
// path (same names, argument meaning and ownership rules), implemented as thin forwards to libosl_b200.so.
// Replaces src/world/octree.cpp:251-389 (Octree), src/world/scene.cpp:87-113 (Scene), the callers' view of
// src/world/svo/svo.cu:584-745, src/rendering/cone_tracing_kernels.cu:157-198, src/rendering/cuda_renderer.cpp:158-171
// and src/sensor/image_kernels.cu:55-58,96-102,217-219.  No CUDA code here: device memory is handled by the C ABI
// and by cudart's C API (cudaMalloc/cudaFree/cudaMemcpy) for the buffers the reference's structs own.
#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdio>
#include <vector>
#include <map>
#include <mutex>

#include <octree_slam/common_types.h>
#include <octree_slam/rendering/cone_tracing_kernels.h>
#include <octree_slam/rendering/cuda_renderer.h>
#include <octree_slam/sensor/image_kernels.h>
#include <octree_slam/sensor/localization_kernels.h>
#include <octree_slam/sensor/rgbd_camera.h>
#include <octree_slam/world/octree.h>
#include <octree_slam/world/scene.h>
#include <octree_slam/world/svo/svo.h>
#include <octree_slam/world/voxelization/voxelization.h>

#include "osl_b200.h"

#include <type_traits>
// The seam takes glm types and SVO by value; with glm 0.9.5.4 they are "non-trivial for the purposes of calls" (passed
// by invisible reference).  Whatever glm this file is compiled against must agree, or the symbols would link and the
// arguments would arrive garbled.
static_assert(!std::is_trivially_copy_constructible<glm::vec2>::value && !std::is_trivially_copy_constructible<glm::vec3>::value &&
              !std::is_trivially_copy_constructible<glm::mat4>::value && !std::is_trivially_copy_constructible<SVO>::value,
              "glm types must have glm 0.9.5's calling convention (user-provided copy constructors)");
static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::mat4) == 64 && sizeof(SVO) == 24 && sizeof(Color256) == 3 &&
              sizeof(BoundingBox) == 24, "layout contract of the reference's POD types");

// ---- common_types.cu:8-52 --------------------------------------------------------------------------------------
bool BoundingBox::contains(const BoundingBox& o) const {
  return bbox0.x <= o.bbox0.x && bbox0.y <= o.bbox0.y && bbox0.z <= o.bbox0.z && bbox1.x >= o.bbox1.x &&
         bbox1.y >= o.bbox1.y && bbox1.z >= o.bbox1.z;
}
// largest overhang of THIS box beyond `o` along any axis (0 when inside), common_types.cu:22-34
float BoundingBox::distanceOutside(const BoundingBox& o) const {
  float d = 0.0f;
  const float lo[3] = {o.bbox0.x - bbox0.x, o.bbox0.y - bbox0.y, o.bbox0.z - bbox0.z};
  const float hi[3] = {bbox1.x - o.bbox1.x, bbox1.y - o.bbox1.y, bbox1.z - o.bbox1.z};
  for (int i = 0; i < 3; i++) d = std::fmax(d, std::fmax(lo[i], hi[i]));
  return d;
}
RawFrame::RawFrame(const int w, const int h) : height(h), width(w) {
  cudaMalloc((void**)&color, (size_t)h * w * sizeof(Color256));
  cudaMalloc((void**)&depth, (size_t)h * w * sizeof(uint16_t));
  timestamp = 0;
}
RawFrame::~RawFrame() {
  cudaFree(color);
  cudaFree(depth);
}
VoxelGrid::~VoxelGrid() {
  if (size > 0) {
    cudaFree(centers);
    cudaFree(colors);
  }
}

// column-major float[16] view of a mat4; works with the reference's glm 0.9.5.4 (where value_ptr lives in a gtc
// header main.cpp's translation units do not all include) and with include/glm/glm.hpp alike
static inline const float* mat_ptr(const glm::mat4& m) { return &m[0].x; }

static void report(osl_status s, const char* where) {
  // the reference returns void and checks nothing on this path; errors are reported, never fatal
  if (s != OSL_OK) fprintf(stderr, "[octree_slam] %s: %s (cuda %d)\n", where, osl_status_string(s), osl_last_cuda_error());
}

namespace octree_slam {

// ---- svo.h -----------------------------------------------------------------------------------------------------
namespace svo {

// pool pointer -> owning handle (the reference hands raw pool pointers around)
static std::map<const void*, osl_svo*>& registry() {
  static std::map<const void*, osl_svo*> r;
  return r;
}
static std::mutex& reg_mutex() {
  static std::mutex m;
  return m;
}

static osl_svo* lookup_or_create(unsigned int* octree, int octree_size, glm::vec3 c, float edge, int max_depth) {
  std::lock_guard<std::mutex> g(reg_mutex());
  if (octree_size != 0 && octree) {
    auto it = registry().find(octree);
    if (it != registry().end()) return it->second;
    fprintf(stderr, "[octree_slam] unknown pool pointer %p: pools are created by svoFrom*() with octree_size == 0\n",
            (void*)octree);
    return nullptr;
  }
  osl_svo* t = nullptr;
  const float cc[3] = {c.x, c.y, c.z};
  int dev = 0;
  cudaGetDevice(&dev);
  report(osl_svo_create(&t, cc, edge, max_depth, 0, dev), "osl_svo_create");
  return t;
}

static void publish(osl_svo* t, unsigned int* old_ptr, unsigned int*& octree, int& octree_size) {
  const uint32_t* pool = nullptr;
  int n = 0;
  report(osl_svo_view(t, &pool, &n, nullptr, nullptr), "osl_svo_view");
  std::lock_guard<std::mutex> g(reg_mutex());
  if (old_ptr) registry().erase(old_ptr);
  registry()[pool] = t;
  octree = const_cast<unsigned int*>(pool);
  octree_size = n;
}

void svoFromPointCloud(const glm::vec3* points, const Color256* colors, const int size, const int max_depth,
                       unsigned int*& octree, int& octree_size, glm::vec3 octree_center, const float edge_length,
                       void*) {
  osl_svo* t = lookup_or_create(octree, octree_size, octree_center, edge_length, max_depth);
  if (!t) return;
  report(osl_integrate_points(t, &points->x, &colors->r, size, nullptr), "osl_integrate_points");
  publish(t, octree_size ? octree : nullptr, octree, octree_size);
}

void svoFromVoxelGrid(const VoxelGrid& grid, const int max_depth, unsigned int*& octree, int& octree_size,
                      glm::vec3 octree_center, const float edge_length, void*) {
  osl_svo* t = lookup_or_create(octree, octree_size, octree_center, edge_length, max_depth);
  if (!t) return;
  report(osl_integrate_voxels(t, &grid.centers->x, &grid.colors->x, grid.size, nullptr), "osl_integrate_voxels");
  publish(t, octree_size ? octree : nullptr, octree, octree_size);
}

void svoFromDepthFrame(const uint16_t* depth, const Color256* colors, int width, int height, glm::vec2 focal_length,
                       const glm::mat4& pose, const int max_depth, unsigned int*& octree, int& octree_size,
                       glm::vec3 octree_center, const float edge_length) {
  osl_svo* t = lookup_or_create(octree, octree_size, octree_center, edge_length, max_depth);
  if (!t) return;
  report(osl_integrate_depth(t, depth, &colors->r, width, height, focal_length.x, focal_length.y,
                             mat_ptr(pose), nullptr), "osl_integrate_depth");
  publish(t, octree_size ? octree : nullptr, octree, octree_size);
}

void extractVoxelGridFromSVO(unsigned int*& octree, int& octree_size, const int max_depth, const glm::vec3,
                             float, VoxelGrid& grid) {
  osl_svo* t = nullptr;
  {
    std::lock_guard<std::mutex> g(reg_mutex());
    auto it = registry().find(octree);
    if (it != registry().end()) t = it->second;
  }
  if (!t || octree_size == 0) { grid.size = 0; return; }
  int64_t n = 0;
  report(osl_extract_voxels(t, max_depth, nullptr, nullptr, nullptr, 0, &n, nullptr), "osl_extract_voxels");
  grid.size = (int)n;
  cudaMalloc((void**)&grid.centers, (size_t)(n > 0 ? n : 1) * sizeof(glm::vec4));  // svo.cu:732-733
  cudaMalloc((void**)&grid.colors, (size_t)(n > 0 ? n : 1) * sizeof(glm::vec4));
  if (n > 0)
    report(osl_extract_voxels(t, max_depth, &grid.centers->x, &grid.colors->x, nullptr, n, &n, nullptr),
           "osl_extract_voxels");
}

void releaseSVO(unsigned int* octree) {
  osl_svo* t = nullptr;
  {
    std::lock_guard<std::mutex> g(reg_mutex());
    auto it = registry().find(octree);
    if (it == registry().end()) return;
    t = it->second;
    registry().erase(it);
  }
  osl_svo_destroy(t);
}

}  // namespace svo

// ---- voxelization (voxelization.cu:90-139 ColorShader, :219-236 createVoxelGrid, :381-405 meshToVoxelGrid) --------
namespace voxelization {

static inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

// per-triangle flat colour, the reference's ColorShader (voxelization.cu:90-139): no texture -> green; no texture
// coordinates -> texel 0; else the texel at the FIRST corner ("TODO: interpolate", voxelization.cu:125-126); quantised
// to 8 bits and returned as byte / 255.0 (createVoxelGrid, voxelization.cu:219-236)
static void triangleColors(const Mesh& m_in, const bmp_texture* tex, std::vector<float>& tri_col) {
  const int n_tri = m_in.ibosize / 3;
  tri_col.assign((size_t)n_tri * 4, 0.0f);
  for (int t = 0; t < n_tri; t++) {
    int r = 0, g = 255, b = 0;
    if (tex && tex->width > 0 && tex->data) {
      glm::vec3 c = tex->data[0];
      if (m_in.tbosize > 0 && m_in.tbo) {
        int tx = (int)(m_in.tbo[6 * t] * tex->width), ty = (int)(m_in.tbo[6 * t + 1] * tex->height);
        tx = tx < 0 ? 0 : (tx >= tex->width ? tex->width - 1 : tx);     // (the reference indexes unchecked)
        ty = ty < 0 ? 0 : (ty >= tex->height ? tex->height - 1 : ty);
        c = tex->data[ty * tex->width + tx];
        r = (int)(clamp01(c.x) * 255.0f); g = (int)(clamp01(c.y) * 255.0f); b = (int)(clamp01(c.z) * 255.0f);
      } else {
        r = (int)(c.x * 255.0); g = (int)(c.y * 255.0); b = (int)(c.z * 255.0);
      }
    }
    tri_col[4 * t] = (float)((r & 0xFF) / 255.0); tri_col[4 * t + 1] = (float)((g & 0xFF) / 255.0);
    tri_col[4 * t + 2] = (float)((b & 0xFF) / 255.0);
  }
}

void meshToVoxelGridAt(const Mesh& m_in, const bmp_texture* tex, const glm::vec3& center, float half_edge, int depth,
                       VoxelGrid& grid_out) {
  const int n_tri = m_in.ibosize / 3, n_vert = m_in.vbosize / 3;
  std::vector<float> tri_col;
  triangleColors(m_in, tex, tri_col);
  float *d_v = nullptr, *d_c = nullptr;
  int* d_i = nullptr;
  cudaMalloc((void**)&d_v, sizeof(float) * 3 * (size_t)(n_vert > 0 ? n_vert : 1));
  cudaMalloc((void**)&d_i, sizeof(int) * 3 * (size_t)(n_tri > 0 ? n_tri : 1));
  cudaMalloc((void**)&d_c, sizeof(float) * 4 * (size_t)(n_tri > 0 ? n_tri : 1));
  cudaMemcpy(d_v, m_in.vbo, sizeof(float) * 3 * (size_t)n_vert, cudaMemcpyHostToDevice);
  cudaMemcpy(d_i, m_in.ibo, sizeof(int) * 3 * (size_t)n_tri, cudaMemcpyHostToDevice);
  cudaMemcpy(d_c, tri_col.data(), sizeof(float) * 4 * (size_t)n_tri, cudaMemcpyHostToDevice);
  if (grid_out.size > 0) { cudaFree(grid_out.centers); cudaFree(grid_out.colors); grid_out.size = 0; }
  float *centers = nullptr, *colors = nullptr;
  int64_t n = 0;
  const float c3[3] = {center.x, center.y, center.z};
  report(osl_voxelize_mesh(d_v, n_vert, d_i, n_tri, d_c, c3, half_edge, depth, &centers, &colors, nullptr, nullptr, &n,
                           nullptr), "osl_voxelize_mesh");
  cudaFree(d_v); cudaFree(d_i); cudaFree(d_c);
  grid_out.centers = reinterpret_cast<glm::vec4*>(centers);
  grid_out.colors = reinterpret_cast<glm::vec4*>(colors);
  grid_out.size = (int)n;
  grid_out.scale = (2.0f * half_edge) / (float)(1 << depth);
  grid_out.bbox = m_in.bbox;
}

// The reference signature = the reference's rule (voxelization.cu:381-405): voxelpipe THIN_RASTER on the dense
// 2^log_N grid over the MESH bounding box; scale and bbox as the reference sets them (voxelization.cu:401-405).  The
// voxels come out ordered by their Morton keys in the cube Scene::voxelizeMeshes builds its Octree on (scene.cpp:78),
// so that svoFromVoxelGrid's colour quirk Q11 is harmless.
void meshToVoxelGrid(const Mesh& m_in, const bmp_texture* tex, VoxelGrid& grid_out) {
  const int n_tri = m_in.ibosize / 3, n_vert = m_in.vbosize / 3;
  std::vector<float> tri_col;
  triangleColors(m_in, tex, tri_col);
  float *d_v = nullptr, *d_c = nullptr;
  int* d_i = nullptr;
  cudaMalloc((void**)&d_v, sizeof(float) * 3 * (size_t)(n_vert > 0 ? n_vert : 1));
  cudaMalloc((void**)&d_i, sizeof(int) * 3 * (size_t)(n_tri > 0 ? n_tri : 1));
  cudaMalloc((void**)&d_c, sizeof(float) * 4 * (size_t)(n_tri > 0 ? n_tri : 1));
  cudaMemcpy(d_v, m_in.vbo, sizeof(float) * 3 * (size_t)n_vert, cudaMemcpyHostToDevice);
  cudaMemcpy(d_i, m_in.ibo, sizeof(int) * 3 * (size_t)n_tri, cudaMemcpyHostToDevice);
  cudaMemcpy(d_c, tri_col.data(), sizeof(float) * 4 * (size_t)n_tri, cudaMemcpyHostToDevice);
  if (grid_out.size > 0) { cudaFree(grid_out.centers); cudaFree(grid_out.colors); grid_out.size = 0; }
  float *centers = nullptr, *colors = nullptr;
  int64_t n = 0;
  const float b0[3] = {m_in.bbox.bbox0.x, m_in.bbox.bbox0.y, m_in.bbox.bbox0.z};
  const float b1[3] = {m_in.bbox.bbox1.x, m_in.bbox.bbox1.y, m_in.bbox.bbox1.z};
  const glm::vec3 mid = (m_in.bbox.bbox1 + m_in.bbox.bbox0) / 2.0f;
  const float c3[3] = {mid.x, mid.y, mid.z};
  report(osl_voxelize_thin(d_v, n_vert, d_i, n_tri, d_c, b0, b1, log_N(), c3, m_in.bbox.bbox1.x, log_N(), &centers,
                           &colors, nullptr, nullptr, &n, nullptr), "osl_voxelize_thin");
  cudaFree(d_v); cudaFree(d_i); cudaFree(d_c);
  grid_out.centers = reinterpret_cast<glm::vec4*>(centers);
  grid_out.colors = reinterpret_cast<glm::vec4*>(colors);
  grid_out.size = (int)n;
  grid_out.scale = (m_in.bbox.bbox1.x - m_in.bbox.bbox0.x) / (float)(1 << log_N()) / 2.0f;  // computeScale, voxelization.cu:80-82
  grid_out.bbox = m_in.bbox;
}

}  // namespace voxelization

// ---- world::Octree (octree.cpp:251-389) ------------------------------------------------------------------------
namespace world {

Octree::Octree(const float resolution, const glm::vec3& center, const float size)
    : svo_(nullptr), center_(center), size_(size), resolution_(resolution) {}

Octree::~Octree() {
  if (svo_) osl_svo_destroy(svo_);
}

int Octree::maxDepth(float resolution) const {
  // octree.cpp:283-284 with node_depth = 0 (the root sub-tree is the GPU tree).  Under g++ the reference's
  // unqualified log() resolves to the double overload; the quotient is taken in double, ceil'd and truncated.
  const float edge_length = size_ / std::pow(2.0f, (float)0);
  return (int)std::ceil(std::log((double)(float)(edge_length / resolution)) / (double)std::log(2.0f));
}

osl_svo* Octree::tree(int max_depth) {
  if (!svo_) {
    const float c[3] = {center_.x, center_.y, center_.z};
    int dev = 0;
    cudaGetDevice(&dev);
    report(osl_svo_create(&svo_, c, size_, max_depth, 0, dev), "osl_svo_create");
  }
  return svo_;
}

void Octree::addCloud(const glm::vec3&, const glm::vec3* points, const Color256* colors, const int size,
                      const BoundingBox&) {
  osl_svo* t = tree(maxDepth(resolution_));
  if (t) report(osl_integrate_points(t, &points->x, &colors->r, size, nullptr), "osl_integrate_points");
}

void Octree::addDepthFrame(const uint16_t* depth, const Color256* colors, int width, int height,
                           glm::vec2 focal_length, const glm::mat4& pose) {
  osl_svo* t = tree(maxDepth(resolution_));
  if (t)
    report(osl_integrate_depth(t, depth, &colors->r, width, height, focal_length.x, focal_length.y,
                               mat_ptr(pose), nullptr), "osl_integrate_depth");
}

void Octree::addDepthFrame(const uint16_t* depth, const Color256* colors, int width, int height,
                           glm::vec2 focal_length, const sensor::RGBDCamera& camera) {
  osl_svo* t = tree(maxDepth(resolution_));
  const float* d_pose = camera.poseDevice();
  if (t && d_pose)
    report(osl_integrate_depth_posed(t, depth, &colors->r, width, height, focal_length.x, focal_length.y, d_pose,
                                     nullptr), "osl_integrate_depth_posed");
}

void Octree::addVoxelGrid(const VoxelGrid& grid) {
  osl_svo* t = tree(maxDepth(resolution_));
  if (t) report(osl_integrate_voxels(t, &grid.centers->x, &grid.colors->x, grid.size, nullptr), "osl_integrate_voxels");
}

void Octree::extractVoxelGrid(VoxelGrid& grid) {
  if (!svo_) { grid.size = 0; return; }
  const int max_depth = maxDepth(grid.scale);  // octree.cpp:330
  int64_t n = 0;
  report(osl_extract_voxels(svo_, max_depth, nullptr, nullptr, nullptr, 0, &n, nullptr), "osl_extract_voxels");
  grid.size = (int)n;
  cudaMalloc((void**)&grid.centers, (size_t)(n > 0 ? n : 1) * sizeof(glm::vec4));
  cudaMalloc((void**)&grid.colors, (size_t)(n > 0 ? n : 1) * sizeof(glm::vec4));
  if (n > 0)
    report(osl_extract_voxels(svo_, max_depth, &grid.centers->x, &grid.colors->x, nullptr, n, &n, nullptr),
           "osl_extract_voxels");
}

SVO Octree::extractSVO(const BoundingBox&) {
  SVO out;
  out.data = nullptr;
  out.center = center_;
  out.size = size_;  // size_/2^node_depth with node_depth = 0 (octree.cpp:357)
  if (svo_) {
    const uint32_t* pool = nullptr;
    report(osl_svo_view(svo_, &pool, nullptr, nullptr, nullptr), "osl_svo_view");
    out.data = const_cast<unsigned int*>(pool);
  }
  return out;
}

BoundingBox Octree::boundingBox() const {
  BoundingBox box;
  box.bbox0 = center_ - glm::vec3(size_, size_, size_);
  box.bbox1 = center_ + glm::vec3(size_, size_, size_);
  return box;
}

void Octree::expandBySize(const float add_size) {
  // octree.cpp:362-378.  Quirk Q10: the reference re-scales size_ although OctreeNode::expand() refuses GPU-backed
  // nodes, silently corrupting the map; here the GPU tree is re-rooted (osl_svo_expand) by enough doublings to hold
  // size_ + add_size, keeping centre and resolution.
  if (!(add_size > 0.0f)) return;
  const int add_layers = (int)std::ceil(std::log2((double)((size_ + add_size) / size_)));
  if (add_layers < 1) return;
  if (svo_) {
    const osl_status rc = osl_svo_expand(svo_, add_layers);
    report(rc, "osl_svo_expand");
    if (rc != OSL_OK) return;
  }
  size_ = std::pow(2.0f, (float)add_layers) * size_;
}

int Octree::nodeCount() const { return svo_ ? osl_svo_size(svo_) : 0; }

// ---- world::Scene (scene.cpp:87-113) ---------------------------------------------------------------------------
Scene::Scene() : voxel_grid_(new VoxelGrid()), tree_(nullptr) {}

Scene::~Scene() {
  delete tree_;
  delete voxel_grid_;
}

void Scene::voxelizeMeshes(const bool octree) {  // scene.cpp:64-85
  if (!meshes_) return;
  const Mesh& m = *meshes_;
  const float scale = m.bbox.bbox1.x / (float)(1 << voxelization::log_N());
  if (!octree) {
    voxelization::meshToVoxelGrid(m, textures_, *voxel_grid_);
    voxel_grid_->scale = scale;
  } else {
    VoxelGrid grid;
    voxelization::meshToVoxelGrid(m, textures_, grid);
    voxel_grid_->scale = scale;
    if (!tree_) tree_ = new Octree(scale, (m.bbox.bbox1 + m.bbox.bbox0) / 2.0f, m.bbox.bbox1.x);
    tree_->addVoxelGrid(grid);
    voxel_grid_->bbox = tree_->boundingBox();
    tree_->extractVoxelGrid(*voxel_grid_);
  }
}

void Scene::extractVoxelGridFromOctree() {
  delete voxel_grid_;
  voxel_grid_ = new VoxelGrid();
  if (!tree_) return;
  voxel_grid_->bbox = tree_->boundingBox();
  voxel_grid_->scale = 0.01f;  // scene.cpp:94
  tree_->extractVoxelGrid(*voxel_grid_);
}

void Scene::addPointCloudToOctree(const glm::vec3& origin, const glm::vec3* points, const Color256* colors,
                                  const int size, const BoundingBox& bbox) {
  if (!tree_) {
    // scene.cpp:100-102: resolution 0.01, centre = bbox mid-point, size = bbox.bbox1.x (sic, quirk Q10)
    tree_ = new Octree(0.01f, (bbox.bbox1 + bbox.bbox0) / 2.0f, bbox.bbox1.x);
  } else if (!tree_->boundingBox().contains(bbox)) {
    tree_->expandBySize(0.0f);
  }
  tree_->addCloud(origin, points, colors, size, bbox);
}

}  // namespace world

// ---- rendering (cone_tracing_kernels.cu:157-198, cuda_renderer.cpp:158-171) --------------------------------------
namespace rendering {

void coneTraceSVO(uchar4* pos, glm::vec2 resolution, float fov, glm::mat4 cameraPose, SVO octree) {
  const float c[3] = {octree.center.x, octree.center.y, octree.center.z};
  report(osl_raycast_pool(octree.data, c, octree.size, &pos->x, (int)resolution.x, (int)resolution.y, fov,
                          mat_ptr(cameraPose), nullptr, nullptr, nullptr), "osl_raycast_pool");
  cudaDeviceSynchronize();  // the reference call is synchronous on return
}

CUDARenderer::CUDARenderer(const int width, const int height) : width_(width), height_(height), d_pixels_(nullptr) {
  cudaMalloc((void**)&d_pixels_, (size_t)width * height * sizeof(uchar4));
}
CUDARenderer::~CUDARenderer() { cudaFree(d_pixels_); }

void CUDARenderer::coneTraceSVO(const SVO& octree, const Camera& camera, const glm::vec3&) {
  rendering::coneTraceSVO(d_pixels_, glm::vec2((float)width_, (float)height_), camera.fov, camera.view, octree);
}

void CUDARenderer::download(uchar4* host_pixels) const {
  cudaMemcpy(host_pixels, d_pixels_, (size_t)width_ * height_ * sizeof(uchar4), cudaMemcpyDeviceToHost);
}

}  // namespace rendering

// ---- sensor (image_kernels.cu:55-58, 96-102, 217-219) ------------------------------------------------------------
namespace sensor {

void generateVertexMap(const uint16_t* depth_pixels, glm::vec3* vertex_map, const int width, const int height,
                       const glm::vec2 focal_length, const int2 img_size) {
  report(osl_generate_vertex_map(depth_pixels, &vertex_map->x, width, height, focal_length.x, focal_length.y,
                                 img_size.x, img_size.y, nullptr), "osl_generate_vertex_map");
  cudaDeviceSynchronize();  // image_kernels.cu:57
}

void transformVertexMap(glm::vec3* vertex_map, const glm::mat4& trans, const int size) {
  report(osl_transform_vertex_map(&vertex_map->x, mat_ptr(trans), size, nullptr), "osl_transform_vertex_map");
}

void computePointCloudBoundingBox(glm::vec3* points, const int num_points, BoundingBox& bbox) {
  float b[6] = {bbox.bbox0.x, bbox.bbox0.y, bbox.bbox0.z, bbox.bbox1.x, bbox.bbox1.y, bbox.bbox1.z};
  report(osl_point_cloud_bbox(&points->x, num_points, b, nullptr), "osl_point_cloud_bbox");
  bbox.bbox0 = glm::vec3(b[0], b[1], b[2]);
  bbox.bbox1 = glm::vec3(b[3], b[4], b[5]);
}

// ---- camera tracking (image_kernels.cu:104-321, localization_kernels.cu, rgbd_camera.cpp) -----------------------
void generateNormalMap(const glm::vec3* vertex_map, glm::vec3* normal_map, const int width, const int height) {
  report(osl_generate_normal_map(&vertex_map->x, &normal_map->x, width, height, nullptr), "osl_generate_normal_map");
  cudaDeviceSynchronize();  // image_kernels.cu:134
}

void bilateralFilter(const uint16_t* depth_in, uint16_t* filtered_out, const int width, const int height) {
  report(osl_bilateral_filter(depth_in, filtered_out, width, height, nullptr), "osl_bilateral_filter");
  cudaDeviceSynchronize();  // image_kernels.cu:175
}

void colorToIntensity(const Color256* color_in, float* intensity_out, const int size) {
  report(osl_color_to_intensity(&color_in->r, intensity_out, size, nullptr), "osl_color_to_intensity");
  cudaDeviceSynchronize();  // image_kernels.cu:191
}

void transformNormalMap(glm::vec3* normal_map, const glm::mat4& trans, const int size) {
  report(osl_transform_normal_map(&normal_map->x, mat_ptr(trans), size, nullptr), "osl_transform_normal_map");
}

// image_kernels.cu:262-277 / 299-313: temporary of a quarter of the size, kernel, copy back over the input
template <>
void subsampleDepth<uint16_t>(uint16_t* data, const int width, const int height) {
  const size_t bytes = sizeof(uint16_t) * (size_t)(width / 2) * (size_t)(height / 2);
  uint16_t* tmp = nullptr;
  if (cudaMalloc((void**)&tmp, bytes ? bytes : 1) != cudaSuccess) return report(OSL_ERR_OOM, "subsampleDepth");
  report(osl_subsample_depth(data, tmp, width, height, nullptr), "osl_subsample_depth");
  cudaMemcpy(data, tmp, bytes, cudaMemcpyDeviceToDevice);
  cudaFree(tmp);
}

template <>
void subsample<float>(float* data, const int width, const int height) {
  const size_t bytes = sizeof(float) * (size_t)(width / 2) * (size_t)(height / 2);
  float* tmp = nullptr;
  if (cudaMalloc((void**)&tmp, bytes ? bytes : 1) != cudaSuccess) return report(OSL_ERR_OOM, "subsample");
  report(osl_subsample_f32(data, tmp, width, height, nullptr), "osl_subsample_f32");
  cudaMemcpy(data, tmp, bytes, cudaMemcpyDeviceToDevice);
  cudaFree(tmp);
}

ICPFrame::ICPFrame(const int w, const int h) : vertex(nullptr), normal(nullptr), width(w), height(h) {
  cudaMalloc((void**)&vertex, (size_t)w * h * sizeof(glm::vec3));
  cudaMalloc((void**)&normal, (size_t)w * h * sizeof(glm::vec3));
}

ICPFrame::~ICPFrame() {
  cudaFree(vertex);
  cudaFree(normal);
}

RGBDFrame::RGBDFrame(const int w, const int h) : intensity(nullptr), vertex(nullptr), width(w), height(h) {
  cudaMalloc((void**)&intensity, (size_t)w * h * sizeof(float));
  cudaMalloc((void**)&vertex, (size_t)w * h * sizeof(glm::vec3));
}

RGBDFrame::~RGBDFrame() {
  cudaFree(intensity);
  cudaFree(vertex);
}

void computeICPCost2(const ICPFrame* last_frame, const ICPFrame& this_frame, float* A, float* b) {
  report(osl_icp_cost(&last_frame->vertex->x, &last_frame->normal->x, &this_frame.vertex->x, &this_frame.normal->x,
                      this_frame.width * this_frame.height, 0, A, b, nullptr, nullptr), "osl_icp_cost");
}

void computeICPCost(const ICPFrame* last_frame, const ICPFrame& this_frame, float* A, float* b) {
  computeICPCost2(last_frame, this_frame, A, b);
}

void computeRGBDCost(const RGBDFrame*, const RGBDFrame&, float*, float*) { cudaDeviceSynchronize(); }

RGBDCamera::RGBDCamera(const int width, const int height, const glm::vec2& focal_length, const bool exact_jacobian)
    : tracker_(nullptr), focal_length_(focal_length), width_(width), height_(height), latest_stamp_(-1) {
  int dev = 0;
  cudaGetDevice(&dev);
  report(osl_tracker_create(&tracker_, width, height, focal_length.x, focal_length.y, exact_jacobian ? 1 : 0, dev),
         "osl_tracker_create");
}

RGBDCamera::~RGBDCamera() { osl_tracker_destroy(tracker_); }

void RGBDCamera::update(const RawFrame* this_frame) {
  if (!tracker_ || !this_frame) return;
  if (this_frame->timestamp <= latest_stamp_) return;  // rgbd_camera.cpp:55-59
  latest_stamp_ = this_frame->timestamp;
  if (this_frame->width != width_ || this_frame->height != height_)
    return report(OSL_ERR_INVALID, "RGBDCamera::update (frame size)");
  report(osl_tracker_update(tracker_, this_frame->depth, nullptr), "osl_tracker_update");
}

const glm::vec3 RGBDCamera::position() const {
  glm::vec3 p;
  if (tracker_) report(osl_tracker_get_pose(tracker_, nullptr, &p.x, nullptr, nullptr, nullptr), "osl_tracker_get_pose");
  return p;
}

const glm::mat3 RGBDCamera::orientation() const {
  float o[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (tracker_) report(osl_tracker_get_pose(tracker_, nullptr, nullptr, o, nullptr, nullptr), "osl_tracker_get_pose");
  glm::mat3 m;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) m[c][r] = o[3 * c + r];
  return m;
}

const glm::mat4 RGBDCamera::pose() const {
  float p[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (tracker_) report(osl_tracker_get_pose(tracker_, p, nullptr, nullptr, nullptr, nullptr), "osl_tracker_get_pose");
  glm::mat4 m;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) m[c][r] = p[4 * c + r];
  return m;
}

const float* RGBDCamera::poseDevice() const {
  const float* p = nullptr;
  if (tracker_) report(osl_tracker_pose_device(tracker_, &p), "osl_tracker_pose_device");
  return p;
}

bool RGBDCamera::lost() const {
  int l = 0;
  if (tracker_) report(osl_tracker_get_pose(tracker_, nullptr, nullptr, nullptr, &l, nullptr), "osl_tracker_get_pose");
  return l != 0;
}

const Camera RGBDCamera::camera() const {  // rgbd_camera.cpp:40-51 (projection is a TODO there: left at identity)
  Camera cam;
  cam.model = glm::mat4(1.0f);
  cam.view = pose();
  cam.modelview = cam.view;  // view * identity
  cam.mvp = cam.view;        // identity projection * modelview
  return cam;
}

}  // namespace sensor
}  // namespace octree_slam
