// rendering::CUDARenderer::coneTraceSVO (include/octree_slam/rendering/cuda_renderer.h:18-26,
// src/rendering/cuda_renderer.cpp:158-171) without the OpenGL plumbing: renders into a device buffer it owns.
#ifndef OSL_B200_CUDA_RENDERER_H_
#define OSL_B200_CUDA_RENDERER_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace rendering {

class CUDARenderer {
 public:
  CUDARenderer(const int width, const int height);
  ~CUDARenderer();
  void coneTraceSVO(const SVO& octree, const Camera& camera, const glm::vec3& light);
  const uchar4* devicePixels() const { return d_pixels_; }
  void download(uchar4* host_pixels) const;
  int width() const { return width_; }
  int height() const { return height_; }

 private:
  int width_, height_;
  uchar4* d_pixels_;
};

}  // namespace rendering
}  // namespace octree_slam
#endif
