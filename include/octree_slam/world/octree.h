// world::Octree with the reference's public interface (include/octree_slam/world/octree.h:83-126).  The CPU
// pointer-octree / paging machinery of OctreeNode is vestigial in the reference (the root never gets CPU children,
// SURVEY.md section 3.1) and is not reproduced: the tree is one GPU pool behind the C ABI.
#ifndef OSL_B200_OCTREE_H_
#define OSL_B200_OCTREE_H_
#include <octree_slam/common_types.h>

struct osl_svo;

namespace octree_slam {
namespace sensor { class RGBDCamera; }
namespace world {

class Octree {
 public:
  Octree(const float resolution, const glm::vec3& center, const float size);
  ~Octree();
  void addCloud(const glm::vec3& origin, const glm::vec3* points, const Color256* colors, const int size,
                const BoundingBox& bbox);
  void addVoxelGrid(const VoxelGrid& grid);
  void extractVoxelGrid(VoxelGrid& grid);
  SVO extractSVO(const BoundingBox& bbox);
  BoundingBox boundingBox() const;
  void expandBySize(const float add_size);  // re-roots the GPU tree (the reference cannot: quirk Q10)
  // fused main.cpp:39-44
  void addDepthFrame(const uint16_t* depth, const Color256* colors, int width, int height, glm::vec2 focal_length,
                     const glm::mat4& pose);
  // the same with the pose taken from `camera` ON THE DEVICE: camera.update(frame) must have been called for this
  // frame; nothing waits for the tracker on the host (osl_integrate_depth_posed)
  void addDepthFrame(const uint16_t* depth, const Color256* colors, int width, int height, glm::vec2 focal_length,
                     const sensor::RGBDCamera& camera);
  int nodeCount() const;

 private:
  osl_svo* tree(int max_depth);
  int maxDepth(float resolution) const;
  osl_svo* svo_;
  glm::vec3 center_;
  float size_;        // half edge length
  float resolution_;
};

}  // namespace world
}  // namespace octree_slam
#endif
