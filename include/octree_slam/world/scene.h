// world::Scene, hot-path methods of include/octree_slam/world/scene.h:22-78 (mesh/texture loading and voxelpipe
// voxelisation are out of scope).
#ifndef OSL_B200_SCENE_H_
#define OSL_B200_SCENE_H_
#include <octree_slam/common_types.h>
#include <octree_slam/world/octree.h>

namespace octree_slam {
namespace world {

class Scene {
 public:
  Scene();
  ~Scene();
  // scene.cpp:26-33 without the OBJ/BMP file readers: the caller hands over a loaded mesh (and texture)
  void addMesh(const Mesh& mesh, const bmp_texture& texture) { meshes_ = &mesh; textures_ = &texture; }
  void voxelizeMeshes(const bool octree);  // scene.cpp:64-85
  void extractVoxelGridFromOctree();
  void addPointCloudToOctree(const glm::vec3& origin, const glm::vec3* points, const Color256* colors, const int size,
                             const BoundingBox& bbox);
  const VoxelGrid& voxel_grid() const { return *voxel_grid_; }
  SVO svo(const BoundingBox& bbox) const { return tree_->extractSVO(bbox); }
  Octree* tree() { return tree_; }

 private:
  VoxelGrid* voxel_grid_;
  Octree* tree_;
  const Mesh* meshes_ = nullptr;
  const bmp_texture* textures_ = nullptr;
};

}  // namespace world
}  // namespace octree_slam
#endif
