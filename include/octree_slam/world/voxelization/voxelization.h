// Drop-in for include/octree_slam/world/voxelization/voxelization.h:19-21 (meshToVoxelGrid only; voxelGridToMesh
// builds cube meshes for the OpenGL voxel view and is out of scope).  Sparse: works at any depth, not the reference's
// compile-time dense 256^3 (voxelization.cu:24).
#ifndef OSL_B200_VOXELIZATION_H_
#define OSL_B200_VOXELIZATION_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace voxelization {

inline int log_N() { return 8; }  // voxelization.cu:24 GRID_RES

// Reference signature = the reference's rule: voxelpipe THIN_RASTER on the dense 2^log_N() grid over the mesh bounding
// box (voxelization.cu:281-285; restated in octree-slam_b200/csrc/osl_voxelize_thin.cu), voxels ordered by their Morton
// keys in the cube Scene::voxelizeMeshes builds its Octree on (scene.cpp:78).
// (`extern "C"` as in the reference, voxelization.h:21)
extern "C" void meshToVoxelGrid(const Mesh& m_in, const bmp_texture* tex, VoxelGrid& grid_out);
// the sparse CONSERVATIVE voxeliser on an explicit cube / depth: every leaf cell the triangle overlaps (new; C++ linkage -- an extern "C" name cannot be overloaded)
void meshToVoxelGridAt(const Mesh& m_in, const bmp_texture* tex, const glm::vec3& center, float half_edge, int depth,
                       VoxelGrid& grid_out);

}  // namespace voxelization
}  // namespace octree_slam
#endif
