// Drop-in replacement for the reference's include/octree_slam/world/svo/svo.h:14-18, forwarding to the C ABI
// (include/osl_b200.h).  Same names, argument order, meaning AND LINKAGE: the reference declares the three seam
// functions `extern "C"` inside the namespace (unmangled symbols svoFromVoxelGrid / svoFromPointCloud /
// extractVoxelGridFromSVO), so an object compiled against the reference's own header links against libosl_host.so
// (tests/test_host_shim.py links the reference's src/world/octree.cpp that way).  glm arguments by value follow
// glm 0.9.5's calling convention (non-trivial copy constructors -> passed by invisible reference), which
// include/glm/glm.hpp reproduces; the node pool is still handed back through
// `octree` / `octree_size`, but it is OWNED by the library: release it with svo::releaseSVO (the reference's
// owner, OctreeNode::~OctreeNode, called cudaFree on it -- octree.cpp:28-32).
#ifndef OSL_B200_SVO_H_
#define OSL_B200_SVO_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace svo {

extern "C" void svoFromVoxelGrid(const VoxelGrid& grid, const int max_depth, unsigned int*& octree, int& octree_size,
                      glm::vec3 octree_center, const float edge_length, void* d_bricks = nullptr);

extern "C" void svoFromPointCloud(const glm::vec3* points, const Color256* colors, const int size, const int max_depth,
                       unsigned int*& octree, int& octree_size, glm::vec3 octree_center, const float edge_length,
                       void* d_bricks = nullptr);

extern "C" void extractVoxelGridFromSVO(unsigned int*& octree, int& octree_size, const int max_depth, const glm::vec3 center,
                             float edge_length, VoxelGrid& grid);

// fused main.cpp:39-44 (generateVertexMap + transformVertexMap + svoFromPointCloud) -- new, optional
void svoFromDepthFrame(const uint16_t* depth, const Color256* colors, int width, int height, glm::vec2 focal_length,
                       const glm::mat4& pose, const int max_depth, unsigned int*& octree, int& octree_size,
                       glm::vec3 octree_center, const float edge_length);

void releaseSVO(unsigned int* octree);

inline int oppositeNode(const int node) { return -(~node); }  // svo.h:20-23

}  // namespace svo
}  // namespace octree_slam
#endif
