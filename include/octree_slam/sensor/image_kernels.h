// Drop-in for the three per-frame functions of include/octree_slam/sensor/image_kernels.h:21,24,52.
#ifndef OSL_B200_IMAGE_KERNELS_H_
#define OSL_B200_IMAGE_KERNELS_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace sensor {
void generateVertexMap(const uint16_t* depth_pixels, glm::vec3* vertex_map, const int width, const int height,
                       const glm::vec2 focal_length, const int2 img_size);
void computePointCloudBoundingBox(glm::vec3* points, const int num_points, BoundingBox& bbox);
void transformVertexMap(glm::vec3* vertex_map, const glm::mat4& trans, const int size);
}  // namespace sensor
}  // namespace octree_slam
#endif
