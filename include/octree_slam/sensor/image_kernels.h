// Drop-in for include/octree_slam/sensor/image_kernels.h:21-57 (the functions the reference defines; `gradient` and
// `difference` are declared there but defined nowhere in the reference).
#ifndef OSL_B200_IMAGE_KERNELS_H_
#define OSL_B200_IMAGE_KERNELS_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace sensor {
// `extern "C"` exactly where the reference has it (image_kernels.h:21,24,27,32,43,52,55): unmangled symbols
extern "C" void generateVertexMap(const uint16_t* depth_pixels, glm::vec3* vertex_map, const int width, const int height,
                       const glm::vec2 focal_length, const int2 img_size);
extern "C" void computePointCloudBoundingBox(glm::vec3* points, const int num_points, BoundingBox& bbox);
extern "C" void transformVertexMap(glm::vec3* vertex_map, const glm::mat4& trans, const int size);
// camera tracking (image_kernels.h:27-54)
extern "C" void generateNormalMap(const glm::vec3* vertex_map, glm::vec3* normal_map, const int width, const int height);
extern "C" void bilateralFilter(const uint16_t* depth_in, uint16_t* filtered_out, const int width, const int height);
extern "C" void colorToIntensity(const Color256* color_in, float* intensity_out, const int size);
extern "C" void transformNormalMap(glm::vec3* normal_map, const glm::mat4& trans, const int size);
// in place like the reference: the first (width/2)*(height/2) elements of `data` receive the result.
// Instantiated for the types the reference instantiates that tracking reads: subsampleDepth<uint16_t>, subsample<float>.
template <class T> void subsample(T* data, const int width, const int height);
template <class T> void subsampleDepth(T* data, const int width, const int height);
}  // namespace sensor
}  // namespace octree_slam
#endif
