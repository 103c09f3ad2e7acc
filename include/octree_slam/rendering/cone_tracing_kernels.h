// Drop-in for include/octree_slam/rendering/cone_tracing_kernels.h:16.
#ifndef OSL_B200_CONE_TRACING_KERNELS_H_
#define OSL_B200_CONE_TRACING_KERNELS_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace rendering {
// pos: DEVICE buffer of resolution.x * resolution.y uchar4 (a mapped GL PBO in the reference)
// (`extern "C"` as in the reference: the symbol is the unmangled `coneTraceSVO`)
extern "C" void coneTraceSVO(uchar4* pos, glm::vec2 resolution, float fov, glm::mat4 cameraPose, SVO octree);
}  // namespace rendering
}  // namespace octree_slam
#endif
