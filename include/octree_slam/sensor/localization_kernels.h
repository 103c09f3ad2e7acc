// Drop-in for include/octree_slam/sensor/localization_kernels.h:17-46.
#ifndef OSL_B200_LOCALIZATION_KERNELS_H_
#define OSL_B200_LOCALIZATION_KERNELS_H_
#include <octree_slam/common_types.h>

namespace octree_slam {
namespace sensor {

struct ICPFrame {  // localization_kernels.h:17-24
  ICPFrame(const int w, const int h);
  ~ICPFrame();
  glm::vec3* vertex;
  glm::vec3* normal;
  int width;
  int height;
};

struct RGBDFrame {  // localization_kernels.h:26-33
  RGBDFrame(const int w, const int h);
  ~RGBDFrame();
  float* intensity;
  glm::vec3* vertex;
  int width;
  int height;
};

// localization_kernels.h:40: normal equations of the point-to-plane ICP, every pixel paired with the same pixel of
// the last frame.  A: 36 floats (row-major 6x6), b: 6 floats, in host memory.
extern "C" void computeICPCost2(const ICPFrame* last_frame, const ICPFrame& this_frame, float* A, float* b);
// localization_kernels.h:37: in the reference this variant compacts the correspondences first and sums the same
// terms without the depth-range test of computeICPCost2; here it forwards to computeICPCost2.
extern "C" void computeICPCost(const ICPFrame* last_frame, const ICPFrame& this_frame, float* A, float* b);
// localization_kernels.h:43: empty in the reference (localization_kernels.cu:332-335); empty here.
extern "C" void computeRGBDCost(const RGBDFrame* last_frame, const RGBDFrame& this_frame, float* A, float* b);

}  // namespace sensor
}  // namespace octree_slam
#endif
