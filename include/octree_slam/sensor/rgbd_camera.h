// Drop-in for include/octree_slam/sensor/rgbd_camera.h:17-82.  update() runs on the GPU without host round trips
// (osl_tracker_update); position() / orientation() wait for it.
#ifndef OSL_B200_RGBD_CAMERA_H_
#define OSL_B200_RGBD_CAMERA_H_
#include <octree_slam/common_types.h>

struct osl_tracker;

namespace octree_slam {
namespace sensor {

class RGBDCamera {
 public:
  // `exact_jacobian` (not in the reference): false reproduces rgbd_camera.cpp bit for bit in its quirks (Q17, Q18),
  // true is the corrected tracker whose pose() is the camera-to-world transform.
  RGBDCamera(const int width, const int height, const glm::vec2& focal_length, const bool exact_jacobian = false);
  ~RGBDCamera();
  const Camera camera() const;                  // rgbd_camera.cpp:40-51
  const glm::vec3 position() const;
  const glm::mat3 orientation() const;
  void update(const RawFrame* this_frame);      // rgbd_camera.cpp:53-191
  // mat4(orientation()) * translate(mat4(1), position()): the matrix main.cpp:40 hands to transformVertexMap
  const glm::mat4 pose() const;
  bool lost() const;
  const float* poseDevice() const;  // pose() in device memory (column-major), rewritten by every update

 private:
  RGBDCamera(const RGBDCamera&);
  RGBDCamera& operator=(const RGBDCamera&);
  osl_tracker* tracker_;
  glm::vec2 focal_length_;
  int width_, height_;
  long long latest_stamp_;
};

}  // namespace sensor
}  // namespace octree_slam
#endif
