// Layout contract of the hot path: the reference's POD types (include/octree_slam/common_types.h:8-79 in the
// reference tree), same names, same memory layout, without the CUDA/GL dependencies.  Device pointers stay raw.
#ifndef OSL_B200_COMMON_TYPES_H_
#define OSL_B200_COMMON_TYPES_H_
#include <stdint.h>

#include "glm/glm.hpp"

#if defined(__has_include) && __has_include(<vector_types.h>)
#include <vector_types.h>      // CUDA's uchar4 / int2 when the toolkit headers are on the include path
#include <vector_functions.h>  // make_int2
#elif !defined(__VECTOR_TYPES_H__)
struct uchar4 { unsigned char x, y, z, w; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
#endif

struct BoundingBox {  // common_types.h:8-17; the two members are defined out of line (common_types.cu:8-34 in the
  glm::vec3 bbox0 = glm::vec3(0.0f);  // reference, osl_host.cpp here) so that objects compiled against the
  glm::vec3 bbox1 = glm::vec3(0.0f);  // reference's header find them in libosl_host.so
  bool contains(const BoundingBox& other) const;
  float distanceOutside(const BoundingBox& other) const;
};

struct Mesh {  // common_types.h:20-32 ("a lighter weight version of obj"): HOST arrays
  int vbosize = 0, nbosize = 0, cbosize = 0, ibosize = 0, tbosize = 0;
  float* vbo = nullptr;  // 3 floats per vertex
  float* nbo = nullptr;
  float* cbo = nullptr;
  int* ibo = nullptr;    // 3 indices per triangle
  float* tbo = nullptr;  // 6 floats per triangle (u, v of its three corners), voxelization.cu:113-118
  BoundingBox bbox;
};

struct bmp_texture {  // common_types.h:34-38
  glm::vec3* data = nullptr;
  int width = 0;
  int height = 0;
};

struct Camera {  // common_types.h:40-47 (only view and fov are raycast inputs)
  glm::mat4 model, view, projection, modelview, mvp;
  float fov = 45.0f;
};

struct Color256 { uint8_t r, g, b; };  // common_types.h:49-53, 3 bytes

struct VoxelGrid {  // common_types.h:55-63; centers/colors are DEVICE arrays of glm::vec4 owned by the grid
  ~VoxelGrid();
  glm::vec4* centers = nullptr;
  glm::vec4* colors = nullptr;
  int size = 0;
  float scale = 0.0f;
  BoundingBox bbox;
};

struct RawFrame {  // common_types.h:65-73; color/depth are DEVICE arrays
  RawFrame(const int w, const int h);
  ~RawFrame();
  Color256* color;
  uint16_t* depth;
  int height;
  int width;
  long long timestamp;
};

struct SVO {  // common_types.h:75-79: aliases the node pool, no ownership; size = HALF edge length
  unsigned int* data;
  glm::vec3 center;
  float size;
};
#endif
