// Minimal stand-in for the subset of glm 0.9.5 the hot-path API uses (vec2/vec3/vec4/mat3/mat4 with the same memory
// layout: column-major mat4, 12-byte vec3).  Written from scratch for this repository; if the real glm is on the
// include path first, it is used instead and everything below still compiles (only .x/.y/.z/[] are touched).
// CALLING CONVENTION: glm 0.9.5.4's vector and matrix types have user-provided copy constructors, so the Itanium C++
// ABI passes them BY INVISIBLE REFERENCE where the reference's seam takes them by value (svoFromPointCloud's
// octree_center, coneTraceSVO's resolution / cameraPose / SVO, generateVertexMap's focal_length).  The types below
// declare their copy operations too, so libosl_host.so built against this header is call-compatible with objects
// built against the reference's glm (static_asserts in octree-slam_b200/host/osl_host.cpp; link-level test in
// tests/test_host_shim.py).
#ifndef OSL_MINI_GLM_HPP_
#define OSL_MINI_GLM_HPP_
namespace glm {
struct vec2 {
  float x, y;
  vec2() : x(0), y(0) {}
  vec2(float a, float b) : x(a), y(b) {}
  vec2(const vec2& o) : x(o.x), y(o.y) {}
  vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
};
struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
  vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
struct vec4 {
  float x, y, z, w;
  vec4() : x(0), y(0), z(0), w(0) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
  vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};
struct mat4 {
  vec4 c[4];  // columns
  mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
  mat4(const mat4& o) { for (int i = 0; i < 4; i++) c[i] = o.c[i]; }
  mat4& operator=(const mat4& o) { for (int i = 0; i < 4; i++) c[i] = o.c[i]; return *this; }
  explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
struct mat3 {
  vec3 c[3];  // columns
  mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
  mat3(const mat3& o) { for (int i = 0; i < 3; i++) c[i] = o.c[i]; }
  mat3& operator=(const mat3& o) { for (int i = 0; i < 3; i++) c[i] = o.c[i]; return *this; }
  explicit mat3(const mat4& m) {
    for (int i = 0; i < 3; i++) c[i] = vec3(m.c[i].x, m.c[i].y, m.c[i].z);
  }
  vec3& operator[](int i) { return c[i]; }
  const vec3& operator[](int i) const { return c[i]; }
};
inline const float* value_ptr(const mat4& m) { return &m.c[0].x; }
}  // namespace glm
#endif
