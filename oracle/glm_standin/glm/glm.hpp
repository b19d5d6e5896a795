// glm stand-in for oracle/ref_binding_check.cpp -- TEST INFRASTRUCTURE ONLY, not glm and not from the reference:
// the handful of types and functions the reference's Camera.h names (Camera.h:17-121), so that its
// RenderKernelLauncher.h -> Config.h -> Scene.h -> Camera.h include chain compiles in an image without glm.
#pragma once
#include <cmath>
namespace glm {
struct ivec2 { int x = 0, y = 0; ivec2() = default; ivec2(int a, int b) : x(a), y(b) {} };
struct vec2 { float x = 0, y = 0; vec2() = default; vec2(float a, float b) : x(a), y(b) {} };
struct vec4;
struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() = default;
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  explicit vec3(const vec4& v);
  vec3& operator+=(const vec3& o) { x += o.x, y += o.y, z += o.z; return *this; }
};
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
struct vec4 {
  float x = 0, y = 0, z = 0, w = 0;
  vec4() = default;
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec4& operator*=(float s) { x *= s, y *= s, z *= s, w *= s; return *this; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
struct mat4 {
  vec4 c[4] = {vec4(1, 0, 0, 0), vec4(0, 1, 0, 0), vec4(0, 0, 1, 0), vec4(0, 0, 0, 1)};
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
  mat4& operator*=(const mat4& o) {
    mat4 r;
    for (int j = 0; j < 4; ++j) {
      const float* b = &o.c[j].x;
      r.c[j] = vec4(c[0].x * b[0] + c[1].x * b[1] + c[2].x * b[2] + c[3].x * b[3], c[0].y * b[0] + c[1].y * b[1] + c[2].y * b[2] + c[3].y * b[3],
                    c[0].z * b[0] + c[1].z * b[1] + c[2].z * b[2] + c[3].z * b[3], c[0].w * b[0] + c[1].w * b[1] + c[2].w * b[2] + c[3].w * b[3]);
    }
    return *this = r;
  }
};
inline bool operator!=(const mat4& a, const mat4& b) {
  for (int i = 0; i < 4; ++i)
    if (a[i].x != b[i].x || a[i].y != b[i].y || a[i].z != b[i].z || a[i].w != b[i].w) return true;
  return false;
}
struct mat3 {
  vec3 c[3];
  mat3() = default;
  explicit mat3(const mat4& m) { for (int i = 0; i < 3; ++i) c[i] = vec3(m[i]); }
};
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline vec3 normalize(const vec3& a) { float s = 1.0f / std::sqrt(dot(a, a)); return vec3(a.x * s, a.y * s, a.z * s); }
inline const float* value_ptr(const mat4& m) { return &m.c[0].x; }
inline float* value_ptr(mat4& m) { return &m.c[0].x; }
}  // namespace glm
