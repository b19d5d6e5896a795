// glm stand-in (see ../glm.hpp): the quaternion pieces Camera.h names
#pragma once
#include "../glm.hpp"
namespace glm {
struct quat {
  float w = 1, x = 0, y = 0, z = 0;
  quat() = default;
  quat(float a, float b, float c, float d) : w(a), x(b), y(c), z(d) {}
};
inline quat operator*(const quat& p, const quat& q) {
  return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
              p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
inline quat normalize(const quat& q) { float s = 1.0f / std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); return quat(q.w * s, q.x * s, q.y * s, q.z * s); }
inline quat angleAxis(float a, const vec3& v) { float s = std::sin(a * 0.5f); return quat(std::cos(a * 0.5f), v.x * s, v.y * s, v.z * s); }
inline mat4 mat4_cast(const quat& q) {
  mat4 m;
  m[0] = vec4(1 - 2 * (q.y * q.y + q.z * q.z), 2 * (q.x * q.y + q.w * q.z), 2 * (q.x * q.z - q.w * q.y), 0);
  m[1] = vec4(2 * (q.x * q.y - q.w * q.z), 1 - 2 * (q.x * q.x + q.z * q.z), 2 * (q.y * q.z + q.w * q.x), 0);
  m[2] = vec4(2 * (q.x * q.z + q.w * q.y), 2 * (q.y * q.z - q.w * q.x), 1 - 2 * (q.x * q.x + q.y * q.y), 0);
  return m;
}
inline quat quat_cast(const mat3& m) {  // trace form; enough for the rotation matrices Camera.h builds
  float t = m.c[0].x + m.c[1].y + m.c[2].z;
  if (t > 0) { float s = std::sqrt(t + 1.0f) * 2; return quat(0.25f * s, (m.c[1].z - m.c[2].y) / s, (m.c[2].x - m.c[0].z) / s, (m.c[0].y - m.c[1].x) / s); }
  return quat(0, 1, 0, 0);
}
}  // namespace glm
