// glm stand-in (see ../glm.hpp): translate
#pragma once
#include "../glm.hpp"
namespace glm {
inline mat4 translate(const vec3& t) { mat4 m; m[3] = vec4(t, 1.0f); return m; }
}
