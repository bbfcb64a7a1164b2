// Uses the real glm when the including project has it (the reference vendors glm 0.9.9.7 and its
// main.cpp includes <glm/glm.hpp> before these headers); otherwise a minimal vec3 so that the headless
// driver and the tests build without any third-party header.
#pragma once
#if defined(GLM_VERSION) || defined(GLM_SETUP_INCLUDED) || defined(SPHE_USE_GLM)
#include <glm/glm.hpp>
#elif defined(__has_include)
#if __has_include(<glm/glm.hpp>)
#include <glm/glm.hpp>
#else
#define SPHE_MINI_GLM 1
#endif
#else
#define SPHE_MINI_GLM 1
#endif

#ifdef SPHE_MINI_GLM
namespace glm {
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
}  // namespace glm
#endif
