// Minimal stand-in for the subset of GLM that crosses the Renderer boundary (Base/GLMInc.h in the reference pulls
// in the real GLM).  Only plain aggregates are needed on this side of the API: ClearStates::clearColor and RGBA.
// When RendererCUDA is built inside the reference tree this header is not used -- the reference's own is.
#pragma once
#include <cstdint>

namespace glm {

struct vec2 {
  float x = 0.f, y = 0.f;
  vec2() = default;
  vec2(float x_, float y_) : x(x_), y(y_) {}
};

struct vec3 {
  float x = 0.f, y = 0.f, z = 0.f;
  vec3() = default;
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct vec4 {
  union { float x; float r; };
  union { float y; float g; };
  union { float z; float b; };
  union { float w; float a; };
  vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
};

struct u8vec4 {
  union { uint8_t x; uint8_t r; };
  union { uint8_t y; uint8_t g; };
  union { uint8_t z; uint8_t b; };
  union { uint8_t w; uint8_t a; };
  u8vec4() : x(0), y(0), z(0), w(0) {}
  u8vec4(uint8_t x_, uint8_t y_, uint8_t z_, uint8_t w_) : x(x_), y(y_), z(z_), w(w_) {}
};

}  // namespace glm

using RGBA = glm::u8vec4;
