// Small vector maths + the exactly-rounded primitives used wherever depth/coverage must be bit-identical
// to the reference binary (GCC -O3 -mavx2 -mfma, gnu++11 => -ffp-contract=fast; SURVEY.md Appendix B and
// DESIGN.md "Arithmetic contract").  Parity-critical code spells every rounding step with x*() helpers so
// that neither nvcc (-fmad) nor a host compiler can re-associate or fuse differently.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SGL_HD __host__ __device__ __forceinline__
#define SGL_HDN static __host__ __device__ __noinline__
#ifdef SGL_SHADE_INLINE
#define SGL_HDN_T __host__ __device__ __forceinline__
#else
#define SGL_HDN_T __host__ __device__ __noinline__
#endif
#else
#define SGL_HD inline
#define SGL_HDN static
#define SGL_HDN_T
#endif

#if defined(__CUDA_ARCH__)
SGL_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
SGL_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
SGL_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
SGL_HD float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SGL_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
// host build (development emulator only): compiled with -ffp-contract=off
SGL_HD float xmul(float a, float b) { volatile float r = a * b; return r; }
SGL_HD float xadd(float a, float b) { volatile float r = a + b; return r; }
SGL_HD float xsub(float a, float b) { volatile float r = a - b; return r; }
SGL_HD float xfma(float a, float b, float c) { return fmaf(a, b, c); }
SGL_HD float xdiv(float a, float b) { volatile float r = a / b; return r; }
#endif

// Shading-only arithmetic (lighting maths of the fragment shaders): colour has a 1/255 tolerance, so the device uses
// the SFU forms (MUFU.RCP / RSQ / LG2 / EX2); -DSGL_PRECISE_SHADING restores IEEE division, sqrt and libm pow.
// Coverage, depth, interpolation and texel arithmetic never go through these.
#if defined(__CUDA_ARCH__) && !defined(SGL_PRECISE_SHADING)
SGL_HD float sdiv(float a, float b) { return __fdividef(a, b); }
SGL_HD float srsqrt(float a) { return rsqrtf(a); }
SGL_HD float spow(float a, float e) { return __powf(a, e); }
#else
SGL_HD float sdiv(float a, float b) { return a / b; }
SGL_HD float srsqrt(float a) { return 1.0f / sqrtf(a); }
SGL_HD float spow(float a, float e) { return powf(a, e); }
#endif

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

SGL_HD V2 v2(float x, float y) { V2 r = {x, y}; return r; }
SGL_HD V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
SGL_HD V3 v3s(float s) { V3 r = {s, s, s}; return r; }
SGL_HD V4 v4(float x, float y, float z, float w) { V4 r = {x, y, z, w}; return r; }
SGL_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SGL_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SGL_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
SGL_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
SGL_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
SGL_HD V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
SGL_HD V3 operator/(V3 a, float s) { return v3(sdiv(a.x, s), sdiv(a.y, s), sdiv(a.z, s)); }
SGL_HD V3 operator/(V3 a, V3 b) { return v3(sdiv(a.x, b.x), sdiv(a.y, b.y), sdiv(a.z, b.z)); }
SGL_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SGL_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
SGL_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
SGL_HD V3 normalize(V3 a) { float inv = srsqrt(dot(a, a)); return a * inv; }
SGL_HD V3 normalizeP(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }   // IEEE form (IBL generation passes)   // glm: v * inversesqrt(dot(v,v))
SGL_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
SGL_HD V3 vmax(V3 a, V3 b) { return v3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
SGL_HD V3 vpow(V3 a, float e) { return v3(spow(a.x, e), spow(a.y, e), spow(a.z, e)); }
SGL_HD V3 vmix(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
SGL_HD V3 reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }

// glm::max / glm::min / glm::clamp keep the reference's NaN behaviour: max(x,y) = (x < y) ? y : x
SGL_HD float gmax(float x, float y) { return (x < y) ? y : x; }
SGL_HD float gmin(float x, float y) { return (y < x) ? y : x; }
SGL_HD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

// column-major 4x4 (GLM storage): m[col*4 + row]
// M * vec4(p, 1) exactly as the oracle binary evaluates it in every vertex shader
// (objdump of VS::shaderMain: vmulps m1,y ; vfmadd m0,x ; vfmadd213 m2,z,+m3 ; vaddps):
//     r = fma(z, m2, m3) + fma(x, m0, rn(y * m1))
SGL_HD V4 xMat4MulPoint(const float *m, float x, float y, float z) {
  V4 r;
  r.x = xadd(xfma(z, m[8], m[12]), xfma(x, m[0], xmul(y, m[4])));
  r.y = xadd(xfma(z, m[9], m[13]), xfma(x, m[1], xmul(y, m[5])));
  r.z = xadd(xfma(z, m[10], m[14]), xfma(x, m[2], xmul(y, m[6])));
  r.w = xadd(xfma(z, m[11], m[15]), xfma(x, m[3], xmul(y, m[7])));
  return r;
}

// _mm_dp_ps(a, b, 0xff) with b.w = 0 term:  (a0*b0 + a1*b1) + (a2*b2 + a3*b3)   each step rounded to fp32
SGL_HD float xDot4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
  return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xadd(xmul(a2, b2), xmul(a3, b3)));
}

// glm::mix(x, y, t) as compiled in RendererSoft::interpolateLinear: fma(x, 1-t, rn(y*t))
SGL_HD float xMix(float x, float y, float t, float oneMinusT) { return xfma(x, oneMinusT, xmul(y, t)); }
