// Texturing: image layouts, wrap modes, nearest / bilinear / mip filtering and cube-face selection,
// reading Linear, Tiled(4x4) and Morton(32x32) storage directly.  Restates, for the device:
//   Buffer/TiledBuffer/MortonBuffer::convertIndex      src/Base/Buffer.h:31-33,151-158,185-202
//   BaseSampler::textureImpl                           src/Render/Software/SamplerSoft.h:118-168
//   BaseSampler::pixelWithWrapMode (incl. the CoordMod macro precedence quirk, :14) :171-213
//   sampleNearest / sampleBilinear / samplePixelBilinear                             :216-266
//   BaseSamplerCube::convertXYZ2UV                                                    :312-373
// Texel arithmetic is 8-bit truncating exactly like glm::mix on u8vec4 in the oracle binary:
//   mix(x, y, a) = u8( fma(float(y), a, rn(float(x) * (1 - a))) )     (objdump of samplePixelBilinear)
#pragma once
#include "sgl_math.h"
#include "sgl_types.h"

#if defined(__CUDA_ARCH__)
#define SGL_LDG(p) __ldg(p)
#else
#define SGL_LDG(p) (*(p))
#endif

// Instrumentation build (-DSGL_TOUCH_BITMAP, tools/gpu/texel_touch.py): every texel load of a sampler marks its 32-byte
// DRAM sector in a per-texture bitmap -- the "touched-tile bitmap" SURVEY 8d asks for to measure the unique texel bytes a
// frame needs (B_tex of the roofline).  Compiled out of the product library.
#if defined(SGL_TOUCH_BITMAP) && defined(__CUDA_ARCH__)
__device__ __forceinline__ void sglTouch(uint32_t *touch, const uint8_t *base, const void *addr) {
  if (!touch) return;
  const size_t sector = (size_t) ((const uint8_t *) addr - base) >> 5;
  atomicOr(&touch[sector >> 5], 1u << (sector & 31));
}
#define SGL_TOUCH_FIELDS uint32_t *touch; const uint8_t *tbase;
#define SGL_TOUCH_SET(view, t) (view).touch = (t)->touch; (view).tbase = (t)->base;
#define SGL_TOUCH_NONE(view) (view).touch = nullptr; (view).tbase = nullptr;
#define SGL_TOUCH(view, addr) sglTouch((view).touch, (view).tbase, (addr))
#else
#define SGL_TOUCH_FIELDS
#define SGL_TOUCH_SET(view, t)
#define SGL_TOUCH_NONE(view)
#define SGL_TOUCH(view, addr)
#endif

SGL_HD uint32_t sglMorton2(uint32_t x, uint32_t y) {   // MortonBuffer::encode16_morton2, x in even bits
  uint32_t res = x | (y << 16);
  res = (res | (res << 4)) & 0x0f0f0f0fu;
  res = (res | (res << 2)) & 0x33333333u;
  res = (res | (res << 1)) & 0x55555555u;
  return (res | (res >> 15)) & 0xffffu;
}

// texel index inside one level (levels are limited to 2^31 texels, checked at creation)
SGL_HD uint32_t sglTexelIndex(int layout, int w, int x, int y) {
  if (layout == SGL_LAYOUT_LINEAR) return (uint32_t) x + (uint32_t) y * (uint32_t) w;
  if (layout == SGL_LAYOUT_TILED) {
    uint32_t tw = (uint32_t) (w + 3) >> 2;
    return ((((uint32_t) y >> 2) * tw + ((uint32_t) x >> 2)) << 4) + (((uint32_t) y & 3u) << 2) + ((uint32_t) x & 3u);
  }
  uint32_t tw = (uint32_t) (w + 31) >> 5;
  return ((((uint32_t) y >> 5) * tw + ((uint32_t) x >> 5)) << 10) + sglMorton2((uint32_t) x & 31u, (uint32_t) y & 31u);
}

SGL_HD size_t sglLevelTexels(int layout, int w, int h) {   // innerWidth * innerHeight (Buffer.h:43-48,143-148,174-179)
  if (layout == SGL_LAYOUT_LINEAR) return (size_t) w * h;
  int ts = layout == SGL_LAYOUT_TILED ? 4 : 32;
  return (size_t) ((w + ts - 1) / ts * ts) * (size_t) ((h + ts - 1) / ts * ts);
}

SGL_HD int sglLevelDim(int d, int level) { int v = d >> level; return v > 1 ? v : 1; }

struct SglSampler {        // BaseSampler state after Sampler2DSoft/SamplerCubeSoft::setTexture (SamplerSoft.h:388-394,428-436)
  const SglTexObj *tex;
  int filter, wrap;
  uint32_t border;
};

// One axis of pixelWithWrapMode (the reference wraps x with the width and y with the height independently).
// Returns 0 = in range, 1 = border colour applies, 2 = Buffer::get bounds check fails -> T(0).
SGL_HD int sglWrapAxis(int &x, int n, int wrap) {
  switch (wrap) {
    case SGL_WRAP_REPEAT:
      // #define CoordMod(i, n) ((i) & ((n) - 1) + (n)) & ((n) - 1)  ==  i & (2n-1) & (n-1): always inside [0, n-1]
      x = (x & ((n - 1) + n)) & (n - 1);
      return 0;
    case SGL_WRAP_MIRRORED_REPEAT:
      x = (x & ((2 * n - 1) + 2 * n)) & (2 * n - 1);
      x -= n;
      x = x >= 0 ? x : (-1 - x);
      x = n - 1 - x;
      return (unsigned) x >= (unsigned) n ? 2 : 0;
    case SGL_WRAP_CLAMP_TO_EDGE:
      x = x < 0 ? 0 : (x >= n ? n - 1 : x);
      return 0;
    case SGL_WRAP_CLAMP_TO_BORDER:
      return (x < 0 || x >= n) ? 1 : 0;
  }
  return (unsigned) x >= (unsigned) n ? 2 : 0;
}

// pixelWithWrapMode on both axes: returns false when the border colour applies
SGL_HD bool sglWrapCoord(int &x, int &y, int w, int h, int wrap) {
  int rx = sglWrapAxis(x, w, wrap), ry = sglWrapAxis(y, h, wrap);
  return !(rx == 1 || ry == 1);
}

// one level of one layer, resolved once per filter footprint
struct SglLevelView {
  const uint32_t *ptr;
  int w, h, layout;
  SGL_TOUCH_FIELDS
};
SGL_HD SglLevelView sglLevelView(const SglTexObj *t, int layer, int level) {
  SglLevelView v;
  v.w = sglLevelDim(t->width, level);
  v.h = sglLevelDim(t->height, level);
  v.layout = t->layout;
  v.ptr = (const uint32_t *) (t->base + (size_t) layer * t->layerStride + t->levelOffset[level]);
  SGL_TOUCH_SET(v, t)
  return v;
}
// texel at wrapped coordinates; rx/ry are the sglWrapAxis results of the two coordinates
SGL_HD uint32_t sglFetchWrapped(const SglLevelView &lv, uint32_t border, int x, int rx, int y, int ry) {
  if (rx == 1 || ry == 1) return border;
  if ((rx | ry) != 0) return 0u;                       // Buffer::get bounds check -> T(0)
  SGL_TOUCH(lv, lv.ptr + sglTexelIndex(lv.layout, lv.w, x, y));
  return SGL_LDG(lv.ptr + sglTexelIndex(lv.layout, lv.w, x, y));
}

// raw 32-bit texel (RGBA8 packed little-endian or float bits) of layer/level at integer coords with wrap
SGL_HD uint32_t sglTexel(const SglSampler &s, int layer, int level, int x, int y) {
  SglLevelView lv = sglLevelView(s.tex, layer, level);
  int rx = sglWrapAxis(x, lv.w, s.wrap), ry = sglWrapAxis(y, lv.h, s.wrap);
  return sglFetchWrapped(lv, s.border, x, rx, y, ry);
}

SGL_HD uint32_t sglMixU8(uint32_t a, uint32_t b, float f, float omf) {
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float x = (float) ((a >> (8 * c)) & 0xffu), y = (float) ((b >> (8 * c)) & 0xffu);
    float m = xfma(y, f, xmul(x, omf));
    r |= ((uint32_t) (int) m & 0xffu) << (8 * c);
  }
  return r;
}

SGL_HD float sglMixF32(uint32_t a, uint32_t b, float f, float omf) {
#if defined(__CUDA_ARCH__)
  float x = __uint_as_float(a), y = __uint_as_float(b);
#else
  float x, y;
  memcpy(&x, &a, 4);
  memcpy(&y, &b, 4);
#endif
  return xfma(x, omf, xmul(y, f));   // scalar glm::mix<float> as compiled: vmulss y*a; vfmadd231ss x*(1-a) (unit KATs f32_16)
}

SGL_HD uint32_t sglFloatBits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
SGL_HD float sglBitsFloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

SGL_HD uint32_t sglMixTexel(int format, uint32_t a, uint32_t b, float f) {
  float omf = xsub(1.0f, f);
  if (format == SGL_FMT_RGBA8) return sglMixU8(a, b, f, omf);
  return sglFloatBits(sglMixF32(a, b, f, omf));
}

// samplePixelBilinear: uv in texel units of the level
SGL_HD uint32_t sglPixelBilinearView(const SglLevelView &lv, int format, int wrap, uint32_t border, float u, float v) {
  float tu = xsub(u, 0.5f), tv = xsub(v, 0.5f);
  float fu = floorf(tu), fv = floorf(tv);
  int x0 = (int) fu, y0 = (int) fv;
  int x1 = x0 + 1, y1 = y0 + 1;
  int rx0 = sglWrapAxis(x0, lv.w, wrap), rx1 = sglWrapAxis(x1, lv.w, wrap);
  int ry0 = sglWrapAxis(y0, lv.h, wrap), ry1 = sglWrapAxis(y1, lv.h, wrap);
  uint32_t s1 = sglFetchWrapped(lv, border, x0, rx0, y0, ry0);
  uint32_t s2 = sglFetchWrapped(lv, border, x1, rx1, y0, ry0);
  uint32_t s3 = sglFetchWrapped(lv, border, x0, rx0, y1, ry1);
  uint32_t s4 = sglFetchWrapped(lv, border, x1, rx1, y1, ry1);
  float fx = xsub(tu, fu), fy = xsub(tv, fv);     // glm::fract(x) = x - floor(x)
  return sglMixTexel(format, sglMixTexel(format, s1, s2, fx), sglMixTexel(format, s3, s4, fx), fy);
}
SGL_HD uint32_t sglPixelBilinear(const SglSampler &s, int layer, int level, float u, float v) {
  return sglPixelBilinearView(sglLevelView(s.tex, layer, level), s.tex->format, s.wrap, s.border, u, v);
}

SGL_HD uint32_t sglSampleNearest(const SglSampler &s, int layer, int level, float u, float v, int ox, int oy) {
  SglLevelView lv = sglLevelView(s.tex, layer, level);
  int x = (int) floorf(xmul(u, (float) lv.w)) + ox;
  int y = (int) floorf(xmul(v, (float) lv.h)) + oy;
  int rx = sglWrapAxis(x, lv.w, s.wrap), ry = sglWrapAxis(y, lv.h, s.wrap);
  return sglFetchWrapped(lv, s.border, x, rx, y, ry);
}

SGL_HD uint32_t sglSampleBilinear(const SglSampler &s, int layer, int level, float u, float v, int ox, int oy) {
  SglLevelView lv = sglLevelView(s.tex, layer, level);
  float tu = xadd(xmul(u, (float) lv.w), (float) ox);
  float tv = xadd(xmul(v, (float) lv.h), (float) oy);
  return sglPixelBilinearView(lv, s.tex->format, s.wrap, s.border, tu, tv);
}

// BaseSampler::textureImpl
SGL_HDN uint32_t sglTextureImpl(const SglSampler &s, int layer, float u, float v, float lod, int ox, int oy) {
  const SglTexObj *t = s.tex;
  if (t == nullptr || t->base == nullptr) return 0u;
  int f = s.filter;
  if (f == SGL_FILTER_NEAREST) return sglSampleNearest(s, layer, 0, u, v, ox, oy);
  if (f == SGL_FILTER_LINEAR) return sglSampleBilinear(s, layer, 0, u, v, ox, oy);
  int maxLevel = t->levels - 1;
  if (f == SGL_FILTER_NEAREST_MIPMAP_NEAREST || f == SGL_FILTER_LINEAR_MIPMAP_NEAREST) {
    int level = (int) ceilf(xadd(lod, 0.5f)) - 1;
    level = level < 0 ? 0 : (level > maxLevel ? maxLevel : level);
    if (f == SGL_FILTER_NEAREST_MIPMAP_NEAREST) return sglSampleNearest(s, layer, level, u, v, ox, oy);
    return sglSampleBilinear(s, layer, level, u, v, ox, oy);
  }
  int hi = (int) floorf(lod);
  hi = hi < 0 ? 0 : (hi > maxLevel ? maxLevel : hi);
  int lo = hi + 1;
  lo = lo < 0 ? 0 : (lo > maxLevel ? maxLevel : lo);
  bool nearest = f == SGL_FILTER_NEAREST_MIPMAP_LINEAR;
  uint32_t thi = nearest ? sglSampleNearest(s, layer, hi, u, v, ox, oy) : sglSampleBilinear(s, layer, hi, u, v, ox, oy);
  if (hi == lo) return thi;
  uint32_t tlo = nearest ? sglSampleNearest(s, layer, lo, u, v, ox, oy) : sglSampleBilinear(s, layer, lo, u, v, ox, oy);
  float fr = xsub(lod, floorf(lod));
  return sglMixTexel(t->format, thi, tlo, fr);
}

// BaseSamplerCube::convertXYZ2UV -- an if-chain where later matches override earlier ones on ties
template<bool FAST>
SGL_HD void sglCubeFaceT(float x, float y, float z, int &index, float &u, float &v) {
  float absX = fabsf(x), absY = fabsf(y), absZ = fabsf(z);
  bool xp = x > 0, yp = y > 0, zp = z > 0;
  float maxAxis = 0.f, uc = 0.f, vc = 0.f;
  index = 0;
  if (xp && absX >= absY && absX >= absZ) { maxAxis = absX; uc = -z; vc = y; index = 0; }
  if (!xp && absX >= absY && absX >= absZ) { maxAxis = absX; uc = z; vc = y; index = 1; }
  if (yp && absY >= absX && absY >= absZ) { maxAxis = absY; uc = x; vc = -z; index = 2; }
  if (!yp && absY >= absX && absY >= absZ) { maxAxis = absY; uc = x; vc = z; index = 3; }
  if (zp && absZ >= absX && absZ >= absY) { maxAxis = absZ; uc = x; vc = y; index = 4; }
  if (!zp && absZ >= absX && absZ >= absY) { maxAxis = absZ; uc = -x; vc = y; index = 5; }
  vc = -vc;
  u = 0.5f * ((FAST ? sdiv(uc, maxAxis) : uc / maxAxis) + 1.0f);
  v = 0.5f * ((FAST ? sdiv(vc, maxAxis) : vc / maxAxis) + 1.0f);
}
SGL_HD void sglCubeFace(float x, float y, float z, int &index, float &u, float &v) { sglCubeFaceT<false>(x, y, z, index, u, v); }

// float(c) / 255.f for c in 0..255, bit-exact without an IEEE division: with r = RN(1/255), q0 = RN(c*r),
// RN(q0 + r * (c - 255*q0)) is the correctly rounded quotient (Markstein); checked exhaustively for all 256 inputs
// (tests/test_capi_and_host.py::test_div255_identity)
SGL_HD float sglDiv255(float c) {
  const float r = 1.0f / 255.0f;
  float q0 = xmul(c, r);
  return xfma(xfma(-q0, 255.0f, c), r, q0);
}
SGL_HD V4 sglUnpackRGBA(uint32_t p) {   // vec4(u8vec4) / 255.f   (ShaderSoft.h:80-111)
  return v4(sglDiv255((float) (p & 0xffu)), sglDiv255((float) ((p >> 8) & 0xffu)), sglDiv255((float) ((p >> 16) & 0xffu)),
            sglDiv255((float) (p >> 24)));
}

SGL_HD V4 sglTexture2D(const SglSampler &s, V2 uv, float lod) {
  return sglUnpackRGBA(sglTextureImpl(s, 0, uv.x, uv.y, lod, 0, 0));
}
SGL_HD V4 sglTexture2DOffset(const SglSampler &s, V2 uv, float lod, int ox, int oy) {
  return sglUnpackRGBA(sglTextureImpl(s, 0, uv.x, uv.y, lod, ox, oy));
}
SGL_HD float sglTexture2DFloat(const SglSampler &s, V2 uv) {
  return sglBitsFloat(sglTextureImpl(s, 0, uv.x, uv.y, 0.f, 0, 0));
}
template<bool FAST>
SGL_HD V4 sglTextureCubeT(const SglSampler &s, V3 dir, float lod) {
  int face;
  float u, v;
  sglCubeFaceT<FAST>(dir.x, dir.y, dir.z, face, u, v);
  if (s.tex == nullptr || face >= s.tex->layers) return v4(0, 0, 0, 0);
  return sglUnpackRGBA(sglTextureImpl(s, face, u, v, lod, 0, 0));
}
// per-frame shaders (PBR, skybox) take the SFU division for the face coordinates; IBL generation keeps IEEE arithmetic
SGL_HD V4 sglTextureCube(const SglSampler &s, V3 dir, float lod) { return sglTextureCubeT<true>(s, dir, lod); }
SGL_HD V4 sglTextureCubeP(const SglSampler &s, V3 dir, float lod) { return sglTextureCubeT<false>(s, dir, lod); }

// ---- split-phase bilinear taps for the straight-line shader fast paths ----------------------------------------
// A "simple" sampler = RGBA8, one sample, LINEAR layout, wrap REPEAT or CLAMP_TO_EDGE (checked per draw on the host,
// SglDrawRec::fastSamplers).  sglTapIssue computes the footprint exactly like sampleBilinear/samplePixelBilinear and
// issues the four texel loads without consuming them, so that a shader can put the loads of all its maps in flight
// before the first truncating mix; sglTapMix then performs the reference's three u8 mixes.  Same arithmetic, same bits.
struct SglTapView {
  const uint32_t *ptr;   // level base
  int w, h;
  bool clamp;            // CLAMP_TO_EDGE, else REPEAT
  SGL_TOUCH_FIELDS
};
struct SglTap {
  uint32_t s1, s2, s3, s4;
  float fx, fy;
};
SGL_HD SglTapView sglTapView(const SglTexObj *t, int layer, int level, int wrap) {
  SglTapView v;
  v.w = sglLevelDim(t->width, level);
  v.h = sglLevelDim(t->height, level);
  v.ptr = (const uint32_t *) (t->base + (size_t) layer * t->layerStride + t->levelOffset[level]);
  v.clamp = wrap == SGL_WRAP_CLAMP_TO_EDGE;
  SGL_TOUCH_SET(v, t)
  return v;
}
SGL_HD SglTapView sglTapViewDummy(const uint32_t *dummy) {
  SglTapView v;
  v.ptr = dummy; v.w = 1; v.h = 1; v.clamp = true;
  SGL_TOUCH_NONE(v)
  return v;
}
SGL_HD int sglTapWrap(int x, int n, bool clamp) {
  int r = (x & ((n - 1) + n)) & (n - 1);                 // CoordMod (SamplerSoft.h:14)
  int c = x < 0 ? 0 : (x >= n ? n - 1 : x);
  return clamp ? c : r;
}
SGL_HD SglTap sglTapIssue(const SglTapView &t, float u, float v, int ox = 0, int oy = 0) {   // uv normalised, texel offset
  float tu = xsub(xadd(xmul(u, (float) t.w), (float) ox), 0.5f), tv = xsub(xadd(xmul(v, (float) t.h), (float) oy), 0.5f);
  float fu = floorf(tu), fv = floorf(tv);
  int x0 = (int) fu, y0 = (int) fv;
  int xa = sglTapWrap(x0, t.w, t.clamp), xb = sglTapWrap(x0 + 1, t.w, t.clamp);
  int ya = sglTapWrap(y0, t.h, t.clamp), yb = sglTapWrap(y0 + 1, t.h, t.clamp);
  SglTap r;
  SGL_TOUCH(t, t.ptr + (uint32_t) ya * (uint32_t) t.w + (uint32_t) xa);
  SGL_TOUCH(t, t.ptr + (uint32_t) ya * (uint32_t) t.w + (uint32_t) xb);
  SGL_TOUCH(t, t.ptr + (uint32_t) yb * (uint32_t) t.w + (uint32_t) xa);
  SGL_TOUCH(t, t.ptr + (uint32_t) yb * (uint32_t) t.w + (uint32_t) xb);
  r.s1 = SGL_LDG(t.ptr + (uint32_t) ya * (uint32_t) t.w + (uint32_t) xa);
  r.s2 = SGL_LDG(t.ptr + (uint32_t) ya * (uint32_t) t.w + (uint32_t) xb);
  r.s3 = SGL_LDG(t.ptr + (uint32_t) yb * (uint32_t) t.w + (uint32_t) xa);
  r.s4 = SGL_LDG(t.ptr + (uint32_t) yb * (uint32_t) t.w + (uint32_t) xb);
  r.fx = xsub(tu, fu);
  r.fy = xsub(tv, fv);
  return r;
}
SGL_HD uint32_t sglTapMix(const SglTap &a) {
  return sglMixTexel(SGL_FMT_RGBA8, sglMixTexel(SGL_FMT_RGBA8, a.s1, a.s2, a.fx), sglMixTexel(SGL_FMT_RGBA8, a.s3, a.s4, a.fx), a.fy);
}
