// Fixed-function pieces of the pipeline as device functions: post-VS vertex processing, frustum clipping,
// face culling, triangle/line/point setup, per-pixel coverage with the reference's exact float edge
// evaluation, depth test, blending and colour packing.  Restates src/Render/Software/RendererSoft.cpp,
// RendererInternal.h, BlendSoft.h and DepthSoft.h (line numbers per function).
#pragma once
#include <float.h>
#include "sgl_shaders.h"

// ---- vertex post-processing ------------------------------------------------------------------------------
// countFrustumClipMask (RendererSoft.cpp:994-1003)
SGL_HD int sglClipMask(V4 c) {
  int mask = 0;
  if (c.w < c.x) mask |= 1;
  if (c.w < -c.x) mask |= 2;
  if (c.w < c.y) mask |= 4;
  if (c.w < -c.y) mask |= 8;
  if (c.w < c.z) mask |= 16;
  if (c.w < -c.z) mask |= 32;
  return mask;
}

// perspectiveDivideImpl + viewportTransformImpl (RendererSoft.cpp:981-992, setViewPort :92-113):
//   invW = 1/w ; pos *= invW ; pos.w = invW ; pos = pos * innerP + innerO   (mul and add NOT fused in the binary)
SGL_HD V4 sglToScreen(V4 clip, float vpX, float vpY, float vpW, float vpH) {
  float invW = xdiv(1.0f, clip.w);
  V4 p = v4(xmul(clip.x, invW), xmul(clip.y, invW), xmul(clip.z, invW), invW);
  float px = vpW / 2.f, py = vpH / 2.f;           // innerP.xy ; innerP.z = maxDepth - minDepth = 1 ; innerP.w = 1
  float ox = vpX + vpW / 2.f, oy = vpY + vpH / 2.f;  // innerO.xy ; innerO.z = minDepth = 0 ; innerO.w = 0
  V4 r;
  r.x = xadd(xmul(p.x, px), ox);
  r.y = xadd(xmul(p.y, py), oy);
  r.z = xadd(xmul(p.z, 1.0f), 0.0f);
  r.w = xadd(xmul(p.w, 1.0f), 0.0f);
  return r;
}

// processFaceCulling (RendererSoft.cpp:277-299): sign of cross(v1-v0, v2-v0).z, fmsub-contracted in the binary
SGL_HD bool sglFrontFacing(V4 v0, V4 v1, V4 v2) {
  float ax = xsub(v1.x, v0.x), ay = xsub(v1.y, v0.y);
  float bx = xsub(v2.x, v0.x), by = xsub(v2.y, v0.y);
  float nz = xfma(ax, by, -xmul(ay, bx));
  return xadd(nz, 0.0f) > 0.f;
}

// frustum planes in clipping order (Geometry.h:108-133); distance = dot(plane, clipPos) via _mm_dp_ps
SGL_HD float sglPlaneDist(int plane, V4 c) {
  switch (plane) {
    case 0: return xDot4(-1.f, 0.f, 0.f, 1.f, c.x, c.y, c.z, c.w);
    case 1: return xDot4(1.f, 0.f, 0.f, 1.f, c.x, c.y, c.z, c.w);
    case 2: return xDot4(0.f, -1.f, 0.f, 1.f, c.x, c.y, c.z, c.w);
    case 3: return xDot4(0.f, 1.f, 0.f, 1.f, c.x, c.y, c.z, c.w);
    case 4: return xDot4(0.f, 0.f, -1.f, 1.f, c.x, c.y, c.z, c.w);
    default: return xDot4(0.f, 0.f, 1.f, 1.f, c.x, c.y, c.z, c.w);
  }
}

SGL_HD bool sglSignBit(float f) { return (sglFloatBits(f) >> 31) != 0; }

// ---- pixel domain of a triangle ------------------------------------------------------------------------------
// One axis of rasterizationTriangle's block/quad loops (RendererSoft.cpp:724-763):
//   bmin = max(min - .5, 0) - 1 ; bmax = min(max + .5, dim - 1)
//   blocks bx < (int)((bmax - bmin + 32 - 1) / 32) ; start = (int)(bmin + bx*32) ; quads at start+1, +3, ... while
//   q < start + 32 && q <= bmax ; each quad visits pixels q and q+1.
// Returns false if nothing is visited.  `regular` = the visited set is the contiguous range [first, last] with
// quad origins first + 2k (always true unless the float add in `start` rounds across an integer).
SGL_HD bool sglAxisRange(float vmin, float vmax, float dim, int &first, int &last, bool &regular) {
  float mn = gmax(xsub(vmin, 0.5f), 0.f);
  float mx = gmin(xadd(vmax, 0.5f), xsub(dim, 1.f));
  float bmin = xsub(mn, 1.f);
  int cnt = (int) xdiv(xsub(xadd(xsub(mx, bmin), 32.f), 1.f), 32.f);
  first = 0;
  last = -1;
  regular = true;
  bool any = false;
  int expect = 0;
  for (int b = 0; b < cnt; b++) {
    int start = (int) xadd(bmin, (float) (b * SGL_RASTER_BLOCK));
    int q0 = start + 1;
    if (!((float) q0 <= mx)) continue;                 // first quad of the block already beyond bmax
    int lim = (int) floorf(mx);                        // (float)q <= mx  <=>  q <= floor(mx) for integer q
    int lastQ = q0 + SGL_RASTER_BLOCK - 2;             // q < start + 32
    if (lastQ > lim) lastQ = q0 + ((lim - q0) >> 1) * 2;
    if (any && q0 != expect) regular = false;
    if (!any) first = q0;
    if (q0 < first) first = q0;
    if (lastQ + 1 > last) last = lastQ + 1;
    any = true;
    expect = lastQ + 2;
  }
  return any;
}

// exact membership / quad origin for the irregular case: re-walks the blocks
SGL_HD bool sglAxisVisitedExact(float vmin, float vmax, float dim, int p, int &origin) {
  float mn = gmax(xsub(vmin, 0.5f), 0.f);
  float mx = gmin(xadd(vmax, 0.5f), xsub(dim, 1.f));
  float bmin = xsub(mn, 1.f);
  int cnt = (int) xdiv(xsub(xadd(xsub(mx, bmin), 32.f), 1.f), 32.f);
  bool hit = false;
  for (int b = 0; b < cnt; b++) {
    int start = (int) xadd(bmin, (float) (b * SGL_RASTER_BLOCK));
    for (int q = start + 1; q < start + SGL_RASTER_BLOCK && (float) q <= mx; q += 2) {
      if (p == q || p == q + 1) { origin = q; hit = true; }   // a later block re-visiting the pixel wins (runs later)
    }
  }
  return hit;
}

SGL_HD float min3f(float a, float b, float c) { return gmin(gmin(a, b), c); }   // std::min(std::min(a,b),c)
SGL_HD float max3f(float a, float b, float c) { return gmax(gmax(a, b), c); }

SGL_HD uint32_t sglStateFlags(const SglRenderStates &rs, bool hasDepth) {
  uint32_t f = SGL_PF_VALID;
  if (rs.depth_test && hasDepth) f |= SGL_PF_DEPTH_TEST;     // processDepthTest: !depthTest || !fboDepth_ => pass
  if (rs.depth_mask) f |= SGL_PF_DEPTH_MASK;
  if (rs.blend) f |= SGL_PF_BLEND;
  f |= ((uint32_t) rs.depth_func & 7u) << SGL_PF_DEPTH_FUNC_SHIFT;
  return f;
}

// triangle setup; returns false when the triangle cannot produce coverage
SGL_HD bool sglSetupTriangle(SglPrim &p, V4 v0, V4 v1, V4 v2, float vpW, float vpH, bool front, uint32_t stateFlags,
                             uint32_t draw) {
  // degenerate reject of barycentric(): |u.z| < FLT_EPSILON is position independent (RendererSoft.cpp:1034)
  float ax = xsub(v2.x, v0.x), ay = xsub(v1.x, v0.x), bx = xsub(v2.y, v0.y), by = xsub(v1.y, v0.y);
  float uz = xfma(ax, by, -xmul(ay, bx));
  if (fabsf(uz) < FLT_EPSILON) return false;
  int fx, lx, fy, ly;
  bool rx, ry;
  if (!sglAxisRange(min3f(v0.x, v1.x, v2.x), max3f(v0.x, v1.x, v2.x), vpW, fx, lx, rx)) return false;
  if (!sglAxisRange(min3f(v0.y, v1.y, v2.y), max3f(v0.y, v1.y, v2.y), vpH, fy, ly, ry)) return false;
  p.v[0][0] = v0.x; p.v[0][1] = v0.y; p.v[0][2] = v0.z; p.v[0][3] = v0.w;
  p.v[1][0] = v1.x; p.v[1][1] = v1.y; p.v[1][2] = v1.z; p.v[1][3] = v1.w;
  p.v[2][0] = v2.x; p.v[2][1] = v2.y; p.v[2][2] = v2.z; p.v[2][3] = v2.w;
  p.bx0 = (int16_t) (fx < -32768 ? -32768 : fx);
  p.by0 = (int16_t) (fy < -32768 ? -32768 : fy);
  p.bx1 = (int16_t) (lx > 32767 ? 32767 : lx);
  p.by1 = (int16_t) (ly > 32767 ? 32767 : ly);
  p.flags = stateFlags | SGL_PK_TRIANGLE | (front ? SGL_PF_FRONT : 0u) | ((rx && ry) ? 0u : SGL_PF_IRREGULAR);
  p.draw = draw;
  return true;
}

// ---- coverage ----------------------------------------------------------------------------------------------
// sample positions (PixelContext::GetSampleLocation4X + centre, RendererInternal.h:61-101)
SGL_HD void sglSampleOffset(int ns, int i, float &ox, float &oy) {
  if (ns == 1 || i == 4) { ox = 0.5f; oy = 0.5f; return; }
  switch (i) {
    case 0: ox = 0.375f; oy = 0.875f; break;
    case 1: ox = 0.875f; oy = 0.625f; break;
    case 2: ox = 0.125f; oy = 0.375f; break;
    default: ox = 0.625f; oy = 0.125f; break;
  }
}

struct SglTriEdge {   // per-triangle part of barycentric() + constants of the conservative outside test
  float ax, ay, bx, by, uz, x0, y0;
  float sax, say, sbx, sby;   // a/b premultiplied by sign(uz)
  float auz;                  // |uz|
  float tA, tB;               // |ay| + |ax|, |by| + |bx|
};

SGL_HD SglTriEdge sglTriEdge(const SglPrim &p) {
  SglTriEdge e;
  e.x0 = p.v[0][0];
  e.y0 = p.v[0][1];
  e.ax = xsub(p.v[2][0], e.x0);
  e.ay = xsub(p.v[1][0], e.x0);
  e.bx = xsub(p.v[2][1], e.y0);
  e.by = xsub(p.v[1][1], e.y0);
  e.uz = xfma(e.ax, e.by, -xmul(e.ay, e.bx));
  const float sg = e.uz < 0.f ? -1.f : 1.f;
  e.sax = sg * e.ax; e.say = sg * e.ay; e.sbx = sg * e.bx; e.sby = sg * e.by;
  e.auz = fabsf(e.uz);
  e.tA = fabsf(e.ay) + fabsf(e.ax);
  e.tB = fabsf(e.by) + fabsf(e.bx);
  return e;
}

// Conservative test: true only if barycentric() is CERTAIN to report "outside" for every sample position inside the
// rectangle (cx +- hx, cy +- hy).  The reference decides on the signs of its float numerators
//   ux = fma(ay, bz, -rn(az*by)),  uy = fma(az, bx, -rn(ax*bz))        (b2 = ux/uz, b1 = uy/uz, b0 = 1 - (b1 + b2))
// whose distance from the real-valued affine functions is below 2^-23 * (|ay||bz| + |az||by|) (resp. ax/bx); the test
// evaluates the same affine functions at the centre in plain float (same error bound), bounds their variation over
// the rectangle and keeps a 2^-20 relative guard (8x the combined error), so every true outcome here implies the
// reference's own comparison is negative.  NaNs make every comparison false => "not sure" => exact path.
SGL_HD bool sglTriSurelyOutside(const SglTriEdge &e, float cx, float cy, float hx, float hy) {
  const float K = 9.5367431640625e-7f;   // 2^-20
  float azc = e.x0 - cx, bzc = e.y0 - cy;
  float sux = e.say * bzc - azc * e.sby;              // sign(uz) * ux at the centre
  float suy = azc * e.sbx - e.sax * bzc;              // sign(uz) * uy at the centre
  float aaz = fabsf(azc) + hx + 1.f, abz = fabsf(bzc) + hy + 1.f;
  float err = K * (e.tA * abz + e.tB * aaz);
  float rux = hx * fabsf(e.by) + hy * fabsf(e.ay);    // variation of ux / uy over the rectangle
  float ruy = hx * fabsf(e.bx) + hy * fabsf(e.ax);
  if (sux < -(rux + err)) return true;                // b2 < 0 everywhere
  if (suy < -(ruy + err)) return true;                // b1 < 0 everywhere
  if ((sux + suy) - e.auz > rux + ruy + 2.f * err + K * e.auz) return true;   // b1 + b2 > 1 everywhere
  return false;
}

// Mirror image of sglTriSurelyOutside: true only if barycentric() is CERTAIN to report "inside" (b0, b1, b2 all > 0)
// for every sample position inside the rectangle (cx +- hx, cy +- hy); same error model, same 2^-20 guard.
SGL_HD bool sglTriSurelyInside(const SglTriEdge &e, float cx, float cy, float hx, float hy) {
  const float K = 9.5367431640625e-7f;   // 2^-20
  float azc = e.x0 - cx, bzc = e.y0 - cy;
  float sux = e.say * bzc - azc * e.sby;
  float suy = azc * e.sbx - e.sax * bzc;
  float aaz = fabsf(azc) + hx + 1.f, abz = fabsf(bzc) + hy + 1.f;
  float err = K * (e.tA * abz + e.tB * aaz);
  float rux = hx * fabsf(e.by) + hy * fabsf(e.ay);
  float ruy = hx * fabsf(e.bx) + hy * fabsf(e.ax);
  return sux > rux + err && suy > ruy + err && (sux + suy) - e.auz < -(rux + ruy + 2.f * err + K * e.auz);
}

// All three vertex depths are exactly +0 (a skybox under reversed-Z, SkyboxSoft.h:67-76).  Then the interpolated depth of
// every INSIDE sample is exactly +0 as well: z = (b0*0 + b1*0) + (b2*0 + 0*0) with finite b >= 0; b0 = 1 - (b1 + b2) is
// never -0, so the first pair is +0, the second pair ends in + (+0), and (+0) + (+0) = +0.  For pixels that are surely
// inside, coverage and depth are therefore known without evaluating barycentric() at all.
SGL_HD bool sglTriFlatZeroDepth(const SglPrim &p) {
  return (sglFloatBits(p.v[0][2]) | sglFloatBits(p.v[1][2]) | sglFloatBits(p.v[2][2])) == 0u;
}

// RendererSoft::barycentric (RendererSoft.cpp:1021-1056) in the oracle binary's association:
//   a = (x2-x0, x1-x0, x0-px) ; b = (y2-y0, y1-y0, y0-py)
//   u = fma(a.yzx, b.zxy, -rn(a.zxy * b.yzx)) ; u /= u.z ; bc = (1 - (u.x + u.y), u.y, u.x)
//   inside <=> !(bc.x < 0 || bc.y < 0 || bc.z < 0)
SGL_HD bool sglBarycentric(const SglTriEdge &e, float spx, float spy, float &b0, float &b1, float &b2) {
  float az = xsub(e.x0, spx), bz = xsub(e.y0, spy);
  float ux = xfma(e.ay, bz, -xmul(az, e.by));
  float uy = xfma(az, e.bx, -xmul(e.ax, bz));
  float qx = xdiv(ux, e.uz), qy = xdiv(uy, e.uz);
  b0 = xsub(1.f, xadd(qx, qy));
  b1 = qy;
  b2 = qx;
  return !(b0 < 0 || b1 < 0 || b2 < 0);
}

// z and 1/w at a sample: interpolateBarycentric(&position.z, vertZ, 2, bc) -> glm::dot == _mm_dp_ps 0xff (:794,:1110-1112)
SGL_HD float sglInterpZ(const SglPrim &p, int comp, float b0, float b1, float b2) {
  return xDot4(b0, b1, b2, 0.f, p.v[0][comp], p.v[1][comp], p.v[2][comp], 0.f);
}

// DepthTest (DepthSoft.h:13-25)
SGL_HD bool sglDepthTest(float a, float b, int func) {
  switch (func) {
    case 0: return false;
    case 1: return a < b;
    case 2: return fabsf(a - b) <= FLT_EPSILON;
    case 3: return a <= b;
    case 4: return a > b;
    case 5: return fabsf(a - b) > FLT_EPSILON;
    case 6: return a >= b;
    case 7: return true;
  }
  return a < b;
}

// Coverage of one triangle at one pixel: for each of NS samples, geometric coverage + depth-range clip +
// depth test (earlyZTest/processDepthTest, RendererSoft.cpp:377-395,853-878).  Returns the mask of samples that
// pass; zOut[s] is the (clamped) depth to write; shadeIdx = index of the shading sample (4 = pixel centre).
template<int NS>
SGL_HD uint32_t sglCoverTriangle(const SglPrim &p, const SglTriEdge &e, int px, int py, const float *depth, bool hasDepth,
                                 float *zOut, int &shadeIdx) {
  float fx = (float) px, fy = (float) py;
  // all sample positions (and the centre) lie within +-0.375 of the pixel centre
  if (sglTriSurelyOutside(e, fx + 0.5f, fy + 0.5f, NS > 1 ? 0.375f : 0.f, NS > 1 ? 0.375f : 0.f)) return 0;
  if (NS > 1 && sglTriFlatZeroDepth(p) && sglTriSurelyInside(e, fx + 0.5f, fy + 0.5f, 0.375f, 0.375f)) {
    // every sample and the centre are covered, depth is +0 at each of them (see sglTriFlatZeroDepth)
    shadeIdx = 4;
    const uint32_t fl = p.flags;
    const bool dt = (fl & SGL_PF_DEPTH_TEST) != 0;
    const int fn = (fl >> SGL_PF_DEPTH_FUNC_SHIFT) & 7;
    uint32_t pass = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      if (dt && (!hasDepth || !sglDepthTest(0.f, depth[s], fn))) continue;
      zOut[s] = 0.f;
      pass |= 1u << s;
    }
    return pass;
  }
  uint32_t geo = 0;
  float bc[NS][3];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    float ox, oy;
    sglSampleOffset(NS, s, ox, oy);
    if (sglBarycentric(e, xadd(ox, fx), xadd(oy, fy), bc[s][0], bc[s][1], bc[s][2])) geo |= 1u << s;
  }
  if (geo == 0) return 0;
  shadeIdx = 0;
  if (NS > 1) {
    float c0, c1, c2;
    bool centre = sglBarycentric(e, xadd(fx, 0.5f), xadd(fy, 0.5f), c0, c1, c2);
    if (centre) shadeIdx = 4;
    else {
      shadeIdx = 0;
      while (!((geo >> shadeIdx) & 1u)) shadeIdx++;
    }
  }
  const uint32_t flags = p.flags;
  const bool dtest = (flags & SGL_PF_DEPTH_TEST) != 0;
  const int func = (flags >> SGL_PF_DEPTH_FUNC_SHIFT) & 7;
  uint32_t pass = 0;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    if (!((geo >> s) & 1u)) continue;
    float z = sglInterpZ(p, 2, bc[s][0], bc[s][1], bc[s][2]);
    // depth-range clipping only removes samples on the multisample path; with one sample per pixel
    // earlyZTest overwrites sample.inside with the (clamped) depth-test result (RendererSoft.cpp:797,871-874)
    if (NS > 1 && (z < 0.f || z > 1.f)) continue;
    z = gclamp(z, 0.f, 1.f);
    if (dtest) {
      if (!hasDepth) continue;
      if (!sglDepthTest(z, depth[s], func)) continue;
    }
    zOut[s] = z;
    pass |= 1u << s;
  }
  return pass;
}

// ---- colour packing / blending ----------------------------------------------------------------------------------
// setFrameColor(x, y, clamp(c,0,1) * 255.f) -> u8vec4 truncation (RendererSoft.cpp:368-374,949-954)
SGL_HD uint32_t sglPackColor(V4 c) {
  uint32_t r = (uint32_t) (int) (gclamp(c.x, 0.f, 1.f) * 255.f) & 0xffu;
  uint32_t g = (uint32_t) (int) (gclamp(c.y, 0.f, 1.f) * 255.f) & 0xffu;
  uint32_t b = (uint32_t) (int) (gclamp(c.z, 0.f, 1.f) * 255.f) & 0xffu;
  uint32_t a = (uint32_t) (int) (gclamp(c.w, 0.f, 1.f) * 255.f) & 0xffu;
  return r | (g << 8) | (b << 16) | (a << 24);
}

// after blending the reference stores u8vec4(colour * 255.f) WITHOUT clamping again (RendererSoft.cpp:371-374): the
// float->u8 conversion of an out-of-range value keeps the low byte of the truncated int32 (x86 cvttps2dq + pack)
SGL_HD uint32_t sglPackColorWrap(V4 c) {
  uint32_t r = (uint32_t) (int) (c.x * 255.f) & 0xffu;
  uint32_t g = (uint32_t) (int) (c.y * 255.f) & 0xffu;
  uint32_t b = (uint32_t) (int) (c.z * 255.f) & 0xffu;
  uint32_t a = (uint32_t) (int) (c.w * 255.f) & 0xffu;
  return r | (g << 8) | (b << 16) | (a << 24);
}

SGL_HD float sglBlendFactorA(float src, float srcA, float dst, float dstA, int factor) {
  switch (factor) {   // calcBlendFactor<float> (BlendSoft.h:14-30)
    case 0: return 0.f;
    case 1: return 1.f;
    case 2: return src;
    case 3: return srcA;
    case 4: return dst;
    case 5: return dstA;
    case 6: return 1.f - src;
    case 7: return 1.f - srcA;
    case 8: return 1.f - dst;
    case 9: return 1.f - dstA;
  }
  return 0.f;
}
SGL_HD float sglBlendFunc(float s, float d, int func) {   // calcBlendFunc (BlendSoft.h:32-42)
  switch (func) {
    case 0: return s + d;
    case 1: return s - d;
    case 2: return d - s;
    case 3: return gmin(s, d);
    case 4: return gmax(s, d);
  }
  return s + d;
}
// processColorBlending + calcBlendColor (RendererSoft.cpp:397-407, BlendSoft.h:44-56); src already clamped
SGL_HD V4 sglBlend(const SglRenderStates &rs, V4 src, uint32_t dstPacked) {
  V4 dst = sglUnpackRGBA(dstPacked);
  V4 r;
  float s[3] = {src.x, src.y, src.z}, dd[3] = {dst.x, dst.y, dst.z}, o[3];
  for (int i = 0; i < 3; i++) {
    float sf = sglBlendFactorA(s[i], src.w, dd[i], dst.w, rs.blend_src_rgb);
    float df = sglBlendFactorA(s[i], src.w, dd[i], dst.w, rs.blend_dst_rgb);
    o[i] = sglBlendFunc(s[i] * sf, dd[i] * df, rs.blend_func_rgb);
  }
  float sa = sglBlendFactorA(src.w, src.w, dst.w, dst.w, rs.blend_src_alpha);
  float da = sglBlendFactorA(src.w, src.w, dst.w, dst.w, rs.blend_dst_alpha);
  r = v4(o[0], o[1], o[2], sglBlendFunc(src.w * sa, dst.w * da, rs.blend_func_alpha));
  return r;
}

// ---- shading of a triangle fragment ------------------------------------------------------------------------------
// barycentrics of the shading sample of pixel (px,py): centre if inside else first covered sample; perspective
// correction only if that sample is geometrically inside (RendererSoft.cpp:776-803, RendererInternal.h:116-124)
template<int NS>
SGL_HD void sglShadingBarycentric(const SglPrim &p, const SglTriEdge &e, int px, int py, int shadeIdx, float *bc,
                                  float &spx, float &spy, float &z, float &w) {
  float fx = (float) px, fy = (float) py;
  float ox, oy;
  sglSampleOffset(NS, shadeIdx, ox, oy);
  if (NS > 1 && shadeIdx == 4) { spx = xadd(fx, 0.5f); spy = xadd(fy, 0.5f); }
  else { spx = xadd(ox, fx); spy = xadd(oy, fy); }
  bool inside = sglBarycentric(e, spx, spy, bc[0], bc[1], bc[2]);
  z = 0.f;
  w = 0.f;
  if (inside) {
    z = sglInterpZ(p, 2, bc[0], bc[1], bc[2]);
    w = sglInterpZ(p, 3, bc[0], bc[1], bc[2]);
    float s = xdiv(1.f, w);   // bc *= (1/w) * vertW   (vertW = per-vertex 1/w)
    bc[0] = xmul(xmul(s, p.v[0][3]), bc[0]);
    bc[1] = xmul(xmul(s, p.v[1][3]), bc[1]);
    bc[2] = xmul(xmul(s, p.v[2][3]), bc[2]);
  }
}

// shading sample selection for an arbitrary (helper) pixel of the quad: coverage over the NS samples + centre
template<int NS>
SGL_HD int sglPickShadingSample(const SglTriEdge &e, int px, int py) {
  if (NS == 1) return 0;
  float fx = (float) px, fy = (float) py;
  float b0, b1, b2;
  if (sglBarycentric(e, xadd(fx, 0.5f), xadd(fy, 0.5f), b0, b1, b2)) return 4;
  for (int s = 0; s < 4; s++) {
    float ox, oy;
    sglSampleOffset(NS, s, ox, oy);
    if (sglBarycentric(e, xadd(ox, fx), xadd(oy, fy), b0, b1, b2)) return s;
  }
  return 4;   // nothing covered: sampleShading stays at the centre sample
}

// interpolateBarycentric[SIMD] (RendererSoft.cpp:1085-1163): out = fma(in2, b2, fma(in1, b1, in0 * b0))
SGL_HD void sglInterpVaryings(float *out, const float *in0, const float *in1, const float *in2, int n, const float *bc) {
  for (int i = 0; i < n; i++) out[i] = xfma(in2[i], bc[2], xfma(in1[i], bc[1], xmul(in0[i], bc[0])));
}
// fixed-size form (N multiple of 4, rows are 32-byte aligned): 128-bit loads, fully unrolled so the result stays in registers
template<int N>
SGL_HD void sglInterpVaryingsN(float *out, const float *in0, const float *in1, const float *in2, const float *bc) {
#if defined(__CUDA_ARCH__)
  const float4 *a = reinterpret_cast<const float4 *>(in0), *b = reinterpret_cast<const float4 *>(in1), *c = reinterpret_cast<const float4 *>(in2);
#pragma unroll
  for (int i = 0; i < N / 4; i++) {
    float4 x = __ldg(a + i), y = __ldg(b + i), z = __ldg(c + i);
    out[4 * i + 0] = xfma(z.x, bc[2], xfma(y.x, bc[1], xmul(x.x, bc[0])));
    out[4 * i + 1] = xfma(z.y, bc[2], xfma(y.y, bc[1], xmul(x.y, bc[0])));
    out[4 * i + 2] = xfma(z.z, bc[2], xfma(y.z, bc[1], xmul(x.z, bc[0])));
    out[4 * i + 3] = xfma(z.w, bc[2], xfma(y.w, bc[1], xmul(x.w, bc[0])));
  }
#else
  sglInterpVaryings(out, in0, in1, in2, N, bc);
#endif
}

SGL_HD bool sglDrawNeedsDeriv(const SglDrawRec &d) {
  if (d.shader != SGL_SHADER_PBR && d.shader != SGL_SHADER_BLINNPHONG) return false;
  for (int s = 0; s < 4; s++) {   // albedo / normal / emissive / ao are the samplers that get a lodFunc
    if ((d.defines >> s) & 1u)
      if (d.samplers[s].tex >= 0 && d.samplers[s].filter > SGL_FILTER_LINEAR) return true;
  }
  return false;
}

template<int NS>
SGL_HD V4 sglShadeTriangle(const SglDrawRec &d, const SglTexObj *textures, const SglPrim &p, const SglPrimVerts &pv, int px,
                           int py, int shadeIdx) {
  SglTriEdge e = sglTriEdge(p);
  float bc[3], spx, spy, z, w;
  sglShadingBarycentric<NS>(p, e, px, py, shadeIdx, bc, spx, spy, z, w);
  const int stride = d.varyingStride;
  float vary[32];
  const float *in0 = d.varyings + (size_t) pv.i0 * stride;
  const float *in1 = d.varyings + (size_t) pv.i1 * stride;
  const float *in2 = d.varyings + (size_t) pv.i2 * stride;
  switch (d.shader) {   // sizeof(ShaderVaryings) of each program, rounded up to whole float4s (rows are padded to 8 floats)
    case SGL_SHADER_PBR: sglInterpVaryingsN<28>(vary, in0, in1, in2, bc); break;
    case SGL_SHADER_BLINNPHONG: sglInterpVaryingsN<32>(vary, in0, in1, in2, bc); break;
    case SGL_SHADER_BASIC: break;
    default: sglInterpVaryingsN<4>(vary, in0, in1, in2, bc); break;   // skybox / IBL: vec3 position; FXAA: vec2
  }
  SglFsCtx c;
  c.draw = &d;
  c.textures = textures;
  c.derivValid = false;
  if (sglDrawNeedsDeriv(d)) {
    // triangle-anchored quad of this pixel (RendererSoft.cpp:756-763); texcoord is varying 0..1 in both shaders
    int qx, qy;
    if (p.flags & SGL_PF_IRREGULAR) {
      sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, qx);
      sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, qy);
    } else {
      qx = p.bx0 + ((px - p.bx0) & ~1);
      qy = p.by0 + ((py - p.by0) & ~1);
    }
    V2 uvq[3];
    for (int k = 0; k < 3; k++) {
      int hx = qx + (k == 1 ? 1 : 0), hy = qy + (k == 2 ? 1 : 0);
      if (hx == px && hy == py) { uvq[k] = v2(vary[0], vary[1]); continue; }
      int si = sglPickShadingSample<NS>(e, hx, hy);
      float hb[3], a, b2, zz, ww;
      sglShadingBarycentric<NS>(p, e, hx, hy, si, hb, a, b2, zz, ww);
      uvq[k].x = xfma(in2[0], hb[2], xfma(in1[0], hb[1], xmul(in0[0], hb[0])));
      uvq[k].y = xfma(in2[1], hb[2], xfma(in1[1], hb[1], xmul(in0[1], hb[0])));
    }
    c.derivValid = true;
    c.uv0 = uvq[0];
    c.uv1 = uvq[1];
    c.uv2 = uvq[2];
  }
  return sglFragmentShader(c, vary);
}
