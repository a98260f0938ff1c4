// Per-pixel processing of one binned primitive and deferred shading ("flush").
// Each thread of the tile kernel owns one pixel: NS depth values, NS packed colours and NS "owners" live in
// registers.  Opaque fragments only record their owner (primitive slot + shading sample); the fragment shader
// runs once per distinct owner when the pixel is flushed -- at the end of the pass, or earlier when a blended
// fragment needs the destination colour.  Because colour is a pure function of (primitive, pixel) this is
// exactly the reference's result (rasterizationPixelQuad + processPerSampleOperations, RendererSoft.cpp:771-851,
// 358-407) without shading overdraw.
#pragma once
#include "sgl_setup.h"

template<int NS>
struct SglPixelState {
  float depth[NS];
  uint32_t color[NS];
  uint32_t owner[NS];
};

SGL_HD uint32_t sglOwner(uint32_t slot, int shadeIdx) { return slot | ((uint32_t) shadeIdx << 29); }

// colour of the fragment of primitive `slot` at pixel (px,py) with shading sample shadeIdx (triangles only)
// (not inlined: the tile kernel reaches the fragment shaders from three places -- flush, blending, points/lines)
struct SglShadeEnv {
  const SglDrawRec *draws;
  const SglPrim *prims;
  const SglPrimVerts *primVerts;
  const SglTexObj *textures;
};
SGL_HD SglShadeEnv sglShadeEnv(const SglPassParams &P) {
  SglShadeEnv e = {P.draws, P.prims, P.primVerts, P.textures};
  return e;
}

template<int NS>
SGL_HDN_T V4 sglShadeSlotImpl(SglShadeEnv env, uint32_t slot, int shadeIdx, int px, int py) {
  SglPrim p = env.prims[slot];
  if ((p.flags & SGL_PF_KIND_MASK) != SGL_PK_TRIANGLE) {
    // deferred point / line fragment: only recorded for programs without varyings (colour independent of the step)
    float none[1] = {0.f};
    SglFsCtx c;
    c.draw = &env.draws[p.draw];
    c.textures = env.textures;
    c.derivValid = false;
    return sglFragmentShader(c, none);
  }
  return sglShadeTriangle<NS>(env.draws[p.draw], env.textures, p, env.primVerts[slot], px, py, shadeIdx);
}
template<int NS>
SGL_HD V4 sglShadeSlot(const SglPassParams &P, uint32_t slot, int shadeIdx, int px, int py) {
  return sglShadeSlotImpl<NS>(sglShadeEnv(P), slot, shadeIdx, px, py);
}
// fragment shader on ready-made varyings (points / line steps)
SGL_HDN V4 sglRunFragmentShader(const SglDrawRec *draw, const SglTexObj *textures, const float *vary) {
  SglFsCtx c;
  c.draw = draw;
  c.textures = textures;
  c.derivValid = false;
  return sglFragmentShader(c, vary);
}

template<int NS>
SGL_HD void sglFlushPixel(const SglPassParams &P, int px, int py, SglPixelState<NS> &st) {
#pragma unroll
  for (int s = 0; s < NS; s++) {
    uint32_t o = st.owner[s];
    if (o == SGL_OWNER_NONE) continue;
    uint32_t c = sglPackColor(sglShadeSlot<NS>(P, o & 0x1fffffffu, (int) (o >> 29), px, py));
#pragma unroll
    for (int t = s; t < NS; t++)
      if (st.owner[t] == o) { st.color[t] = c; st.owner[t] = SGL_OWNER_NONE; }
  }
}

// write `src` (unclamped FS output) into the samples of `mask`, blending when the draw asks for it
template<int NS>
SGL_HD void sglWriteImmediate(const SglPassParams &P, const SglPrim &p, V4 src, uint32_t mask, int px, int py,
                              SglPixelState<NS> &st) {
  V4 c = v4(gclamp(src.x, 0.f, 1.f), gclamp(src.y, 0.f, 1.f), gclamp(src.z, 0.f, 1.f), gclamp(src.w, 0.f, 1.f));
  if (p.flags & SGL_PF_BLEND) {
    sglFlushPixel<NS>(P, px, py, st);
    const SglRenderStates &rs = P.draws[p.draw].rs;
#pragma unroll
    for (int s = 0; s < NS; s++)
      if ((mask >> s) & 1u) st.color[s] = sglPackColorWrap(sglBlend(rs, c, st.color[s]));
  } else {
    uint32_t pc = sglPackColor(c);
#pragma unroll
    for (int s = 0; s < NS; s++)
      if ((mask >> s) & 1u) { st.color[s] = pc; st.owner[s] = SGL_OWNER_NONE; }
  }
}

// depth test + write of a flat-depth fragment (points / line steps write every sample, RendererSoft.cpp:655-657)
template<int NS>
SGL_HD uint32_t sglFlatDepth(const SglPrim &p, float z, bool hasDepth, SglPixelState<NS> &st) {
  const uint32_t flags = p.flags;
  const bool dtest = (flags & SGL_PF_DEPTH_TEST) != 0;
  const int func = (flags >> SGL_PF_DEPTH_FUNC_SHIFT) & 7;
  uint32_t mask = 0;
  float zc = gclamp(z, 0.f, 1.f);
#pragma unroll
  for (int s = 0; s < NS; s++) {
    if (dtest) {
      if (!sglDepthTest(zc, st.depth[s], func)) continue;
      if (flags & SGL_PF_DEPTH_MASK) st.depth[s] = zc;
    }
    mask |= 1u << s;
  }
  return mask;
}

// rasterizationLine (RendererSoft.cpp:663-718) seen from one pixel: calls f(t, 1-t, z) for every Bresenham step whose
// lineWidth square covers (px,py), in step order
template<class F>
SGL_HD void sglLineVisit(const SglPrim &p, int px, int py, F &&f) {
  if (!sglLineNearRect(p, px, py, px, py)) return;
  const uint32_t flags = p.flags;
  int x0, y0, x1, y1;
  memcpy(&x0, &p.v[0][0], 4); memcpy(&y0, &p.v[0][1], 4); memcpy(&x1, &p.v[0][2], 4); memcpy(&y1, &p.v[0][3], 4);
  const bool steep = (flags & SGL_PF_STEEP) != 0;
  const float width = p.v[2][0];
  const int major = steep ? py : px, minor = steep ? px : py;
  const int dx = x1 - x0, dy = y1 - y0, ady = dy < 0 ? -dy : dy, sy = y1 > y0 ? 1 : -1;
  int reach = (int) ceilf(fabsf(width)) + 1;
  int k0 = major - x0 - reach, k1 = major - x0 + reach;
  if (k0 < 0) k0 = 0;
  if (k1 > dx) k1 = dx;
  for (int k = k0; k <= k1; k++) {
    int lo, hi;
    sglPointSpan((float) (x0 + k), width, lo, hi);       // the cheap axis first: of the 2*reach + 1 candidate steps usually one
    if (major < lo || major > hi) continue;              // survives, and only that one pays for the division in sglLineYSteps
    const int cy = y0 + sy * sglLineYSteps(k, dx, ady);
    sglPointSpan((float) cy, width, lo, hi);
    if (minor < lo || minor > hi) continue;
    float t = xdiv((float) k, (float) dx);           // (float)(x - x0) / (float)dx ; 0/0 = NaN for single-column lines
    float omt = xsub(1.f, t);
    f(t, omt, xMix(p.v[1][0], p.v[1][1], t, omt));
  }
}

template<int NS>
SGL_HD void sglPixelPrim(const SglPassParams &P, const SglPrim &p, uint32_t slot, int px, int py, SglPixelState<NS> &st,
                         bool hasColor, bool hasDepth) {
  if (px < p.bx0 || px > p.bx1 || py < p.by0 || py > p.by1) return;
  const uint32_t flags = p.flags;
  const uint32_t kind = flags & SGL_PF_KIND_MASK;
  if (kind == SGL_PK_TRIANGLE) {
    if (flags & SGL_PF_IRREGULAR) {
      const SglDrawRec &d = P.draws[p.draw];
      int q;
      if (!sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, q)) return;
      if (!sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, q)) return;
    }
    SglTriEdge e = sglTriEdge(p);
    float z[NS];
    int shadeIdx = 0;
    uint32_t mask = sglCoverTriangle<NS>(p, e, px, py, st.depth, hasDepth, z, shadeIdx);
    if (!mask) return;
    if ((flags & SGL_PF_DEPTH_TEST) && (flags & SGL_PF_DEPTH_MASK)) {
#pragma unroll
      for (int s = 0; s < NS; s++)
        if ((mask >> s) & 1u) st.depth[s] = z[s];
    }
    if (!hasColor) return;
    if (flags & SGL_PF_BLEND) {
      V4 c = sglShadeSlot<NS>(P, slot, shadeIdx, px, py);
      sglWriteImmediate<NS>(P, p, c, mask, px, py, st);
    } else {
      uint32_t o = sglOwner(slot, shadeIdx);
#pragma unroll
      for (int s = 0; s < NS; s++)
        if ((mask >> s) & 1u) st.owner[s] = o;
    }
    return;
  }
  if (!hasColor) return;
  const SglDrawRec &d = P.draws[p.draw];
  const SglPrimVerts pv = P.primVerts[slot];
  if (kind == SGL_PK_POINT) {
    uint32_t mask = sglFlatDepth<NS>(p, p.v[0][2], hasDepth, st);
    if (!mask) return;
    float vary[32];
    for (int i = 0; i < d.varyingCount; i++) vary[i] = d.varyings[(size_t) pv.i0 * d.varyingStride + i];
    sglWriteImmediate<NS>(P, p, sglRunFragmentShader(&d, P.textures, vary), mask, px, py, st);
    return;
  }
  // line: every Bresenham step k draws a lineWidth square; visit the steps whose square covers this pixel in order
  sglLineVisit(p, px, py, [&](float t, float omt, float z) {
    uint32_t mask = sglFlatDepth<NS>(p, z, hasDepth, st);
    if (!mask) return;
    float vary[32];
    const bool sw = (flags & SGL_PF_SWAPPED) != 0;
    const float *va = d.varyings + (size_t) (sw ? pv.i1 : pv.i0) * d.varyingStride;
    const float *vb = d.varyings + (size_t) (sw ? pv.i0 : pv.i1) * d.varyingStride;
    for (int i = 0; i < d.varyingCount; i++) vary[i] = xMix(va[i], vb[i], t, omt);
    sglWriteImmediate<NS>(P, p, sglRunFragmentShader(&d, P.textures, vary), mask, px, py, st);
  });
}

// ---- multisample colour storage ---------------------------------------------------------------------------------
// A pixel whose four samples hold the same colour (everything but primitive edges) keeps only its resolved colour -- which
// then equals each sample exactly, sum / 4 of four equal bytes -- and a 0 in the per-pixel mask; the 16-byte per-sample
// record is written for the other pixels only.  In a frame that starts with a clear and ends in the resolve nobody reads the
// per-sample image, and it was 33 MB of the shading kernel's 84 MB of DRAM traffic on config 2.  Readers (a pass that loads
// the attachment, the fused tail of a split pass) go through sglLoadMsColor; the per-sample read-back expands first
// (sglMsExpandKernel).  P.colorMask == null: every record is written, as before.
#if defined(__CUDACC__)
__device__ __forceinline__ void sglLoadMsColor(const SglPassParams &P, size_t pix, uint32_t *color) {
  if (P.colorMask && P.colorMask[pix] == 0) {
    const uint32_t r = reinterpret_cast<const uint32_t *>(P.resolveBase)[pix];
    color[0] = color[1] = color[2] = color[3] = r;
    return;
  }
  const uint4 cq = reinterpret_cast<const uint4 *>(P.colorBase)[pix];
  color[0] = cq.x; color[1] = cq.y; color[2] = cq.z; color[3] = cq.w;
}
// per-sample record (when needed) + mask; the caller writes the resolved colour
__device__ __forceinline__ void sglStoreMsColor(const SglPassParams &P, size_t pix, const uint32_t *color) {
  if (P.colorMask) {
    const bool same = color[0] == color[1] && color[1] == color[2] && color[2] == color[3];
    P.colorMask[pix] = same ? 0 : 1;
    if (same) return;
  }
  reinterpret_cast<uint4 *>(P.colorBase)[pix] = make_uint4(color[0], color[1], color[2], color[3]);
}
#endif
