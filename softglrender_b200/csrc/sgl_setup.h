// Primitive assembly, frustum clipping, face culling and primitive setup for ONE input primitive.
// Restates processPrimitiveAssembly / processClipping / clippingTriangle / clippingLine / clippingPoint /
// clippingNewVertex / interpolateVertex / processFaceCulling / rasterizationPolygons{Point,Line}
// (RendererSoft.cpp:192-257,409-559,575-622,956-969,1058-1070) as a device function that the setup kernel
// calls with one thread per input primitive.  `Alloc` supplies the two arena counters (atomicAdd on the GPU).
#pragma once
#include "sgl_raster.h"

struct SglSetupOut {          // where the setup kernel writes
  SglPrim *prims;
  SglPrimVerts *primVerts;
  uint32_t *primKeys;
};

SGL_HD const float *sglAttrPtr(const SglDrawRec &d, int i) {
  return i < d.vertexCount ? d.vertexIn + (size_t) i * 16 : d.vertexOut + (size_t) (i - d.vertexCount) * 16;
}
SGL_HD V4 sglLoadV4(const float *base, int i) {
  const float *p = base + (size_t) i * 4;
  return v4(p[0], p[1], p[2], p[3]);
}
SGL_HD void sglStoreV4(float *base, int i, V4 v) {
  float *p = base + (size_t) i * 4;
  p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
}

// vertexShaderImpl + perspective divide + viewport for vertex slot `idx` whose attributes are at `attr`
SGL_HD V4 sglProcessVertex(const SglDrawRec &d, int idx, const float *attr) {
  float vary[32];
  V4 clip = sglVertexShader(d, attr, vary);
  for (int k = 0; k < d.varyingStride; k++) d.varyings[(size_t) idx * d.varyingStride + k] = vary[k];
  sglStoreV4(d.clipPos, idx, clip);
  d.clipMask[idx] = sglClipMask(clip);
  sglStoreV4(d.fragPos, idx, sglToScreen(clip, d.vpX, d.vpY, d.vpW, d.vpH));
  return clip;
}

// clippingNewVertex: VS(mix(attributes of idx0, idx1, t)); returns the new index or -1 when the arena is full
template<class Alloc>
SGL_HD int sglClipNewVertex(const SglDrawRec &d, Alloc &alloc, int idx0, int idx1, float t, V4 *clipOut = nullptr) {
  int idx = alloc.newVertex(d);
  if (idx < 0) return -1;
  const float *a0 = sglAttrPtr(d, idx0), *a1 = sglAttrPtr(d, idx1);
  float *out = d.vertexOut + (size_t) (idx - d.vertexCount) * 16;
  float omt = xsub(1.f, t);
  float attr[16];
  for (int k = 0; k < 16; k++) attr[k] = out[k] = xMix(a0[k], a1[k], t, omt);
  V4 clip = sglProcessVertex(d, idx, attr);
  if (clipOut) *clipOut = clip;
  return idx;
}

// clippingTriangle: Sutherland-Hodgman over the 6 planes in fixed order.  poly[] receives up to 9 indices.
// returns the polygon size (0 = discarded, -1 = arena overflow)
template<class Alloc>
SGL_HD int sglClipTriangle(const SglDrawRec &d, Alloc &alloc, int i0, int i1, int i2, int *poly) {
  int mask = d.clipMask[i0] | d.clipMask[i1] | d.clipMask[i2];
  poly[0] = i0; poly[1] = i1; poly[2] = i2;
  if (mask == 0) return 3;
  int in[12], out[12];
  V4 cin[12], cout[12];          // clip positions of the working polygon (no store -> load round trips through memory)
  int nin = 3, nout = 0;
  in[0] = i0; in[1] = i1; in[2] = i2;
  cin[0] = sglLoadV4(d.clipPos, i0); cin[1] = sglLoadV4(d.clipPos, i1); cin[2] = sglLoadV4(d.clipPos, i2);
  for (int plane = 0; plane < 6; plane++) {
    if (!(mask & (1 << plane))) continue;
    if (nin < 3) return 0;                       // fullClip
    nout = 0;
    int idxPre = in[0];
    V4 cPre = cin[0];
    float dPre = sglPlaneDist(plane, cPre);
    in[nin] = idxPre;
    cin[nin] = cPre;
    for (int i = 1; i <= nin; i++) {
      int idx = in[i];
      V4 cc = cin[i];
      float dd = sglPlaneDist(plane, cc);
      if (dPre >= 0) { out[nout] = idxPre; cout[nout++] = cPre; }
      if (sglSignBit(dPre) != sglSignBit(dd)) {
        float t = dd < 0 ? xdiv(dPre, xsub(dPre, dd)) : xdiv(-dPre, xsub(dd, dPre));
        V4 cn;
        int nv = sglClipNewVertex(d, alloc, idxPre, idx, t, &cn);
        if (nv < 0) return -1;
        out[nout] = nv;
        cout[nout++] = cn;
      }
      idxPre = idx;
      cPre = cc;
      dPre = dd;
    }
    nin = nout;
    for (int i = 0; i < nout; i++) { in[i] = out[i]; cin[i] = cout[i]; }
  }
  if (nin < 3) return 0;
  for (int i = 0; i < nin; i++) poly[i] = in[i];
  return nin;
}

// clippingLine (RendererSoft.cpp:443-487); returns false when fully clipped; a/b are replaced by new vertices
template<class Alloc>
SGL_HD int sglClipLine(const SglDrawRec &d, Alloc &alloc, int &a, int &b) {
  int m0 = d.clipMask[a], m1 = d.clipMask[b];
  V4 c0 = sglLoadV4(d.clipPos, a), c1 = sglLoadV4(d.clipPos, b);
  float t0 = 0.f, t1 = 1.f;
  int mask = m0 | m1;
  if (mask != 0) {
    for (int i = 0; i < 6; i++) {
      if (!(mask & (1 << i))) continue;
      float d0 = sglPlaneDist(i, c0), d1 = sglPlaneDist(i, c1);
      if (d0 < 0 && d1 < 0) return 0;
      if (d0 < 0) {
        float t = xdiv(-d0, xsub(d1, d0));
        t0 = gmax(t0, t);                       // std::max(t0, t)
      } else {
        float t = xdiv(d0, xsub(d0, d1));
        t1 = gmin(t1, t);                       // std::min(t1, t)
      }
    }
  }
  if (m0) {
    int nv = sglClipNewVertex(d, alloc, a, b, t0);
    if (nv < 0) return -1;
    a = nv;
  }
  if (m1) {
    int nv = sglClipNewVertex(d, alloc, a, b, t1);   // relative to the already replaced first vertex (:481-486)
    if (nv < 0) return -1;
    b = nv;
  }
  return 1;
}

SGL_HD int16_t sglClamp16(int v) { return (int16_t) (v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

// closed form of the Bresenham error recurrence of rasterizationLine (RendererSoft.cpp:693-716):
// number of y steps taken before column k is emitted
SGL_HD int sglLineYSteps(int k, int dx, int ady) {
  if (dx <= 0) return 0;
  // 2*ady*k + dx - 1 <= 2*32767^2 + 32766 < 2^32: one 32-bit unsigned division covers every on-screen line
  if (dx < 32768 && ady < 32768 && k >= 0 && k <= dx) return (int) ((2u * (unsigned) ady * (unsigned) k + (unsigned) (dx - 1)) / (2u * (unsigned) dx));
  return (int) (((long long) 2 * ady * k + dx - 1) / ((long long) 2 * dx));
}

// Conservative test "can any step square of line `p` touch the pixel rectangle [rx0,rx1] x [ry0,ry1]?".  Step k sits at
// minor = y0 + sy*floor((2*ady*k + dx - 1) / (2*dx)), i.e. within 1 of the ideal line y0 + sy*slope*k (slope <= 1), and its
// square reaches `reach` pixels, so everything farther than 2*reach + 3 from the ideal line is never written.
SGL_HD bool sglLineNearRect(const SglPrim &p, int rx0, int ry0, int rx1, int ry1) {
  int x0, y0, y1;
  memcpy(&x0, &p.v[0][0], 4); memcpy(&y0, &p.v[0][1], 4); memcpy(&y1, &p.v[0][3], 4);
  const bool steep = (p.flags & SGL_PF_STEEP) != 0;
  const int a0 = steep ? ry0 : rx0, a1 = steep ? ry1 : rx1, b0 = steep ? rx0 : ry0, b1 = steep ? rx1 : ry1;
  const float slope = p.v[2][1];
  const float margin = (float) (2 * ((int) ceilf(fabsf(p.v[2][0])) + 1) + 3);
  const float e0 = slope * (float) (a0 - x0), e1 = slope * (float) (a1 - x0);     // slope >= 0: e0 <= e1
  const float d0 = (float) (y1 > y0 ? b0 - y0 : y0 - b1), d1 = (float) (y1 > y0 ? b1 - y0 : y0 - b0);
  return !(d1 < e0 - margin || d0 > e1 + margin);
}

// pixel range covered by rasterizationPoint(centre c, size s) on one axis: [(int)left, (int)right)
SGL_HD void sglPointSpan(float c, float size, int &lo, int &hi) {
  float left = xadd(xsub(c, xmul(size, 0.5f)), 0.5f);
  float right = xadd(left, size);
  lo = (int) left;
  hi = (int) right - 1;
}

SGL_HD bool sglSetupPoint(SglPrim &p, V4 fp, float size, uint32_t stateFlags, uint32_t draw) {
  int x0, x1, y0, y1;
  sglPointSpan(fp.x, size, x0, x1);
  sglPointSpan(fp.y, size, y0, y1);
  if (x1 < x0 || y1 < y0) return false;
  p.v[0][0] = fp.x; p.v[0][1] = fp.y; p.v[0][2] = fp.z; p.v[0][3] = fp.w;
  p.v[1][0] = size; p.v[1][1] = p.v[1][2] = p.v[1][3] = 0.f;
  p.v[2][0] = p.v[2][1] = p.v[2][2] = p.v[2][3] = 0.f;
  p.bx0 = sglClamp16(x0); p.bx1 = sglClamp16(x1); p.by0 = sglClamp16(y0); p.by1 = sglClamp16(y1);
  p.flags = stateFlags | SGL_PK_POINT | SGL_PF_FRONT;
  p.draw = draw;
  return true;
}

// rasterizationLine setup (RendererSoft.cpp:663-692)
SGL_HD bool sglSetupLine(SglPrim &p, V4 f0, V4 f1, float width, uint32_t stateFlags, uint32_t draw) {
  int x0 = (int) f0.x, y0 = (int) f0.y, x1 = (int) f1.x, y1 = (int) f1.y;
  float z0 = f0.z, z1 = f1.z, w0 = f0.w, w1 = f1.w;
  uint32_t fl = 0;
  int adx = x0 - x1; adx = adx < 0 ? -adx : adx;
  int ady = y0 - y1; ady = ady < 0 ? -ady : ady;
  if (adx < ady) {
    int t;
    t = x0; x0 = y0; y0 = t;
    t = x1; x1 = y1; y1 = t;
    fl |= SGL_PF_STEEP;
  }
  if (x0 > x1) {
    int t;
    t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
    float f;
    f = z0; z0 = z1; z1 = f;
    f = w0; w0 = w1; w1 = f;
    fl |= SGL_PF_SWAPPED;
  }
  memcpy(&p.v[0][0], &x0, 4); memcpy(&p.v[0][1], &y0, 4); memcpy(&p.v[0][2], &x1, 4); memcpy(&p.v[0][3], &y1, 4);
  p.v[1][0] = z0; p.v[1][1] = z1; p.v[1][2] = w0; p.v[1][3] = w1;
  p.v[2][0] = width; p.v[2][1] = p.v[2][2] = p.v[2][3] = 0.f;
  if (x1 > x0) p.v[2][1] = (float) (y1 > y0 ? y1 - y0 : y0 - y1) / (float) (x1 - x0);   // slope of the ideal line (culling only)
  // conservative pixel bbox of all step squares
  int ylo = y0 < y1 ? y0 : y1, yhi = y0 < y1 ? y1 : y0;
  int a0, a1, b0, b1, t0, t1;
  sglPointSpan((float) x0, width, a0, t0);
  sglPointSpan((float) x1, width, t1, a1);
  sglPointSpan((float) ylo, width, b0, t0);
  sglPointSpan((float) yhi, width, t1, b1);
  if (a1 < a0 || b1 < b0) return false;
  if (fl & SGL_PF_STEEP) { int t; t = a0; a0 = b0; b0 = t; t = a1; a1 = b1; b1 = t; }
  p.bx0 = sglClamp16(a0); p.bx1 = sglClamp16(a1); p.by0 = sglClamp16(b0); p.by1 = sglClamp16(b1);
  p.flags = stateFlags | SGL_PK_LINE | SGL_PF_FRONT | fl;
  p.draw = draw;
  return true;
}

// Emits one finished primitive into slot `slot` with order key `key`.
template<class Alloc>
SGL_HD void sglEmitPrim(const SglSetupOut &o, Alloc &alloc, const SglDrawRec &d, int slot, uint32_t key, const SglPrim &p,
                        int i0, int i1, int i2) {
  if (!Alloc::kRecords) { alloc.consume(d, p); return; }   // immediate-mode consumers (depth-only atomic path)
  const int binned = alloc.binPrim(slot, p);     // 0: counted into its tiles, 1: big list, 2: touches no tile this rank renders
  if (binned == 2) return;                       // the slot stays invalid (flags = 0), nobody will ask for its varyings
  SglPrimVerts pv = {(uint32_t) i0, (uint32_t) i1, (uint32_t) i2, 0u};
  o.primVerts[slot] = pv;
  o.primKeys[slot] = key;
  SglPrim q = p;
  if (binned == 1) q.flags |= SGL_PF_BIG;
  o.prims[slot] = q;
  if (d.vertexUsed) {                            // clip-generated vertices (index >= vertexCount) already have their varyings
    if (i0 < d.vertexCount) d.vertexUsed[i0] = 1;
    if (i1 < d.vertexCount) d.vertexUsed[i1] = 1;
    if (i2 < d.vertexCount) d.vertexUsed[i2] = 1;
  }
}

// Everything RendererSoft::draw() does for input primitive `i` of draw `drawIdx` between the vertex stage and
// rasterisation.  Slots/keys: originals live at primBase + i*slotsPerPrim + e with key keyBase + i*slotsPerPrim + e;
// fan triangles produced by clipping are appended (RendererSoft.cpp:227) -> slots from the append arena, keys
// keyBase + inputPrims*slotsPerPrim + 6*i + (k-1), which preserves "appended in source order after all originals".
template<class Alloc>
SGL_HD void sglProcessInputPrim(const SglDrawRec &d, uint32_t drawIdx, int i, bool hasDepth, const SglSetupOut &o, Alloc &alloc) {
  const SglRenderStates &rs = d.rs;
  const uint32_t sf = sglStateFlags(rs, hasDepth);
  const int baseSlot = d.primBase + i * d.slotsPerPrim;
  const uint32_t baseKey = (uint32_t) d.keyBase + (uint32_t) (i * d.slotsPerPrim);
  // invalidate this input primitive's original slots first
  for (int e = 0; Alloc::kRecords && e < d.slotsPerPrim; e++) {
    o.prims[baseSlot + e].flags = 0;
    o.primKeys[baseSlot + e] = baseKey + e;
  }
  SglPrim p;
  if (rs.primitive_type == SGL_PRIM_POINT) {
    int a = d.indices[i];
    if (!d.hasColor) return;                                  // rasterizationPoint returns without colour buffer
    if (d.clipMask[a] != 0) return;                           // clippingPoint
    if (sglSetupPoint(p, sglLoadV4(d.fragPos, a), d.pointSize, sf, drawIdx)) sglEmitPrim(o, alloc, d, baseSlot, baseKey, p, a, a, a);
    return;
  }
  if (rs.primitive_type == SGL_PRIM_LINE) {
    int a = d.indices[2 * i], b = d.indices[2 * i + 1];
    if (!d.hasColor) return;
    int r = sglClipLine(d, alloc, a, b);
    if (r <= 0) { if (r < 0) alloc.overflow(); return; }
    if (sglSetupLine(p, sglLoadV4(d.fragPos, a), sglLoadV4(d.fragPos, b), rs.line_width, sf, drawIdx))
      sglEmitPrim(o, alloc, d, baseSlot, baseKey, p, a, b, b);
    return;
  }
  int i0 = d.indices[3 * i], i1 = d.indices[3 * i + 1], i2 = d.indices[3 * i + 2];
  if (rs.polygon_mode == SGL_POLY_FILL) {
    int poly[12];
    int n = sglClipTriangle(d, alloc, i0, i1, i2, poly);
    if (n <= 0) { if (n < 0) alloc.overflow(); return; }
    int appended = n - 3;
    int appendSlot = appended > 0 ? alloc.newAppendSlots(d, appended) : 0;
    if (appended > 0 && appendSlot < 0) { alloc.overflow(); appended = 0; }
    for (int k = 0; k <= appended; k++) {
      int a = poly[0], b = poly[k + 1], c = poly[k + 2];
      V4 fa = sglLoadV4(d.fragPos, a), fb = sglLoadV4(d.fragPos, b), fc = sglLoadV4(d.fragPos, c);
      bool front = sglFrontFacing(fa, fb, fc);
      int slot = k == 0 ? baseSlot : appendSlot + (k - 1);
      uint32_t key = k == 0 ? baseKey : (uint32_t) d.keyBase + (uint32_t) (d.inputPrims * d.slotsPerPrim) + (uint32_t) (6 * i + (k - 1));
      if (Alloc::kRecords && k > 0) { o.prims[slot].flags = 0; o.primKeys[slot] = key; }
      if (rs.cull_face && !front) continue;
      if (sglSetupTriangle(p, fa, fb, fc, d.vpW, d.vpH, front, sf, drawIdx)) sglEmitPrim(o, alloc, d, slot, key, p, a, b, c);
    }
    return;
  }
  // polygon mode LINE / POINT: no triangle clipping, but face culling still applies (RendererSoft.cpp:221-224,277-299)
  {
    V4 fa = sglLoadV4(d.fragPos, i0), fb = sglLoadV4(d.fragPos, i1), fc = sglLoadV4(d.fragPos, i2);
    bool front = sglFrontFacing(fa, fb, fc);
    if (rs.cull_face && !front) return;
    if (!d.hasColor) return;
    int tri[3] = {i0, i1, i2};
    for (int e = 0; e < 3; e++) {
      if (rs.polygon_mode == SGL_POLY_POINT) {
        int a = tri[e];
        if (d.clipMask[a] != 0) continue;
        if (sglSetupPoint(p, sglLoadV4(d.fragPos, a), d.pointSize, sf, drawIdx))
          sglEmitPrim(o, alloc, d, baseSlot + e, baseKey + e, p, a, a, a);
      } else {
        int a = tri[e], b = tri[(e + 1) % 3];
        int r = sglClipLine(d, alloc, a, b);
        if (r <= 0) { if (r < 0) alloc.overflow(); continue; }
        if (sglSetupLine(p, sglLoadV4(d.fragPos, a), sglLoadV4(d.fragPos, b), rs.line_width, sf, drawIdx))
          sglEmitPrim(o, alloc, d, baseSlot + e, baseKey + e, p, a, b, b);
      }
    }
  }
}
