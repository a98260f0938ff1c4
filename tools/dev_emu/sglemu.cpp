// DEVELOPMENT TOOL -- NOT PART OF THE PRODUCT, NOT AN ORACLE, NOT A FALLBACK.
// Compiles the device headers of softglrender_b200/csrc for the host and runs them serially behind the same C ABI,
// so that arithmetic of the device functions can be debugged against oracle/_ref in a container without a GPU.
// It is built only by tools/dev_emu/build.sh, is never loaded by the package, tests or bench, and brute-forces
// every pixel against every primitive (no binning, no tiles).  The shipped library has no CPU path (sgl_init fails
// without a CUDA device).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../softglrender_b200/csrc/sgl_pixel.h"

namespace {
struct Buf { std::vector<uint8_t> d; };
struct Tex { bool alive = false; SglTextureDesc desc{}; SglTexObj obj{}; std::vector<uint8_t> mem, res; };
std::vector<Buf> buffers(1);
std::vector<Tex> textures(1);
std::vector<SglTexObj> texTable(1);
std::vector<SglDrawRec> draws;
bool inPass = false;
int colorTex, colorLayer, colorLevel, depthTex, clrC, clrD;
float clearColor[4], clearDepth, vpX, vpY, vpW, vpH;
std::string err;
unsigned long long overflowCount = 0;

struct HostAlloc {
  static constexpr bool kRecords = true;
  void consume(const SglDrawRec &, const SglPrim &) {}
  std::vector<int> *vc, *ac;
  int newVertex(const SglDrawRec &d) { int e = (*d.vertexCounter)++; int i = d.vertexCount + e; return i < d.vertexCap ? i : -1; }
  int newAppendSlots(const SglDrawRec &d, int n) { int a = *d.appendCounter; *d.appendCounter += n; return a + n <= d.appendCap ? d.appendBase + a : -1; }
  void overflow() { overflowCount++; }
  bool binPrim(int, const SglPrim &) { return false; }
};
int levelCount(const SglTextureDesc &d) {
  if (!d.use_mipmaps) return 1;
  int m = std::max(d.width, d.height), n = 0;
  while ((1 << (n + 1)) <= m) n++;
  return n + 1;
}
size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

extern "C" {
const char *sgl_last_error(void) { return err.c_str(); }
int sgl_init(int, int, int) { return 0; }
int sgl_shutdown(void) { return 0; }
int sgl_set_stream(void *) { return 0; }
int sgl_wait_idle(void) { return 0; }
int sgl_get_counters(SglCounters *o) { memset(o, 0, sizeof(*o)); o->clip_overflow = overflowCount; return 0; }
int sgl_reset_counters(void) { return 0; }
int sgl_timer_begin(void) { return 0; }
int sgl_timer_end(float *ms) { *ms = 0; return 0; }
int sgl_tile_size(void) { return SGL_TILE; }
int sgl_set_tile_owner_map(const uint8_t *, int, int) { return 0; }
int sgl_texture_set_shard_halo(int, int) { return 0; }

static const char *kBlocks[8][4] = {{}, {"UniformsModel", "UniformsMaterial"}, {"UniformsModel", "UniformsScene", "UniformsMaterial"},
  {"UniformsModel", "UniformsScene", "UniformsMaterial"}, {"UniformsModel"}, {"UniformsModel"}, {"UniformsModel", "UniformsPrefilter"}, {"UniformsQuadFilter"}};
static const int kBlockOff[8][4] = {{}, {0, 256}, {0, 256, 320}, {0, 256, 320}, {0}, {0}, {0, 256}, {0}};
static const char *kSamplers[8][8] = {{}, {}, {"u_albedoMap", "u_normalMap", "u_emissiveMap", "u_aoMap", "u_shadowMap"},
  {"u_albedoMap", "u_normalMap", "u_emissiveMap", "u_aoMap", "u_metalRoughnessMap", "u_irradianceMap", "u_prefilterMap"},
  {"u_equirectangularMap", "u_cubeMap"}, {"u_cubeMap"}, {"u_cubeMap"}, {"u_screenTexture"}};
static const char *kDefines[8][8] = {{}, {}, {"ALBEDO_MAP", "NORMAL_MAP", "EMISSIVE_MAP", "AO_MAP"},
  {"ALBEDO_MAP", "NORMAL_MAP", "EMISSIVE_MAP", "AO_MAP", "METALROUGHNESS_MAP"}, {"EQUIRECTANGULAR_MAP"}, {}, {}, {}};
int sgl_shader_uniform_offset(int s, const char *n) { if (s < 1 || s > 7) return -1; for (int i = 0; i < 4 && kBlocks[s][i]; i++) if (!strcmp(kBlocks[s][i], n)) return kBlockOff[s][i]; return -1; }
int sgl_shader_sampler_slot(int s, const char *n) { if (s < 1 || s > 7) return -1; for (int i = 0; i < 8 && kSamplers[s][i]; i++) if (!strcmp(kSamplers[s][i], n)) return i; return -1; }
int sgl_shader_define_bit(int s, const char *n) { if (s < 1 || s > 7) return -1; for (int i = 0; i < 8 && kDefines[s][i]; i++) if (!strcmp(kDefines[s][i], n)) return i; return -1; }
int sgl_shader_uniform_size(int s) { return (s < 1 || s > 7) ? -1 : sglShaderInfo(s).uniformBytes; }
int sgl_shader_varying_floats(int s) { return (s < 1 || s > 7) ? -1 : sglShaderInfo(s).varyingCount; }

int sgl_buffer_create(size_t bytes, const void *data, int *h) {
  Buf b; b.d.resize(std::max<size_t>(bytes, 16)); if (data) memcpy(b.d.data(), data, bytes);
  buffers.push_back(std::move(b)); *h = (int) buffers.size() - 1; return 0;
}
int sgl_buffer_upload(int h, size_t off, size_t bytes, const void *data) { memcpy(buffers[h].d.data() + off, data, std::min(bytes, buffers[h].d.size() - off)); return 0; }
int sgl_buffer_destroy(int h) { buffers[h].d.clear(); return 0; }

int sgl_texture_create(const SglTextureDesc *desc, int *h) {
  textures.emplace_back();
  Tex &t = textures.back();
  t.alive = true; t.desc = *desc;
  SglTexObj &o = t.obj;
  memset(&o, 0, sizeof(o));
  o.width = desc->width; o.height = desc->height; o.levels = levelCount(*desc); o.layers = desc->type == SGL_TEX_CUBE ? 6 : 1;
  o.format = desc->format; o.samples = desc->multi_sample ? 4 : 1; o.layout = desc->layout;
  size_t off = 0;
  for (int l = 0; l < o.levels; l++) { o.levelOffset[l] = off; off += alignUp(sglLevelTexels(o.layout, sglLevelDim(o.width, l), sglLevelDim(o.height, l)) * 4 * o.samples, 256); }
  o.layerStride = off;
  t.mem.assign(off * o.layers, 0);
  if (desc->multi_sample && desc->format == SGL_FMT_RGBA8) t.res.assign((size_t) o.width * o.height * 4, 0);
  *h = (int) textures.size() - 1;
  return 0;
}
static uint32_t dummyTexel[64];
static void fixPtrs() {
  texTable.resize(textures.size());
  {
    SglTexObj &o = texTable[0];
    memset(&o, 0, sizeof(o));
    o.base = (uint8_t *) dummyTexel;
    o.width = o.height = o.levels = o.layers = o.samples = 1;
    o.format = SGL_FMT_RGBA8; o.layout = SGL_LAYOUT_LINEAR;
  }
  for (size_t i = 1; i < textures.size(); i++) {
    textures[i].obj.base = textures[i].mem.data();
    textures[i].obj.resolve = textures[i].res.empty() ? nullptr : textures[i].res.data();
    texTable[i] = textures[i].obj;
  }
}
int sgl_texture_destroy(int h) { textures[h].alive = false; textures[h].mem.clear(); return 0; }
int sgl_texture_level_size(int h, int level, int *w, int *hh) { *w = sglLevelDim(textures[h].obj.width, level); *hh = sglLevelDim(textures[h].obj.height, level); return 0; }
int sgl_texture_upload(int h, int layer, int level, const void *data) {
  fixPtrs();
  Tex &t = textures[h];
  int w = sglLevelDim(t.obj.width, level), hh = sglLevelDim(t.obj.height, level);
  uint32_t *dst = (uint32_t *) (t.obj.base + (size_t) layer * t.obj.layerStride + t.obj.levelOffset[level]);
  const uint32_t *src = (const uint32_t *) data;
  for (int y = 0; y < hh; y++) for (int x = 0; x < w; x++) dst[sglTexelIndex(t.obj.layout, w, x, y)] = src[(size_t) y * w + x];
  return 0;
}
int sgl_texture_gen_mips(int h) {
  fixPtrs();
  Tex &t = textures[h];
  for (int layer = 0; layer < t.obj.layers; layer++)
    for (int level = 1; level < t.obj.levels; level++) {
      int ow = sglLevelDim(t.obj.width, level), oh = sglLevelDim(t.obj.height, level);
      int iw = sglLevelDim(t.obj.width, level - 1), ih = sglLevelDim(t.obj.height, level - 1);
      float rx = xdiv((float) iw, (float) ow), ry = xdiv((float) ih, (float) oh);
      SglSampler s; s.tex = &t.obj; s.filter = SGL_FILTER_LINEAR; s.wrap = SGL_WRAP_CLAMP_TO_EDGE; s.border = 0;
      uint32_t *dst = (uint32_t *) (t.obj.base + (size_t) layer * t.obj.layerStride + t.obj.levelOffset[level]);
      for (int y = 0; y < oh; y++) for (int x = 0; x < ow; x++) {
        float u = xadd(xmul((float) x, rx), xmul(0.5f, rx)), v = xadd(xmul((float) y, ry), xmul(0.5f, ry));
        dst[sglTexelIndex(t.obj.layout, ow, x, y)] = sglPixelBilinear(s, layer, level - 1, u, v);
      }
    }
  return 0;
}
int sgl_texture_readback(int h, int layer, int level, int kind, void *out, size_t bytes) {
  fixPtrs();
  Tex &t = textures[h];
  int w = sglLevelDim(t.obj.width, level), hh = sglLevelDim(t.obj.height, level);
  if (kind == 1) { memcpy(out, t.res.data(), std::min(bytes, t.res.size())); return 0; }
  const uint32_t *src = (const uint32_t *) (t.obj.base + (size_t) layer * t.obj.layerStride + t.obj.levelOffset[level]);
  uint32_t *dst = (uint32_t *) out;
  if (t.obj.samples > 1) { memcpy(out, src, (size_t) w * hh * 16); return 0; }
  for (int y = 0; y < hh; y++) for (int x = 0; x < w; x++) dst[(size_t) y * w + x] = src[sglTexelIndex(t.obj.layout, w, x, y)];
  return 0;
}
int sgl_texture_device_ptr(int, int, int, int, void **, size_t *) { return -1; }

int sgl_pass_begin(int c, int cl, int clv, int d, int fc, int fd, const float cc[4], float cd) {
  inPass = true; colorTex = c; colorLayer = cl; colorLevel = clv; depthTex = d; clrC = fc; clrD = fd;
  memcpy(clearColor, cc, 16); clearDepth = cd; draws.clear(); return 0;
}
int sgl_set_viewport(int x, int y, int w, int h) { vpX = x; vpY = y; vpW = w; vpH = h; return 0; }
int sgl_draw(const SglDraw *draw) {
  SglDrawRec r; memset(&r, 0, sizeof(r));
  memcpy(r.uniforms, draw->uniforms, std::min<size_t>(draw->uniform_bytes, SGL_MAX_UNIFORM_BYTES));
  for (int s = 0; s < 8; s++) {
    const SglSamplerBinding &b = draw->samplers[s];
    bool ok = b.texture > 0 && b.texture < (int) textures.size() && textures[b.texture].alive;
    r.samplers[s].tex = ok ? b.texture : -1; r.samplers[s].filter = b.filter_min; r.samplers[s].wrap = b.wrap;
    float bc = b.border == SGL_BORDER_WHITE ? 1.f : 0.f;
    if (ok && textures[b.texture].obj.format == SGL_FMT_FLOAT32) memcpy(&r.samplers[s].border, &bc, 4);
    else r.samplers[s].border = b.border == SGL_BORDER_WHITE ? 0xFFFFFFFFu : 0u;
    if (ok) { fixPtrs(); if (sglSamplerIsSimple(draw->shader, s, textures[b.texture].obj, b.filter_min, b.wrap)) r.fastSamplers |= 1u << s; }
  }
  r.rs = draw->states; r.shader = draw->shader; r.defines = draw->defines;
  r.vpX = vpX; r.vpY = vpY; r.vpW = vpW; r.vpH = vpH;
  r.vertexIn = (const float *) buffers[draw->vertex_buffer].d.data();
  r.indices = (const int32_t *) buffers[draw->index_buffer].d.data();
  r.vertexCount = draw->vertex_count; r.indexCount = draw->index_count;
  SglShaderInfo info = sglShaderInfo(draw->shader);
  r.varyingStride = info.varyingStride; r.varyingCount = info.varyingCount;
  r.pointSize = 1.f; if (draw->shader == SGL_SHADER_BASIC) memcpy(&r.pointSize, r.uniforms + 268, 4);
  r.hasColor = colorTex != 0;
  draws.push_back(r);
  return 0;
}

}  // extern "C"
template<int NS>
static void rasterAll(SglPassParams &P, std::vector<uint32_t> &order) {
  const bool hasColor = P.colorBase != nullptr, hasDepth = P.depthBase != nullptr;
  std::vector<SglPixelState<NS>> state((size_t) P.fbW * P.fbH);
  for (int py = 0; py < P.fbH; py++)
    for (int px = 0; px < P.fbW; px++) {
      size_t pix = (size_t) py * P.fbW + px;
      SglPixelState<NS> &st = state[pix];
      for (int s = 0; s < NS; s++) {
        st.depth[s] = (hasDepth && !P.clearDepthFlag) ? P.depthBase[pix * NS + s] : P.clearDepth;
        st.color[s] = (hasColor && !P.clearColorFlag) ? ((uint32_t *) P.colorBase)[pix * NS + s] : P.clearColor;
        st.owner[s] = SGL_OWNER_NONE;
      }
    }
  for (uint32_t slot : order) {   // primitive-major: same per-pixel order as the tile kernel, O(sum of bbox areas)
    const SglPrim &p = P.prims[slot];
    int x0 = std::max<int>(p.bx0, 0), y0 = std::max<int>(p.by0, 0), x1 = std::min<int>(p.bx1, P.fbW - 1), y1 = std::min<int>(p.by1, P.fbH - 1);
    for (int py = y0; py <= y1; py++)
      for (int px = x0; px <= x1; px++) sglPixelPrim<NS>(P, p, slot, px, py, state[(size_t) py * P.fbW + px], hasColor, hasDepth);
  }
  for (int py = 0; py < P.fbH; py++)
    for (int px = 0; px < P.fbW; px++) {
      size_t pix = (size_t) py * P.fbW + px;
      SglPixelState<NS> &st = state[pix];
      if (hasColor) sglFlushPixel<NS>(P, px, py, st);
      for (int s = 0; s < NS; s++) {
        if (hasDepth) P.depthBase[pix * NS + s] = st.depth[s];
        if (hasColor) ((uint32_t *) P.colorBase)[pix * NS + s] = st.color[s];
      }
      if (hasColor && NS == 4 && P.resolveBase) {
        uint32_t r = 0;
        for (int c = 0; c < 4; c++) { uint32_t sum = 0; for (int s = 0; s < NS; s++) sum += (st.color[s] >> (8 * c)) & 0xffu; r |= (sum / NS) << (8 * c); }
        ((uint32_t *) P.resolveBase)[pix] = r;
      }
    }
}

extern "C" {
int sgl_pass_end(void) {
  inPass = false;
  fixPtrs();
  Tex *ct = colorTex ? &textures[colorTex] : nullptr, *dt = depthTex ? &textures[depthTex] : nullptr;
  int fbW = ct ? sglLevelDim(ct->obj.width, colorLevel) : dt->obj.width, fbH = ct ? sglLevelDim(ct->obj.height, colorLevel) : dt->obj.height;
  int samples = ct ? ct->obj.samples : dt->obj.samples;
  int nDraws = (int) draws.size(), primSlots = 0, keyBase = 0;
  std::vector<std::vector<float>> clip(nDraws), frag(nDraws), vout(nDraws), vary(nDraws);
  std::vector<std::vector<int32_t>> mask(nDraws);
  std::vector<int32_t> counters(2 * std::max(nDraws, 1), 0);
  for (int i = 0; i < nDraws; i++) {
    SglDrawRec &r = draws[i];
    int pt = r.rs.primitive_type, per = pt == SGL_PRIM_TRIANGLE ? 3 : (pt == SGL_PRIM_LINE ? 2 : 1);
    r.inputPrims = r.indexCount / per;
    bool fill = pt == SGL_PRIM_TRIANGLE && r.rs.polygon_mode == SGL_POLY_FILL;
    r.slotsPerPrim = (pt == SGL_PRIM_TRIANGLE && !fill) ? 3 : 1;
    int extra = fill ? 12 * r.inputPrims : (pt == SGL_PRIM_POINT ? 0 : (pt == SGL_PRIM_LINE ? 2 : 6) * r.inputPrims);
    r.vertexCap = r.vertexCount + extra;
    r.appendCap = fill ? 6 * r.inputPrims : 0;
    r.primBase = primSlots; r.appendBase = primSlots + r.inputPrims * r.slotsPerPrim; primSlots = r.appendBase + r.appendCap;
    r.keyBase = keyBase; keyBase += r.inputPrims * r.slotsPerPrim + (fill ? 6 * r.inputPrims : 0);
    clip[i].assign((size_t) r.vertexCap * 4, 0); frag[i].assign((size_t) r.vertexCap * 4, 0); mask[i].assign(r.vertexCap, 0);
    vout[i].assign((size_t) std::max(extra, 1) * 16, 0); vary[i].assign((size_t) r.vertexCap * std::max(r.varyingStride, 1), 0);
    r.clipPos = clip[i].data(); r.fragPos = frag[i].data(); r.clipMask = mask[i].data(); r.vertexOut = vout[i].data(); r.varyings = vary[i].data();
    r.vertexCounter = &counters[2 * i]; r.appendCounter = &counters[2 * i + 1];
  }
  std::vector<SglPrim> prims(std::max(primSlots, 1));
  for (auto &p : prims) p.flags = 0;
  std::vector<SglPrimVerts> pverts(std::max(primSlots, 1));
  std::vector<uint32_t> keys(std::max(primSlots, 1), 0xFFFFFFFFu);
  SglSetupOut so = {prims.data(), pverts.data(), keys.data()};
  HostAlloc alloc;
  for (int i = 0; i < nDraws; i++) {
    SglDrawRec &r = draws[i];
    for (int v = 0; v < r.vertexCount; v++) sglProcessVertex(r, v, r.vertexIn + (size_t) v * 16);
    for (int p = 0; p < r.inputPrims; p++) sglProcessInputPrim(r, (uint32_t) i, p, dt != nullptr, so, alloc);
  }
  std::vector<uint32_t> order;
  for (int s = 0; s < primSlots; s++) if (prims[s].flags & SGL_PF_VALID) order.push_back((uint32_t) s);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  SglPassParams P; memset(&P, 0, sizeof(P));
  P.colorBase = ct ? ct->obj.base + (size_t) colorLayer * ct->obj.layerStride + ct->obj.levelOffset[colorLevel] : nullptr;
  P.depthBase = dt ? (float *) dt->obj.base : nullptr;
  P.resolveBase = (ct && samples > 1) ? ct->obj.resolve : nullptr;
  P.fbW = fbW; P.fbH = fbH; P.samples = samples; P.clearColorFlag = clrC; P.clearDepthFlag = clrD;
  uint32_t c = 0; for (int k = 0; k < 4; k++) c |= ((uint32_t) (uint8_t) (int) (clearColor[k] * 255.f)) << (8 * k);
  P.clearColor = c; P.clearDepth = clearDepth;
  P.draws = draws.data(); P.drawCount = nDraws; P.prims = prims.data(); P.primVerts = pverts.data(); P.primKeys = keys.data(); P.primSlots = primSlots;
  P.textures = texTable.data();
  if (getenv("SGLEMU_VERBOSE")) {
    fprintf(stderr, "[emu] pass %dx%d s%d draws %d prims %zu\n", fbW, fbH, samples, nDraws, order.size());
    int tX = (fbW + SGL_TILE - 1) / SGL_TILE, tY = (fbH + SGL_TILE - 1) / SGL_TILE;
    std::vector<int> cnt(tX * tY, 0), cntBox(tX * tY, 0);
    long long big = 0;
    for (uint32_t slot : order) {
      const SglPrim &p = prims[slot];
      int x0 = std::max<int>(p.bx0, 0), y0 = std::max<int>(p.by0, 0), x1 = std::min<int>(p.bx1, fbW - 1), y1 = std::min<int>(p.by1, fbH - 1);
      if (x1 < x0 || y1 < y0) continue;
      int n = (x1 / SGL_TILE - x0 / SGL_TILE + 1) * (y1 / SGL_TILE - y0 / SGL_TILE + 1);
      if (n > SGL_BIG_PRIM_TILES) big++;
      for (int ty = y0 / SGL_TILE; ty <= y1 / SGL_TILE; ty++)
        for (int tx = x0 / SGL_TILE; tx <= x1 / SGL_TILE; tx++) {
          cntBox[ty * tX + tx]++;
          bool near = true;
          if ((p.flags & SGL_PF_KIND_MASK) == SGL_PK_TRIANGLE) {
            SglTriEdge e = sglTriEdge(p);
            near = !sglTriSurelyOutside(e, tx * SGL_TILE + 8.f, ty * SGL_TILE + 8.f, 8.f, 8.f);
          } else if ((p.flags & SGL_PF_KIND_MASK) == SGL_PK_LINE) near = sglLineNearRect(p, tx * SGL_TILE, ty * SGL_TILE, tx * SGL_TILE + 15, ty * SGL_TILE + 15);
          if (near) cnt[ty * tX + tx]++;
        }
    }
    auto stats = [&](std::vector<int> v, const char *nm) {
      std::sort(v.begin(), v.end());
      long long sum = 0; for (int c : v) sum += c;
      fprintf(stderr, "[emu]   %s per tile: sum %lld mean %.1f p50 %d p90 %d p99 %d max %d  (tiles %zu, big prims %lld)\n", nm, sum, (double) sum / v.size(),
              v[v.size() / 2], v[v.size() * 9 / 10], v[v.size() * 99 / 100], v.back(), v.size(), big);
    };
    stats(cntBox, "bbox "); stats(cnt, "culled");
    {
      std::vector<int> areas;
      for (uint32_t slot : order) {
        const SglPrim &p = prims[slot];
        int x0 = std::max<int>(p.bx0, 0), y0 = std::max<int>(p.by0, 0), x1 = std::min<int>(p.bx1, fbW - 1), y1 = std::min<int>(p.by1, fbH - 1);
        if (x1 >= x0 && y1 >= y0) areas.push_back((x1 - x0 + 1) * (y1 - y0 + 1));
      }
      std::sort(areas.begin(), areas.end());
      long long sum = 0; for (int a : areas) sum += a;
      int c256 = 0, c4096 = 0, c64k = 0;
      for (int a : areas) { c256 += a > 256; c4096 += a > 4096; c64k += a > 65536; }
      if (!areas.empty()) fprintf(stderr, "[emu]   pixel-range area: n %zu sum %lld p50 %d p90 %d p99 %d max %d  >256: %d >4096: %d >65536: %d\n", areas.size(), sum,
              areas[areas.size() / 2], areas[areas.size() * 9 / 10], areas[areas.size() * 99 / 100], areas.back(), c256, c4096, c64k);
    }
  }
  if (samples == 4) rasterAll<4>(P, order); else rasterAll<1>(P, order);
  draws.clear();
  return 0;
}
int sgl_kat_barycentric(const float *, const float *, int, float *, int *, float *) { return -1; }
int sgl_kat_sample(int, int, int, int, const float *, const float *, const int32_t *, int, int, uint32_t *) { return -1; }
int sgl_kat_blend(const SglRenderStates *, const float *, const float *, int, float *) { return -1; }
int sgl_kat_depth(int, const float *, const float *, int, int *) { return -1; }
}
