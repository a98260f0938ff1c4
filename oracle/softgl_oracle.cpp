// ORACLE -- test infrastructure, NOT product code (see softgl_oracle.h).
// Scalar CPU restatement of the reference's software pipeline.  All file:line citations are relative to
// /root/reference/src.  Compile with -ffp-contract=off; fused operations of the reference binary are explicit fmaf().
#include "softgl_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace SoftGL {
namespace {

int gNextId = 0;

// ------------------------------------------------------------------------------------------------ small maths
struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
f3 cross3(f3 a, f3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
f3 norm3(f3 a) { return a * (1.0f / std::sqrt(dot3(a, a))); }
float gmaxf(float x, float y) { return (x < y) ? y : x; }   // glm::max / std::max
float gminf(float x, float y) { return (y < x) ? y : x; }   // glm::min / std::min
float gclampf(float x, float lo, float hi) { return gminf(gmaxf(x, lo), hi); }
float bitsToFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
uint32_t floatToBits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// glm mat4 * vec4(p,1) as emitted for every VS (objdump: vmulps m1,y; vfmadd m0,x; vfmadd213 m2,z,+m3; vaddps)
f4 mulPoint(const float *m, f3 p) {
  f4 r;
  float *o = &r.x;
  for (int i = 0; i < 4; i++) o[i] = fmaf(p.z, m[8 + i], m[12 + i]) + fmaf(p.x, m[i], p.y * m[4 + i]);
  return r;
}
// glm::dot on aligned vec4 = _mm_dp_ps 0xff: (a0b0 + a1b1) + (a2b2 + a3b3)
float dpps(const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]); }

// ------------------------------------------------------------------------------------------------ sampling
struct Sampler {   // Sampler2DSoft / SamplerCubeSoft state (Render/Software/SamplerSoft.h:382-453)
  TextureOracle *tex = nullptr;
  int filter = Filter_LINEAR, wrap = Wrap_CLAMP_TO_EDGE;
  uint32_t border = 0;
};

int levelDim(int d, int l) { return std::max(1, d >> l); }

// BaseSampler::pixelWithWrapMode (SamplerSoft.h:171-213); CoordMod(i,n) expands to i & (2n-1) & (n-1) (:14)
uint32_t pixelWrapped(const Sampler &s, int layer, int level, int x, int y) {
  TextureOracle *t = s.tex;
  int w = levelDim(t->width, level), h = levelDim(t->height, level);
  switch (s.wrap) {
    case Wrap_REPEAT:
      x = (x & (w - 1 + w)) & (w - 1);
      y = (y & (h - 1 + h)) & (h - 1);
      break;
    case Wrap_MIRRORED_REPEAT:
      x = (x & (2 * w - 1 + 2 * w)) & (2 * w - 1);
      y = (y & (2 * h - 1 + 2 * h)) & (2 * h - 1);
      x -= w; y -= h;
      x = x >= 0 ? x : (-1 - x);
      y = y >= 0 ? y : (-1 - y);
      x = w - 1 - x; y = h - 1 - y;
      break;
    case Wrap_CLAMP_TO_EDGE:
      x = std::min(std::max(x, 0), w - 1);
      y = std::min(std::max(y, 0), h - 1);
      break;
    case Wrap_CLAMP_TO_BORDER:
      if (x < 0 || x >= w || y < 0 || y >= h) return s.border;
      break;
  }
  if ((unsigned) x >= (unsigned) w || (unsigned) y >= (unsigned) h) return 0;   // Buffer::get -> nullptr -> T(0)
  return t->levels[layer][level][(size_t) y * w + x];
}

// glm::mix as compiled (objdump of samplePixelBilinear<RGBA> / <float> in the reference binary):
//   u8vec4: fma(y, a, x * (1 - a)), u8 conversion truncates;   float: fma(x, 1 - a, y * a)
uint32_t mixTexel(int format, uint32_t a, uint32_t b, float f) {
  float omf = 1.0f - f;
  if (format == TextureFormat_FLOAT32) return floatToBits(fmaf(bitsToFloat(a), omf, bitsToFloat(b) * f));
  uint32_t r = 0;
  for (int c = 0; c < 4; c++) {
    float x = (float) ((a >> (8 * c)) & 255u), y = (float) ((b >> (8 * c)) & 255u);
    r |= ((uint32_t) (int) fmaf(y, f, x * omf) & 255u) << (8 * c);
  }
  return r;
}

// BaseSampler::samplePixelBilinear (SamplerSoft.h:255-266)
uint32_t pixelBilinear(const Sampler &s, int layer, int level, float u, float v) {
  float tu = u - 0.5f, tv = v - 0.5f;
  float fu = std::floor(tu), fv = std::floor(tv);
  int x = (int) fu, y = (int) fv;
  uint32_t s1 = pixelWrapped(s, layer, level, x, y), s2 = pixelWrapped(s, layer, level, x + 1, y);
  uint32_t s3 = pixelWrapped(s, layer, level, x, y + 1), s4 = pixelWrapped(s, layer, level, x + 1, y + 1);
  float fx = tu - fu, fy = tv - fv;
  int fmt = s.tex->format;
  return mixTexel(fmt, mixTexel(fmt, s1, s2, fx), mixTexel(fmt, s3, s4, fx), fy);
}

uint32_t sampleLevel(const Sampler &s, bool nearest, int layer, int level, float u, float v, int ox, int oy) {
  float w = (float) levelDim(s.tex->width, level), h = (float) levelDim(s.tex->height, level);
  if (nearest)   // sampleNearest (SamplerSoft.h:216-226)
    return pixelWrapped(s, layer, level, (int) std::floor(u * w) + ox, (int) std::floor(v * h) + oy);
  return pixelBilinear(s, layer, level, u * w + (float) ox, v * h + (float) oy);   // sampleBilinear (:229-238)
}

// BaseSampler::textureImpl (SamplerSoft.h:118-168)
uint32_t textureImpl(const Sampler &s, int layer, float u, float v, float lod, int ox = 0, int oy = 0) {
  if (!s.tex || s.tex->levels.empty()) return 0;
  int f = s.filter;
  if (f == Filter_NEAREST) return sampleLevel(s, true, layer, 0, u, v, ox, oy);
  if (f == Filter_LINEAR) return sampleLevel(s, false, layer, 0, u, v, ox, oy);
  int maxLevel = s.tex->levelCount() - 1;
  if (f == Filter_NEAREST_MIPMAP_NEAREST || f == Filter_LINEAR_MIPMAP_NEAREST) {
    int level = std::min(std::max((int) std::ceil(lod + 0.5f) - 1, 0), maxLevel);
    return sampleLevel(s, f == Filter_NEAREST_MIPMAP_NEAREST, layer, level, u, v, ox, oy);
  }
  int hi = std::min(std::max((int) std::floor(lod), 0), maxLevel);
  int lo = std::min(std::max(hi + 1, 0), maxLevel);
  bool nearest = f == Filter_NEAREST_MIPMAP_LINEAR;
  uint32_t thi = sampleLevel(s, nearest, layer, hi, u, v, ox, oy);
  if (hi == lo) return thi;
  uint32_t tlo = sampleLevel(s, nearest, layer, lo, u, v, ox, oy);
  return mixTexel(s.tex->format, thi, tlo, lod - std::floor(lod));
}

// BaseSamplerCube::convertXYZ2UV (SamplerSoft.h:312-373): later matches override earlier ones
void cubeFaceUV(float x, float y, float z, int &face, float &u, float &v) {
  float ax = std::fabs(x), ay = std::fabs(y), az = std::fabs(z);
  bool xp = x > 0, yp = y > 0, zp = z > 0;
  float m = 0, uc = 0, vc = 0;
  face = 0;
  if (xp && ax >= ay && ax >= az) { m = ax; uc = -z; vc = y; face = 0; }
  if (!xp && ax >= ay && ax >= az) { m = ax; uc = z; vc = y; face = 1; }
  if (yp && ay >= ax && ay >= az) { m = ay; uc = x; vc = -z; face = 2; }
  if (!yp && ay >= ax && ay >= az) { m = ay; uc = x; vc = z; face = 3; }
  if (zp && az >= ax && az >= ay) { m = az; uc = x; vc = y; face = 4; }
  if (!zp && az >= ax && az >= ay) { m = az; uc = -x; vc = y; face = 5; }
  vc = -vc;
  u = 0.5f * (uc / m + 1.0f);
  v = 0.5f * (vc / m + 1.0f);
}

f4 rgba(uint32_t p) {   // vec4(u8vec4) / 255.f (ShaderSoft.h:80-111)
  return {(float) (p & 255u) / 255.f, (float) ((p >> 8) & 255u) / 255.f, (float) ((p >> 16) & 255u) / 255.f,
          (float) (p >> 24) / 255.f};
}
f4 tex2D(const Sampler &s, f2 uv, float lod, int ox = 0, int oy = 0) { return rgba(textureImpl(s, 0, uv.x, uv.y, lod, ox, oy)); }
f4 texCube(const Sampler &s, f3 d, float lod) {
  int face;
  float u, v;
  cubeFaceUV(d.x, d.y, d.z, face, u, v);
  if (!s.tex || face >= s.tex->layers()) return {0, 0, 0, 0};
  return rgba(textureImpl(s, face, u, v, lod));
}

// ------------------------------------------------------------------------------------------------ objects
struct VaoOracle : VertexArrayObject {   // VertexArrayObjectSoft (Render/Software/VertexSoft.h:14-46)
  int id = gNextId++;
  std::vector<float> verts;
  std::vector<int32_t> indices;
  int getId() const override { return id; }
  void updateVertexData(void *data, size_t len) override { memcpy(verts.data(), data, std::min(len, verts.size() * 4)); }
};

struct FboOracle : FrameBuffer {
  int id = gNextId++;
  explicit FboOracle(bool off) : FrameBuffer(off) {}
  int getId() const override { return id; }
  bool isValid() override { return colorReady_ || depthReady_; }
};

struct ProgramOracle : ShaderProgram {   // ShaderProgramSoft (Render/Software/ShaderProgramSoft.h:18-128)
  int id = gNextId++;
  int shading = 0;
  std::vector<std::string> defineNames;
  uint32_t defines = 0;
  uint8_t uniforms[512] = {};
  Sampler slots[8];
  int getId() const override { return id; }
  void addDefine(const std::string &d) override { defineNames.push_back(d); }
};

// name tables = getUniformsDesc()/getDefines() of the shader classes under Viewer/Shader/Software
struct Meta {
  const char *blocks[4];
  int offsets[4];
  const char *samplers[8];
  const char *defines[8];
  int uniformBytes, varyings;
};
const Meta *metaOf(int shading) {
  static const Meta basic = {{"UniformsModel", "UniformsMaterial"}, {0, 256}, {}, {}, 304, 0};             // BasicSoft.h:18-61
  static const Meta blinn = {{"UniformsModel", "UniformsScene", "UniformsMaterial"}, {0, 256, 320},        // BlinnPhongSoft.h:14-98
                             {"u_albedoMap", "u_normalMap", "u_emissiveMap", "u_aoMap", "u_shadowMap"},
                             {"ALBEDO_MAP", "NORMAL_MAP", "EMISSIVE_MAP", "AO_MAP"}, 368, 32};
  static const Meta pbr = {{"UniformsModel", "UniformsScene", "UniformsMaterial"}, {0, 256, 320},          // PbrSoft.h:14-105
                           {"u_albedoMap", "u_normalMap", "u_emissiveMap", "u_aoMap", "u_metalRoughnessMap", "u_irradianceMap", "u_prefilterMap"},
                           {"ALBEDO_MAP", "NORMAL_MAP", "EMISSIVE_MAP", "AO_MAP", "METALROUGHNESS_MAP"}, 368, 28};
  static const Meta sky = {{"UniformsModel"}, {0}, {"u_equirectangularMap", "u_cubeMap"}, {"EQUIRECTANGULAR_MAP"}, 256, 4};   // SkyboxSoft.h:14-62
  static const Meta irr = {{"UniformsModel"}, {0}, {"u_cubeMap"}, {}, 256, 4};                            // IBLIrradianceSoft.h
  static const Meta pre = {{"UniformsModel", "UniformsPrefilter"}, {0, 256}, {"u_cubeMap"}, {}, 264, 4};  // IBLPrefilterSoft.h
  static const Meta fxaa = {{"UniformsQuadFilter"}, {0}, {"u_screenTexture"}, {}, 8, 2};                  // FxaaSoft.h:14-53
  switch (shading) {
    case 1: return &basic; case 2: return &blinn; case 3: return &pbr; case 4: return &sky;
    case 5: return &irr; case 6: return &pre; case 7: return &fxaa;
  }
  return nullptr;
}

struct BlockOracle : UniformBlock {   // UniformBlockSoft (Render/Software/UniformSoft.h:17-41)
  std::vector<uint8_t> bytes;
  BlockOracle(const std::string &n, int size) : UniformBlock(n, size), bytes((size_t) size) {}
  int getLocation(ShaderProgram &p) override {
    auto *pr = dynamic_cast<ProgramOracle *>(&p);
    const Meta *m = pr ? metaOf(pr->shading) : nullptr;
    if (!m) return -1;
    for (int i = 0; i < 4 && m->blocks[i]; i++)
      if (name == m->blocks[i]) return m->offsets[i];
    return -1;
  }
  void bindProgram(ShaderProgram &p, int loc) override {
    auto *pr = dynamic_cast<ProgramOracle *>(&p);
    if (pr && loc >= 0 && loc < 512) memcpy(pr->uniforms + loc, bytes.data(), std::min(bytes.size(), (size_t) (512 - loc)));
  }
  void setSubData(void *d, int len, int off) override {
    if (off >= 0 && (size_t) off < bytes.size()) memcpy(bytes.data() + off, d, std::min((size_t) len, bytes.size() - (size_t) off));
  }
  void setData(void *d, int len) override { setSubData(d, len, 0); }
};

struct SamplerUniformOracle : UniformSampler {   // UniformSamplerSoft (UniformSoft.h:43-89)
  Sampler s;
  SamplerUniformOracle(const std::string &n, TextureType t, TextureFormat f) : UniformSampler(n, t, f) {}
  int getLocation(ShaderProgram &p) override {
    auto *pr = dynamic_cast<ProgramOracle *>(&p);
    const Meta *m = pr ? metaOf(pr->shading) : nullptr;
    if (!m) return -1;
    for (int i = 0; i < 8 && m->samplers[i]; i++)
      if (name == m->samplers[i]) return 1000 + i;
    return -1;
  }
  void bindProgram(ShaderProgram &p, int loc) override {
    auto *pr = dynamic_cast<ProgramOracle *>(&p);
    if (pr && loc >= 1000 && loc < 1008) pr->slots[loc - 1000] = s;
  }
  void setTexture(const std::shared_ptr<Texture> &t) override {   // Sampler2DSoft::setTexture (SamplerSoft.h:388-394)
    auto *to = dynamic_cast<TextureOracle *>(t.get());
    if (!to) return;
    s.tex = to;
    s.filter = to->sampler.filterMin;
    s.wrap = to->sampler.wrapS;
    bool white = to->sampler.borderColor == Border_WHITE;   // TextureSoft::getBorderColor (TextureSoft.h:158-164)
    s.border = to->format == TextureFormat_FLOAT32 ? floatToBits(white ? 1.f : 0.f) : (white ? 0xFFFFFFFFu : 0u);
  }
};

// ------------------------------------------------------------------------------------------------ shaders
struct ShaderEnv {
  const ProgramOracle *p;
  const float *q0, *q1, *q2;   // varyings of quad pixels p0,p1,p2 (DerivativeContext, ShaderSoft.h:20-25) or null
};
float uF(const ProgramOracle *p, int off) { float f; memcpy(&f, p->uniforms + off, 4); return f; }
int uI(const ProgramOracle *p, int off) { int i; memcpy(&i, p->uniforms + off, 4); return i; }
f3 uV3(const ProgramOracle *p, int off) { return {uF(p, off), uF(p, off + 4), uF(p, off + 8)}; }
const float *uM(const ProgramOracle *p, int off) { return (const float *) (p->uniforms + off); }

f3 mat3Mul(const float *c0, const float *c1, const float *c2, f3 v) {
  return {fmaf(v.z, c2[0], fmaf(v.x, c0[0], v.y * c1[0])), fmaf(v.z, c2[1], fmaf(v.x, c0[1], v.y * c1[1])),
          fmaf(v.z, c2[2], fmaf(v.x, c0[2], v.y * c1[2]))};
}

// vertex shaders: BasicSoft.h:66-69, BlinnPhongSoft.h:103-121, PbrSoft.h:110-127, SkyboxSoft.h:67-76,
// FxaaSoft.h:58-61, IBLIrradianceSoft.h:62-67, IBLPrefilterSoft.h:69-74
f4 vertexShader(const ProgramOracle *p, const float *a, float *v) {
  f3 pos = {a[0], a[1], a[2]};
  f4 clip = mulPoint(uM(p, 80), pos);
  for (int i = 0; i < 32; i++) v[i] = 0.f;
  switch (p->shading) {
    case 1: return clip;
    case 7: v[0] = a[4]; v[1] = a[5]; return {pos.x, pos.y, pos.z, 1.f};
    case 4: {
      v[0] = pos.x; v[1] = pos.y; v[2] = pos.z;
      f4 r = {clip.x, clip.y, clip.w, clip.w};
      if (uI(p, 0)) r.z = 0.f;
      return r;
    }
    case 5: case 6: v[0] = pos.x; v[1] = pos.y; v[2] = pos.z; return {clip.x, clip.y, clip.w, clip.w};
  }
  const float *model = uM(p, 16);
  f4 wp = mulPoint(model, pos);
  f3 n = {a[8], a[9], a[10]};
  f3 nv = mat3Mul(model, model + 4, model + 8, n);
  v[0] = a[4]; v[1] = a[5];
  v[4] = nv.x; v[5] = nv.y; v[6] = nv.z;
  v[8] = wp.x; v[9] = wp.y; v[10] = wp.z;
  f3 cam = uV3(p, 272), light = uV3(p, 288);
  v[12] = cam.x - wp.x; v[13] = cam.y - wp.y; v[14] = cam.z - wp.z;
  v[16] = light.x - wp.x; v[17] = light.y - wp.y; v[18] = light.z - wp.z;
  int no = 20;
  if (p->shading == 2) {
    f4 sp = mulPoint(uM(p, 192), pos);
    v[20] = sp.x; v[21] = sp.y; v[22] = sp.z; v[23] = sp.w;
    no = 24;
  }
  if (p->defines & 2u) {
    const float *it = uM(p, 144);
    f3 N = norm3(mat3Mul(it, it + 4, it + 8, n));
    f3 T = norm3(mat3Mul(it, it + 4, it + 8, {a[12], a[13], a[14]}));
    f3 T2 = norm3(T - N * dot3(T, N));
    v[no] = N.x; v[no + 1] = N.y; v[no + 2] = N.z;
    v[no + 4] = T2.x; v[no + 5] = T2.y; v[no + 6] = T2.z;
  }
  return clip;
}

// ShaderSoft::getSampler2DLod (Render/Software/ShaderSoft.h:117-133) through BaseSampler2D::texture2DImpl (SamplerSoft.h:73-79)
float implicitLod(const ShaderEnv &e, const Sampler &s) {
  if (s.filter <= Filter_LINEAR || !e.q0 || !s.tex) return 0.f;
  float w = (float) s.tex->width, h = (float) s.tex->height;
  f2 dx = {(e.q1[0] - e.q0[0]) * w, (e.q1[1] - e.q0[1]) * h}, dy = {(e.q2[0] - e.q0[0]) * w, (e.q2[1] - e.q0[1]) * h};
  float d = gmaxf(dx.x * dx.x + dx.y * dx.y, dy.x * dy.x + dy.y * dy.y);
  return gmaxf(0.5f * std::log2(d), 0.0f);
}
f4 texLod(const ShaderEnv &e, int slot, f2 uv) {
  const Sampler &s = e.p->slots[slot];
  return tex2D(s, uv, implicitLod(e, s));
}

f3 normalFromMap(const ShaderEnv &e, const float *v, int no, f2 uv) {   // BlinnPhongSoft.h:149-163 / PbrSoft.h:153-167
  if (e.p->defines & 2u) {
    f3 N = norm3({v[no], v[no + 1], v[no + 2]});
    f3 T = norm3({v[no + 4], v[no + 5], v[no + 6]});
    T = norm3(T - N * dot3(T, N));
    f3 B = cross3(T, N);
    f4 t = texLod(e, 1, uv);
    f3 tn = {t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f};
    return norm3(T * tn.x + B * tn.y + N * tn.z);
  }
  return norm3({v[4], v[5], v[6]});
}

f4 fsBlinnPhong(const ShaderEnv &e, const float *v) {   // BlinnPhongSoft.h:165-242
  const ProgramOracle *p = e.p;
  f2 uv = {v[0], v[1]};
  f4 base = (p->defines & 1u) ? texLod(e, 0, uv) : f4{uF(p, 352), uF(p, 356), uF(p, 360), uF(p, 364)};
  f3 N = normalFromMap(e, v, 24, uv);
  float ao = (p->defines & 8u) ? texLod(e, 3, uv).x : 1.f;
  f3 b3 = {base.x, base.y, base.z};
  f3 ambient = b3 * uV3(p, 256) * ao, diffuse = {0, 0, 0}, specular = {0, 0, 0}, emissive = {0, 0, 0};
  f3 lv = {v[16], v[17], v[18]};
  if (uI(p, 320)) {
    f3 ld = lv * (1.0f / 5.f);
    float att = std::min(std::max(1.0f - dot3(ld, ld), 0.0f), 1.0f);
    f3 L = norm3(lv);
    diffuse = uV3(p, 304) * b3 * std::max(dot3(N, L), 0.0f) * att;
    f3 H = norm3(L + norm3({v[12], v[13], v[14]}));
    float sp = uF(p, 336) * std::pow(std::max(dot3(N, H), 0.0f), 128.f);
    specular = {sp, sp, sp};
    if (uI(p, 328)) {   // ShadowCalculation, 3x3 PCF
      float sh = 0.f;
      f4 fp = {v[20], v[21], v[22], v[23]};
      f3 pc = {fp.x / fp.w, fp.y / fp.w, fp.z / fp.w};
      const Sampler &sm = p->slots[4];
      if (!(pc.z < 0.f || pc.z > 1.f) && sm.tex) {
        float bias = gmaxf(0.00025f * (1.0f - dot3(N, norm3(lv))), 0.00005f);
        f2 po = {1.0f / (float) sm.tex->width, 1.0f / (float) sm.tex->height};
        for (int x = -1; x <= 1; ++x)
          for (int y = -1; y <= 1; ++y) {
            float d = bitsToFloat(textureImpl(sm, 0, pc.x + (float) x * po.x, pc.y + (float) y * po.y, 0.f));
            if (uI(p, 0)) sh += (pc.z + bias < d) ? 1.0f : 0.0f;
            else sh += (pc.z - bias > d) ? 1.0f : 0.0f;
          }
        sh /= 9.0f;
      }
      diffuse = diffuse * (1.0f - sh);
      specular = specular * (1.0f - sh);
    }
  }
  if (p->defines & 4u) { f4 em = texLod(e, 2, uv); emissive = {em.x, em.y, em.z}; }
  f3 c = ambient + diffuse + specular + emissive;
  return {c.x, c.y, c.z, base.w};
}

const float kPi = 3.14159265359f;
float ggxD(f3 N, f3 H, float r) {
  float a = r * r, a2 = a * a, nh = std::max(dot3(N, H), 0.0f);
  float d = nh * nh * (a2 - 1.0f) + 1.0f;
  return a2 / (kPi * d * d);
}
float ggxG1(float nv, float r) { float k = (r + 1.0f) * (r + 1.0f) / 8.0f; return nv / (nv * (1.0f - k) + k); }

f4 fsPbr(const ShaderEnv &e, const float *v) {   // PbrSoft.h:169-322
  const ProgramOracle *p = e.p;
  f2 uv = {v[0], v[1]};
  f4 arga = (p->defines & 1u) ? texLod(e, 0, uv) : f4{uF(p, 352), uF(p, 356), uF(p, 360), uF(p, 364)};
  f3 albedo = {std::pow(arga.x, 2.2f), std::pow(arga.y, 2.2f), std::pow(arga.z, 2.2f)};
  float metallic = 0.f, rough = 1.f;
  if (p->defines & 16u) { f4 mr = tex2D(p->slots[4], uv, 0.f); metallic = mr.z; rough = mr.y; }
  float ao = (p->defines & 8u) ? texLod(e, 3, uv).x : 1.f;
  f3 N = normalFromMap(e, v, 20, uv);
  f3 V = norm3({v[12], v[13], v[14]});
  f3 I = V * -1.f;
  f3 R = I - N * dot3(N, I) * 2.0f;
  f3 F0 = f3{0.04f, 0.04f, 0.04f} * (1.0f - metallic) + albedo * metallic;
  f3 Lo = {0, 0, 0};
  f3 lv = {v[16], v[17], v[18]};
  f3 one = {1.f, 1.f, 1.f};
  if (uI(p, 320)) {
    f3 L = norm3(lv), H = norm3(V + L);
    f3 ld = lv * (1.0f / 5.f);
    float att = std::min(std::max(1.0f - dot3(ld, ld), 0.0f), 1.0f);
    f3 radiance = uV3(p, 304) * att;
    float NDF = ggxD(N, H, rough);
    float G = ggxG1(std::max(dot3(N, L), 0.0f), rough) * ggxG1(std::max(dot3(N, V), 0.0f), rough);
    float p5 = std::pow(std::min(std::max(1.0f - std::max(dot3(H, V), 0.0f), 0.0f), 1.0f), 5.0f);
    f3 F = F0 + (one - F0) * p5;
    f3 spec = F * (NDF * G) / (4.0f * std::max(dot3(N, V), 0.0f) * std::max(dot3(N, L), 0.0f) + 0.0001f);
    f3 kD = (one - F) * (1.0f - metallic);
    Lo = Lo + (kD * albedo / kPi + spec) * radiance * std::max(dot3(N, L), 0.0f);
  }
  f3 ambient;
  if (uI(p, 324)) {
    float nv = std::max(dot3(N, V), 0.0f);
    float p5 = std::pow(std::min(std::max(1.0f - nv, 0.0f), 1.0f), 5.0f);
    float omr = 1.0f - rough;
    f3 Fm = {std::max(omr, F0.x), std::max(omr, F0.y), std::max(omr, F0.z)};
    f3 F = F0 + (Fm - F0) * p5;
    f3 kD = (one - F) * (1.0f - metallic);
    f4 irr = texCube(p->slots[5], N, 0.f);
    f4 pre = texCube(p->slots[6], R, rough * 4.0f);
    // EnvBRDFApprox (PbrSoft.h:208-223)
    f4 r = {rough * -1.f + 1.f, rough * -0.0275f + 0.0425f, rough * -0.572f + 1.04f, rough * 0.022f + -0.04f};
    float a004 = std::min(r.x * r.x, std::exp2(-9.28f * nv)) * r.x + r.y;
    float ABx = -1.04f * a004 + r.z, ABy = (1.04f * a004 + r.w) * std::max(0.f, std::min(1.f, 50.0f * F.y));
    f3 env = F * ABx + f3{ABy, ABy, ABy};
    ambient = (kD * (f3{irr.x, irr.y, irr.z} * albedo) + f3{pre.x, pre.y, pre.z} * env) * ao;
  } else {
    ambient = uV3(p, 256) * albedo * ao;
  }
  f3 c = ambient + Lo;
  c = {std::pow(c.x, 1.0f / 2.2f), std::pow(c.y, 1.0f / 2.2f), std::pow(c.z, 1.0f / 2.2f)};
  if (p->defines & 4u) { f4 em = texLod(e, 2, uv); c = c + f3{em.x, em.y, em.z}; }
  return {c.x, c.y, c.z, arga.w};
}

f4 fsSkybox(const ShaderEnv &e, const float *v) {   // SkyboxSoft.h:79-98
  f3 wp = {v[0], v[1], v[2]};
  if (e.p->defines & 1u) {
    f3 d = norm3(wp);
    f2 uv = {std::atan2(d.z, d.x) * 0.1591f + 0.5f, std::asin(-d.y) * 0.3183f + 0.5f};
    return tex2D(e.p->slots[0], uv, 0.f);
  }
  return texCube(e.p->slots[1], wp, 0.f);
}

float luma(f4 c) { return c.x * 0.299f + c.y * 0.587f + c.z * 0.114f; }
float fxaaQuality(int i) { float q = (float) i; return q < 5.f ? 1.0f : (q > 5.f ? (q < 10.f ? 2.0f : (q < 11.f ? 4.0f : 8.0f)) : 1.5f); }

f4 fsFxaa(const ShaderEnv &e, const float *v) {   // FxaaSoft.h:63-266
  const Sampler &s = e.p->slots[0];
  f2 uv = {v[0], v[1]}, inv = {1.0f / uF(e.p, 0), 1.0f / uF(e.p, 4)};
  f4 cc = tex2D(s, uv, 0.f);
  float lc = luma(cc), ld = luma(tex2D(s, uv, 0.f, 0, -1)), lu = luma(tex2D(s, uv, 0.f, 0, 1));
  float ll = luma(tex2D(s, uv, 0.f, -1, 0)), lr = luma(tex2D(s, uv, 0.f, 1, 0));
  float lmin = std::min(lc, std::min(std::min(ld, lu), std::min(ll, lr)));
  float lmax = std::max(lc, std::max(std::max(ld, lu), std::max(ll, lr)));
  float range = lmax - lmin;
  if (range < std::max(0.0312f, lmax * 0.125f)) return {cc.x, cc.y, cc.z, 1.f};
  float ldl = luma(tex2D(s, uv, 0.f, -1, -1)), lur = luma(tex2D(s, uv, 0.f, 1, 1));
  float lul = luma(tex2D(s, uv, 0.f, -1, 1)), ldr = luma(tex2D(s, uv, 0.f, 1, -1));
  float ldu = ld + lu, llr = ll + lr, lcl = ldl + lul, lcd = ldl + ldr, lcr = ldr + lur, lcu = lur + lul;
  float eh = std::fabs(-2.0f * ll + lcl) + std::fabs(-2.0f * lc + ldu) * 2.0f + std::fabs(-2.0f * lr + lcr);
  float ev = std::fabs(-2.0f * lu + lcu) + std::fabs(-2.0f * lc + llr) * 2.0f + std::fabs(-2.0f * ld + lcd);
  bool horiz = eh >= ev;
  float step = horiz ? inv.y : inv.x;
  float l1 = horiz ? ld : ll, l2 = horiz ? lu : lr;
  float g1 = l1 - lc, g2 = l2 - lc;
  bool steep1 = std::fabs(g1) >= std::fabs(g2);
  float gs = 0.25f * std::max(std::fabs(g1), std::fabs(g2));
  float avg;
  if (steep1) { step = -step; avg = 0.5f * (l1 + lc); } else avg = 0.5f * (l2 + lc);
  f2 cur = uv;
  if (horiz) cur.y += step * 0.5f; else cur.x += step * 0.5f;
  f2 off = horiz ? f2{inv.x, 0.f} : f2{0.f, inv.y};
  float q0 = fxaaQuality(0);
  f2 a = {cur.x - off.x * q0, cur.y - off.y * q0}, b = {cur.x + off.x * q0, cur.y + off.y * q0};
  float e1 = 0.f, e2 = 0.f;
  bool r1 = false, r2 = false;
  for (int i = 1; i < 12; i++) {
    if (!r1) { e1 = luma(tex2D(s, a, 0.f)) - avg; r1 = std::fabs(e1) >= gs; }
    if (!r2) { e2 = luma(tex2D(s, b, 0.f)) - avg; r2 = std::fabs(e2) >= gs; }
    float q = fxaaQuality(i);
    if (!r1) { a.x -= off.x * q; a.y -= off.y * q; }
    if (!r2) { b.x += off.x * q; b.y += off.y * q; }
    if (r1 && r2) break;
  }
  float d1 = horiz ? (uv.x - a.x) : (uv.y - a.y), d2 = horiz ? (b.x - uv.x) : (b.y - uv.y);
  bool dir1 = d1 < d2;
  float dmin = std::min(d1, d2);
  bool smaller = lc < avg;
  bool ok = dir1 ? ((e1 < 0.0f) != smaller) : ((e2 < 0.0f) != smaller);
  float pixOff = -dmin / (d1 + d2) + 0.5f;
  float fin = ok ? pixOff : 0.0f;
  float lavg = (1.0f / 12.0f) * (2.0f * (ldu + llr) + lcl + lcr);
  float s1 = std::min(std::max(std::fabs(lavg - lc) / range, 0.0f), 1.0f);
  float s2 = (-2.0f * s1 + 3.0f) * s1 * s1;
  fin = std::max(fin, s2 * s2 * 0.75f);
  f2 fuv = uv;
  if (horiz) fuv.y += fin * step; else fuv.x += fin * step;
  f4 fc = tex2D(s, fuv, 0.f);
  return {fc.x, fc.y, fc.z, 1.f};
}

f4 fsIrradiance(const ShaderEnv &e, const float *v) {   // IBLIrradianceSoft.h:74-104
  const Sampler &s = e.p->slots[0];
  f3 N = norm3({v[0], v[1], v[2]}), irr = {0, 0, 0}, up = {0.f, 1.f, 0.f};
  f3 right = norm3(cross3(up, N));
  up = norm3(cross3(N, right));
  float n = 0.f;
  for (float phi = 0.0f; phi < 2.0f * kPi; phi += 0.025f)
    for (float th = 0.0f; th < 0.5f * kPi; th += 0.025f) {
      f3 ts = {std::sin(th) * std::cos(phi), std::sin(th) * std::sin(phi), std::cos(th)};
      f3 sv = right * ts.x + up * ts.y + N * ts.z;
      f4 t = texCube(s, sv, 0.f);
      irr = irr + f3{t.x, t.y, t.z} * std::cos(th) * std::sin(th);
      n += 1.0f;
    }
  irr = irr * kPi * (1.0f / n);
  return {irr.x, irr.y, irr.z, 1.0f};
}

f4 fsPrefilter(const ShaderEnv &e, const float *v) {   // IBLPrefilterSoft.h:76-168
  const Sampler &s = e.p->slots[0];
  float res = uF(e.p, 256), rough = uF(e.p, 260), a = rough * rough;
  f3 N = norm3({v[0], v[1], v[2]}), V = N, col = {0, 0, 0};
  float tw = 0.f;
  f3 upv = std::fabs(N.z) < 0.999f ? f3{0, 0, 1} : f3{1, 0, 0};
  f3 tg = norm3(cross3(upv, N)), bt = cross3(N, tg);
  for (uint32_t i = 0; i < 1024u; ++i) {
    uint32_t b = (i << 16u) | (i >> 16u);
    b = ((b & 0x55555555u) << 1u) | ((b & 0xAAAAAAAAu) >> 1u);
    b = ((b & 0x33333333u) << 2u) | ((b & 0xCCCCCCCCu) >> 2u);
    b = ((b & 0x0F0F0F0Fu) << 4u) | ((b & 0xF0F0F0F0u) >> 4u);
    b = ((b & 0x00FF00FFu) << 8u) | ((b & 0xFF00FF00u) >> 8u);
    f2 Xi = {(float) i / 1024.f, (float) ((double) b * 2.3283064365386963e-10)};
    float phi = 2.0f * kPi * Xi.x;
    float ct = std::sqrt((1.0f - Xi.y) / (1.0f + (a * a - 1.0f) * Xi.y)), st = std::sqrt(1.0f - ct * ct);
    f3 H = norm3(tg * (std::cos(phi) * st) + bt * (std::sin(phi) * st) + N * ct);
    f3 L = norm3(H * (2.0f * dot3(V, H)) - V);
    float nl = std::max(dot3(N, L), 0.0f);
    if (nl > 0.0f) {
      float D = ggxD(N, H, rough), nh = std::max(dot3(N, H), 0.0f), hv = std::max(dot3(H, V), 0.0f);
      float pdf = D * nh / (4.0f * hv) + 0.0001f;
      float saTexel = 4.0f * kPi / (6.0f * res * res), saSample = 1.0f / (1024.f * pdf + 0.0001f);
      float mip = rough == 0.0f ? 0.0f : 0.5f * std::log2(saSample / saTexel);
      f4 t = texCube(s, L, mip);
      col = col + f3{t.x, t.y, t.z} * nl;
      tw += nl;
    }
  }
  col = col / tw;
  return {col.x, col.y, col.z, 1.0f};
}

f4 fragmentShader(const ShaderEnv &e, const float *v) {
  switch (e.p->shading) {
    case 1: return {uF(e.p, 288), uF(e.p, 292), uF(e.p, 296), uF(e.p, 300)};   // BasicSoft.h:76-78
    case 2: return fsBlinnPhong(e, v);
    case 3: return fsPbr(e, v);
    case 4: return fsSkybox(e, v);
    case 5: return fsIrradiance(e, v);
    case 6: return fsPrefilter(e, v);
    case 7: return fsFxaa(e, v);
  }
  return {0, 0, 0, 0};
}

bool needsDeriv(const ProgramOracle *p) {
  if (p->shading != 2 && p->shading != 3) return false;
  for (int s = 0; s < 4; s++)
    if (((p->defines >> s) & 1u) && p->slots[s].tex && p->slots[s].filter > Filter_LINEAR) return true;
  return false;
}

// ------------------------------------------------------------------------------------------------ renderer
struct VertexO {   // VertexHolder (Render/Software/RendererInternal.h:30-43)
  float attr[16];
  float vary[32];
  f4 clip, frag;
  int mask;
};
struct PrimO { int i[3]; bool discard, front; };

// BlendSoft.h:14-56
float blendFactor(float s, float sa, float d, float da, int f) {
  switch (f) {
    case 0: return 0.f; case 1: return 1.f; case 2: return s; case 3: return sa; case 4: return d; case 5: return da;
    case 6: return 1.f - s; case 7: return 1.f - sa; case 8: return 1.f - d; case 9: return 1.f - da;
  }
  return 0.f;
}
float blendFunc(float s, float d, int fn) {
  switch (fn) { case 0: return s + d; case 1: return s - d; case 2: return d - s; case 3: return gminf(s, d); case 4: return gmaxf(s, d); }
  return s + d;
}
// DepthSoft.h:13-25
bool depthTestFn(float a, float b, int fn) {
  switch (fn) {
    case 0: return false; case 1: return a < b; case 2: return std::fabs(a - b) <= FLT_EPSILON; case 3: return a <= b;
    case 4: return a > b; case 5: return std::fabs(a - b) > FLT_EPSILON; case 6: return a >= b; case 7: return true;
  }
  return a < b;
}

// barycentric (RendererSoft.cpp:1021-1056), SIMD association
bool barycentricO(const f4 *v, float px, float py, float *bc) {
  float ax = v[2].x - v[0].x, ay = v[1].x - v[0].x, az = v[0].x - px;
  float bx = v[2].y - v[0].y, by = v[1].y - v[0].y, bz = v[0].y - py;
  float ux = fmaf(ay, bz, -(az * by)), uy = fmaf(az, bx, -(ax * bz)), uz = fmaf(ax, by, -(ay * bx));
  if (std::fabs(uz) < FLT_EPSILON) return false;
  ux = ux / uz; uy = uy / uz;
  bc[0] = 1.f - (ux + uy); bc[1] = uy; bc[2] = ux;
  return !(bc[0] < 0 || bc[1] < 0 || bc[2] < 0);
}

class RendererOracle : public Renderer {
 public:
  RendererType type() override { return Renderer_SOFT; }
  std::shared_ptr<FrameBuffer> createFrameBuffer(bool off) override { return std::make_shared<FboOracle>(off); }
  std::shared_ptr<Texture> createTexture(const TextureDesc &d) override { return std::make_shared<TextureOracle>(d); }
  std::shared_ptr<VertexArrayObject> createVertexArrayObject(const VertexArray &va) override {
    auto v = std::make_shared<VaoOracle>();
    v->verts.resize(va.vertexesBufferLength / 4);
    memcpy(v->verts.data(), va.vertexesBuffer, va.vertexesBufferLength);
    v->indices.resize(va.indexBufferLength / 4);
    memcpy(v->indices.data(), va.indexBuffer, va.indexBufferLength);
    return v;
  }
  std::shared_ptr<ShaderProgram> createShaderProgram() override { return std::make_shared<ProgramOracle>(); }
  std::shared_ptr<PipelineStates> createPipelineStates(const RenderStates &rs) override { return std::make_shared<PipelineStates>(rs); }
  std::shared_ptr<UniformBlock> createUniformBlock(const std::string &n, int size) override { return std::make_shared<BlockOracle>(n, size); }
  std::shared_ptr<UniformSampler> createUniformSampler(const std::string &n, const TextureDesc &d) override {
    return std::make_shared<SamplerUniformOracle>(n, d.type, d.format);
  }

  // RendererSoft::beginRenderPass (Render/Software/RendererSoft.cpp:61-90)
  void beginRenderPass(std::shared_ptr<FrameBuffer> &fb, const ClearStates &cs) override {
    fbo_ = dynamic_cast<FboOracle *>(fb.get());
    if (!fbo_) return;
    resolveAttachments();
    if (cs.colorFlag && color_) {
      uint32_t c = 0;
      float ch[4] = {cs.clearColor.r, cs.clearColor.g, cs.clearColor.b, cs.clearColor.a};
      for (int k = 0; k < 4; k++) c |= ((uint32_t) (uint8_t) (int) (ch[k] * 255.f)) << (8 * k);
      std::fill(color_->begin(), color_->end(), c);
    }
    if (cs.depthFlag && depth_) std::fill(depth_->begin(), depth_->end(), floatToBits(cs.clearDepth));
  }
  void setViewPort(int x, int y, int w, int h) override { vpX_ = (float) x; vpY_ = (float) y; vpW_ = (float) w; vpH_ = (float) h; }
  void setVertexArrayObject(std::shared_ptr<VertexArrayObject> &v) override { vao_ = dynamic_cast<VaoOracle *>(v.get()); }
  void setShaderProgram(std::shared_ptr<ShaderProgram> &p) override { prog_ = dynamic_cast<ProgramOracle *>(p.get()); }
  void setShaderResources(std::shared_ptr<ShaderResources> &r) override { if (r && prog_) prog_->bindResources(*r); }
  void setPipelineStates(std::shared_ptr<PipelineStates> &s) override { rs_ = &s->renderStates; }
  void endRenderPass() override {}
  void waitIdle() override {}

  // RendererSoft::draw (RendererSoft.cpp:136-164)
  void draw() override {
    if (!fbo_ || !vao_ || !prog_ || !rs_) return;
    resolveAttachments();
    ns_ = colorTex_ ? colorTex_->samples() : (depthTex_ ? depthTex_->samples() : 1);
    const Meta *m = metaOf(prog_->shading);
    if (!m) return;
    nvary_ = m->varyings;
    // processVertexShader (:170-190): every VAO vertex
    size_t nv = vao_->verts.size() / 16;
    verts_.assign(nv, VertexO());
    pointSize_ = 1.f;
    for (size_t i = 0; i < nv; i++) {
      memcpy(verts_[i].attr, &vao_->verts[i * 16], 64);
      shadeVertex(verts_[i]);
    }
    // processPrimitiveAssembly (:192-204,409-437)
    int per = rs_->primitiveType == Primitive_TRIANGLE ? 3 : (rs_->primitiveType == Primitive_LINE ? 2 : 1);
    prims_.clear();
    for (size_t i = 0; i + per <= vao_->indices.size(); i += per) {
      PrimO p{};
      for (int k = 0; k < per; k++) p.i[k] = vao_->indices[i + k];
      p.front = true;
      prims_.push_back(p);
    }
    // processClipping (:206-257)
    size_t cnt = prims_.size();
    for (size_t i = 0; i < cnt; i++) {
      if (rs_->primitiveType == Primitive_POINT) prims_[i].discard = verts_[prims_[i].i[0]].mask != 0;
      else if (rs_->primitiveType == Primitive_LINE) clipLine(prims_[i], false);
      else if (rs_->polygonMode == PolygonMode_FILL) clipTriangle(i);
    }
    // processPerspectiveDivide / processViewportTransform (:259-275): new vertices made later do their own
    for (auto &v : verts_) toScreen(v);
    // processFaceCulling (:277-299)
    if (rs_->primitiveType == Primitive_TRIANGLE)
      for (auto &t : prims_) {
        if (t.discard) continue;
        f4 a = verts_[t.i[0]].frag, b = verts_[t.i[1]].frag, c = verts_[t.i[2]].frag;
        float ax = b.x - a.x, ay = b.y - a.y, bx = c.x - a.x, by = c.y - a.y;
        t.front = (fmaf(ax, by, -(ay * bx)) + 0.f) > 0.f;
        if (rs_->cullFace) t.discard = !t.front;
      }
    // processRasterization (:301-340)
    for (size_t i = 0; i < prims_.size(); i++) {
      PrimO t = prims_[i];
      if (t.discard) continue;
      if (rs_->primitiveType == Primitive_POINT) rasterPoint(verts_[t.i[0]].frag, verts_[t.i[0]].vary, pointSize_);
      else if (rs_->primitiveType == Primitive_LINE) rasterLine(t.i[0], t.i[1]);
      else if (rs_->polygonMode == PolygonMode_FILL) rasterTriangle(t);
      else if (rs_->polygonMode == PolygonMode_LINE) {   // rasterizationPolygonsLine (:598-622)
        for (int e = 0; e < 3; e++) {
          PrimO l{};
          l.i[0] = t.i[e]; l.i[1] = t.i[(e + 1) % 3];
          clipLine(l, true);
          if (!l.discard) rasterLine(l.i[0], l.i[1]);
        }
      } else {                                           // rasterizationPolygonsPoint (:575-596)
        for (int e = 0; e < 3; e++)
          if (verts_[t.i[e]].mask == 0) rasterPoint(verts_[t.i[e]].frag, verts_[t.i[e]].vary, pointSize_);
      }
    }
    if (colorTex_ && colorTex_->multiSample) resolve();   // multiSampleResolve (:880-912)
  }

 private:
  void resolveAttachments() {   // FrameBufferSoft::getColorBuffer/getDepthBuffer (FramebufferSoft.h:27-41)
    colorTex_ = depthTex_ = nullptr;
    color_ = depth_ = nullptr;
    if (fbo_->isColorReady()) {
      colorTex_ = dynamic_cast<TextureOracle *>(fbo_->getColorAttachment().tex.get());
      if (colorTex_) {
        if (colorTex_->levels.empty()) colorTex_->initImageData();
        int lv = (int) fbo_->getColorAttachment().level;
        color_ = &colorTex_->levels[fbo_->getColorAttachment().layer][lv];
        fbW_ = levelDim(colorTex_->width, lv); fbH_ = levelDim(colorTex_->height, lv);
      }
    }
    if (fbo_->isDepthReady()) {
      depthTex_ = dynamic_cast<TextureOracle *>(fbo_->getDepthAttachment().tex.get());
      if (depthTex_) {
        if (depthTex_->levels.empty()) depthTex_->initImageData();
        depth_ = &depthTex_->levels[0][0];
        if (!colorTex_) { fbW_ = depthTex_->width; fbH_ = depthTex_->height; }
      }
    }
  }

  void shadeVertex(VertexO &v) {   // vertexShaderImpl (:971-979) + countFrustumClipMask (:994-1003)
    v.clip = vertexShader(prog_, v.attr, v.vary);
    if (prog_->shading == 1) pointSize_ = uF(prog_, 268);
    const f4 &c = v.clip;
    v.mask = (c.w < c.x ? 1 : 0) | (c.w < -c.x ? 2 : 0) | (c.w < c.y ? 4 : 0) | (c.w < -c.y ? 8 : 0) |
             (c.w < c.z ? 16 : 0) | (c.w < -c.z ? 32 : 0);
  }
  void toScreen(VertexO &v) {   // perspectiveDivideImpl + viewportTransformImpl (:981-992), setViewPort (:92-113)
    float inv = 1.f / v.clip.w;
    f4 p = {v.clip.x * inv, v.clip.y * inv, v.clip.z * inv, inv};
    v.frag = {p.x * (vpW_ / 2.f) + (vpX_ + vpW_ / 2.f), p.y * (vpH_ / 2.f) + (vpY_ + vpH_ / 2.f), p.z * 1.f + 0.f, p.w * 1.f + 0.f};
  }
  int newVertex(int i0, int i1, float t, bool post) {   // clippingNewVertex + interpolateVertex (:956-969,1058-1070)
    VertexO n{};
    float omt = 1.f - t;
    for (int k = 0; k < 16; k++) n.attr[k] = fmaf(verts_[i0].attr[k], omt, verts_[i1].attr[k] * t);
    shadeVertex(n);
    if (post) toScreen(n);
    verts_.push_back(n);
    return (int) verts_.size() - 1;
  }
  float planeDist(int plane, const f4 &c) {   // Geometry.h:117-133
    static const float P[6][4] = {{-1, 0, 0, 1}, {1, 0, 0, 1}, {0, -1, 0, 1}, {0, 1, 0, 1}, {0, 0, -1, 1}, {0, 0, 1, 1}};
    return dpps(P[plane], &c.x);
  }
  void clipTriangle(size_t ti) {   // clippingTriangle (:489-559)
    int mask = verts_[prims_[ti].i[0]].mask | verts_[prims_[ti].i[1]].mask | verts_[prims_[ti].i[2]].mask;
    if (!mask) return;
    std::vector<int> in = {prims_[ti].i[0], prims_[ti].i[1], prims_[ti].i[2]}, out;
    bool full = false;
    for (int pl = 0; pl < 6; pl++) {
      if (!(mask & (1 << pl))) continue;
      if (in.size() < 3) { full = true; break; }
      out.clear();
      int pre = in[0];
      float dPre = planeDist(pl, verts_[pre].clip);
      in.push_back(pre);
      for (size_t k = 1; k < in.size(); k++) {
        int idx = in[k];
        float d = planeDist(pl, verts_[idx].clip);
        if (dPre >= 0) out.push_back(pre);
        if (std::signbit(dPre) != std::signbit(d)) {
          float t = d < 0 ? dPre / (dPre - d) : -dPre / (d - dPre);
          out.push_back(newVertex(pre, idx, t, false));
        }
        pre = idx;
        dPre = d;
      }
      in.swap(out);
    }
    if (full || in.size() < 3) { prims_[ti].discard = true; return; }
    prims_[ti].i[0] = in[0]; prims_[ti].i[1] = in[1]; prims_[ti].i[2] = in[2];
    for (size_t k = 3; k < in.size(); k++) {
      PrimO p{};
      p.i[0] = in[0]; p.i[1] = in[k - 1]; p.i[2] = in[k];
      p.front = prims_[ti].front;
      prims_.push_back(p);   // appended after all originals (:227)
    }
  }
  void clipLine(PrimO &l, bool post) {   // clippingLine (:443-487)
    int m0 = verts_[l.i[0]].mask, m1 = verts_[l.i[1]].mask;
    f4 c0 = verts_[l.i[0]].clip, c1 = verts_[l.i[1]].clip;
    float t0 = 0.f, t1 = 1.f;
    int mask = m0 | m1;
    for (int pl = 0; pl < 6 && mask; pl++) {
      if (!(mask & (1 << pl))) continue;
      float d0 = planeDist(pl, c0), d1 = planeDist(pl, c1);
      if (d0 < 0 && d1 < 0) { l.discard = true; return; }
      if (d0 < 0) t0 = gmaxf(t0, -d0 / (d1 - d0));
      else t1 = gminf(t1, d0 / (d0 - d1));
    }
    if (m0) l.i[0] = newVertex(l.i[0], l.i[1], t0, post);
    if (m1) l.i[1] = newVertex(l.i[0], l.i[1], t1, post);
  }

  uint32_t *colorAt(int x, int y, int s) {
    if (!color_ || (unsigned) x >= (unsigned) fbW_ || (unsigned) y >= (unsigned) fbH_) return nullptr;
    return &(*color_)[((size_t) y * fbW_ + x) * ns_ + s];
  }
  uint32_t *depthAt(int x, int y, int s) {
    if (!depth_ || (unsigned) x >= (unsigned) fbW_ || (unsigned) y >= (unsigned) fbH_) return nullptr;
    return &(*depth_)[((size_t) y * fbW_ + x) * ns_ + s];
  }
  bool depthTest(int x, int y, float z, int s, bool skipWrite) {   // processDepthTest (:377-395)
    if (!rs_->depthTest || !depth_) return true;
    z = gclampf(z, 0.f, 1.f);
    uint32_t *p = depthAt(x, y, s);
    if (p && depthTestFn(z, bitsToFloat(*p), rs_->depthFunc)) {
      if (!skipWrite && rs_->depthMask) *p = floatToBits(z);
      return true;
    }
    return false;
  }
  void perSample(int x, int y, float z, f4 c, int s) {   // processPerSampleOperations + blending (:358-407)
    if (!depthTest(x, y, z, s, false)) return;
    if (!color_) return;
    c = {gclampf(c.x, 0.f, 1.f), gclampf(c.y, 0.f, 1.f), gclampf(c.z, 0.f, 1.f), gclampf(c.w, 0.f, 1.f)};
    uint32_t *dst = colorAt(x, y, s);
    if (rs_->blend) {
      f4 d = dst ? rgba(*dst) : f4{0, 0, 0, 0};
      const BlendParameters &bp = rs_->blendParams;
      float sc[3] = {c.x, c.y, c.z}, dc[3] = {d.x, d.y, d.z}, o[3];
      for (int k = 0; k < 3; k++)
        o[k] = blendFunc(sc[k] * blendFactor(sc[k], c.w, dc[k], d.w, bp.blendSrcRgb), dc[k] * blendFactor(sc[k], c.w, dc[k], d.w, bp.blendDstRgb), bp.blendFuncRgb);
      float oa = blendFunc(c.w * blendFactor(c.w, c.w, d.w, d.w, bp.blendSrcAlpha), d.w * blendFactor(c.w, c.w, d.w, d.w, bp.blendDstAlpha), bp.blendFuncAlpha);
      c = {o[0], o[1], o[2], oa};
    }
    if (dst) {   // setFrameColor(color * 255.f) -> u8 truncation (:949-954)
      float ch[4] = {c.x * 255.f, c.y * 255.f, c.z * 255.f, c.w * 255.f};
      uint32_t pk = 0;
      for (int k = 0; k < 4; k++) pk |= ((uint32_t) (int) ch[k] & 255u) << (8 * k);
      *dst = pk;
    }
  }

  void rasterPoint(f4 &fp, const float *vary, float size) {   // rasterizationPoint (:636-661)
    if (!color_) return;
    float left = fp.x - size / 2.f + 0.5f, right = left + size, top = fp.y - size / 2.f + 0.5f, bottom = top + size;
    ShaderEnv env = {prog_, nullptr, nullptr, nullptr};
    for (int x = (int) left; x < (int) right; x++)
      for (int y = (int) top; y < (int) bottom; y++) {
        f4 c = fragmentShader(env, vary);
        for (int s = 0; s < ns_; s++) perSample(x, y, fp.z, c, s);
      }
  }
  void rasterLine(int ia, int ib) {   // rasterizationLine (:663-718)
    const VertexO &A = verts_[ia], &B = verts_[ib];
    int x0 = (int) A.frag.x, y0 = (int) A.frag.y, x1 = (int) B.frag.x, y1 = (int) B.frag.y;
    float z0 = A.frag.z, z1 = B.frag.z, w0 = A.frag.w, w1 = B.frag.w;
    const float *va = A.vary, *vb = B.vary;
    bool steep = false;
    if (std::abs(x0 - x1) < std::abs(y0 - y1)) { std::swap(x0, y0); std::swap(x1, y1); steep = true; }
    if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); std::swap(z0, z1); std::swap(w0, w1); std::swap(va, vb); }
    int dx = x1 - x0, dy = y1 - y0, err = 0, dErr = 2 * std::abs(dy), y = y0;
    float vary[32];
    for (int x = x0; x <= x1; x++) {
      float t = (float) (x - x0) / (float) dx, omt = 1.f - t;
      f4 fp = {(float) x, (float) y, fmaf(z0, omt, z1 * t), fmaf(w0, omt, w1 * t)};
      if (steep) std::swap(fp.x, fp.y);
      for (int k = 0; k < nvary_; k++) vary[k] = fmaf(va[k], omt, vb[k] * t);
      rasterPoint(fp, vary, rs_->lineWidth);
      err += dErr;
      if (err > dx) { y += (y1 > y0 ? 1 : -1); err -= 2 * dx; }
    }
  }

  // ---- triangles: rasterizationTriangle + rasterizationPixelQuad (:720-851) ----
  struct SampleO { bool inside; int fx, fy; float x, y, z, w; float bc[3]; };
  struct PixelO { bool inside; SampleO s[5]; int shade; int count; int coverage; float vary[32]; };

  bool barycentric(const f4 *v, float px, float py, float *bc) { return barycentricO(v, px, py, bc); }

  void rasterTriangle(const PrimO &t) {
    const VertexO *vx[3] = {&verts_[t.i[0]], &verts_[t.i[1]], &verts_[t.i[2]]};
    f4 v[3] = {vx[0]->frag, vx[1]->frag, vx[2]->frag};
    // triangleBoundingBox (:1005-1019) then bounds.min -= 1 (:725)
    float minX = gmaxf(gminf(gminf(v[0].x, v[1].x), v[2].x) - 0.5f, 0.f), minY = gmaxf(gminf(gminf(v[0].y, v[1].y), v[2].y) - 0.5f, 0.f);
    float maxX = gminf(gmaxf(gmaxf(v[0].x, v[1].x), v[2].x) + 0.5f, vpW_ - 1.f), maxY = gminf(gmaxf(gmaxf(v[0].y, v[1].y), v[2].y) + 0.5f, vpH_ - 1.f);
    minX -= 1.f; minY -= 1.f;
    const int bs = 32;
    int cx = (int) ((maxX - minX + (float) bs - 1.f) / (float) bs), cy = (int) ((maxY - minY + (float) bs - 1.f) / (float) bs);
    for (int by = 0; by < cy; by++)
      for (int bx = 0; bx < cx; bx++) {
        int sx = (int) (minX + (float) (bx * bs)), sy = (int) (minY + (float) (by * bs));
        for (int y = sy + 1; y < sy + bs && (float) y <= maxY; y += 2)
          for (int x = sx + 1; x < sx + bs && (float) x <= maxX; x += 2) pixelQuad(v, vx, x, y);
      }
  }

  void pixelQuad(const f4 *v, const VertexO *const *vx, int qx, int qy) {
    PixelO px[4];
    static const float loc[4][2] = {{0.375f, 0.875f}, {0.875f, 0.625f}, {0.125f, 0.375f}, {0.625f, 0.125f}};   // RendererInternal.h:61-69
    bool any = false;
    for (int k = 0; k < 4; k++) {
      PixelO &p = px[k];
      int x = qx + (k & 1), y = qy + (k >> 1);   // p0 (x,y) p1 (x+1,y) p2 (x,y+1) p3 (x+1,y+1)  (:147-164)
      p.count = ns_ > 1 ? ns_ + 1 : 1;
      for (int s = 0; s < p.count; s++) {
        SampleO &sm = p.s[s];
        sm.fx = x; sm.fy = y;
        if (ns_ > 1 && s < 4) { sm.x = loc[s][0] + (float) x; sm.y = loc[s][1] + (float) y; }
        else { sm.x = (float) x + 0.5f; sm.y = (float) y + 0.5f; }
        sm.z = sm.w = 0.f;
        sm.bc[0] = sm.bc[1] = sm.bc[2] = 0.f;
        sm.inside = barycentric(v, sm.x, sm.y, sm.bc);
      }
      // InitCoverage + InitShadingSample (RendererInternal.h:103-124)
      p.shade = p.count - 1;
      if (ns_ > 1) {
        p.coverage = 0;
        for (int s = 0; s < 4; s++) p.coverage += p.s[s].inside ? 1 : 0;
        p.inside = p.coverage > 0;
      } else {
        p.coverage = 1;
        p.inside = p.s[0].inside;
      }
      if (!p.s[p.shade].inside)
        for (int s = 0; s < p.count; s++)
          if (p.s[s].inside) { p.shade = s; break; }
      any = any || p.inside;
    }
    if (!any) return;
    for (auto &p : px)
      for (int s = 0; s < p.count; s++) {
        SampleO &sm = p.s[s];
        if (!sm.inside) continue;
        float bz[4] = {sm.bc[0], sm.bc[1], sm.bc[2], 0.f};
        float zs[4] = {v[0].z, v[1].z, v[2].z, 0.f}, ws[4] = {v[0].w, v[1].w, v[2].w, 0.f};
        sm.z = dpps(bz, zs);   // interpolateBarycentric(&position.z, vertZ, 2, bc) -> glm::dot (:794,1110-1112)
        sm.w = dpps(bz, ws);
        if (sm.z < 0.f || sm.z > 1.f) sm.inside = false;   // depth clipping (:797)
        float inv = 1.f / sm.w;                            // barycentric correction (:802)
        sm.bc[0] = (inv * v[0].w) * sm.bc[0];
        sm.bc[1] = (inv * v[1].w) * sm.bc[1];
        sm.bc[2] = (inv * v[2].w) * sm.bc[2];
      }
    if (rs_->depthTest) {   // earlyZTest (:853-878)
      any = false;
      for (auto &p : px) {
        if (!p.inside) continue;
        if (p.count > 1) {
          bool in = false;
          for (int s = 0; s < 4; s++) {
            if (!p.s[s].inside) continue;
            p.s[s].inside = depthTest(p.s[s].fx, p.s[s].fy, p.s[s].z, s, true);
            in = in || p.s[s].inside;
          }
          p.inside = in;
        } else {
          SampleO &sm = p.s[p.shade];
          sm.inside = depthTest(sm.fx, sm.fy, sm.z, 0, true);
          p.inside = sm.inside;
        }
        any = any || p.inside;
      }
      if (!any) return;
    }
    for (auto &p : px) {   // varyings for all four pixels (:815-820), SIMD association (:1130-1139)
      const float *bc = p.s[p.shade].bc;
      for (int k = 0; k < nvary_; k++) p.vary[k] = fmaf(vx[2]->vary[k], bc[2], fmaf(vx[1]->vary[k], bc[1], vx[0]->vary[k] * bc[0]));
    }
    bool deriv = needsDeriv(prog_);
    ShaderEnv env = {prog_, deriv ? px[0].vary : nullptr, px[1].vary, px[2].vary};
    for (auto &p : px) {
      if (!p.inside) continue;
      f4 c = {0, 0, 0, 0};
      if (color_) c = fragmentShader(env, p.vary);   // processFragmentShader returns early without colour buffer (:346-348)
      if (p.count > 1) {
        for (int s = 0; s < 4; s++)
          if (p.s[s].inside) perSample(p.s[s].fx, p.s[s].fy, p.s[s].z, c, s);
      } else {
        SampleO &sm = p.s[p.shade];
        perSample(sm.fx, sm.fy, sm.z, c, 0);
      }
    }
  }

  void resolve() {   // multiSampleResolve (:880-912)
    colorTex_->resolved.resize((size_t) colorTex_->width * colorTex_->height);
    const std::vector<uint32_t> &ms = colorTex_->levels[0][0];
    for (size_t i = 0; i < colorTex_->resolved.size(); i++) {
      uint32_t r = 0;
      for (int c = 0; c < 4; c++) {
        float sum = 0.f;
        for (int s = 0; s < 4; s++) sum += (float) ((ms[i * 4 + s] >> (8 * c)) & 255u);
        r |= ((uint32_t) (int) (sum / 4.f) & 255u) << (8 * c);
      }
      colorTex_->resolved[i] = r;
    }
  }

  FboOracle *fbo_ = nullptr;
  VaoOracle *vao_ = nullptr;
  ProgramOracle *prog_ = nullptr;
  const RenderStates *rs_ = nullptr;
  TextureOracle *colorTex_ = nullptr, *depthTex_ = nullptr;
  std::vector<uint32_t> *color_ = nullptr, *depth_ = nullptr;
  int fbW_ = 0, fbH_ = 0, ns_ = 1, nvary_ = 0;
  float vpX_ = 0, vpY_ = 0, vpW_ = 0, vpH_ = 0, pointSize_ = 1.f;
  std::vector<VertexO> verts_;
  std::vector<PrimO> prims_;
};

}  // namespace

// ------------------------------------------------------------------------------------------------ TextureOracle
TextureOracle::TextureOracle(const TextureDesc &d) : id_(gNextId++) {
  width = d.width; height = d.height; type = d.type; format = d.format; usage = d.usage;
  useMipmaps = d.useMipmaps; multiSample = d.multiSample; tag = d.tag;
}

void TextureOracle::allocate(bool withMips) {   // TextureSoft::initImageData / generateMipmaps(false) (TextureSoft.h:133-141, SamplerSoft.h:90-110)
  int nl = 1;
  if (withMips) nl = (int) std::floor(std::log2((double) std::max(width, height))) + 1;
  levels.assign(layers(), {});
  for (auto &layer : levels) {
    layer.resize(nl);
    for (int l = 0; l < nl; l++) layer[l].assign((size_t) levelDim(width, l) * levelDim(height, l) * samples(), 0u);
  }
}
void TextureOracle::initImageData() { allocate(useMipmaps); }

void TextureOracle::generateMipmaps() {   // sampleBufferBilinear (SamplerSoft.h:241-252)
  Sampler s;
  s.tex = this;
  s.filter = Filter_LINEAR;
  s.wrap = Wrap_CLAMP_TO_EDGE;
  for (int layer = 0; layer < layers(); layer++)
    for (int l = 1; l < levelCount(); l++) {
      int ow = levelDim(width, l), oh = levelDim(height, l), iw = levelDim(width, l - 1), ih = levelDim(height, l - 1);
      float rx = (float) iw / (float) ow, ry = (float) ih / (float) oh;
      for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++)
          levels[layer][l][(size_t) y * ow + x] = pixelBilinear(s, layer, l - 1, (float) x * rx + 0.5f * rx, (float) y * ry + 0.5f * ry);
    }
}

template<typename T>
static void uploadTo(TextureOracle &t, const std::vector<std::shared_ptr<Buffer<T>>> &b) {   // TextureSoft::setImageData (TextureSoft.h:110-131)
  if (t.multiSample || b.empty() || (size_t) t.width != b[0]->getWidth() || (size_t) t.height != b[0]->getHeight()) return;
  t.allocate(t.useMipmaps);
  for (int i = 0; i < t.layers() && i < (int) b.size(); i++) memcpy(t.levels[i][0].data(), b[i]->getRawDataPtr(), (size_t) t.width * t.height * 4);
  if (t.useMipmaps) t.generateMipmaps();
}
void TextureOracle::setImageData(const std::vector<std::shared_ptr<Buffer<RGBA>>> &b) { uploadTo(*this, b); }
void TextureOracle::setImageData(const std::vector<std::shared_ptr<Buffer<float>>> &b) { uploadTo(*this, b); }

std::shared_ptr<Renderer> createRendererOracle() { return std::make_shared<RendererOracle>(); }

bool oracleLoadShaders(ShaderProgram &program, int shading) {   // ShaderProgramSoft::SetShaders (ShaderProgramSoft.h:26-59)
  auto *p = dynamic_cast<ProgramOracle *>(&program);
  const Meta *m = metaOf(shading);
  if (!p || !m) return false;
  p->shading = shading;
  p->defines = 0;
  for (auto &d : p->defineNames)
    for (int i = 0; i < 8 && m->defines[i]; i++)
      if (d == m->defines[i]) p->defines |= 1u << i;
  return true;
}


// ---- unit-level known-answer entry points (same file protocol as oracle/ref_kat.cpp, so that the restatement can be
//      checked against the vectors the reference produced: tests/golden/unit_kats.npz)
static std::vector<uint8_t> katReadFile(const char *path) {
  std::vector<uint8_t> v;
  FILE *f = fopen(path, "rb");
  if (!f) return v;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize(n);
  if (n && fread(v.data(), 1, n, f) != (size_t) n) v.clear();
  fclose(f);
  return v;
}

int oracleKatMain(int argc, char **argv) {
  if (argc >= 4 && !strcmp(argv[1], "sample")) {
    struct { int32_t w, h, layers, format, mips, filter, wrap, border, n, hasOffset; } H;
    std::vector<uint8_t> blob = katReadFile(argv[2]);
    if (blob.size() < sizeof(H)) return 2;
    memcpy(&H, blob.data(), sizeof(H));
    const uint8_t *p = blob.data() + sizeof(H);
    TextureDesc d;
    d.width = H.w; d.height = H.h; d.type = H.layers == 6 ? TextureType_CUBE : TextureType_2D;
    d.format = (TextureFormat) H.format; d.useMipmaps = H.mips != 0; d.multiSample = false;
    TextureOracle tex(d);
    tex.allocate(d.useMipmaps);
    const size_t texels = (size_t) H.w * H.h;
    for (int l = 0; l < H.layers; l++) memcpy(tex.levels[l][0].data(), p + (size_t) l * texels * 4, texels * 4);
    if (d.useMipmaps) tex.generateMipmaps();
    p += texels * 4 * H.layers;
    const int comps = H.layers == 6 ? 3 : 2;
    const float *coords = (const float *) p, *lod = coords + (size_t) comps * H.n;
    const int32_t *offs = (const int32_t *) (lod + H.n);
    Sampler s;
    s.tex = &tex; s.filter = H.filter; s.wrap = H.wrap;
    float bf = H.border == Border_WHITE ? 1.f : 0.f;          // TextureSoft::getBorderColor (TextureSoft.h:158-164)
    s.border = H.format == TextureFormat_FLOAT32 ? floatToBits(bf) : (H.border == Border_WHITE ? 0xFFFFFFFFu : 0u);
    std::vector<uint32_t> out(H.n);
    for (int i = 0; i < H.n; i++) {
      if (H.layers == 6) {
        int face; float u, v;
        cubeFaceUV(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], face, u, v);
        out[i] = textureImpl(s, face, u, v, lod[i]);
      } else {
        out[i] = textureImpl(s, 0, coords[2 * i], coords[2 * i + 1], lod[i], H.hasOffset ? offs[2 * i] : 0, H.hasOffset ? offs[2 * i + 1] : 0);
      }
    }
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 2;
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "bary")) {
    std::vector<uint8_t> blob = katReadFile(argv[2]);
    if (blob.size() < 8) return 2;
    int32_t nTri, nPer;
    memcpy(&nTri, blob.data(), 4);
    memcpy(&nPer, blob.data() + 4, 4);
    const float *p = (const float *) (blob.data() + 8);
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 2;
    for (int t = 0; t < nTri; t++) {
      f4 v[3];
      for (int k = 0; k < 3; k++) v[k] = {p[4 * k], p[4 * k + 1], p[4 * k + 2], p[4 * k + 3]};
      p += 12;
      for (int sidx = 0; sidx < nPer; sidx++, p += 2) {
        float bc[3] = {0.f, 0.f, 0.f}, zw[2] = {0.f, 0.f};
        bool in = barycentricO(v, p[0], p[1], bc);
        if (in) {   // interpolateBarycentric(&position.z, vertZ, 2, bc) (RendererSoft.cpp:794,1110-1112)
          float bz[4] = {bc[0], bc[1], bc[2], 0.f}, zs[4] = {v[0].z, v[1].z, v[2].z, 0.f}, ws[4] = {v[0].w, v[1].w, v[2].w, 0.f};
          zw[0] = dpps(bz, zs);
          zw[1] = dpps(bz, ws);
        }
        int32_t in32 = in ? 1 : 0;
        fwrite(&in32, 4, 1, f);
        fwrite(bc, 4, 3, f);
        fwrite(zw, 4, 2, f);
      }
    }
    fclose(f);
    return 0;
  }
  fprintf(stderr, "usage: oracle_kat sample in out | bary in out\n");
  return 1;
}

}  // namespace SoftGL
