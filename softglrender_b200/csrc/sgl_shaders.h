// The reference's software shader programs as device functions (north_star "Shading").
// Each block cites the C++ "shader" it restates under src/Viewer/Shader/Software/.
// Uniforms are read from the draw's snapshot of the program's uniform byte buffer, laid out exactly
// like the reference's ShaderUniforms structs (GLM aligned types: vec3 = 16 B, mat3 = 48 B):
//   UniformsModel    @0    {u32 reverseZ @0, mat4 model @16, mat4 mvp @80, mat3 invT @144, mat4 shadowMVP @192}
//   UniformsScene    @256  {vec3 ambient, vec3 cameraPos @272, vec3 lightPos @288, vec3 lightColor @304}
//   UniformsMaterial @320  {u32 enableLight, enableIBL, enableShadow, f32 pointSize @+12, kSpecular @+16, vec4 baseColor @+32}
//   (ShaderBasic has no scene block: material @256.)
#pragma once
#include <string.h>
#include "sgl_texture.h"

// ---- reflection tables (ShaderSoft::getDefines / getUniformsDesc) --------------------------------------
struct SglShaderInfo {
  int varyingCount;        // sizeof(ShaderVaryings)/4
  int varyingStride;       // MemoryUtils::alignedSize => multiple of 8 floats (RendererSoft.cpp:172-174)
  int uniformBytes;        // block part of ShaderUniforms
  int materialOffset;      // offset of UniformsMaterial or -1
};

SGL_HD SglShaderInfo sglShaderInfo(int shader) {
  SglShaderInfo i = {0, 0, 0, -1};
  switch (shader) {
    case SGL_SHADER_BASIC: i.varyingCount = 0; i.varyingStride = 0; i.uniformBytes = 304; i.materialOffset = 256; break;
    case SGL_SHADER_BLINNPHONG: i.varyingCount = 32; i.varyingStride = 32; i.uniformBytes = 368; i.materialOffset = 320; break;
    case SGL_SHADER_PBR: i.varyingCount = 28; i.varyingStride = 32; i.uniformBytes = 368; i.materialOffset = 320; break;
    case SGL_SHADER_SKYBOX: i.varyingCount = 4; i.varyingStride = 8; i.uniformBytes = 256; break;
    case SGL_SHADER_IBL_IRRADIANCE: i.varyingCount = 4; i.varyingStride = 8; i.uniformBytes = 256; break;
    case SGL_SHADER_IBL_PREFILTER: i.varyingCount = 4; i.varyingStride = 8; i.uniformBytes = 264; break;
    case SGL_SHADER_FXAA: i.varyingCount = 2; i.varyingStride = 8; i.uniformBytes = 8; break;
  }
  return i;
}

// define bits
#define SGL_DEF_ALBEDO_MAP 1u
#define SGL_DEF_NORMAL_MAP 2u
#define SGL_DEF_EMISSIVE_MAP 4u
#define SGL_DEF_AO_MAP 8u
#define SGL_DEF_METALROUGHNESS_MAP 16u
#define SGL_DEF_EQUIRECTANGULAR_MAP 1u

// sampler slots
enum { SGL_SLOT_ALBEDO = 0, SGL_SLOT_NORMAL = 1, SGL_SLOT_EMISSIVE = 2, SGL_SLOT_AO = 3,
       SGL_SLOT_BP_SHADOW = 4, SGL_SLOT_PBR_METALROUGH = 4, SGL_SLOT_PBR_IRRADIANCE = 5, SGL_SLOT_PBR_PREFILTER = 6,
       SGL_SLOT_SKY_EQUIRECT = 0, SGL_SLOT_SKY_CUBE = 1, SGL_SLOT_FXAA_SCREEN = 0, SGL_SLOT_IBL_CUBE = 0 };

SGL_HD float uF(const SglDrawRec &d, int off) { float f; memcpy(&f, d.uniforms + off, 4); return f; }
SGL_HD int uI(const SglDrawRec &d, int off) { int v; memcpy(&v, d.uniforms + off, 4); return v; }
SGL_HD V3 uV3(const SglDrawRec &d, int off) { return v3(uF(d, off), uF(d, off + 4), uF(d, off + 8)); }
SGL_HD V4 uV4(const SglDrawRec &d, int off) { return v4(uF(d, off), uF(d, off + 4), uF(d, off + 8), uF(d, off + 12)); }
SGL_HD const float *uMat(const SglDrawRec &d, int off) { return (const float *) (d.uniforms + off); }

// ---- vertex shaders -----------------------------------------------------------------------------------
// vin: 16 floats {pos@0, uv@4, normal@8, tangent@12}; vout: varyingStride floats (zero-filled where the
// reference leaves memory untouched).  Returns gl_Position.  gl_PointSize is per draw (d.pointSize).
SGL_HD V3 sglMat3MulCols(const float *c0, const float *c1, const float *c2, V3 v) {
  // glm mat3 * vec3 as compiled: fma(z, m2, fma(x, m0, y*m1))
  return v3(fmaf(v.z, c2[0], fmaf(v.x, c0[0], v.y * c1[0])), fmaf(v.z, c2[1], fmaf(v.x, c0[1], v.y * c1[1])),
            fmaf(v.z, c2[2], fmaf(v.x, c0[2], v.y * c1[2])));
}

// gl_Position alone (same arithmetic as sglVertexShader): what the geometry stage needs before it knows which vertices
// belong to primitives that survive clipping / culling / tile ownership.  Reads vin[0..2] only.
SGL_HD V4 sglVertexPosition(const SglDrawRec &d, const float *vin) {
  float px = vin[0], py = vin[1], pz = vin[2];
  if (d.shader == SGL_SHADER_FXAA) return v4(px, py, pz, 1.0f);
  V4 pos = xMat4MulPoint(uMat(d, 80), px, py, pz);
  if (d.shader == SGL_SHADER_SKYBOX) {
    V4 r = v4(pos.x, pos.y, pos.w, pos.w);
    if (uI(d, 0)) r.z = 0.f;
    return r;
  }
  if (d.shader == SGL_SHADER_IBL_IRRADIANCE || d.shader == SGL_SHADER_IBL_PREFILTER) return v4(pos.x, pos.y, pos.w, pos.w);
  return pos;
}

SGL_HD V4 sglVertexShader(const SglDrawRec &d, const float *vin, float *vout) {
  const float *mvp = uMat(d, 80);
  float px = vin[0], py = vin[1], pz = vin[2];
  V4 pos = xMat4MulPoint(mvp, px, py, pz);
  switch (d.shader) {
    case SGL_SHADER_BASIC:            // BasicSoft.h:66-69
      return pos;
    case SGL_SHADER_FXAA: {           // FxaaSoft.h:58-61: gl_Position = vec4(a_position, 1)
#pragma unroll
      for (int i = 0; i < 8; i++) vout[i] = 0.f;
      vout[0] = vin[4];
      vout[1] = vin[5];
      return v4(px, py, pz, 1.0f);
    }
    case SGL_SHADER_SKYBOX: {         // SkyboxSoft.h:67-76: pos.xyww, z = 0 when reverseZ
#pragma unroll
      for (int i = 0; i < 8; i++) vout[i] = 0.f;
      vout[0] = px; vout[1] = py; vout[2] = pz;
      V4 r = v4(pos.x, pos.y, pos.w, pos.w);
      if (uI(d, 0)) r.z = 0.f;
      return r;
    }
    case SGL_SHADER_IBL_IRRADIANCE:   // IBLIrradianceSoft.h:62-67
    case SGL_SHADER_IBL_PREFILTER: {  // IBLPrefilterSoft.h:69-74: Position.z = pos.w
#pragma unroll
      for (int i = 0; i < 8; i++) vout[i] = 0.f;
      vout[0] = px; vout[1] = py; vout[2] = pz;
      return v4(pos.x, pos.y, pos.w, pos.w);
    }
    default: break;
  }
  // BlinnPhongSoft.h:103-121 / PbrSoft.h:110-127
  bool bp = d.shader == SGL_SHADER_BLINNPHONG;
#pragma unroll
  for (int i = 0; i < 32; i++) vout[i] = 0.f;
  const float *model = uMat(d, 16);
  vout[0] = vin[4];
  vout[1] = vin[5];
  V4 wp = xMat4MulPoint(model, px, py, pz);
  V3 nrm = v3(vin[8], vin[9], vin[10]);
  V3 nv = sglMat3MulCols(model, model + 4, model + 8, nrm);
  vout[4] = nv.x; vout[5] = nv.y; vout[6] = nv.z;
  vout[8] = wp.x; vout[9] = wp.y; vout[10] = wp.z;
  V3 cam = uV3(d, 272), light = uV3(d, 288);
  vout[12] = cam.x - wp.x; vout[13] = cam.y - wp.y; vout[14] = cam.z - wp.z;
  vout[16] = light.x - wp.x; vout[17] = light.y - wp.y; vout[18] = light.z - wp.z;
  if (bp) {
    V4 sp = xMat4MulPoint(uMat(d, 192), px, py, pz);
    vout[20] = sp.x; vout[21] = sp.y; vout[22] = sp.z; vout[23] = sp.w;
  }
  if (d.defines & SGL_DEF_NORMAL_MAP) {
    const float *it = uMat(d, 144);
    V3 N = normalize(sglMat3MulCols(it, it + 4, it + 8, nrm));
    V3 T = normalize(sglMat3MulCols(it, it + 4, it + 8, v3(vin[12], vin[13], vin[14])));
    V3 T2 = normalize(T - dot(T, N) * N);
    // constant indices in both arms: the varyings of a thread stay in registers (a runtime offset sends the array to local memory)
    if (bp) {
      vout[24] = N.x; vout[25] = N.y; vout[26] = N.z;
      vout[28] = T2.x; vout[29] = T2.y; vout[30] = T2.z;
    } else {
      vout[20] = N.x; vout[21] = N.y; vout[22] = N.z;
      vout[24] = T2.x; vout[25] = T2.y; vout[26] = T2.z;
    }
  }
  return pos;
}

// ---- fragment shader context -----------------------------------------------------------------------------
struct SglFsCtx {
  const SglDrawRec *draw;
  const SglTexObj *textures;   // device texture table; entry 0 is a 1x1 dummy texture (unbound maps of the fast paths)
  // texture coordinates of quad pixels p0, p1, p2 (DerivativeContext, ShaderSoft.h:20-25); valid only when a
  // mip-filtered 2D sampler is bound
  bool derivValid;
  V2 uv0, uv1, uv2;
};

SGL_HD SglSampler sglSlot(const SglFsCtx &c, int slot) {
  SglSampler s;
  const SglSamplerSlot &b = c.draw->samplers[slot];
  s.tex = b.tex >= 0 ? &c.textures[b.tex] : nullptr;
  s.filter = b.filter;
  s.wrap = b.wrap;
  s.border = b.border;
  return s;
}

// ShaderSoft::getSampler2DLod (ShaderSoft.h:117-133), applied by BaseSampler2D::texture2DImpl only when the
// sampler's min filter uses mipmaps (SamplerSoft.h:73-79)
SGL_HD float sglImplicitLod(const SglFsCtx &c, const SglSampler &s) {
  if (s.filter <= SGL_FILTER_LINEAR || !c.derivValid || s.tex == nullptr) return 0.f;
  float w = (float) s.tex->width, h = (float) s.tex->height;
  V2 dx = v2((c.uv1.x - c.uv0.x) * w, (c.uv1.y - c.uv0.y) * h);
  V2 dy = v2((c.uv2.x - c.uv0.x) * w, (c.uv2.y - c.uv0.y) * h);
  float dd = gmax(dot(dx, dx), dot(dy, dy));
  return gmax(0.5f * log2f(dd), 0.0f);
}

SGL_HD V4 sglTexLod2D(const SglFsCtx &c, int slot, V2 uv) {   // texture(sampler2D with lodFunc, uv)
  SglSampler s = sglSlot(c, slot);
  return sglTexture2D(s, uv, sglImplicitLod(c, s));
}

// ---- shared pieces of BlinnPhong / PBR ---------------------------------------------------------------
SGL_HD V3 sglNormalFromMap(const SglFsCtx &c, const float *v, int nOff, V2 uv) {
  // GetNormalFromMap (BlinnPhongSoft.h:149-163, PbrSoft.h:153-167)
  if (c.draw->defines & SGL_DEF_NORMAL_MAP) {
    V3 N = normalize(v3(v[nOff], v[nOff + 1], v[nOff + 2]));
    V3 T = normalize(v3(v[nOff + 4], v[nOff + 5], v[nOff + 6]));
    T = normalize(T - dot(T, N) * N);
    V3 B = cross(T, N);
    V4 t = sglTexLod2D(c, SGL_SLOT_NORMAL, uv);
    V3 tn = v3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
    return normalize(T * tn.x + B * tn.y + N * tn.z);
  }
  return normalize(v3(v[4], v[5], v[6]));
}

// ---- ShaderBlinnPhong::FS (BlinnPhongSoft.h:123-242) -----------------------------------------------------
SGL_HD float sglShadowCalc(const SglFsCtx &c, V4 fragPos, V3 normal, V3 lightDir) {
  V3 proj = v3(fragPos.x / fragPos.w, fragPos.y / fragPos.w, fragPos.z / fragPos.w);
  float cur = proj.z;
  if (cur < 0.f || cur > 1.f) return 0.0f;
  float bias = gmax(0.00025f * (1.0f - dot(normal, normalize(lightDir))), 0.00005f);
  SglSampler sm = sglSlot(c, SGL_SLOT_BP_SHADOW);
  if (sm.tex == nullptr) return 0.0f;
  V2 po = v2(1.0f / (float) sm.tex->width, 1.0f / (float) sm.tex->height);
  bool rev = uI(*c.draw, 0) != 0;
  float shadow = 0.0f;
  if (sm.filter == SGL_FILTER_NEAREST && sm.tex->format == SGL_FMT_FLOAT32 && sm.tex->layout == SGL_LAYOUT_LINEAR &&
      sm.tex->samples == 1 && sm.tex->base != nullptr) {
    // the shadow map as the Viewer binds it (NEAREST float texture): same nine sampleNearest taps, with the level view
    // resolved once and all nine texel loads in flight before the first comparison
    const int w = sm.tex->width, h = sm.tex->height;
    const uint32_t *base = (const uint32_t *) (sm.tex->base + sm.tex->levelOffset[0]);
    uint32_t tap[9];
    int k = 0;
    for (int x = -1; x <= 1; ++x) {
      for (int y = -1; y <= 1; ++y, ++k) {
        int ix = (int) floorf(xmul(proj.x + (float) x * po.x, (float) w));
        int iy = (int) floorf(xmul(proj.y + (float) y * po.y, (float) h));
        const int rx = sglWrapAxis(ix, w, sm.wrap), ry = sglWrapAxis(iy, h, sm.wrap);
        const bool border = rx == 1 || ry == 1, oob = (rx | ry) != 0;
#if defined(SGL_TOUCH_BITMAP) && defined(__CUDA_ARCH__)
        if (!(border || oob)) sglTouch(sm.tex->touch, sm.tex->base, base + (uint32_t) iy * (uint32_t) w + (uint32_t) ix);
#endif
        const uint32_t t = SGL_LDG(base + (border || oob ? 0 : (uint32_t) iy * (uint32_t) w + (uint32_t) ix));
        tap[k] = border ? sm.border : (oob ? 0u : t);
      }
    }
    for (k = 0; k < 9; k++) {
      const float pcf = sglBitsFloat(tap[k]);
      if (rev) shadow += (cur + bias < pcf) ? 1.0f : 0.0f;
      else shadow += (cur - bias > pcf) ? 1.0f : 0.0f;
    }
    return shadow / 9.0f;
  }
  for (int x = -1; x <= 1; ++x) {
    for (int y = -1; y <= 1; ++y) {
      float pcf = sglTexture2DFloat(sm, v2(proj.x + (float) x * po.x, proj.y + (float) y * po.y));
      if (rev) shadow += (cur + bias < pcf) ? 1.0f : 0.0f;
      else shadow += (cur - bias > pcf) ? 1.0f : 0.0f;
    }
  }
  return shadow / 9.0f;
}

SGL_HD V4 sglFsBlinnPhong(const SglFsCtx &c, const float *v) {
  const SglDrawRec &d = *c.draw;
  V2 uv = v2(v[0], v[1]);
  V4 base = (d.defines & SGL_DEF_ALBEDO_MAP) ? sglTexLod2D(c, SGL_SLOT_ALBEDO, uv) : uV4(d, 352);
  V3 N = sglNormalFromMap(c, v, 24, uv);
  float ao = 1.f;
  if (d.defines & SGL_DEF_AO_MAP) ao = sglTexLod2D(c, SGL_SLOT_AO, uv).x;
  V3 baseRgb = v3(base.x, base.y, base.z);
  V3 ambient = baseRgb * uV3(d, 256) * ao;
  V3 diffuse = v3s(0.f), specular = v3s(0.f), emissive = v3s(0.f);
  V3 lightVec = v3(v[16], v[17], v[18]);
  if (uI(d, 320)) {
    V3 lDir = lightVec * (1.0f / 5.f);
    float atten = clampf(1.0f - dot(lDir, lDir), 0.0f, 1.0f);
    V3 L = normalize(lightVec);
    float diff = fmaxf(dot(N, L), 0.0f);
    diffuse = uV3(d, 304) * baseRgb * diff * atten;
    V3 camDir = normalize(v3(v[12], v[13], v[14]));
    V3 H = normalize(L + camDir);
    float sa = fmaxf(dot(N, H), 0.0f);
    specular = v3s(uF(d, 336) * spow(sa, 128.f));
    if (uI(d, 328)) {
      float shadow = 1.0f - sglShadowCalc(c, v4(v[20], v[21], v[22], v[23]), N, lightVec);
      diffuse = diffuse * shadow;
      specular = specular * shadow;
    }
  }
  if (d.defines & SGL_DEF_EMISSIVE_MAP) {
    V4 e = sglTexLod2D(c, SGL_SLOT_EMISSIVE, uv);
    emissive = v3(e.x, e.y, e.z);
  }
  V3 col = ambient + diffuse + specular + emissive;
  return v4(col.x, col.y, col.z, base.w);
}

// ---- ShaderPbrIBL::FS (PbrSoft.h:129-322) -------------------------------------------------------------------
#define SGL_PI 3.14159265359f
SGL_HD float sglDistributionGGX(V3 N, V3 H, float roughness) {
  float a = roughness * roughness;
  float a2 = a * a;
  float NdotH = fmaxf(dot(N, H), 0.0f);
  float NdotH2 = NdotH * NdotH;
  float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
  denom = SGL_PI * denom * denom;
  return sdiv(a2, denom);
}
SGL_HD float sglDistributionGGXP(V3 N, V3 H, float roughness) {   // IEEE division (IBL prefilter pass)
  float a = roughness * roughness;
  float a2 = a * a;
  float NdotH = fmaxf(dot(N, H), 0.0f);
  float NdotH2 = NdotH * NdotH;
  float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
  denom = SGL_PI * denom * denom;
  return a2 / denom;
}
SGL_HD float sglGeometrySchlickGGX(float NdotV, float roughness) {
  float r = roughness + 1.0f;
  float k = (r * r) / 8.0f;
  return sdiv(NdotV, NdotV * (1.0f - k) + k);
}
SGL_HD float sglGeometrySmith(V3 N, V3 V, V3 L, float roughness) {
  float NdotV = fmaxf(dot(N, V), 0.0f), NdotL = fmaxf(dot(N, L), 0.0f);
  return sglGeometrySchlickGGX(NdotL, roughness) * sglGeometrySchlickGGX(NdotV, roughness);
}
SGL_HD V3 sglEnvBRDFApprox(V3 spec, float rough, float NdotV) {
  V4 r = v4(rough * -1.f + 1.f, rough * -0.0275f + 0.0425f, rough * -0.572f + 1.04f, rough * 0.022f + -0.04f);
  float a004 = fminf(r.x * r.x, exp2f(-9.28f * NdotV)) * r.x + r.y;
  V2 AB = v2(-1.04f * a004 + r.z, 1.04f * a004 + r.w);
  AB.y *= fmaxf(0.f, fminf(1.f, 50.0f * spec.y));
  return spec * AB.x + v3s(AB.y);
}

SGL_HD V4 sglFsPbr(const SglFsCtx &c, const float *v) {
  const SglDrawRec &d = *c.draw;
  V2 uv = v2(v[0], v[1]);
  V4 albedoRgba = (d.defines & SGL_DEF_ALBEDO_MAP) ? sglTexLod2D(c, SGL_SLOT_ALBEDO, uv) : uV4(d, 352);
  V3 albedo = vpow(v3(albedoRgba.x, albedoRgba.y, albedoRgba.z), 2.2f);
  float metallic = 0.0f, roughness = 1.0f;
  if (d.defines & SGL_DEF_METALROUGHNESS_MAP) {
    V4 mr = sglTexture2D(sglSlot(c, SGL_SLOT_PBR_METALROUGH), uv, 0.f);   // no lodFunc on this sampler (PbrSoft.h:138-151)
    metallic = mr.z;
    roughness = mr.y;
  }
  float ao = 1.f;
  if (d.defines & SGL_DEF_AO_MAP) ao = sglTexLod2D(c, SGL_SLOT_AO, uv).x;
  V3 N = sglNormalFromMap(c, v, 20, uv);
  V3 V = normalize(v3(v[12], v[13], v[14]));
  V3 R = reflect(-V, N);
  V3 F0 = vmix(v3s(0.04f), albedo, metallic);
  V3 Lo = v3s(0.0f);
  V3 lightVec = v3(v[16], v[17], v[18]);
  if (uI(d, 320)) {
    V3 L = normalize(lightVec);
    V3 H = normalize(V + L);
    V3 lDir = lightVec * (1.0f / 5.f);
    float atten = clampf(1.0f - dot(lDir, lDir), 0.0f, 1.0f);
    V3 radiance = uV3(d, 304) * atten;
    float NDF = sglDistributionGGX(N, H, roughness);
    float G = sglGeometrySmith(N, V, L, roughness);
    float p5 = spow(clampf(1.0f - fmaxf(dot(H, V), 0.0f), 0.0f, 1.0f), 5.0f);
    V3 F = F0 + (v3s(1.0f) - F0) * p5;
    V3 numerator = F * (NDF * G);
    float denominator = 4.0f * fmaxf(dot(N, V), 0.0f) * fmaxf(dot(N, L), 0.0f) + 0.0001f;
    V3 specular = numerator / denominator;
    V3 kD = (v3s(1.0f) - F) * (1.0f - metallic);
    float NdotL = fmaxf(dot(N, L), 0.0f);
    Lo = Lo + (kD * albedo / SGL_PI + specular) * radiance * NdotL;
  }
  V3 ambient;
  if (uI(d, 324)) {
    float NdotV = fmaxf(dot(N, V), 0.0f);
    float p5 = spow(clampf(1.0f - NdotV, 0.0f, 1.0f), 5.0f);
    V3 F = F0 + (vmax(v3s(1.0f - roughness), F0) - F0) * p5;
    V3 kD = (v3s(1.0f) - F) * (1.0f - metallic);
    V4 irr = sglTextureCube(sglSlot(c, SGL_SLOT_PBR_IRRADIANCE), N, 0.f);
    V3 diffuse = v3(irr.x, irr.y, irr.z) * albedo;
    V4 pre = sglTextureCube(sglSlot(c, SGL_SLOT_PBR_PREFILTER), R, roughness * 4.0f);
    V3 specular = v3(pre.x, pre.y, pre.z) * sglEnvBRDFApprox(F, roughness, NdotV);
    ambient = (kD * diffuse + specular) * ao;
  } else {
    ambient = uV3(d, 256) * albedo * ao;
  }
  V3 color = vpow(ambient + Lo, 1.0f / 2.2f);
  V3 emissive = v3s(0.f);
  if (d.defines & SGL_DEF_EMISSIVE_MAP) {
    V4 e = sglTexLod2D(c, SGL_SLOT_EMISSIVE, uv);
    emissive = v3(e.x, e.y, e.z);
  }
  color = color + emissive;
  return v4(color.x, color.y, color.z, albedoRgba.w);
}


// ---- ShaderPbrIBL::FS, straight-line form for "simple" samplers ------------------------------------------------
// Same expressions as sglFsPbr in the same order; the only difference is that the texel loads of all material maps
// are issued before the first use (and unbound maps read a 1x1 dummy texture instead of branching), which turns seven
// dependent memory round trips per pixel into two.
SGL_HD SglTapView sglSlotTapView(const SglFsCtx &c, int slot, bool enabled, int layer, int level) {
  const SglSamplerSlot &b = c.draw->samplers[slot];
  const SglTexObj *t = &c.textures[enabled ? b.tex : 0];
  return sglTapView(t, enabled ? layer : 0, enabled ? level : 0, enabled ? b.wrap : SGL_WRAP_CLAMP_TO_EDGE);
}

SGL_HD V4 sglFsPbrFast(const SglFsCtx &c, const float *v) {
  const SglDrawRec &d = *c.draw;
  V2 uv = v2(v[0], v[1]);
  const bool hasAlbedo = (d.defines & SGL_DEF_ALBEDO_MAP) != 0, hasNormal = (d.defines & SGL_DEF_NORMAL_MAP) != 0;
  const bool hasEmissive = (d.defines & SGL_DEF_EMISSIVE_MAP) != 0, hasAo = (d.defines & SGL_DEF_AO_MAP) != 0;
  const bool hasMr = (d.defines & SGL_DEF_METALROUGHNESS_MAP) != 0;
  SglTap tAlbedo = sglTapIssue(sglSlotTapView(c, SGL_SLOT_ALBEDO, hasAlbedo, 0, 0), uv.x, uv.y);
  SglTap tMr = sglTapIssue(sglSlotTapView(c, SGL_SLOT_PBR_METALROUGH, hasMr, 0, 0), uv.x, uv.y);
  SglTap tAo = sglTapIssue(sglSlotTapView(c, SGL_SLOT_AO, hasAo, 0, 0), uv.x, uv.y);
  SglTap tNormal = sglTapIssue(sglSlotTapView(c, SGL_SLOT_NORMAL, hasNormal, 0, 0), uv.x, uv.y);
  SglTap tEmissive = sglTapIssue(sglSlotTapView(c, SGL_SLOT_EMISSIVE, hasEmissive, 0, 0), uv.x, uv.y);

  V4 albedoRgba = hasAlbedo ? sglUnpackRGBA(sglTapMix(tAlbedo)) : uV4(d, 352);
  V3 albedo = vpow(v3(albedoRgba.x, albedoRgba.y, albedoRgba.z), 2.2f);
  float metallic = 0.0f, roughness = 1.0f;
  {
    V4 mr = sglUnpackRGBA(sglTapMix(tMr));
    metallic = hasMr ? mr.z : metallic;
    roughness = hasMr ? mr.y : roughness;
  }
  float ao = hasAo ? sglUnpackRGBA(sglTapMix(tAo)).x : 1.f;
  V3 N;
  {
    V3 Nn = normalize(v3(v[20], v[21], v[22]));
    V3 T = normalize(v3(v[24], v[25], v[26]));
    T = normalize(T - dot(T, Nn) * Nn);
    V3 B = cross(T, Nn);
    V4 t = sglUnpackRGBA(sglTapMix(tNormal));
    V3 tn = v3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
    V3 mapped = normalize(T * tn.x + B * tn.y + Nn * tn.z);
    V3 plain = normalize(v3(v[4], v[5], v[6]));
    N = hasNormal ? mapped : plain;
  }
  V3 V = normalize(v3(v[12], v[13], v[14]));
  V3 R = reflect(-V, N);
  // IBL taps (irradiance: level 0; prefilter: two levels around roughness * 4), issued as soon as N, R and roughness exist
  const bool ibl = uI(d, 324) != 0;
  SglTap tIrr, tPreHi, tPreLo;
  float preFr = 0.f;
  bool preSame = true;
  {
    int face;
    float fu, fv;
    sglCubeFaceT<true>(N.x, N.y, N.z, face, fu, fv);
    tIrr = sglTapIssue(sglSlotTapView(c, SGL_SLOT_PBR_IRRADIANCE, ibl, face, 0), fu, fv);
    sglCubeFaceT<true>(R.x, R.y, R.z, face, fu, fv);
    const SglSamplerSlot &b = d.samplers[SGL_SLOT_PBR_PREFILTER];
    const SglTexObj *t = &c.textures[ibl ? b.tex : 0];
    float lod = roughness * 4.0f;
    int maxLevel = (ibl && b.filter == SGL_FILTER_LINEAR_MIPMAP_LINEAR) ? t->levels - 1 : 0;
    int hi = (int) floorf(lod);
    hi = hi < 0 ? 0 : (hi > maxLevel ? maxLevel : hi);
    int lo = hi + 1;
    lo = lo > maxLevel ? maxLevel : lo;
    preSame = hi == lo;
    preFr = xsub(lod, floorf(lod));
    tPreHi = sglTapIssue(sglSlotTapView(c, SGL_SLOT_PBR_PREFILTER, ibl, face, hi), fu, fv);
    tPreLo = sglTapIssue(sglSlotTapView(c, SGL_SLOT_PBR_PREFILTER, ibl, face, lo), fu, fv);
  }
  V3 F0 = vmix(v3s(0.04f), albedo, metallic);
  V3 Lo = v3s(0.0f);
  V3 lightVec = v3(v[16], v[17], v[18]);
  if (uI(d, 320)) {
    V3 L = normalize(lightVec);
    V3 H = normalize(V + L);
    V3 lDir = lightVec * (1.0f / 5.f);
    float atten = clampf(1.0f - dot(lDir, lDir), 0.0f, 1.0f);
    V3 radiance = uV3(d, 304) * atten;
    float NDF = sglDistributionGGX(N, H, roughness);
    float G = sglGeometrySmith(N, V, L, roughness);
    float p5 = spow(clampf(1.0f - fmaxf(dot(H, V), 0.0f), 0.0f, 1.0f), 5.0f);
    V3 F = F0 + (v3s(1.0f) - F0) * p5;
    V3 numerator = F * (NDF * G);
    float denominator = 4.0f * fmaxf(dot(N, V), 0.0f) * fmaxf(dot(N, L), 0.0f) + 0.0001f;
    V3 specular = numerator / denominator;
    V3 kD = (v3s(1.0f) - F) * (1.0f - metallic);
    float NdotL = fmaxf(dot(N, L), 0.0f);
    Lo = Lo + (kD * albedo / SGL_PI + specular) * radiance * NdotL;
  }
  V3 ambient;
  if (ibl) {
    float NdotV = fmaxf(dot(N, V), 0.0f);
    float p5 = spow(clampf(1.0f - NdotV, 0.0f, 1.0f), 5.0f);
    V3 F = F0 + (vmax(v3s(1.0f - roughness), F0) - F0) * p5;
    V3 kD = (v3s(1.0f) - F) * (1.0f - metallic);
    V4 irr = sglUnpackRGBA(sglTapMix(tIrr));
    V3 diffuse = v3(irr.x, irr.y, irr.z) * albedo;
    uint32_t pHi = sglTapMix(tPreHi), pLo = sglTapMix(tPreLo);
    V4 pre = sglUnpackRGBA(preSame ? pHi : sglMixTexel(SGL_FMT_RGBA8, pHi, pLo, preFr));
    V3 specular = v3(pre.x, pre.y, pre.z) * sglEnvBRDFApprox(F, roughness, NdotV);
    ambient = (kD * diffuse + specular) * ao;
  } else {
    ambient = uV3(d, 256) * albedo * ao;
  }
  V3 color = vpow(ambient + Lo, 1.0f / 2.2f);
  V4 e = sglUnpackRGBA(sglTapMix(tEmissive));
  V3 emissive = hasEmissive ? v3(e.x, e.y, e.z) : v3s(0.f);
  color = color + emissive;
  return v4(color.x, color.y, color.z, albedoRgba.w);
}

// host side of SglDrawRec::fastSamplers: is sampler slot `slot` of program `shader`, bound to texture t, "simple"?
SGL_HD bool sglSamplerIsSimple(int shader, int slot, const SglTexObj &t, int filter, int wrap) {
  if (t.base == nullptr || t.format != SGL_FMT_RGBA8 || t.samples != 1 || t.layout != SGL_LAYOUT_LINEAR) return false;
  if (wrap != SGL_WRAP_REPEAT && wrap != SGL_WRAP_CLAMP_TO_EDGE) return false;
  const bool cubeSlot = (shader == SGL_SHADER_PBR && (slot == SGL_SLOT_PBR_IRRADIANCE || slot == SGL_SLOT_PBR_PREFILTER)) ||
                        (shader == SGL_SHADER_SKYBOX && slot == SGL_SLOT_SKY_CUBE);
  if (t.layers != (cubeSlot ? 6 : 1)) return false;
  if (filter == SGL_FILTER_LINEAR) return true;
  return shader == SGL_SHADER_PBR && slot == SGL_SLOT_PBR_PREFILTER && filter == SGL_FILTER_LINEAR_MIPMAP_LINEAR;
}

// true when every sampler the PBR program is going to use for this draw is "simple" (SglDrawRec::fastSamplers)
SGL_HD bool sglPbrFastEligible(const SglDrawRec &d) {
  uint32_t need = d.defines & 0x1fu;
  if (uI(d, 324)) need |= (1u << SGL_SLOT_PBR_IRRADIANCE) | (1u << SGL_SLOT_PBR_PREFILTER);
  // slot order == define order for the five material maps except metalRoughness (define bit 4 -> slot 4): identical
  return (d.fastSamplers & need) == need;
}

// ---- ShaderSkybox::FS (SkyboxSoft.h:79-98) --------------------------------------------------------------------
SGL_HD V4 sglFsSkybox(const SglFsCtx &c, const float *v) {
  V3 wp = v3(v[0], v[1], v[2]);
  if (c.draw->defines & SGL_DEF_EQUIRECTANGULAR_MAP) {
    V3 dir = normalizeP(wp);   // IEEE form: this branch also performs the one-off equirect -> cube conversion
    V2 uv = v2(atan2f(dir.z, dir.x), asinf(-dir.y));
    uv = v2(uv.x * 0.1591f + 0.5f, uv.y * 0.3183f + 0.5f);
    return sglTexture2D(sglSlot(c, SGL_SLOT_SKY_EQUIRECT), uv, 0.f);   // no lodFunc installed for the skybox sampler
  }
  if ((c.draw->fastSamplers >> SGL_SLOT_SKY_CUBE) & 1u) {
    // "simple" cube (linear RGBA8, LINEAR filter): the same bilinear footprint through the branch-free split-phase tap
    int face;
    float fu, fv;
    sglCubeFaceT<true>(wp.x, wp.y, wp.z, face, fu, fv);
    const SglSamplerSlot &b = c.draw->samplers[SGL_SLOT_SKY_CUBE];
    return sglUnpackRGBA(sglTapMix(sglTapIssue(sglTapView(&c.textures[b.tex], face, 0, b.wrap), fu, fv)));
  }
  return sglTextureCube(sglSlot(c, SGL_SLOT_SKY_CUBE), wp, 0.f);
}

// ---- ShaderFXAA::FS (FxaaSoft.h:63-266) ----------------------------------------------------------------------
SGL_HD float sglLuma(V4 c) { return c.x * 0.299f + c.y * 0.587f + c.z * 0.114f; }
SGL_HD float sglFxaaQuality(int i) {
  float q = (float) i;
  return q < 5.f ? 1.0f : (q > 5.f ? (q < 10.f ? 2.0f : (q < 11.f ? 4.0f : 8.0f)) : 1.5f);
}

// screen-texture fetch of the FXAA pass: the split-phase tap when the sampler is "simple" (linear RGBA8, LINEAR filter:
// what the Viewer binds), the generic sampler otherwise -- same footprint, same truncating mixes
struct SglFxaaTex {
  SglSampler s;
  SglTapView tv;
  bool fast;
  SGL_HD V4 fetch(V2 uv, int ox, int oy) const {
    if (fast) return sglUnpackRGBA(sglTapMix(sglTapIssue(tv, uv.x, uv.y, ox, oy)));
    return sglTexture2DOffset(s, uv, 0.f, ox, oy);
  }
};

SGL_HD V4 sglFsFxaa(const SglFsCtx &c, const float *v) {
  const SglDrawRec &d = *c.draw;
  SglFxaaTex s;
  s.s = sglSlot(c, SGL_SLOT_FXAA_SCREEN);
  s.fast = ((d.fastSamplers >> SGL_SLOT_FXAA_SCREEN) & 1u) != 0;
  s.tv = sglTapView(&c.textures[s.fast ? d.samplers[SGL_SLOT_FXAA_SCREEN].tex : 0], 0, 0, d.samplers[SGL_SLOT_FXAA_SCREEN].wrap);
  V2 uv = v2(v[0], v[1]);
  V2 inv = v2(1.0f / uF(d, 0), 1.0f / uF(d, 4));
  V4 colorCenter;
  float lumaDown, lumaUp, lumaLeft, lumaRight;
  if (s.fast) {   // the five taps of the early-out test in flight together
    SglTap tc = sglTapIssue(s.tv, uv.x, uv.y), td = sglTapIssue(s.tv, uv.x, uv.y, 0, -1), tu = sglTapIssue(s.tv, uv.x, uv.y, 0, 1);
    SglTap tl = sglTapIssue(s.tv, uv.x, uv.y, -1, 0), tr = sglTapIssue(s.tv, uv.x, uv.y, 1, 0);
    colorCenter = sglUnpackRGBA(sglTapMix(tc));
    lumaDown = sglLuma(sglUnpackRGBA(sglTapMix(td)));
    lumaUp = sglLuma(sglUnpackRGBA(sglTapMix(tu)));
    lumaLeft = sglLuma(sglUnpackRGBA(sglTapMix(tl)));
    lumaRight = sglLuma(sglUnpackRGBA(sglTapMix(tr)));
  } else {
    colorCenter = s.fetch(uv, 0, 0);
    lumaDown = sglLuma(s.fetch(uv, 0, -1));
    lumaUp = sglLuma(s.fetch(uv, 0, 1));
    lumaLeft = sglLuma(s.fetch(uv, -1, 0));
    lumaRight = sglLuma(s.fetch(uv, 1, 0));
  }
  float lumaCenter = sglLuma(colorCenter);
  float lumaMin = fminf(lumaCenter, fminf(fminf(lumaDown, lumaUp), fminf(lumaLeft, lumaRight)));
  float lumaMax = fmaxf(lumaCenter, fmaxf(fmaxf(lumaDown, lumaUp), fmaxf(lumaLeft, lumaRight)));
  float lumaRange = lumaMax - lumaMin;
  if (lumaRange < fmaxf(0.0312f, lumaMax * 0.125f)) return v4(colorCenter.x, colorCenter.y, colorCenter.z, 1.f);
  float lumaDownLeft = sglLuma(s.fetch(uv, -1, -1));
  float lumaUpRight = sglLuma(s.fetch(uv, 1, 1));
  float lumaUpLeft = sglLuma(s.fetch(uv, -1, 1));
  float lumaDownRight = sglLuma(s.fetch(uv, 1, -1));
  float lumaDownUp = lumaDown + lumaUp;
  float lumaLeftRight = lumaLeft + lumaRight;
  float lumaLeftCorners = lumaDownLeft + lumaUpLeft;
  float lumaDownCorners = lumaDownLeft + lumaDownRight;
  float lumaRightCorners = lumaDownRight + lumaUpRight;
  float lumaUpCorners = lumaUpRight + lumaUpLeft;
  float edgeHorizontal = fabsf(-2.0f * lumaLeft + lumaLeftCorners) + fabsf(-2.0f * lumaCenter + lumaDownUp) * 2.0f +
                         fabsf(-2.0f * lumaRight + lumaRightCorners);
  float edgeVertical = fabsf(-2.0f * lumaUp + lumaUpCorners) + fabsf(-2.0f * lumaCenter + lumaLeftRight) * 2.0f +
                       fabsf(-2.0f * lumaDown + lumaDownCorners);
  bool isHorizontal = edgeHorizontal >= edgeVertical;
  float stepLength = isHorizontal ? inv.y : inv.x;
  float luma1 = isHorizontal ? lumaDown : lumaLeft;
  float luma2 = isHorizontal ? lumaUp : lumaRight;
  float gradient1 = luma1 - lumaCenter, gradient2 = luma2 - lumaCenter;
  bool is1Steepest = fabsf(gradient1) >= fabsf(gradient2);
  float gradientScaled = 0.25f * fmaxf(fabsf(gradient1), fabsf(gradient2));
  float lumaLocalAverage;
  if (is1Steepest) {
    stepLength = -stepLength;
    lumaLocalAverage = 0.5f * (luma1 + lumaCenter);
  } else {
    lumaLocalAverage = 0.5f * (luma2 + lumaCenter);
  }
  V2 currentUv = uv;
  if (isHorizontal) currentUv.y += stepLength * 0.5f;
  else currentUv.x += stepLength * 0.5f;
  V2 offset = isHorizontal ? v2(inv.x, 0.0f) : v2(0.0f, inv.y);
  float q0 = sglFxaaQuality(0);
  V2 uv1 = v2(currentUv.x - offset.x * q0, currentUv.y - offset.y * q0);
  V2 uv2 = v2(currentUv.x + offset.x * q0, currentUv.y + offset.y * q0);
  float lumaEnd1 = 0.f, lumaEnd2 = 0.f;
  bool reached1 = false, reached2 = false, reachedBoth = false;
  for (int i = 1; i < 12; i++) {
    if (!reached1) {
      lumaEnd1 = sglLuma(s.fetch(uv1, 0, 0)) - lumaLocalAverage;
      reached1 = fabsf(lumaEnd1) >= gradientScaled;
    }
    if (!reached2) {
      lumaEnd2 = sglLuma(s.fetch(uv2, 0, 0)) - lumaLocalAverage;
      reached2 = fabsf(lumaEnd2) >= gradientScaled;
    }
    reachedBoth = reached1 && reached2;
    float q = sglFxaaQuality(i);
    if (!reached1) { uv1.x -= offset.x * q; uv1.y -= offset.y * q; }
    if (!reached2) { uv2.x += offset.x * q; uv2.y += offset.y * q; }
    if (reachedBoth) break;
  }
  float distance1 = isHorizontal ? (uv.x - uv1.x) : (uv.y - uv1.y);
  float distance2 = isHorizontal ? (uv2.x - uv.x) : (uv2.y - uv.y);
  bool isDirection1 = distance1 < distance2;
  float distanceFinal = fminf(distance1, distance2);
  bool isLumaCenterSmaller = lumaCenter < lumaLocalAverage;
  bool correctVariation1 = (lumaEnd1 < 0.0f) != isLumaCenterSmaller;
  bool correctVariation2 = (lumaEnd2 < 0.0f) != isLumaCenterSmaller;
  bool correctVariation = isDirection1 ? correctVariation1 : correctVariation2;
  float edgeLength = distance1 + distance2;
  float pixelOffset = -distanceFinal / edgeLength + 0.5f;
  float finalOffset = correctVariation ? pixelOffset : 0.0f;
  float lumaAverage = (1.0f / 12.0f) * (2.0f * (lumaDownUp + lumaLeftRight) + lumaLeftCorners + lumaRightCorners);
  float sub1 = clampf(fabsf(lumaAverage - lumaCenter) / lumaRange, 0.0f, 1.0f);
  float sub2 = (-2.0f * sub1 + 3.0f) * sub1 * sub1;
  float subFinal = sub2 * sub2 * 0.75f;
  finalOffset = fmaxf(finalOffset, subFinal);
  V2 finalUv = uv;
  if (isHorizontal) finalUv.y += finalOffset * stepLength;
  else finalUv.x += finalOffset * stepLength;
  V4 fc = s.fetch(finalUv, 0, 0);
  return v4(fc.x, fc.y, fc.z, 1.f);
}

// ---- ShaderIBLIrradiance::FS (IBLIrradianceSoft.h:74-104) --------------------------------------------------------
SGL_HD V4 sglFsIrradiance(const SglFsCtx &c, const float *v) {
  SglSampler s = sglSlot(c, SGL_SLOT_IBL_CUBE);
  V3 N = normalizeP(v3(v[0], v[1], v[2]));
  V3 irr = v3s(0.f);
  V3 up = v3(0.f, 1.f, 0.f);
  V3 right = normalizeP(cross(up, N));
  up = normalizeP(cross(N, right));
  const float sampleDelta = 0.025f;
  float nr = 0.0f;
  for (float phi = 0.0f; phi < 2.0f * SGL_PI; phi += sampleDelta) {
    float sp = sinf(phi), cp = cosf(phi);
    for (float theta = 0.0f; theta < 0.5f * SGL_PI; theta += sampleDelta) {
      float st = sinf(theta), ct = cosf(theta);
      V3 ts = v3(st * cp, st * sp, ct);
      V3 sv = ts.x * right + ts.y * up + ts.z * N;
      V4 t = sglTextureCubeP(s, sv, 0.f);
      irr = irr + v3(t.x, t.y, t.z) * ct * st;
      nr += 1.0f;
    }
  }
  irr = SGL_PI * irr * (1.0f / nr);
  return v4(irr.x, irr.y, irr.z, 1.0f);
}

// ---- ShaderIBLPrefilter::FS (IBLPrefilterSoft.h:76-168) -------------------------------------------------------------
SGL_HD float sglRadicalInverse(uint32_t bits) {
  bits = (bits << 16u) | (bits >> 16u);
  bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
  bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
  bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
  bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
  return (float) ((double) bits * 2.3283064365386963e-10);
}

SGL_HD V4 sglFsPrefilter(const SglFsCtx &c, const float *v) {
  const SglDrawRec &d = *c.draw;
  SglSampler s = sglSlot(c, SGL_SLOT_IBL_CUBE);
  float srcRes = uF(d, 256), rough = uF(d, 260);
  V3 N = normalizeP(v3(v[0], v[1], v[2]));
  V3 V = N;
  V3 col = v3s(0.f);
  float totalWeight = 0.f;
  float a = rough * rough;
  V3 upv = fabsf(N.z) < 0.999f ? v3(0.f, 0.f, 1.f) : v3(1.f, 0.f, 0.f);
  V3 tangent = normalizeP(cross(upv, N));
  V3 bitangent = cross(N, tangent);
  for (uint32_t i = 0u; i < 1024u; ++i) {
    V2 Xi = v2((float) i / 1024.f, sglRadicalInverse(i));
    float phi = 2.0f * SGL_PI * Xi.x;
    float cosTheta = sqrtf((1.0f - Xi.y) / (1.0f + (a * a - 1.0f) * Xi.y));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    V3 Hh = v3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
    V3 H = normalizeP(tangent * Hh.x + bitangent * Hh.y + N * Hh.z);
    V3 L = normalizeP(2.0f * dot(V, H) * H - V);
    float NdotL = fmaxf(dot(N, L), 0.0f);
    if (NdotL > 0.0f) {
      float D = sglDistributionGGXP(N, H, rough);
      float NdotH = fmaxf(dot(N, H), 0.0f);
      float HdotV = fmaxf(dot(H, V), 0.0f);
      float pdf = D * NdotH / (4.0f * HdotV) + 0.0001f;
      float saTexel = 4.0f * SGL_PI / (6.0f * srcRes * srcRes);
      float saSample = 1.0f / (1024.f * pdf + 0.0001f);
      float mip = rough == 0.0f ? 0.0f : 0.5f * log2f(saSample / saTexel);
      V4 t = sglTextureCubeP(s, L, mip);
      col = col + v3(t.x, t.y, t.z) * NdotL;
      totalWeight += NdotL;
    }
  }
  col = v3(col.x / totalWeight, col.y / totalWeight, col.z / totalWeight);
  return v4(col.x, col.y, col.z, 1.0f);
}

// ---- dispatch ------------------------------------------------------------------------------------------------
SGL_HD V4 sglFragmentShader(const SglFsCtx &c, const float *varyings) {
  switch (c.draw->shader) {
    case SGL_SHADER_BASIC: return uV4(*c.draw, 288);          // BasicSoft.h:76-78: FragColor = u_baseColor
    case SGL_SHADER_BLINNPHONG: return sglFsBlinnPhong(c, varyings);
    case SGL_SHADER_PBR: return sglPbrFastEligible(*c.draw) ? sglFsPbrFast(c, varyings) : sglFsPbr(c, varyings);
    case SGL_SHADER_SKYBOX: return sglFsSkybox(c, varyings);
    case SGL_SHADER_FXAA: return sglFsFxaa(c, varyings);
    case SGL_SHADER_IBL_IRRADIANCE: return sglFsIrradiance(c, varyings);
    case SGL_SHADER_IBL_PREFILTER: return sglFsPrefilter(c, varyings);
  }
  return v4(0.f, 0.f, 0.f, 0.f);
}
