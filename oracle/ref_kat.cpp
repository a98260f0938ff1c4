// ORACLE (test infrastructure, not product code): unit-level known-answer vectors produced by the REFERENCE itself.
// Compiled against the reference headers and linked with the reference's RendererSoft.o by oracle/Makefile
// (-> oracle/_ref/ref_kat); tests/golden/make_unit_kats.py drives it and commits the vectors (tests/golden/unit_kats.npz).
//
//   ref_kat layout W H out.bin        raw storage of a W x H RGBA image (texel (x,y) = x | y<<16) in TiledBuffer and
//                                     MortonBuffer (Base/Buffer.h:141-213): u32 innerW, innerH, data... twice
//   ref_kat sample in.bin out.bin     BaseSampler::textureImpl through Sampler2DSoft / SamplerCubeSoft
//                                     (SamplerSoft.h:118-168,289-373) for n coordinates
//   ref_kat bary in.bin out.bin       RendererSoft::barycentric + z / 1/w interpolation exactly as
//                                     rasterizationPixelQuad calls them (RendererSoft.cpp:741-744,771-797,1021-1056)
//
// `#define private public` only widens access for this harness (barycentric is a private member without state); it does
// not change the reference's code or object layout.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define private public
#include "Render/Software/RendererSoft.h"
#undef private
#include "Render/Software/SamplerSoft.h"

using namespace SoftGL;

static std::vector<uint8_t> readFile(const char *path) {
  std::vector<uint8_t> v;
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize(n);
  if (n && fread(v.data(), 1, n, f) != (size_t) n) exit(2);
  fclose(f);
  return v;
}

template<typename B>
static void dumpLayout(FILE *f, int w, int h) {
  B buf;
  buf.create(w, h);
  buf.clear();
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t v = (uint32_t) x | ((uint32_t) y << 16);
      RGBA px;
      memcpy(&px, &v, 4);
      buf.set(x, y, px);
    }
  uint32_t hdr[2];
  size_t texels = buf.getRawDataSize();
  // innerWidth_/innerHeight_ are protected; derive the padded size from the tile size of the class
  hdr[0] = (uint32_t) texels;
  hdr[1] = (uint32_t) buf.getLayout();
  fwrite(hdr, 4, 2, f);
  fwrite(buf.getRawDataPtr(), 4, texels, f);
}

static int cmdLayout(int w, int h, const char *out) {
  FILE *f = fopen(out, "wb");
  if (!f) return 2;
  dumpLayout<TiledBuffer<RGBA>>(f, w, h);
  dumpLayout<MortonBuffer<RGBA>>(f, w, h);
  fclose(f);
  return 0;
}

struct SampleHeader {
  int32_t w, h, layers, format, mips, filter, wrap, border, n, hasOffset;
};

static int cmdSample(const char *in, const char *out) {
  std::vector<uint8_t> blob = readFile(in);
  SampleHeader H;
  memcpy(&H, blob.data(), sizeof(H));
  const uint8_t *p = blob.data() + sizeof(H);
  RendererSoft renderer;
  TextureDesc desc;
  desc.width = H.w;
  desc.height = H.h;
  desc.type = H.layers == 6 ? TextureType_CUBE : TextureType_2D;
  desc.format = (TextureFormat) H.format;
  desc.usage = TextureUsage_Sampler | TextureUsage_UploadData;
  desc.useMipmaps = H.mips != 0;
  desc.multiSample = false;
  auto tex = renderer.createTexture(desc);
  SamplerDesc sd;
  sd.filterMin = (FilterMode) H.filter;
  sd.filterMag = Filter_LINEAR;
  sd.wrapS = sd.wrapT = sd.wrapR = (WrapMode) H.wrap;
  sd.borderColor = (BorderColor) H.border;
  tex->setSamplerDesc(sd);
  const size_t texels = (size_t) H.w * H.h;
  std::vector<uint32_t> result(H.n);
  const int comps = H.layers == 6 ? 3 : 2;
  if (H.format == TextureFormat_RGBA8) {
    std::vector<std::shared_ptr<Buffer<RGBA>>> bufs;
    for (int l = 0; l < H.layers; l++) {
      auto b = Buffer<RGBA>::makeDefault(H.w, H.h);
      for (int y = 0; y < H.h; y++)
        for (int x = 0; x < H.w; x++) {
          RGBA px;
          memcpy(&px, p + ((size_t) l * texels + (size_t) y * H.w + x) * 4, 4);
          b->set(x, y, px);
        }
      bufs.push_back(b);
    }
    tex->setImageData(bufs);
    p += texels * 4 * H.layers;
    const float *coords = (const float *) p;
    const float *lod = coords + (size_t) comps * H.n;
    const int32_t *offs = (const int32_t *) (lod + H.n);
    if (H.layers == 6) {
      SamplerCubeSoft<RGBA> s;
      s.setTexture(tex);
      for (int i = 0; i < H.n; i++) {
        RGBA c = s.textureCubeLod(glm::vec3(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]), lod[i]);
        memcpy(&result[i], &c, 4);
      }
    } else {
      Sampler2DSoft<RGBA> s;
      s.setTexture(tex);
      for (int i = 0; i < H.n; i++) {
        glm::ivec2 o = H.hasOffset ? glm::ivec2(offs[2 * i], offs[2 * i + 1]) : glm::ivec2(0);
        RGBA c = s.texture2DLodOffset(glm::vec2(coords[2 * i], coords[2 * i + 1]), lod[i], o);
        memcpy(&result[i], &c, 4);
      }
    }
  } else {
    std::vector<std::shared_ptr<Buffer<float>>> bufs;
    auto b = Buffer<float>::makeDefault(H.w, H.h);
    for (int y = 0; y < H.h; y++)
      for (int x = 0; x < H.w; x++) {
        float v;
        memcpy(&v, p + ((size_t) y * H.w + x) * 4, 4);
        b->set(x, y, v);
      }
    bufs.push_back(b);
    tex->setImageData(bufs);
    p += texels * 4;
    const float *coords = (const float *) p;
    const float *lod = coords + (size_t) comps * H.n;
    const int32_t *offs = (const int32_t *) (lod + H.n);
    Sampler2DSoft<float> s;
    s.setTexture(tex);
    for (int i = 0; i < H.n; i++) {
      glm::ivec2 o = H.hasOffset ? glm::ivec2(offs[2 * i], offs[2 * i + 1]) : glm::ivec2(0);
      float c = s.texture2DLodOffset(glm::vec2(coords[2 * i], coords[2 * i + 1]), lod[i], o);
      memcpy(&result[i], &c, 4);
    }
  }
  FILE *f = fopen(out, "wb");
  if (!f) return 2;
  fwrite(result.data(), 4, result.size(), f);
  fclose(f);
  return 0;
}

// in: i32 nTri, i32 nPerTri; per triangle 12 floats (3 x fragPos xyzw) then nPerTri x 2 floats sample positions.
// out per sample: i32 inside, f32 bc[3], f32 z, f32 w   (z, w = interpolateBarycentric over fragPos.z/.w of the
// three vertices with the un-corrected barycentrics, RendererSoft.cpp:794; for samples outside: bc as computed, z = w = 0)
static int cmdBary(const char *in, const char *out) {
  std::vector<uint8_t> blob = readFile(in);
  int32_t nTri, nPer;
  memcpy(&nTri, blob.data(), 4);
  memcpy(&nPer, blob.data() + 4, 4);
  const float *p = (const float *) (blob.data() + 8);
  RendererSoft renderer;
  FILE *f = fopen(out, "wb");
  if (!f) return 2;
  for (int t = 0; t < nTri; t++) {
    glm::aligned_vec4 vertPos[3];
    for (int v = 0; v < 3; v++) vertPos[v] = glm::aligned_vec4(p[4 * v], p[4 * v + 1], p[4 * v + 2], p[4 * v + 3]);
    p += 12;
    glm::aligned_vec4 flat[4];
    flat[0] = {vertPos[2].x, vertPos[1].x, vertPos[0].x, 0.f};
    flat[1] = {vertPos[2].y, vertPos[1].y, vertPos[0].y, 0.f};
    flat[2] = {vertPos[0].z, vertPos[1].z, vertPos[2].z, 0.f};
    flat[3] = {vertPos[0].w, vertPos[1].w, vertPos[2].w, 0.f};
    const float *vertZ[3] = {&vertPos[0].z, &vertPos[1].z, &vertPos[2].z};
    for (int s = 0; s < nPer; s++) {
      glm::aligned_vec4 pos(p[0], p[1], 0.f, 0.f);
      p += 2;
      glm::aligned_vec4 bc(0.f);
      bool inside = renderer.barycentric(flat, vertPos[0], pos, bc);
      float zw[2] = {0.f, 0.f};
      if (inside) {
        glm::aligned_vec4 tmp = pos;
        renderer.interpolateBarycentric(&tmp.z, vertZ, 2, bc);
        zw[0] = tmp.z;
        zw[1] = tmp.w;
      }
      int32_t in32 = inside ? 1 : 0;
      fwrite(&in32, 4, 1, f);
      fwrite(&bc.x, 4, 3, f);
      fwrite(zw, 4, 2, f);
    }
  }
  fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 5 && !strcmp(argv[1], "layout")) return cmdLayout(atoi(argv[2]), atoi(argv[3]), argv[4]);
  if (argc >= 4 && !strcmp(argv[1], "sample")) return cmdSample(argv[2], argv[3]);
  if (argc >= 4 && !strcmp(argv[1], "bary")) return cmdBary(argv[2], argv[3]);
  fprintf(stderr, "usage: ref_kat layout W H out | sample in out | bary in out\n");
  return 1;
}
