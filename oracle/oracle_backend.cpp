// ORACLE -- test infrastructure, NOT product code.  Trace-player backend for the CPU restatement.
#include "trace_player.h"

#include <cstdio>
#include <cstring>

using namespace SoftGL;

namespace PlayerBackend {

const char *name() { return "oracle(restatement)"; }
std::shared_ptr<Renderer> createRenderer() { return createRendererOracle(); }
bool loadShaders(ShaderProgram &program, int shading) { return oracleLoadShaders(program, shading); }
int nativeHandle(Texture &) { return -1; }

bool readback(Texture &tex, int layer, int level, int kind, Blob &out) {
  auto *t = dynamic_cast<TextureOracle *>(&tex);
  if (!t || t->levels.empty()) return false;
  out.format = tex.format;
  out.width = std::max(1, tex.width >> level);
  out.height = std::max(1, tex.height >> level);
  const std::vector<uint32_t> *src;
  if (kind == 1) {
    if (t->resolved.empty()) return false;
    src = &t->resolved;
    out.samples = 1;
  } else {
    src = &t->levels[layer][level];
    out.samples = t->samples();
  }
  out.data.resize(src->size() * 4);
  memcpy(out.data.data(), src->data(), out.data.size());
  return true;
}

// raw .tex cache format of TextureSoft::loadFromFile / storeToFile (TextureSoft.h:166-215): layers x levels
bool loadRaw(Texture &tex, const char *path) {
  auto *t = dynamic_cast<TextureOracle *>(&tex);
  if (!t) return false;
  if (t->levels.empty()) t->initImageData();
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  bool ok = true;
  for (auto &layer : t->levels)
    for (auto &lv : layer)
      if (fread(lv.data(), 4, lv.size(), f) != lv.size()) ok = false;
  fclose(f);
  return ok;
}

bool storeRaw(Texture &tex, const char *path) {
  auto *t = dynamic_cast<TextureOracle *>(&tex);
  if (!t || t->levels.empty()) return false;
  FILE *f = fopen(path, "wb");
  if (!f) return false;
  for (auto &layer : t->levels)
    for (auto &lv : layer) fwrite(lv.data(), 4, lv.size(), f);
  fclose(f);
  return true;
}

}  // namespace PlayerBackend
