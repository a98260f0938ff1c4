// PlayerBackend for RendererCUDA: what a ViewerCUDA subclass provides around the renderer -- createRenderer,
// loadShaders (Viewer.h:29,38-39) and read-back of attachments (swapBuffer role, ViewerSoftware.h:30-44).
#include "trace_player.h"

#include <cstdlib>

using namespace SoftGL;

namespace PlayerBackend {

const char *name() { return "RendererCUDA"; }

std::shared_ptr<Renderer> createRenderer() {
  auto r = std::make_shared<RendererCUDA>();
  const char *dev = getenv("SGL_DEVICE"), *rank = getenv("SGL_RANK"), *world = getenv("SGL_WORLD"), *layout = getenv("SGL_TEXTURE_LAYOUT");
  r->setDevice(dev ? atoi(dev) : 0, rank ? atoi(rank) : 0, world ? atoi(world) : 1);
  if (layout) r->setTextureLayout(atoi(layout));
  if (!r->create()) return nullptr;
  return r;
}

bool loadShaders(ShaderProgram &program, int shading) {
  auto *p = dynamic_cast<ShaderProgramCUDA *>(&program);
  return p && p->setShadingModel(shading);
}

bool readback(Texture &tex, int layer, int level, int kind, Blob &out) {
  auto *t = dynamic_cast<TextureCUDA *>(&tex);
  if (!t) return false;
  out.format = tex.format;
  out.samples = (tex.multiSample && kind == 0) ? 4 : 1;
  return t->readPixels((uint32_t) layer, (uint32_t) level, kind, out.data, out.width, out.height);
}

int nativeHandle(Texture &tex) {
  auto *t = dynamic_cast<TextureCUDA *>(&tex);
  return t ? t->handle() : -1;
}

bool loadRaw(Texture &tex, const char *path) {
  auto *t = dynamic_cast<TextureCUDA *>(&tex);
  return t && t->loadFromFile(path);
}

bool storeRaw(Texture &tex, const char *path) {
  auto *t = dynamic_cast<TextureCUDA *>(&tex);
  return t && t->storeToFile(path);
}

}  // namespace PlayerBackend
