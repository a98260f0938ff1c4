// ORACLE (test infrastructure, not product code): PlayerBackend for the reference RendererSoft.
// Mirrors what ViewerSoftware does around the renderer (src/Viewer/ViewerSoftware.h:30-69):
// createRenderer -> RendererSoft, loadShaders -> the software shader classes, read-back of
// TextureSoft buffers.  Requires the default (linear) Buffer layout build.
#include "trace_player.h"
#include "Viewer/Shader/Software/ShaderSoft.h"

using namespace SoftGL;

// layout facts the CUDA side relies on
static_assert(sizeof(ShaderPbrIBL::ShaderAttributes) == 64, "vertex stride");
static_assert(sizeof(ShaderPbrIBL::ShaderVaryings) == 112, "pbr varyings");
static_assert(sizeof(ShaderBlinnPhong::ShaderVaryings) == 128, "blinn-phong varyings");
static_assert(sizeof(ShaderSkybox::ShaderVaryings) == 16, "skybox varyings");
static_assert(sizeof(ShaderFXAA::ShaderVaryings) == 8, "fxaa varyings");
static_assert(offsetof(ShaderPbrIBL::ShaderUniforms, u_modelViewProjectionMatrix) == 80, "mvp offset");
static_assert(offsetof(ShaderPbrIBL::ShaderUniforms, u_ambientColor) == 256, "scene block offset");
static_assert(offsetof(ShaderPbrIBL::ShaderUniforms, u_enableLight) == 320, "material block offset");
static_assert(offsetof(ShaderPbrIBL::ShaderUniforms, u_baseColor) == 352, "base colour offset");
static_assert(offsetof(ShaderPbrIBL::ShaderUniforms, u_albedoMap) == 368, "sampler offset");
static_assert(offsetof(ShaderBasic::ShaderUniforms, u_enableLight) == 256, "basic material offset");
static_assert(offsetof(ShaderFXAA::ShaderUniforms, u_screenTexture) == 8, "fxaa sampler offset");
static_assert(offsetof(ShaderIBLPrefilter::ShaderUniforms, u_srcResolution) == 256, "prefilter block offset");

namespace PlayerBackend {

const char *name() { return "RendererSoft(reference)"; }

std::shared_ptr<Renderer> createRenderer() {
  auto r = std::make_shared<RendererSoft>();
  if (!r->create()) return nullptr;
  return r;
}

template<typename S>
static bool setShaders(ShaderProgram &program) {
  auto *soft = dynamic_cast<ShaderProgramSoft *>(&program);
  return soft && soft->SetShaders(std::make_shared<typename S::VS>(), std::make_shared<typename S::FS>());
}

namespace {
struct NsBasic { using VS = ShaderBasic::VS; using FS = ShaderBasic::FS; };
struct NsBlinn { using VS = ShaderBlinnPhong::VS; using FS = ShaderBlinnPhong::FS; };
struct NsPbr { using VS = ShaderPbrIBL::VS; using FS = ShaderPbrIBL::FS; };
struct NsSky { using VS = ShaderSkybox::VS; using FS = ShaderSkybox::FS; };
struct NsFxaa { using VS = ShaderFXAA::VS; using FS = ShaderFXAA::FS; };
struct NsIrr { using VS = ShaderIBLIrradiance::VS; using FS = ShaderIBLIrradiance::FS; };
struct NsPre { using VS = ShaderIBLPrefilter::VS; using FS = ShaderIBLPrefilter::FS; };
}

bool loadShaders(ShaderProgram &program, int shading) {
  switch (shading) {   // View::ShadingModel values (src/Viewer/Material.h:25-34)
    case 1: return setShaders<NsBasic>(program);
    case 2: return setShaders<NsBlinn>(program);
    case 3: return setShaders<NsPbr>(program);
    case 4: return setShaders<NsSky>(program);
    case 5: return setShaders<NsIrr>(program);
    case 6: return setShaders<NsPre>(program);
    case 7: return setShaders<NsFxaa>(program);
    default: return false;
  }
}

template<typename T>
static bool readbackT(Texture &tex, int layer, int level, int kind, Blob &out) {
  auto *t = dynamic_cast<TextureSoft<T> *>(&tex);
  if (!t) return false;
  auto &img = t->getImage(layer).getBuffer(level);
  out.width = img->width;
  out.height = img->height;
  out.format = tex.format;
  if (img->multiSample && kind == 0) {
    out.samples = img->sampleCnt;
    auto *p = (const uint8_t *) img->bufferMs4x->getRawDataPtr();
    out.data.assign(p, p + img->bufferMs4x->getRawDataBytesSize());
    return true;
  }
  out.samples = 1;
  if (!img->buffer) return false;
  auto *p = (const uint8_t *) img->buffer->getRawDataPtr();
  out.data.assign(p, p + img->buffer->getRawDataBytesSize());
  return true;
}

bool readback(Texture &tex, int layer, int level, int kind, Blob &out) {
  if (tex.format == TextureFormat_RGBA8) return readbackT<RGBA>(tex, layer, level, kind, out);
  return readbackT<float>(tex, layer, level, kind, out);
}

int nativeHandle(Texture &) { return -1; }

bool loadRaw(Texture &tex, const char *path) {
  if (tex.format == TextureFormat_RGBA8) return dynamic_cast<TextureSoft<RGBA> *>(&tex)->loadFromFile(path);
  return dynamic_cast<TextureSoft<float> *>(&tex)->loadFromFile(path);
}

bool storeRaw(Texture &tex, const char *path) {
  if (tex.format == TextureFormat_RGBA8) dynamic_cast<TextureSoft<RGBA> *>(&tex)->storeToFile(path);
  else dynamic_cast<TextureSoft<float> *>(&tex)->storeToFile(path);
  return true;
}

}  // namespace PlayerBackend
