// Headless harness around the REFERENCE's own Viewer (src/Viewer/Viewer.cpp, ModelLoader.cpp, Environment.cpp, QuadFilter.cpp,
// assimp) with either backend behind it:
//     --renderer soft   ViewerSoftware's role (RendererSoft + the software shaders; ViewerSoftware.h itself pulls in GL)
//     --renderer cuda   ViewerCUDA (softglrender_b200/host/Viewer/ViewerCUDA.h) = RendererCUDA over the C ABI
// It loads a bundled asset with the reference's ModelLoader, draws frames exactly as ViewerManager::drawFrame does
// (ViewerManager.h:95-119 minus window / imgui / orbit controller) and writes the attachments in the trace players' output
// format, plus one JSON line of submission counters.  Test infrastructure / integration proof -- not part of the library.
//
//   viewer_headless --renderer cuda|soft --model Cube --width 1000 --height 800 [--skybox Room] [--aa none|msaa|fxaa]
//                   [--reverse-z] [--ibl] [--blinnphong] [--frames N] [--out file] [--assets dir]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "Viewer/Camera.h"
#include "Viewer/Config.h"
#include "Viewer/ModelLoader.h"
#include "Viewer/Viewer.h"
#include "Render/Software/RendererSoft.h"
#include "Render/Software/TextureSoft.h"
#include "Viewer/Shader/Software/ShaderSoft.h"
#ifdef WITH_CUDA_BACKEND
#include "Viewer/ViewerCUDA.h"
#include "sglcuda.h"
#endif

using namespace SoftGL;
using namespace SoftGL::View;

namespace {

// ViewerSoftware without its GL upload (ViewerSoftware.h:22-69)
class ViewerSoftHeadless : public Viewer {
 public:
  ViewerSoftHeadless(Config &config, Camera &camera) : Viewer(config, camera) {}
  void configRenderer() override {
    camera_->setReverseZ(config_.reverseZ);
    cameraDepth_->setReverseZ(config_.reverseZ);
  }
  int swapBuffer() override { return outTexId_; }
  std::shared_ptr<Renderer> createRenderer() override {
    auto r = std::make_shared<RendererSoft>();
    return r->create() ? r : nullptr;
  }
#define CASE_SOFT(shading, source) case shading: return soft->SetShaders(std::make_shared<source::VS>(), std::make_shared<source::FS>())
  bool loadShaders(ShaderProgram &program, ShadingModel shading) override {
    auto *soft = dynamic_cast<ShaderProgramSoft *>(&program);
    switch (shading) {
      CASE_SOFT(Shading_BaseColor, ShaderBasic);
      CASE_SOFT(Shading_BlinnPhong, ShaderBlinnPhong);
      CASE_SOFT(Shading_PBR, ShaderPbrIBL);
      CASE_SOFT(Shading_Skybox, ShaderSkybox);
      CASE_SOFT(Shading_FXAA, ShaderFXAA);
      CASE_SOFT(Shading_IBL_Irradiance, ShaderIBLIrradiance);
      CASE_SOFT(Shading_IBL_Prefilter, ShaderIBLPrefilter);
      default: break;
    }
    return false;
  }
  Texture *colorTexture() { return texColorMain_.get(); }
  Texture *depthTexture() { return texDepthMain_.get(); }
  Texture *shadowTexture() { return texDepthShadow_.get(); }
};

void writeRecord(FILE *f, const std::string &tag, int w, int h, int format, int samples, const void *data, size_t bytes) {
  uint32_t n = (uint32_t) tag.size();
  fwrite(&n, 4, 1, f);
  fwrite(tag.data(), 1, n, f);
  int32_t hdr[4] = {w, h, format, samples};
  fwrite(hdr, 4, 4, f);
  uint32_t nb = (uint32_t) bytes;
  fwrite(&nb, 4, 1, f);
  fwrite(data, 1, bytes, f);
}

template<typename T>
void dumpSoft(FILE *f, const std::string &tag, Texture *tex) {
  auto *t = dynamic_cast<TextureSoft<T> *>(tex);
  if (!t) return;
  auto &img = t->getImage().getBuffer();
  if (img->multiSample) {
    writeRecord(f, tag + ".ms", img->width, img->height, tex->format, img->sampleCnt, img->bufferMs4x->getRawDataPtr(), img->bufferMs4x->getRawDataBytesSize());
    if (img->buffer) writeRecord(f, tag, img->width, img->height, tex->format, 1, img->buffer->getRawDataPtr(), img->buffer->getRawDataBytesSize());
  } else {
    writeRecord(f, tag, img->width, img->height, tex->format, 1, img->buffer->getRawDataPtr(), img->buffer->getRawDataBytesSize());
  }
}

#ifdef WITH_CUDA_BACKEND
void dumpCuda(FILE *f, const std::string &tag, Texture *tex) {
  auto *t = dynamic_cast<TextureCUDA *>(tex);
  if (!t) return;
  std::vector<uint8_t> px;
  int w = 0, h = 0;
  if (t->multiSample) {
    if (t->readPixels(0, 0, 0, px, w, h)) writeRecord(f, tag + ".ms", w, h, tex->format, 4, px.data(), px.size());
    if (tex->format == TextureFormat_RGBA8 && t->readPixels(0, 0, 1, px, w, h)) writeRecord(f, tag, w, h, tex->format, 1, px.data(), px.size());
  } else if (t->readPixels(0, 0, 0, px, w, h)) {
    writeRecord(f, tag, w, h, tex->format, 1, px.data(), px.size());
  }
}
#endif

}  // namespace

int main(int argc, char **argv) {
  std::string renderer = "cuda", model = "Cube", skybox, aa = "none", out, assets = "./assets/";
  int width = 1000, height = 800, frames = 1;
  bool reverseZ = false, ibl = false, blinnphong = false;
  for (int i = 1; i < argc; i++) {
    auto is = [&](const char *k) { return !strcmp(argv[i], k); };
    if (is("--renderer") && i + 1 < argc) renderer = argv[++i];
    else if (is("--model") && i + 1 < argc) model = argv[++i];
    else if (is("--skybox") && i + 1 < argc) skybox = argv[++i];
    else if (is("--aa") && i + 1 < argc) aa = argv[++i];
    else if (is("--out") && i + 1 < argc) out = argv[++i];
    else if (is("--assets") && i + 1 < argc) assets = argv[++i];
    else if (is("--width") && i + 1 < argc) width = atoi(argv[++i]);
    else if (is("--height") && i + 1 < argc) height = atoi(argv[++i]);
    else if (is("--frames") && i + 1 < argc) frames = atoi(argv[++i]);
    else if (is("--reverse-z")) reverseZ = true;
    else if (is("--ibl")) ibl = true;
    else if (is("--blinnphong")) blinnphong = true;
    else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
  }
  if (assets.back() != '/') assets += '/';

  // what ViewerManager::create sets up (ViewerManager.h:27-59), without the window
  auto camera = std::make_shared<Camera>();
  camera->setPerspective(glm::radians(CAMERA_FOV), (float) width / (float) height, CAMERA_NEAR, CAMERA_FAR);
  camera->lookAt(glm::vec3(-1.5f, 3.f, 3.f), glm::vec3(0.f, 1.f, 0.f), glm::vec3(0.f, 1.f, 0.f));   // OrbitController.cpp:14-16
  camera->update();
  Config config;
  config.aaType = aa == "msaa" ? AAType_MSAA : (aa == "fxaa" ? AAType_FXAA : AAType_NONE);
  config.reverseZ = reverseZ;
  config.showSkybox = !skybox.empty();
  config.pbrIbl = ibl;
  ModelLoader loader(config);
  // asset paths as ConfigPanel::loadConfig resolves them (ConfigPanel.cpp:241-290): ASSETS_DIR + the assets.json entry
  static const char *models[][2] = {{"Cube", "Cube/Cube.gltf"}, {"DamagedHelmet", "DamagedHelmet/DamagedHelmet.gltf"}, {"BoomBox", "BoomBox/BoomBox.gltf"},
                                    {"GlassTable", "GlassTable/scene.gltf"}, {"Robot", "Robot/scene.gltf"},
                                    {"AfricanHead", "AfricanHead/african_head.obj"}, {"Brickwall", "Brickwall/brickwall.obj"}};
  std::string modelPath;
  for (auto &m : models) if (model == m[0]) modelPath = assets + m[1];
  if (modelPath.empty() || !loader.loadModel(modelPath)) { fprintf(stderr, "cannot load model %s\n", model.c_str()); return 1; }
  if (!skybox.empty()) {
    std::string p = skybox == "Room" ? assets + "Skybox/Room.jpeg" : assets + "Skybox/" + skybox + "/";
    if (!loader.loadSkybox(p)) { fprintf(stderr, "cannot load skybox %s\n", skybox.c_str()); return 1; }
  }
  if (blinnphong) {   // glTF assets load as Shading_PBR (ModelLoader.cpp:326-332); BASELINE config 1 wants Blinn-Phong
    std::function<void(ModelNode &)> walk = [&](ModelNode &n) {
      for (auto &m : n.meshes) m.material->shadingModel = Shading_BlinnPhong;
      for (auto &c : n.children) walk(c);
    };
    walk(loader.getScene().model->rootNode);
  }

  std::shared_ptr<Viewer> viewer;
  ViewerSoftHeadless *soft = nullptr;
#ifdef WITH_CUDA_BACKEND
  ViewerCUDA *cuda = nullptr;
  if (renderer == "cuda") {
    auto v = std::make_shared<ViewerCUDA>(config, *camera);
    cuda = v.get();
    viewer = v;
  }
#endif
  if (renderer == "soft") {
    auto v = std::make_shared<ViewerSoftHeadless>(config, *camera);
    soft = v.get();
    viewer = v;
  }
  if (!viewer) { fprintf(stderr, "renderer %s not available in this build\n", renderer.c_str()); return 2; }
  if (!viewer->create(width, height, 0)) { fprintf(stderr, "Viewer::create failed\n"); return 1; }

  {  // ConfigPanel::update (ConfigPanel.cpp:216-224, default angle ConfigPanel.h:68) + the update-light callback
     // (ViewerManager.h:86-91): the point light sits at 2 * (sin 235 deg, 1.2, cos 235 deg)
    const float angle = glm::radians(235.f);
    config.pointLightPosition = 2.f * glm::vec3(glm::sin(angle), 1.2f, glm::cos(angle));
    auto &scene = loader.getScene();
    scene.pointLight.vertexes[0].a_position = config.pointLightPosition;
    scene.pointLight.UpdateVertexes();
    scene.pointLight.material->baseColor = glm::vec4(config.pointLightColor, 1.f);
  }
  double ms = 0.0;
  for (int f = 0; f < frames; f++) {   // ViewerManager::drawFrame (ViewerManager.h:95-119)
#ifdef WITH_CUDA_BACKEND
    if (cuda && f == frames - 1) { sgl_wait_idle(); sgl_reset_counters(); }
#endif
    auto t0 = std::chrono::steady_clock::now();
    viewer->configRenderer();
    viewer->drawFrame(loader.getScene());
    viewer->swapBuffer();
    viewer->waitRenderIdle();
    ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }

  if (!out.empty()) {
    FILE *f = fopen(out.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    if (soft) {
      dumpSoft<RGBA>(f, "color", soft->colorTexture());
      dumpSoft<float>(f, "depth", soft->depthTexture());
      if (soft->shadowTexture()) dumpSoft<float>(f, "shadow", soft->shadowTexture());
    }
#ifdef WITH_CUDA_BACKEND
    if (cuda) {
      dumpCuda(f, "color", cuda->colorTexture());
      dumpCuda(f, "depth", cuda->depthTexture());
      if (cuda->shadowTexture()) dumpCuda(f, "shadow", cuda->shadowTexture());
    }
#endif
    fclose(f);
  }
  unsigned long long passes = 0, draws = 0, verts = 0, idx = 0, prims = 0, launches = 0;
#ifdef WITH_CUDA_BACKEND
  if (cuda) {
    SglCounters c;
    if (sgl_get_counters(&c) == SGL_OK) {
      passes = c.passes; draws = c.draws; verts = c.vertices_in; idx = c.indices_in; prims = c.primitives_in; launches = c.kernel_launches;
    }
  }
#endif
  printf("{\"renderer\": \"%s\", \"model\": \"%s\", \"width\": %d, \"height\": %d, \"frames\": %d, \"last_frame_ms\": %.3f, "
         "\"last_frame\": {\"passes\": %llu, \"draws\": %llu, \"vertices_in\": %llu, \"indices_in\": %llu, \"primitives_in\": %llu, \"kernel_launches\": %llu}}\n",
         renderer.c_str(), model.c_str(), width, height, frames, ms, passes, draws, verts, idx, prims, launches);
  viewer->destroy();
  return 0;
}
