// Headless offscreen harness: replays a Renderer-API trace on whatever backend implements
// SoftGL::Renderer.  The SAME source is compiled twice:
//   * against the reference's own headers + RendererSoft   (oracle/_ref/ref_player, the oracle)
//   * against softglrender_b200/host (mirrored interface) + RendererCUDA (the product)
// which is the drop-in claim of BASELINE.json's north_star in executable form.
// PLAYER_BACKEND_HEADER must declare the SoftGL Render API and namespace PlayerBackend.
#pragma once
#include PLAYER_BACKEND_HEADER

#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace PlayerBackend {
struct Blob {
  int width = 0, height = 0, format = 0, samples = 1;
  std::vector<uint8_t> data;
};
const char *name();
std::shared_ptr<SoftGL::Renderer> createRenderer();
// role of Viewer::loadShaders (src/Viewer/Viewer.h:39, ViewerSoftware.h:54-69); shading = View::ShadingModel value
bool loadShaders(SoftGL::ShaderProgram &program, int shading);
// kind 0: the attachment itself (per-sample data when multisampled); kind 1: resolved colour of an MS texture
bool readback(SoftGL::Texture &tex, int layer, int level, int kind, Blob &out);
bool loadRaw(SoftGL::Texture &tex, const char *path);
bool storeRaw(SoftGL::Texture &tex, const char *path);
int nativeHandle(SoftGL::Texture &tex);   // C-ABI handle for RendererCUDA, -1 elsewhere
}  // namespace PlayerBackend

class TracePlayer {
 public:
  ~TracePlayer();   // Viewer::destroy order (Viewer.cpp:59-88): waitIdle, release every resource, then Renderer::destroy
  bool load(const std::string &path);
  void setDataDir(const std::string &dir) { dataDir_ = dir; }
  void setOutput(const std::string &path) { outPath_ = path; }
  // executes [0, frameBegin) : resource creation + warm-up frame
  bool runSetup();
  // executes the FRAME section once (steady-state frame); ends with waitIdle when sync==true
  bool runFrame(bool sync);
  // executes everything after FRAME_END (read-backs)
  bool runTail();
  bool hasFrameSection() const { return frameBegin_ >= 0; }
  std::shared_ptr<SoftGL::Renderer> renderer() { return renderer_; }
  SoftGL::Texture *texture(int id) { return id >= 0 && id < (int) textures_.size() ? textures_[id].get() : nullptr; }
  // read-back by tag recorded in the trace tail (bench/e2e): returns false if the tag is unknown
  bool readbackTagged(const std::string &tag, PlayerBackend::Blob &out);
  // backend-specific integer handle of the texture read back under `tag` (-1 if unknown / not supported)
  int textureHandleTagged(const std::string &tag);

 private:
  struct Cmd {
    uint32_t op;
    const uint8_t *p;
    uint32_t n;
  };
  bool exec(const Cmd &c);
  bool execRange(int from, int to);
  void writeRecord(const std::string &tag, const PlayerBackend::Blob &b);

  std::vector<uint8_t> bytes_;
  std::vector<Cmd> cmds_;
  int frameBegin_ = -1, frameEnd_ = -1;
  std::string dataDir_ = ".", outPath_;
  FILE *out_ = nullptr;

  std::shared_ptr<SoftGL::Renderer> renderer_;
  std::vector<std::shared_ptr<SoftGL::Texture>> textures_;
  std::vector<std::shared_ptr<SoftGL::VertexArrayObject>> vaos_;
  std::vector<std::vector<uint8_t>> vaoVertexBytes_;
  std::vector<std::shared_ptr<SoftGL::ShaderProgram>> programs_;
  std::vector<std::shared_ptr<SoftGL::UniformBlock>> blocks_;
  std::vector<std::shared_ptr<SoftGL::UniformSampler>> samplers_;
  std::vector<std::shared_ptr<SoftGL::PipelineStates>> pipelines_;
  std::vector<std::shared_ptr<SoftGL::FrameBuffer>> fbos_;
  // ShaderResources objects are cached per (block ids, sampler ids) like MaterialObject::shaderResources
  std::map<std::vector<int>, std::shared_ptr<SoftGL::ShaderResources>> resources_;
};
