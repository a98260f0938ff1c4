// RendererCUDA: the B200 backend behind SoftGLRender's abstract Renderer interface -- the drop-in for
// RendererSoft (src/Render/Software/RendererSoft.h:18-132).  Every object owns handles into libsglcuda.so and
// talks to it exclusively through the C ABI of include/sglcuda.h.
//
// Execution model: calls between beginRenderPass and endRenderPass are recorded (with snapshots of uniform
// bytes, sampler bindings and render states taken at draw() time, exactly the data RendererSoft reads at that
// moment) and executed on the GPU at endRenderPass; results are complete at waitIdle() or at any read-back.
#pragma once
#include <string>
#include <vector>
#include "Render/Renderer.h"

namespace SoftGL {

// RendererType of this backend.  The reference's enum (Render/Renderer.h:18-22) ends at Renderer_Vulkan; a tree that adds a
// `Renderer_CUDA` enumerator there defines SGL_HAVE_RENDERER_CUDA_ENUM, otherwise the next free
// value is used, so this header compiles against the UNMODIFIED reference headers (tests/test_boundary_compiles.py).
#ifndef SGL_HAVE_RENDERER_CUDA_ENUM
static const RendererType Renderer_CUDA = static_cast<RendererType>(Renderer_Vulkan + 1);
#endif

// View::ShadingModel values (src/Viewer/Material.h:25-34) == SGL_SHADER_* ids
enum ShaderKindCUDA {
  ShaderCUDA_None = 0, ShaderCUDA_BaseColor = 1, ShaderCUDA_BlinnPhong = 2, ShaderCUDA_PBR = 3, ShaderCUDA_Skybox = 4,
  ShaderCUDA_IBLIrradiance = 5, ShaderCUDA_IBLPrefilter = 6, ShaderCUDA_FXAA = 7
};

class TextureCUDA : public Texture {
 public:
  explicit TextureCUDA(const TextureDesc &desc, int layout);
  ~TextureCUDA() override;
  int getId() const override { return id_; }
  void setSamplerDesc(SamplerDesc &sampler) override { samplerDesc_ = sampler; }
  void initImageData() override;
  void setImageData(const std::vector<std::shared_ptr<Buffer<RGBA>>> &buffers) override;
  void setImageData(const std::vector<std::shared_ptr<Buffer<float>>> &buffers) override;
  void dumpImage(const char *path, uint32_t layer, uint32_t level) override;

  const SamplerDesc &getSamplerDesc() const { return samplerDesc_; }
  int handle() const { return handle_; }
  int levelCount() const;
  int layerCount() const { return type == TextureType_CUBE ? 6 : 1; }
  // raw .tex cache format of TextureSoft::loadFromFile/storeToFile (TextureSoft.h:166-215): layers x levels, linear
  bool loadFromFile(const char *path);
  bool storeToFile(const char *path);
  // kind 0: attachment texels ([y][x][sample]); kind 1: resolved colour of a multisample texture
  bool readPixels(uint32_t layer, uint32_t level, int kind, std::vector<uint8_t> &out, int &w, int &h);

 private:
  bool ensureAllocated();
  template<typename T> void upload(const std::vector<std::shared_ptr<Buffer<T>>> &buffers);
  int id_;
  int handle_ = 0;
  int layout_ = 0;
  SamplerDesc samplerDesc_;
};

class FrameBufferCUDA : public FrameBuffer {
 public:
  explicit FrameBufferCUDA(bool offscreen);
  int getId() const override { return id_; }
  bool isValid() override { return colorReady_ || depthReady_; }

 private:
  int id_;
};

class VertexArrayObjectCUDA : public VertexArrayObject {
 public:
  explicit VertexArrayObjectCUDA(const VertexArray &va);
  ~VertexArrayObjectCUDA();   // not an override: the reference's VertexArrayObject has no virtual destructor (Vertex.h:15-19)
  int getId() const override { return id_; }
  void updateVertexData(void *data, size_t length) override;
  int vertexBuffer = 0, indexBuffer = 0;
  size_t vertexCount = 0, indexCount = 0, vertexStride = 0;

 private:
  int id_;
};

class ShaderProgramCUDA : public ShaderProgram {
 public:
  ShaderProgramCUDA();
  int getId() const override { return id_; }
  void addDefine(const std::string &def) override { defines_.push_back(def); }
  // role of ShaderProgramSoft::SetShaders (ShaderProgramSoft.h:26-59): select the device shader pair, resolve defines
  bool setShadingModel(int shading);
  int shader() const { return shader_; }
  uint32_t defineMask() const { return defineMask_; }
  int getUniformLocation(const std::string &name) const;      // blocks: byte offset; samplers: kSamplerBase + slot
  void bindUniformBlockBuffer(const void *data, size_t len, int location);
  void bindUniformSampler(int texture, int filterMin, int wrap, int border, int location);
  const std::vector<uint8_t> &uniformBytes() const { return uniforms_; }
  struct SamplerSlot { int texture = 0, filterMin = 0, wrap = 0, border = 0; };
  const SamplerSlot *samplerSlots() const { return slots_; }
  static const int kSamplerBase = 1 << 20;

 private:
  int id_;
  int shader_ = 0;
  uint32_t defineMask_ = 0;
  std::vector<std::string> defines_;
  std::vector<uint8_t> uniforms_;
  SamplerSlot slots_[8];
};

class UniformBlockCUDA : public UniformBlock {
 public:
  UniformBlockCUDA(const std::string &name, int size) : UniformBlock(name, size), bytes_((size_t) size) {}
  int getLocation(ShaderProgram &program) override;
  void bindProgram(ShaderProgram &program, int location) override;
  void setSubData(void *data, int len, int offset) override;
  void setData(void *data, int len) override { setSubData(data, len, 0); }

 private:
  std::vector<uint8_t> bytes_;
};

class UniformSamplerCUDA : public UniformSampler {
 public:
  UniformSamplerCUDA(const std::string &name, TextureType type, TextureFormat format) : UniformSampler(name, type, format) {}
  int getLocation(ShaderProgram &program) override;
  void bindProgram(ShaderProgram &program, int location) override;
  void setTexture(const std::shared_ptr<Texture> &tex) override;

 private:
  int texture_ = 0, filterMin_ = 0, wrap_ = 0, border_ = 0;
};

class RendererCUDA : public Renderer {
 public:
  RendererType type() override { return Renderer_CUDA; }
  bool create() override;
  void destroy() override;
  std::shared_ptr<FrameBuffer> createFrameBuffer(bool offscreen) override;
  std::shared_ptr<Texture> createTexture(const TextureDesc &desc) override;
  std::shared_ptr<VertexArrayObject> createVertexArrayObject(const VertexArray &vertexArray) override;
  std::shared_ptr<ShaderProgram> createShaderProgram() override;
  std::shared_ptr<PipelineStates> createPipelineStates(const RenderStates &renderStates) override;
  std::shared_ptr<UniformBlock> createUniformBlock(const std::string &name, int size) override;
  std::shared_ptr<UniformSampler> createUniformSampler(const std::string &name, const TextureDesc &desc) override;
  void beginRenderPass(std::shared_ptr<FrameBuffer> &frameBuffer, const ClearStates &states) override;
  void setViewPort(int x, int y, int width, int height) override;
  void setVertexArrayObject(std::shared_ptr<VertexArrayObject> &vao) override;
  void setShaderProgram(std::shared_ptr<ShaderProgram> &program) override;
  void setShaderResources(std::shared_ptr<ShaderResources> &resources) override;
  void setPipelineStates(std::shared_ptr<PipelineStates> &states) override;
  void draw() override;
  void endRenderPass() override;
  void waitIdle() override;

  // device / layout selection for objects created afterwards (SGL_LAYOUT_*; the reference picks at compile time)
  void setDevice(int ordinal, int rank = 0, int world = 1) { device_ = ordinal; rank_ = rank; world_ = world; }
  void setTextureLayout(int layout) { textureLayout_ = layout; }
  // Multi-GPU tile sharding: how far around the pixel it shades a later pass samples an attachment rendered earlier
  // (pixels; < 0 = anywhere, the attachment is then rendered whole on every rank -- the default).  A Viewer whose only
  // screen-space consumer is the FXAA filter declares 32 (FxaaSoft.h:73-74,169-210: 18.5 px + bilinear footprint).
  static void setSampledAttachmentHalo(int pixels);
  static int sampledAttachmentHalo();

 private:
  int device_ = 0, rank_ = 0, world_ = 1;
  int textureLayout_ = 0;
  bool passOpen_ = false;
  bool created_ = false;
  VertexArrayObjectCUDA *vao_ = nullptr;
  ShaderProgramCUDA *program_ = nullptr;
  const RenderStates *states_ = nullptr;
};

}  // namespace SoftGL
