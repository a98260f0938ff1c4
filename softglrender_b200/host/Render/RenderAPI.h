// The abstract Renderer boundary, re-declared from scratch so that RendererCUDA and the headless harness build
// without the reference tree.  Names, signatures, enum values and default values are those of the reference's
// Render/*.h (Renderer.h:18-59, Texture.h:17-105, Framebuffer.h:14-91, Vertex.h:15-37, Uniform.h:18-64,
// ShaderProgram.h:15-58, PipelineStates.h:13-22, RenderStates.h:13-99) -- that is what makes the backend a drop-in;
// a caller written against the reference headers (e.g. Viewer.cpp, or harness/trace_player.cpp) compiles against
// either set.  Inside the reference tree these declarations are NOT used: RendererCUDA then includes the
// reference's own headers.
#pragma once
#include <algorithm>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>
#include "Base/Buffer.h"
#include "Base/GLMInc.h"

namespace SoftGL {

// ---- Texture.h ---------------------------------------------------------------------------------------------------
enum WrapMode { Wrap_REPEAT, Wrap_MIRRORED_REPEAT, Wrap_CLAMP_TO_EDGE, Wrap_CLAMP_TO_BORDER };
enum FilterMode {
  Filter_NEAREST, Filter_LINEAR, Filter_NEAREST_MIPMAP_NEAREST, Filter_LINEAR_MIPMAP_NEAREST,
  Filter_NEAREST_MIPMAP_LINEAR, Filter_LINEAR_MIPMAP_LINEAR
};
enum CubeMapFace {
  TEXTURE_CUBE_MAP_POSITIVE_X = 0, TEXTURE_CUBE_MAP_NEGATIVE_X = 1, TEXTURE_CUBE_MAP_POSITIVE_Y = 2,
  TEXTURE_CUBE_MAP_NEGATIVE_Y = 3, TEXTURE_CUBE_MAP_POSITIVE_Z = 4, TEXTURE_CUBE_MAP_NEGATIVE_Z = 5
};
enum BorderColor { Border_BLACK = 0, Border_WHITE };
enum TextureType { TextureType_2D, TextureType_CUBE };
enum TextureFormat { TextureFormat_RGBA8 = 0, TextureFormat_FLOAT32 = 1 };
enum TextureUsage {
  TextureUsage_Sampler = 1 << 0, TextureUsage_UploadData = 1 << 1, TextureUsage_AttachmentColor = 1 << 2,
  TextureUsage_AttachmentDepth = 1 << 3, TextureUsage_RendererOutput = 1 << 4
};

struct SamplerDesc {
  FilterMode filterMin = Filter_NEAREST, filterMag = Filter_NEAREST;
  WrapMode wrapS = Wrap_CLAMP_TO_EDGE, wrapT = Wrap_CLAMP_TO_EDGE, wrapR = Wrap_CLAMP_TO_EDGE;
  BorderColor borderColor = Border_BLACK;
};

struct TextureDesc {
  int width = 0, height = 0;
  TextureType type = TextureType_2D;
  TextureFormat format = TextureFormat_RGBA8;
  uint32_t usage = TextureUsage_Sampler;
  bool useMipmaps = false, multiSample = false;
  std::string tag;
};

class Texture : public TextureDesc {
 public:
  virtual ~Texture() = default;
  uint32_t getLevelWidth(uint32_t level) { return (uint32_t) std::max(1, width >> level); }
  uint32_t getLevelHeight(uint32_t level) { return (uint32_t) std::max(1, height >> level); }
  virtual int getId() const = 0;
  virtual void setSamplerDesc(SamplerDesc &sampler) {}
  virtual void initImageData() {}
  virtual void setImageData(const std::vector<std::shared_ptr<Buffer<RGBA>>> &buffers) {}
  virtual void setImageData(const std::vector<std::shared_ptr<Buffer<float>>> &buffers) {}
  virtual void dumpImage(const char *path, uint32_t layer, uint32_t level) = 0;
};

// ---- Framebuffer.h -------------------------------------------------------------------------------------------------
struct FrameBufferAttachment {
  std::shared_ptr<Texture> tex = nullptr;
  uint32_t layer = 0, level = 0;
};

class FrameBuffer {
 public:
  explicit FrameBuffer(bool offscreen) : offscreen_(offscreen) {}
  virtual int getId() const = 0;
  virtual bool isValid() = 0;

  virtual void setColorAttachment(std::shared_ptr<Texture> &color, int level) {
    colorAttachment_.tex = color; colorAttachment_.layer = 0; colorAttachment_.level = (uint32_t) level; colorReady_ = true;
  }
  virtual void setColorAttachment(std::shared_ptr<Texture> &color, CubeMapFace face, int level) {
    colorAttachment_.tex = color; colorAttachment_.layer = (uint32_t) face; colorAttachment_.level = (uint32_t) level; colorReady_ = true;
  }
  virtual void setDepthAttachment(std::shared_ptr<Texture> &depth) {
    depthAttachment_.tex = depth; depthAttachment_.layer = 0; depthAttachment_.level = 0; depthReady_ = true;
  }
  const FrameBufferAttachment &getColorAttachment() const { return colorAttachment_; }
  const FrameBufferAttachment &getDepthAttachment() const { return depthAttachment_; }
  bool isColorReady() const { return colorReady_; }
  bool isDepthReady() const { return depthReady_; }
  bool isMultiSample() const {
    if (colorReady_) return colorAttachment_.tex->multiSample;
    if (depthReady_) return depthAttachment_.tex->multiSample;
    return false;
  }
  bool isOffscreen() const { return offscreen_; }
  void setOffscreen(bool offscreen) { offscreen_ = offscreen; }

 protected:
  bool offscreen_ = false, colorReady_ = false, depthReady_ = false;
  FrameBufferAttachment colorAttachment_{}, depthAttachment_{};
};

// ---- Vertex.h ---------------------------------------------------------------------------------------------------------
class VertexArrayObject {
 public:
  virtual int getId() const = 0;
  virtual void updateVertexData(void *data, size_t length) = 0;
};

struct VertexAttributeDesc {
  size_t size, stride, offset;
};

struct VertexArray {
  size_t vertexSize = 0;
  std::vector<VertexAttributeDesc> vertexesDesc;
  uint8_t *vertexesBuffer = nullptr;
  size_t vertexesBufferLength = 0;
  int32_t *indexBuffer = nullptr;
  size_t indexBufferLength = 0;
};

// ---- Uniform.h ----------------------------------------------------------------------------------------------------------
class ShaderProgram;

class Uniform {
 public:
  explicit Uniform(std::string n) : name(std::move(n)), hash_(nextHash()++) {}
  int getHash() const { return hash_; }
  virtual int getLocation(ShaderProgram &program) = 0;
  virtual void bindProgram(ShaderProgram &program, int location) = 0;
  std::string name;

 private:
  static int &nextHash() { static int h = 0; return h; }
  int hash_;
};

class UniformBlock : public Uniform {
 public:
  UniformBlock(const std::string &n, int size) : Uniform(n), blockSize(size) {}
  virtual void setSubData(void *data, int len, int offset) = 0;
  virtual void setData(void *data, int len) = 0;

 protected:
  int blockSize;
};

class UniformSampler : public Uniform {
 public:
  UniformSampler(const std::string &n, TextureType t, TextureFormat f) : Uniform(n), type(t), format(f) {}
  virtual void setTexture(const std::shared_ptr<Texture> &tex) = 0;

 protected:
  TextureType type;
  TextureFormat format;
};

class ShaderResources {
 public:
  std::unordered_map<int, std::shared_ptr<UniformBlock>> blocks;
  std::unordered_map<int, std::shared_ptr<UniformSampler>> samplers;
};

// ---- ShaderProgram.h -------------------------------------------------------------------------------------------------------
class ShaderProgram {
 public:
  virtual int getId() const = 0;
  virtual void addDefine(const std::string &def) = 0;
  virtual void addDefines(const std::set<std::string> &defs) {
    for (auto &d : defs) addDefine(d);
  }
  virtual void bindResources(ShaderResources &resources) {
    for (auto &kv : resources.blocks) bindUniform(*kv.second);
    for (auto &kv : resources.samplers) bindUniform(*kv.second);
  }

 protected:
  virtual bool bindUniform(Uniform &uniform) {
    int location;
    auto it = uniformLocations_.find(uniform.getHash());
    if (it == uniformLocations_.end()) {
      location = uniform.getLocation(*this);
      uniformLocations_[uniform.getHash()] = location;
    } else {
      location = it->second;
    }
    if (location < 0) return false;
    uniform.bindProgram(*this, location);
    return true;
  }
  std::unordered_map<int, int> uniformLocations_;
};

// ---- RenderStates.h / PipelineStates.h ---------------------------------------------------------------------------------------
enum DepthFunction {
  DepthFunc_NEVER, DepthFunc_LESS, DepthFunc_EQUAL, DepthFunc_LEQUAL, DepthFunc_GREATER, DepthFunc_NOTEQUAL,
  DepthFunc_GEQUAL, DepthFunc_ALWAYS
};
enum BlendFactor {
  BlendFactor_ZERO, BlendFactor_ONE, BlendFactor_SRC_COLOR, BlendFactor_SRC_ALPHA, BlendFactor_DST_COLOR,
  BlendFactor_DST_ALPHA, BlendFactor_ONE_MINUS_SRC_COLOR, BlendFactor_ONE_MINUS_SRC_ALPHA,
  BlendFactor_ONE_MINUS_DST_COLOR, BlendFactor_ONE_MINUS_DST_ALPHA
};
enum BlendFunction { BlendFunc_ADD, BlendFunc_SUBTRACT, BlendFunc_REVERSE_SUBTRACT, BlendFunc_MIN, BlendFunc_MAX };
enum PolygonMode { PolygonMode_POINT, PolygonMode_LINE, PolygonMode_FILL };
enum PrimitiveType { Primitive_POINT, Primitive_LINE, Primitive_TRIANGLE };

struct BlendParameters {
  BlendFunction blendFuncRgb = BlendFunc_ADD;
  BlendFactor blendSrcRgb = BlendFactor_ONE, blendDstRgb = BlendFactor_ZERO;
  BlendFunction blendFuncAlpha = BlendFunc_ADD;
  BlendFactor blendSrcAlpha = BlendFactor_ONE, blendDstAlpha = BlendFactor_ZERO;
  void SetBlendFactor(BlendFactor src, BlendFactor dst) {
    blendSrcRgb = blendSrcAlpha = src;
    blendDstRgb = blendDstAlpha = dst;
  }
  void SetBlendFunc(BlendFunction func) { blendFuncRgb = blendFuncAlpha = func; }
};

struct RenderStates {
  bool blend = false;
  BlendParameters blendParams;
  bool depthTest = false, depthMask = true;
  DepthFunction depthFunc = DepthFunc_LESS;
  bool cullFace = false;
  PrimitiveType primitiveType = Primitive_TRIANGLE;
  PolygonMode polygonMode = PolygonMode_FILL;
  float lineWidth = 1.f;
};

struct ClearStates {
  bool depthFlag = false, colorFlag = false;
  glm::vec4 clearColor = glm::vec4(0.f);
  float clearDepth = 1.f;
};

class PipelineStates {
 public:
  explicit PipelineStates(const RenderStates &states) : renderStates(states) {}
  virtual ~PipelineStates() = default;
  RenderStates renderStates;
};

// ---- Renderer.h --------------------------------------------------------------------------------------------------------------
enum RendererType { Renderer_SOFT, Renderer_OPENGL, Renderer_Vulkan };   // Renderer.h:18-22 (Renderer_CUDA: RendererCUDA.h)

class Renderer {
 public:
  virtual RendererType type() = 0;
  virtual bool create() { return true; }
  virtual void destroy() {}
  virtual std::shared_ptr<FrameBuffer> createFrameBuffer(bool offscreen) = 0;
  virtual std::shared_ptr<Texture> createTexture(const TextureDesc &desc) = 0;
  virtual std::shared_ptr<VertexArrayObject> createVertexArrayObject(const VertexArray &vertexArray) = 0;
  virtual std::shared_ptr<ShaderProgram> createShaderProgram() = 0;
  virtual std::shared_ptr<PipelineStates> createPipelineStates(const RenderStates &renderStates) = 0;
  virtual std::shared_ptr<UniformBlock> createUniformBlock(const std::string &name, int size) = 0;
  virtual std::shared_ptr<UniformSampler> createUniformSampler(const std::string &name, const TextureDesc &desc) = 0;
  virtual void beginRenderPass(std::shared_ptr<FrameBuffer> &frameBuffer, const ClearStates &states) = 0;
  virtual void setViewPort(int x, int y, int width, int height) = 0;
  virtual void setVertexArrayObject(std::shared_ptr<VertexArrayObject> &vao) = 0;
  virtual void setShaderProgram(std::shared_ptr<ShaderProgram> &program) = 0;
  virtual void setShaderResources(std::shared_ptr<ShaderResources> &uniforms) = 0;
  virtual void setPipelineStates(std::shared_ptr<PipelineStates> &states) = 0;
  virtual void draw() = 0;
  virtual void endRenderPass() = 0;
  virtual void waitIdle() = 0;
};

}  // namespace SoftGL
