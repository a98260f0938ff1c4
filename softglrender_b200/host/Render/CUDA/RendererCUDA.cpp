#include "RendererCUDA.h"
#include <cstdlib>

#include <cstdio>
#include <cstring>
#include "sglcuda.h"

namespace SoftGL {

namespace {
int nextId() {
  static int id = 0;
  return id++;
}
// same convention as the reference: log and carry on (Base/Logger.h LOGE; RendererSoft.cpp:33,64-66,137-139)
void logError(const char *what) { fprintf(stderr, "[RendererCUDA] %s: %s\n", what, sgl_last_error()); }
}  // namespace

// ---- TextureCUDA ------------------------------------------------------------------------------------------------
TextureCUDA::TextureCUDA(const TextureDesc &desc, int layout) : id_(nextId()), layout_(layout) {
  width = desc.width;
  height = desc.height;
  type = desc.type;
  format = desc.format;
  usage = desc.usage;
  useMipmaps = desc.useMipmaps;
  multiSample = desc.multiSample;
  tag = desc.tag;
  // attachments are rendered through linear addressing; Tiled/Morton apply to sampled images
  if (multiSample || (usage & (TextureUsage_AttachmentColor | TextureUsage_AttachmentDepth))) layout_ = SGL_LAYOUT_LINEAR;
}

TextureCUDA::~TextureCUDA() {
  if (handle_) sgl_texture_destroy(handle_);
}

int TextureCUDA::levelCount() const {
  if (!useMipmaps) return 1;
  int m = std::max(width, height), n = 0;
  while ((1 << (n + 1)) <= m) n++;
  return n + 1;
}

bool TextureCUDA::ensureAllocated() {
  if (handle_) return true;
  SglTextureDesc d{};
  d.width = width;
  d.height = height;
  d.type = type == TextureType_CUBE ? SGL_TEX_CUBE : SGL_TEX_2D;
  d.format = format == TextureFormat_FLOAT32 ? SGL_FMT_FLOAT32 : SGL_FMT_RGBA8;
  d.use_mipmaps = useMipmaps;
  d.multi_sample = multiSample;
  d.layout = layout_;
  if (sgl_texture_create(&d, &handle_) != SGL_OK) {
    logError("createTexture");
    handle_ = 0;
    return false;
  }
  // An attachment that is also sampled (shadow map, FXAA input, IBL cubes): under sort-first tile sharding a later pass
  // reads pixels other ranks own, so passes into it render more than the owned tiles -- everything by default, or the
  // halo the application declared (RendererCUDA::setSampledAttachmentHalo)
  if ((usage & TextureUsage_Sampler) && (usage & (TextureUsage_AttachmentColor | TextureUsage_AttachmentDepth)))
    sgl_texture_set_shard_halo(handle_, RendererCUDA::sampledAttachmentHalo());
  return true;
}

void TextureCUDA::initImageData() { ensureAllocated(); }

template<typename T>
void TextureCUDA::upload(const std::vector<std::shared_ptr<Buffer<T>>> &buffers) {
  if (multiSample) {
    fprintf(stderr, "[RendererCUDA] setImageData not support: multi sample texture\n");
    return;
  }
  if (buffers.empty() || (size_t) width != buffers[0]->getWidth() || (size_t) height != buffers[0]->getHeight()) {
    fprintf(stderr, "[RendererCUDA] setImageData error: size not match\n");
    return;
  }
  if (!ensureAllocated()) return;
  int layers = layerCount();
  for (int i = 0; i < layers && i < (int) buffers.size(); i++) {
    if (sgl_texture_upload(handle_, i, 0, buffers[i]->getRawDataPtr()) != SGL_OK) logError("setImageData");
  }
  if (useMipmaps && sgl_texture_gen_mips(handle_) != SGL_OK) logError("generateMipmap");
}

void TextureCUDA::setImageData(const std::vector<std::shared_ptr<Buffer<RGBA>>> &buffers) { upload<RGBA>(buffers); }
void TextureCUDA::setImageData(const std::vector<std::shared_ptr<Buffer<float>>> &buffers) { upload<float>(buffers); }

bool TextureCUDA::readPixels(uint32_t layer, uint32_t level, int kind, std::vector<uint8_t> &out, int &w, int &h) {
  if (!handle_) return false;
  if (sgl_texture_level_size(handle_, (int) level, &w, &h) != SGL_OK) return false;
  size_t bytes = (size_t) w * h * 4 * ((multiSample && kind == 0) ? 4 : 1);
  out.resize(bytes);
  if (sgl_texture_readback(handle_, (int) layer, (int) level, kind, out.data(), bytes) != SGL_OK) {
    logError("readPixels");
    return false;
  }
  return true;
}

void TextureCUDA::dumpImage(const char *path, uint32_t layer, uint32_t level) {
  if (multiSample) return;   // same early-out as TextureSoft::dumpImageSoft (TextureSoft.h:230-233)
  std::vector<uint8_t> px;
  int w = 0, h = 0;
  if (!readPixels(layer, level, 0, px, w, h)) return;
  FILE *f = fopen(path, "wb");
  if (!f) return;
  // binary PAM (RGBA), rows flipped like the reference's PNG dump; float textures are written as grey
  fprintf(f, "P7\nWIDTH %d\nHEIGHT %d\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n", w, h);
  for (int y = h - 1; y >= 0; y--) {
    for (int x = 0; x < w; x++) {
      uint8_t o[4];
      if (format == TextureFormat_FLOAT32) {
        float v;
        memcpy(&v, &px[((size_t) y * w + x) * 4], 4);
        uint8_t g = (uint8_t) (std::min(std::max(v, 0.f), 1.f) * 255.f);
        o[0] = o[1] = o[2] = g;
        o[3] = 255;
      } else {
        memcpy(o, &px[((size_t) y * w + x) * 4], 4);
      }
      fwrite(o, 1, 4, f);
    }
  }
  fclose(f);
}

bool TextureCUDA::loadFromFile(const char *path) {
  if (!ensureAllocated()) return false;
  FILE *f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "[RendererCUDA] failed to open file: %s\n", path);
    return false;
  }
  bool ok = true;
  std::vector<uint8_t> buf;
  for (int layer = 0; layer < layerCount() && ok; layer++)
    for (int level = 0; level < levelCount() && ok; level++) {
      int w, h;
      sgl_texture_level_size(handle_, level, &w, &h);
      buf.resize((size_t) w * h * 4);
      if (fread(buf.data(), 1, buf.size(), f) != buf.size()) ok = false;
      else if (sgl_texture_upload(handle_, layer, level, buf.data()) != SGL_OK) ok = false;
    }
  fclose(f);
  if (!ok) fprintf(stderr, "[RendererCUDA] failed to read file: %s\n", path);
  return ok;
}

bool TextureCUDA::storeToFile(const char *path) {
  if (!handle_) return false;
  FILE *f = fopen(path, "wb");
  if (!f) {
    fprintf(stderr, "[RendererCUDA] failed to open file: %s\n", path);
    return false;
  }
  std::vector<uint8_t> buf;
  for (int layer = 0; layer < layerCount(); layer++)
    for (int level = 0; level < levelCount(); level++) {
      int w, h;
      if (readPixels(layer, level, 0, buf, w, h)) fwrite(buf.data(), 1, buf.size(), f);
    }
  fclose(f);
  return true;
}

// ---- small objects ---------------------------------------------------------------------------------------------------
FrameBufferCUDA::FrameBufferCUDA(bool offscreen) : FrameBuffer(offscreen), id_(nextId()) {}

VertexArrayObjectCUDA::VertexArrayObjectCUDA(const VertexArray &va) : id_(nextId()) {
  // VertexArrayObjectSoft copies vertices and indices at creation (VertexSoft.h:16-27)
  vertexStride = va.vertexesDesc.empty() ? va.vertexSize : va.vertexesDesc[0].stride;
  if (vertexStride != SGL_VERTEX_STRIDE) {
    fprintf(stderr, "[RendererCUDA] unsupported vertex stride %zu (expected %d)\n", vertexStride, SGL_VERTEX_STRIDE);
    return;
  }
  vertexCount = va.vertexesBufferLength / vertexStride;
  indexCount = va.indexBufferLength / sizeof(int32_t);
  if (sgl_buffer_create(va.vertexesBufferLength, va.vertexesBuffer, &vertexBuffer) != SGL_OK) logError("createVertexArrayObject");
  if (sgl_buffer_create(va.indexBufferLength, va.indexBuffer, &indexBuffer) != SGL_OK) logError("createVertexArrayObject");
}

VertexArrayObjectCUDA::~VertexArrayObjectCUDA() {
  if (vertexBuffer) sgl_buffer_destroy(vertexBuffer);
  if (indexBuffer) sgl_buffer_destroy(indexBuffer);
}

void VertexArrayObjectCUDA::updateVertexData(void *data, size_t length) {
  if (vertexBuffer && sgl_buffer_upload(vertexBuffer, 0, length, data) != SGL_OK) logError("updateVertexData");
}

// ---- ShaderProgramCUDA ---------------------------------------------------------------------------------------------------
ShaderProgramCUDA::ShaderProgramCUDA() : id_(nextId()) {}

bool ShaderProgramCUDA::setShadingModel(int shading) {
  int size = sgl_shader_uniform_size(shading);
  if (size < 0) return false;
  shader_ = shading;
  uniforms_.assign((size_t) size, 0);
  defineMask_ = 0;
  for (auto &d : defines_) {   // names not in the shader's define list are ignored (ShaderProgramSoft.h:37-45)
    int bit = sgl_shader_define_bit(shading, d.c_str());
    if (bit >= 0) defineMask_ |= 1u << bit;
  }
  return true;
}

int ShaderProgramCUDA::getUniformLocation(const std::string &name) const {
  int off = sgl_shader_uniform_offset(shader_, name.c_str());
  if (off >= 0) return off;
  int slot = sgl_shader_sampler_slot(shader_, name.c_str());
  if (slot >= 0) return kSamplerBase + slot;
  return -1;
}

void ShaderProgramCUDA::bindUniformBlockBuffer(const void *data, size_t len, int location) {
  if (location < 0 || location >= kSamplerBase || (size_t) location >= uniforms_.size()) return;
  len = std::min(len, uniforms_.size() - (size_t) location);
  memcpy(uniforms_.data() + location, data, len);
}

void ShaderProgramCUDA::bindUniformSampler(int texture, int filterMin, int wrap, int border, int location) {
  int slot = location - kSamplerBase;
  if (slot < 0 || slot >= 8) return;
  slots_[slot].texture = texture;
  slots_[slot].filterMin = filterMin;
  slots_[slot].wrap = wrap;
  slots_[slot].border = border;
}

// ---- uniforms ------------------------------------------------------------------------------------------------------------
int UniformBlockCUDA::getLocation(ShaderProgram &program) {
  auto *p = dynamic_cast<ShaderProgramCUDA *>(&program);
  if (!p) return -1;
  int loc = p->getUniformLocation(name);
  return loc >= ShaderProgramCUDA::kSamplerBase ? -1 : loc;
}
void UniformBlockCUDA::bindProgram(ShaderProgram &program, int location) {
  auto *p = dynamic_cast<ShaderProgramCUDA *>(&program);
  if (p) p->bindUniformBlockBuffer(bytes_.data(), bytes_.size(), location);
}
void UniformBlockCUDA::setSubData(void *data, int len, int offset) {
  if (offset < 0 || len <= 0 || (size_t) offset >= bytes_.size()) return;
  memcpy(bytes_.data() + offset, data, std::min((size_t) len, bytes_.size() - (size_t) offset));
}

int UniformSamplerCUDA::getLocation(ShaderProgram &program) {
  auto *p = dynamic_cast<ShaderProgramCUDA *>(&program);
  if (!p) return -1;
  int loc = p->getUniformLocation(name);
  return loc >= ShaderProgramCUDA::kSamplerBase ? loc : -1;
}
void UniformSamplerCUDA::bindProgram(ShaderProgram &program, int location) {
  auto *p = dynamic_cast<ShaderProgramCUDA *>(&program);
  if (p) p->bindUniformSampler(texture_, filterMin_, wrap_, border_, location);
}
void UniformSamplerCUDA::setTexture(const std::shared_ptr<Texture> &tex) {
  // Sampler2DSoft/SamplerCubeSoft::setTexture copy filterMin, wrapS and the border colour (SamplerSoft.h:388-394,428-436)
  auto *t = dynamic_cast<TextureCUDA *>(tex.get());
  if (!t) return;
  texture_ = t->handle();
  filterMin_ = (int) t->getSamplerDesc().filterMin;
  wrap_ = (int) t->getSamplerDesc().wrapS;
  border_ = (int) t->getSamplerDesc().borderColor;
}

// ---- RendererCUDA ----------------------------------------------------------------------------------------------------------
static int gSampledAttachmentHalo = -1;
void RendererCUDA::setSampledAttachmentHalo(int pixels) { gSampledAttachmentHalo = pixels; }
int RendererCUDA::sampledAttachmentHalo() { return gSampledAttachmentHalo; }

bool RendererCUDA::create() {
  if (const char *h = getenv("SGL_SHARD_HALO")) gSampledAttachmentHalo = atoi(h);   // harness knob (tests, bench)
  if (sgl_init(device_, rank_, world_) != SGL_OK) {
    logError("create");
    return false;
  }
  created_ = true;
  return true;
}

// Renderer::destroy (Renderer.h:28): releases this renderer's reference on the device context (sgl_shutdown frees every
// device resource once the last renderer is gone)
void RendererCUDA::destroy() {
  if (!created_) return;
  created_ = false;
  sgl_wait_idle();
  sgl_shutdown();
}

std::shared_ptr<FrameBuffer> RendererCUDA::createFrameBuffer(bool offscreen) { return std::make_shared<FrameBufferCUDA>(offscreen); }

std::shared_ptr<Texture> RendererCUDA::createTexture(const TextureDesc &desc) {
  switch (desc.format) {
    case TextureFormat_RGBA8:
    case TextureFormat_FLOAT32: return std::make_shared<TextureCUDA>(desc, textureLayout_);
  }
  return nullptr;
}

std::shared_ptr<VertexArrayObject> RendererCUDA::createVertexArrayObject(const VertexArray &va) {
  return std::make_shared<VertexArrayObjectCUDA>(va);
}
std::shared_ptr<ShaderProgram> RendererCUDA::createShaderProgram() { return std::make_shared<ShaderProgramCUDA>(); }
std::shared_ptr<PipelineStates> RendererCUDA::createPipelineStates(const RenderStates &rs) { return std::make_shared<PipelineStates>(rs); }
std::shared_ptr<UniformBlock> RendererCUDA::createUniformBlock(const std::string &name, int size) {
  return std::make_shared<UniformBlockCUDA>(name, size);
}
std::shared_ptr<UniformSampler> RendererCUDA::createUniformSampler(const std::string &name, const TextureDesc &desc) {
  return std::make_shared<UniformSamplerCUDA>(name, desc.type, desc.format);
}

void RendererCUDA::beginRenderPass(std::shared_ptr<FrameBuffer> &frameBuffer, const ClearStates &states) {
  auto *fbo = dynamic_cast<FrameBufferCUDA *>(frameBuffer.get());
  passOpen_ = false;
  if (!fbo) return;
  int color = 0, depth = 0, layer = 0, level = 0;
  if (fbo->isColorReady()) {
    auto *t = dynamic_cast<TextureCUDA *>(fbo->getColorAttachment().tex.get());
    if (t) {
      t->initImageData();
      color = t->handle();
      layer = (int) fbo->getColorAttachment().layer;
      level = (int) fbo->getColorAttachment().level;
    }
  }
  if (fbo->isDepthReady()) {
    auto *t = dynamic_cast<TextureCUDA *>(fbo->getDepthAttachment().tex.get());
    if (t) {
      t->initImageData();
      depth = t->handle();
    }
  }
  float cc[4] = {states.clearColor.r, states.clearColor.g, states.clearColor.b, states.clearColor.a};
  if (sgl_pass_begin(color, layer, level, depth, states.colorFlag ? 1 : 0, states.depthFlag ? 1 : 0, cc, states.clearDepth) != SGL_OK) {
    logError("beginRenderPass");
    return;
  }
  passOpen_ = true;
}

void RendererCUDA::setViewPort(int x, int y, int width, int height) { sgl_set_viewport(x, y, width, height); }
void RendererCUDA::setVertexArrayObject(std::shared_ptr<VertexArrayObject> &vao) { vao_ = dynamic_cast<VertexArrayObjectCUDA *>(vao.get()); }
void RendererCUDA::setShaderProgram(std::shared_ptr<ShaderProgram> &program) { program_ = dynamic_cast<ShaderProgramCUDA *>(program.get()); }
void RendererCUDA::setShaderResources(std::shared_ptr<ShaderResources> &resources) {
  if (!resources) return;
  if (program_) program_->bindResources(*resources);
}
void RendererCUDA::setPipelineStates(std::shared_ptr<PipelineStates> &states) { states_ = &states->renderStates; }

void RendererCUDA::draw() {
  if (!passOpen_ || !vao_ || !program_ || !states_ || !program_->shader()) return;
  SglDraw d;
  memset(&d, 0, sizeof(d));
  d.vertex_buffer = vao_->vertexBuffer;
  d.index_buffer = vao_->indexBuffer;
  d.vertex_count = (int32_t) vao_->vertexCount;
  d.index_count = (int32_t) vao_->indexCount;
  d.shader = program_->shader();
  d.defines = program_->defineMask();
  const RenderStates &rs = *states_;
  d.states.blend = rs.blend;
  d.states.blend_func_rgb = rs.blendParams.blendFuncRgb;
  d.states.blend_src_rgb = rs.blendParams.blendSrcRgb;
  d.states.blend_dst_rgb = rs.blendParams.blendDstRgb;
  d.states.blend_func_alpha = rs.blendParams.blendFuncAlpha;
  d.states.blend_src_alpha = rs.blendParams.blendSrcAlpha;
  d.states.blend_dst_alpha = rs.blendParams.blendDstAlpha;
  d.states.depth_test = rs.depthTest;
  d.states.depth_mask = rs.depthMask;
  d.states.depth_func = rs.depthFunc;
  d.states.cull_face = rs.cullFace;
  d.states.primitive_type = rs.primitiveType;
  d.states.polygon_mode = rs.polygonMode;
  d.states.line_width = rs.lineWidth;
  const auto &ub = program_->uniformBytes();
  d.uniform_bytes = (uint32_t) std::min<size_t>(ub.size(), SGL_MAX_UNIFORM_BYTES);
  memcpy(d.uniforms, ub.data(), d.uniform_bytes);
  const ShaderProgramCUDA::SamplerSlot *slots = program_->samplerSlots();
  for (int s = 0; s < SGL_MAX_SAMPLER_SLOTS; s++) {
    d.samplers[s].texture = slots[s].texture;
    d.samplers[s].filter_min = slots[s].filterMin;
    d.samplers[s].wrap = slots[s].wrap;
    d.samplers[s].border = slots[s].border;
  }
  if (sgl_draw(&d) != SGL_OK) logError("draw");
}

void RendererCUDA::endRenderPass() {
  if (!passOpen_) return;
  passOpen_ = false;
  if (sgl_pass_end() != SGL_OK) logError("endRenderPass");
}

void RendererCUDA::waitIdle() {
  if (sgl_wait_idle() != SGL_OK) logError("waitIdle");
}

}  // namespace SoftGL
