#include "trace_player.h"
#include "trace_format.h"

#include <cstring>
#include <set>

using namespace SoftGL;
using namespace sgltrace;

namespace {
struct Rd {
  const uint8_t *p;
  const uint8_t *e;
  int32_t i32() { int32_t v; memcpy(&v, p, 4); p += 4; return v; }
  uint32_t u32() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  float f32() { float v; memcpy(&v, p, 4); p += 4; return v; }
  std::string str() { uint32_t n = u32(); std::string s((const char *) p, n); p += n; return s; }
  const uint8_t *raw(size_t n) { const uint8_t *r = p; p += n; return r; }
};
}  // namespace

TracePlayer::~TracePlayer() {
  if (out_) fclose(out_);
  out_ = nullptr;
  if (!renderer_) return;
  renderer_->waitIdle();
  resources_.clear();
  fbos_.clear();
  samplers_.clear();
  blocks_.clear();
  pipelines_.clear();
  programs_.clear();
  vaos_.clear();
  textures_.clear();
  renderer_->destroy();
  renderer_ = nullptr;
}

bool TracePlayer::load(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "trace_player: cannot open %s\n", path.c_str());
    return false;
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bytes_.resize((size_t) n);
  if (fread(bytes_.data(), 1, (size_t) n, f) != (size_t) n) {
    fclose(f);
    return false;
  }
  fclose(f);
  if (n < 8 || memcmp(bytes_.data(), "SGLT", 4) != 0) {
    fprintf(stderr, "trace_player: bad magic\n");
    return false;
  }
  size_t off = 8;
  while (off + 8 <= bytes_.size()) {
    uint32_t op, len;
    memcpy(&op, &bytes_[off], 4);
    memcpy(&len, &bytes_[off + 4], 4);
    off += 8;
    if (off + len > bytes_.size()) {
      fprintf(stderr, "trace_player: truncated command\n");
      return false;
    }
    if (op == OP_FRAME_BEGIN) frameBegin_ = (int) cmds_.size();
    if (op == OP_FRAME_END) frameEnd_ = (int) cmds_.size();
    cmds_.push_back({op, &bytes_[off], len});
    off += len;
  }
  return true;
}

bool TracePlayer::execRange(int from, int to) {
  for (int i = from; i < to; i++) {
    if (!exec(cmds_[i])) {
      fprintf(stderr, "trace_player: command %d (op %u) failed\n", i, cmds_[i].op);
      return false;
    }
  }
  return true;
}

bool TracePlayer::runSetup() {
  if (!renderer_) {
    renderer_ = PlayerBackend::createRenderer();
    if (!renderer_) return false;
  }
  int end = frameBegin_ >= 0 ? frameBegin_ : (int) cmds_.size();
  return execRange(0, end);
}

bool TracePlayer::runFrame(bool sync) {
  if (frameBegin_ < 0) return true;
  bool ok = execRange(frameBegin_ + 1, frameEnd_);
  if (sync) renderer_->waitIdle();
  return ok;
}

bool TracePlayer::runTail() {
  if (frameEnd_ < 0) return true;
  if (!outPath_.empty() && !out_) out_ = fopen(outPath_.c_str(), "wb");
  bool ok = execRange(frameEnd_ + 1, (int) cmds_.size());
  if (out_) {
    fclose(out_);
    out_ = nullptr;
  }
  return ok;
}

void TracePlayer::writeRecord(const std::string &tag, const PlayerBackend::Blob &b) {
  if (!out_) return;
  uint32_t n = (uint32_t) tag.size();
  fwrite(&n, 4, 1, out_);
  fwrite(tag.data(), 1, n, out_);
  int32_t hdr[4] = {b.width, b.height, b.format, b.samples};
  fwrite(hdr, 4, 4, out_);
  uint32_t nb = (uint32_t) b.data.size();
  fwrite(&nb, 4, 1, out_);
  fwrite(b.data.data(), 1, nb, out_);
}

bool TracePlayer::readbackTagged(const std::string &tag, PlayerBackend::Blob &out) {
  for (int i = frameEnd_ < 0 ? 0 : frameEnd_; i < (int) cmds_.size(); i++) {
    if (cmds_[i].op != OP_READBACK) continue;
    Rd r{cmds_[i].p, cmds_[i].p + cmds_[i].n};
    int tex = r.i32(), layer = r.i32(), level = r.i32();
    if (r.str() != tag) continue;
    Texture *t = texture(tex);
    if (!t) return false;
    // resolved colour for multisample colour targets; the attachment itself (per-sample data) otherwise
    return PlayerBackend::readback(*t, layer, level, (t->multiSample && t->format == TextureFormat_RGBA8) ? 1 : 0, out);
  }
  return false;
}

int TracePlayer::textureHandleTagged(const std::string &tag) {
  for (int i = frameEnd_ < 0 ? 0 : frameEnd_; i < (int) cmds_.size(); i++) {
    if (cmds_[i].op != OP_READBACK) continue;
    Rd r{cmds_[i].p, cmds_[i].p + cmds_[i].n};
    int tex = r.i32();
    r.i32();
    r.i32();
    if (r.str() != tag) continue;
    Texture *t = texture(tex);
    return t ? PlayerBackend::nativeHandle(*t) : -1;
  }
  return -1;
}

bool TracePlayer::exec(const Cmd &c) {
  Rd r{c.p, c.p + c.n};
  switch (c.op) {
    case OP_CREATE_TEXTURE: {
      TextureDesc d{};
      d.width = r.i32();
      d.height = r.i32();
      d.type = (TextureType) r.i32();
      d.format = (TextureFormat) r.i32();
      d.usage = (uint32_t) r.i32();
      d.useMipmaps = r.i32() != 0;
      d.multiSample = r.i32() != 0;
      auto t = renderer_->createTexture(d);
      textures_.push_back(t);
      return t != nullptr;
    }
    case OP_TEX_SET_SAMPLER: {
      int id = r.i32();
      SamplerDesc s{};
      s.filterMin = (FilterMode) r.i32();
      s.filterMag = (FilterMode) r.i32();
      s.wrapS = (WrapMode) r.i32();
      s.wrapT = (WrapMode) r.i32();
      s.wrapR = (WrapMode) r.i32();
      s.borderColor = (BorderColor) r.i32();
      textures_[id]->setSamplerDesc(s);
      return true;
    }
    case OP_TEX_INIT: {
      textures_[r.i32()]->initImageData();
      return true;
    }
    case OP_TEX_SET_DATA: {
      int id = r.i32();
      int n = r.i32();
      auto &tex = textures_[id];
      if (tex->format == TextureFormat_RGBA8) {
        std::vector<std::shared_ptr<Buffer<RGBA>>> bufs;
        for (int i = 0; i < n; i++) {
          uint32_t w = r.u32(), h = r.u32();
          auto b = Buffer<RGBA>::makeDefault(w, h);
          const uint8_t *src = r.raw((size_t) w * h * 4);
          for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
              RGBA px;
              memcpy(&px, src + ((size_t) y * w + x) * 4, 4);
              b->set(x, y, px);
            }
          bufs.push_back(b);
        }
        tex->setImageData(bufs);
      } else {
        std::vector<std::shared_ptr<Buffer<float>>> bufs;
        for (int i = 0; i < n; i++) {
          uint32_t w = r.u32(), h = r.u32();
          auto b = Buffer<float>::makeDefault(w, h);
          const uint8_t *src = r.raw((size_t) w * h * 4);
          for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
              float px;
              memcpy(&px, src + ((size_t) y * w + x) * 4, 4);
              b->set(x, y, px);
            }
          bufs.push_back(b);
        }
        tex->setImageData(bufs);
      }
      return true;
    }
    case OP_TEX_LOAD_RAW: {
      int id = r.i32();
      std::string p = dataDir_ + "/" + r.str();
      renderer_->waitIdle();
      return PlayerBackend::loadRaw(*textures_[id], p.c_str());
    }
    case OP_TEX_STORE_RAW: {
      int id = r.i32();
      std::string p = dataDir_ + "/" + r.str();
      renderer_->waitIdle();
      return PlayerBackend::storeRaw(*textures_[id], p.c_str());
    }
    case OP_CREATE_VAO: {
      uint32_t vb = r.u32();
      const uint8_t *v = r.raw(vb);
      uint32_t ib = r.u32();
      const uint8_t *idx = r.raw(ib);
      // 64-byte Vertex {vec3 pos@0, vec2 uv@16, vec3 normal@32, vec3 tangent@48} (src/Viewer/Model.h:20-46)
      vaoVertexBytes_.emplace_back(v, v + vb);
      std::vector<int32_t> indices(ib / 4);
      if (ib) memcpy(indices.data(), idx, ib);
      VertexArray va{};
      va.vertexSize = 64;
      va.vertexesDesc = {{3, 64, 0}, {2, 64, 16}, {3, 64, 32}, {3, 64, 48}};
      va.vertexesBuffer = vaoVertexBytes_.back().data();
      va.vertexesBufferLength = vb;
      va.indexBuffer = indices.data();
      va.indexBufferLength = ib;
      auto vao = renderer_->createVertexArrayObject(va);
      vaos_.push_back(vao);
      return vao != nullptr;
    }
    case OP_VAO_UPDATE: {
      int id = r.i32();
      uint32_t vb = r.u32();
      const uint8_t *v = r.raw(vb);
      vaoVertexBytes_[id].assign(v, v + vb);
      vaos_[id]->updateVertexData(vaoVertexBytes_[id].data(), vb);
      return true;
    }
    case OP_CREATE_PROGRAM: {
      int shading = r.i32();
      uint32_t n = r.u32();
      auto prog = renderer_->createShaderProgram();
      if (!prog) return false;
      std::set<std::string> defs;
      for (uint32_t i = 0; i < n; i++) defs.insert(r.str());
      prog->addDefines(defs);
      bool ok = PlayerBackend::loadShaders(*prog, shading);
      programs_.push_back(prog);
      return ok;
    }
    case OP_CREATE_BLOCK: {
      std::string name = r.str();
      int size = r.i32();
      blocks_.push_back(renderer_->createUniformBlock(name, size));
      return blocks_.back() != nullptr;
    }
    case OP_CREATE_SAMPLER: {
      std::string name = r.str();
      TextureDesc d{};
      d.type = (TextureType) r.i32();
      d.format = (TextureFormat) r.i32();
      samplers_.push_back(renderer_->createUniformSampler(name, d));
      return samplers_.back() != nullptr;
    }
    case OP_CREATE_PIPELINE: {
      RenderStates rs;
      rs.blend = r.i32() != 0;
      rs.blendParams.blendFuncRgb = (BlendFunction) r.i32();
      rs.blendParams.blendSrcRgb = (BlendFactor) r.i32();
      rs.blendParams.blendDstRgb = (BlendFactor) r.i32();
      rs.blendParams.blendFuncAlpha = (BlendFunction) r.i32();
      rs.blendParams.blendSrcAlpha = (BlendFactor) r.i32();
      rs.blendParams.blendDstAlpha = (BlendFactor) r.i32();
      rs.depthTest = r.i32() != 0;
      rs.depthMask = r.i32() != 0;
      rs.depthFunc = (DepthFunction) r.i32();
      rs.cullFace = r.i32() != 0;
      rs.primitiveType = (PrimitiveType) r.i32();
      rs.polygonMode = (PolygonMode) r.i32();
      rs.lineWidth = r.f32();
      pipelines_.push_back(renderer_->createPipelineStates(rs));
      return pipelines_.back() != nullptr;
    }
    case OP_CREATE_FBO: {
      fbos_.push_back(renderer_->createFrameBuffer(r.i32() != 0));
      return fbos_.back() != nullptr;
    }
    case OP_FBO_COLOR: {
      int fbo = r.i32(), tex = r.i32(), face = r.i32(), level = r.i32();
      if (face < 0) {
        fbos_[fbo]->setColorAttachment(textures_[tex], level);
      } else {
        fbos_[fbo]->setColorAttachment(textures_[tex], (CubeMapFace) face, level);
      }
      return true;
    }
    case OP_FBO_DEPTH: {
      int fbo = r.i32(), tex = r.i32();
      fbos_[fbo]->setDepthAttachment(textures_[tex]);
      return true;
    }
    case OP_FBO_OFFSCREEN: {
      int fbo = r.i32();
      fbos_[fbo]->setOffscreen(r.i32() != 0);
      return true;
    }
    case OP_BEGIN_PASS: {
      int fbo = r.i32();
      ClearStates cs{};
      cs.colorFlag = r.i32() != 0;
      cs.depthFlag = r.i32() != 0;
      float c0 = r.f32(), c1 = r.f32(), c2 = r.f32(), c3 = r.f32();
      cs.clearColor = glm::vec4(c0, c1, c2, c3);
      cs.clearDepth = r.f32();
      renderer_->beginRenderPass(fbos_[fbo], cs);
      return true;
    }
    case OP_VIEWPORT: {
      int x = r.i32(), y = r.i32(), w = r.i32(), h = r.i32();
      renderer_->setViewPort(x, y, w, h);
      return true;
    }
    case OP_BLOCK_DATA: {
      int id = r.i32(), off = r.i32(), len = r.i32();
      blocks_[id]->setSubData((void *) r.raw((size_t) len), len, off);
      return true;
    }
    case OP_SAMPLER_TEX: {
      int id = r.i32(), tex = r.i32();
      samplers_[id]->setTexture(textures_[tex]);
      return true;
    }
    case OP_DRAW: {
      int vao = r.i32(), prog = r.i32(), pipe = r.i32();
      std::vector<int> key;
      int nb = r.i32();
      key.push_back(nb);
      for (int i = 0; i < 2 * nb; i++) key.push_back(r.i32());
      int ns = r.i32();
      key.push_back(ns);
      for (int i = 0; i < 2 * ns; i++) key.push_back(r.i32());
      auto it = resources_.find(key);
      if (it == resources_.end()) {
        auto res = std::make_shared<ShaderResources>();
        for (int i = 0; i < nb; i++) res->blocks[key[1 + 2 * i]] = blocks_[key[2 + 2 * i]];
        for (int i = 0; i < ns; i++) res->samplers[key[2 + 2 * nb + 2 * i]] = samplers_[key[3 + 2 * nb + 2 * i]];
        it = resources_.emplace(key, res).first;
      }
      renderer_->setVertexArrayObject(vaos_[vao]);
      renderer_->setShaderProgram(programs_[prog]);
      renderer_->setShaderResources(it->second);
      renderer_->setPipelineStates(pipelines_[pipe]);
      renderer_->draw();
      return true;
    }
    case OP_END_PASS:
      renderer_->endRenderPass();
      return true;
    case OP_WAIT_IDLE:
      renderer_->waitIdle();
      return true;
    case OP_READBACK: {
      int tex = r.i32(), layer = r.i32(), level = r.i32();
      std::string tag = r.str();
      renderer_->waitIdle();
      PlayerBackend::Blob b;
      auto &t = textures_[tex];
      if (!PlayerBackend::readback(*t, layer, level, 0, b)) return false;
      writeRecord(t->multiSample ? tag + ".ms" : tag, b);
      if (t->multiSample && t->format == TextureFormat_RGBA8) {
        PlayerBackend::Blob rb;
        if (PlayerBackend::readback(*t, layer, level, 1, rb)) writeRecord(tag, rb);
      }
      return true;
    }
    case OP_FRAME_BEGIN:
    case OP_FRAME_END:
      return true;
    default:
      fprintf(stderr, "trace_player: unknown opcode %u\n", c.op);
      return false;
  }
}
