// Renderer-API trace ("scene pack") opcodes.  Writer: softglrender_b200/scene/trace.py.
// Each command replays one call of the abstract Renderer interface of the reference
// (src/Render/Renderer.h:24-59 and companions), cited per opcode.
#pragma once
#include <cstdint>

namespace sgltrace {

constexpr uint32_t kMagic = 0x544c4753u;  // "SGLT"
constexpr uint32_t kVersion = 1;

enum Op : uint32_t {
  OP_CREATE_TEXTURE = 1,    // Renderer::createTexture(TextureDesc)            Renderer.h:34
  OP_TEX_SET_SAMPLER = 2,   // Texture::setSamplerDesc                         Texture.h:100
  OP_TEX_INIT = 3,          // Texture::initImageData                          Texture.h:101
  OP_TEX_SET_DATA = 4,      // Texture::setImageData(vector<Buffer<T>>)        Texture.h:102-103
  OP_TEX_LOAD_RAW = 6,      // TextureSoft::loadFromFile (.tex cache)          TextureSoft.h:166-193
  OP_TEX_STORE_RAW = 7,     // TextureSoft::storeToFile                        TextureSoft.h:195-215
  OP_CREATE_VAO = 10,       // Renderer::createVertexArrayObject               Renderer.h:37
  OP_VAO_UPDATE = 11,       // VertexArrayObject::updateVertexData             Vertex.h:18
  OP_CREATE_PROGRAM = 12,   // Renderer::createShaderProgram + addDefines + Viewer::loadShaders
  OP_CREATE_BLOCK = 13,     // Renderer::createUniformBlock                    Renderer.h:46
  OP_CREATE_SAMPLER = 14,   // Renderer::createUniformSampler                  Renderer.h:47
  OP_CREATE_PIPELINE = 15,  // Renderer::createPipelineStates                  Renderer.h:43
  OP_CREATE_FBO = 16,       // Renderer::createFrameBuffer                     Renderer.h:31
  OP_FBO_COLOR = 20,        // FrameBuffer::setColorAttachment                 Framebuffer.h:27-39
  OP_FBO_DEPTH = 21,        // FrameBuffer::setDepthAttachment                 Framebuffer.h:41-46
  OP_FBO_OFFSCREEN = 22,    // FrameBuffer::setOffscreen                       Framebuffer.h:78
  OP_BEGIN_PASS = 30,       // Renderer::beginRenderPass                       Renderer.h:50
  OP_VIEWPORT = 31,         // Renderer::setViewPort                           Renderer.h:51
  OP_BLOCK_DATA = 32,       // UniformBlock::setSubData                        Uniform.h:43
  OP_SAMPLER_TEX = 33,      // UniformSampler::setTexture                      Uniform.h:54
  OP_DRAW = 34,             // setVertexArrayObject/ShaderProgram/ShaderResources/PipelineStates + draw()
  OP_END_PASS = 35,         // Renderer::endRenderPass                         Renderer.h:57
  OP_WAIT_IDLE = 36,        // Renderer::waitIdle                              Renderer.h:58
  OP_READBACK = 40,         // harness read-back of an attachment (swapBuffer role, ViewerSoftware.h:30-44)
  OP_FRAME_BEGIN = 50,
  OP_FRAME_END = 51,
};

}  // namespace sgltrace
