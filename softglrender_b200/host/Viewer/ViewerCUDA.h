// ViewerCUDA: the Viewer subclass that puts RendererCUDA behind the reference's own Viewer (src/Viewer/Viewer.h:29,38-39),
// the twin of ViewerSoftware (src/Viewer/ViewerSoftware.h:22-69).  Compiles only inside / against the reference tree
// (it includes the reference's Viewer/Viewer.h); integration/Makefile builds it together with the reference's Viewer.cpp,
// ModelLoader.cpp and assimp into a headless harness, which is how tests/test_viewer_integration*.py prove the drop-in.
//
// swapBuffer(): the software viewer uploads its CPU frame into a GL texture (ViewerSoftware.h:30-44).  Here the frame lives
// in HBM; this header reads the resolved colour back into a host buffer (headless use, PNG dump) and -- in a windowed build,
// -DSGL_VIEWER_CUDA_PRESENT_GL, which is what the reference's main.cpp / ViewerManager needs -- hands it to the same GL
// texture with the same glTexSubImage2D call as ViewerSoftware, so the imgui window presents it unchanged.  (CUDA-GL
// interop on outTexId_ would save the PCIe round trip; the reference's present path is 8 MB per frame either way.)
#pragma once
#include <vector>
#include "Viewer/Viewer.h"
#ifdef SGL_VIEWER_CUDA_PRESENT_GL
#include "Render/OpenGL/OpenGLUtils.h"
#endif
#include "Render/CUDA/RendererCUDA.h"

namespace SoftGL {
namespace View {

class ViewerCUDA : public Viewer {
 public:
  ViewerCUDA(Config &config, Camera &camera) : Viewer(config, camera) {}

  void configRenderer() override {                       // ViewerSoftware.h:25-28
    camera_->setReverseZ(config_.reverseZ);
    cameraDepth_->setReverseZ(config_.reverseZ);
  }

  int swapBuffer() override {
    auto *tex = dynamic_cast<TextureCUDA *>(texColorMain_.get());
    if (tex) tex->readPixels(0, 0, tex->multiSample ? 1 : 0, frame_, frameWidth_, frameHeight_);
#ifdef SGL_VIEWER_CUDA_PRESENT_GL
    if (tex && !frame_.empty()) {                        // ViewerSoftware.h:33-43, same target, same format
      GL_CHECK(glBindTexture(GL_TEXTURE_2D, outTexId_));
      GL_CHECK(glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, frameWidth_, frameHeight_, GL_RGBA, GL_UNSIGNED_BYTE, frame_.data()));
    }
#endif
    return outTexId_;
  }

  std::shared_ptr<Renderer> createRenderer() override {
    auto renderer = std::make_shared<RendererCUDA>();
    // the FXAA filter is the only consumer that samples an attachment around the pixel it shades (multi-GPU tile sharding)
    RendererCUDA::setSampledAttachmentHalo(32);
    if (!renderer->create()) return nullptr;
    return renderer;
  }

  bool loadShaders(ShaderProgram &program, ShadingModel shading) override {   // ViewerSoftware.h:54-69
    auto *p = dynamic_cast<ShaderProgramCUDA *>(&program);
    return p && p->setShadingModel((int) shading);
  }

  // the frame swapBuffer() read back: RGBA8, row 0 = bottom row, like the software viewer's buffer
  const std::vector<uint8_t> &frame() const { return frame_; }
  int frameWidth() const { return frameWidth_; }
  int frameHeight() const { return frameHeight_; }
  // attachments, for harnesses that compare more than the presented frame
  Texture *colorTexture() { return texColorMain_.get(); }
  Texture *depthTexture() { return texDepthMain_.get(); }
  Texture *shadowTexture() { return texDepthShadow_.get(); }

 private:
  std::vector<uint8_t> frame_;
  int frameWidth_ = 0, frameHeight_ = 0;
};

}  // namespace View
}  // namespace SoftGL
