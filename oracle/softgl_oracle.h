// ORACLE -- test infrastructure, NOT product code.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg may build, link or execute anything under oracle/.
//
// A scalar, single-threaded CPU restatement of SoftGLRender's software pipeline (RendererSoft + the software
// shaders + SamplerSoft), structured like the reference -- immediate mode, draw by draw, triangle by triangle,
// 32x32 blocks, 2x2 quads -- i.e. deliberately NOT like the tile-deferred CUDA implementation it checks.
// Every function cites the reference file:line it follows.  Floating-point association follows the reference
// *binary* (GCC -O3 -mavx2 -mfma, -ffp-contract=fast): fused operations are spelled with fmaf(), everything
// else relies on this file being compiled with -ffp-contract=off.
//
// Pinning: validated against the compiled reference (oracle/_ref, built from /root/reference by oracle/Makefile)
// on the traces of tests/golden/make_golden.py; the resulting hashes/images are committed under tests/golden/.
#pragma once
#include "Render/Renderer.h"

#include <cstdint>
#include <vector>

namespace SoftGL {

class TextureOracle : public Texture {
 public:
  explicit TextureOracle(const TextureDesc &d);
  int getId() const override { return id_; }
  void setSamplerDesc(SamplerDesc &s) override { sampler = s; }
  void initImageData() override;
  void setImageData(const std::vector<std::shared_ptr<Buffer<RGBA>>> &b) override;
  void setImageData(const std::vector<std::shared_ptr<Buffer<float>>> &b) override;
  void dumpImage(const char *, uint32_t, uint32_t) override {}

  int layers() const { return type == TextureType_CUBE ? 6 : 1; }
  int levelCount() const { return (int) levels.empty() ? 0 : (int) levels[0].size(); }
  int samples() const { return multiSample ? 4 : 1; }
  // levels[layer][level] = w*h*samples 32-bit texels, linear, [y][x][sample]
  std::vector<std::vector<std::vector<uint32_t>>> levels;
  std::vector<uint32_t> resolved;   // the reference's `buffer` beside `bufferMs4x`
  SamplerDesc sampler;
  void allocate(bool withMips);
  void generateMipmaps();

 private:
  int id_;
};

std::shared_ptr<Renderer> createRendererOracle();
bool oracleLoadShaders(ShaderProgram &program, int shading);
int oracleKatMain(int argc, char **argv);   // unit-level KAT entry (same protocol as oracle/ref_kat.cpp)

}  // namespace SoftGL

// trace-player backend header: the player only needs the Render API + PlayerBackend
