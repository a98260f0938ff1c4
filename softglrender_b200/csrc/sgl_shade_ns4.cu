// Deferred shading kernel for MSAA 4x targets (own translation unit: the fragment shaders dominate ptxas time).
#include "sgl_vis.cuh"
extern "C" int sglLaunchShade4(const SglPassParams *P, int nTiles, void *stream) {
  sglShadeKernel<4><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  return (int) cudaGetLastError();
}
