// Depth-only (shadow map) pass kernels: prim-parallel setup + atomic rasterisation, tile-parallel walk of large triangles.
#include "sgl_depth.cuh"
extern "C" int sglLaunchDepthOnly(int samples, const SglDepthPass *D, int maxPrims, int nDraws, int nTiles, void *stream) {
  cudaStream_t st = (cudaStream_t) stream;
  dim3 grid((maxPrims + 127) / 128, nDraws);
  const dim3 persistent(148 * 6);   // 6 resident CTAs of 8 warps per SM
  if (samples == 4) {
    sglDepthSetupKernel<4><<<grid, dim3(128), 0, st>>>(*D);
    sglDepthRasterKernel<4><<<persistent, dim3(256), 0, st>>>(*D);
    sglDepthLargeKernel<4><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, st>>>(*D);
  } else {
    sglDepthSetupKernel<1><<<grid, dim3(128), 0, st>>>(*D);
    sglDepthRasterKernel<1><<<persistent, dim3(256), 0, st>>>(*D);
    sglDepthLargeKernel<1><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, st>>>(*D);
  }
  return (int) cudaGetLastError();
}
