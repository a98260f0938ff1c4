// Depth-only (shadow map) pass kernels: prim-parallel setup + atomic rasterisation, tile-parallel walk of large triangles.
#include "sgl_depth.cuh"
// The setup kernel only touches the pass arena (clip vertices, work queues): it belongs to the GEOMETRY stage and runs before
// the pass has to wait for earlier pixel work; the rasterisers write the depth attachment (pixel stage).
extern "C" int sglLaunchDepthSetup(int samples, const SglDepthPass *D, int maxPrims, int nDraws, void *stream) {
  cudaStream_t st = (cudaStream_t) stream;
  dim3 grid((maxPrims + 127) / 128, nDraws);
  if (samples == 4) sglDepthSetupKernel<4><<<grid, dim3(128), 0, st>>>(*D);
  else sglDepthSetupKernel<1><<<grid, dim3(128), 0, st>>>(*D);
  return (int) cudaGetLastError();
}
extern "C" int sglLaunchDepthRaster(int samples, const SglDepthPass *D, int nTiles, void *stream) {
  cudaStream_t st = (cudaStream_t) stream;
  const dim3 persistent(148 * 6);   // 6 resident CTAs of 8 warps per SM
  if (samples == 4) {
    sglDepthRasterKernel<4><<<persistent, dim3(256), 0, st>>>(*D);
    sglDepthLargeKernel<4><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, st>>>(*D);
  } else {
    sglDepthRasterKernel<1><<<persistent, dim3(256), 0, st>>>(*D);
    sglDepthLargeKernel<1><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, st>>>(*D);
  }
  return (int) cudaGetLastError();
}
