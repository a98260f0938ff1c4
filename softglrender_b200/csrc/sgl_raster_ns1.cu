// Tile rasteriser instantiation for single-sample targets (own translation unit: ptxas time is dominated by this kernel).
#define SGL_RASTER_ONLY
#include "sgl_kernels.cuh"
extern "C" int sglLaunchRaster1(const SglPassParams *P, int nTiles, void *stream) {
  sglRasterKernel<1><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  return (int) cudaGetLastError();
}
