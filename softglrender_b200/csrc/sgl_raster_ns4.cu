// Tile rasteriser instantiation for MSAA 4x targets (own translation unit: ptxas time is dominated by this kernel).
#define SGL_RASTER_ONLY
#include "sgl_kernels.cuh"
extern "C" int sglLaunchRaster4(const SglPassParams *P, int nTiles, void *stream) {
  sglRasterKernel<4><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  return (int) cudaGetLastError();
}
