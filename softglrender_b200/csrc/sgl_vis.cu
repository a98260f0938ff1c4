// Visibility kernels of the deferred path (no shading code in this translation unit).
#include "sgl_vis.cuh"
extern "C" int sglLaunchVis(int samples, const SglPassParams *P, int nTiles, void *stream) {
  // MSAA: the grid has room for the quarter-tile CTAs of up to P->splitCap heavy tiles (sglTileOfBlock)
  if (samples == 4) sglVisKernel<4><<<dim3(nTiles + 3 * (P->tileOrder ? P->splitCap : 0)), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  else sglVisKernel<1><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  return (int) cudaGetLastError();
}
