// Visibility kernels of the deferred path (no shading code in this translation unit).
#include "sgl_vis.cuh"
extern "C" int sglLaunchVis(int samples, const SglPassParams *P, int nTiles, void *stream) {
  if (samples == 4) sglVisKernel<4><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  else sglVisKernel<1><<<dim3(nTiles), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  return (int) cudaGetLastError();
}
