// Visibility kernels of the deferred path (no shading code in this translation unit).
#include "sgl_vis.cuh"
// one CTA per work item: at most one per tile plus three more per split heavy MSAA tile (sglVisWorkCounts)
extern "C" int sglLaunchVis(int samples, const SglPassParams *P, int nTiles, void *stream) {
  const int grid = nTiles + 3 * P->splitCap;
  if (samples == 4) sglVisKernel<4><<<dim3(grid), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  else sglVisKernel<1><<<dim3(grid), dim3(SGL_TILE_THREADS), 0, (cudaStream_t) stream>>>(*P);
  return (int) cudaGetLastError();
}
