// Parameters of a depth-only pass (see sgl_depth.cuh)
#pragma once
#include "sgl_types.h"

struct SglDepthPass {
  const SglDrawRec *draws;
  float *depthBase;
  int fbW, fbH;
  int useMin;               // 1: LESS/LEQUAL (atomicMin), 0: GREATER/GEQUAL (atomicMax)
  SglPrim *queue;           // work items of the raster kernel (triangle records narrowed to a row band)
  uint32_t *queueCount;
  uint32_t queueCapacity;
  SglPrim *large;           // triangles with a huge pixel range (tile-parallel kernel)
  uint32_t *largeCount;
  uint32_t largeCapacity;
  unsigned long long *counters;
  unsigned int *overflowHost;   // pinned host word, set when a triangle had to be dropped
  const uint8_t *tileOwner;
  int tilesX, rank;
};
