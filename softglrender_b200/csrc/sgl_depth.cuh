// Depth-only render passes (shadow maps): no colour attachment, every draw = filled triangles with depth test and
// depth write on and depth functions of one direction (LESS/LEQUAL or GREATER/GEQUAL).  The final depth of a sample
// is then min (resp. max) over the initial value and the clamped z of every fragment covering it -- independent of
// submission order -- so the pass needs no binning and no sorting:
//   sglDepthSetupKernel<NS>   one thread per input triangle: assembly, clipping, culling, setup (the same
//                             sglProcessInputPrim as the tile path); each triangle's pixel range is cut into row bands
//                             of <= 1024 pixels that are appended to a work queue (huge triangles: separate queue)
//   sglDepthRasterKernel<NS>  persistent grid, one warp per queue entry, lanes = 8x4 pixels per step: exact coverage
//                             and atomicMin/Max on the depth bits (non-negative floats order like their bit patterns)
//   sglDepthLargeKernel<NS>   one CTA per 16x16 tile, one thread per pixel, walks the queue of huge triangles
// Same per-sample arithmetic as the tile kernels (sglCoverTriangle), hence the same bits.
#pragma once
#include <cuda_runtime.h>
#include "sgl_vis.cuh"


#include "sgl_depth_pass.h"

// exact geometric coverage + clamped z of triangle p at pixel (px,py); no depth test (DEPTH_TEST flag must be cleared)
template<int NS>
__device__ __forceinline__ void sglDepthPixel(const SglPrim &p, const SglTriEdge &e, int px, int py, const SglDepthPass &D) {
  float z[NS];
  int shadeIdx;
  float dummy[NS];
  uint32_t mask = sglCoverTriangle<NS>(p, e, px, py, dummy, true, z, shadeIdx);
  if (!mask) return;
  int *dst = reinterpret_cast<int *>(D.depthBase) + ((size_t) py * D.fbW + px) * NS;
#pragma unroll
  for (int s = 0; s < NS; s++)
    if ((mask >> s) & 1u) {
      if (!(z[s] >= 0.f)) continue;                      // NaN never passes a depth test
      if (D.useMin) atomicMin(dst + s, __float_as_int(z[s]));
      else atomicMax(dst + s, __float_as_int(z[s]));
    }
}

#define SGL_DEPTH_CHUNK_AREA 1024    // work item of the raster kernel: one warp, at most this many pixels of one triangle
#define SGL_DEPTH_LARGE_AREA 65536   // pixel-range area above which a triangle goes to the tile-parallel kernel

template<int NS>
struct SglDepthAlloc {
  static constexpr bool kRecords = false;
  SglDepthPass D;
  __device__ int newVertex(const SglDrawRec &d) {
    int extra = atomicAdd(d.vertexCounter, 1);
    int idx = d.vertexCount + extra;
    return idx < d.vertexCap ? idx : -1;
  }
  __device__ int newAppendSlots(const SglDrawRec &d, int) { return d.appendBase; }
  __device__ void overflow() {
    atomicAdd(D.counters + 7, 1ull);
    *(volatile unsigned int *) D.overflowHost = 1u;
  }
  __device__ bool binPrim(int, const SglPrim &) { return false; }
  __device__ void pushLarge(const SglPrim &p) {
    uint32_t q = atomicAdd(D.largeCount, 1u);
    if (q < D.largeCapacity) D.large[q] = p; else overflow();
  }
  // splits the triangle's pixel range into row bands of <= SGL_DEPTH_CHUNK_AREA pixels, one queue entry (= the record
  // with by0/by1 narrowed to the band) per band; coverage itself never depends on the range
  __device__ void consume(const SglDrawRec &, const SglPrim &pin) {
    SglPrim p = pin;
    p.flags &= ~SGL_PF_DEPTH_TEST;                        // coverage + z only; the test is the atomic
    int x0 = p.bx0 < 0 ? 0 : p.bx0, y0 = p.by0 < 0 ? 0 : p.by0;
    int x1 = p.bx1 >= D.fbW ? D.fbW - 1 : p.bx1, y1 = p.by1 >= D.fbH ? D.fbH - 1 : p.by1;
    if (x1 < x0 || y1 < y0) return;
    const int w = x1 - x0 + 1, h = y1 - y0 + 1;
    if ((long long) w * h > SGL_DEPTH_LARGE_AREA) { pushLarge(p); return; }
    int rows = SGL_DEPTH_CHUNK_AREA / w;
    rows = rows < 4 ? 4 : (rows & ~3);                      // bands are whole 8x4 steps
    const int bands = (h + rows - 1) / rows;
    uint32_t q = atomicAdd(D.queueCount, (uint32_t) bands);
    if (q + bands > D.queueCapacity) {
      // queue full: the triangle goes to the tile-parallel kernel.  The reservation is NOT rolled back (a roll-back races
      // with concurrent reservations); the part of it that lies inside the queue is filled with empty records instead.
      SglPrim none = p;
      none.by0 = 1; none.by1 = 0;
      for (uint32_t k = q; k < D.queueCapacity && k < q + (uint32_t) bands; k++) D.queue[k] = none;
      pushLarge(p);
      return;
    }
    p.bx0 = (int16_t) x0; p.bx1 = (int16_t) x1;
    for (int b = 0; b < bands; b++) {
      p.by0 = (int16_t) (y0 + b * rows);
      int e = y0 + b * rows + rows - 1;
      p.by1 = (int16_t) (e > y1 ? y1 : e);
      D.queue[q + b] = p;
    }
  }
};

// grid = (ceil(maxInputPrims/128), drawCount): one thread per input triangle (assembly .. setup), output = work queue
template<int NS>
__global__ void __launch_bounds__(128) sglDepthSetupKernel(SglDepthPass D) {
  const SglDrawRec &d = D.draws[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.inputPrims) return;
  SglDepthAlloc<NS> alloc;
  alloc.D = D;
  SglSetupOut none = {nullptr, nullptr, nullptr};
  sglProcessInputPrim(d, blockIdx.y, i, true, none, alloc);
  if (i == 0) atomicAdd(D.counters + 2, (unsigned long long) d.inputPrims);
}

// persistent grid: each warp takes queue entries round-robin; lanes = 8x4 pixels per step
template<int NS>
__global__ void __launch_bounds__(256) sglDepthRasterKernel(SglDepthPass D) {
  uint32_t n = *D.queueCount;
  if (n > D.queueCapacity) n = D.queueCapacity;
  const int lane = threadIdx.x & 31, lx = lane & 7, ly = lane >> 3;
  const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
  for (uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n; k += warpsTotal) {
    SglPrim p;
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(D.queue + k);
      uint4 *dst = reinterpret_cast<uint4 *>(&p);
#pragma unroll
      for (int q = 0; q < 4; q++) dst[q] = __ldg(src + q);
    }
    const SglTriEdge e = sglTriEdge(p);
    const bool irregular = (p.flags & SGL_PF_IRREGULAR) != 0;
    for (int by = p.by0; by <= p.by1; by += 4)
      for (int bx = p.bx0; bx <= p.bx1; bx += 8) {
        const int px = bx + lx, py = by + ly;
        if (px > p.bx1 || py > p.by1) continue;
        if (D.tileOwner && D.tileOwner[(py / SGL_TILE) * D.tilesX + px / SGL_TILE] != D.rank) continue;
        if (irregular) {
          const SglDrawRec &d = D.draws[p.draw];
          int q;
          if (!sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, q)) continue;
          if (!sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, q)) continue;
        }
        sglDepthPixel<NS>(p, e, px, py, D);
      }
  }
}

// grid = tiles; usually the queue is empty and every CTA leaves after one load
template<int NS>
__global__ void __launch_bounds__(SGL_TILE_THREADS) sglDepthLargeKernel(SglDepthPass D) {
  uint32_t n = *D.largeCount;
  if (n > D.largeCapacity) n = D.largeCapacity;
  if (n == 0) return;
  const int tile = blockIdx.x;
  if (D.tileOwner && D.tileOwner[tile] != D.rank) return;
  const int tx = tile % D.tilesX, ty = tile / D.tilesX;
  const int px = tx * SGL_TILE + (threadIdx.x & (SGL_TILE - 1)), py = ty * SGL_TILE + (threadIdx.x / SGL_TILE);
  const bool inFb = px < D.fbW && py < D.fbH;
  const int tx0 = tx * SGL_TILE, ty0 = ty * SGL_TILE, tx1 = tx0 + SGL_TILE - 1, ty1 = ty0 + SGL_TILE - 1;
  __shared__ SglVisPrim sp;
  __shared__ int sHit;
  for (uint32_t k = 0; k < n; k++) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const SglPrim &q = D.large[k];
      sHit = (q.bx0 <= tx1 && q.bx1 >= tx0 && q.by0 <= ty1 && q.by1 >= ty0 && sglPrimNearTile(q, tx, ty)) ? 1 : 0;
      if (sHit) { sp.p = q; sp.e = sglTriEdge(q); }
    }
    __syncthreads();
    if (!sHit || !inFb) continue;
    const SglPrim &p = sp.p;
    if (px < p.bx0 || px > p.bx1 || py < p.by0 || py > p.by1) continue;
    if (p.flags & SGL_PF_IRREGULAR) {
      const SglDrawRec &d = D.draws[p.draw];
      int q;
      if (!sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, q)) continue;
      if (!sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, q)) continue;
    }
    sglDepthPixel<NS>(p, sp.e, px, py, D);
  }
}
