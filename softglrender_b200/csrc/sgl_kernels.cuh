// sm_100a kernels of RendererCUDA.  Geometry stage of a pass:
//   sglVertexKernel   vertex shading + clip mask + perspective divide + viewport, one thread per VAO vertex
//   sglSetupKernel    assembly, clipping (VS re-execution), culling, primitive setup, tile counting
//   sglTileScanKernel exclusive scan of per-tile counts + heavy-first tile classes (single CTA)
//   sglBinFillKernel  scatter primitive slots into per-tile lists
// Pixel stage: the deferred pair sglVisKernel / sglShadeKernel (sgl_vis.cuh) for opaque draws, or
//   sglRasterKernel   the fused ordered form for blending and for lines/points of programs with varyings: one CTA per
//                     16x16 tile, order-sorted list, coverage/depth in registers, immediate shading + blend for blended
//                     fragments, deferred owners for opaque ones, MSAA resolve, write-back
// plus the depth-only kernels of sgl_depth.cuh, upload/utility kernels and the multi-GPU gather helpers.
#pragma once
#include <cuda_runtime.h>
#include "sgl_pixel.h"
#include "sgl_vis.cuh"

#ifndef SGL_RASTER_ONLY
// ---------------------------------------------------------------------------------------------------------
struct SglSetupShared {
  uint32_t *tileCount;
  uint32_t *bigList;
  uint32_t *bigCount;
  uint32_t bigCapacity;
  uint32_t *binReserved;       // running upper bound of bin entries handed out (zero-initialised with the counters)
  uint32_t binCapacity;
  unsigned long long *counters;
  unsigned int *overflowHost;  // pinned host word: set when geometry had to be DROPPED (clip arena full) -- read at sync points
  int tilesX, tilesY, fbW, fbH;
  const uint8_t *tileOwner;
  int rank;
};

// tile rectangle of a primitive (clamped to the framebuffer); false if empty
__device__ __forceinline__ bool sglPrimTiles(const SglPrim &p, int fbW, int fbH, int &tx0, int &ty0, int &tx1, int &ty1) {
  int x0 = p.bx0 < 0 ? 0 : p.bx0, y0 = p.by0 < 0 ? 0 : p.by0;
  int x1 = p.bx1 >= fbW ? fbW - 1 : p.bx1, y1 = p.by1 >= fbH ? fbH - 1 : p.by1;
  if (x1 < x0 || y1 < y0) return false;
  tx0 = x0 / SGL_TILE; ty0 = y0 / SGL_TILE; tx1 = x1 / SGL_TILE; ty1 = y1 / SGL_TILE;
  return true;
}

// atomicAdd(counter, n) for the lanes that are converged here, as ONE atomic per warp (the reservation counter of the bins
// is a single address: 10 M primitives would otherwise serialise on it)
__device__ __forceinline__ uint32_t sglWarpAggregatedAdd(uint32_t *counter, uint32_t n) {
  const uint32_t active = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
  uint32_t incl = n;
  if (active == 0xffffffffu) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
  } else {   // partially converged warp: sum the active lanes at or below this one
    incl = 0;
    for (uint32_t m = active; m; m &= m - 1) {
      const int src = __ffs(m) - 1;
      const uint32_t v = __shfl_sync(active, n, src);
      if (src <= lane) incl += v;
    }
  }
  const uint32_t total = __shfl_sync(active, incl, 31 - __clz(active));
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, total);
  base = __shfl_sync(active, base, leader);
  return base + incl - n;
}

struct SglDeviceAlloc {
  static constexpr bool kRecords = true;
  __device__ void consume(const SglDrawRec &, const SglPrim &) {}
  SglSetupShared S;
  __device__ int newVertex(const SglDrawRec &d) {
    int extra = atomicAdd(d.vertexCounter, 1);
    int idx = d.vertexCount + extra;
    return idx < d.vertexCap ? idx : -1;
  }
  __device__ int newAppendSlots(const SglDrawRec &d, int n) {
    int a = atomicAdd(d.appendCounter, n);
    return a + n <= d.appendCap ? d.appendBase + a : -1;
  }
  __device__ void overflow() {
    atomicAdd(S.counters + 7, 1ull);
    *(volatile unsigned int *) S.overflowHost = 1u;
  }
  // Counts the primitive into the bins of the tiles it can touch.  Returns 1 when it goes to the pass-wide big list
  // instead (SGL_PF_BIG): more than SGL_BIG_PRIM_TILES tiles, or the bin region is exhausted -- the big list holds one
  // entry per primitive slot at most, so binning can never drop geometry, it only gets slower.  Returns 2 when the pass is
  // tile-sharded and none of the tiles the primitive can touch is rendered by this rank: the caller does not emit it.
  __device__ int binPrim(int slot, const SglPrim &p) {
    int tx0, ty0, tx1, ty1;
    if (!sglPrimTiles(p, S.fbW, S.fbH, tx0, ty0, tx1, ty1)) return S.tileOwner ? 2 : 0;
    int n = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
    bool big = n > SGL_BIG_PRIM_TILES;
    if (!big && S.tileOwner) {     // sharded: look before reserving, most primitives belong to other ranks
      bool mine = false;
      for (int ty = ty0; ty <= ty1 && !mine; ty++)
        for (int tx = tx0; tx <= tx1; tx++)
          if (S.tileOwner[ty * S.tilesX + tx] == S.rank && sglPrimNearTile(p, tx, ty)) { mine = true; break; }
      if (!mine) return 2;
    }
    if (!big) {
      const uint32_t r = sglWarpAggregatedAdd(S.binReserved, (uint32_t) n);
      if (r + (uint32_t) n > S.binCapacity || r + (uint32_t) n < r) { big = true; atomicAdd(S.counters + 1, 1ull); }
    }
    if (big) {   // sglBigBinKernel bins it (or leaves it in the residual list)
      uint32_t b = atomicAdd(S.bigCount, 1u);
      if (b < S.bigCapacity) S.bigList[b] = (uint32_t) slot;
      return 1;
    }
    for (int ty = ty0; ty <= ty1; ty++)
      for (int tx = tx0; tx <= tx1; tx++) {
        int t = ty * S.tilesX + tx;
        if (S.tileOwner && S.tileOwner[t] != S.rank) continue;
        if (!sglPrimNearTile(p, tx, ty)) continue;
        atomicAdd(&S.tileCount[t], 1u);
      }
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------------------
// processVertexShader + perspective divide + viewport transform (RendererSoft.cpp:170-190,259-275,971-992)
// grid = (ceil(maxVertices/128), drawCount); vertex loads are 4 x LDG.128 (64-byte Vertex, Model.h:20-25).
// POSITION_ONLY: gl_Position, clip mask and screen position from the first 16 bytes of the vertex, no varyings -- depth-only
// passes (nothing reads varyings) and passes with lazy varyings (SglDrawRec::vertexUsed, sglVaryingKernel).
template<bool POSITION_ONLY>
__global__ void __launch_bounds__(128) sglVertexKernel(const SglDrawRec *draws) {
  const SglDrawRec &d = draws[blockIdx.y];
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= d.vertexCount) return;
  const float4 *src = reinterpret_cast<const float4 *>(d.vertexIn) + (size_t) v * 4;
  V4 clip;
  if (POSITION_ONLY) {
    const float4 q = __ldg(src);
    const float pos[3] = {q.x, q.y, q.z};
    clip = sglVertexPosition(d, pos);
  } else {
    float attr[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float4 q = __ldg(src + i);
      attr[4 * i] = q.x; attr[4 * i + 1] = q.y; attr[4 * i + 2] = q.z; attr[4 * i + 3] = q.w;
    }
    float vary[32];
    clip = sglVertexShader(d, attr, vary);
    float4 *vo = reinterpret_cast<float4 *>(d.varyings + (size_t) v * d.varyingStride);
    const int nq = d.varyingStride / 4;
#pragma unroll
    for (int k = 0; k < 8; k++)      // fully unrolled + predicated: constant indices keep vary[] in registers
      if (k < nq) vo[k] = make_float4(vary[4 * k], vary[4 * k + 1], vary[4 * k + 2], vary[4 * k + 3]);
  }
  reinterpret_cast<float4 *>(d.clipPos)[v] = make_float4(clip.x, clip.y, clip.z, clip.w);
  d.clipMask[v] = sglClipMask(clip);
  V4 f = sglToScreen(clip, d.vpX, d.vpY, d.vpW, d.vpH);
  reinterpret_cast<float4 *>(d.fragPos)[v] = make_float4(f.x, f.y, f.z, f.w);
}

// Lazy varyings: the full vertex shader for the vertices the setup kernel marked (SglDrawRec::vertexUsed) -- in a
// tile-sharded pass a rank needs the varyings of the primitives that reach its own tiles only.  A thread looks at eight
// flags with one 8-byte load (the flag array of a draw is 16-byte aligned and padded): grid = (ceil(maxVertices/1024), draws).
#define SGL_VARYING_PER_THREAD 8
__global__ void __launch_bounds__(128) sglVaryingKernel(const SglDrawRec *draws) {
  const SglDrawRec &d = draws[blockIdx.y];
  const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * SGL_VARYING_PER_THREAD;
  if (v0 >= d.vertexCount || !d.vertexUsed || d.varyingStride == 0) return;
  const uint2 f = *reinterpret_cast<const uint2 *>(d.vertexUsed + v0);
  if ((f.x | f.y) == 0u) return;
#pragma unroll 1
  for (int k = 0; k < SGL_VARYING_PER_THREAD; k++) {
    const uint32_t w = k < 4 ? f.x : f.y;
    if (!((w >> (8 * (k & 3))) & 0xffu) || v0 + k >= d.vertexCount) continue;
    const int v = v0 + k;
    const float4 *src = reinterpret_cast<const float4 *>(d.vertexIn) + (size_t) v * 4;
    float attr[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float4 q = __ldg(src + i);
      attr[4 * i] = q.x; attr[4 * i + 1] = q.y; attr[4 * i + 2] = q.z; attr[4 * i + 3] = q.w;
    }
    float vary[32];
    sglVertexShader(d, attr, vary);
    float4 *vo = reinterpret_cast<float4 *>(d.varyings + (size_t) v * d.varyingStride);
    const int nq = d.varyingStride / 4;
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (q < nq) vo[q] = make_float4(vary[4 * q], vary[4 * q + 1], vary[4 * q + 2], vary[4 * q + 3]);
  }
}

// grid = (ceil(maxInputPrims/128), drawCount)
__global__ void __launch_bounds__(128) sglSetupKernel(const SglDrawRec *draws, SglSetupOut out, SglSetupShared S, int hasDepth) {
  const SglDrawRec &d = draws[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.inputPrims) return;
  SglDeviceAlloc alloc;
  alloc.S = S;
  sglProcessInputPrim(d, blockIdx.y, i, hasDepth != 0, out, alloc);
  if (i == 0) atomicAdd(S.counters + 2, (unsigned long long) d.inputPrims);
}

// Binning of the big primitives (SglPassParams::bigAll): one CTA per big primitive (strided), threads over the tiles of
// its pixel range, same conservative near-tile test as everywhere.  FILL = 0 (before the scan): reserves the exact number
// of entries; if the bins cannot take them the primitive moves to the residual list (bigList) that every tile kernel
// scans, else the per-tile counts are raised.  FILL = 1 (after the scan): writes the slots.  Residual primitives are
// marked in bigAll by their top bit.
#define SGL_BIGBIN_THREADS 1024
template<int FILL>
__global__ void __launch_bounds__(SGL_BIGBIN_THREADS) sglBigBinKernel(SglPassParams P) {
  __shared__ uint32_t sWarp[SGL_BIGBIN_THREADS / 32];
  __shared__ uint32_t sOk;
  uint32_t nBig = *P.bigAllCount;
  if (nBig > P.bigCapacity) nBig = P.bigCapacity;
  const int tid = threadIdx.x;
  for (uint32_t i = blockIdx.x; i < nBig; i += gridDim.x) {
    const uint32_t entry = P.bigAll[i];
    if (FILL && (entry >> 31)) continue;
    const uint32_t slot = entry & 0x7fffffffu;
    const SglPrim p = P.prims[slot];
    int tx0, ty0, tx1, ty1;
    if (!sglPrimTiles(p, P.fbW, P.fbH, tx0, ty0, tx1, ty1)) continue;    // block-uniform
    const int w = tx1 - tx0 + 1, n = w * (ty1 - ty0 + 1);
    if (!FILL) {
      uint32_t cnt = 0;
      for (int k = tid; k < n; k += SGL_BIGBIN_THREADS) {
        const int tx = tx0 + k % w, ty = ty0 + k / w, t = ty * P.tilesX + tx;
        if (P.tileOwner && P.tileOwner[t] != P.rank) continue;
        if (sglPrimNearTile(p, tx, ty)) cnt++;
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      __syncthreads();
      if ((tid & 31) == 0) sWarp[tid >> 5] = cnt;
      __syncthreads();
      if (tid == 0) {
        uint32_t total = 0;
        for (int k = 0; k < SGL_BIGBIN_THREADS / 32; k++) total += sWarp[k];
        const uint32_t r = atomicAdd(P.binReserved, total);
        const bool ok = r + total <= P.binCapacity && r + total >= r;
        if (!ok) {
          P.bigAll[i] = entry | 0x80000000u;
          const uint32_t b = atomicAdd(P.bigCount, 1u);
          if (b < P.bigCapacity) P.bigList[b] = slot;
          atomicAdd(P.counters + 1, 1ull);
        }
        sOk = ok ? 1u : 0u;
      }
      __syncthreads();
      if (!sOk) continue;
    }
    for (int k = tid; k < n; k += SGL_BIGBIN_THREADS) {
      const int tx = tx0 + k % w, ty = ty0 + k / w, t = ty * P.tilesX + tx;
      if (P.tileOwner && P.tileOwner[t] != P.rank) continue;
      if (!sglPrimNearTile(p, tx, ty)) continue;
      if (FILL) P.binSlots[P.tileOffset[t] + atomicAdd(&P.tileCursor[t], 1u)] = slot;
      else atomicAdd(&P.tileCount[t], 1u);
    }
  }
}

// tileOffset = exclusive scan(tileCount); tileOffset[nTiles] = total.  Single pass over any number of tiles: one CTA per
// 1024 tiles, chained by decoupled look-back (each CTA publishes its aggregate, then its inclusive prefix, in a 64-bit word
// {status, value}; a CTA sums the aggregates of its predecessors until it meets a published prefix).  CTAs take their
// position from a ticket counter, so every predecessor of a waiting CTA is already running.
// Also classifies the tiles by bin length (heavy first: SglPassParams::tileOrder; each CTA reserves its share of every class
// list with one atomic) and marks every tile's pre-sorted list as absent until sglTileSortKernel has prepared it.
// `scanState` (ceil(nTiles / 1024) + 1 words, the last one is the ticket) must be zero on entry.
__global__ void __launch_bounds__(1024) sglTileScanKernel(const uint32_t *tileCount, uint32_t *tileOffset, int nTiles,
                                                         unsigned long long *counters, uint32_t *tileOrder, uint32_t *tileClassCount,
                                                         uint32_t *tileSortedCount, const uint8_t *tileOwner, int rank,
                                                         unsigned long long *scanState) {
  __shared__ uint32_t sWarp[32];
  __shared__ uint32_t sPrefix, sPart;
  __shared__ uint32_t sClass[SGL_TILE_CLASSES], sClassBase[SGL_TILE_CLASSES];
  const int nParts = (nTiles + 1023) / 1024;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) sPart = (uint32_t) atomicAdd(&scanState[nParts], 1ull);
  if (threadIdx.x < SGL_TILE_CLASSES) sClass[threadIdx.x] = 0;
  __syncthreads();
  const int part = (int) sPart;
  const int i = part * 1024 + (int) threadIdx.x;
  const uint32_t v = i < nTiles ? tileCount[i] : 0u;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sWarp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = sWarp[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    sWarp[lane] = winc - w;   // exclusive
    if (lane == 31) {
      // ---- decoupled look-back (one thread: the chain is short, ceil(nTiles / 1024) links)
      const unsigned long long AGG = 1ull << 62, PRE = 2ull << 62, VAL = (1ull << 62) - 1ull;
      const uint32_t aggregate = winc;
      volatile unsigned long long *st = scanState;
      if (part > 0) {
        st[part] = AGG | aggregate;
        __threadfence();
      }
      uint32_t prefix = 0;
      for (int p = part - 1; p >= 0; p--) {
        unsigned long long s;
        while (((s = st[p]) >> 62) == 0ull) { }
        prefix += (uint32_t) (s & VAL);
        if ((s >> 62) == 2ull) break;
      }
      st[part] = PRE | (unsigned long long) (prefix + aggregate);
      __threadfence();
      sPrefix = prefix;
      if (part == nParts - 1) {
        tileOffset[nTiles] = prefix + aggregate;
        atomicAdd(counters + 3, (unsigned long long) (prefix + aggregate));
      }
    }
  }
  __syncthreads();
  const uint32_t excl = sPrefix + sWarp[warp] + inc - v;
  int cls = -1;
  uint32_t pos = 0;
  if (i < nTiles) {
    tileOffset[i] = excl;
    if (tileOrder && !(tileOwner && tileOwner[i] != rank)) {
      tileSortedCount[i] = SGL_TILE_UNSORTED;
      cls = v >= 184u ? 0 : (v >= 40u ? 1 : (v >= 8u ? 2 : 3));
      pos = atomicAdd(&sClass[cls], 1u);
    }
  }
  __syncthreads();
  if (tileOrder && threadIdx.x < SGL_TILE_CLASSES) sClassBase[threadIdx.x] = sClass[threadIdx.x] ? atomicAdd(&tileClassCount[threadIdx.x], sClass[threadIdx.x]) : 0u;
  __syncthreads();
  if (cls >= 0) tileOrder[(size_t) cls * nTiles + sClassBase[cls] + pos] = (uint32_t) i;
}

// grid = (ceil(maxSlotsPerDraw/256), drawCount): one thread per primitive slot (originals, then appended)
__global__ void __launch_bounds__(256) sglBinFillKernel(SglPassParams P) {
  const SglDrawRec &d = P.draws[blockIdx.y];
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int originals = d.inputPrims * d.slotsPerPrim;
  int slot;
  if (j < originals) slot = d.primBase + j;
  else {
    int a = j - originals;
    int used = *d.appendCounter;
    if (used > d.appendCap) used = d.appendCap;
    if (a >= used) return;
    slot = d.appendBase + a;
  }
  const SglPrim p = P.prims[slot];
  if (!(p.flags & SGL_PF_VALID)) return;
  int tx0, ty0, tx1, ty1;
  if (!sglPrimTiles(p, P.fbW, P.fbH, tx0, ty0, tx1, ty1)) return;
  if (p.flags & SGL_PF_BIG) return;     // lives in the pass-wide big list (bigCapacity >= primSlots, never overflows)
  // <= SGL_BIG_PRIM_TILES (64) tiles: first the set of tiles as a bit mask, then the cursor atomics four at a time -- the
  // returned positions are needed for the stores only, so four round trips overlap instead of one per tile in sequence
  const int w = tx1 - tx0 + 1, n = w * (ty1 - ty0 + 1);
  unsigned long long m = 0ull;
  for (int k = 0; k < n && k < 64; k++) {
    const int tx = tx0 + k % w, ty = ty0 + k / w, t = ty * P.tilesX + tx;
    if (P.tileOwner && P.tileOwner[t] != P.rank) continue;
    if (sglPrimNearTile(p, tx, ty)) m |= 1ull << k;
  }
  while (m) {
    int t[4];
    uint32_t c[4], o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      t[q] = -1;
      if (m) {
        const int k = __ffsll((long long) m) - 1;
        m &= m - 1ull;
        t[q] = (ty0 + k / w) * P.tilesX + tx0 + k % w;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (t[q] >= 0) { c[q] = atomicAdd(&P.tileCursor[t[q]], 1u); o[q] = P.tileOffset[t[q]]; }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (t[q] >= 0) P.binSlots[o[q] + c[q]] = (uint32_t) slot;   // < binCapacity: the setup kernel reserved every entry it counted
  }
}

#endif  // SGL_RASTER_ONLY

// ---------------------------------------------------------------------------------------------------------
#ifndef SGL_SORT_CAP
#define SGL_SORT_CAP 2048
#endif
#define SGL_PRIM_BATCH 64

template<int NS>
__global__ void __launch_bounds__(SGL_TILE_THREADS) sglRasterKernel(SglPassParams P) {
  __shared__ uint32_t sKeys[SGL_SORT_CAP];
  __shared__ uint32_t sSlots[SGL_SORT_CAP];
  __shared__ __align__(16) SglPrim sPrims[SGL_PRIM_BATCH];
  __shared__ int sCount;
  __shared__ unsigned int sShaded;

  int quarter;
  const int tile = sglTileOfBlock(P, blockIdx.x, 0, quarter);   // heavy tiles first
  if (tile < 0) return;
  const int tx = tile % P.tilesX, ty = tile / P.tilesX;
  const int tid = threadIdx.x;
  // a warp = a 16x2 pixel strip: this kernel mostly sees few, large (blended) primitives, where the row-major strip's
  // 64-byte colour segments beat the 8x4 block + warp-level cull of the visibility kernel (measured on config 3)
  const int px = tx * SGL_TILE + (tid & (SGL_TILE - 1));
  const int py = ty * SGL_TILE + (tid / SGL_TILE);
  const bool inFb = px < P.fbW && py < P.fbH;
  const bool hasColor = P.colorBase != nullptr, hasDepth = P.depthBase != nullptr;
  const size_t pix = (size_t) py * P.fbW + px;

  SglPixelState<NS> st;
#pragma unroll
  for (int s = 0; s < NS; s++) { st.depth[s] = P.clearDepth; st.color[s] = P.clearColor; st.owner[s] = SGL_OWNER_NONE; }
  bool loaded = false;
  auto loadState = [&]() {   // attachments that are not cleared are read once, after the first gather (empty tiles may skip it)
    if (loaded) return;
    loaded = true;
    if (!inFb) return;
    if (hasDepth && !P.clearDepthFlag) {
      if (NS == 4) {
        float4 dq = reinterpret_cast<const float4 *>(P.depthBase)[pix];
        st.depth[0] = dq.x; st.depth[NS > 1 ? 1 : 0] = dq.y; st.depth[NS > 2 ? 2 : 0] = dq.z; st.depth[NS > 3 ? 3 : 0] = dq.w;
      } else st.depth[0] = P.depthBase[pix];
    }
    if (hasColor && !P.clearColorFlag) {
      if (NS == 4) {
        uint32_t c4[4];
        sglLoadMsColor(P, pix, c4);
        st.color[0] = c4[0]; st.color[NS > 1 ? 1 : 0] = c4[1]; st.color[NS > 2 ? 2 : 0] = c4[2]; st.color[NS > 3 ? 3 : 0] = c4[3];
      } else st.color[0] = reinterpret_cast<const uint32_t *>(P.colorBase)[pix];
    }
  };

  const uint32_t off = P.tileOffset[tile];
  uint32_t nList = P.tileOffset[tile + 1] - off;
  if (off + nList > P.binCapacity) nList = off < P.binCapacity ? P.binCapacity - off : 0;
  uint32_t nBig = *P.bigCount;
  if (nBig > P.bigCapacity) nBig = P.bigCapacity;
  const int tx0 = tx * SGL_TILE, ty0 = ty * SGL_TILE, tx1 = tx0 + SGL_TILE - 1, ty1 = ty0 + SGL_TILE - 1;
  unsigned int shaded = 0;

  // key windows: the common case (everything fits) is one window covering all keys
  uint32_t lo = 0;
  const uint32_t keyEnd = 0xFFFFFFFFu;
  bool fits = (nList + nBig) <= SGL_SORT_CAP;
  while (true) {
    uint32_t hi = keyEnd;
    // ---- gather candidates with lo <= key < hi
    while (true) {
      if (tid == 0) sCount = 0;
      __syncthreads();
      for (uint32_t i = tid; i < nList; i += SGL_TILE_THREADS) {
        uint32_t slot = P.binSlots[off + i];
        uint32_t key = P.primKeys[slot];
        if (key >= lo && key < hi) {
          int idx = atomicAdd(&sCount, 1);
          if (idx < SGL_SORT_CAP) { sKeys[idx] = key; sSlots[idx] = slot; }
        }
      }
      for (uint32_t i = tid; i < nBig; i += SGL_TILE_THREADS) {
        uint32_t slot = P.bigList[i];
        uint32_t key = P.primKeys[slot];
        if (key >= lo && key < hi) {
          const SglPrim &bp = P.prims[slot];
          if (bp.bx0 <= tx1 && bp.bx1 >= tx0 && bp.by0 <= ty1 && bp.by1 >= ty0 && sglPrimNearTile(bp, tx, ty)) {
            int idx = atomicAdd(&sCount, 1);
            if (idx < SGL_SORT_CAP) { sKeys[idx] = key; sSlots[idx] = slot; }
          }
        }
      }
      __syncthreads();
      if (sCount <= SGL_SORT_CAP) break;
      hi = lo + (hi - lo) / 2;          // too many: halve the key window and retry
      __syncthreads();
    }
    const int n = sCount;
    if (P.skipEmptyTiles && !loaded && n == 0 && hi == keyEnd) return;   // nothing touches this tile: it keeps its content
    loadState();
    sglSortTileList(sKeys, sSlots, n);
    // ---- process in order
    for (int b0 = 0; b0 < n; b0 += SGL_PRIM_BATCH) {
      int nb = n - b0 < SGL_PRIM_BATCH ? n - b0 : SGL_PRIM_BATCH;
      __syncthreads();
      {  // 64 records x 4 x uint4 = one 16-byte load per thread
        int r = tid >> 2, q = tid & 3;
        if (r < nb) {
          const uint4 *src = reinterpret_cast<const uint4 *>(P.prims + sSlots[b0 + r]);
          reinterpret_cast<uint4 *>(sPrims + r)[q] = __ldg(src + q);
        }
      }
      __syncthreads();
      if (inFb) {
        for (int k = 0; k < nb; k++) sglPixelPrim<NS>(P, sPrims[k], sSlots[b0 + k], px, py, st, hasColor, hasDepth);
      }
    }
    __syncthreads();
    if (fits || hi == keyEnd) break;
    lo = hi;
  }

  // ---- deferred shading, warp-coherent by draw: lanes shade together when their pending owner is in the draw
  //      of the warp's smallest pending slot (slots are laid out draw by draw)
  if (hasColor) {
    while (true) {
      // pick this lane's pending owner with the smallest slot
      uint32_t best = SGL_OWNER_NONE, bestSlot = 0xFFFFFFFFu;
#pragma unroll
      for (int s = 0; s < NS; s++) {
        uint32_t o = st.owner[s];
        if (o != SGL_OWNER_NONE && (o & 0x1fffffffu) < bestSlot) { best = o; bestSlot = o & 0x1fffffffu; }
      }
      uint32_t wmin = __reduce_min_sync(0xffffffffu, bestSlot);
      if (wmin == 0xFFFFFFFFu) break;
      uint32_t wdraw = P.prims[wmin].draw;
      if (best != SGL_OWNER_NONE && P.prims[bestSlot].draw == wdraw) {
        uint32_t c = sglPackColor(sglShadeSlot<NS>(P, bestSlot, (int) (best >> 29), px, py));
        shaded++;
#pragma unroll
        for (int s = 0; s < NS; s++)
          if (st.owner[s] == best) { st.color[s] = c; st.owner[s] = SGL_OWNER_NONE; }
      }
    }
  }

  // ---- write-back: per-sample colour/depth and the resolved colour (multiSampleResolve, RendererSoft.cpp:880-912)
  if (inFb) {
    if (hasDepth) {
      if (NS == 4) reinterpret_cast<float4 *>(P.depthBase)[pix] =
          make_float4(st.depth[0], st.depth[NS > 1 ? 1 : 0], st.depth[NS > 2 ? 2 : 0], st.depth[NS > 3 ? 3 : 0]);
      else P.depthBase[pix] = st.depth[0];
    }
    if (hasColor) {
      if (NS == 4) {
        {
          const uint32_t c4[4] = {st.color[0], st.color[NS > 1 ? 1 : 0], st.color[NS > 2 ? 2 : 0], st.color[NS > 3 ? 3 : 0]};
          sglStoreMsColor(P, pix, c4);
        }
        if (P.resolveBase) {
          uint32_t r = 0;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            uint32_t sum = 0;
#pragma unroll
            for (int s = 0; s < NS; s++) sum += (st.color[s] >> (8 * c)) & 0xffu;
            r |= (sum / NS) << (8 * c);     // u8vec4(sum / 4.f) truncation: exact integer division
          }
          reinterpret_cast<uint32_t *>(P.resolveBase)[pix] = r;
          if (P.mirrorBase) reinterpret_cast<uint32_t *>(P.mirrorBase)[pix] = r;
        }
      } else {
        reinterpret_cast<uint32_t *>(P.colorBase)[pix] = st.color[0];
        if (P.mirrorBase) reinterpret_cast<uint32_t *>(P.mirrorBase)[pix] = st.color[0];
      }
    }
  }
  // counters: one atomic per CTA
  if (tid == 0) sShaded = 0;
  __syncthreads();
  shaded = __reduce_add_sync(0xffffffffu, shaded);
  if ((tid & 31) == 0 && shaded) atomicAdd(&sShaded, shaded);
  __syncthreads();
  if (tid == 0 && sShaded) atomicAdd(P.fragCounters + (blockIdx.x & 31), (unsigned long long) sShaded);
}

#ifndef SGL_RASTER_ONLY
// ---------------------------------------------------------------------------------------------------------
// mip generation: BaseSampler::sampleBufferBilinear (SamplerSoft.h:241-252), one thread per output texel
__global__ void sglMipKernel(SglTexObj tex, int layer, int level) {
  int ow = sglLevelDim(tex.width, level), oh = sglLevelDim(tex.height, level);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  int iw = sglLevelDim(tex.width, level - 1), ih = sglLevelDim(tex.height, level - 1);
  float rx = xdiv((float) iw, (float) ow), ry = xdiv((float) ih, (float) oh);
  float u = xadd(xmul((float) x, rx), xmul(0.5f, rx)), v = xadd(xmul((float) y, ry), xmul(0.5f, ry));
  SglSampler s;
  s.tex = &tex;
  s.filter = SGL_FILTER_LINEAR;
  s.wrap = SGL_WRAP_CLAMP_TO_EDGE;
  s.border = 0;
  uint32_t t = sglPixelBilinear(s, layer, level - 1, u, v);
  uint32_t *dst = (uint32_t *) (tex.base + (size_t) layer * tex.layerStride + tex.levelOffset[level]);
  dst[sglTexelIndex(tex.layout, ow, x, y)] = t;
}

// linear host image <-> texture layout (upload / read-back of Tiled and Morton textures)
__global__ void sglRelayoutKernel(uint32_t *dst, const uint32_t *src, int w, int h, int layout, int toLayout) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  size_t lin = (size_t) y * w + x, til = sglTexelIndex(layout, w, x, y);
  if (toLayout) dst[til] = src[lin];
  else dst[lin] = src[til];
}

// ---- multi-GPU gather helpers ------------------------------------------------------------------------
// Tiles owned by `rank` in tile-index order <-> dense [n][SGL_TILE][SGL_TILE] RGBA8 staging buffer (what NCCL moves).
// grid = tiles, block = SGL_TILE_THREADS; `prefix[t]` = number of tiles owned by `rank` before tile t.
__global__ void __launch_bounds__(SGL_TILE_THREADS) sglTilePackKernel(uint32_t *image, uint32_t *packed, const uint8_t *owner,
                                                                     const uint32_t *prefix, int rank, int tilesX, int w, int h,
                                                                     int unpack) {
  const int tile = blockIdx.x;
  if (owner[tile] != rank) return;
  const int tx = tile % tilesX, ty = tile / tilesX;
  const int px = tx * SGL_TILE + (threadIdx.x & (SGL_TILE - 1)), py = ty * SGL_TILE + threadIdx.x / SGL_TILE;
  if (px >= w || py >= h) return;
  const size_t pi = (size_t) prefix[tile] * SGL_TILE_THREADS + threadIdx.x, ii = (size_t) py * w + px;
  if (unpack) image[ii] = packed[pi];
  else packed[pi] = image[ii];
}

// exclusive count of owned tiles (single CTA; tile maps are a few thousand entries)
__global__ void __launch_bounds__(1024) sglTileOwnerPrefixKernel(const uint8_t *owner, uint32_t *prefix, int nTiles, int rank) {
  __shared__ uint32_t sWarp[32];
  __shared__ uint32_t sCarry;
  __shared__ uint32_t sClass[SGL_TILE_CLASSES];   // this CTA is the only writer of the class lists
  if (threadIdx.x == 0) sCarry = 0;
  if (threadIdx.x < SGL_TILE_CLASSES) sClass[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nTiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < nTiles && owner[i] == rank) ? 1u : 0u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, v);
    const uint32_t excl = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 31) sWarp[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
      uint32_t w = sWarp[lane], inc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      sWarp[lane] = inc - w;
    }
    __syncthreads();
    const uint32_t e = sCarry + sWarp[warp] + excl;
    if (i < nTiles) prefix[i] = e;
    __syncthreads();
    if (threadIdx.x == 1023) sCarry = e + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) prefix[nTiles] = sCarry;
}

// Peer-memory flags of the direct-store gather: a release store at system scope after the frame's kernels (stream order),
// and a bounded spin on `count` consecutive 32-bit flags (stride 64 bytes) until each is >= value.
__global__ void sglPeerSignalKernel(uint32_t *flag, uint32_t value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void sglPeerWaitKernel(const uint32_t *flags, int count, uint32_t value, long long timeoutCycles,
                                  unsigned long long *counters) {
  const int i = threadIdx.x;
  if (i >= count) return;
  const uint32_t *f = flags + (size_t) i * 16;
  const long long t0 = clock64();
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if ((int32_t) (v - value) >= 0) break;
    if (clock64() - t0 > timeoutCycles) {
      atomicAdd(counters + 6, 1ull);   // peer wait timed out: surfaced by sgl_peer_check
      break;
    }
    __nanosleep(200);
  }
}

// Rank 0's side of the direct-store gather in ONE launch: wait until `count` done-flags reach `value`, then publish
// `value` into every peer's consumed-flag (peers[0] = rank 0 itself is skipped).
struct SglPeerPtrs {
  uint32_t *p[64];
};
__global__ void sglPeerCollectKernel(const uint32_t *flags, int count, uint32_t value, SglPeerPtrs peers, long long timeoutCycles,
                                     unsigned long long *counters) {
  const int i = threadIdx.x;
  if (i < count) {
    const uint32_t *f = flags + (size_t) i * 16;
    const long long t0 = clock64();
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int32_t) (v - value) >= 0) break;
      if (clock64() - t0 > timeoutCycles) {
        atomicAdd(counters + 6, 1ull);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (i >= 1 && i < count && peers.p[i]) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[i]), "r"(value) : "memory");
  }
}

__global__ void sglFill32Kernel(uint32_t *dst, uint32_t value, size_t n) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = value;
}

// Head of a pass's geometry chain in ONE launch: zero the counter block and fetch the draw records straight from the
// slot's pinned (UVA-mapped) staging buffer.  A memset node + a host-to-device copy node did the same through a copy
// engine, where the 5 KB of records can queue behind the 8 MB device-to-host copy of the previous frame's read-back.
__global__ void sglPassHeadKernel(uint4 *zero, size_t nZero16, uint4 *dst, const uint4 *srcHost, size_t nCopy16) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t k = i; k < nCopy16; k += stride) dst[k] = srcHost[k];
  for (size_t k = i; k < nZero16; k += stride) zero[k] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void sglCopy16Kernel(uint4 *dst, const uint4 *src, size_t n16) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (; i < n16; i += stride) dst[i] = src[i];
}

// device texture table: a renamed depth texture's entry learns its new backing store (sglcuda.cu, "Renaming")
__global__ void sglSetTexBaseKernel(SglTexObj *entry, uint8_t *base) { entry->base = base; }

// per-sample records of the pixels whose mask says "samples equal" (sgl_pixel.h, multisample colour storage), before the
// per-sample image leaves the library (read-back, sgl_texture_device_ptr)
__global__ void sglMsExpandKernel(uint4 *color, const uint32_t *resolve, uint8_t *mask, size_t n) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (mask[i] == 0) {
      const uint32_t r = resolve[i];
      color[i] = make_uint4(r, r, r, r);
      mask[i] = 1;
    }
}

// ---- known-answer-test kernels (wrap the device functions above) ---------------------------------------------
__global__ void sglKatBarycentricKernel(const float *tri, const float *xy, int n, float *bcOut, int *insideOut, float *zwOut) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SglPrim p;
  for (int v = 0; v < 3; v++)
    for (int c = 0; c < 4; c++) p.v[v][c] = tri[v * 4 + c];
  SglTriEdge e = sglTriEdge(p);
  float b0 = 0, b1 = 0, b2 = 0;
  bool in = false;
  if (!(fabsf(e.uz) < FLT_EPSILON)) in = sglBarycentric(e, xy[2 * i], xy[2 * i + 1], b0, b1, b2);
  bcOut[3 * i] = b0; bcOut[3 * i + 1] = b1; bcOut[3 * i + 2] = b2;
  insideOut[i] = in ? 1 : 0;
  zwOut[2 * i] = sglInterpZ(p, 2, b0, b1, b2);
  zwOut[2 * i + 1] = sglInterpZ(p, 3, b0, b1, b2);
}

__global__ void sglKatSampleKernel(const SglTexObj *textures, int tex, int filter, int wrap, uint32_t border, const float *coords,
                                   const float *lod, const int32_t *offs, int n, int splitPhase, uint32_t *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SglSampler s;
  s.tex = &textures[tex];
  s.filter = filter;
  s.wrap = wrap;
  s.border = border;
  const float l = lod ? lod[i] : 0.f;
  const int ox = offs ? offs[2 * i] : 0, oy = offs ? offs[2 * i + 1] : 0;
  int face = 0;
  float u, v;
  if (s.tex->layers == 6) sglCubeFace(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], face, u, v);
  else { u = coords[2 * i]; v = coords[2 * i + 1]; }
  if (splitPhase) out[i] = sglTapMix(sglTapIssue(sglTapView(s.tex, face, 0, wrap), u, v, ox, oy));
  else out[i] = sglTextureImpl(s, face, u, v, l, ox, oy);
}

__global__ void sglKatBlendKernel(SglRenderStates rs, const float *src, const float *dst, int n, float *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V4 s = v4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
  // destination is given as floats in [0,1]; the pipeline reads it back from RGBA8
  uint32_t dp = sglPackColor(v4(dst[4 * i], dst[4 * i + 1], dst[4 * i + 2], dst[4 * i + 3]));
  V4 r = sglBlend(rs, s, dp);
  out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

__global__ void sglKatDepthKernel(int func, const float *a, const float *b, int n, int *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = sglDepthTest(a[i], b[i], func) ? 1 : 0;
}
#endif  // SGL_RASTER_ONLY
