// Deferred ("visibility buffer") form of the tile rasteriser, used for every render pass whose draws are all opaque:
//   sglVisKernel<NS>    one CTA per 16x16 tile, one thread per pixel: ordered walk over the tile's primitives doing
//                       exact coverage + depth only; writes per-sample depth and per-sample OWNER (primitive slot +
//                       shading sample) -- no shading code, so it runs at 4 CTAs/SM instead of 2
//   sglShadeKernel<NS>  one thread per pixel: shades each distinct owner once (fragment shaders + texturing),
//                       packs, resolves MSAA and writes colour -- no barriers, no shared memory
// Exactness argument: DESIGN.md section 5 (an opaque fragment's colour is a pure function of (primitive, pixel), and the
// last fragment to pass the depth test per sample is the one whose colour survives in the reference).
// Passes with blending or with point/line draws of programs that have varyings use the fused sglRasterKernel instead.
#pragma once
#include <cuda_runtime.h>
#include "sgl_pixel.h"

#ifndef SGL_SORT_CAP
#define SGL_SORT_CAP 2048
#endif
#define SGL_VIS_BATCH 64
#ifndef SGL_SHADE_MIN_BLOCKS
#define SGL_SHADE_MIN_BLOCKS 4
#endif
// five CTAs per SM (48 registers, 64 B of spills) beat four (64 registers): the kernel waits on fixed-latency dependencies
// of the exact barycentric sequence, and the fifth CTA's warps fill them -- A/B on B200: config 2 vis 108 -> 103 us,
// config 3 338 -> 312 us, 2 M-triangle soup 3.71 -> 3.55 ms; six CTAs (40 registers) spill too much (105 us).  The shading
// kernel is the other way round: 5 / 6 CTAs per SM cost 160 / 166 us against 150 (its callee already spills at 64).
#ifndef SGL_VIS_MIN_BLOCKS
#define SGL_VIS_MIN_BLOCKS 5
#endif

// conservative "can primitive p write into tile (tx,ty)?" beyond the bbox overlap; MUST be the same function in the
// counting pass (sglSetupKernel), the fill pass (sglBinFillKernel) and the big-list scan of the tile kernels
__device__ __forceinline__ bool sglPrimNearTile(const SglPrim &p, int tx, int ty) {
  const uint32_t kind = p.flags & SGL_PF_KIND_MASK;
  if (kind == SGL_PK_LINE)
    return sglLineNearRect(p, tx * SGL_TILE, ty * SGL_TILE, tx * SGL_TILE + SGL_TILE - 1, ty * SGL_TILE + SGL_TILE - 1);
  if (kind == SGL_PK_TRIANGLE) {
    // every sample position of the tile lies within +-SGL_TILE/2 of its centre
    SglTriEdge e = sglTriEdge(p);
    const float h = 0.5f * SGL_TILE;
    return !sglTriSurelyOutside(e, (float) (tx * SGL_TILE) + h, (float) (ty * SGL_TILE) + h, h, h);
  }
  return true;
}

struct __align__(16) SglVisPrim {   // stream / shared-memory form of a primitive: record + per-triangle edge constants + slot
  SglPrim p;
  SglTriEdge e;
  uint32_t slot;
  float zBound;     // conservative Hi-Z bound of a depth-tested triangle (sglZBound), NaN = never cull
};

// Conservative bound of the depth any fragment of triangle p can have, on the side its depth test compares against:
// LESS / LEQUAL fragments are >= min(z0,z1,z2) - pad, GREATER / GEQUAL fragments <= max + pad.  The interpolated
// z = (b0 z0 + b1 z1) + b2 z2 has b_i >= 0 and sum b = 1 up to a few ulps (b0 = 1 - (b1 + b2)), so it leaves the vertex range
// by at most a few 1e-7 * scale; the pad is an order of magnitude above that.  Clamping to [0,1] keeps the bound valid
// against stored depths (all in [0,1]).  NaN (never cull) for everything else.
__device__ __forceinline__ float sglZBound(const SglPrim &p) {
  const uint32_t fl = p.flags;
  if ((fl & (SGL_PF_KIND_MASK | SGL_PF_DEPTH_TEST)) != (SGL_PK_TRIANGLE | SGL_PF_DEPTH_TEST)) return __int_as_float(0x7fc00000);
  const uint32_t f = (fl >> SGL_PF_DEPTH_FUNC_SHIFT) & 7u;
  const float z0 = p.v[0][2], z1 = p.v[1][2], z2 = p.v[2][2];
  const float pad = 4e-6f * (1.f + fmaxf(fabsf(z0), fmaxf(fabsf(z1), fabsf(z2))));
  if (f == 1u || f == 3u) return fminf(z0, fminf(z1, z2)) - pad;
  if (f == 4u || f == 6u) return fmaxf(z0, fmaxf(z1, z2)) + pad;
  return __int_as_float(0x7fc00000);
}
static_assert(sizeof(SglVisPrim) == 128, "stream entries are 128 bytes (cp.async.bulk: 16-byte granules)");

// ---- TMA-style bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ----------------------------------
__device__ __forceinline__ uint32_t sglSmemAddr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void sglMbarInit(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sglSmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sglMbarExpectTx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sglSmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sglBulkCopyG2S(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sglSmemAddr(dst)), "l"(src), "r"(bytes), "r"(sglSmemAddr(bar)) : "memory");
}
__device__ __forceinline__ void sglMbarWait(unsigned long long *bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nSGL_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra SGL_DONE_%=;\nbra SGL_WAIT_%=;\nSGL_DONE_%=:\n}"
               ::"r"(sglSmemAddr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void sglBitonicSortKV(uint32_t *keys, uint32_t *vals, int n /* power of two */) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint32_t a = keys[i], b = keys[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b; keys[ixj] = a;
            uint32_t t = vals[i]; vals[i] = vals[ixj]; vals[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Sorts (keys, vals)[0..n) by key; keys are unique.  n <= blockDim.x: rank sort (each thread counts the keys below its
// own: n broadcast shared-memory reads, two barriers); larger lists: bitonic network.  Ends with a barrier.
__device__ __forceinline__ void sglSortTileList(uint32_t *keys, uint32_t *vals, int n) {
  if (n <= 1) return;
  if (n <= (int) blockDim.x) {
    const int i = threadIdx.x;
    uint32_t k = 0, v = 0;
    int rank = 0;
    if (i < n) {
      k = keys[i];
      v = vals[i];
      for (int j = 0; j < n; j++) rank += keys[j] < k ? 1 : 0;
    }
    __syncthreads();
    if (i < n) { keys[rank] = k; vals[rank] = v; }
    __syncthreads();
    return;
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) { keys[i] = 0xFFFFFFFFu; vals[i] = 0; }
  __syncthreads();
  sglBitonicSortKV(keys, vals, n2);
}

// How many heavy tiles run as four quarter-tile items (splitCap = 0: none).  MSAA: the first splitCap of them.  One sample per
// pixel: all of them while they fit splitCap (few heavy tiles = stragglers), none otherwise (mostly heavy tiles = throughput).
// "Few" is relative to the tiles THIS rank renders (ownedTiles): a rank of a tile-sharded triangle soup owns an eighth of the
// tiles, all heavy -- measured against the whole frame they looked few and were all split (visibility 2.5x instead of 8x faster).
__device__ __forceinline__ uint32_t sglVisSplitCount(const SglPassParams &P, uint32_t heavyTiles, uint32_t ownedTiles) {
  const uint32_t cap = (uint32_t) P.splitCap;
  if (P.samples == 1) return (heavyTiles <= cap && heavyTiles * 8u <= ownedTiles) ? heavyTiles : 0u;
  return heavyTiles < cap ? heavyTiles : cap;
}

// Geometry-stage preparation of the pixel stage's input: one WARP per tile (in heavy-first work order) sorts the tile's bin
// by order key and writes (a) the sorted slots (tileSorted, read by the fused kernel), (b) the tile's packed record STREAM:
// 128-byte entries {primitive record, edge constants, slot} in submission order, contiguous, so that the visibility kernel
// fetches a whole batch with one cp.async.bulk, and (c) the work descriptor(s) of the tile.  Big primitives are in the bins
// already (sglBigBinKernel).  Tiles whose list is longer than SGL_TILE_SORT_CAP or does not fit the stream, and every tile
// while the residual big list is not empty, get SGL_TILE_UNSORTED and take the in-kernel path.
// grid = ceil(tiles / 8), block = 256.  (Compiled into the translation unit that launches it only.)
#ifdef SGL_WITH_TILE_SORT
#define SGL_TILE_SORT_WARPS 8
#define SGL_TILE_SORT_CAP 512     // longest list a warp sorts; longer ones are flagged SGL_TILE_UNSORTED
__device__ __forceinline__ void sglStreamStore(SglVisPrim *dst, const SglPrim *prims, uint32_t slot) {
  SglVisPrim v;
  const uint4 *src = reinterpret_cast<const uint4 *>(prims + slot);
#pragma unroll
  for (int q = 0; q < 4; q++) reinterpret_cast<uint4 *>(&v.p)[q] = __ldg(src + q);
  if ((v.p.flags & SGL_PF_KIND_MASK) == SGL_PK_TRIANGLE) v.e = sglTriEdge(v.p);
  else memset(&v.e, 0, sizeof(v.e));
  v.slot = slot;
  v.zBound = sglZBound(v.p);
#pragma unroll
  for (int q = 0; q < 8; q++) reinterpret_cast<uint4 *>(dst)[q] = reinterpret_cast<const uint4 *>(&v)[q];
}

__global__ void __launch_bounds__(32 * SGL_TILE_SORT_WARPS) sglTileSortKernel(SglPassParams P) {
  __shared__ uint32_t sKeys[SGL_TILE_SORT_WARPS][SGL_TILE_SORT_CAP];
  __shared__ uint32_t sSlots[SGL_TILE_SORT_WARPS][SGL_TILE_SORT_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t nTiles = (size_t) P.tilesX * P.tilesY;
  // work order: the class lists of sglTileScanKernel, heaviest class first
  uint32_t cnt[SGL_TILE_CLASSES];
#pragma unroll
  for (int c = 0; c < SGL_TILE_CLASSES; c++) cnt[c] = P.tileClassCount[c];
  const uint32_t owned = cnt[0] + cnt[1] + cnt[2] + cnt[3];
  const uint32_t o = (uint32_t) (blockIdx.x * SGL_TILE_SORT_WARPS + warp);
  if (o >= owned) return;
  const uint32_t heavyTiles = cnt[0] + cnt[1];
  const uint32_t split = sglVisSplitCount(P, heavyTiles, owned);
  const uint32_t ctaItems = 4u * split + (heavyTiles - split);
  int tile = -1;
  {
    uint32_t b = o;
#pragma unroll
    for (int c = 0; c < SGL_TILE_CLASSES; c++) {
      if (tile < 0 && b < cnt[c]) tile = (int) P.tileOrder[(size_t) c * nTiles + b];
      if (tile < 0) b -= cnt[c];
    }
  }
  const uint32_t off = P.tileOffset[tile];
  const uint32_t n = P.tileOffset[tile + 1] - off;
  bool prepared = n <= SGL_TILE_SORT_CAP && *P.bigCount == 0;   // residual big primitives must be merged by key in-kernel
  uint32_t streamOff = 0;
  if (prepared && n > 0) {
    if (lane == 0) streamOff = atomicAdd(P.streamCursor, n);
    streamOff = __shfl_sync(0xffffffffu, streamOff, 0);
    if (streamOff + n > P.streamCapacity || streamOff + n < streamOff) prepared = false;
  }
  if (prepared) {
    uint32_t *dst = P.tileSorted + off;
    SglVisPrim *sdst = P.stream + streamOff;
    if (n <= 32) {
      // the common case by far: one element per lane, rank by counting smaller keys across the warp
      uint32_t slot = 0, key = 0xFFFFFFFFu;
      if ((uint32_t) lane < n) {
        slot = P.binSlots[off + lane];
        key = P.primKeys[slot];
      }
      int rank = 0;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const uint32_t kj = __shfl_sync(0xffffffffu, key, j);
        rank += (kj < key) ? 1 : 0;
      }
      if ((uint32_t) lane < n) {
        dst[rank] = slot;
        sglStreamStore(sdst + rank, P.prims, slot);
      }
    } else {
      uint32_t *keys = sKeys[warp], *slots = sSlots[warp];
      for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t slot = P.binSlots[off + i];
        keys[i] = P.primKeys[slot];
        slots[i] = slot;
      }
      __syncwarp();
      // rank sort (keys are unique): each lane ranks its elements against the whole list, then scatters to global memory
      for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t k = keys[i];
        int rank = 0;
        for (uint32_t j = 0; j < n; j++) rank += keys[j] < k ? 1 : 0;
        dst[rank] = slots[i];
        sglStreamStore(sdst + rank, P.prims, slots[i]);
      }
    }
    if (lane == 0) P.tileSortedCount[tile] = n;
  }
  // work descriptors (sglVisWorkCounts): split heavy tiles = four quarter items, other heavy tiles = one CTA item, light
  // tiles = one warp item each
  const uint32_t none = 0xFFFFFFFFu;
  if (o < split) {
    if (lane < 4) {
      SglVisWork w;
      if (prepared) { w.tile = (uint32_t) tile; w.streamOff = streamOff; w.count = n; w.quarter = (uint32_t) lane; }
      else if (lane == 0) { w.tile = (uint32_t) tile; w.streamOff = 0; w.count = SGL_TILE_UNSORTED; w.quarter = none; }
      else { w.tile = none; w.streamOff = 0; w.count = 0; w.quarter = none; }
      P.work[4u * o + lane] = w;
    }
  } else if (lane == 0) {
    SglVisWork w;
    w.tile = (uint32_t) tile; w.streamOff = streamOff; w.count = prepared ? n : SGL_TILE_UNSORTED; w.quarter = none;
    P.work[o < heavyTiles ? 4u * split + (o - split) : ctaItems + (o - heavyTiles)] = w;
  }
}
#endif  // SGL_WITH_TILE_SORT

// blockIdx -> tile through the class lists of sglTileSortKernel (heaviest class first); -1 = nothing to do.
// With `splitCap` > 0 (MSAA visibility kernel) the first min(heavy, splitCap) tiles of classes 0-1 occupy FOUR blocks
// each: block b -> tile b / 4, quarter b % 4 (an 8x8 pixel quarter processed with one SAMPLE per lane); quarter = -1
// means "whole tile, one pixel per lane".
__device__ __forceinline__ int sglTileOfBlock(const SglPassParams &P, int block, int splitCap, int &quarter) {
  quarter = -1;
  if (!P.tileOrder) return (P.tileOwner && P.tileOwner[block] != P.rank) ? -1 : block;
  const size_t nTiles = (size_t) P.tilesX * P.tilesY;
  uint32_t cnt[SGL_TILE_CLASSES];
#pragma unroll
  for (int c = 0; c < SGL_TILE_CLASSES; c++) cnt[c] = P.tileClassCount[c];
  uint32_t b = (uint32_t) block;
  if (splitCap > 0) {
    uint32_t heavy = cnt[0] + cnt[1];
    if (heavy > (uint32_t) splitCap) heavy = (uint32_t) splitCap;
    if (b < 4u * heavy) {
      const uint32_t h = b >> 2;
      quarter = (int) (b & 3u);
      return (int) (h < cnt[0] ? P.tileOrder[h] : P.tileOrder[nTiles + (h - cnt[0])]);
    }
    b = b - 4u * heavy + heavy;      // position in the concatenated class lists, past the split tiles
  }
#pragma unroll
  for (int c = 0; c < SGL_TILE_CLASSES; c++) {
    if (b < cnt[c]) return (int) P.tileOrder[(size_t) c * nTiles + b];
    b -= cnt[c];
  }
  return -1;   // tiles of other ranks are in no class
}

// Sample-per-lane form of sglVisPixelPrim for the quarters of heavy MSAA tiles: the four lanes of a pixel each own one
// sample (depth + owner in one register each); geometric coverage of the pixel is exchanged with a ballot.  Same
// arithmetic as sglCoverTriangle / sglVisPixelPrim, evaluated per sample.  Must be called by all 32 lanes.
__device__ __forceinline__ void sglVisSamplePrim(const SglPassParams &P, const SglVisPrim &vp, uint32_t slot, int px, int py, int s,
                                                 int lane, bool inFb, float &depth, uint32_t &owner, bool hasColor, bool hasDepth) {
  const SglPrim &p = vp.p;
  const uint32_t flags = p.flags;
  const uint32_t kind = flags & SGL_PF_KIND_MASK;
  bool cand = inFb && !(px < p.bx0 || px > p.bx1 || py < p.by0 || py > p.by1);
  if (kind == SGL_PK_TRIANGLE) {   // warp-uniform
    if (cand && (flags & SGL_PF_IRREGULAR)) {
      const SglDrawRec &d = P.draws[p.draw];
      int q;
      if (!sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, q)) cand = false;
      else if (!sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, q)) cand = false;
    }
    const float fx = (float) px, fy = (float) py;
    if (cand && sglTriSurelyOutside(vp.e, fx + 0.5f, fy + 0.5f, 0.375f, 0.375f)) cand = false;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    bool in = false;
    // flat +0 depth and surely inside: all four samples and the centre covered, z == +0 (sglTriFlatZeroDepth)
    const bool flat = cand && sglTriFlatZeroDepth(p) && sglTriSurelyInside(vp.e, fx + 0.5f, fy + 0.5f, 0.375f, 0.375f);
    if (flat) in = true;
    else if (cand) {
      float ox, oy;
      sglSampleOffset(4, s, ox, oy);
      in = sglBarycentric(vp.e, xadd(ox, fx), xadd(oy, fy), b0, b1, b2);
    }
    const uint32_t g4 = (__ballot_sync(0xffffffffu, in) >> (lane & 28)) & 0xFu;   // geometric coverage of this pixel
    if (g4 == 0 || !in) return;
    int shadeIdx = 4;
    float z = 0.f;
    if (!flat) {
      float c0, c1, c2;
      shadeIdx = sglBarycentric(vp.e, xadd(fx, 0.5f), xadd(fy, 0.5f), c0, c1, c2) ? 4 : (__ffs(g4) - 1);
      z = sglInterpZ(p, 2, b0, b1, b2);
    }
    if (z < 0.f || z > 1.f) return;                    // depth-range clipping (multisample path)
    z = gclamp(z, 0.f, 1.f);
    if (flags & SGL_PF_DEPTH_TEST) {
      if (!hasDepth) return;
      if (!sglDepthTest(z, depth, (flags >> SGL_PF_DEPTH_FUNC_SHIFT) & 7)) return;
      if (flags & SGL_PF_DEPTH_MASK) depth = z;
    }
    owner = sglOwner(slot, shadeIdx);
    return;
  }
  if (!hasColor || !cand) return;
  SglPixelState<1> st;
  st.depth[0] = depth;
  uint32_t wrote = 0;
  if (kind == SGL_PK_POINT) wrote = sglFlatDepth<1>(p, p.v[0][2], hasDepth, st);
  else sglLineVisit(p, px, py, [&](float, float, float z) { wrote |= sglFlatDepth<1>(p, z, hasDepth, st); });
  depth = st.depth[0];
  if (wrote & 1u) owner = sglOwner(slot, 0);
}

// Single-sample triangle against one pixel WITHOUT the depth test: the part of sglCoverTriangle<1> / sglVisPixelPrim<1> that
// does not depend on what earlier primitives left in the pixel.  Returns "covered"; z = the clamped depth the test and the
// write use (one sample per pixel: no depth-range clipping, RendererSoft.cpp:797,871-874).
__device__ __forceinline__ bool sglVisEvalTriangle1(const SglPassParams &P, const SglVisPrim &vp, int px, int py, float &z) {
  const SglPrim &p = vp.p;
  if (px < p.bx0 || px > p.bx1 || py < p.by0 || py > p.by1) return false;
  if (p.flags & SGL_PF_IRREGULAR) {
    const SglDrawRec &d = P.draws[p.draw];
    int q;
    if (!sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, q)) return false;
    if (!sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, q)) return false;
  }
  const float fx = (float) px, fy = (float) py;
  if (sglTriSurelyOutside(vp.e, fx + 0.5f, fy + 0.5f, 0.f, 0.f)) return false;
  float b0, b1, b2;
  if (!sglBarycentric(vp.e, xadd(0.5f, fx), xadd(0.5f, fy), b0, b1, b2)) return false;
  z = gclamp(sglInterpZ(p, 2, b0, b1, b2), 0.f, 1.f);
  return true;
}

// one primitive against one pixel: coverage + depth, owners instead of colours
template<int NS>
__device__ __forceinline__ void sglVisPixelPrim(const SglPassParams &P, const SglVisPrim &vp, uint32_t slot, int px, int py,
                                                float (&depth)[NS], uint32_t (&owner)[NS], bool hasColor, bool hasDepth) {
  const SglPrim &p = vp.p;
  if (px < p.bx0 || px > p.bx1 || py < p.by0 || py > p.by1) return;
  const uint32_t flags = p.flags;
  const uint32_t kind = flags & SGL_PF_KIND_MASK;
  if (kind == SGL_PK_TRIANGLE) {
    if (flags & SGL_PF_IRREGULAR) {
      const SglDrawRec &d = P.draws[p.draw];
      int q;
      if (!sglAxisVisitedExact(min3f(p.v[0][0], p.v[1][0], p.v[2][0]), max3f(p.v[0][0], p.v[1][0], p.v[2][0]), d.vpW, px, q)) return;
      if (!sglAxisVisitedExact(min3f(p.v[0][1], p.v[1][1], p.v[2][1]), max3f(p.v[0][1], p.v[1][1], p.v[2][1]), d.vpH, py, q)) return;
    }
    float z[NS];
    int shadeIdx = 0;
    uint32_t mask = sglCoverTriangle<NS>(p, vp.e, px, py, depth, hasDepth, z, shadeIdx);
    if (!mask) return;
    const bool wr = (flags & SGL_PF_DEPTH_TEST) && (flags & SGL_PF_DEPTH_MASK);
    const uint32_t o = sglOwner(slot, shadeIdx);
#pragma unroll
    for (int s = 0; s < NS; s++)
      if ((mask >> s) & 1u) {
        if (wr) depth[s] = z[s];
        owner[s] = o;
      }
    return;
  }
  if (!hasColor) return;   // rasterizationPoint returns without a colour buffer (RendererSoft.cpp:636-640)
  SglPixelState<NS> st;    // only the depth part is live
#pragma unroll
  for (int s = 0; s < NS; s++) st.depth[s] = depth[s];
  uint32_t wrote = 0;
  if (kind == SGL_PK_POINT) {
    wrote = sglFlatDepth<NS>(p, p.v[0][2], hasDepth, st);
  } else {
    sglLineVisit(p, px, py, [&](float, float, float z) { wrote |= sglFlatDepth<NS>(p, z, hasDepth, st); });
  }
#pragma unroll
  for (int s = 0; s < NS; s++) {
    depth[s] = st.depth[s];
    if ((wrote >> s) & 1u) owner[s] = sglOwner(slot, 0);
  }
}

// Number of work items that need a whole CTA (tiles of the two heavy classes: four quarter items each while they fit
// splitCap, one whole-tile item each beyond it); the items after them are light tiles (< 40 primitives) that ONE WARP
// rasterises on its own.  Same arithmetic in sglTileSortKernel (writer) and sglVisKernel (reader).
__device__ __forceinline__ void sglVisWorkCounts(const SglPassParams &P, uint32_t &ctaItems, uint32_t &allItems, uint32_t &heavy, uint32_t &split) {
  uint32_t cnt[SGL_TILE_CLASSES];
#pragma unroll
  for (int c = 0; c < SGL_TILE_CLASSES; c++) cnt[c] = P.tileClassCount[c];
  heavy = cnt[0] + cnt[1];
  split = sglVisSplitCount(P, heavy, heavy + cnt[2] + cnt[3]);
  ctaItems = 4u * split + (heavy - split);
  allItems = ctaItems + cnt[2] + cnt[3];
}

// Visibility kernel: one CTA per work item of the heavy-first list the geometry stage prepared (a whole tile, or -- heavy
// MSAA tiles -- an 8x8 quarter with one sample per lane).  The CTA reads its 16-byte descriptor, then streams the item's
// packed records in batches of 64 through a double-buffered shared-memory ring: thread 0 issues the cp.async.bulk of batch
// b + 1 onto an mbarrier while all eight warps rasterise batch b.  Descriptor -> records is the whole dependent-load chain
// of a tile (it used to be class counts -> tile order -> offsets -> slots -> keys -> records, with a sort in between).
// Tiles without a prepared stream (SGL_TILE_UNSORTED: list longer than the sort cap, stream region full, residual big
// primitives) take the in-kernel gather / sort path.
template<int NS>
__global__ void __launch_bounds__(SGL_TILE_THREADS, SGL_VIS_MIN_BLOCKS) sglVisKernel(SglPassParams P) {
  constexpr int kRingBytes = 2 * SGL_VIS_BATCH * (int) sizeof(SglVisPrim);
  __shared__ __align__(128) unsigned char sPool[kRingBytes];
  SglVisPrim (*sRec)[SGL_VIS_BATCH] = reinterpret_cast<SglVisPrim (*)[SGL_VIS_BATCH]>(sPool);
  __shared__ uint32_t sKeys[SGL_SORT_CAP];      // in-kernel path only
  __shared__ uint32_t sSlots[SGL_SORT_CAP];
  __shared__ __align__(8) unsigned long long sBar[2];
  __shared__ int sCount;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const SglVisWork wk = P.work[blockIdx.x];     // in bounds for every CTA of the grid; valid iff blockIdx.x < allItems
  uint32_t ctaItems, allItems, heavyTiles, splitTiles;
  sglVisWorkCounts(P, ctaItems, allItems, heavyTiles, splitTiles);
  if (blockIdx.x >= allItems) return;
  if (wk.tile == 0xFFFFFFFFu) return;
  const int tile = (int) wk.tile;
  const bool streamed = wk.count != SGL_TILE_UNSORTED;
  const uint32_t count = streamed ? wk.count : 0u;
  if (tid == 0 && streamed && count > 0) {   // first two batches in flight before anything else
    sglMbarInit(&sBar[0], 1);
    sglMbarInit(&sBar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t n0 = count < SGL_VIS_BATCH ? count : SGL_VIS_BATCH;
    sglMbarExpectTx(&sBar[0], n0 * (uint32_t) sizeof(SglVisPrim));
    sglBulkCopyG2S(&sRec[0][0], P.stream + wk.streamOff, n0 * (uint32_t) sizeof(SglVisPrim), &sBar[0]);
    if (count > SGL_VIS_BATCH) {
      const uint32_t n1 = count - SGL_VIS_BATCH < SGL_VIS_BATCH ? count - SGL_VIS_BATCH : SGL_VIS_BATCH;
      sglMbarExpectTx(&sBar[1], n1 * (uint32_t) sizeof(SglVisPrim));
      sglBulkCopyG2S(&sRec[1][0], P.stream + wk.streamOff + SGL_VIS_BATCH, n1 * (uint32_t) sizeof(SglVisPrim), &sBar[1]);
    }
  }
  const bool hasColor = P.colorBase != nullptr, hasDepth = P.depthBase != nullptr;
  const int tx = tile % P.tilesX, ty = tile / P.tilesX;

  // ---- pixel state (registers): one pixel per lane (a warp owns an 8x4 block), or -- quarter items -- four lanes per pixel
  //      (a warp owns a 4x2 block, depth[0] / owner[0] only): with MSAA each of the four lanes owns one SAMPLE; with one sample
  //      per pixel the four lanes hold the same state and evaluate four consecutive surviving triangles at once, whose
  //      results are then applied in submission order (the depth test is the only step that depends on the order)
  const bool quarterMode = wk.quarter != 0xFFFFFFFFu;
  int wbx, wby, px, py, smp = 0;
  if (quarterMode) {
    wbx = tx * SGL_TILE + (int) (wk.quarter & 1u) * 8 + (warp & 1) * 4;
    wby = ty * SGL_TILE + (int) (wk.quarter >> 1) * 8 + (warp >> 1) * 2;
    px = wbx + ((lane >> 2) & 3);
    py = wby + (lane >> 4);
    smp = lane & 3;
  } else {
    wbx = tx * SGL_TILE + (warp & 1) * 8;
    wby = ty * SGL_TILE + (warp >> 1) * 4;
    px = wbx + (lane & 7);
    py = wby + (lane >> 3);
  }
  const bool inFb = px < P.fbW && py < P.fbH;
  const size_t pix = (size_t) py * P.fbW + px;
  float depth[NS];
  uint32_t owner[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) { depth[s] = P.clearDepth; owner[s] = SGL_OWNER_NONE; }
  if (inFb && hasDepth && !P.clearDepthFlag) {
    if (quarterMode) depth[0] = NS == 4 ? P.depthBase[pix * 4 + smp] : P.depthBase[pix];
    else if (NS == 4) {
      float4 dq = reinterpret_cast<const float4 *>(P.depthBase)[pix];
      depth[0] = dq.x; depth[NS > 1 ? 1 : 0] = dq.y; depth[NS > 2 ? 2 : 0] = dq.z; depth[NS > 3 ? 3 : 0] = dq.w;
    } else depth[0] = P.depthBase[pix];
  }
  const bool stampIt = P.tileTimes && tid == 0 && (!quarterMode || wk.quarter == 0);
  if (stampIt) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    P.tileTimes[2 * tile] = t;
  }

  // Hi-Z pays for itself where there is depth complexity: lists of a few primitives (sky, floor, a line) skip its bookkeeping
  const bool hiz = hasDepth && (!streamed || count >= 32u);
  // cull one staged batch (<= 64 records) against the warp's pixel block, then every pixel (sample) visits the survivors
  auto processBatch = [&](const SglVisPrim *recs, int n) {
    uint32_t rel[2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int idx = hh * 32 + lane;
      bool r = false;
      if (idx < n) {
        const SglVisPrim &vp = recs[idx];
        const uint32_t kind = vp.p.flags & SGL_PF_KIND_MASK;
        if (quarterMode) {
          r = vp.p.bx0 <= wbx + 3 && vp.p.bx1 >= wbx && vp.p.by0 <= wby + 1 && vp.p.by1 >= wby;
          if (r) {
            if (kind == SGL_PK_TRIANGLE) r = !sglTriSurelyOutside(vp.e, (float) wbx + 2.f, (float) wby + 1.f, 1.875f, 0.875f);
            else if (kind == SGL_PK_LINE) r = sglLineNearRect(vp.p, wbx, wby, wbx + 3, wby + 1);
          }
        } else {
          r = vp.p.bx0 <= wbx + 7 && vp.p.bx1 >= wbx && vp.p.by0 <= wby + 3 && vp.p.by1 >= wby;
          if (r) {
            if (kind == SGL_PK_TRIANGLE) r = !sglTriSurelyOutside(vp.e, (float) wbx + 4.f, (float) wby + 2.f, 3.875f, 1.875f);
            else if (kind == SGL_PK_LINE) r = sglLineNearRect(vp.p, wbx, wby, wbx + 7, wby + 3);
          }
        }
      }
      rel[hh] = __ballot_sync(0xffffffffu, r);
    }
    // Hi-Z over the warp's pixel block: a depth-tested triangle whose conservative depth bound (zBound) cannot pass against
    // the farthest (LESS / LEQUAL) resp. nearest (GREATER / GEQUAL) depth currently stored in the block fails the test at
    // every sample of the block -- skipping it changes nothing, in submission order or not.  Stored depths are in [0,1]:
    // their bit patterns order like the values (-0 mapped to +0); lanes outside the framebuffer are neutral.
    float wFar, wNear;
    auto blockDepthRange = [&]() {
      uint32_t hi = 0u, lo = 0xFFFFFFFFu;
      if (inFb) {
        if (NS == 4 && quarterMode) {
          const uint32_t b = __float_as_uint(depth[0]) == 0x80000000u ? 0u : __float_as_uint(depth[0]);
          hi = lo = b;
        } else {
#pragma unroll
          for (int s2 = 0; s2 < NS; s2++) {
            const uint32_t b = __float_as_uint(depth[s2]) == 0x80000000u ? 0u : __float_as_uint(depth[s2]);
            hi = b > hi ? b : hi;
            lo = b < lo ? b : lo;
          }
        }
      }
      wFar = __uint_as_float(__reduce_max_sync(0xffffffffu, hi));
      wNear = __uint_as_float(__reduce_min_sync(0xffffffffu, lo));
    };
    if (hiz) blockDepthRange();
    // single-sample quarter items: survivors wait in `pack` (8 bits each) until four triangles are there
    uint32_t pack = 0;
    int g = 0;
    auto runGroup = [&]() {
      if (g == 0) return;
      bool in = false;
      float z = 0.f;
      if (smp < g && inFb) in = sglVisEvalTriangle1(P, recs[(pack >> (8 * smp)) & 0xffu], px, py, z);
      bool wroteDepth = false;
      for (int j = 0; j < g; j++) {                      // g is warp-uniform
        const int srcLane = (lane & 28) | j;
        const bool inj = __shfl_sync(0xffffffffu, in ? 1 : 0, srcLane) != 0;
        const float zj = __shfl_sync(0xffffffffu, z, srcLane);
        const SglVisPrim &vj = recs[(pack >> (8 * j)) & 0xffu];
        const uint32_t fl = vj.p.flags;
        wroteDepth = wroteDepth || ((fl & SGL_PF_DEPTH_TEST) && (fl & SGL_PF_DEPTH_MASK));
        if (!inj) continue;
        if (fl & SGL_PF_DEPTH_TEST) {
          if (!hasDepth || !sglDepthTest(zj, depth[0], (fl >> SGL_PF_DEPTH_FUNC_SHIFT) & 7)) continue;
          if (fl & SGL_PF_DEPTH_MASK) depth[0] = zj;
        }
        owner[0] = sglOwner(vj.slot, 0);
      }
      pack = 0;
      g = 0;
      if (hiz && wroteDepth) blockDepthRange();
    };
    // ONE copy of the per-primitive body for both halves of the batch (it used to be unrolled over the two ballot words: the
    // kernel is 170 KB of SASS and stalled 2.4 cycles per issue waiting for instructions)
    {
      unsigned long long m = (unsigned long long) rel[0] | ((unsigned long long) rel[1] << 32);
      while (m) {
        const int k = __ffsll((long long) m) - 1;
        m &= m - 1;
        if (hiz) {
          const float zb = recs[k].zBound;                       // NaN: every comparison below is false
          const uint32_t f = (recs[k].p.flags >> SGL_PF_DEPTH_FUNC_SHIFT) & 7u;
          if ((f == 1u && zb >= wFar) || (f == 3u && zb > wFar) || (f == 4u && zb <= wNear) || (f == 6u && zb < wNear)) continue;
        }
        if (NS == 1 && quarterMode) {
          if ((recs[k].p.flags & SGL_PF_KIND_MASK) == SGL_PK_TRIANGLE) {
            pack |= (uint32_t) k << (8 * g);
            if (++g == 4) runGroup();
            continue;
          }
          runGroup();      // a point / line: in order after the triangles before it; the four lanes of a pixel do the same work
        }
        if (NS == 4 && quarterMode) sglVisSamplePrim(P, recs[k], recs[k].slot, px, py, smp, lane, inFb, depth[0], owner[0], hasColor, hasDepth);
        else if (inFb) sglVisPixelPrim<NS>(P, recs[k], recs[k].slot, px, py, depth, owner, hasColor, hasDepth);
        if (hiz && (recs[k].p.flags & SGL_PF_DEPTH_MASK)) blockDepthRange();
      }
    }
    if (NS == 1 && quarterMode) runGroup();
  };

  if (streamed) {
    uint32_t phase0 = 0, phase1 = 0;
    if (count > 0) __syncthreads();     // thread 0's mbarrier initialisation is visible before anybody waits on them
    for (uint32_t b0 = 0, u = 0; b0 < count; b0 += SGL_VIS_BATCH, u++) {
      const int slot = (int) (u & 1u);
      const int nb = count - b0 < SGL_VIS_BATCH ? (int) (count - b0) : SGL_VIS_BATCH;
      if (slot == 0) { sglMbarWait(&sBar[0], phase0); phase0 ^= 1u; }
      else { sglMbarWait(&sBar[1], phase1); phase1 ^= 1u; }
      processBatch(&sRec[slot][0], nb);
      const uint32_t b2 = b0 + 2 * SGL_VIS_BATCH;      // this ring slot's next batch
      if (b2 < count) {
        __syncthreads();                               // everybody is done reading the slot
        if (tid == 0) {
          const uint32_t n2 = count - b2 < SGL_VIS_BATCH ? count - b2 : SGL_VIS_BATCH;
          sglMbarExpectTx(&sBar[slot], n2 * (uint32_t) sizeof(SglVisPrim));
          sglBulkCopyG2S(&sRec[slot][0], P.stream + wk.streamOff + b2, n2 * (uint32_t) sizeof(SglVisPrim), &sBar[slot]);
        }
      }
    }
  } else {
    // ---- in-kernel path: gather the tile's bin and the residual big primitives in key windows, sort, stage, process
    SglVisPrim *stage = &sRec[0][0];
    const uint32_t off = P.tileOffset[tile];
    uint32_t nList = P.tileOffset[tile + 1] - off;
    if (off + nList > P.binCapacity) nList = off < P.binCapacity ? P.binCapacity - off : 0;
    uint32_t nBig = *P.bigCount;
    if (nBig > P.bigCapacity) nBig = P.bigCapacity;
    const int tx0 = tx * SGL_TILE, ty0 = ty * SGL_TILE, tx1 = tx0 + SGL_TILE - 1, ty1 = ty0 + SGL_TILE - 1;
    uint32_t lo = 0;
    const uint32_t keyEnd = 0xFFFFFFFFu;
    const bool fits = (nList + nBig) <= SGL_SORT_CAP;
    while (true) {
      uint32_t hi = keyEnd;
      while (true) {   // gather candidates with lo <= key < hi
        if (tid == 0) sCount = 0;
        __syncthreads();
        for (uint32_t i = tid; i < nList; i += SGL_TILE_THREADS) {
          uint32_t s2 = P.binSlots[off + i];
          uint32_t key = P.primKeys[s2];
          if (key >= lo && key < hi) {
            int idx = atomicAdd(&sCount, 1);
            if (idx < SGL_SORT_CAP) { sKeys[idx] = key; sSlots[idx] = s2; }
          }
        }
        for (uint32_t i = tid; i < nBig; i += SGL_TILE_THREADS) {
          uint32_t s2 = P.bigList[i];
          uint32_t key = P.primKeys[s2];
          if (key >= lo && key < hi) {
            const SglPrim &bp = P.prims[s2];
            if (bp.bx0 <= tx1 && bp.bx1 >= tx0 && bp.by0 <= ty1 && bp.by1 >= ty0 && sglPrimNearTile(bp, tx, ty)) {
              int idx = atomicAdd(&sCount, 1);
              if (idx < SGL_SORT_CAP) { sKeys[idx] = key; sSlots[idx] = s2; }
            }
          }
        }
        __syncthreads();
        if (sCount <= SGL_SORT_CAP) break;
        hi = lo + (hi - lo) / 2;          // too many: halve the key window and retry
        __syncthreads();
      }
      const int n = sCount;
      sglSortTileList(sKeys, sSlots, n);
      for (int b0 = 0; b0 < n; b0 += SGL_VIS_BATCH) {   // process in order
        const int nbb = n - b0 < SGL_VIS_BATCH ? n - b0 : SGL_VIS_BATCH;
        __syncthreads();
        {  // 64 records x 4 x uint4 = one 16-byte load per thread, then one thread per record derives the edge constants
          const int r = tid >> 2, q = tid & 3;
          if (r < nbb) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.prims + sSlots[b0 + r]);
            reinterpret_cast<uint4 *>(&stage[r].p)[q] = __ldg(src + q);
          }
        }
        __syncthreads();
        if (tid < nbb) {
          if ((stage[tid].p.flags & SGL_PF_KIND_MASK) == SGL_PK_TRIANGLE) stage[tid].e = sglTriEdge(stage[tid].p);
          stage[tid].slot = sSlots[b0 + tid];
          stage[tid].zBound = sglZBound(stage[tid].p);
        }
        __syncthreads();
        processBatch(stage, nbb);
      }
      __syncthreads();
      if (fits || hi == keyEnd) break;
      lo = hi;
    }
  }

  if (inFb) {
    if (quarterMode && NS == 4) {
      if (hasDepth) P.depthBase[pix * 4 + smp] = depth[0];
      if (hasColor) P.vis[pix * 4 + smp] = owner[0];
    } else if (quarterMode) {
      if (smp == 0) {      // the four lanes of the pixel hold the same state
        if (hasDepth) P.depthBase[pix] = depth[0];
        if (hasColor) P.vis[pix] = owner[0];
      }
    } else {
      if (hasDepth) {
        if (NS == 4) reinterpret_cast<float4 *>(P.depthBase)[pix] =
            make_float4(depth[0], depth[NS > 1 ? 1 : 0], depth[NS > 2 ? 2 : 0], depth[NS > 3 ? 3 : 0]);
        else P.depthBase[pix] = depth[0];
      }
      if (hasColor) {
        if (NS == 4) reinterpret_cast<uint4 *>(P.vis)[pix] =
            make_uint4(owner[0], owner[NS > 1 ? 1 : 0], owner[NS > 2 ? 2 : 0], owner[NS > 3 ? 3 : 0]);
        else P.vis[pix] = owner[0];
      }
    }
  }
  if (stampIt) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    P.tileTimes[2 * tile + 1] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// grid = tiles (same ownership rule as the visibility kernel); a warp shades an 8x4 pixel block
template<int NS>
__global__ void __launch_bounds__(SGL_TILE_THREADS, SGL_SHADE_MIN_BLOCKS) sglShadeKernel(SglPassParams P) {
  const int tile = blockIdx.x;
  if (P.tileOwner && P.tileOwner[tile] != P.rank) return;
  const int tx = tile % P.tilesX, ty = tile / P.tilesX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int px = tx * SGL_TILE + (warp & 1) * 8 + (lane & 7);
  const int py = ty * SGL_TILE + (warp >> 1) * 4 + (lane >> 3);
  const bool inFb = px < P.fbW && py < P.fbH;
  const size_t pix = (size_t) py * P.fbW + px;

  uint32_t owner[NS], color[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) { owner[s] = SGL_OWNER_NONE; color[s] = P.clearColor; }
  if (inFb) {
    if (NS == 4) {
      uint4 o = reinterpret_cast<const uint4 *>(P.vis)[pix];
      owner[0] = o.x; owner[NS > 1 ? 1 : 0] = o.y; owner[NS > 2 ? 2 : 0] = o.z; owner[NS > 3 ? 3 : 0] = o.w;
      if (!P.clearColorFlag) {
        uint32_t c4[4];
        sglLoadMsColor(P, pix, c4);
        color[0] = c4[0]; color[NS > 1 ? 1 : 0] = c4[1]; color[NS > 2 ? 2 : 0] = c4[2]; color[NS > 3 ? 3 : 0] = c4[3];
      }
    } else {
      owner[0] = P.vis[pix];
      if (!P.clearColorFlag) color[0] = reinterpret_cast<const uint32_t *>(P.colorBase)[pix];
    }
  }

  // warp-coherent by draw: lanes shade together when their pending owner belongs to the draw of the warp's smallest
  // pending slot (slots are laid out draw by draw)
  unsigned int shaded = 0;
  while (true) {
    uint32_t best = SGL_OWNER_NONE, bestSlot = 0xFFFFFFFFu;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      uint32_t o = owner[s];
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
        if (owner[s] == best) { color[s] = c; owner[s] = SGL_OWNER_NONE; }
    }
  }

  if (inFb) {
    if (NS == 4) {
      {
        const uint32_t c4[4] = {color[0], color[NS > 1 ? 1 : 0], color[NS > 2 ? 2 : 0], color[NS > 3 ? 3 : 0]};
        sglStoreMsColor(P, pix, c4);
      }
      if (P.resolveBase) {   // multiSampleResolve (RendererSoft.cpp:880-912): u8vec4(sum / 4.f) == exact integer division
        uint32_t r = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          uint32_t sum = 0;
#pragma unroll
          for (int s = 0; s < NS; s++) sum += (color[s] >> (8 * c)) & 0xffu;
          r |= (sum / NS) << (8 * c);
        }
        reinterpret_cast<uint32_t *>(P.resolveBase)[pix] = r;
        if (P.mirrorBase) reinterpret_cast<uint32_t *>(P.mirrorBase)[pix] = r;
      }
    } else {
      reinterpret_cast<uint32_t *>(P.colorBase)[pix] = color[0];
      if (P.mirrorBase) reinterpret_cast<uint32_t *>(P.mirrorBase)[pix] = color[0];
    }
  }
  shaded = __reduce_add_sync(0xffffffffu, shaded);
  if (lane == 0 && shaded) atomicAdd(P.fragCounters + ((blockIdx.x + warp) & 31), (unsigned long long) shaded);
}
