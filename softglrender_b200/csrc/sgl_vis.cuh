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
#ifndef SGL_VIS_MIN_BLOCKS
#define SGL_VIS_MIN_BLOCKS 4
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

struct __align__(16) SglVisPrim {   // shared-memory form of a primitive: record + per-triangle edge constants
  SglPrim p;
  SglTriEdge e;
  float pad[2];
};

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

// Geometry-stage preparation of the tile lists: gathers a tile's bin plus the big primitives that can touch it, sorts by
// order key and stores the slots, so that the pixel-stage kernels stream a ready list instead of running the gather /
// sort latency chain with 256 mostly idle threads per tile.  Tiles whose list does not fit are flagged SGL_TILE_UNSORTED
// and take the in-kernel path.  grid = tiles, block = 128.  (Compiled into the translation unit that launches it only.)
#ifdef SGL_WITH_TILE_SORT
#define SGL_TILE_SORT_WARPS 8
#define SGL_TILE_SORT_CAP 512     // longest list a warp sorts; longer ones are flagged SGL_TILE_UNSORTED
// one WARP per heavy tile: grid = ceil(splitCap / 8), block = 256
__global__ void __launch_bounds__(32 * SGL_TILE_SORT_WARPS) sglTileSortKernel(SglPassParams P) {
  __shared__ uint32_t sKeys[SGL_TILE_SORT_WARPS][SGL_TILE_SORT_CAP];
  __shared__ uint32_t sSlots[SGL_TILE_SORT_WARPS][SGL_TILE_SORT_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nTiles = P.tilesX * P.tilesY;
  // only the heavy classes (0 and 1 of sglTileScanKernel) get a prepared list: they are the tiles the visibility kernel
  // splits into quarter-tile CTAs; light tiles keep the in-kernel gather (thousands of them hide each other's latency)
  const uint32_t h = (uint32_t) (blockIdx.x * SGL_TILE_SORT_WARPS + warp);
  const uint32_t c0 = P.tileClassCount[0], c1 = P.tileClassCount[1];
  if (h >= c0 + c1) return;
  const int tile = (int) (h < c0 ? P.tileOrder[h] : P.tileOrder[(size_t) nTiles + (h - c0)]);
  uint32_t *keys = sKeys[warp], *slots = sSlots[warp];
  const int tx = tile % P.tilesX, ty = tile / P.tilesX;
  const uint32_t off = P.tileOffset[tile];
  uint32_t nList = P.tileOffset[tile + 1] - off;
  if (off + nList > P.binCapacity) nList = off < P.binCapacity ? P.binCapacity - off : 0;
  uint32_t nBig = *P.bigCount;
  if (nBig > P.bigCapacity) nBig = P.bigCapacity;
  bool fits = nList <= SGL_TILE_SORT_CAP - SGL_BIG_PER_TILE;
  int n = 0;
  if (fits) {
    for (uint32_t i = lane; i < nList; i += 32) {
      const uint32_t slot = P.binSlots[off + i];
      keys[i] = P.primKeys[slot];
      slots[i] = slot;
    }
    n = (int) nList;
    const int tx0 = tx * SGL_TILE, ty0 = ty * SGL_TILE, tx1 = tx0 + SGL_TILE - 1, ty1 = ty0 + SGL_TILE - 1;
    int big = 0;
    for (uint32_t i0 = 0; i0 < nBig; i0 += 32) {      // warp-uniform trip count
      const uint32_t i = i0 + lane;
      bool near = false;
      uint32_t slot = 0;
      if (i < nBig) {
        slot = P.bigList[i];
        const SglPrim &bp = P.prims[slot];
        near = bp.bx0 <= tx1 && bp.bx1 >= tx0 && bp.by0 <= ty1 && bp.by1 >= ty0 && sglPrimNearTile(bp, tx, ty);
      }
      const uint32_t m = __ballot_sync(0xffffffffu, near);
      const int pos = big + __popc(m & ((1u << lane) - 1u));
      if (near && pos < SGL_BIG_PER_TILE) {
        keys[n + pos] = P.primKeys[slot];
        slots[n + pos] = slot;
      }
      big += __popc(m);
    }
    if (big > SGL_BIG_PER_TILE) fits = false;
    n += big;
  }
  if (!fits) return;   // stays SGL_TILE_UNSORTED
  __syncwarp();
  // rank sort (keys are unique): each lane ranks its elements against the whole list, then scatters to global memory
  uint32_t *dst = P.tileSorted + off + (size_t) tile * SGL_BIG_PER_TILE;
  for (int i = lane; i < n; i += 32) {
    const uint32_t k = keys[i];
    int rank = 0;
    for (int j = 0; j < n; j++) rank += keys[j] < k ? 1 : 0;
    dst[rank] = slots[i];
  }
  if (lane == 0) P.tileSortedCount[tile] = (uint32_t) n;
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

template<int NS>
__global__ void __launch_bounds__(SGL_TILE_THREADS, SGL_VIS_MIN_BLOCKS) sglVisKernel(SglPassParams P) {
  __shared__ uint32_t sKeys[SGL_SORT_CAP];
  __shared__ uint32_t sSlots[SGL_SORT_CAP];
  __shared__ SglVisPrim sPrims[SGL_VIS_BATCH];
  __shared__ int sCount;

  int quarter;
  const int tile = sglTileOfBlock(P, blockIdx.x, NS == 4 ? P.splitCap : 0, quarter);
  if (tile < 0) return;
  const int tx = tile % P.tilesX, ty = tile / P.tilesX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (P.tileTimes && tid == 0 && quarter <= 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    P.tileTimes[2 * tile] = t;
  }
  if (NS == 4 && quarter >= 0) {
    const uint32_t pre = P.tileSortedCount[tile];
    if (pre == SGL_TILE_UNSORTED) {
      if (quarter != 0) return;       // list not prepared: quarter 0 takes the whole tile below
    } else {
      // ---- heavy tile, one 8x8 pixel quarter per CTA, one SAMPLE per lane: a warp owns a 4x2 pixel block
      const int wbx = tx * SGL_TILE + (quarter & 1) * 8 + (warp & 1) * 4, wby = ty * SGL_TILE + (quarter >> 1) * 8 + (warp >> 1) * 2;
      const int px = wbx + ((lane >> 2) & 3), py = wby + (lane >> 4), smp = lane & 3;
      const bool inFb = px < P.fbW && py < P.fbH;
      const bool hasColor = P.colorBase != nullptr, hasDepth = P.depthBase != nullptr;
      const size_t idx = ((size_t) py * P.fbW + px) * 4 + smp;
      float depth = P.clearDepth;
      uint32_t owner = SGL_OWNER_NONE;
      if (inFb && hasDepth && !P.clearDepthFlag) depth = P.depthBase[idx];
      const uint32_t *list = P.tileSorted + P.tileOffset[tile] + (size_t) tile * SGL_BIG_PER_TILE;
      for (uint32_t b0 = 0; b0 < pre; b0 += SGL_VIS_BATCH) {
        const int nb = pre - b0 < SGL_VIS_BATCH ? (int) (pre - b0) : SGL_VIS_BATCH;
        __syncthreads();
        {
          const int r = tid >> 2, q = tid & 3;
          if (r < nb) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.prims + __ldg(list + b0 + r));
            reinterpret_cast<uint4 *>(&sPrims[r].p)[q] = __ldg(src + q);
          }
          if (tid < nb) sSlots[tid] = __ldg(list + b0 + tid);
        }
        __syncthreads();
        if (tid < nb && (sPrims[tid].p.flags & SGL_PF_KIND_MASK) == SGL_PK_TRIANGLE) sPrims[tid].e = sglTriEdge(sPrims[tid].p);
        __syncthreads();
        uint32_t rel[2];
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int i = hh * 32 + lane;
          bool r = false;
          if (i < nb) {
            const SglVisPrim &vp = sPrims[i];
            r = vp.p.bx0 <= wbx + 3 && vp.p.bx1 >= wbx && vp.p.by0 <= wby + 1 && vp.p.by1 >= wby;
            if (r) {
              const uint32_t kind = vp.p.flags & SGL_PF_KIND_MASK;
              if (kind == SGL_PK_TRIANGLE) r = !sglTriSurelyOutside(vp.e, (float) wbx + 2.f, (float) wby + 1.f, 1.875f, 0.875f);
              else if (kind == SGL_PK_LINE) r = sglLineNearRect(vp.p, wbx, wby, wbx + 3, wby + 1);
            }
          }
          rel[hh] = __ballot_sync(0xffffffffu, r);
        }
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          uint32_t m = rel[hh];
          while (m) {
            const int k = hh * 32 + __ffs(m) - 1;
            m &= m - 1;
            sglVisSamplePrim(P, sPrims[k], sSlots[k], px, py, smp, lane, inFb, depth, owner, hasColor, hasDepth);
          }
        }
      }
      if (inFb) {
        if (hasDepth) P.depthBase[idx] = depth;
        if (hasColor) P.vis[idx] = owner;
      }
      if (P.tileTimes) {
        __syncthreads();
        if (tid == 0 && quarter == 0) {
          unsigned long long t;
          asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
          P.tileTimes[2 * tile + 1] = t;
        }
      }
      return;
    }
  }
  // a warp owns an 8x4 pixel block of the tile (same mapping as the shading kernel): small triangles then concern one
  // or two warps of the CTA instead of most 16x2 strips
  const int wbx = tx * SGL_TILE + (warp & 1) * 8, wby = ty * SGL_TILE + (warp >> 1) * 4;
  const int px = wbx + (lane & 7);
  const int py = wby + (lane >> 3);
  const bool inFb = px < P.fbW && py < P.fbH;
  const bool hasColor = P.colorBase != nullptr, hasDepth = P.depthBase != nullptr;
  const size_t pix = (size_t) py * P.fbW + px;

  float depth[NS];
  uint32_t owner[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) { depth[s] = P.clearDepth; owner[s] = SGL_OWNER_NONE; }
  if (inFb && hasDepth && !P.clearDepthFlag) {
    if (NS == 4) {
      float4 dq = reinterpret_cast<const float4 *>(P.depthBase)[pix];
      depth[0] = dq.x; depth[NS > 1 ? 1 : 0] = dq.y; depth[NS > 2 ? 2 : 0] = dq.z; depth[NS > 3 ? 3 : 0] = dq.w;
    } else depth[0] = P.depthBase[pix];
  }

  const uint32_t off = P.tileOffset[tile];

  // one batch of <= SGL_VIS_BATCH records whose slots sit in slots[0..nb): stage the records in shared memory, derive the
  // edge constants, cull per warp, then every pixel visits the survivors in order
  auto runBatch = [&](const uint32_t *slots, int nb) {
    {  // 64 records x 4 x uint4 = one 16-byte load per thread, then one thread per record derives the edge constants
      const int r = tid >> 2, q = tid & 3;
      if (r < nb) {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.prims + slots[r]);
        reinterpret_cast<uint4 *>(&sPrims[r].p)[q] = __ldg(src + q);
      }
    }
    __syncthreads();
    if (tid < nb && (sPrims[tid].p.flags & SGL_PF_KIND_MASK) == SGL_PK_TRIANGLE) sPrims[tid].e = sglTriEdge(sPrims[tid].p);
    __syncthreads();
    // warp-level cull: each lane tests two records of the batch against the warp's 8x4 block (bbox, then the same
    // conservative outside test the pixels use, over the block's sample positions); the warp then visits only the
    // surviving records, in order
    uint32_t rel[2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int idx = hh * 32 + lane;
      bool r = false;
      if (idx < nb) {
        const SglVisPrim &vp = sPrims[idx];
        r = vp.p.bx0 <= wbx + 7 && vp.p.bx1 >= wbx && vp.p.by0 <= wby + 3 && vp.p.by1 >= wby;
        if (r) {
          const uint32_t kind = vp.p.flags & SGL_PF_KIND_MASK;
          if (kind == SGL_PK_TRIANGLE) r = !sglTriSurelyOutside(vp.e, (float) wbx + 4.f, (float) wby + 2.f, 3.875f, 1.875f);
          else if (kind == SGL_PK_LINE) r = sglLineNearRect(vp.p, wbx, wby, wbx + 7, wby + 3);
        }
      }
      rel[hh] = __ballot_sync(0xffffffffu, r);
    }
    if (inFb) {
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        uint32_t m = rel[hh];
        while (m) {
          const int k = hh * 32 + __ffs(m) - 1;
          m &= m - 1;
          sglVisPixelPrim<NS>(P, sPrims[k], slots[k], px, py, depth, owner, hasColor, hasDepth);
        }
      }
    }
  };

  const uint32_t pre = P.tileSortedCount ? P.tileSortedCount[tile] : SGL_TILE_UNSORTED;
  if (pre != SGL_TILE_UNSORTED) {
    // list prepared by sglTileSortKernel: stream it
    const uint32_t *list = P.tileSorted + off + (size_t) tile * SGL_BIG_PER_TILE;
    for (uint32_t b0 = 0; b0 < pre; b0 += SGL_VIS_BATCH) {
      const int nb = pre - b0 < SGL_VIS_BATCH ? (int) (pre - b0) : SGL_VIS_BATCH;
      __syncthreads();
      if (tid < nb) sSlots[tid] = __ldg(list + b0 + tid);
      __syncthreads();
      runBatch(sSlots, nb);
    }
  } else {
  uint32_t nList = P.tileOffset[tile + 1] - off;
  if (off + nList > P.binCapacity) nList = off < P.binCapacity ? P.binCapacity - off : 0;
  uint32_t nBig = *P.bigCount;
  if (nBig > P.bigCapacity) nBig = P.bigCapacity;
  const int tx0 = tx * SGL_TILE, ty0 = ty * SGL_TILE, tx1 = tx0 + SGL_TILE - 1, ty1 = ty0 + SGL_TILE - 1;
  // key windows: the common case (everything fits) is one window covering all keys
  uint32_t lo = 0;
  const uint32_t keyEnd = 0xFFFFFFFFu;
  const bool fits = (nList + nBig) <= SGL_SORT_CAP;
  while (true) {
    uint32_t hi = keyEnd;
    while (true) {   // gather candidates with lo <= key < hi
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
    sglSortTileList(sKeys, sSlots, n);
    for (int b0 = 0; b0 < n; b0 += SGL_VIS_BATCH) {   // process in order
      const int nb = n - b0 < SGL_VIS_BATCH ? n - b0 : SGL_VIS_BATCH;
      __syncthreads();
      runBatch(sSlots + b0, nb);
    }
    __syncthreads();
    if (fits || hi == keyEnd) break;
    lo = hi;
  }
  }

  if (inFb) {
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
  if (P.tileTimes) {
    __syncthreads();
    if (tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      P.tileTimes[2 * tile + 1] = t;
    }
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
        uint4 cq = reinterpret_cast<const uint4 *>(P.colorBase)[pix];
        color[0] = cq.x; color[NS > 1 ? 1 : 0] = cq.y; color[NS > 2 ? 2 : 0] = cq.z; color[NS > 3 ? 3 : 0] = cq.w;
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
      reinterpret_cast<uint4 *>(P.colorBase)[pix] = make_uint4(color[0], color[NS > 1 ? 1 : 0], color[NS > 2 ? 2 : 0], color[NS > 3 ? 3 : 0]);
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
  if (lane == 0 && shaded) atomicAdd(P.counters + 4, (unsigned long long) shaded);
}
