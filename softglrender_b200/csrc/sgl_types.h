// Device-visible PODs shared by the kernels and the C-ABI implementation.
#pragma once
#include <stdint.h>
#include "../../include/sglcuda.h"

#define SGL_TILE 16                 // screen tile = SGL_TILE x SGL_TILE pixels = one CTA of 256 threads
#define SGL_TILE_THREADS (SGL_TILE * SGL_TILE)
#define SGL_MAX_LEVELS 16
#define SGL_OWNER_NONE 0xFFFFFFFFu
#define SGL_BIG_PRIM_TILES 64       // primitives touching more tiles than this go to the pass-wide "big" list
#define SGL_TILE_UNSORTED 0xFFFFFFFFu
#define SGL_TILE_CLASSES 4          // list length >= 192, >= 48, >= 12, rest
#define SGL_RASTER_BLOCK 32         // RendererSoft::rasterBlockSize_ (RendererSoft.h:128)

// texture / attachment descriptor (TextureSoft<T> + ImageBufferSoft<T>, TextureSoft.h:20-255)
struct SglTexObj {
  uint8_t *base;                           // layers x levels, each level in `layout`
  uint8_t *resolve;                        // resolved RGBA8 (the reference's `buffer` next to `bufferMs4x`) or null
  unsigned long long levelOffset[SGL_MAX_LEVELS];  // bytes from the start of a layer
  unsigned long long layerStride;          // bytes
  int32_t width, height, levels, layers, format, samples, layout, pad;
#ifdef SGL_TOUCH_BITMAP
  uint32_t *touch;                         // instrumentation build: 1 bit per 32-byte sector of `base` that a sampler read
#endif
};

// primitive kinds after assembly / polygon-mode expansion
enum { SGL_PK_TRIANGLE = 0, SGL_PK_LINE = 1, SGL_PK_POINT = 2 };

// flags of SglPrim
#define SGL_PF_KIND_MASK 0x3u
#define SGL_PF_FRONT (1u << 2)
#define SGL_PF_IRREGULAR (1u << 3)   // block starts not contiguous (float rounding in RendererSoft.cpp:756-757): exact slow path
#define SGL_PF_DEPTH_TEST (1u << 4)
#define SGL_PF_DEPTH_MASK (1u << 5)
#define SGL_PF_BLEND (1u << 6)
#define SGL_PF_DEPTH_FUNC_SHIFT 7    // 3 bits
#define SGL_PF_VALID (1u << 10)
#define SGL_PF_STEEP (1u << 11)      // line: x/y swapped (RendererSoft.cpp:675-679)
#define SGL_PF_SWAPPED (1u << 12)    // line: endpoints swapped (RendererSoft.cpp:683-689)
#define SGL_PF_BIG (1u << 13)        // lives in the pass-wide big list (touches > SGL_BIG_PRIM_TILES tiles, or the bins were full)

// 64-byte primitive record, written by the setup kernel, read by the tile rasteriser
struct __attribute__((aligned(16))) SglPrim {
  // triangle: screen-space fragPos of the 3 vertices (x, y, z, 1/w)      (VertexHolder::fragPos, RendererInternal.h:39)
  // line:     v[0] = (x0, y0, x1, y1) as int bits after steep/order swaps, v[1] = (z0, z1, w0, w1), v[2].x = lineWidth
  // point:    v[0] = fragPos, v[1].x = pointSize
  float v[3][4];
  int16_t bx0, by0, bx1, by1;   // visited pixel range, inclusive (bbox + quad anchoring, RendererSoft.cpp:724-763)
  uint32_t flags;
  uint32_t draw;                // index into the pass' draw table
};

// vertex indices of a primitive (for varying interpolation at shading time)
struct SglPrimVerts {
  uint32_t i0, i1, i2, pad;
};

struct SglSamplerSlot {
  int32_t tex;       // index into the device texture table, -1 = unbound
  int32_t filter;
  int32_t wrap;
  uint32_t border;   // RGBA8 packed, or float bits for FLOAT32 textures
};

// per-draw record in device memory
struct __attribute__((aligned(16))) SglDrawRec {
  uint8_t uniforms[SGL_MAX_UNIFORM_BYTES];
  SglSamplerSlot samplers[SGL_MAX_SAMPLER_SLOTS];
  SglRenderStates rs;
  int32_t shader;
  uint32_t defines;
  float vpX, vpY, vpW, vpH;               // Viewport (RendererInternal.h:14-28)
  // geometry in
  const float *vertexIn;                  // 16 floats per vertex
  const int32_t *indices;
  int32_t vertexCount, indexCount;
  // vertex stage out (capacity = vertexCap, first vertexCount are the VAO vertices, rest clip-generated)
  float *clipPos;                         // float4
  float *fragPos;                         // float4
  int32_t *clipMask;
  float *vertexOut;                       // 16 floats per vertex: attributes of clip-generated vertices (for re-clipping)
  float *varyings;                        // varyingStride floats per vertex
  int32_t varyingStride, varyingCount;
  int32_t vertexCap;
  int32_t *vertexCounter;                 // next free vertex slot (starts at vertexCount)
  // primitives
  int32_t inputPrims;                     // points / lines / triangles assembled from the index buffer
  int32_t slotsPerPrim;                   // 1, or 3 for polygon-mode LINE/POINT
  int32_t primBase;                       // first slot in the pass' primitive arrays
  int32_t appendBase, appendCap;          // fan triangles produced by clipping: slots [appendBase, appendBase+appendCap)
  int32_t *appendCounter;
  int32_t keyBase;                        // pass-global order key of slot 0
  int32_t hasColor;
  float pointSize;
  uint32_t fastSamplers;                  // bit s: sampler slot s is "simple" (sgl_texture.h split-phase taps); prefilter
                                          // slot: LINEAR or LINEAR_MIPMAP_LINEAR
  uint8_t *vertexUsed;                    // lazy varyings (tile-sharded passes): one byte per VAO vertex, set by the setup kernel for
                                          // the vertices of emitted primitives; sglVaryingKernel then runs the full vertex shader for
                                          // those only.  null: the vertex kernel writes every vertex's varyings itself
};

// One unit of work of the visibility kernel, written in heavy-first order by sglTileSortKernel (geometry stage)
struct __attribute__((aligned(16))) SglVisWork {
  uint32_t tile;        // 0xFFFFFFFF = nothing to do (quarters 1-3 of a heavy tile whose stream could not be prepared)
  uint32_t streamOff;   // first entry of the tile's packed record stream (SglPassParams::stream)
  uint32_t count;       // entries of the stream; SGL_TILE_UNSORTED = no stream: in-kernel gather / sort path
  uint32_t quarter;     // 0..3: one 8x8 pixel quarter, one SAMPLE per lane (heavy MSAA tiles); 0xFFFFFFFF: whole tile
};
struct SglVisPrim;      // sgl_vis.cuh: primitive record + edge constants + slot, 128 bytes

// pass-level parameters
struct SglPassParams {
  // attachments
  uint8_t *colorBase;       // RGBA8 [y][x][sample] of the attached layer/level, or null
  float *depthBase;         // float [y][x][sample], or null
  uint8_t *resolveBase;     // resolved colour (MS only), or null
  uint8_t *colorMask;       // MS only, or null: one byte per pixel, 0 = the four samples are equal and only the resolved colour
                            // (== each sample) is stored, 1 = the 16-byte per-sample record in colorBase is valid (sglStoreMsColor)
  uint8_t *mirrorBase;      // optional second destination of the final colour (resolved colour for MS targets), linear
                            // RGBA8 [y][x]; may be peer memory of another GPU (multi-GPU gather by direct store), or null
  uint32_t *vis;            // visibility buffer of the deferred path: owner (slot | shading sample << 29) [y][x][sample]
  int32_t fbW, fbH, samples;
  int32_t clearColorFlag, clearDepthFlag;
  int32_t skipEmptyTiles;   // fused kernel, tail of a split pass (both clear flags off): tiles without primitives keep their content
  uint32_t clearColor;      // RGBA8 packed (RendererSoft.cpp:72-75)
  float clearDepth;
  // tiles
  int32_t tilesX, tilesY;
  const uint8_t *tileOwner; // null = own all
  int32_t rank;
  // work
  const SglDrawRec *draws;
  int32_t drawCount;
  const SglPrim *prims;
  const SglPrimVerts *primVerts;
  const uint32_t *primKeys;     // order key per primitive slot
  int32_t primSlots;            // total slots in this pass
  // bins
  uint32_t *tileCount;          // [tiles]
  uint32_t *tileOffset;         // [tiles+1]
  uint32_t *tileCursor;         // [tiles]
  uint32_t *binSlots;           // primitive slots per tile (unordered)
  uint32_t binCapacity;
  uint32_t *tileSorted;         // per tile: its bin in submission order, written by sglTileSortKernel at tileOffset[t];
                                // null = not prepared
  uint32_t *tileSortedCount;    // [tiles] entries of the sorted list, SGL_TILE_UNSORTED = use the in-kernel gather
  uint32_t *tileOrder;          // [SGL_TILE_CLASSES][tiles]: tiles by descending list length class (heavy tiles are
  uint32_t *tileClassCount;     // [SGL_TILE_CLASSES]           launched first so that they cannot become stragglers)
  int32_t splitCap;             // visibility kernel: up to splitCap heavy tiles run as four quarter-tile CTAs
  // big primitives (> SGL_BIG_PRIM_TILES tiles in their pixel range, or the bins were exhausted when they arrived): the
  // setup kernel lists them in bigAll; sglBigBinKernel bins them tile by tile like everybody else.  Only what does not fit
  // the bins then stays in the RESIDUAL list bigList/bigCount, which every tile kernel scans (normally empty).
  uint32_t *bigAll;
  uint32_t *bigAllCount;
  uint32_t *bigList;
  uint32_t *bigCount;
  uint32_t bigCapacity;         // of both lists (>= primitive slots: they never overflow)
  uint32_t *binReserved;        // running count of bin entries handed out
  // visibility-kernel work list (one CTA per item) and the packed per-tile record streams it copies with cp.async.bulk
  SglVisWork *work;
  SglVisPrim *stream;
  uint32_t *streamCursor;
  uint32_t streamCapacity;      // entries
  const SglTexObj *textures;
  unsigned long long *counters; // device-side SglCounters mirror
  unsigned long long *fragCounters;  // [32] fragments shaded, spread over 32 words (one hot word would serialise 65 k warps at one L2 slice)
  unsigned long long *tileTimes; // instrumentation (normally null): [tiles][2] globaltimer ns at CTA start / end of the visibility kernel
};
