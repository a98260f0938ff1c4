// libsglcuda.so -- C ABI (include/sglcuda.h) over the sm_100a kernels.
// Host-side responsibilities: resource tables (buffers, textures), recording the draws of a render pass with
// snapshots of uniforms/sampler bindings/states, laying out the pass' transient arena in HBM, and launching
// the five kernels of sgl_kernels.cuh at sgl_pass_end (tile-based deferred execution, the model the reference's
// own Vulkan backend uses behind the same API -- Render/Vulkan/RendererVulkan.cpp:73-205).
#define SGL_WITH_TILE_SORT
#include "sgl_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// tile rasteriser instantiations live in their own translation units (sgl_raster_ns{1,4}.cu)
extern "C" int sglLaunchRaster1(const SglPassParams *P, int nTiles, void *stream);
extern "C" int sglLaunchRaster4(const SglPassParams *P, int nTiles, void *stream);
// deferred path: visibility kernels (sgl_vis.cu) and shading kernels (sgl_shade_ns{1,4}.cu)
extern "C" int sglLaunchVis(int samples, const SglPassParams *P, int nTiles, void *stream);
extern "C" int sglLaunchShade1(const SglPassParams *P, int nTiles, void *stream);
extern "C" int sglLaunchShade4(const SglPassParams *P, int nTiles, void *stream);
// depth-only passes (sgl_depth.cu)
#include "sgl_depth_pass.h"
extern "C" int sglLaunchDepthSetup(int samples, const SglDepthPass *D, int maxPrims, int nDraws, void *stream);
extern "C" int sglLaunchDepthRaster(int samples, const SglDepthPass *D, int nTiles, void *stream);

namespace {

struct BufferRec {
  void *d = nullptr;
  size_t bytes = 0;
};

struct TextureRec {
  bool alive = false;
  SglTextureDesc desc{};
  SglTexObj obj{};
  size_t bytes = 0;
  void *mirror = nullptr;   // sgl_texture_set_mirror: second destination of the final colour (may be peer memory)
  int shardHalo = 0;        // sgl_texture_set_shard_halo: pixels around owned tiles that passes into this texture also render
                            // under a tile-owner map (< 0: every tile -- the texture is replicated)
  cudaEvent_t rbDone = nullptr;   // completion of the last sgl_texture_readback_async on the copy stream
  bool rbPending = false;         // a later pass that overwrites the colour image must wait for rbDone on the device
  uint8_t *cmask = nullptr;       // multisample colour textures: per-pixel "per-sample record valid" mask (sgl_pixel.h)
  bool cmaskOff = false;          // the per-sample image was handed out (sgl_texture_device_ptr): every record is written from now on
  void *altBase = nullptr;        // second backing store of a depth texture whose depth-only passes are renamed (see runPass)
  bool exposed = false;           // sgl_texture_device_ptr handed the storage out: work the library cannot see may touch it in stream order
};

#define SGL_ARENAS 8
#define SGL_SMALL_ARENA_BYTES ((size_t) 32 << 20)
struct Ctx {
  bool ready = false;
  int refs = 0;              // sgl_init calls not yet matched by sgl_shutdown (several Renderer objects may share the context)
  int device = 0, rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  cudaStream_t copyStream = nullptr;   // asynchronous read-backs overlap the next frame's geometry / visibility work
  void *rbStage[2] = {nullptr, nullptr};   // device-side snapshots of an image on its way to the host (see sgl_texture_readback_async)
  size_t rbStageCap[2] = {0, 0};
  int rbStageNext = 0;
  cudaEvent_t copyReady = nullptr;
  std::vector<BufferRec> buffers{1};
  std::vector<TextureRec> textures{1};
  SglTexObj *dTextures = nullptr;
  int dTexCap = 0;
  // pass
  bool inPass = false;
  int colorTex = 0, colorLayer = 0, colorLevel = 0, depthTex = 0;
  int clearColorFlag = 0, clearDepthFlag = 0;
  float clearColor[4] = {0, 0, 0, 0};
  float clearDepth = 1.f;
  float vpX = 0, vpY = 0, vpW = 0, vpH = 0;
  std::vector<SglDrawRec> draws;
  // pass arenas: a ring, so that the geometry stages of a pass (vertex, setup, scan, bin fill -- they touch only the
  // arena) run on their own stream while the pixel stages (visibility, shading) of the previous passes are still busy
  struct Arena {
    uint8_t *mem = nullptr;
    size_t cap = 0;
    cudaEvent_t geomDone = nullptr, pixelDone = nullptr;
    bool used = false;
    void *stagingHost = nullptr;   // pinned: this slot's draw records (fixed address: the geometry graph's copy node reads it)
    size_t stagingCap = 0;
  };
  // small passes (arena <= SGL_SMALL_ARENA_BYTES) rotate over all SGL_ARENAS slots, large ones over the first three: many
  // small dependent kernels per pass (config 5: 512x512 views) need more passes in flight to hide the geometry chain's latency
  Arena arenas[SGL_ARENAS];
  int arenaNext = 0;
  int ringSize = 4;            // arena slots that ordinary passes rotate over (SGL_RING=3..8).  Four: in a frame of two passes (shadow + main) the main
                               // pass always reuses the slot of the main pass two frames back; with three its geometry waited for the PREVIOUS frame's
                               // shadow raster (which waits for the frame before to finish shading) two frames out of three -- e2e +4.8 % on config 2
  int smallStreak = 0;         // consecutive passes whose arena is small (<= SGL_SMALL_ARENA_BYTES, <= 4096 tiles)
  unsigned long long *dTileTimes = nullptr;        // sgl_debug_tile_times
  size_t tileTimesCap = 0;
  int tileTiming = 0;
  const uint32_t *lastTileSortedCount = nullptr;   // of the most recent non-depth-only pass (instrumentation)
  int lastTilesX = 0, lastTilesY = 0;
  cudaStream_t geomStreams[SGL_ARENAS] = {};   // one per arena slot: geometry stages of consecutive passes are independent
  // depth-only passes (shadow maps) run their pixel stage on an auxiliary stream: they start when all earlier pixel work
  // is done and only the next kernel that SAMPLES textures (or touches their depth texture) waits for them, so the
  // shadow pass of a frame overlaps the visibility kernel of its main pass
  cudaStream_t peerStream = nullptr;   // rank 0's gather bookkeeping (sgl_peer_collect), decoupled from its own rendering
  cudaStream_t auxStream = nullptr;
  cudaEvent_t auxReady = nullptr, auxDone = nullptr;
  bool auxPending = false;
  std::vector<int> auxDepthTex;   // depth textures written by the auxiliary-stream passes not joined yet
  int noGraphs = 0;            // SGL_NO_GRAPHS=1: every kernel of a stage is launched individually (A/B runs, tests)
  int noOverlap = 0;           // SGL_NO_OVERLAP=1: geometry and pixel stages on one stream (A/B runs)
  int noPassSplit = 0;         // SGL_NO_PASS_SPLIT=1: a pass with a blended tail runs entirely in the fused kernel (A/B runs)
  int noLazyVaryings = 0;      // SGL_NO_LAZY_VARYINGS=1: tile-sharded passes shade every vertex up front (A/B runs)
  int rbMode = 0;              // SGL_RB_MODE=1: read-back snapshot by a copy kernel on the rendering stream (A/B runs: 3 % slower on config 2)
  cudaEvent_t rbStageFree[2] = {nullptr, nullptr};   // PCIe copy out of rbStage[k] finished
  int ceUpload = 0;            // SGL_CE_UPLOAD=1: counter memset + draw-record upload as copy-engine nodes (A/B runs)
  int fewArenas = 0;           // SGL_FEW_ARENAS=1: three arena slots for every pass (A/B runs)
  int noSplit1 = 0;            // SGL_NO_SPLIT1=1: single-sample heavy tiles are not split (A/B runs)
  int noSplit = 0;             // SGL_NO_SPLIT=1: heavy MSAA tiles are not split into quarter-tile CTAs (A/B runs)
  void *dummyTexels = nullptr; // backing store of texture table entry 0
  uint32_t *vis[2] = {nullptr, nullptr};   // visibility buffers of the deferred path (two: see "early visibility" in runPass)
  size_t visCap[2] = {0, 0};
  int visNext = 0;
  // Early visibility: the visibility kernel of pass n+1 may start while the shading kernel of pass n still runs (it reads the
  // other visibility buffer and touches neither the colour attachment nor anything pass n samples).  It is launched on visStream
  // behind preShade -- the position of the rendering stream just BEFORE pass n's shading kernel -- which is only legal while
  // nothing else has been queued on the rendering stream since that kernel: mainSeq counts every such submission.
  cudaStream_t visStream = nullptr;
  cudaEvent_t preShade = nullptr, visDone = nullptr;
  unsigned long long mainSeq = 1, seqAfterShade = 0;
  std::vector<int> lastShadeSampled;   // textures bound to the draws of the pass whose shading kernel was launched last
  // Renaming: a depth-only pass that clears a single-level depth texture (the shadow map of every frame) renders into the
  // texture's OTHER backing store, so it does not have to wait for the shading kernel that still samples the previous
  // contents; the device texture table learns the new address in stream order before the next kernel that samples.
  std::vector<int> pendingTexBase;   // handles whose table entry still shows the previous backing store
  int noMsMask = 0;            // SGL_NO_MS_MASK=1: multisample colour always stored per sample (A/B runs, tests)
  int noRename = 0;            // SGL_NO_RENAME=1 (A/B runs, tests)
  unsigned long long hostRenames = 0;
  int noEarlyVis = 0;          // SGL_NO_EARLY_VIS=1: every visibility kernel on the rendering stream (A/B runs, tests)
  int forceFused = 0;          // SGL_FORCE_FUSED=1: always use the fused tile kernel (A/B runs, tests)
  // multi-GPU
  uint8_t *dTileOwner = nullptr;
  int ownerTilesX = 0, ownerTilesY = 0;
  std::vector<uint8_t> hostTileOwner;
  // dilated copies of the owner map for passes into textures with a shard halo: entry == rank for owned tiles and tiles
  // within the halo, 255 elsewhere (cached per halo width in tiles; dropped when the map or the rank changes)
  struct HaloMap { int tiles; uint8_t *d; };
  std::vector<HaloMap> haloMaps;
  uint32_t *dOwnerPrefix = nullptr;   // exclusive count of tiles owned by prefixRank (sgl_tiles_pack / unpack)
  int prefixRank = -1;
  std::vector<void *> peerAllocs, peerMaps;
  // counters
  unsigned long long *dCounters = nullptr;
  unsigned int *hOverflow = nullptr;   // pinned, device-visible: set by a kernel that had to DROP geometry (clip arena full)
  unsigned int *dOverflow = nullptr;   // device alias of hOverflow
  int clipScale = 1;                   // grows after an overflow: clip-vertex / fan arenas of later passes are this much larger
  long long binCapLimit = 0, clipMinVerts = 65536, clipMinFans = 32768;   // sgl_debug_set_limits (tests shrink them)
  unsigned long long hostEarlyVis = 0;   // visibility kernels that started ahead of the previous pass's shading kernel
  unsigned long long hostLaunches = 0, hostPasses = 0, hostDraws = 0, hostH2D = 0, hostD2H = 0, hostNsPassEnd = 0, hostNsDraw = 0, hostVertices = 0, hostIndices = 0, hostNsWaitGpu = 0;
  cudaEvent_t evBegin = nullptr, evEnd = nullptr;
  std::string err;
};

Ctx g;
// handles are never reused across a shutdown / init cycle: objects that outlive their renderer (the reference's scene
// caches do, Viewer.cpp:59-66) then release a dead handle instead of somebody else's resource
size_t gRetiredBuffers = 1, gRetiredTextures = 1;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g.err = buf;
  return code;
}

void joinAux();

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(SGL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// Entry points that only record host state (pass_begin / set_viewport / draw / pass_end) use NEED_CTX_RECORDING; every
// other entry point first orders the main stream behind a depth-only pass still running on the auxiliary stream.
#define NEED_CTX_RECORDING() \
  if (!g.ready) return fail(SGL_ERR_STATE, "sgl_init has not been called (or failed)")
#define NEED_CTX()        \
  NEED_CTX_RECORDING();   \
  g.mainSeq++;            \
  if (!g.pendingTexBase.empty()) { int rcTb = flushTexBases(); if (rcTb) return rcTb; } \
  joinAux()

// queues (on the rendering stream, i.e. behind every kernel that still samples the old contents) the table updates of
// renamed depth textures; must precede any kernel that samples textures
int flushTexBases();
void joinAux() {
  if (!g.auxPending) return;
  g.mainSeq++;     // later rendering-stream work is ordered behind the auxiliary passes from here on; an early visibility kernel would not be
  cudaStreamWaitEvent(g.stream, g.auxDone, 0);
  g.auxPending = false;
  g.auxDepthTex.clear();
}

// API calls whose rendering-stream work (if any) touches nothing a visibility kernel reads or writes -- gather flags,
// mirror pointers -- do not end the "nothing since the last shading kernel" state that early visibility and renaming need:
// they restore mainSeq on success (unless they had to join the auxiliary stream).
struct SeqKeeper {
  unsigned long long seq = g.mainSeq;
  bool auxWasPending = g.auxPending;
  int done(int rc) const {
    if (rc == SGL_OK && !auxWasPending) g.mainSeq = seq;
    return rc;
  }
};

size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

int levelCount(const SglTextureDesc &d) {
  if (!d.use_mipmaps) return 1;
  int m = std::max(d.width, d.height), n = 0;
  while ((1 << (n + 1)) <= m) n++;          // floor(log2(max(w,h))) + 1 levels (SamplerSoft.h:98)
  return n + 1;
}

int uploadTexObj(int handle) {
  if (handle >= g.dTexCap) {
    int ncap = std::max(64, g.dTexCap * 2);
    while (ncap <= handle) ncap *= 2;
    SglTexObj *n = nullptr;
    CU(cudaMalloc(&n, sizeof(SglTexObj) * ncap));
    CU(cudaMemset(n, 0, sizeof(SglTexObj) * ncap));
    if (g.dTextures) {
      CU(cudaStreamSynchronize(g.stream));
      CU(cudaMemcpy(n, g.dTextures, sizeof(SglTexObj) * g.dTexCap, cudaMemcpyDeviceToDevice));
      CU(cudaFree(g.dTextures));
    }
    g.dTextures = n;
    g.dTexCap = ncap;
  }
  CU(cudaMemcpyAsync(g.dTextures + handle, &g.textures[handle].obj, sizeof(SglTexObj), cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));   // obj lives in a std::vector that may move
  return SGL_OK;
}

TextureRec *tex(int h) {
  if (h <= 0 || h >= (int) g.textures.size() || !g.textures[h].alive) return nullptr;
  return &g.textures[h];
}

uint8_t *levelPtr(const TextureRec &t, int layer, int level) {
  return t.obj.base + (size_t) layer * t.obj.layerStride + t.obj.levelOffset[level];
}

int syncAll() {
  CU(cudaStreamSynchronize(g.stream));
  for (cudaStream_t gs : g.geomStreams) if (gs) CU(cudaStreamSynchronize(gs));
  if (g.auxStream) CU(cudaStreamSynchronize(g.auxStream));
  if (g.visStream) CU(cudaStreamSynchronize(g.visStream));
  if (g.peerStream) CU(cudaStreamSynchronize(g.peerStream));
  g.auxPending = false;
  g.auxDepthTex.clear();
  return SGL_OK;
}

// Called after the streams have been drained: a pass that ran out of clip-arena space dropped primitives.  That is
// reported once as SGL_ERR_OVERFLOW (the reference never drops geometry, so the frame must not pass as good) and the
// arenas of later passes are enlarged, so a caller that simply retries the frame succeeds.
int checkOverflow() {
  if (!g.hOverflow || !*(volatile unsigned int *) g.hOverflow) return SGL_OK;
  *(volatile unsigned int *) g.hOverflow = 0;
  if (g.clipScale < 64) g.clipScale *= 2;
  return fail(SGL_ERR_OVERFLOW, "a render pass ran out of clip-vertex arena space and dropped primitives "
                                "(sgl_get_counters().clip_overflow); later passes get a %dx larger arena -- resubmit the frame", g.clipScale);
}

void dropHaloMaps() {
  for (auto &h : g.haloMaps) if (h.d) cudaFree(h.d);
  g.haloMaps.clear();
}

// owner map for a pass into a texture with `haloPixels` of shard halo (0: the plain map; < 0: null = render every tile)
int ownerMapForHalo(int haloPixels, const uint8_t **out) {
  *out = nullptr;
  if (haloPixels < 0) return SGL_OK;
  if (haloPixels == 0) { *out = g.dTileOwner; return SGL_OK; }
  const int ht = (haloPixels + SGL_TILE - 1) / SGL_TILE;
  for (auto &h : g.haloMaps) if (h.tiles == ht) { *out = h.d; return SGL_OK; }
  const int tx = g.ownerTilesX, ty = g.ownerTilesY;
  std::vector<uint8_t> m((size_t) tx * ty, 255);
  for (int y = 0; y < ty; y++)
    for (int x = 0; x < tx; x++) {
      if (g.hostTileOwner[(size_t) y * tx + x] != g.rank) continue;
      for (int yy = std::max(0, y - ht); yy <= std::min(ty - 1, y + ht); yy++)
        for (int xx = std::max(0, x - ht); xx <= std::min(tx - 1, x + ht); xx++) m[(size_t) yy * tx + xx] = (uint8_t) g.rank;
    }
  Ctx::HaloMap h;
  h.tiles = ht;
  h.d = nullptr;
  CU(cudaMalloc(&h.d, m.size()));
  CU(cudaMemcpy(h.d, m.data(), m.size(), cudaMemcpyHostToDevice));
  g.haloMaps.push_back(h);
  *out = h.d;
  return SGL_OK;
}

int ensureArena(Ctx::Arena &a, size_t bytes) {
  if (!a.geomDone) {
    CU(cudaEventCreateWithFlags(&a.geomDone, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&a.pixelDone, cudaEventDisableTiming));
  }
  if (bytes <= a.cap) return SGL_OK;
  int rc = syncAll();
  if (rc) return rc;
  if (a.mem) CU(cudaFree(a.mem));
  a.mem = nullptr;
  a.cap = 0;
  size_t ncap = alignUp(bytes + bytes / 4, 1 << 20);
  cudaError_t e = cudaMalloc(&a.mem, ncap);
  if (e != cudaSuccess) return fail(SGL_ERR_OOM, "pass arena of %zu bytes: %s", ncap, cudaGetErrorString(e));
  a.cap = ncap;
  return SGL_OK;
}

struct ShaderMeta {
  const char *blocks[4];
  int blockOffsets[4];
  const char *samplers[8];
  const char *defines[8];
};

const ShaderMeta *shaderMeta(int shader) {
  // ShaderSoft::getUniformsDesc / getDefines of each program (e.g. PbrSoft.h:77-103, BlinnPhongSoft.h:71-96)
  static const ShaderMeta basic = {{"UniformsModel", "UniformsMaterial"}, {0, 256}, {nullptr}, {nullptr}};
  static const ShaderMeta blinn = {{"UniformsModel", "UniformsScene", "UniformsMaterial"}, {0, 256, 320},
                                   {"u_albedoMap", "u_normalMap", "u_emissiveMap", "u_aoMap", "u_shadowMap"},
                                   {"ALBEDO_MAP", "NORMAL_MAP", "EMISSIVE_MAP", "AO_MAP"}};
  static const ShaderMeta pbr = {{"UniformsModel", "UniformsScene", "UniformsMaterial"}, {0, 256, 320},
                                 {"u_albedoMap", "u_normalMap", "u_emissiveMap", "u_aoMap", "u_metalRoughnessMap",
                                  "u_irradianceMap", "u_prefilterMap"},
                                 {"ALBEDO_MAP", "NORMAL_MAP", "EMISSIVE_MAP", "AO_MAP", "METALROUGHNESS_MAP"}};
  static const ShaderMeta sky = {{"UniformsModel"}, {0}, {"u_equirectangularMap", "u_cubeMap"}, {"EQUIRECTANGULAR_MAP"}};
  static const ShaderMeta irr = {{"UniformsModel"}, {0}, {"u_cubeMap"}, {nullptr}};
  static const ShaderMeta pre = {{"UniformsModel", "UniformsPrefilter"}, {0, 256}, {"u_cubeMap"}, {nullptr}};
  static const ShaderMeta fxaa = {{"UniformsQuadFilter"}, {0}, {"u_screenTexture"}, {nullptr}};
  switch (shader) {
    case SGL_SHADER_BASIC: return &basic;
    case SGL_SHADER_BLINNPHONG: return &blinn;
    case SGL_SHADER_PBR: return &pbr;
    case SGL_SHADER_SKYBOX: return &sky;
    case SGL_SHADER_IBL_IRRADIANCE: return &irr;
    case SGL_SHADER_IBL_PREFILTER: return &pre;
    case SGL_SHADER_FXAA: return &fxaa;
  }
  return nullptr;
}

struct ProfEvent {
  const char *name;
  cudaEvent_t a, b;
};
std::vector<ProfEvent> gProf;
std::vector<cudaEvent_t> gProfPool;
bool gProfiling = false;

cudaEvent_t profEvent() {
  cudaEvent_t e = nullptr;
  if (!gProfPool.empty()) {
    e = gProfPool.back();
    gProfPool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
cudaStream_t gCur = nullptr;   // stream of the stage being issued (null = the context's main stream)
cudaStream_t curStream() { return gCur ? gCur : g.stream; }
void profBegin(const char *name) {
  if (!gProfiling) return;
  ProfEvent p = {name, profEvent(), profEvent()};
  cudaEventRecord(p.a, curStream());
  gProf.push_back(p);
}
void profEnd() {
  if (!gProfiling) return;
  cudaEventRecord(gProf.back().b, curStream());
}

template<typename... Args>
int launch(const char *name, void (*kernel)(Args...), dim3 grid, dim3 block, Args... args) {
  profBegin(name);
  kernel<<<grid, block, 0, curStream()>>>(args...);
  profEnd();
  g.hostLaunches++;
  if (curStream() == g.stream) g.mainSeq++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SGL_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
  return SGL_OK;
}

// the per-sample image of a multisample colour texture is about to be read by something other than the pixel kernels
int expandMsColor(TextureRec *t) {
  if (!t->cmask || t->obj.samples == 1) return SGL_OK;
  struct CurGuard { cudaStream_t saved; ~CurGuard() { gCur = saved; } } guard{gCur};
  gCur = g.stream;
  const size_t n = (size_t) t->obj.width * t->obj.height;
  return launch("sglMsExpandKernel", sglMsExpandKernel, dim3((unsigned) std::min<size_t>((n + 255) / 256, 148 * 8)), dim3(256),
                (uint4 *) t->obj.base, (const uint32_t *) t->obj.resolve, t->cmask, n);
}

int flushTexBases() {
  struct CurGuard { cudaStream_t saved; ~CurGuard() { gCur = saved; } } guard{gCur};
  gCur = g.stream;
  for (int h : g.pendingTexBase) {
    int rc = launch("sglSetTexBaseKernel", sglSetTexBaseKernel, dim3(1), dim3(1), g.dTextures + h, g.textures[h].obj.base);
    if (rc) return rc;
  }
  g.pendingTexBase.clear();
  return SGL_OK;
}

// ---- stage graphs (submission cost) ---------------------------------------------------------------
// The kernels of one stage of a pass (memset + record upload + vertex/setup/binning/sort kernels; or the fill + raster
// kernels of a depth-only pass) are one stream-ordered chain whose launch parameters only depend on the pass "signature"
// (arena slot, attachment pointers, sizes, draw counts).  The first pass with a given signature captures the chain into a
// CUDA graph; every later one replays it with ONE cudaGraphLaunch -- the per-draw data (uniform snapshots, pointers) is not
// part of the graph: it travels in the pinned staging buffer the graph's copy node reads.  Events and cross-stream waits
// stay outside the graphs, so the scheduling is exactly that of the individually launched kernels.
// SGL_HOST_PROFILE=1: where sgl_pass_end's CPU time goes (printed at shutdown)
unsigned long long gHostSec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
const char *gHostSecName[8] = {"layout+records", "staging wait+copy", "geometry stage", "events", "pixel launches", "depth pixel stage", "arena", "-"};
struct SecTimer {
  int k;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit SecTimer(int k_) : k(k_) {}
  void switchTo(int k2) {
    auto t1 = std::chrono::steady_clock::now();
    gHostSec[k] += (unsigned long long) std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    k = k2;
    t0 = t1;
  }
  ~SecTimer() { switchTo(k); }
};

struct StageGraph {
  uint64_t key;
  cudaGraphExec_t exec;
  unsigned launches;
  uint64_t lastUse;
};
std::vector<StageGraph> gStageGraphs;
uint64_t gStageClock = 0;

uint64_t hashBytes(uint64_t h, const void *p, size_t n) {   // FNV-1a
  const unsigned char *b = (const unsigned char *) p;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
template<typename T> uint64_t hashPod(uint64_t h, const T &v) { return hashBytes(h, &v, sizeof(T)); }

void dropStageGraphs() {
  for (auto &e : gStageGraphs) cudaGraphExecDestroy(e.exec);
  gStageGraphs.clear();
}

// issue() queues the chain on stream `s` through launch() / cudaMemcpyAsync(..., curStream()) only
template<class F>
int runStage(cudaStream_t s, uint64_t key, F issue) {
  struct CurGuard { cudaStream_t saved; ~CurGuard() { gCur = saved; } } guard{gCur};
  gCur = s;
  if (s == g.stream) g.mainSeq++;
  if (g.noGraphs || gProfiling) return issue();
  for (auto &e : gStageGraphs)
    if (e.key == key) {
      e.lastUse = ++gStageClock;
      g.hostLaunches += e.launches;
      CU(cudaGraphLaunch(e.exec, s));
      return SGL_OK;
    }
  const unsigned long long before = g.hostLaunches;
  CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = issue();
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != SGL_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(SGL_ERR_CUDA, "stage graph capture failed: %s", cudaGetErrorString(e));
  StageGraph sg;
  sg.key = key;
  sg.exec = nullptr;
  sg.launches = (unsigned) (g.hostLaunches - before);
  sg.lastUse = ++gStageClock;
  e = cudaGraphInstantiate(&sg.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(SGL_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  if (gStageGraphs.size() >= 96) {   // bounded cache: drop the least recently used graph
    size_t lru = 0;
    for (size_t i = 1; i < gStageGraphs.size(); i++) if (gStageGraphs[i].lastUse < gStageGraphs[lru].lastUse) lru = i;
    cudaGraphExecDestroy(gStageGraphs[lru].exec);
    gStageGraphs.erase(gStageGraphs.begin() + lru);
  }
  gStageGraphs.push_back(sg);
  CU(cudaGraphLaunch(sg.exec, s));
  return SGL_OK;
}

}  // namespace

namespace {
template<typename T>
struct DevTmp {
  T *p = nullptr;
  ~DevTmp() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};
}  // namespace

extern "C" {

const char *sgl_last_error(void) { return g.err.c_str(); }

int sgl_init(int device_ordinal, int rank, int world) {
  if (g.ready) {
    g.refs++;
    return SGL_OK;
  }
  g.buffers.resize(gRetiredBuffers);
  g.textures.resize(gRetiredTextures);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SGL_ERR_NO_DEVICE, "no CUDA device (%s); RendererCUDA has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device_ordinal < 0 || device_ordinal >= n) return fail(SGL_ERR_INVALID, "device ordinal %d out of range", device_ordinal);
  CU(cudaSetDevice(device_ordinal));
  g.device = device_ordinal;
  g.rank = rank;
  g.world = world < 1 ? 1 : world;
  CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  g.ownStream = true;
  CU(cudaMalloc(&g.dCounters, 40 * sizeof(unsigned long long)));   // 8 counters + 32 spread fragment counters
  CU(cudaMemset(g.dCounters, 0, 40 * sizeof(unsigned long long)));
  CU(cudaEventCreate(&g.evBegin));
  CU(cudaEventCreate(&g.evEnd));
  CU(cudaHostAlloc((void **) &g.hOverflow, 64, cudaHostAllocMapped));
  memset(g.hOverflow, 0, 64);
  CU(cudaHostGetDevicePointer((void **) &g.dOverflow, g.hOverflow, 0));
  {  // geometry kernels are small and feed the pixel stage of the NEXT pass: give them priority over resident pixel work
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (auto &gs : g.geomStreams) CU(cudaStreamCreateWithPriority(&gs, cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&g.auxStream, cudaStreamNonBlocking, hi));   // small kernels next to a visibility kernel
    CU(cudaStreamCreateWithFlags(&g.visStream, cudaStreamNonBlocking));          // early visibility kernels (same priority as the rendering stream)
  }
  CU(cudaEventCreateWithFlags(&g.preShade, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&g.visDone, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&g.auxReady, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&g.auxDone, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&g.copyStream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&g.peerStream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&g.copyReady, cudaEventDisableTiming));
  {  // texture table entry 0 = 1x1 RGBA8 dummy: what unbound maps read in the straight-line shader paths
    CU(cudaMalloc(&g.dummyTexels, 256));
    CU(cudaMemset(g.dummyTexels, 0, 256));
    SglTexObj &o = g.textures[0].obj;
    memset(&o, 0, sizeof(o));
    o.base = (uint8_t *) g.dummyTexels;
    o.layerStride = 0;
    o.width = o.height = o.levels = o.layers = o.samples = 1;
    o.format = SGL_FMT_RGBA8;
    o.layout = SGL_LAYOUT_LINEAR;
    int rc0 = uploadTexObj(0);
    if (rc0) return rc0;
  }
  {
    const char *ff = getenv("SGL_FORCE_FUSED");
    g.forceFused = (ff && atoi(ff) != 0) ? 1 : 0;
    const char *ng = getenv("SGL_NO_GRAPHS");
    g.noGraphs = (ng && atoi(ng) != 0) ? 1 : 0;
    const char *no = getenv("SGL_NO_OVERLAP");
    g.noOverlap = (no && atoi(no) != 0) ? 1 : 0;
    const char *nps = getenv("SGL_NO_PASS_SPLIT");
    g.noPassSplit = (nps && atoi(nps) != 0) ? 1 : 0;
    const char *ns = getenv("SGL_NO_SPLIT");
    g.noSplit = (ns && atoi(ns) != 0) ? 1 : 0;
    const char *fa = getenv("SGL_FEW_ARENAS");
    g.fewArenas = (fa && atoi(fa) != 0) ? 1 : 0;
    const char *nmm = getenv("SGL_NO_MS_MASK");
    g.noMsMask = (nmm && atoi(nmm) != 0) ? 1 : 0;
    const char *nrn = getenv("SGL_NO_RENAME");
    g.noRename = (nrn && atoi(nrn) != 0) ? 1 : 0;
    const char *nev = getenv("SGL_NO_EARLY_VIS");
    g.noEarlyVis = (nev && atoi(nev) != 0) ? 1 : 0;
    const char *rg = getenv("SGL_RING");
    g.ringSize = rg ? std::min(std::max(atoi(rg), 3), SGL_ARENAS) : 4;
    const char *rbm = getenv("SGL_RB_MODE");
    g.rbMode = rbm ? atoi(rbm) : 0;
    const char *ceu = getenv("SGL_CE_UPLOAD");
    g.ceUpload = (ceu && atoi(ceu) != 0) ? 1 : 0;
    const char *ns1 = getenv("SGL_NO_SPLIT1");
    g.noSplit1 = (ns1 && atoi(ns1) != 0) ? 1 : 0;
    const char *nl = getenv("SGL_NO_LAZY_VARYINGS");
    g.noLazyVaryings = (nl && atoi(nl) != 0) ? 1 : 0;
  }
  g.ready = true;
  g.refs = 1;
  g.err.clear();
  return SGL_OK;
}

int sgl_shutdown(void) {
  if (!g.ready) return SGL_OK;
  if (--g.refs > 0) return SGL_OK;
  gRetiredBuffers = g.buffers.size();
  gRetiredTextures = g.textures.size();
  cudaStreamSynchronize(g.stream);
  for (cudaStream_t gs : g.geomStreams) if (gs) cudaStreamSynchronize(gs);
  for (auto &b : g.buffers)
    if (b.d) cudaFree(b.d);
  if (g.copyStream) cudaStreamSynchronize(g.copyStream);
  for (auto &t : g.textures) {
    if (t.alive && t.obj.base) cudaFree(t.obj.base);
    if (t.alive && t.altBase) cudaFree(t.altBase);
    if (t.alive && t.cmask) cudaFree(t.cmask);
    if (t.alive && t.obj.resolve) cudaFree(t.obj.resolve);
    if (t.rbDone) cudaEventDestroy(t.rbDone);
  }
  for (void *p : g.rbStage) if (p) cudaFree(p);
  for (auto &e : g.rbStageFree) { if (e) cudaEventDestroy(e); e = nullptr; }
  if (g.copyStream) cudaStreamDestroy(g.copyStream);
  if (g.copyReady) cudaEventDestroy(g.copyReady);
  if (g.dTextures) cudaFree(g.dTextures);
  dropStageGraphs();
  for (auto &a : g.arenas) {
    if (a.stagingHost) cudaFreeHost(a.stagingHost);
    if (a.mem) cudaFree(a.mem);
    if (a.geomDone) cudaEventDestroy(a.geomDone);
    if (a.pixelDone) cudaEventDestroy(a.pixelDone);
  }
  for (cudaStream_t gs : g.geomStreams) if (gs) cudaStreamDestroy(gs);
  if (g.auxStream) { cudaStreamSynchronize(g.auxStream); cudaStreamDestroy(g.auxStream); }
  if (g.peerStream) { cudaStreamSynchronize(g.peerStream); cudaStreamDestroy(g.peerStream); }
  if (g.visStream) { cudaStreamSynchronize(g.visStream); cudaStreamDestroy(g.visStream); }
  if (g.preShade) cudaEventDestroy(g.preShade);
  if (g.visDone) cudaEventDestroy(g.visDone);
  if (g.auxReady) cudaEventDestroy(g.auxReady);
  if (g.auxDone) cudaEventDestroy(g.auxDone);
  for (auto &v : g.vis) if (v) cudaFree(v);
  if (g.dummyTexels) cudaFree(g.dummyTexels);
  dropHaloMaps();
  if (g.dTileOwner) cudaFree(g.dTileOwner);
  if (g.dOwnerPrefix) cudaFree(g.dOwnerPrefix);
  if (g.dTileTimes) cudaFree(g.dTileTimes);
  for (void *m : g.peerMaps) cudaIpcCloseMemHandle(m);
  for (void *a : g.peerAllocs) cudaFree(a);
  if (g.dCounters) cudaFree(g.dCounters);
  if (g.hOverflow) cudaFreeHost(g.hOverflow);
  if (g.evBegin) cudaEventDestroy(g.evBegin);
  if (g.evEnd) cudaEventDestroy(g.evEnd);
  if (g.ownStream && g.stream) cudaStreamDestroy(g.stream);
  g = Ctx();
  return SGL_OK;
}

int sgl_set_stream(void *cuda_stream) {
  NEED_CTX();
  { int rc = syncAll(); if (rc) return rc; }
  if (g.ownStream && g.stream) cudaStreamDestroy(g.stream);
  if (cuda_stream) {
    g.stream = (cudaStream_t) cuda_stream;
    g.ownStream = false;
  } else {
    CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    g.ownStream = true;
  }
  return SGL_OK;
}

int sgl_wait_idle(void) {
  NEED_CTX();
  { int rc = syncAll(); if (rc) return rc; }
  CU(cudaStreamSynchronize(g.copyStream));
  return checkOverflow();
}

int sgl_get_counters(SglCounters *out) {
  NEED_CTX();
  unsigned long long c[40];
  { int rc = syncAll(); if (rc) return rc; }
  CU(cudaMemcpy(c, g.dCounters, sizeof(c), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 32; k++) c[4] += c[8 + k];
  out->passes = g.hostPasses;
  out->draws = g.hostDraws;
  out->primitives_in = c[2];
  out->primitives_binned = c[3];
  out->fragments_shaded = c[4];
  out->samples_written = c[5];
  out->kernel_launches = g.hostLaunches;
  out->clip_overflow = c[7];
  out->h2d_bytes = g.hostH2D;
  out->d2h_bytes = g.hostD2H;
  out->host_ns_pass_end = g.hostNsPassEnd;
  out->host_ns_draw = g.hostNsDraw;
  if (getenv("SGL_HOST_PROFILE") && g.hostPasses) {
    fprintf(stderr, "[sgl host profile] passes %llu:", g.hostPasses);
    for (int k = 0; k < 7; k++) fprintf(stderr, " %s %.1f us/pass;", gHostSecName[k], gHostSec[k] / 1e3 / g.hostPasses);
    fprintf(stderr, "\n");
  }
  out->bin_spills = c[1];
  out->vertices_in = g.hostVertices;
  out->host_ns_wait_gpu = g.hostNsWaitGpu;
  out->indices_in = g.hostIndices;
  out->early_vis = g.hostEarlyVis;
  out->renamed_passes = g.hostRenames;
  return checkOverflow();
}

int sgl_reset_counters(void) {
  NEED_CTX();
  { int rc = syncAll(); if (rc) return rc; }
  CU(cudaMemset(g.dCounters, 0, 40 * sizeof(unsigned long long)));
  for (auto &v : gHostSec) v = 0;
  g.hostEarlyVis = g.hostRenames = 0;
  g.hostLaunches = g.hostPasses = g.hostDraws = g.hostH2D = g.hostD2H = g.hostNsPassEnd = g.hostNsDraw = g.hostVertices = g.hostIndices = g.hostNsWaitGpu = 0;
  return SGL_OK;
}

int sgl_timer_begin(void) {
  NEED_CTX();
  CU(cudaEventRecord(g.evBegin, g.stream));
  return SGL_OK;
}

int sgl_timer_end(float *ms_out) {
  NEED_CTX();
  CU(cudaEventRecord(g.evEnd, g.stream));
  CU(cudaEventSynchronize(g.evEnd));
  CU(cudaEventElapsedTime(ms_out, g.evBegin, g.evEnd));
  return SGL_OK;
}

int sgl_set_profiling(int on) {
  NEED_CTX();
  gProfiling = on != 0;
  return SGL_OK;
}

int sgl_get_kernel_times(SglKernelTime *out, int capacity) {
  if (!g.ready) return 0;
  cudaStreamSynchronize(g.stream);
  for (cudaStream_t gs : g.geomStreams) if (gs) cudaStreamSynchronize(gs);
  int n = 0;
  const bool dump = getenv("SGL_PROFILE_OVERLAP") != nullptr && !gProf.empty();
  for (auto &p : gProf) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    if (dump) {   // timeline relative to the first recorded launch
      float t0 = 0.f;
      cudaEventElapsedTime(&t0, gProf.front().a, p.a);
      fprintf(stderr, "[sgl timeline] %-22s start %9.1f us  dur %7.1f us\n", p.name, t0 * 1e3f, ms * 1e3f);
    }
    int k = 0;
    for (; k < n; k++)
      if (!strcmp(out[k].name, p.name)) break;
    if (k == n) {
      if (n >= capacity) continue;
      memset(&out[n], 0, sizeof(SglKernelTime));
      strncpy(out[n].name, p.name, sizeof(out[n].name) - 1);
      n++;
    }
    out[k].launches++;
    out[k].total_ms += ms;
    gProfPool.push_back(p.a);
    gProfPool.push_back(p.b);
  }
  gProf.clear();
  return n;
}

// ---- reflection ---------------------------------------------------------------------------------------------
int sgl_shader_uniform_offset(int shader, const char *name) {
  const ShaderMeta *m = shaderMeta(shader);
  if (!m || !name) return -1;
  for (int i = 0; i < 4 && m->blocks[i]; i++)
    if (!strcmp(m->blocks[i], name)) return m->blockOffsets[i];
  return -1;
}
int sgl_shader_sampler_slot(int shader, const char *name) {
  const ShaderMeta *m = shaderMeta(shader);
  if (!m || !name) return -1;
  for (int i = 0; i < 8 && m->samplers[i]; i++)
    if (!strcmp(m->samplers[i], name)) return i;
  return -1;
}
int sgl_shader_define_bit(int shader, const char *name) {
  const ShaderMeta *m = shaderMeta(shader);
  if (!m || !name) return -1;
  for (int i = 0; i < 8 && m->defines[i]; i++)
    if (!strcmp(m->defines[i], name)) return i;
  return -1;
}
int sgl_shader_uniform_size(int shader) { return shaderMeta(shader) ? sglShaderInfo(shader).uniformBytes : -1; }
int sgl_shader_varying_floats(int shader) { return shaderMeta(shader) ? sglShaderInfo(shader).varyingCount : -1; }

// ---- buffers --------------------------------------------------------------------------------------------------
int sgl_buffer_create(size_t bytes, const void *host_data, int *handle_out) {
  NEED_CTX();
  BufferRec b;
  b.bytes = bytes;
  CU(cudaMalloc(&b.d, std::max<size_t>(bytes, 16)));
  if (host_data && bytes) CU(cudaMemcpyAsync(b.d, host_data, bytes, cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));   // caller's memory may go away (VertexArrayObjectSoft copies too, VertexSoft.h:16-27)
  g.buffers.push_back(b);
  *handle_out = (int) g.buffers.size() - 1;
  return SGL_OK;
}

int sgl_buffer_upload(int handle, size_t offset, size_t bytes, const void *host_data) {
  NEED_CTX();
  if (handle <= 0 || handle >= (int) g.buffers.size() || !g.buffers[handle].d) return fail(SGL_ERR_INVALID, "bad buffer handle %d", handle);
  BufferRec &b = g.buffers[handle];
  if (offset > b.bytes) return SGL_OK;
  bytes = std::min(bytes, b.bytes - offset);   // updateVertexData clamps to the buffer size (VertexSoft.h:29-31)
  { int rc = syncAll(); if (rc) return rc; }   // passes already submitted read the old contents
  CU(cudaMemcpyAsync((uint8_t *) b.d + offset, host_data, bytes, cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return SGL_OK;
}

int sgl_buffer_destroy(int handle) {
  NEED_CTX();
  if (handle <= 0 || handle >= (int) g.buffers.size()) return fail(SGL_ERR_INVALID, "bad buffer handle %d", handle);
  { int rc = syncAll(); if (rc) return rc; }
  if (g.buffers[handle].d) CU(cudaFree(g.buffers[handle].d));
  g.buffers[handle] = BufferRec();
  return SGL_OK;
}

// ---- textures -------------------------------------------------------------------------------------------------
int sgl_texture_create(const SglTextureDesc *desc, int *handle_out) {
  NEED_CTX();
  if (!desc || desc->width <= 0 || desc->height <= 0) return fail(SGL_ERR_INVALID, "texture size %dx%d", desc ? desc->width : 0, desc ? desc->height : 0);
  if (desc->multi_sample && desc->layout != SGL_LAYOUT_LINEAR) return fail(SGL_ERR_INVALID, "multisample textures must be linear");
  TextureRec t;
  t.desc = *desc;
  t.alive = true;
  SglTexObj &o = t.obj;
  o.width = desc->width;
  o.height = desc->height;
  o.levels = levelCount(*desc);
  if (o.levels > SGL_MAX_LEVELS) return fail(SGL_ERR_INVALID, "too many mip levels");
  o.layers = desc->type == SGL_TEX_CUBE ? 6 : 1;
  o.format = desc->format;
  o.samples = desc->multi_sample ? 4 : 1;
  o.layout = desc->layout;
  size_t offBytes = 0;
  for (int l = 0; l < o.levels; l++) {
    o.levelOffset[l] = offBytes;
    size_t texels = sglLevelTexels(o.layout, sglLevelDim(o.width, l), sglLevelDim(o.height, l));
    offBytes += alignUp(texels * 4 * o.samples, 256);
  }
  o.layerStride = offBytes;
  t.bytes = offBytes * o.layers;
  cudaError_t e = cudaMalloc(&o.base, t.bytes);
  if (e != cudaSuccess) return fail(SGL_ERR_OOM, "texture of %zu bytes: %s", t.bytes, cudaGetErrorString(e));
  CU(cudaMemsetAsync(o.base, 0, t.bytes, g.stream));
  if (desc->multi_sample && desc->format == SGL_FMT_RGBA8) {
    CU(cudaMalloc(&o.resolve, (size_t) o.width * o.height * 4));
    CU(cudaMemsetAsync(o.resolve, 0, (size_t) o.width * o.height * 4, g.stream));
    if (o.levels == 1 && o.layers == 1 && !g.noMsMask) {   // 0 everywhere = "samples equal the resolved colour" = the zeroed image
      CU(cudaMalloc(&t.cmask, (size_t) o.width * o.height));
      CU(cudaMemsetAsync(t.cmask, 0, (size_t) o.width * o.height, g.stream));
    }
  }
#ifdef SGL_TOUCH_BITMAP
  {
    const size_t words = t.bytes / 32 / 32 + 2;
    CU(cudaMalloc(&t.obj.touch, words * 4));
    CU(cudaMemsetAsync(t.obj.touch, 0, words * 4, g.stream));
  }
#endif
  g.textures.push_back(t);
  int h = (int) g.textures.size() - 1;
  *handle_out = h;
  return uploadTexObj(h);
}

// instrumentation build only (-DSGL_TOUCH_BITMAP): bytes of `handle` (whole 32-byte sectors) that samplers have read since
// the last reset; handle 0 = sum over all textures.  The product library answers SGL_ERR_STATE.
int sgl_debug_texel_touch(int handle, int reset, unsigned long long *bytes_out) {
  NEED_CTX();
#ifdef SGL_TOUCH_BITMAP
  { int rc = syncAll(); if (rc) return rc; }
  unsigned long long total = 0;
  for (int h = 1; h < (int) g.textures.size(); h++) {
    TextureRec &t = g.textures[h];
    if (!t.alive || !t.obj.touch || (handle != 0 && handle != h)) continue;
    const size_t words = t.bytes / 32 / 32 + 2;
    std::vector<uint32_t> bits(words);
    CU(cudaMemcpy(bits.data(), t.obj.touch, words * 4, cudaMemcpyDeviceToHost));
    for (uint32_t w : bits) total += (unsigned long long) __builtin_popcount(w) * 32ull;
    if (reset) CU(cudaMemset(t.obj.touch, 0, words * 4));
  }
  if (bytes_out) *bytes_out = total;
  return SGL_OK;
#else
  (void) handle; (void) reset; (void) bytes_out;
  return fail(SGL_ERR_STATE, "sgl_debug_texel_touch needs the instrumentation build (python -m softglrender_b200.build --variant touch -DSGL_TOUCH_BITMAP)");
#endif
}

int sgl_texture_destroy(int handle) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t) return fail(SGL_ERR_INVALID, "bad texture handle %d", handle);
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaStreamSynchronize(g.copyStream));
  if (t->obj.base) CU(cudaFree(t->obj.base));
  if (t->altBase) CU(cudaFree(t->altBase));
  t->altBase = nullptr;
  if (t->cmask) CU(cudaFree(t->cmask));
  t->cmask = nullptr;
  if (t->obj.resolve) CU(cudaFree(t->obj.resolve));
#ifdef SGL_TOUCH_BITMAP
  if (t->obj.touch) CU(cudaFree(t->obj.touch));
#endif
  if (t->rbDone) cudaEventDestroy(t->rbDone);
  *t = TextureRec();
  return SGL_OK;
}

int sgl_texture_level_size(int handle, int level, int *w_out, int *h_out) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t || level < 0 || level >= t->obj.levels) return fail(SGL_ERR_INVALID, "bad texture/level %d/%d", handle, level);
  *w_out = sglLevelDim(t->obj.width, level);
  *h_out = sglLevelDim(t->obj.height, level);
  return SGL_OK;
}

int sgl_texture_upload(int handle, int layer, int level, const void *host_data) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t || layer < 0 || layer >= t->obj.layers || level < 0 || level >= t->obj.levels) return fail(SGL_ERR_INVALID, "bad texture upload target");
  if (t->obj.samples != 1) return fail(SGL_ERR_INVALID, "setImageData not supported on multisample textures");   // TextureSoft.h:115-118
  if (t->rbPending) {   // an asynchronous read-back still reads the old contents
    CU(cudaStreamWaitEvent(g.stream, t->rbDone, 0));
    t->rbPending = false;
  }
  int w = sglLevelDim(t->obj.width, level), h = sglLevelDim(t->obj.height, level);
  size_t bytes = (size_t) w * h * 4;
  uint8_t *dst = levelPtr(*t, layer, level);
  g.hostH2D += bytes;
  if (t->obj.layout == SGL_LAYOUT_LINEAR) {
    CU(cudaMemcpyAsync(dst, host_data, bytes, cudaMemcpyHostToDevice, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return SGL_OK;
  }
  void *tmp = nullptr;
  CU(cudaMalloc(&tmp, bytes));
  CU(cudaMemcpyAsync(tmp, host_data, bytes, cudaMemcpyHostToDevice, g.stream));
  dim3 blk(16, 16), grd((w + 15) / 16, (h + 15) / 16);
  int rc = launch("sglRelayoutKernel", sglRelayoutKernel, grd, blk, (uint32_t *) dst, (const uint32_t *) tmp, w, h, t->obj.layout, 1);
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaFree(tmp));
  return rc;
}

int sgl_texture_gen_mips(int handle) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t) return fail(SGL_ERR_INVALID, "bad texture handle %d", handle);
  for (int layer = 0; layer < t->obj.layers; layer++)
    for (int level = 1; level < t->obj.levels; level++) {
      int w = sglLevelDim(t->obj.width, level), h = sglLevelDim(t->obj.height, level);
      dim3 blk(16, 16), grd((w + 15) / 16, (h + 15) / 16);
      int rc = launch("sglMipKernel", sglMipKernel, grd, blk, t->obj, layer, level);
      if (rc) return rc;
    }
  return SGL_OK;
}

int sgl_texture_device_ptr(int handle, int layer, int level, int kind, void **ptr_out, size_t *bytes_out) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t || layer < 0 || layer >= t->obj.layers || level < 0 || level >= t->obj.levels) return fail(SGL_ERR_INVALID, "bad texture/layer/level");
  int w = sglLevelDim(t->obj.width, level), h = sglLevelDim(t->obj.height, level);
  t->exposed = true;
  if (kind == 1) {
    if (!t->obj.resolve) return fail(SGL_ERR_INVALID, "texture has no resolved colour buffer");
    *ptr_out = t->obj.resolve;
    *bytes_out = (size_t) w * h * 4;
  } else {
    if (t->cmask) {
      int rc = expandMsColor(t);
      if (rc) return rc;
      t->cmaskOff = true;
    }
    *ptr_out = levelPtr(*t, layer, level);
    *bytes_out = sglLevelTexels(t->obj.layout, w, h) * 4 * t->obj.samples;
  }
  return SGL_OK;
}

int sgl_texture_readback(int handle, int layer, int level, int kind, void *host_out, size_t bytes) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t || layer < 0 || layer >= t->obj.layers || level < 0 || level >= t->obj.levels) return fail(SGL_ERR_INVALID, "bad texture/layer/level");
  int w = sglLevelDim(t->obj.width, level), h = sglLevelDim(t->obj.height, level);
  if (kind != 1 && t->cmask) { int rc = expandMsColor(t); if (rc) return rc; }
  CU(cudaStreamSynchronize(g.stream));
  { int rc = checkOverflow(); if (rc) return rc; }
  if (kind == 1) {
    if (!t->obj.resolve) return fail(SGL_ERR_INVALID, "texture has no resolved colour buffer");
    size_t need = (size_t) w * h * 4;
    if (bytes < need) return fail(SGL_ERR_INVALID, "readback buffer too small");
    CU(cudaMemcpy(host_out, t->obj.resolve, need, cudaMemcpyDeviceToHost));
    g.hostD2H += need;
    return SGL_OK;
  }
  size_t need = kind == 2 ? sglLevelTexels(t->obj.layout, w, h) * 4 * t->obj.samples : (size_t) w * h * 4 * t->obj.samples;
  if (bytes < need) return fail(SGL_ERR_INVALID, "readback buffer too small");
  uint8_t *src = levelPtr(*t, layer, level);
  g.hostD2H += need;
  if (t->obj.layout == SGL_LAYOUT_LINEAR || kind == 2) {
    CU(cudaMemcpy(host_out, src, need, cudaMemcpyDeviceToHost));
    return SGL_OK;
  }
  void *tmp = nullptr;
  CU(cudaMalloc(&tmp, need));
  dim3 blk(16, 16), grd((w + 15) / 16, (h + 15) / 16);
  int rc = launch("sglRelayoutKernel", sglRelayoutKernel, grd, blk, (uint32_t *) tmp, (const uint32_t *) src, w, h, t->obj.layout, 0);
  if (rc) return rc;
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaMemcpy(host_out, tmp, need, cudaMemcpyDeviceToHost));
  CU(cudaFree(tmp));
  return SGL_OK;
}

// Pipelined read-back: the copy is queued on a second stream behind everything submitted so far and runs while the
// next frame's vertex / setup / binning / visibility kernels execute; a later pass that overwrites the image waits for
// the copy on the device (sgl_pass_end), so the caller never has to fence.  The host buffer should be pinned.
int sgl_texture_readback_async(int handle, int layer, int level, int kind, void *host_out, size_t bytes) {
  // a copy-stream read of an RGBA8 image puts no work on the rendering stream and no visibility kernel touches colour: the
  // next pass may still start its visibility kernel early (mainSeq restored below)
  const unsigned long long seqAtEntry = g.mainSeq;
  const bool auxWasPending = g.auxPending;
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t || layer < 0 || layer >= t->obj.layers || level < 0 || level >= t->obj.levels) return fail(SGL_ERR_INVALID, "bad texture/layer/level");
  if (t->obj.format != SGL_FMT_RGBA8 || t->obj.layout != SGL_LAYOUT_LINEAR) return fail(SGL_ERR_INVALID, "async read-back needs a linear RGBA8 texture");
  int w = sglLevelDim(t->obj.width, level), h = sglLevelDim(t->obj.height, level);
  const uint8_t *src;
  size_t need;
  if (kind == 1) {
    if (!t->obj.resolve) return fail(SGL_ERR_INVALID, "texture has no resolved colour buffer");
    src = t->obj.resolve;
    need = (size_t) w * h * 4;
  } else {
    src = levelPtr(*t, layer, level);
    need = (size_t) w * h * 4 * t->obj.samples;
  }
  if (bytes < need) return fail(SGL_ERR_INVALID, "readback buffer too small");
  bool expanded = false;
  if (kind != 1 && t->cmask) { int rc = expandMsColor(t); if (rc) return rc; expanded = true; }
  if (!t->rbDone) CU(cudaEventCreateWithFlags(&t->rbDone, cudaEventDisableTiming));
  cudaPointerAttributes pa;
  const bool toDevice = cudaPointerGetAttributes(&pa, host_out) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();   // unregistered host memory reports an error on old drivers: not ours to keep
  if (!toDevice && g.rbMode == 1 && need % 16 == 0) {
    // to the host, snapshot taken by a copy kernel on the RENDERING stream (a few microseconds at HBM speed, in stream order
    // behind the pass that produced the image, so no later pass ever waits for anything); only the PCIe copy of the
    // snapshot runs on the copy stream.  The staging buffer is reused two read-backs later, after its PCIe copy.
    const int k = g.rbStageNext;
    g.rbStageNext ^= 1;
    if (g.rbStageCap[k] < need) {
      CU(cudaStreamSynchronize(g.copyStream));
      if (g.rbStage[k]) CU(cudaFree(g.rbStage[k]));
      g.rbStage[k] = nullptr;
      g.rbStageCap[k] = 0;
      cudaError_t e = cudaMalloc(&g.rbStage[k], need);
      if (e != cudaSuccess) return fail(SGL_ERR_OOM, "read-back staging of %zu bytes: %s", need, cudaGetErrorString(e));
      g.rbStageCap[k] = need;
    }
    if (!g.rbStageFree[k]) CU(cudaEventCreateWithFlags(&g.rbStageFree[k], cudaEventDisableTiming));
    else CU(cudaStreamWaitEvent(g.stream, g.rbStageFree[k], 0));
    const size_t n16 = need / 16;
    int rc = launch("sglCopy16Kernel", sglCopy16Kernel, dim3((unsigned) std::min<size_t>((n16 + 255) / 256, 148 * 8)), dim3(256),
                    (uint4 *) g.rbStage[k], (const uint4 *) src, n16);
    if (rc) return rc;
    CU(cudaEventRecord(g.copyReady, g.stream));
    CU(cudaStreamWaitEvent(g.copyStream, g.copyReady, 0));
    CU(cudaMemcpyAsync(host_out, g.rbStage[k], need, cudaMemcpyDeviceToHost, g.copyStream));
    CU(cudaEventRecord(g.rbStageFree[k], g.copyStream));
    g.hostD2H += need;
    return SGL_OK;
  }
  CU(cudaEventRecord(g.copyReady, g.stream));
  CU(cudaStreamWaitEvent(g.copyStream, g.copyReady, 0));
  if (toDevice) {
    // the destination is device memory -- another GPU's frame store mapped with sgl_peer_open (copy-engine form of the
    // multi-GPU gather): one copy
    CU(cudaMemcpyAsync(host_out, src, need, cudaMemcpyDefault, g.copyStream));
    CU(cudaEventRecord(t->rbDone, g.copyStream));
  } else {
    // to the host: the image is first snapshotted into a device staging buffer (microseconds at HBM speed) and the PCIe
    // copy reads the snapshot.  The next pass that overwrites the image only waits for the snapshot -- with one resolve
    // buffer and a 170-us PCIe copy of a 1080p frame the shading kernel of frame f+1 otherwise stalls behind the copy of frame f
    const int k = g.rbStageNext;
    g.rbStageNext ^= 1;
    if (g.rbStageCap[k] < need) {
      CU(cudaStreamSynchronize(g.copyStream));
      if (g.rbStage[k]) CU(cudaFree(g.rbStage[k]));
      g.rbStage[k] = nullptr;
      g.rbStageCap[k] = 0;
      cudaError_t e = cudaMalloc(&g.rbStage[k], need);
      if (e != cudaSuccess) return fail(SGL_ERR_OOM, "read-back staging of %zu bytes: %s", need, cudaGetErrorString(e));
      g.rbStageCap[k] = need;
    }
    CU(cudaMemcpyAsync(g.rbStage[k], src, need, cudaMemcpyDeviceToDevice, g.copyStream));
    CU(cudaEventRecord(t->rbDone, g.copyStream));
    CU(cudaMemcpyAsync(host_out, g.rbStage[k], need, cudaMemcpyDeviceToHost, g.copyStream));
    g.hostD2H += need;
  }
  t->rbPending = true;
  if (!auxWasPending && !expanded) g.mainSeq = seqAtEntry;
  return SGL_OK;
}

int sgl_readback_wait(void) {
  NEED_CTX();
  CU(cudaStreamSynchronize(g.copyStream));
  return checkOverflow();
}

// ---- render pass ----------------------------------------------------------------------------------------------
int sgl_pass_begin(int color_tex, int color_layer, int color_level, int depth_tex, int clear_color_flag,
                   int clear_depth_flag, const float clear_color[4], float clear_depth) {
  NEED_CTX_RECORDING();
  if (g.inPass) return fail(SGL_ERR_STATE, "sgl_pass_begin inside a pass");
  if (color_tex && !tex(color_tex)) return fail(SGL_ERR_INVALID, "bad colour attachment %d", color_tex);
  if (depth_tex && !tex(depth_tex)) return fail(SGL_ERR_INVALID, "bad depth attachment %d", depth_tex);
  if (!color_tex && !depth_tex) return fail(SGL_ERR_INVALID, "render pass without attachments");
  if (color_tex) {
    TextureRec *t = tex(color_tex);
    if (t->obj.format != SGL_FMT_RGBA8 || t->obj.layout != SGL_LAYOUT_LINEAR) return fail(SGL_ERR_INVALID, "colour attachment must be linear RGBA8");
    if (color_layer < 0 || color_layer >= t->obj.layers || color_level < 0 || color_level >= t->obj.levels)
      return fail(SGL_ERR_INVALID, "colour attachment layer/level out of range");
  }
  if (depth_tex) {
    TextureRec *t = tex(depth_tex);
    if (t->obj.format != SGL_FMT_FLOAT32 || t->obj.layout != SGL_LAYOUT_LINEAR) return fail(SGL_ERR_INVALID, "depth attachment must be linear FLOAT32");
  }
  if (color_tex && depth_tex) {
    TextureRec *c = tex(color_tex), *d = tex(depth_tex);
    if (c->obj.samples != d->obj.samples) return fail(SGL_ERR_INVALID, "attachment sample counts differ");
    if (sglLevelDim(c->obj.width, color_level) != d->obj.width || sglLevelDim(c->obj.height, color_level) != d->obj.height)
      return fail(SGL_ERR_INVALID, "attachment sizes differ");
  }
  g.inPass = true;
  g.colorTex = color_tex;
  g.colorLayer = color_layer;
  g.colorLevel = color_level;
  g.depthTex = depth_tex;
  g.clearColorFlag = clear_color_flag;
  g.clearDepthFlag = clear_depth_flag;
  if (clear_color) memcpy(g.clearColor, clear_color, 16);
  g.clearDepth = clear_depth;
  g.draws.clear();
  return SGL_OK;
}

int sgl_set_viewport(int x, int y, int width, int height) {
  NEED_CTX_RECORDING();
  g.vpX = (float) x;
  g.vpY = (float) y;
  g.vpW = (float) width;
  g.vpH = (float) height;
  return SGL_OK;
}

namespace {
struct HostTimer {
  unsigned long long &acc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit HostTimer(unsigned long long &a) : acc(a) {}
  ~HostTimer() { acc += (unsigned long long) std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};
}  // namespace

int sgl_draw(const SglDraw *draw) {
  NEED_CTX_RECORDING();
  HostTimer hostTimer(g.hostNsDraw);
  if (!g.inPass) return fail(SGL_ERR_STATE, "sgl_draw outside a render pass");
  if (!draw || !shaderMeta(draw->shader)) return fail(SGL_ERR_INVALID, "unknown shader %d", draw ? draw->shader : -1);
  if (draw->vertex_buffer <= 0 || draw->vertex_buffer >= (int) g.buffers.size() || !g.buffers[draw->vertex_buffer].d)
    return fail(SGL_ERR_INVALID, "bad vertex buffer");
  if (draw->index_buffer <= 0 || draw->index_buffer >= (int) g.buffers.size() || !g.buffers[draw->index_buffer].d)
    return fail(SGL_ERR_INVALID, "bad index buffer");
  if ((size_t) draw->vertex_count * SGL_VERTEX_STRIDE > g.buffers[draw->vertex_buffer].bytes ||
      (size_t) draw->index_count * 4 > g.buffers[draw->index_buffer].bytes)
    return fail(SGL_ERR_INVALID, "draw exceeds buffer size");
  SglDrawRec r;
  memset(&r, 0, sizeof(r));
  size_t ub = std::min<size_t>(draw->uniform_bytes, SGL_MAX_UNIFORM_BYTES);
  memcpy(r.uniforms, draw->uniforms, ub);
  for (int s = 0; s < SGL_MAX_SAMPLER_SLOTS; s++) {
    const SglSamplerBinding &b = draw->samplers[s];
    TextureRec *t = b.texture ? tex(b.texture) : nullptr;
    r.samplers[s].tex = t ? b.texture : -1;
    r.samplers[s].filter = b.filter_min;
    r.samplers[s].wrap = b.wrap;
    // TextureSoft::getBorderColor (TextureSoft.h:158-164): float -> clamp(r), RGBA -> clamp(c*255)
    float bc = b.border == SGL_BORDER_WHITE ? 1.f : 0.f;
    if (t && t->obj.format == SGL_FMT_FLOAT32) memcpy(&r.samplers[s].border, &bc, 4);
    else r.samplers[s].border = b.border == SGL_BORDER_WHITE ? 0xFFFFFFFFu : 0u;
    if (t && sglSamplerIsSimple(draw->shader, s, t->obj, b.filter_min, b.wrap)) r.fastSamplers |= 1u << s;
  }
  r.rs = draw->states;
  r.shader = draw->shader;
  r.defines = draw->defines;
  r.vpX = g.vpX; r.vpY = g.vpY; r.vpW = g.vpW; r.vpH = g.vpH;
  r.vertexIn = (const float *) g.buffers[draw->vertex_buffer].d;
  r.indices = (const int32_t *) g.buffers[draw->index_buffer].d;
  r.vertexCount = draw->vertex_count;
  r.indexCount = draw->index_count;
  SglShaderInfo info = sglShaderInfo(draw->shader);
  r.varyingStride = info.varyingStride;
  r.varyingCount = info.varyingCount;
  // gl_PointSize: only ShaderBasic::VS writes it (BasicSoft.h:68); other programs keep the builtin's 1.0 (ShaderSoft.h:30)
  r.pointSize = 1.f;
  if (draw->shader == SGL_SHADER_BASIC) memcpy(&r.pointSize, r.uniforms + 268, 4);
  r.hasColor = g.colorTex != 0;
  g.draws.push_back(r);
  return SGL_OK;
}

namespace {
// One (sub-)pass over `draws` into the current attachments.  A render pass is normally one call; a pass whose tail
// needs the ordered fused kernel (blending, lines/points of programs with varyings) is run as two: the opaque head takes
// the deferred visibility + shading path, the tail loads what the head wrote (clear flags off).
int runPass(std::vector<SglDrawRec> &draws, int clearColorFlag, int clearDepthFlag, bool tailOfSplit) {
  TextureRec *ct = g.colorTex ? tex(g.colorTex) : nullptr;
  TextureRec *dt = g.depthTex ? tex(g.depthTex) : nullptr;
  const int fbW = ct ? sglLevelDim(ct->obj.width, g.colorLevel) : dt->obj.width;
  const int fbH = ct ? sglLevelDim(ct->obj.height, g.colorLevel) : dt->obj.height;
  const int samples = ct ? ct->obj.samples : dt->obj.samples;
  const int tilesX = (fbW + SGL_TILE - 1) / SGL_TILE, tilesY = (fbH + SGL_TILE - 1) / SGL_TILE;
  const int nTiles = tilesX * tilesY;
  const int nDraws = (int) draws.size();

  // Depth-only passes whose result is order independent (filled triangles, depth test + write on, one direction of
  // depth function) skip binning entirely: prim-parallel setup with atomicMin/Max rasterisation (sgl_depth.cuh).
  bool depthOnly = !g.forceFused && !ct && dt && nDraws > 0;
  int dir = 0;
  for (int i = 0; i < nDraws && depthOnly; i++) {
    const SglRenderStates &rs = draws[i].rs;
    int f = rs.depth_func;
    int d = (f == 1 || f == 3) ? 1 : ((f == 4 || f == 6) ? 2 : 0);
    if (rs.primitive_type != SGL_PRIM_TRIANGLE || rs.polygon_mode != SGL_POLY_FILL || !rs.depth_test || !rs.depth_mask || d == 0 ||
        (dir != 0 && d != dir))
      depthOnly = false;
    dir = d;
  }
  // sort-first sharding: this rank renders its own tiles -- plus, for an attachment that later passes sample around the
  // pixel they shade (FXAA input), the tiles within the texture's halo; a texture marked "replicated" is rendered whole
  const uint8_t *tileOwner = nullptr;
  if (g.dTileOwner && g.ownerTilesX == tilesX && g.ownerTilesY == tilesY) {
    int halo = 0;
    if (ct && ct->shardHalo != 0) halo = ct->shardHalo;
    if (dt && dt->shardHalo != 0 && halo >= 0) halo = dt->shardHalo < 0 ? -1 : std::max(halo, dt->shardHalo);
    int rc0 = ownerMapForHalo(halo, &tileOwner);
    if (rc0) return rc0;
  }
  // lazy varyings: in a sharded pass most primitives reach none of this rank's tiles, so the vertex kernel computes
  // positions only and sglVaryingKernel runs the full vertex shader for the vertices of the primitives that were emitted
  const bool lazyVaryings = tileOwner != nullptr && !depthOnly && !g.noLazyVaryings;
  // renaming (see Ctx): only in the plain steady-state pattern -- nothing but event traffic on the rendering stream since the
  // last shading kernel, no other auxiliary pass in flight, whole single-level texture cleared and rendered by this rank
  // Only in a run of small passes (small frames, view farms: smallStreak), where the chain shading -> next shadow pass ->
  // next shading is what limits throughput (config 5 +13 %, config 1 +11 %); next to a large main pass the shadow pass hides
  // behind the visibility kernel anyway and an earlier start only takes SMs from the shading kernel (config 2 -3 %).
  bool renamed = false;
  if (depthOnly && clearDepthFlag && !g.noRename && g.smallStreak >= 2 && !g.noOverlap && !gProfiling && !tileOwner && !g.auxPending && g.preShade &&
      g.seqAfterShade == g.mainSeq && !dt->exposed && dt->obj.levels == 1 && dt->obj.layers == 1 && dt->obj.samples == 1 && !dt->rbPending &&
      std::find(g.pendingTexBase.begin(), g.pendingTexBase.end(), g.depthTex) == g.pendingTexBase.end()) {
    if (!dt->altBase) {
      cudaError_t e = cudaMalloc(&dt->altBase, dt->bytes);     // once per texture
      if (e != cudaSuccess) { cudaGetLastError(); dt->altBase = nullptr; }
    }
    if (dt->altBase) {
      void *cur = dt->obj.base;
      dt->obj.base = (uint8_t *) dt->altBase;
      dt->altBase = cur;
      g.pendingTexBase.push_back(g.depthTex);
      g.hostRenames++;
      renamed = true;
    }
  }
  // ---- arena layout
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = alignUp(off + bytes, 256); return o; };
  size_t oDraws = take(sizeof(SglDrawRec) * std::max(nDraws, 1));
  // zero-initialised region: per-draw counters, big count, tile counts, tile cursors
  size_t oZero = off;
  size_t oDrawCounters = take(sizeof(int32_t) * 2 * std::max(nDraws, 1));
  size_t oBigCount = take(sizeof(uint32_t));
  size_t oBigAllCount = take(sizeof(uint32_t));
  size_t oBinReserved = take(sizeof(uint32_t));
  size_t oStreamCursor = take(sizeof(uint32_t));
  size_t oTileCount = take(sizeof(uint32_t) * nTiles);
  size_t oTileCursor = take(sizeof(uint32_t) * nTiles);
  size_t oTileClassCount = take(sizeof(uint32_t) * SGL_TILE_CLASSES);
  size_t oScanState = take(sizeof(unsigned long long) * ((nTiles + 1023) / 1024 + 1));   // look-back scan: one word per 1024 tiles + ticket
  size_t zeroBytes = off - oZero;
  size_t oTileOffset = take(sizeof(uint32_t) * (nTiles + 1));
  int primSlots = 0, keyBase = 0, maxVerts = 0, maxPrims = 0, maxSlots = 0;
  std::vector<size_t> oClip(nDraws), oFrag(nDraws), oMask(nDraws), oVout(nDraws), oVary(nDraws), oUsed(nDraws);
  size_t usedBytes = 0;
  for (int i = 0; i < nDraws; i++) {
    SglDrawRec &r = draws[i];
    const int pt = r.rs.primitive_type;
    const int per = pt == SGL_PRIM_TRIANGLE ? 3 : (pt == SGL_PRIM_LINE ? 2 : 1);
    r.inputPrims = r.indexCount / per;
    const bool fill = pt == SGL_PRIM_TRIANGLE && r.rs.polygon_mode == SGL_POLY_FILL;
    r.slotsPerPrim = (pt == SGL_PRIM_TRIANGLE && !fill) ? 3 : 1;
    // clip arena: a filled triangle allocates at most 12 vertices (2 per frustum plane) and appends at most 6 fan
    // triangles; that worst case is reserved outright for draws of up to ~5 k triangles, larger draws get 2 vertices + 1
    // fan triangle per input triangle (x clipScale, which doubles after an overflow was reported -- checkOverflow)
    int extraVerts;
    if (fill) extraVerts = (int) std::min<long long>(12LL * r.inputPrims, std::max<long long>(g.clipMinVerts, 2LL * g.clipScale * r.inputPrims));
    else if (pt == SGL_PRIM_POINT) extraVerts = 0;
    else extraVerts = (pt == SGL_PRIM_LINE ? 2 : 6) * r.inputPrims;
    r.vertexCap = r.vertexCount + extraVerts;
    r.appendCap = fill ? (int) std::min<long long>(6LL * r.inputPrims, std::max<long long>(g.clipMinFans, (long long) g.clipScale * r.inputPrims)) : 0;
    r.primBase = primSlots;
    r.appendBase = primSlots + r.inputPrims * r.slotsPerPrim;
    primSlots = r.appendBase + r.appendCap;
    r.keyBase = keyBase;
    keyBase += r.inputPrims * r.slotsPerPrim + (fill ? 6 * r.inputPrims : 0);
    oClip[i] = take((size_t) r.vertexCap * 16);
    oFrag[i] = take((size_t) r.vertexCap * 16);
    oMask[i] = take((size_t) r.vertexCap * 4);
    oVout[i] = take((size_t) std::max(extraVerts, 1) * 64);
    oVary[i] = take((size_t) r.vertexCap * std::max(r.varyingStride, 1) * 4);
    oUsed[i] = usedBytes;
    if (lazyVaryings) usedBytes = alignUp(usedBytes + (size_t) r.vertexCount, 16);     // sglVaryingKernel reads 8 flags at a time
    g.hostVertices += (unsigned long long) r.vertexCount;
    g.hostIndices += (unsigned long long) r.indexCount;
    maxVerts = std::max(maxVerts, r.vertexCount);
    maxPrims = std::max(maxPrims, r.inputPrims);
    maxSlots = std::max(maxSlots, r.inputPrims * r.slotsPerPrim + r.appendCap);
  }
  if (primSlots >= (1 << 29)) return fail(SGL_ERR_OVERFLOW, "too many primitive slots in one pass (%d)", primSlots);
  size_t oUsedAll = take(std::max<size_t>(usedBytes, 1));
  size_t oPrims = take(sizeof(SglPrim) * std::max(primSlots, 1));
  size_t oPrimVerts = take(sizeof(SglPrimVerts) * std::max(primSlots, 1));
  size_t oPrimKeys = take(sizeof(uint32_t) * std::max(primSlots, 1));
  size_t oBigList = take(sizeof(uint32_t) * std::max(primSlots, 1));
  size_t oBigAll = take(sizeof(uint32_t) * std::max(primSlots, 1));
  size_t binCapacity = std::min<size_t>(std::max<size_t>((size_t) primSlots * 8, 1 << 20), (size_t) 1 << 29);
  if (g.binCapLimit > 0) binCapacity = std::min<size_t>(binCapacity, (size_t) g.binCapLimit);
  // depth-only path: the region holds 64-byte work items instead (>= 2 per primitive slot + one per 512 framebuffer pixels)
  if (depthOnly) binCapacity = ((size_t) primSlots * 2 + (size_t) fbW * fbH / 512 + 65536) * (sizeof(SglPrim) / sizeof(uint32_t));
  size_t oBins = take(sizeof(uint32_t) * binCapacity);
  // pre-sorted tile lists (sglTileSortKernel)
  size_t oTileSorted = depthOnly ? 0 : take(sizeof(uint32_t) * binCapacity);
  size_t oTileSortedCount = depthOnly ? 0 : take(sizeof(uint32_t) * nTiles);
  size_t oTileOrder = depthOnly ? 0 : take(sizeof(uint32_t) * nTiles * SGL_TILE_CLASSES);
  // visibility-kernel work items (heavy MSAA tiles: four quarter items) and the packed per-tile record streams
  // heavy tiles (>= 40 entries) run as four quarter-tile CTAs.  MSAA: one sample per lane, at most a quarter of the tiles.
  // One sample per pixel: four triangles per pixel at a time -- this shortens the longest tiles when FEW tiles are heavy (a
  // model in front of a floor: the kernel ends with its stragglers); when most tiles are heavy (config 4's triangle soup) the
  // kernel is throughput-bound and four CTAs culling the same list are a loss, so nothing is split once more than an eighth
  // of the tiles is heavy (sglVisSplitCount)
  const int splitCap = (!depthOnly && !g.noSplit) ? (samples == 4 ? std::max(nTiles / 4, 1) : (g.noSplit1 ? 0 : std::max(nTiles / 8, 1))) : 0;
  const size_t streamCapacity = depthOnly ? 0 : std::min<size_t>((size_t) 2 * std::max(primSlots, 1) + (size_t) 16 * nTiles, binCapacity);
  size_t oWork = depthOnly ? 0 : take(sizeof(SglVisWork) * ((size_t) nTiles + 3 * (size_t) splitCap));
  size_t oStream = depthOnly ? 0 : take((size_t) 128 * std::max<size_t>(streamCapacity, 1));
  SecTimer sec(6);
  auto section = [&](int k) { sec.switchTo(k); };
  // the deep ring only for a run of small passes (a view farm); one large pass (config 2's main pass) and the ring is three
  // deep again -- with eight frames of geometry kernels in flight next to a long pixel stage config 2 lost 5 %
  g.smallStreak = (off <= SGL_SMALL_ARENA_BYTES && nTiles <= 4096) ? std::min(g.smallStreak + 1, 1 << 20) : 0;
  const int arenaSlot = g.arenaNext % (g.smallStreak > SGL_ARENAS && !g.fewArenas ? SGL_ARENAS : g.ringSize);
  Ctx::Arena &arena = g.arenas[arenaSlot];
  const cudaStream_t geomStream = g.geomStreams[arenaSlot];
  g.arenaNext = (g.arenaNext + 1) % (SGL_ARENAS * 3 * 5 * 7);   // a multiple of every ring size
  int rc = ensureArena(arena, off);
  if (rc) return rc;
  uint8_t *A = arena.mem;
  // geometry stages go to their own stream unless overlap is off or per-kernel profiling wants clean timings;
  // the arena slot is recycled only after the pixel stages of the pass that used it last have finished
  static const bool profileOverlapped = getenv("SGL_PROFILE_OVERLAP") != nullptr;   // timeline dumps (development)
  const bool overlap = !g.noOverlap && (!gProfiling || profileOverlapped);
  struct StageGuard { ~StageGuard() { gCur = nullptr; } } stageGuard;
  gCur = overlap ? geomStream : g.stream;
  if (arena.used && overlap) CU(cudaStreamWaitEvent(geomStream, arena.pixelDone, 0));
  auto toPixelStage = [&]() -> int {   // everything issued so far on the geometry stream precedes what follows
    CU(cudaEventRecord(arena.geomDone, gCur ? gCur : g.stream));   // also what the slot's staging buffer waits for on reuse
    if (gCur != g.stream) CU(cudaStreamWaitEvent(g.stream, arena.geomDone, 0));
    gCur = g.stream;
    return SGL_OK;
  };
  auto passDone = [&]() -> int {
    CU(cudaEventRecord(arena.pixelDone, g.stream));
    arena.used = true;
    g.hostDraws += nDraws;
    return SGL_OK;
  };

  for (int i = 0; i < nDraws; i++) {
    SglDrawRec &r = draws[i];
    r.clipPos = (float *) (A + oClip[i]);
    r.fragPos = (float *) (A + oFrag[i]);
    r.clipMask = (int32_t *) (A + oMask[i]);
    r.vertexOut = (float *) (A + oVout[i]);
    r.varyings = (float *) (A + oVary[i]);
    r.vertexUsed = lazyVaryings ? A + oUsedAll + oUsed[i] : nullptr;
    r.vertexCounter = (int32_t *) (A + oDrawCounters) + 2 * i;
    r.appendCounter = (int32_t *) (A + oDrawCounters) + 2 * i + 1;
  }
  section(1);
  // this slot's pinned staging buffer (fixed address: the geometry graph's copy node reads it).  The previous pass that
  // used the slot recorded geomDone behind its copy, so one event wait makes the buffer safe to overwrite.
  const size_t recBytes = sizeof(SglDrawRec) * (size_t) nDraws;
  if (nDraws) {
    if (arena.used) {   // normally long complete; when the GPU is the bottleneck this is where the host is throttled
      HostTimer waitTimer(g.hostNsWaitGpu);
      CU(cudaEventSynchronize(arena.geomDone));
    }
    if (arena.stagingCap < recBytes) {
      if (arena.stagingHost) CU(cudaFreeHost(arena.stagingHost));
      arena.stagingHost = nullptr;
      arena.stagingCap = alignUp(recBytes * 2, 1 << 16);
      CU(cudaMallocHost(&arena.stagingHost, arena.stagingCap));
    }
    memcpy(arena.stagingHost, draws.data(), recBytes);
    g.hostH2D += recBytes;
  }
  const cudaStream_t geomS = gCur;
  auto issueUpload = [&]() -> int {   // head of the geometry chain: zero the counters, upload the draw records
    if (lazyVaryings && usedBytes) CU(cudaMemsetAsync(A + oUsedAll, 0, usedBytes, curStream()));
    if (g.ceUpload) {
      CU(cudaMemsetAsync(A + oZero, 0, zeroBytes, curStream()));
      if (nDraws) CU(cudaMemcpyAsync(A + oDraws, arena.stagingHost, recBytes, cudaMemcpyHostToDevice, curStream()));
      return SGL_OK;
    }
    // one kernel: no copy-engine work in the geometry chain (the offsets are 256-byte aligned, the sizes multiples of 16)
    const size_t n16 = std::max(zeroBytes, recBytes) / 16;
    return launch("sglPassHeadKernel", sglPassHeadKernel, dim3((unsigned) std::min<size_t>((n16 + 255) / 256, 148 * 4)), dim3(256),
                  (uint4 *) (A + oZero), zeroBytes / 16, (uint4 *) (A + oDraws), (const uint4 *) arena.stagingHost, recBytes / 16);
  };

  section(0);
  SglPassParams P;
  memset(&P, 0, sizeof(P));
  P.colorBase = ct ? levelPtr(*ct, g.colorLayer, g.colorLevel) : nullptr;
  P.depthBase = dt ? (float *) levelPtr(*dt, 0, 0) : nullptr;
  P.resolveBase = (ct && samples > 1) ? ct->obj.resolve : nullptr;
  P.colorMask = (ct && samples > 1 && P.resolveBase && ct->cmask && !ct->cmaskOff && g.colorLayer == 0 && g.colorLevel == 0) ? ct->cmask : nullptr;
  P.mirrorBase = (ct && g.colorLayer == 0 && g.colorLevel == 0) ? (uint8_t *) ct->mirror : nullptr;
  P.fbW = fbW; P.fbH = fbH; P.samples = samples;
  P.clearColorFlag = clearColorFlag;
  P.clearDepthFlag = clearDepthFlag;
  P.skipEmptyTiles = (tailOfSplit && !clearColorFlag && !clearDepthFlag) ? 1 : 0;
  {  // RGBA(clear * 255) truncation (RendererSoft.cpp:72-75)
    uint32_t c = 0;
    for (int k = 0; k < 4; k++) c |= ((uint32_t) (uint8_t) (int) (g.clearColor[k] * 255.f)) << (8 * k);
    P.clearColor = c;
  }
  P.clearDepth = g.clearDepth;
  P.tilesX = tilesX; P.tilesY = tilesY;
  P.tileOwner = tileOwner;
  P.rank = g.rank;
  P.draws = (const SglDrawRec *) (A + oDraws);
  P.drawCount = nDraws;
  P.prims = (const SglPrim *) (A + oPrims);
  P.primVerts = (const SglPrimVerts *) (A + oPrimVerts);
  P.primKeys = (const uint32_t *) (A + oPrimKeys);
  P.primSlots = primSlots;
  P.tileCount = (uint32_t *) (A + oTileCount);
  P.tileOffset = (uint32_t *) (A + oTileOffset);
  P.tileCursor = (uint32_t *) (A + oTileCursor);
  P.binSlots = (uint32_t *) (A + oBins);
  P.binCapacity = (uint32_t) binCapacity;
  P.tileSorted = depthOnly ? nullptr : (uint32_t *) (A + oTileSorted);
  P.tileSortedCount = depthOnly ? nullptr : (uint32_t *) (A + oTileSortedCount);
  if (!depthOnly) { g.lastTileSortedCount = (const uint32_t *) (A + oTileSortedCount); g.lastTilesX = tilesX; g.lastTilesY = tilesY; }
  P.tileOrder = depthOnly ? nullptr : (uint32_t *) (A + oTileOrder);
  P.tileClassCount = (uint32_t *) (A + oTileClassCount);
  P.splitCap = splitCap;
  P.work = depthOnly ? nullptr : (SglVisWork *) (A + oWork);
  P.stream = depthOnly ? nullptr : (SglVisPrim *) (A + oStream);
  P.streamCursor = (uint32_t *) (A + oStreamCursor);
  P.streamCapacity = (uint32_t) streamCapacity;
  P.bigList = (uint32_t *) (A + oBigList);
  P.bigCount = (uint32_t *) (A + oBigCount);
  P.bigAll = (uint32_t *) (A + oBigAll);
  P.bigAllCount = (uint32_t *) (A + oBigAllCount);
  P.bigCapacity = (uint32_t) std::max(primSlots, 1);
  P.binReserved = (uint32_t *) (A + oBinReserved);
  P.textures = g.dTextures;
  P.counters = g.dCounters;
  P.fragCounters = g.dCounters + 8;
  if (g.tileTiming && ct && (size_t) nTiles * 4 <= g.tileTimesCap) P.tileTimes = g.dTileTimes;

  // signature of the pass for the stage graphs: everything the captured launches depend on
  uint64_t sig = hashPod(1469598103934665603ull, P);
  sig = hashPod(sig, A);
  sig = hashPod(sig, arena.stagingHost);
  {
    const long long dims[] = {(long long) oZero, (long long) zeroBytes, (long long) oDraws, (long long) recBytes, nDraws, maxVerts, maxPrims, maxSlots,
                              nTiles, primSlots, dt ? 1 : 0, depthOnly ? 1 : 0, (long long) oPrims, (long long) oPrimVerts, (long long) oPrimKeys,
                              (long long) oScanState, (long long) oBinReserved, (long long) oBins, dir, (long long) oUsedAll, (long long) usedBytes,
                              lazyVaryings ? 1 : 0};
    sig = hashBytes(sig, dims, sizeof(dims));
  }

  section(2);
  if (depthOnly) {
    SglDepthPass D;
    memset(&D, 0, sizeof(D));
    D.draws = P.draws;
    D.depthBase = P.depthBase;
    D.fbW = fbW; D.fbH = fbH;
    D.useMin = dir == 1;
    // work queue in the bin region (4 bytes x binCapacity), huge-triangle queue in the primitive-record region
    D.queue = (SglPrim *) (A + oBins);
    D.queueCount = P.bigCount;
    D.queueCapacity = (uint32_t) (binCapacity * sizeof(uint32_t) / sizeof(SglPrim));
    D.large = (SglPrim *) (A + oPrims);
    D.largeCount = P.tileCount;          // zero-initialised with the rest of the counter block
    D.largeCapacity = (uint32_t) std::max(primSlots, 1);
    D.counters = g.dCounters;
    D.overflowHost = g.dOverflow;
    D.tileOwner = P.tileOwner;
    D.tilesX = tilesX;
    D.rank = g.rank;
    // geometry stage: vertex shading (positions only: nothing in a depth-only pass reads varyings) and the primitive-parallel
    // setup (assembly, clipping, culling, work queues) -- arena only, so it does not wait for earlier pixel work
    rc = runStage(geomS, hashPod(hashPod(sig, D), 0x67656f6dull), [&]() -> int {
      int r2 = issueUpload();
      if (r2) return r2;
      if (maxVerts > 0) {
        r2 = launch("sglVertexKernel<1>", sglVertexKernel<true>, dim3((maxVerts + 127) / 128, nDraws), dim3(128), P.draws);
        if (r2) return r2;
      }
      if (maxPrims > 0) {
        profBegin(samples == 4 ? "sglDepthSetup<4>" : "sglDepthSetup<1>");
        int e = sglLaunchDepthSetup(samples, &D, maxPrims, nDraws, (void *) curStream());
        profEnd();
        g.hostLaunches += 1;
        if (e != 0) return fail(SGL_ERR_CUDA, "depth-only setup launch failed: %s", cudaGetErrorString((cudaError_t) e));
      }
      return SGL_OK;
    });
    if (rc) return rc;
    section(3);
    // pixel stage (the atomic rasteriser writes the depth attachment): on the auxiliary stream when overlap is on
    const bool aux = overlap;
    cudaStream_t pix = aux ? g.auxStream : g.stream;
    if (aux) {
      if (renamed) {
        // the other backing store was last sampled by the shading kernel BEFORE the one that may still be running
        CU(cudaStreamWaitEvent(g.auxStream, g.preShade, 0));
      } else {
        CU(cudaEventRecord(g.auxReady, g.stream));           // all earlier pixel work (it may sample this depth texture) first
        CU(cudaStreamWaitEvent(g.auxStream, g.auxReady, 0));
      }
      CU(cudaEventRecord(arena.geomDone, geomStream));
      CU(cudaStreamWaitEvent(g.auxStream, arena.geomDone, 0));
      gCur = g.auxStream;
    } else {
      rc = toPixelStage();
      if (rc) return rc;
    }
    section(5);
    rc = runStage(pix, hashPod(hashPod(sig, D), 0x70697865ull), [&]() -> int {
      if (clearDepthFlag) {
        uint32_t bits;
        memcpy(&bits, &g.clearDepth, 4);
        size_t n = (size_t) fbW * fbH * samples;
        int r2 = launch("sglFill32Kernel", sglFill32Kernel, dim3((unsigned) std::min<size_t>((n + 1023) / 1024, 148 * 8)), dim3(256), (uint32_t *) P.depthBase, bits, n);
        if (r2) return r2;
      }
      if (maxPrims > 0) {
        profBegin(samples == 4 ? "sglDepthOnly<4>" : "sglDepthOnly<1>");
        int e = sglLaunchDepthRaster(samples, &D, nTiles, (void *) curStream());
        profEnd();
        g.hostLaunches += 2;
        if (e != 0) return fail(SGL_ERR_CUDA, "depth-only kernel launch failed: %s", cudaGetErrorString((cudaError_t) e));
      }
      return SGL_OK;
    });
    if (rc) return rc;
    section(3);
    if (aux) {
      CU(cudaEventRecord(g.auxDone, g.auxStream));
      CU(cudaEventRecord(arena.pixelDone, g.auxStream));
      g.auxPending = true;
      g.auxDepthTex.push_back(g.depthTex);
      arena.used = true;
      g.hostDraws += nDraws;
      return SGL_OK;
    }
    return passDone();
  }

  rc = runStage(geomS, hashPod(sig, 0x67656f6dull), [&]() -> int {
    int r2 = issueUpload();
    if (r2) return r2;
    if (nDraws) {
      if (maxVerts > 0) {
        r2 = lazyVaryings ? launch("sglVertexKernel<1>", sglVertexKernel<true>, dim3((maxVerts + 127) / 128, nDraws), dim3(128), P.draws)
                          : launch("sglVertexKernel<0>", sglVertexKernel<false>, dim3((maxVerts + 127) / 128, nDraws), dim3(128), P.draws);
        if (r2) return r2;
      }
      if (maxPrims > 0) {
        SglSetupOut so = {(SglPrim *) (A + oPrims), (SglPrimVerts *) (A + oPrimVerts), (uint32_t *) (A + oPrimKeys)};
        SglSetupShared ss;
        ss.tileCount = P.tileCount; ss.bigList = P.bigAll; ss.bigCount = P.bigAllCount; ss.bigCapacity = P.bigCapacity;
        ss.binReserved = (uint32_t *) (A + oBinReserved); ss.binCapacity = P.binCapacity; ss.overflowHost = g.dOverflow;
        ss.counters = g.dCounters; ss.tilesX = tilesX; ss.tilesY = tilesY; ss.fbW = fbW; ss.fbH = fbH;
        ss.tileOwner = P.tileOwner; ss.rank = g.rank;
        r2 = launch("sglSetupKernel", sglSetupKernel, dim3((maxPrims + 127) / 128, nDraws), dim3(128), P.draws, so, ss, dt ? 1 : 0);
        if (r2) return r2;
        if (lazyVaryings && maxVerts > 0) {
          r2 = launch("sglVaryingKernel", sglVaryingKernel, dim3((maxVerts + 128 * SGL_VARYING_PER_THREAD - 1) / (128 * SGL_VARYING_PER_THREAD), nDraws),
                      dim3(128), P.draws);
          if (r2) return r2;
        }
      }
    }
    const bool anyPrims = nDraws && maxSlots > 0;
    const int bigGrid = std::min(std::max(primSlots, 1), 148 * 2);
    if (anyPrims) {   // big primitives: exact per-tile counts before the scan
      r2 = launch("sglBigBinKernel<0>", sglBigBinKernel<0>, dim3(bigGrid), dim3(SGL_BIGBIN_THREADS), P);
      if (r2) return r2;
    }
    r2 = launch("sglTileScanKernel", sglTileScanKernel, dim3((nTiles + 1023) / 1024), dim3(1024), (const uint32_t *) P.tileCount, P.tileOffset, nTiles, g.dCounters,
                P.tileOrder, P.tileClassCount, P.tileSortedCount, P.tileOwner, g.rank, (unsigned long long *) (A + oScanState));
    if (r2) return r2;
    if (anyPrims) {
      r2 = launch("sglBinFillKernel", sglBinFillKernel, dim3((maxSlots + 255) / 256, nDraws), dim3(256), P);
      if (r2) return r2;
      r2 = launch("sglBigBinKernel<1>", sglBigBinKernel<1>, dim3(bigGrid), dim3(SGL_BIGBIN_THREADS), P);
      if (r2) return r2;
    }
    // every tile's list in submission order (one warp per tile)
    return launch("sglTileSortKernel", sglTileSortKernel, dim3((nTiles + SGL_TILE_SORT_WARPS - 1) / SGL_TILE_SORT_WARPS), dim3(32 * SGL_TILE_SORT_WARPS), P);
  });
  if (rc) return rc;
  section(3);
  rc = toPixelStage();
  if (rc) return rc;
  const bool needAuxJoin = g.auxPending && (!overlap || std::find(g.auxDepthTex.begin(), g.auxDepthTex.end(), g.depthTex) != g.auxDepthTex.end());
  if (needAuxJoin) joinAux();   // this pass writes a depth texture an auxiliary-stream pass is still producing
  // Deferred (visibility + shading) path for passes made of opaque draws whose point/line programs have no varyings;
  // everything else (blending, wireframe with a lit program) takes the fused tile kernel.
  bool deferred = !g.forceFused;
  for (int i = 0; i < nDraws && deferred; i++) {
    const SglDrawRec &r = draws[i];
    const bool fill = r.rs.primitive_type == SGL_PRIM_TRIANGLE && r.rs.polygon_mode == SGL_POLY_FILL;
    if (r.rs.blend && ct) deferred = false;
    if (!fill && r.varyingCount != 0) deferred = false;
  }
  section(4);
  if (deferred) {
    const int vk = ct ? (g.visNext ^= 1) : 0;
    if (ct) {
      size_t need = (size_t) fbW * fbH * samples * sizeof(uint32_t);
      if (need > g.visCap[vk]) {
        CU(cudaStreamSynchronize(g.stream));
        if (g.visStream) CU(cudaStreamSynchronize(g.visStream));
        if (g.vis[vk]) CU(cudaFree(g.vis[vk]));
        g.vis[vk] = nullptr;
        g.visCap[vk] = 0;
        cudaError_t e = cudaMalloc(&g.vis[vk], need);
        if (e != cudaSuccess) return fail(SGL_ERR_OOM, "visibility buffer of %zu bytes: %s", need, cudaGetErrorString(e));
        g.visCap[vk] = need;
      }
      P.vis = g.vis[vk];
    }
    // early visibility (see Ctx): nothing but event traffic has reached the rendering stream since the last shading kernel,
    // that kernel reads the other visibility buffer, its pass does not sample this pass's depth texture, and no auxiliary
    // pass or outside party touches that texture
    bool early = overlap && ct && !g.noEarlyVis && !needAuxJoin && g.visStream && g.seqAfterShade == g.mainSeq && !P.tileTimes;
    if (early && dt) {
      if (dt->exposed || std::find(g.lastShadeSampled.begin(), g.lastShadeSampled.end(), g.depthTex) != g.lastShadeSampled.end()) early = false;
    }
    cudaStream_t visS = early ? g.visStream : g.stream;
    if (early) {
      CU(cudaStreamWaitEvent(g.visStream, arena.geomDone, 0));
      CU(cudaStreamWaitEvent(g.visStream, g.preShade, 0));
      gCur = g.visStream;
    }
    profBegin(samples == 4 ? "sglVisKernel<4>" : "sglVisKernel<1>");
    int e = sglLaunchVis(samples, &P, nTiles, (void *) visS);
    profEnd();
    g.hostLaunches++;
    if (e != 0) return fail(SGL_ERR_CUDA, "visibility kernel launch failed: %s", cudaGetErrorString((cudaError_t) e));
    if (early) {
      CU(cudaEventRecord(g.visDone, g.visStream));
      CU(cudaStreamWaitEvent(g.stream, g.visDone, 0));
      gCur = g.stream;
      g.hostEarlyVis++;
    }
    g.mainSeq++;
    if (!g.pendingTexBase.empty()) { int rcTb = flushTexBases(); if (rcTb) return rcTb; }
    joinAux();   // the shading kernel samples textures (shadow maps): depth-only passes on the auxiliary stream first
    if (ct) {
      if (ct->rbPending) {   // an asynchronous read-back of this image is still in flight: overwrite only after it
        CU(cudaStreamWaitEvent(g.stream, ct->rbDone, 0));
        ct->rbPending = false;
      }
      if (g.preShade) CU(cudaEventRecord(g.preShade, g.stream));
      profBegin(samples == 4 ? "sglShadeKernel<4>" : "sglShadeKernel<1>");
      e = samples == 4 ? sglLaunchShade4(&P, nTiles, (void *) g.stream) : sglLaunchShade1(&P, nTiles, (void *) g.stream);
      profEnd();
      g.hostLaunches++;
      if (e != 0) return fail(SGL_ERR_CUDA, "shading kernel launch failed: %s", cudaGetErrorString((cudaError_t) e));
      g.seqAfterShade = ++g.mainSeq;
      g.lastShadeSampled.clear();
      for (int i = 0; i < nDraws; i++)
        for (const SglSamplerSlot &sl : draws[i].samplers)
          if (sl.tex > 0 && std::find(g.lastShadeSampled.begin(), g.lastShadeSampled.end(), sl.tex) == g.lastShadeSampled.end())
            g.lastShadeSampled.push_back(sl.tex);
    }
  } else {
    if (!g.pendingTexBase.empty()) { int rcTb = flushTexBases(); if (rcTb) return rcTb; }
    joinAux();
    if (ct && ct->rbPending) {
      CU(cudaStreamWaitEvent(g.stream, ct->rbDone, 0));
      ct->rbPending = false;
    }
    profBegin(samples == 4 ? "sglRasterKernel<4>" : "sglRasterKernel<1>");
    int e = samples == 4 ? sglLaunchRaster4(&P, nTiles, (void *) g.stream) : sglLaunchRaster1(&P, nTiles, (void *) g.stream);
    profEnd();
    g.hostLaunches++;
    g.mainSeq++;
    if (e != 0) return fail(SGL_ERR_CUDA, "raster kernel launch failed: %s", cudaGetErrorString((cudaError_t) e));
  }
  return passDone();
}

// can this draw take the deferred (visibility + shading) path?  Blending is order dependent; point/line fragments of
// programs with varyings carry per-step varyings that the visibility buffer cannot reproduce.
bool drawIsDeferrable(const SglDrawRec &r, bool hasColor) {
  const bool fill = r.rs.primitive_type == SGL_PRIM_TRIANGLE && r.rs.polygon_mode == SGL_POLY_FILL;
  if (r.rs.blend && hasColor) return false;
  if (!fill && r.varyingCount != 0) return false;
  return true;
}
}  // namespace

int sgl_pass_end(void) {
  NEED_CTX_RECORDING();
  HostTimer hostTimer(g.hostNsPassEnd);
  if (!g.inPass) return fail(SGL_ERR_STATE, "sgl_pass_end outside a pass");
  g.inPass = false;
  TextureRec *ct = g.colorTex ? tex(g.colorTex) : nullptr;
  TextureRec *dt = g.depthTex ? tex(g.depthTex) : nullptr;
  if ((g.colorTex && !ct) || (g.depthTex && !dt)) return fail(SGL_ERR_INVALID, "attachment destroyed during pass");
  const int nDraws = (int) g.draws.size();
  int rc;
  // split point: the first draw that cannot be deferred, when a deferrable head of real size precedes it
  int k = 0;
  while (k < nDraws && drawIsDeferrable(g.draws[k], ct != nullptr)) k++;
  if (ct && !g.forceFused && !g.noPassSplit && k > 0 && k < nDraws) {
    std::vector<SglDrawRec> head(g.draws.begin(), g.draws.begin() + k), tail(g.draws.begin() + k, g.draws.end());
    rc = runPass(head, g.clearColorFlag, g.clearDepthFlag, false);
    if (rc == SGL_OK) rc = runPass(tail, 0, 0, true);
  } else {
    rc = runPass(g.draws, g.clearColorFlag, g.clearDepthFlag, false);
  }
  g.hostPasses++;
  g.draws.clear();
  return rc;
}

// testing aid: shrink the tile-bin region and the minimum clip arenas so that the spill / overflow paths can be exercised
// with small inputs (0 / negative = default)
int sgl_debug_set_limits(long long bin_capacity, long long clip_min_vertices, long long clip_min_fans) {
  NEED_CTX();
  { int rc = syncAll(); if (rc) return rc; }
  g.binCapLimit = bin_capacity > 0 ? bin_capacity : 0;
  g.clipMinVerts = clip_min_vertices > 0 ? clip_min_vertices : 65536;
  g.clipMinFans = clip_min_fans > 0 ? clip_min_fans : 32768;
  g.clipScale = 1;
  return SGL_OK;
}

// instrumentation: per-tile primitive list lengths of the most recent colour pass (SGL_TILE_UNSORTED = overflow tile)
int sgl_get_tile_list_sizes(uint32_t *out, int capacity, int *tiles_x_out, int *tiles_y_out) {
  NEED_CTX();
  if (!g.lastTileSortedCount) return fail(SGL_ERR_STATE, "no pass recorded yet");
  { int rc = syncAll(); if (rc) return rc; }
  int n = g.lastTilesX * g.lastTilesY;
  if (tiles_x_out) *tiles_x_out = g.lastTilesX;
  if (tiles_y_out) *tiles_y_out = g.lastTilesY;
  if (out) CU(cudaMemcpy(out, g.lastTileSortedCount, sizeof(uint32_t) * std::min(n, capacity), cudaMemcpyDeviceToHost));
  return SGL_OK;
}

// instrumentation: per-tile start/end time (globaltimer ns) of the visibility kernel of the most recent colour pass
int sgl_debug_tile_times(int enable, unsigned long long *out, int capacity_tiles) {
  NEED_CTX();
  { int rc = syncAll(); if (rc) return rc; }
  if (out && g.dTileTimes) CU(cudaMemcpy(out, g.dTileTimes, sizeof(unsigned long long) * 2 * std::min<size_t>(capacity_tiles, g.tileTimesCap / 2), cudaMemcpyDeviceToHost));
  g.tileTiming = enable;
  if (enable && !g.dTileTimes) {
    g.tileTimesCap = 2 * 1024 * 1024;
    CU(cudaMalloc(&g.dTileTimes, g.tileTimesCap * sizeof(unsigned long long)));
    CU(cudaMemset(g.dTileTimes, 0, g.tileTimesCap * sizeof(unsigned long long)));
  }
  return SGL_OK;
}

// ---- multi-GPU --------------------------------------------------------------------------------------------------
int sgl_tile_size(void) { return SGL_TILE; }

int sgl_set_tile_owner_map(const uint8_t *owner, int tiles_x, int tiles_y) {
  NEED_CTX();
  { int rc = syncAll(); if (rc) return rc; }
  dropHaloMaps();
  if (g.dTileOwner) CU(cudaFree(g.dTileOwner));
  if (g.dOwnerPrefix) CU(cudaFree(g.dOwnerPrefix));
  g.dTileOwner = nullptr;
  g.dOwnerPrefix = nullptr;
  g.prefixRank = -1;
  g.ownerTilesX = g.ownerTilesY = 0;
  g.hostTileOwner.clear();
  if (!owner) return SGL_OK;
  if (tiles_x <= 0 || tiles_y <= 0) return fail(SGL_ERR_INVALID, "bad tile map size");
  const size_t n = (size_t) tiles_x * tiles_y;
  CU(cudaMalloc(&g.dTileOwner, n));
  CU(cudaMemcpy(g.dTileOwner, owner, n, cudaMemcpyHostToDevice));
  CU(cudaMalloc(&g.dOwnerPrefix, (n + 1) * sizeof(uint32_t)));
  g.hostTileOwner.assign(owner, owner + n);
  g.ownerTilesX = tiles_x;
  g.ownerTilesY = tiles_y;
  return SGL_OK;
}

int sgl_set_rank(int rank, int world) {
  NEED_CTX();
  if (world < 1 || rank < 0 || rank >= world) return fail(SGL_ERR_INVALID, "rank %d of %d", rank, world);
  if (rank != g.rank) {
    int rc = syncAll();
    if (rc) return rc;
    dropHaloMaps();
  }
  g.rank = rank;
  g.world = world;
  return SGL_OK;
}

int sgl_texture_set_shard_halo(int handle, int pixels) {
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t) return fail(SGL_ERR_INVALID, "bad texture handle %d", handle);
  t->shardHalo = pixels;
  return SGL_OK;
}

int sgl_texture_set_mirror(int handle, void *device_ptr) {
  SeqKeeper keep;
  NEED_CTX();
  TextureRec *t = tex(handle);
  if (!t) return fail(SGL_ERR_INVALID, "bad texture handle %d", handle);
  if (t->obj.format != SGL_FMT_RGBA8 || t->obj.layout != SGL_LAYOUT_LINEAR || t->obj.layers != 1)
    return fail(SGL_ERR_INVALID, "mirror target needs a linear 2D RGBA8 texture");
  t->mirror = device_ptr;
  return keep.done(SGL_OK);
}

namespace {
// image (layer 0 / level 0, resolved colour for MS textures) of a tile-gather operand, and the owner-prefix table
int tileGatherOperand(int handle, int owner_rank, uint32_t **image, int *w, int *h, int *owned) {
  TextureRec *t = tex(handle);
  if (!t) return fail(SGL_ERR_INVALID, "bad texture handle %d", handle);
  if (t->obj.format != SGL_FMT_RGBA8 || t->obj.layout != SGL_LAYOUT_LINEAR) return fail(SGL_ERR_INVALID, "tile gather needs a linear RGBA8 texture");
  *w = t->obj.width;
  *h = t->obj.height;
  *image = (uint32_t *) (t->obj.samples > 1 ? t->obj.resolve : t->obj.base);
  if (!*image) return fail(SGL_ERR_INVALID, "texture has no single-sample colour image");
  const int tilesX = (*w + SGL_TILE - 1) / SGL_TILE, tilesY = (*h + SGL_TILE - 1) / SGL_TILE;
  if (!g.dTileOwner || g.ownerTilesX != tilesX || g.ownerTilesY != tilesY)
    return fail(SGL_ERR_STATE, "no tile owner map for a %dx%d-tile target", tilesX, tilesY);
  const int n = tilesX * tilesY;
  if (g.prefixRank != owner_rank) {
    int rc = launch("sglTileOwnerPrefixKernel", sglTileOwnerPrefixKernel, dim3(1), dim3(1024), (const uint8_t *) g.dTileOwner, g.dOwnerPrefix, n, owner_rank);
    if (rc) return rc;
    g.prefixRank = owner_rank;
  }
  int c = 0;
  for (int i = 0; i < n; i++) c += g.hostTileOwner[i] == owner_rank;
  *owned = c;
  return SGL_OK;
}
}  // namespace

int sgl_tiles_owned(int owner_rank, int *tiles_out) {
  NEED_CTX();
  if (g.hostTileOwner.empty()) return fail(SGL_ERR_STATE, "no tile owner map");
  int c = 0;
  for (uint8_t o : g.hostTileOwner) c += o == owner_rank;
  *tiles_out = c;
  return SGL_OK;
}

int sgl_tiles_pack(int handle, int owner_rank, void *dst_device, size_t dst_bytes, int *tiles_out) {
  NEED_CTX();
  uint32_t *image;
  int w, h, owned;
  int rc = tileGatherOperand(handle, owner_rank, &image, &w, &h, &owned);
  if (rc) return rc;
  if (tiles_out) *tiles_out = owned;
  if ((size_t) owned * SGL_TILE_THREADS * 4 > dst_bytes) return fail(SGL_ERR_INVALID, "tile staging buffer too small (%d tiles)", owned);
  return launch("sglTilePackKernel", sglTilePackKernel, dim3(g.ownerTilesX * g.ownerTilesY), dim3(SGL_TILE_THREADS), image, (uint32_t *) dst_device,
                (const uint8_t *) g.dTileOwner, (const uint32_t *) g.dOwnerPrefix, owner_rank, g.ownerTilesX, w, h, 0);
}

int sgl_tiles_unpack(int handle, int owner_rank, const void *src_device, size_t src_bytes) {
  NEED_CTX();
  uint32_t *image;
  int w, h, owned;
  int rc = tileGatherOperand(handle, owner_rank, &image, &w, &h, &owned);
  if (rc) return rc;
  if ((size_t) owned * SGL_TILE_THREADS * 4 > src_bytes) return fail(SGL_ERR_INVALID, "tile staging buffer too small (%d tiles)", owned);
  return launch("sglTilePackKernel", sglTilePackKernel, dim3(g.ownerTilesX * g.ownerTilesY), dim3(SGL_TILE_THREADS), image, (uint32_t *) src_device,
                (const uint8_t *) g.dTileOwner, (const uint32_t *) g.dOwnerPrefix, owner_rank, g.ownerTilesX, w, h, 1);
}

// ---- peer memory (one process per GPU; buffers are shared through CUDA IPC handles) --------------------------------
int sgl_peer_alloc(size_t bytes, void **ptr_out, uint8_t ipc_handle_out[64]) {
  NEED_CTX();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 256));
  if (e != cudaSuccess) return fail(SGL_ERR_OOM, "peer buffer of %zu bytes: %s", bytes, cudaGetErrorString(e));
  CU(cudaMemset(p, 0, std::max<size_t>(bytes, 256)));
  cudaIpcMemHandle_t hd;
  e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(SGL_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle_out, &hd, 64);
  g.peerAllocs.push_back(p);
  *ptr_out = p;
  return SGL_OK;
}

int sgl_peer_free(void *ptr) {
  NEED_CTX();
  auto it = std::find(g.peerAllocs.begin(), g.peerAllocs.end(), ptr);
  if (it == g.peerAllocs.end()) return fail(SGL_ERR_INVALID, "not a peer allocation");
  { int rc = syncAll(); if (rc) return rc; }
  CU(cudaStreamSynchronize(g.copyStream));
  CU(cudaFree(ptr));
  g.peerAllocs.erase(it);
  return SGL_OK;
}

int sgl_peer_open(const uint8_t ipc_handle[64], void **ptr_out) {
  NEED_CTX();
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, 64);
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(SGL_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  g.peerMaps.push_back(p);
  *ptr_out = p;
  return SGL_OK;
}

int sgl_peer_close(void *ptr) {
  NEED_CTX();
  auto it = std::find(g.peerMaps.begin(), g.peerMaps.end(), ptr);
  if (it == g.peerMaps.end()) return fail(SGL_ERR_INVALID, "not a peer mapping");
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaIpcCloseMemHandle(ptr));
  g.peerMaps.erase(it);
  return SGL_OK;
}

int sgl_peer_signal(void *flag_device_ptr, uint32_t value) {
  SeqKeeper keep;
  NEED_CTX();
  return keep.done(launch("sglPeerSignalKernel", sglPeerSignalKernel, dim3(1), dim3(1), (uint32_t *) flag_device_ptr, value));
}

int sgl_peer_signal_after_copies(void *flag_device_ptr, uint32_t value) {
  SeqKeeper keep;
  NEED_CTX();
  gCur = g.copyStream;   // ordered behind every sgl_texture_readback_async queued so far
  int rc = launch("sglPeerSignalKernel", sglPeerSignalKernel, dim3(1), dim3(1), (uint32_t *) flag_device_ptr, value);
  gCur = nullptr;
  return keep.done(rc);
}

int sgl_peer_collect(const void *done_flags, int count, uint32_t value, void *const *consumed_flags, int timeout_ms, int side_stream) {
  SeqKeeper keep;
  NEED_CTX();
  if (count < 1 || count > 64) return fail(SGL_ERR_INVALID, "peer collect over %d ranks", count);
  SglPeerPtrs pp;
  memset(&pp, 0, sizeof(pp));
  for (int i = 0; i < count; i++) pp.p[i] = consumed_flags ? (uint32_t *) consumed_flags[i] : nullptr;
  long long cycles = (long long) std::max(timeout_ms, 1) * 2000000LL;
  gCur = side_stream ? g.peerStream : g.stream;
  int rc = launch("sglPeerCollectKernel", sglPeerCollectKernel, dim3(1), dim3(64), (const uint32_t *) done_flags, count, value, pp, cycles, g.dCounters);
  gCur = nullptr;
  return keep.done(rc);
}

int sgl_peer_wait(const void *flags_device_ptr, int count, uint32_t value, int timeout_ms) {
  SeqKeeper keep;
  NEED_CTX();
  if (count < 1 || count > 1024) return fail(SGL_ERR_INVALID, "peer wait on %d flags", count);
  long long cycles = (long long) std::max(timeout_ms, 1) * 2000000LL;   // ~2 GHz SM clock
  return keep.done(launch("sglPeerWaitKernel", sglPeerWaitKernel, dim3(1), dim3((unsigned) ((count + 31) / 32 * 32)), (const uint32_t *) flags_device_ptr, count, value,
                          cycles, g.dCounters));
}

int sgl_peer_timeouts(uint64_t *count_out) {
  NEED_CTX();
  unsigned long long c = 0;
  { int rc = syncAll(); if (rc) return rc; }
  CU(cudaStreamSynchronize(g.copyStream));
  CU(cudaMemcpy(&c, g.dCounters + 6, sizeof(c), cudaMemcpyDeviceToHost));
  *count_out = c;
  return SGL_OK;
}

// ---- KATs -------------------------------------------------------------------------------------------------------

int sgl_kat_barycentric(const float *tri_xyzw, const float *sample_xy, int n, float *bc_out, int *inside_out, float *zw_out) {
  NEED_CTX();
  DevTmp<float> dTri, dXy, dBc, dZw;
  DevTmp<int> dIn;
  CU(dTri.alloc(12)); CU(dXy.alloc(2 * n)); CU(dBc.alloc(3 * n)); CU(dZw.alloc(2 * n)); CU(dIn.alloc(n));
  CU(cudaMemcpy(dTri.p, tri_xyzw, 48, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dXy.p, sample_xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice));
  int rc = launch("sglKatBarycentricKernel", sglKatBarycentricKernel, dim3((n + 127) / 128), dim3(128), (const float *) dTri.p, (const float *) dXy.p, n, dBc.p, dIn.p, dZw.p);
  if (rc) return rc;
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaMemcpy(bc_out, dBc.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(inside_out, dIn.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(zw_out, dZw.p, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
  return SGL_OK;
}

int sgl_kat_sample(int texture, int filter_min, int wrap, int border, const float *coords, const float *lod,
                   const int32_t *offsets_xy, int n, int split_phase, uint32_t *out) {
  NEED_CTX();
  TextureRec *t = tex(texture);
  if (!t) return fail(SGL_ERR_INVALID, "bad texture handle %d", texture);
  int comps = t->obj.layers == 6 ? 3 : 2;
  if (split_phase && !(t->obj.format == SGL_FMT_RGBA8 && t->obj.samples == 1 && t->obj.layout == SGL_LAYOUT_LINEAR &&
                       filter_min == SGL_FILTER_LINEAR && (wrap == SGL_WRAP_REPEAT || wrap == SGL_WRAP_CLAMP_TO_EDGE)))
    return fail(SGL_ERR_INVALID, "split-phase taps need a simple sampler");
  DevTmp<float> dC, dL;
  DevTmp<uint32_t> dO;
  DevTmp<int32_t> dOff;
  CU(dC.alloc((size_t) comps * n)); CU(dL.alloc(n)); CU(dO.alloc(n)); CU(dOff.alloc((size_t) 2 * n));
  CU(cudaMemcpy(dC.p, coords, sizeof(float) * comps * n, cudaMemcpyHostToDevice));
  if (lod) CU(cudaMemcpy(dL.p, lod, sizeof(float) * n, cudaMemcpyHostToDevice));
  if (offsets_xy) CU(cudaMemcpy(dOff.p, offsets_xy, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice));
  uint32_t b;
  float bf = border == SGL_BORDER_WHITE ? 1.f : 0.f;
  if (t->obj.format == SGL_FMT_FLOAT32) memcpy(&b, &bf, 4);
  else b = border == SGL_BORDER_WHITE ? 0xFFFFFFFFu : 0u;
  int rc = launch("sglKatSampleKernel", sglKatSampleKernel, dim3((n + 127) / 128), dim3(128), (const SglTexObj *) g.dTextures, texture, filter_min, wrap, b,
                  (const float *) dC.p, lod ? (const float *) dL.p : (const float *) nullptr,
                  offsets_xy ? (const int32_t *) dOff.p : (const int32_t *) nullptr, n, split_phase, dO.p);
  if (rc) return rc;
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaMemcpy(out, dO.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  return SGL_OK;
}

int sgl_kat_blend(const SglRenderStates *states, const float *src_rgba, const float *dst_rgba, int n, float *out_rgba) {
  NEED_CTX();
  DevTmp<float> dS, dD, dO;
  CU(dS.alloc(4 * n)); CU(dD.alloc(4 * n)); CU(dO.alloc(4 * n));
  CU(cudaMemcpy(dS.p, src_rgba, sizeof(float) * 4 * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dD.p, dst_rgba, sizeof(float) * 4 * n, cudaMemcpyHostToDevice));
  int rc = launch("sglKatBlendKernel", sglKatBlendKernel, dim3((n + 127) / 128), dim3(128), *states, (const float *) dS.p, (const float *) dD.p, n, dO.p);
  if (rc) return rc;
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaMemcpy(out_rgba, dO.p, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost));
  return SGL_OK;
}

int sgl_kat_depth(int func, const float *a, const float *b, int n, int *pass_out) {
  NEED_CTX();
  DevTmp<float> dA, dB;
  DevTmp<int> dO;
  CU(dA.alloc(n)); CU(dB.alloc(n)); CU(dO.alloc(n));
  CU(cudaMemcpy(dA.p, a, sizeof(float) * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dB.p, b, sizeof(float) * n, cudaMemcpyHostToDevice));
  int rc = launch("sglKatDepthKernel", sglKatDepthKernel, dim3((n + 127) / 128), dim3(128), func, (const float *) dA.p, (const float *) dB.p, n, dO.p);
  if (rc) return rc;
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaMemcpy(pass_out, dO.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
  return SGL_OK;
}

}  // extern "C"
