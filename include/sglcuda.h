/*
 * sglcuda.h -- C ABI of libsglcuda.so, the B200 (sm_100a) implementation of SoftGLRender's software
 * pipeline.  Plain C, plain pointers and sizes; no C++/torch types cross this boundary.
 *
 * Every entry point replaces one piece of the reference's CPU path (file:line under
 * keith2018/SoftGLRender src/).  The C++ host class RendererCUDA (softglrender_b200/host) implements the
 * reference's abstract `Renderer` interface (Render/Renderer.h:24-59) on top of these calls; a reference
 * maintainer binds them the same way.
 *
 * Conventions: every function returns SGL_OK (0) or a negative SglStatus; sgl_last_error() gives the text.
 * Handles are small positive ints, 0 is "none".  All work is queued on one CUDA stream per context and is
 * complete after sgl_wait_idle() or any *_readback call.  There is NO CPU fallback: without a CUDA device
 * sgl_init fails.
 */
#ifndef SGLCUDA_H_
#define SGLCUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum SglStatus {
  SGL_OK = 0,
  SGL_ERR_NO_DEVICE = -1,
  SGL_ERR_CUDA = -2,
  SGL_ERR_INVALID = -3,
  SGL_ERR_OOM = -4,
  SGL_ERR_STATE = -5,
  SGL_ERR_OVERFLOW = -6
} SglStatus;

/* enum values are those of the reference headers (Render/Texture.h:17-77, Render/RenderStates.h:13-79) */
enum { SGL_WRAP_REPEAT = 0, SGL_WRAP_MIRRORED_REPEAT = 1, SGL_WRAP_CLAMP_TO_EDGE = 2, SGL_WRAP_CLAMP_TO_BORDER = 3 };
enum { SGL_FILTER_NEAREST = 0, SGL_FILTER_LINEAR = 1, SGL_FILTER_NEAREST_MIPMAP_NEAREST = 2,
       SGL_FILTER_LINEAR_MIPMAP_NEAREST = 3, SGL_FILTER_NEAREST_MIPMAP_LINEAR = 4, SGL_FILTER_LINEAR_MIPMAP_LINEAR = 5 };
enum { SGL_BORDER_BLACK = 0, SGL_BORDER_WHITE = 1 };
enum { SGL_TEX_2D = 0, SGL_TEX_CUBE = 1 };
enum { SGL_FMT_RGBA8 = 0, SGL_FMT_FLOAT32 = 1 };
/* image layouts of Base/Buffer.h:21,141,172 -- run-time selectable here, compile-time in the reference */
enum { SGL_LAYOUT_LINEAR = 0, SGL_LAYOUT_TILED = 1, SGL_LAYOUT_MORTON = 2 };
enum { SGL_PRIM_POINT = 0, SGL_PRIM_LINE = 1, SGL_PRIM_TRIANGLE = 2 };
enum { SGL_POLY_POINT = 0, SGL_POLY_LINE = 1, SGL_POLY_FILL = 2 };
/* shader programs = the reference's software shaders (Viewer/Shader/Software/*.h); ids = View::ShadingModel */
enum { SGL_SHADER_BASIC = 1, SGL_SHADER_BLINNPHONG = 2, SGL_SHADER_PBR = 3, SGL_SHADER_SKYBOX = 4,
       SGL_SHADER_IBL_IRRADIANCE = 5, SGL_SHADER_IBL_PREFILTER = 6, SGL_SHADER_FXAA = 7 };

#define SGL_MAX_SAMPLER_SLOTS 8
#define SGL_MAX_UNIFORM_BYTES 512
#define SGL_VERTEX_STRIDE 64 /* Viewer/Model.h:20-25: vec3 pos@0, vec2 uv@16, vec3 normal@32, vec3 tangent@48 */

/* Render/Texture.h:67-76 TextureDesc (+ layout) */
typedef struct SglTextureDesc {
  int32_t width, height;
  int32_t type;        /* SGL_TEX_* */
  int32_t format;      /* SGL_FMT_* */
  int32_t use_mipmaps; /* allocate the full level chain (SamplerSoft.h:90-110) */
  int32_t multi_sample;/* 4x: texel = 4 samples (TextureSoft.h:20-55) */
  int32_t layout;      /* SGL_LAYOUT_* */
} SglTextureDesc;

/* Render/RenderStates.h:80-92 RenderStates as a POD */
typedef struct SglRenderStates {
  int32_t blend;
  int32_t blend_func_rgb, blend_src_rgb, blend_dst_rgb;
  int32_t blend_func_alpha, blend_src_alpha, blend_dst_alpha;
  int32_t depth_test, depth_mask, depth_func;
  int32_t cull_face;
  int32_t primitive_type, polygon_mode;
  float line_width;
} SglRenderStates;

/* sampler slot -> texture binding snapshot (UniformSamplerSoft::setTexture, UniformSoft.h:79-81:
 * filter/wrap/border are copied from the texture when it is bound, SamplerSoft.h:388-394) */
typedef struct SglSamplerBinding {
  int32_t texture;     /* texture handle, 0 = unbound */
  int32_t filter_min;
  int32_t wrap;        /* wrapS, used on both axes like the reference */
  int32_t border;      /* SGL_BORDER_* */
} SglSamplerBinding;

/* One draw() call with everything the asynchronous backend must snapshot (RendererSoft.cpp:115-164). */
typedef struct SglDraw {
  int32_t vertex_buffer;   /* sgl_buffer handle, 64-byte vertices */
  int32_t index_buffer;    /* sgl_buffer handle, int32 indices */
  int32_t vertex_count, index_count;
  int32_t shader;          /* SGL_SHADER_* */
  uint32_t defines;        /* bit i = i-th entry of the shader's define list (sgl_shader_define_bit) */
  SglRenderStates states;
  uint32_t uniform_bytes;
  uint8_t uniforms[SGL_MAX_UNIFORM_BYTES];             /* the program's uniform buffer (ShaderProgramSoft.h:61-70) */
  SglSamplerBinding samplers[SGL_MAX_SAMPLER_SLOTS];   /* slot order = sgl_shader_sampler_slot */
} SglDraw;

/* Geometry is never dropped silently: if the clip-vertex arena of a pass was too small, the next synchronising call
 * (sgl_wait_idle, any read-back, sgl_get_counters) returns SGL_ERR_OVERFLOW once, and later passes get a larger arena. */
typedef struct SglCounters {
  uint64_t passes, draws, primitives_in, primitives_binned, fragments_shaded, samples_written;
  uint64_t kernel_launches;   /* launches of kernels defined in this library */
  uint64_t clip_overflow;     /* primitives dropped because the clip-vertex arena was full (should be 0) */
  uint64_t h2d_bytes;         /* host->device bytes copied by passes (draw records incl. uniform snapshots) and uploads */
  uint64_t d2h_bytes;         /* device->host bytes copied by read-backs */
  uint64_t host_ns_pass_end;  /* CPU time spent inside sgl_pass_end (arena layout, uploads, launches) */
  uint64_t host_ns_draw;      /* CPU time spent inside sgl_draw (state snapshot) */
  uint64_t bin_spills;        /* primitives that went to the pass-wide list because the tile bins were full (correct, slower) */
  uint64_t vertices_in;       /* VAO vertices of all submitted draws (64 B each) ... */
  uint64_t indices_in;        /* ... and their indices (4 B each): B_geom of the roofline = 64 * vertices_in + 4 * indices_in */
  uint64_t host_ns_wait_gpu;  /* part of host_ns_pass_end spent BLOCKED on the GPU (arena ring full): not CPU work */
  uint64_t early_vis;         /* visibility kernels that started while the previous pass's shading kernel was still running */
  uint64_t renamed_passes;    /* depth-only passes that rendered into the texture's other backing store instead of waiting for its readers */
} SglCounters;

/* ---- context ---------------------------------------------------------------------------------------- */
int sgl_init(int device_ordinal, int rank, int world);   /* RendererSoft::create (Renderer.h:27) */
int sgl_shutdown(void);
const char *sgl_last_error(void);
int sgl_set_stream(void *cuda_stream);                   /* run on a caller-owned cudaStream_t (NULL = own stream) */
int sgl_wait_idle(void);                                 /* Renderer::waitIdle (Renderer.h:58) */
int sgl_get_counters(SglCounters *out);
int sgl_reset_counters(void);
/* device-side timing on the library's stream (CUDA events); returns elapsed ms of the last begin/end pair */
int sgl_timer_begin(void);
int sgl_timer_end(float *ms_out);

/* per-kernel device timing for bench.py's roofline block: while profiling is on every kernel launch is bracketed
 * by CUDA events on the library's stream; sgl_get_kernel_times drains them (accumulated per kernel name). */
typedef struct SglKernelTime {
  char name[48];
  uint64_t launches;
  double total_ms;
} SglKernelTime;
/* per-tile primitive list lengths of the most recent colour pass (row-major tiles; 0xFFFFFFFF = list overflowed into the
 * in-kernel path) -- load-balance instrumentation */
int sgl_get_tile_list_sizes(uint32_t *out, int capacity, int *tiles_x_out, int *tiles_y_out);
/* enable != 0: the visibility kernel of later colour passes records per-tile start/end times (globaltimer ns);
 * out (may be NULL) receives [tiles][2] of the most recent such pass */
int sgl_debug_tile_times(int enable, unsigned long long *out, int capacity_tiles);
/* testing aid: cap the tile-bin region (entries) and the minimum clip arenas (vertices, fan triangles) of later passes so
 * that the bin-spill and clip-overflow paths can be exercised with small inputs; 0 restores a default */
int sgl_debug_set_limits(long long bin_capacity, long long clip_min_vertices, long long clip_min_fans);
/* instrumentation build only (-DSGL_TOUCH_BITMAP): bytes (whole 32-byte DRAM sectors) of texture `handle` (0 = all textures)
 * that samplers read since the last reset -- the measured B_tex of the roofline; the product library returns SGL_ERR_STATE */
int sgl_debug_texel_touch(int handle, int reset, unsigned long long *bytes_out);
int sgl_set_profiling(int on);
int sgl_get_kernel_times(SglKernelTime *out, int capacity);   /* returns the number of entries written */

/* ---- shader reflection (ShaderSoft::getUniformsDesc / getDefines, e.g. PbrSoft.h:77-103) -------------- */
int sgl_shader_uniform_offset(int shader, const char *name);   /* byte offset or -1 */
int sgl_shader_sampler_slot(int shader, const char *name);     /* slot or -1 */
int sgl_shader_define_bit(int shader, const char *name);       /* bit index or -1 */
int sgl_shader_uniform_size(int shader);                       /* bytes of the block part of ShaderUniforms */
int sgl_shader_varying_floats(int shader);                     /* sizeof(ShaderVaryings)/4 */

/* ---- buffers: VertexArrayObjectSoft (VertexSoft.h:14-46) ------------------------------------------------ */
int sgl_buffer_create(size_t bytes, const void *host_data, int *handle_out);
int sgl_buffer_upload(int handle, size_t offset, size_t bytes, const void *host_data);
int sgl_buffer_destroy(int handle);

/* ---- textures: TextureSoft / ImageBufferSoft (TextureSoft.h:20-255), SamplerSoft mip generation ------- */
int sgl_texture_create(const SglTextureDesc *desc, int *handle_out);
int sgl_texture_destroy(int handle);
/* host data is row-major w x h of the level, 4 bytes per texel; converted to the texture's layout */
int sgl_texture_upload(int handle, int layer, int level, const void *host_data);
int sgl_texture_gen_mips(int handle);                                   /* SamplerSoft.h:90-110,241-252 */
/* kind 0: attachment (w*h*samples*4 bytes, [y][x][sample]); kind 1: resolved colour of an MS texture; kind 2: the level's
 * raw storage in the texture's own layout, padded to whole tiles exactly like TiledBuffer / MortonBuffer
 * (Base/Buffer.h:143-158,174-202; size = sgl_texture_device_ptr's bytes_out) */
int sgl_texture_readback(int handle, int layer, int level, int kind, void *host_out, size_t bytes);
/* pipelined form: queued behind all submitted work on a second stream; a later pass that overwrites the image waits for
 * the copy on the device, the host waits with sgl_readback_wait() (or sgl_wait_idle()).  host_out should be pinned host
 * memory, or device memory (also a peer GPU's, mapped with sgl_peer_open). */
int sgl_texture_readback_async(int handle, int layer, int level, int kind, void *host_out, size_t bytes);
int sgl_readback_wait(void);
int sgl_texture_level_size(int handle, int level, int *w_out, int *h_out);
int sgl_texture_device_ptr(int handle, int layer, int level, int kind, void **ptr_out, size_t *bytes_out);

/* ---- render pass: beginRenderPass / setViewPort / draw / endRenderPass (RendererSoft.cpp:61-168) ------ */
int sgl_pass_begin(int color_tex, int color_layer, int color_level, int depth_tex,
                   int clear_color_flag, int clear_depth_flag, const float clear_color[4], float clear_depth);
int sgl_set_viewport(int x, int y, int width, int height);
int sgl_draw(const SglDraw *draw);
int sgl_pass_end(void);

/* ---- multi-GPU: sort-first screen-tile ownership --------------------------------- */
/* tile (tx,ty) of SGL_TILE x SGL_TILE pixels is rendered iff ((tx / band) + (ty / band) * k) % world == rank, where
 * the map is supplied explicitly: owner[ty*tiles_x+tx] == rank.  NULL restores "own everything". */
int sgl_set_tile_owner_map(const uint8_t *owner, int tiles_x, int tiles_y);
int sgl_tile_size(void);
/* Passes into `texture` (as colour or depth attachment) also render the tiles within `pixels` of an owned tile; pixels < 0:
 * every tile (the texture is replicated on all ranks).  For attachments that LATER passes sample: a screen-space filter
 * needs the owner's neighbourhood (FXAA reads up to 18.5 px along the edge + its bilinear footprint,
 * Viewer/Shader/Software/FxaaSoft.h:73-74,169-210: 32 px = two tiles), arbitrary look-ups need the whole image.  Attachments
 * of a different size than the owner map's frame (shadow maps) are always rendered whole. */
int sgl_texture_set_shard_halo(int texture, int pixels);
int sgl_set_rank(int rank, int world);                     /* re-label the context (sgl_init's rank/world) */
int sgl_tiles_owned(int owner_rank, int *tiles_out);       /* number of tiles `owner_rank` owns in the current map */
/* Gather of finished tiles to rank 0, NCCL form: tiles of a linear RGBA8 texture (resolved colour for MS targets) owned by
 * `owner_rank`, in tile-index order <-> dense device staging buffer [tiles][SGL_TILE][SGL_TILE] RGBA8 that the caller
 * moves with ncclSend/Recv (torch.distributed); rank 0 unpacks every peer's buffer into its own image.  Queued on the
 * library's stream. */
int sgl_tiles_pack(int texture, int owner_rank, void *dst_device, size_t dst_bytes, int *tiles_out);
int sgl_tiles_unpack(int texture, int owner_rank, const void *src_device, size_t src_bytes);
/* Gather by direct store over NVLink (no copy step): the final colour of every pass that renders into `texture`
 * (the MSAA resolve, RendererSoft.cpp:880-912, or the colour of a 1-sample target) is ALSO stored to
 * device_ptr[y*width+x]; device_ptr may be another GPU's memory mapped with sgl_peer_open.  NULL switches it off. */
int sgl_texture_set_mirror(int texture, void *device_ptr);
/* peer memory between the one-process-per-GPU ranks of a box: cudaMalloc + CUDA IPC handle (64 bytes, sent to the peers by
 * the caller), mapping, and stream-ordered flags: signal = system-scope release store of `value` after all prior work on
 * the library's stream; wait = bounded spin until `count` flags (64-byte stride) are >= value (wrap-safe compare). */
int sgl_peer_alloc(size_t bytes, void **ptr_out, uint8_t ipc_handle_out[64]);
int sgl_peer_free(void *ptr);
int sgl_peer_open(const uint8_t ipc_handle[64], void **ptr_out);
int sgl_peer_close(void *ptr);
int sgl_peer_signal(void *flag_device_ptr, uint32_t value);
int sgl_peer_signal_after_copies(void *flag_device_ptr, uint32_t value);   /* same, on the copy stream: after all queued
                                                                             sgl_texture_readback_async copies (whose destination
                                                                             may be a peer mapping: copy-engine gather) */
int sgl_peer_wait(const void *flags_device_ptr, int count, uint32_t value, int timeout_ms);
/* rank 0's per-frame bookkeeping in one launch: wait until `count` done-flags (64-byte stride) are >= value, then store
 * value into consumed_flags[1..count) (device pointers, usually peer mappings; entry 0 is ignored).  side_stream != 0
 * queues it on an internal stream of its own, so that rank 0's rendering is not serialised behind the other ranks. */
int sgl_peer_collect(const void *done_flags, int count, uint32_t value, void *const *consumed_flags, int timeout_ms, int side_stream);
int sgl_peer_timeouts(uint64_t *count_out);                /* number of waits that gave up (must stay 0) */

/* ---- unit-level entry points used by the known-answer tests (each wraps the device function the pipeline uses) */
/* barycentric + coverage + depth of RendererSoft::barycentric / rasterizationPixelQuad (RendererSoft.cpp:771-804,1021-1056)
 * for one triangle (3 x float4 screen positions) over n sample positions; out: bc[3], inside flag, z, 1/w per sample */
int sgl_kat_barycentric(const float *tri_xyzw, const float *sample_xy, int n, float *bc_out, int *inside_out,
                        float *zw_out);
/* BaseSampler::textureImpl (SamplerSoft.h:118-168) on a bound texture for n coordinates (2D: uv, cube: xyz); lod and
 * offsets_xy (2 ints per coordinate, texture2DLodOffset) may be NULL.  split_phase != 0 evaluates the split-phase bilinear
 * taps the straight-line shader paths use for "simple" samplers (RGBA8, linear layout, LINEAR filter, REPEAT or
 * CLAMP_TO_EDGE; SGL_ERR_INVALID otherwise) instead of the general function -- both must give the same bits. */
int sgl_kat_sample(int texture, int filter_min, int wrap, int border, const float *coords, const float *lod,
                   const int32_t *offsets_xy, int n, int split_phase, uint32_t *rgba_or_float_bits_out);
/* calcBlendColor (BlendSoft.h:44-56) and DepthTest (DepthSoft.h:13-25) tables */
int sgl_kat_blend(const SglRenderStates *states, const float *src_rgba, const float *dst_rgba, int n, float *out_rgba);
int sgl_kat_depth(int func, const float *a, const float *b, int n, int *pass_out);

#ifdef __cplusplus
}
#endif
#endif /* SGLCUDA_H_ */
