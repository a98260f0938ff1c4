"""ctypes bindings: the C ABI of libsglcuda.so (include/sglcuda.h) and the harness entry points of libsglhost.so."""
import ctypes as C
import os

import numpy as np

from . import LIB_DIR

SGL_MAX_SAMPLER_SLOTS = 8
SGL_MAX_UNIFORM_BYTES = 512


class SglTextureDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("type", C.c_int32), ("format", C.c_int32),
                ("use_mipmaps", C.c_int32), ("multi_sample", C.c_int32), ("layout", C.c_int32)]


class SglRenderStates(C.Structure):
    _fields_ = [("blend", C.c_int32), ("blend_func_rgb", C.c_int32), ("blend_src_rgb", C.c_int32),
                ("blend_dst_rgb", C.c_int32), ("blend_func_alpha", C.c_int32), ("blend_src_alpha", C.c_int32),
                ("blend_dst_alpha", C.c_int32), ("depth_test", C.c_int32), ("depth_mask", C.c_int32),
                ("depth_func", C.c_int32), ("cull_face", C.c_int32), ("primitive_type", C.c_int32),
                ("polygon_mode", C.c_int32), ("line_width", C.c_float)]


class SglSamplerBinding(C.Structure):
    _fields_ = [("texture", C.c_int32), ("filter_min", C.c_int32), ("wrap", C.c_int32), ("border", C.c_int32)]


class SglDraw(C.Structure):
    _fields_ = [("vertex_buffer", C.c_int32), ("index_buffer", C.c_int32), ("vertex_count", C.c_int32),
                ("index_count", C.c_int32), ("shader", C.c_int32), ("defines", C.c_uint32),
                ("states", SglRenderStates), ("uniform_bytes", C.c_uint32),
                ("uniforms", C.c_uint8 * SGL_MAX_UNIFORM_BYTES),
                ("samplers", SglSamplerBinding * SGL_MAX_SAMPLER_SLOTS)]


class SglCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("passes", "draws", "primitives_in", "primitives_binned",
                                          "fragments_shaded", "samples_written", "kernel_launches", "clip_overflow",
                                          "h2d_bytes", "d2h_bytes", "host_ns_pass_end", "host_ns_draw", "bin_spills", "vertices_in", "indices_in", "host_ns_wait_gpu", "early_vis", "renamed_passes")]


class SglKernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint64), ("total_ms", C.c_double)]


# every symbol include/sglcuda.h declares (tests check that the library exports each one)
C_ABI_SYMBOLS = [
    "sgl_init", "sgl_shutdown", "sgl_last_error", "sgl_set_stream", "sgl_wait_idle", "sgl_get_counters",
    "sgl_reset_counters", "sgl_timer_begin", "sgl_timer_end", "sgl_set_profiling", "sgl_get_kernel_times",
    "sgl_get_tile_list_sizes", "sgl_debug_tile_times", "sgl_debug_set_limits", "sgl_debug_texel_touch",
    "sgl_shader_uniform_offset", "sgl_shader_sampler_slot",
    "sgl_shader_define_bit", "sgl_shader_uniform_size", "sgl_shader_varying_floats", "sgl_buffer_create",
    "sgl_buffer_upload", "sgl_buffer_destroy", "sgl_texture_create", "sgl_texture_destroy", "sgl_texture_upload",
    "sgl_texture_gen_mips", "sgl_texture_readback", "sgl_texture_readback_async", "sgl_readback_wait", "sgl_texture_level_size", "sgl_texture_device_ptr",
    "sgl_pass_begin", "sgl_set_viewport", "sgl_draw", "sgl_pass_end", "sgl_set_tile_owner_map", "sgl_tile_size",
    "sgl_set_rank", "sgl_texture_set_shard_halo", "sgl_tiles_owned", "sgl_tiles_pack", "sgl_tiles_unpack", "sgl_texture_set_mirror", "sgl_peer_alloc",
    "sgl_peer_free", "sgl_peer_open", "sgl_peer_close", "sgl_peer_signal", "sgl_peer_signal_after_copies", "sgl_peer_wait", "sgl_peer_collect",
    "sgl_peer_timeouts",
    "sgl_kat_barycentric", "sgl_kat_sample", "sgl_kat_blend", "sgl_kat_depth"]

_lib = None
_host = None


def lib_path():
    return os.path.join(LIB_DIR, "libsglcuda.so")


def load():
    """Load libsglcuda.so.  Raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise RuntimeError("libsglcuda.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "RendererCUDA has no CPU fallback" % p)
    _lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    _lib.sgl_last_error.restype = C.c_char_p
    _lib.sgl_timer_end.argtypes = [C.POINTER(C.c_float)]
    _lib.sgl_set_stream.argtypes = [C.c_void_p]
    _lib.sgl_buffer_create.argtypes = [C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]
    _lib.sgl_buffer_upload.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]
    _lib.sgl_texture_create.argtypes = [C.POINTER(SglTextureDesc), C.POINTER(C.c_int)]
    _lib.sgl_texture_upload.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    _lib.sgl_texture_readback.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    _lib.sgl_texture_readback_async.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    _lib.sgl_texture_level_size.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib.sgl_pass_begin.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_float]
    _lib.sgl_draw.argtypes = [C.POINTER(SglDraw)]
    _lib.sgl_get_counters.argtypes = [C.POINTER(SglCounters)]
    _lib.sgl_set_tile_owner_map.argtypes = [C.c_void_p, C.c_int, C.c_int]
    _lib.sgl_tiles_owned.argtypes = [C.c_int, C.POINTER(C.c_int)]
    _lib.sgl_tiles_pack.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
    _lib.sgl_tiles_unpack.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    _lib.sgl_texture_set_mirror.argtypes = [C.c_int, C.c_void_p]
    _lib.sgl_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    _lib.sgl_peer_free.argtypes = [C.c_void_p]
    _lib.sgl_peer_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    _lib.sgl_peer_close.argtypes = [C.c_void_p]
    _lib.sgl_peer_signal.argtypes = [C.c_void_p, C.c_uint32]
    _lib.sgl_peer_signal_after_copies.argtypes = [C.c_void_p, C.c_uint32]
    _lib.sgl_peer_collect.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
    _lib.sgl_peer_wait.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int]
    _lib.sgl_peer_timeouts.argtypes = [C.POINTER(C.c_uint64)]
    _lib.sgl_texture_device_ptr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    _lib.sgl_shader_uniform_offset.argtypes = [C.c_int, C.c_char_p]
    _lib.sgl_shader_sampler_slot.argtypes = [C.c_int, C.c_char_p]
    _lib.sgl_shader_define_bit.argtypes = [C.c_int, C.c_char_p]
    _lib.sgl_kat_barycentric.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib.sgl_kat_sample.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    _lib.sgl_kat_blend.argtypes = [C.POINTER(SglRenderStates), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    _lib.sgl_kat_depth.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    _lib.sgl_get_tile_list_sizes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib.sgl_debug_tile_times.argtypes = [C.c_int, C.c_void_p, C.c_int]
    _lib.sgl_debug_texel_touch.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    _lib.sgl_debug_set_limits.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong]
    _lib.sgl_get_kernel_times.argtypes = [C.POINTER(SglKernelTime), C.c_int]
    return _lib


def check(rc, what="sglcuda"):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, load().sgl_last_error().decode()))


def init(device=0, rank=0, world=1):
    check(load().sgl_init(device, rank, world), "sgl_init")


def counters():
    c = SglCounters()
    check(load().sgl_get_counters(C.byref(c)))
    return {n: int(getattr(c, n)) for n, _ in SglCounters._fields_}


def kernel_times():
    arr = (SglKernelTime * 32)()
    n = load().sgl_get_kernel_times(arr, 32)
    return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].total_ms)) for i in range(n)}


class Player:
    """In-process RendererCUDA trace player (libsglhost.so)."""

    def __init__(self, trace_path, data_dir="."):
        global _host
        load()
        if _host is None:
            p = os.path.join(LIB_DIR, "libsglhost.so")
            if not os.path.exists(p):
                raise RuntimeError("libsglhost.so is missing (%s): build the package first" % p)
            _host = C.CDLL(p, mode=C.RTLD_GLOBAL)
            _host.sglp_create.restype = C.c_void_p
            _host.sglp_destroy.argtypes = [C.c_void_p]
            _host.sglp_load.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
            _host.sglp_setup.argtypes = [C.c_void_p]
            _host.sglp_frame.argtypes = [C.c_void_p, C.c_int]
            _host.sglp_tail.argtypes = [C.c_void_p, C.c_char_p]
            _host.sglp_readback.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.h = _host
        self.p = self.h.sglp_create()
        if self.h.sglp_load(self.p, trace_path.encode(), data_dir.encode()) != 0:
            raise RuntimeError("cannot load trace %s" % trace_path)

    def setup(self):
        if self.h.sglp_setup(self.p) != 0:
            raise RuntimeError("trace setup failed: %s" % load().sgl_last_error().decode())

    def frame(self, sync=False):
        if self.h.sglp_frame(self.p, 1 if sync else 0) != 0:
            raise RuntimeError("trace frame failed: %s" % load().sgl_last_error().decode())

    def tail(self, out_path):
        if self.h.sglp_tail(self.p, out_path.encode()) != 0:
            raise RuntimeError("trace tail failed")

    def readback(self, tag, out=None):
        w, h, f, s = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        if out is None:
            n = self.h.sglp_readback(self.p, tag.encode(), None, 0, C.byref(w), C.byref(h), C.byref(f), C.byref(s))
            if n < 0:
                raise RuntimeError("unknown readback tag %s" % tag)
            out = np.empty(n, np.uint8)
        n = self.h.sglp_readback(self.p, tag.encode(), out.ctypes.data, out.nbytes, C.byref(w), C.byref(h),
                                 C.byref(f), C.byref(s))
        if n < 0:
            raise RuntimeError("readback of %s failed (%d)" % (tag, n))
        return out, (w.value, h.value, f.value, s.value)

    def texture_handle(self, tag):
        self.h.sglp_texture_handle.argtypes = [C.c_void_p, C.c_char_p]
        return int(self.h.sglp_texture_handle(self.p, tag.encode()))

    def close(self):
        if self.p:
            self.h.sglp_destroy(self.p)
            self.p = None
