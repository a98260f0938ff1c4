"""Builds libsglcuda.so (sm_100a kernels + C ABI) and the RendererCUDA trace player in-tree.

nvcc cross-compiles without a GPU; the .so files are git-ignored but travel with the repo snapshot to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CU_SOURCES = ["sglcuda.cu", "sgl_raster_ns1.cu", "sgl_raster_ns4.cu", "sgl_vis.cu", "sgl_depth.cu", "sgl_shade_ns1.cu", "sgl_shade_ns4.cu"]
HEADERS = ["sgl_kernels.cuh", "sgl_vis.cuh", "sgl_depth.cuh", "sgl_pixel.h", "sgl_setup.h", "sgl_raster.h", "sgl_shaders.h", "sgl_texture.h",
           "sgl_math.h", "sgl_types.h", os.path.join(ROOT, "include", "sglcuda.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=True, out=None, extra_flags=()):
    """out / extra_flags: A/B variant of the same sources (e.g. -DSGL_VIS_MIN_BLOCKS=5) into another directory."""
    global OUT
    saved = OUT
    if out:
        OUT = out
    try:
        return _build(force, verbose, list(extra_flags))
    finally:
        OUT = saved


def _build(force, verbose, extra_flags):
    os.makedirs(OUT, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append((["nvcc"] + NVCC_FLAGS + extra_flags + ["-c", s, "-o", o], o + ".ptxas.log"))
    if jobs and verbose:
        print("[build] nvcc: %d translation unit(s) for sm_100a ..." % len(jobs), flush=True)
    with ThreadPoolExecutor(max_workers=len(jobs) or 1) as ex:
        list(ex.map(lambda j: _run(*j), jobs))
    lib = os.path.join(OUT, "libsglcuda.so")
    if force or jobs or not os.path.exists(lib):
        _run(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    # host side: RendererCUDA + harness, linked against the C ABI only
    host_srcs = [os.path.join(HERE, "host", "Render", "CUDA", "RendererCUDA.cpp"),
                 os.path.join(HERE, "harness", "cuda_backend.cpp"),
                 os.path.join(HERE, "harness", "trace_player.cpp")]
    host_hdrs = [os.path.join(HERE, "host", "Render", "CUDA", "RendererCUDA.h"),
                 os.path.join(HERE, "host", "Render", "RenderAPI.h"), os.path.join(HERE, "harness", "trace_player.h"),
                 os.path.join(HERE, "harness", "trace_format.h"), os.path.join(ROOT, "include", "sglcuda.h")]
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "host"), "-I" + os.path.join(HERE, "harness"),
           "-DPLAYER_BACKEND_HEADER=\"cuda_backend.h\""]
    common = ["g++", "-std=c++17", "-O2", "-fPIC"] + inc
    player = os.path.join(OUT, "sgl_player")
    main = os.path.join(HERE, "harness", "player_main.cpp")
    if force or _newer(player, host_srcs + host_hdrs + [main, lib]):
        _run(common + host_srcs + [main, "-o", player, "-L" + OUT, "-lsglcuda", "-Wl,-rpath,$ORIGIN"])
    hostlib = os.path.join(OUT, "libsglhost.so")
    capi = os.path.join(HERE, "harness", "player_capi.cpp")
    if os.path.exists(capi) and (force or _newer(hostlib, host_srcs + host_hdrs + [capi, lib])):
        _run(common + ["-shared"] + host_srcs + [capi, "-o", hostlib, "-L" + OUT, "-lsglcuda", "-Wl,-rpath,$ORIGIN"])
    if verbose:
        print("[build] ok:", lib, flush=True)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python -m softglrender_b200.build --variant NAME -DFOO=1 ...
        i = sys.argv.index("--variant")
        build(force=True, out=os.path.join(HERE, "lib_variants", sys.argv[i + 1]), extra_flags=sys.argv[i + 2:])
    else:
        build(force="--force" in sys.argv)
