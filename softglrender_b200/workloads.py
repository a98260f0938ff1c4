"""Preparation of the BASELINE.json workloads for tests and bench: traces, IBL maps, player invocations.

Nothing here reads /root/reference: assets come from <repo>/assets (copied there by ``__graft_entry__.build()``).
"""
import json
import os
import subprocess

from . import LIB_DIR, REPO_ROOT
from .scene import assets as A
from .scene import scenes

IBL_FILES = dict(cube="ibl_cube.tex", irradiance="ibl_irr.tex", prefilter="ibl_pre.tex")

CUDA_PLAYER = os.path.join(LIB_DIR, "sgl_player")
REF_PLAYER = os.path.join(REPO_ROOT, "oracle", "_ref", "ref_player")          # reference, its own thread pool
REF_PLAYER_ST = os.path.join(REPO_ROOT, "oracle", "_ref", "ref_player_st")    # reference, deterministic order
ORACLE_PLAYER = os.path.join(REPO_ROOT, "oracle", "_build", "oracle_player")  # CPU restatement


def assets_dir():
    d = A.find_assets_dir()
    if d is None:
        raise RuntimeError("assets/ not found next to the package (run __graft_entry__.build() where the reference "
                           "tree is available, or set SGL_ASSETS_DIR)")
    return d


def run_player(binary, trace, out=None, data_dir=".", frames=0, warmup=0, env=None, timeout=1800):
    if not os.path.exists(binary):
        raise RuntimeError("player binary missing: %s" % binary)
    cmd = [binary, trace, "--data-dir", data_dir]
    if out:
        cmd += ["--out", out]
    if frames:
        cmd += ["--frames", str(frames), "--warmup", str(warmup)]
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (" ".join(cmd), r.returncode, r.stderr[-2000:]))
    for line in r.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    return {}


def build_c2(work_dir, width=1920, height=1080, skybox="Room", model="DamagedHelmet", ibl_player=None, **cfg):
    """C2 = DamagedHelmet PBR+IBL, equirect skybox, MSAA4x, reversed-Z.  Returns (trace_path, data_dir).

    The IBL maps (converted cube, irradiance, prefilter) are produced once per work_dir by replaying the
    IBL-generation trace on ``ibl_player`` (a player binary; default: the RendererCUDA player) and are then
    loaded by every renderer under comparison, so all of them shade with identical maps.
    """
    os.makedirs(work_dir, exist_ok=True)
    ad = assets_dir()
    gen = os.path.join(work_dir, "iblgen_%s.sglt" % skybox)
    have = all(os.path.exists(os.path.join(work_dir, f)) for f in IBL_FILES.values())
    if not have:
        scenes.config2_helmet(ad, width=64, height=64, skybox=skybox, model="Cube", ibl_store=IBL_FILES,
                              shadow_map=False).save(gen)
        run_player(ibl_player or CUDA_PLAYER, gen, data_dir=work_dir)
    trace = os.path.join(work_dir, "c2_%s_%dx%d.sglt" % (model, width, height))
    if not os.path.exists(trace):
        scenes.config2_helmet(ad, width=width, height=height, skybox=skybox, model=model, ibl_files=IBL_FILES,
                              **cfg).save(trace)
    return trace, work_dir


def build_c1(work_dir, width=1000, height=800, **cfg):
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c1_%dx%d.sglt" % (width, height))
    if not os.path.exists(trace):
        scenes.config1_cube(assets_dir(), width, height, **cfg).save(trace)
    return trace, work_dir


def build_c3(work_dir, width=3840, height=2160, **cfg):
    """C3 = BoomBox + GlassTable, shadow mapping, alpha blending (GlassTable glass), FXAA pass."""
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c3_%dx%d.sglt" % (width, height))
    if not os.path.exists(trace):
        scenes.config3_boombox_table(assets_dir(), width, height, **cfg).save(trace)
    return trace, work_dir


def build_c4(work_dir, n_tris=100000, width=1920, height=1080, tex_size=1024, **kw):
    """C4 = synthetic triangle soup (no assets needed): mixed sizes, 8 mip-mapped REPEAT textures, Blinn-Phong."""
    from .scene import synth
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c4_%d_%dx%d_t%d.sglt" % (n_tris, width, height, tex_size))
    if not os.path.exists(trace):
        synth.soup_trace(n_tris, width, height, tex_size=tex_size, **kw).save(trace)
    return trace, work_dir


def build_c5(work_dir, model, views, n_total=4096, width=512, height=512, **cfg):
    """C5 = multi-view batch: `views` (indices into the n_total-view Fibonacci sphere) of AfricanHead or Robot."""
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c5_%s_%s_%dx%d.sglt" % (model, "-".join(str(v) for v in views[:6]) + ("+%d" % len(views)), width, height))
    if not os.path.exists(trace):
        scenes.config5_views(assets_dir(), model, list(views), n_total, width, height, **cfg).save(trace)
    return trace, work_dir
