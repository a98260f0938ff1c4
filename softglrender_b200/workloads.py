"""Preparation of the BASELINE.json workloads for tests and bench: traces, IBL maps, player invocations.

Nothing here reads /root/reference: assets come from <repo>/assets (copied there by ``__graft_entry__.build()``).
"""
import hashlib
import json
import os
import subprocess

from . import LIB_DIR, REPO_ROOT
from .scene import assets as A
from .scene import scenes

IBL_FILES = dict(cube="ibl_cube.tex", irradiance="ibl_irr.tex", prefilter="ibl_pre.tex")   # plain names (IBL-generation tests)

CUDA_PLAYER = os.path.join(LIB_DIR, "sgl_player")
REF_PLAYER = os.path.join(REPO_ROOT, "oracle", "_ref", "ref_player")          # reference, its own thread pool
REF_PLAYER_ST = os.path.join(REPO_ROOT, "oracle", "_ref", "ref_player_st")    # reference, deterministic order
ORACLE_PLAYER = os.path.join(REPO_ROOT, "oracle", "_build", "oracle_player")  # CPU restatement


_SCENE_DIGEST = None


def cache_key(*parts):
    """Short digest of the builder arguments AND of the scene-module sources, so that a cached trace / IBL map is never
    replayed after the code or the arguments that produced it changed."""
    global _SCENE_DIGEST
    if _SCENE_DIGEST is None:
        h = hashlib.sha1()
        d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scene")
        for name in sorted(os.listdir(d)):
            if name.endswith(".py"):
                with open(os.path.join(d, name), "rb") as f:
                    h.update(f.read())
        _SCENE_DIGEST = h.hexdigest()
    return hashlib.sha1((repr(parts) + _SCENE_DIGEST).encode()).hexdigest()[:10]


def _atomic_save(writer, path):
    tmp = path + ".tmp%d" % os.getpid()
    writer.save(tmp)
    os.replace(tmp, path)


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


def ibl_file_names(skybox, player=None):
    """IBL cache files of one (skybox, generating player) pair inside a work dir."""
    tag = "%s_%s_%s" % (skybox, os.path.basename(player or CUDA_PLAYER), cache_key("ibl", skybox))
    return dict(cube="ibl_%s_cube.tex" % tag, irradiance="ibl_%s_irr.tex" % tag, prefilter="ibl_%s_pre.tex" % tag)


def build_ibl(work_dir, skybox="Room", ibl_player=None):
    """Runs the IBL-generation passes (equirect -> cube, irradiance, prefilter; Environment.cpp:25-195) of `skybox` on
    `ibl_player` (a player binary; default: the RendererCUDA player) once per work_dir.  Returns the file-name dict."""
    os.makedirs(work_dir, exist_ok=True)
    files = ibl_file_names(skybox, ibl_player)
    if not all(os.path.exists(os.path.join(work_dir, f)) for f in files.values()):
        gen = os.path.join(work_dir, "iblgen_%s_%s.sglt" % (skybox, cache_key("iblgen", skybox, sorted(files.items()))))
        _atomic_save(scenes.config2_helmet(assets_dir(), width=64, height=64, skybox=skybox, model="Cube", ibl_store=files,
                                           shadow_map=False), gen)
        run_player(ibl_player or CUDA_PLAYER, gen, data_dir=work_dir)
        os.remove(gen)
    return files


def build_c2(work_dir, width=1920, height=1080, skybox="Room", model="DamagedHelmet", ibl_player=None, **cfg):
    """C2 = DamagedHelmet PBR+IBL, equirect skybox, MSAA4x, reversed-Z.  Returns (trace_path, data_dir).

    The IBL maps (converted cube, irradiance, prefilter) are produced once per work_dir by `ibl_player` (build_ibl) and are
    then loaded by every renderer under comparison, so all of them shade with identical maps.  The parity tests pass the
    compiled reference as `ibl_player`; tests/test_parity_gpu.py::test_equirect_to_cube_and_ibl_match_reference compares
    the maps the two renderers generate."""
    files = build_ibl(work_dir, skybox, ibl_player)
    trace = os.path.join(work_dir, "c2_%s_%dx%d_%s.sglt" % (model, width, height,
                                                            cache_key("c2", width, height, skybox, model, sorted(files.items()), sorted(cfg.items()))))
    if not os.path.exists(trace):
        _atomic_save(scenes.config2_helmet(assets_dir(), width=width, height=height, skybox=skybox, model=model, ibl_files=files,
                                           **cfg), trace)
    return trace, work_dir


def build_c1(work_dir, width=1000, height=800, **cfg):
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c1_%dx%d_%s.sglt" % (width, height, cache_key("c1", width, height, sorted(cfg.items()))))
    if not os.path.exists(trace):
        _atomic_save(scenes.config1_cube(assets_dir(), width, height, **cfg), trace)
    return trace, work_dir


def build_c3(work_dir, width=3840, height=2160, **cfg):
    """C3 = BoomBox + GlassTable, shadow mapping, alpha blending (GlassTable glass), FXAA pass."""
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c3_%dx%d_%s.sglt" % (width, height, cache_key("c3", width, height, sorted(cfg.items()))))
    if not os.path.exists(trace):
        _atomic_save(scenes.config3_boombox_table(assets_dir(), width, height, **cfg), trace)
    return trace, work_dir


def build_c4(work_dir, n_tris=100000, width=1920, height=1080, tex_size=1024, **kw):
    """C4 = synthetic triangle soup (no assets needed): mixed sizes, 8 mip-mapped REPEAT textures, Blinn-Phong."""
    from .scene import synth
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c4_%d_%dx%d_t%d_%s.sglt" % (n_tris, width, height, tex_size,
                                                                 cache_key("c4", n_tris, width, height, tex_size, sorted(kw.items()))))
    if not os.path.exists(trace):
        _atomic_save(synth.soup_trace(n_tris, width, height, tex_size=tex_size, **kw), trace)
    return trace, work_dir


def build_c5(work_dir, model, views, n_total=4096, width=512, height=512, **cfg):
    """C5 = multi-view batch: `views` (indices into the n_total-view Fibonacci sphere) of AfricanHead or Robot."""
    os.makedirs(work_dir, exist_ok=True)
    trace = os.path.join(work_dir, "c5_%s_%dx%d_%s.sglt" % (model, width, height,
                                                             cache_key("c5", model, list(views), n_total, width, height, sorted(cfg.items()))))
    if not os.path.exists(trace):
        _atomic_save(scenes.config5_views(assets_dir(), model, list(views), n_total, width, height, **cfg), trace)
    return trace, work_dir
