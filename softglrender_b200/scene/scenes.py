"""The BASELINE.json workloads as Renderer-API traces (SURVEY.md section 8d "Concrete inputs").

Each builder returns a TraceWriter whose stream is:
  resource creation + one warm-up frame, FRAME_BEGIN, one steady-state frame, FRAME_END, READBACKs.
Players execute the stream once; benchmark drivers re-execute the FRAME section per timed step.
"""
import os
import numpy as np

from . import trace as T
from . import assets as A
from .viewer import Config, Viewer, Scene, load_skybox, AA_NONE, AA_MSAA, AA_FXAA


def _finish(w, viewer, scene, frames=1):
    viewer.draw_frame(scene)              # creates resources, IBL etc.
    w.frame_begin()
    for _ in range(frames):
        viewer.draw_frame(scene)
    w.frame_end()
    viewer.readback_all()
    return w


def _model_path(assets_dir, name):
    import json
    with open(os.path.join(assets_dir, "assets.json")) as f:
        idx = json.load(f)
    return os.path.join(assets_dir, idx["model"][name]["path"])


def _skybox_path(assets_dir, name):
    import json
    with open(os.path.join(assets_dir, "assets.json")) as f:
        idx = json.load(f)
    return os.path.join(assets_dir, idx["skybox"][name]["path"])


def build_viewer_scene(assets_dir, model_name, width, height, config, skybox_name=None,
                       force_blinnphong=False, ibl_files=None, ibl_store=None, eye=(-1.5, 3, 3),
                       center=(0, 1, 0), model=None):
    w = T.TraceWriter()
    cache = {}
    if model is None:
        model = A.load_model(_model_path(assets_dir, model_name), cache)
    if force_blinnphong:
        def walk(n):
            for m in n.meshes:
                m.shading = "blinnphong"
            for c in n.children:
                walk(c)
        walk(model.root)
    sky = load_skybox(_skybox_path(assets_dir, skybox_name), cache) if skybox_name else None
    scene = Scene(config, model, sky)
    viewer = Viewer(w, config, width, height, eye=eye, center=center)
    viewer.ibl_files = ibl_files
    viewer.ibl_store = ibl_store
    return w, viewer, scene


def config1_cube(assets_dir, width=1000, height=800, **cfg):
    """C1: Cube.gltf forced to Blinn-Phong, 1000x800, no AA, non-reversed Z, Config defaults."""
    config = Config(**cfg)
    w, v, s = build_viewer_scene(assets_dir, "Cube", width, height, config, force_blinnphong=True)
    return _finish(w, v, s)


def config2_helmet(assets_dir, width=1920, height=1080, ibl_files=None, ibl_store=None, skybox="Room",
                   aa=AA_MSAA, model="DamagedHelmet", **cfg):
    """C2: DamagedHelmet PBR+IBL, equirect skybox (Room.jpeg), 1920x1080 MSAA4x, reversed-Z."""
    base = dict(show_skybox=True, pbr_ibl=True, reverse_z=True, aa_type=aa)
    base.update(cfg)
    config = Config(**base)
    w, v, s = build_viewer_scene(assets_dir, model, width, height, config, skybox_name=skybox,
                                 ibl_files=ibl_files, ibl_store=ibl_store)
    return _finish(w, v, s)


def compose_models(models, offsets):
    """Harness-composed Model with one child node per input model (SURVEY 8c gotcha 4)."""
    out = A.Model()
    for m, off in zip(models, offsets):
        n = A.Node()
        t = np.eye(4, dtype=np.float32)
        t[:3, 3] = off
        n.transform = (t @ m.centered).astype(np.float32)
        n.children.append(m.root)
        out.root.children.append(n)
        out.tri_count += m.tri_count
        out.vertex_count += m.vertex_count
    out.centered = np.eye(4, dtype=np.float32)
    return out


def config3_boombox_table(assets_dir, width=3840, height=2160, **cfg):
    """C3: BoomBox + GlassTable, shadow mapping, alpha blending, FXAA pass."""
    base = dict(aa_type=AA_FXAA)
    base.update(cfg)
    config = Config(**base)
    cache = {}
    bb = A.load_model(_model_path(assets_dir, "BoomBox"), cache)
    gt = A.load_model(_model_path(assets_dir, "GlassTable"), cache)
    model = compose_models([bb, gt], [(-0.9, 0.0, 0.0), (0.9, 0.0, 0.0)])
    w, v, s = build_viewer_scene(assets_dir, None, width, height, config, model=model)
    return _finish(w, v, s)


def simple_model(assets_dir, name, width, height, **cfg):
    config = Config(**cfg)
    w, v, s = build_viewer_scene(assets_dir, name, width, height, config)
    return _finish(w, v, s)


def fibonacci_eye(v, n_total=4096, radius=3.9, center=(0.0, 1.0, 0.0)):
    """Eye of view v of n_total on a Fibonacci sphere around `center`."""
    golden = np.pi * (3.0 - np.sqrt(5.0))
    y = 1.0 - 2.0 * (v + 0.5) / n_total
    r = np.sqrt(max(0.0, 1.0 - y * y))
    th = golden * v
    return (center[0] + radius * r * np.cos(th), center[1] + radius * y, center[2] + radius * r * np.sin(th))


def config5_views(assets_dir, name, views, n_total=4096, width=512, height=512, **cfg):
    """C5: a batch of views of one model (AfricanHead: views 0-2047, Robot: 2048-4095 of 4096), 512x512, no AA.

    FRAME section = every view of `views` back to back (the throughput unit of the render farm); the tail renders each
    view again and reads it back as color_v<k> / depth_v<k> (parity)."""
    config = Config(**cfg)
    w, viewer, scene = build_viewer_scene(assets_dir, name, width, height, config, eye=fibonacci_eye(views[0], n_total))
    viewer.draw_frame(scene)
    w.frame_begin()
    for v in views:
        viewer.cam_main.look_at(fibonacci_eye(v, n_total), (0, 1, 0), (0, 1, 0))
        viewer.draw_frame(scene)
    w.frame_end()
    for k, v in enumerate(views):
        viewer.cam_main.look_at(fibonacci_eye(v, n_total), (0, 1, 0), (0, 1, 0))
        viewer.draw_frame(scene)
        w.wait_idle()
        w.readback(viewer.tex_color_main, "color_v%d" % k)
        w.readback(viewer.tex_depth_main, "depth_v%d" % k)
    return w
