"""GPU part of the integration proof: the reference's own Viewer (unmodified Viewer.cpp / ModelLoader.cpp / assimp, see
integration/Makefile) drives RendererCUDA through ViewerCUDA and RendererSoft through ViewerSoftware's role, same scene, same
camera, same process image -- the frames must meet the parity bar (depth bit-exact, colour within 1/255 on >= 99.9 %), and the
submission the Viewer makes (passes, draws, vertices, indices per frame) equals what the Python scene builder's trace submits."""
import os

import numpy as np
import pytest

from conftest import compare_outputs
from test_viewer_integration import need_viewer, run_viewer

pytestmark = pytest.mark.gpu


def _both(work, tag, *args):
    from softglrender_b200.scene.trace import read_outputs
    out_s, out_c = os.path.join(work, tag + "_soft.out"), os.path.join(work, tag + "_cuda.out")
    run_viewer(work, "soft", out_s, *args)
    info = run_viewer(work, "cuda", out_c, *args, "--frames", 2)
    return read_outputs(out_s), read_outputs(out_c), info


def test_config1_through_the_reference_viewer(work_dir):
    """BASELINE config 1: Cube forced to Blinn-Phong, 1000x800, no AA, shadow pass + main pass."""
    need_viewer()
    work = os.path.join(work_dir, "integ")
    os.makedirs(work, exist_ok=True)
    soft, cuda, info = _both(work, "c1", "--model", "Cube", "--blinnphong", "--width", 1000, "--height", 800)
    rep = compare_outputs(soft, cuda)
    print("config 1 through the reference Viewer:", rep, info)
    assert soft["color"].std() > 5.0
    assert info["last_frame"]["kernel_launches"] > 0


def test_config2_through_the_reference_viewer_and_its_submission_matches_the_trace(work_dir):
    """BASELINE config 2 (DamagedHelmet PBR + IBL generated from Room.jpeg by each backend's own IBLGenerator passes, equirect
    skybox, MSAA 4x, reversed-Z) at 960x540: frames at the parity bar (the soft side runs its real thread pool, so colour is
    compared at 99.9 %, depth bit for bit), and one Viewer frame submits exactly what one frame of scene/viewer.py's trace does."""
    from softglrender_b200 import capi, workloads
    need_viewer()
    work = os.path.join(work_dir, "integ")
    os.makedirs(work, exist_ok=True)
    args = ("--model", "DamagedHelmet", "--skybox", "Room", "--ibl", "--aa", "msaa", "--reverse-z", "--width", 960, "--height", 540)
    soft, cuda, info = _both(work, "c2", *args)
    rep = compare_outputs({k: soft[k] for k in ("color", "depth.ms", "shadow")}, cuda)
    print("config 2 through the reference Viewer:", rep, info)
    # the Python scene builder's steady-state frame of the same configuration
    trace, data = workloads.build_c2(os.path.join(work_dir, "c2"), 960, 540)
    capi.init(0)
    lib = capi.load()
    p = capi.Player(trace, data)
    try:
        p.setup()
        p.frame(sync=True)
        capi.check(lib.sgl_reset_counters())
        p.frame(sync=True)
        c = capi.counters()
    finally:
        p.close()
    want = info["last_frame"]
    got = {k: c[k] for k in ("passes", "draws", "vertices_in", "indices_in", "primitives_in")}
    assert got == {k: want[k] for k in got}, (got, want)
