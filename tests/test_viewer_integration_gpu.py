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


def _compare_with_racy_reference(work, tag, args, soft, cuda, keys=None):
    """The soft side of this harness is the reference with its REAL thread pool, and that renderer races with itself: blocks of
    different triangles test and write the same depth sample without ordering, so its own runs are not identical (measured here:
    4 of 60 runs of the config-2 frame differ from the first one in one depth or colour sample).  RendererCUDA is deterministic.
    So the bit-exact depth bar is kept, but a mismatch of a handful of samples gets the reference rendered again (twice at
    most): a defect on the CUDA side fails every time, a lost update inside the reference does not repeat."""
    from softglrender_b200.scene.trace import read_outputs
    last = None
    for attempt in range(3):
        ref = soft if keys is None else {k: soft[k] for k in keys}
        try:
            return compare_outputs(ref, cuda)
        except AssertionError as e:
            last = e
            racy = sum(int((soft[k].view(np.uint32) != cuda[k].view(np.uint32)).sum()) for k in ref if soft[k].dtype != np.uint8
                       and k in cuda and soft[k].shape == cuda[k].shape)
            if "depth samples differ" not in str(e) or racy > 8:
                raise
            out_s = os.path.join(work, "%s_soft_retry%d.out" % (tag, attempt))
            run_viewer(work, "soft", out_s, *args)
            soft = read_outputs(out_s)
            os.remove(out_s)
    raise last


def test_config1_through_the_reference_viewer(work_dir):
    """BASELINE config 1: Cube forced to Blinn-Phong, 1000x800, no AA, shadow pass + main pass."""
    need_viewer()
    work = os.path.join(work_dir, "integ")
    os.makedirs(work, exist_ok=True)
    args = ("--model", "Cube", "--blinnphong", "--width", 1000, "--height", 800)
    soft, cuda, info = _both(work, "c1", *args)
    rep = _compare_with_racy_reference(work, "c1", args, soft, cuda)
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
    rep = _compare_with_racy_reference(work, "c2", args, soft, cuda, keys=("color", "depth.ms", "shadow"))
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
