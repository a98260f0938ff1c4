"""GPU: capacity limits of the binning / clipping arenas never change a frame silently.

* tile bins exhausted -> primitives spill to the pass-wide list: the frame stays bit-identical to the reference, only
  SglCounters.bin_spills tells;
* clip-vertex / fan arena exhausted -> primitives would be dropped: the next synchronising call returns SGL_ERR_OVERFLOW,
  later passes get a larger arena and the resubmitted frame matches the reference;
* depth-only (shadow) work queue exhausted -> triangles go to the tile-parallel kernel, same depth map."""
import os

import numpy as np
import pytest

from conftest import compare_outputs

pytestmark = pytest.mark.gpu

SGL_ERR_OVERFLOW = -6


def _frame_outputs(trace, data, out):
    from softglrender_b200 import capi
    from softglrender_b200.scene.trace import read_outputs
    p = capi.Player(trace, data)
    p.setup()
    p.frame(sync=False)
    rc = capi.load().sgl_wait_idle()
    res = None
    if rc == 0:
        p.tail(out)
        res = read_outputs(out)
    return p, rc, res


@pytest.fixture()
def limits():
    from softglrender_b200 import capi
    capi.init(0)
    lib = capi.load()
    yield lambda b=0, v=0, f=0: capi.check(lib.sgl_debug_set_limits(b, v, f))
    lib.sgl_wait_idle()
    capi.check(lib.sgl_debug_set_limits(0, 0, 0))
    lib.sgl_reset_counters()      # the overflow counts provoked here must not leak into other tests' assertions


@pytest.mark.parametrize("msaa", [False, True])
def test_full_bins_spill_without_changing_the_frame(msaa, limits, checker_player, work_dir):
    from softglrender_b200 import capi, workloads
    from softglrender_b200.scene import synth
    from softglrender_b200.scene.trace import read_outputs
    trace = os.path.join(work_dir, "stress_mid_%d.sglt" % msaa)
    synth.stress_trace("mid", n_tris=600, msaa=msaa).save(trace)
    ref_out = os.path.join(work_dir, "stress_mid.ref.out")
    workloads.run_player(checker_player, trace, out=ref_out, data_dir=work_dir)
    ref = read_outputs(ref_out)
    lib = capi.load()
    capi.check(lib.sgl_reset_counters())
    p, rc, base = _frame_outputs(trace, work_dir, os.path.join(work_dir, "stress_mid.a.out"))
    assert rc == 0
    c = capi.counters()
    assert c["bin_spills"] == 0 and c["primitives_binned"] > 20 * 400      # 20+ tiles per visible triangle on average
    p.close()
    compare_outputs(ref, base)
    limits(4096)                                                            # room for ~100 of the 600 triangles
    capi.check(lib.sgl_reset_counters())
    p, rc, small = _frame_outputs(trace, work_dir, os.path.join(work_dir, "stress_mid.b.out"))
    assert rc == 0, capi.load().sgl_last_error()
    c = capi.counters()
    assert c["bin_spills"] > 100 and c["clip_overflow"] == 0
    p.close()
    for k in base:
        assert np.array_equal(base[k], small[k]), k
    os.remove(trace)


def test_clip_arena_overflow_is_reported_and_the_retry_is_right(limits, checker_player, work_dir):
    from softglrender_b200 import capi, workloads
    from softglrender_b200.scene import synth
    from softglrender_b200.scene.trace import read_outputs
    trace = os.path.join(work_dir, "stress_clipped.sglt")
    synth.stress_trace("clipped", width=320, height=240, n_tris=480).save(trace)
    ref_out = os.path.join(work_dir, "stress_clipped.ref.out")
    workloads.run_player(checker_player, trace, out=ref_out, data_dir=work_dir)
    ref = read_outputs(ref_out)
    lib = capi.load()
    # default arenas hold the worst case of a draw this small: no overflow, frame right first time
    p, rc, got = _frame_outputs(trace, work_dir, os.path.join(work_dir, "stress_clipped.a.out"))
    assert rc == 0
    p.close()
    compare_outputs(ref, got)
    # tiny arenas: the first submissions overflow LOUDLY, each report doubles the arena, then the frame is right
    limits(0, 16, 4)
    capi.check(lib.sgl_reset_counters())
    rcs = []
    got = None
    for attempt in range(8):
        p, rc, got = _frame_outputs(trace, work_dir, os.path.join(work_dir, "stress_clipped.b.out"))
        rcs.append(rc)
        p.close()
        if rc == 0:
            break
        assert rc == SGL_ERR_OVERFLOW and b"clip" in lib.sgl_last_error()
    assert rcs[0] == SGL_ERR_OVERFLOW and rcs[-1] == 0, rcs
    compare_outputs(ref, got)
    os.remove(trace)


def test_depth_only_queue_overflow_falls_back_to_the_tile_kernel(limits, work_dir):
    """Shadow pass of config 1 with a work queue far too small: same shadow map, same frame."""
    from softglrender_b200 import capi, workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    trace, data = workloads.build_c1(os.path.join(work_dir, "c1ovf"), 400, 320)
    p, rc, base = _frame_outputs(trace, data, os.path.join(work_dir, "c1ovf.a.out"))
    assert rc == 0
    p.close()
    limits(64)    # the depth-only path carves its queue out of the bin region
    p, rc, small = _frame_outputs(trace, data, os.path.join(work_dir, "c1ovf.b.out"))
    assert rc == 0, capi.load().sgl_last_error()
    p.close()
    for k in base:
        assert np.array_equal(base[k], small[k]), k
