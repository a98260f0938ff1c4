"""The multi-stream execution model must be invisible: geometry / pixel / auxiliary (shadow) / copy streams, pass
splitting and quarter-tile CTAs are scheduling choices, so every combination of the A/B switches has to produce the
same bytes; the pipelined read-back and the mirror target must deliver exactly what the blocking read-back returns."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _outputs(trace, data, out, env):
    from softglrender_b200 import workloads
    from softglrender_b200.scene.trace import read_outputs
    workloads.run_player(workloads.CUDA_PLAYER, trace, out=out, data_dir=data, env=env)
    r = read_outputs(out)
    os.remove(out)
    return r


@pytest.mark.parametrize("env", [{"SGL_NO_OVERLAP": "1"}, {"SGL_NO_SPLIT": "1"}, {"SGL_NO_PASS_SPLIT": "1"}, {"SGL_FORCE_FUSED": "1"}, {"SGL_NO_GRAPHS": "1"},
                                 {"SGL_NO_EARLY_VIS": "1"}, {"SGL_NO_RENAME": "1"}, {"SGL_RING": "3"}, {"SGL_CE_UPLOAD": "1"}, {"SGL_NO_MS_MASK": "1"}])
def test_scheduling_switches_do_not_change_the_frame(env, work_dir):
    """Config 2 at 960x540 MSAA4x (shadow pass on the auxiliary stream, heavy tiles split) and a blended KAT trace
    (pass split into deferred head + fused tail) against the same traces with one mechanism switched off."""
    from softglrender_b200 import workloads
    from softglrender_b200.scene import synth
    cases = []
    if workloads.A.find_assets_dir() is not None:
        cases.append(workloads.build_c2(os.path.join(work_dir, "c2s"), 960, 540))
    kat = os.path.join(work_dir, "kat_streams.sglt")
    synth.kat_trace(320, 200, msaa=True, reverse_z=True, seed=31).save(kat)
    cases.append((kat, work_dir))
    for trace, data in cases:
        base = _outputs(trace, data, os.path.join(data, "sw_base.out"), None)
        other = _outputs(trace, data, os.path.join(data, "sw_other.out"), env)
        for k in base:
            assert np.array_equal(base[k], other[k]), (os.path.basename(trace), env, k)


def test_async_readback_and_mirror_match_blocking_readback(work_dir):
    import torch
    from softglrender_b200 import capi, workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    capi.init(0)
    lib = capi.load()
    trace, data = workloads.build_c2(os.path.join(work_dir, "c2s"), 960, 540)
    p = capi.Player(trace, data)
    try:
        p.setup()
        tex = p.texture_handle("color")
        n = 960 * 540 * 4
        mirror = torch.zeros(n, dtype=torch.uint8, device="cuda")
        capi.check(lib.sgl_texture_set_mirror(tex, mirror.data_ptr()))
        pinned = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
        # frames submitted back to back; each read-back is queued behind its frame and overlaps the next one
        for f in range(6):
            p.frame(sync=False)
            capi.check(lib.sgl_texture_readback_async(tex, 0, 0, 1, pinned[f & 1].data_ptr(), n))
        capi.check(lib.sgl_readback_wait())
        capi.check(lib.sgl_wait_idle())
        ref = np.zeros(n, np.uint8)
        capi.check(lib.sgl_texture_readback(tex, 0, 0, 1, ref.ctypes.data, n))
        assert ref.any()
        assert np.array_equal(pinned[0].numpy(), ref) and np.array_equal(pinned[1].numpy(), ref)
        assert np.array_equal(mirror.cpu().numpy(), ref)
        # a too-small buffer is refused, nothing is queued
        assert lib.sgl_texture_readback_async(tex, 0, 0, 1, pinned[0].data_ptr(), 16) != 0
    finally:
        lib.sgl_texture_set_mirror(tex, None)
        p.close()


def test_depth_readback_right_after_a_shadow_pass(work_dir):
    """A depth-only pass runs on the auxiliary stream; any API call that touches its result must be ordered behind it
    (config 1: the shadow map read back immediately after the frame equals the one read after a full sync)."""
    from softglrender_b200 import capi, workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    capi.init(0)
    lib = capi.load()
    trace, data = workloads.build_c1(os.path.join(work_dir, "c1s"), 500, 400)
    p = capi.Player(trace, data)
    try:
        p.setup()
        for _ in range(3):
            p.frame(sync=False)
        a = p.readback("shadow")[0].copy()          # no explicit sync before it
        capi.check(lib.sgl_wait_idle())
        b = p.readback("shadow")[0].copy()
        assert np.array_equal(a, b) and a.view(np.float32).min() < 1.0
    finally:
        p.close()


def test_early_visibility_keeps_back_to_back_views_exact(work_dir):
    """Views submitted back to back (config 5, a different camera per view, same attachments): the visibility kernel of view
    n+1 starts while view n is still being shaded and its shadow pass renders into the shadow map's other backing store
    (SglCounters.early_vis / renamed_passes say that they did).  The last view's colour and depth
    must equal the same view rendered alone behind a full sync (the trace's tail), and the asynchronous read-back queued
    between two frames must not disturb it."""
    import torch
    from softglrender_b200 import capi, workloads
    from softglrender_b200.scene.trace import read_outputs
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    capi.init(0)
    lib = capi.load()
    views = [0, 300, 900, 1500, 2000]
    trace, data = workloads.build_c5(os.path.join(work_dir, "c5e"), "AfricanHead", views, n_total=4096, width=256, height=256)
    p = capi.Player(trace, data)
    try:
        p.setup()
        tex = p.texture_handle("color_v0")
        n = 256 * 256 * 4
        pinned = torch.zeros(n, dtype=torch.uint8).pin_memory()
        p.frame(sync=False)
        capi.check(lib.sgl_wait_idle())
        capi.check(lib.sgl_reset_counters())
        for _ in range(3):
            p.frame(sync=False)
            capi.check(lib.sgl_texture_readback_async(tex, 0, 0, 0, pinned.data_ptr(), n))
        capi.check(lib.sgl_wait_idle())
        ctr = capi.counters()
        early, renamed = ctr["early_vis"], ctr["renamed_passes"]
        last = len(views) - 1
        color = p.readback("color_v%d" % last)[0].copy()
        depth = p.readback("depth_v%d" % last)[0].copy()
        out = os.path.join(work_dir, "c5e_tail.out")
        p.tail(out)
        ref = read_outputs(out)
        os.remove(out)
        assert np.array_equal(color, ref["color_v%d" % last].reshape(-1).view(np.uint8))
        assert np.array_equal(depth, ref["depth_v%d" % last].reshape(-1).view(np.uint8))
        assert np.array_equal(pinned.numpy(), color)
        if not os.environ.get("SGL_NO_EARLY_VIS") and not os.environ.get("SGL_NO_OVERLAP"):
            assert early >= 3 * (len(views) - 1), early
        if not os.environ.get("SGL_NO_RENAME") and not os.environ.get("SGL_NO_OVERLAP"):
            assert renamed >= 3 * (len(views) - 1), renamed      # the shadow map of every view but the first
    finally:
        p.close()
