"""GPU side of the N > 1 path: tile-sharded rendering reproduces the 1-GPU frame byte for byte, the pack /
unpack kernels follow the host specification, and -- on a box with >= 2 GPUs -- the NCCL and direct-store gathers over a
real 2-rank process group (tests/mgpu_worker.py under torchrun)."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _play_every_rank(trace, data, world, policy, extra_targets=(), check_depth=True):
    """One GPU plays every rank of a `world`-way tile-sharded render in turn.  Before each rank's frame the colour target
    and every intermediate target in `extra_targets` are filled with junk, so anything the frame needs from tiles this
    rank does not render would show.  Returns (full unsharded frame, frame assembled from the ranks' owned tiles)."""
    import torch
    from softglrender_b200 import capi, multigpu as M
    capi.init(0)
    lib = capi.load()
    p = capi.Player(trace, data)
    try:
        p.setup()
        tex = p.texture_handle("color")
        others = [p.texture_handle(t) for t in extra_targets]
        p.frame(sync=True)
        buf, (w, h, _, _) = p.readback("color")
        full = buf.reshape(h, w, 4).copy()
        depth_full = p.readback("depth")[0].copy() if check_depth else None
        g = M.TileGather(w, h, 0, world, policy)
        g.install(lib)
        n = g.max_count * g.tile_bytes
        junk = torch.full((n,), 0xAB, dtype=torch.uint8, device="cuda")
        stage = torch.zeros(n, dtype=torch.uint8, device="cuda")
        assembled = np.zeros_like(full)
        for r in range(world):
            capi.check(lib.sgl_set_rank(r, world))
            for q in range(world):
                for t in [tex] + others:
                    capi.check(lib.sgl_tiles_unpack(t, q, junk.data_ptr(), n))
            p.frame(sync=False)
            cnt = C.c_int()
            capi.check(lib.sgl_tiles_pack(tex, r, stage.data_ptr(), n, C.byref(cnt)))
            capi.check(lib.sgl_wait_idle())
            assert cnt.value == g.counts[r]
            packed = stage.cpu().numpy()[:cnt.value * g.tile_bytes].reshape(cnt.value, M.TILE, M.TILE, 4)
            M.unpack_tiles_host(assembled, packed, g.owner, r)
            other = np.repeat(np.repeat(g.owner != r, M.TILE, axis=0), M.TILE, axis=1)[:h, :w]
            # tiles of other ranks must be untouched in the final target (still junk) ...
            got = p.readback("color")[0].reshape(h, w, 4)
            assert (got[other] == 0xAB).all(), "rank %d wrote tiles it does not own" % r
            # ... and the depth of owned pixels is bit-identical to the unsharded frame
            if check_depth:
                d = p.readback("depth")[0].reshape(h, w, -1)
                assert np.array_equal(d[~other], depth_full.reshape(h, w, -1)[~other]), "rank %d of %d" % (r, world)
        return full, assembled
    finally:
        capi.check(lib.sgl_set_tile_owner_map(None, 0, 0))
        capi.check(lib.sgl_set_rank(0, 1))
        p.close()


@pytest.mark.parametrize("name,world,policy", [("kat_ms4_revz", 2, "interleave"), ("kat_ms4_revz", 3, "bands"), ("kat_1x", 4, "interleave")])
def test_one_gpu_plays_every_rank_in_turn(name, world, policy, work_dir):
    """Each rank renders only the tiles it owns; the frame assembled from the ranks' packed tiles == the unsharded frame,
    for colour and (through the attachment read-back) per-sample depth."""
    trace, _ = make_golden.build_trace(name, work_dir)
    full, assembled = _play_every_rank(trace, work_dir, world, policy)
    assert np.array_equal(assembled, full)


@pytest.mark.parametrize("world,policy", [(2, "bands"), (4, "interleave")])
def test_config2_tile_sharded_equals_unsharded(world, policy, work_dir):
    """Config 2 at full size, tile-sharded: the shadow map the Blinn-Phong floor samples is a different-size attachment and
    is therefore rendered whole by every rank; the 1920x1080 MSAA frame assembled from the ranks' tiles is byte-identical."""
    from softglrender_b200 import workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    trace, data = workloads.build_c2(os.path.join(work_dir, "c2"))
    full, assembled = _play_every_rank(trace, data, world, policy)
    assert full.std() > 5.0
    assert np.array_equal(assembled, full)


@pytest.mark.parametrize("world,policy,halo", [(2, "bands", 32), (4, "interleave", 32), (2, "bands", -1), (3, "stripes", 32)])
def test_config3_fxaa_tile_sharded_needs_and_gets_its_halo(world, policy, halo, work_dir, monkeypatch):
    """Config 3 (3840x2160: shadow pass, opaque + blended main pass into the FXAA input, FXAA pass into the output),
    tile-sharded: the FXAA pass reads up to 18.5 px + a bilinear footprint around the pixel it shades
    (FxaaSoft.h:73-74,169-210), so every rank renders a 32-px halo of the FXAA input around its own tiles
    (sgl_texture_set_shard_halo; halo -1 = the whole input).  The assembled frame is byte-identical to the unsharded one."""
    from softglrender_b200 import workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    monkeypatch.setenv("SGL_SHARD_HALO", str(halo))
    trace, data = workloads.build_c3(os.path.join(work_dir, "c3"))
    full, assembled = _play_every_rank(trace, data, world, policy, extra_targets=("color_prefxaa",))
    assert full.std() > 5.0
    assert np.array_equal(assembled, full)


def test_config3_without_halo_is_wrong_at_tile_borders(work_dir, monkeypatch):
    """The negative of the test above: with no halo the FXAA pass reads junk across ownership borders (and the test's junk
    fill makes that visible) -- the halo is what makes sharded config 3 correct, not luck."""
    from softglrender_b200 import workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    monkeypatch.setenv("SGL_SHARD_HALO", "0")
    trace, data = workloads.build_c3(os.path.join(work_dir, "c3"), 1920, 1080)
    full, assembled = _play_every_rank(trace, data, 2, "bands", extra_targets=("color_prefxaa",))
    assert not np.array_equal(assembled, full)


def test_unpack_kernel_matches_host_specification():
    import torch
    from softglrender_b200 import capi, multigpu as M
    capi.init(0)
    lib = capi.load()
    w, h, world = 257, 131, 3
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, (h, w, 4)).astype(np.uint8)
    desc = capi.SglTextureDesc(width=w, height=h, type=0, format=0, use_mipmaps=0, multi_sample=0, layout=0)
    tex = C.c_int()
    capi.check(lib.sgl_texture_create(C.byref(desc), C.byref(tex)))
    try:
        owner = M.tile_owner_map(w, h, world)
        capi.check(lib.sgl_set_tile_owner_map(owner.ctypes.data, owner.shape[1], owner.shape[0]))
        for r in range(world):
            packed = torch.from_numpy(M.pack_tiles_host(img, owner, r)).cuda().contiguous()
            capi.check(lib.sgl_tiles_unpack(tex.value, r, packed.data_ptr(), packed.numel()))
        out = np.zeros((h, w, 4), np.uint8)
        capi.check(lib.sgl_texture_readback(tex.value, 0, 0, 0, out.ctypes.data, out.nbytes))
        assert np.array_equal(out, img)
        small = torch.zeros(16, dtype=torch.uint8, device="cuda")
        assert lib.sgl_tiles_pack(tex.value, 0, small.data_ptr(), 16, None) != 0      # staging buffer too small: refused
    finally:
        capi.check(lib.sgl_set_tile_owner_map(None, 0, 0))
        capi.check(lib.sgl_texture_destroy(tex.value))


def test_two_ranks_nccl_and_direct_store_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:]
