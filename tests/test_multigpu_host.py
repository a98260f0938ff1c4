"""Host-side logic of the N > 1 path on CPU: ownership maps, shard arithmetic, pack/unpack index maths, and the gather
protocol over a real world_size-2 process group (gloo).  The device kernels that replace the numpy pack/unpack are
checked against the same functions in tests/test_multigpu_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _image(w, h, seed=0):
    rng = np.random.RandomState(seed)
    return rng.randint(0, 256, (h, w, 4)).astype(np.uint8)


@pytest.mark.parametrize("policy", ["interleave", "bands"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("size", [(1920, 1080), (257, 131), (16, 16), (500, 17)])
def test_owner_map_is_a_partition(policy, world, size):
    from softglrender_b200 import multigpu as M
    w, h = size
    owner = M.tile_owner_map(w, h, world, policy)
    tx, ty = M.tiles_of(w, h)
    assert owner.shape == (ty, tx) and owner.dtype == np.uint8
    assert owner.max() < world
    idx = np.concatenate([M.owned_tile_indices(owner, r) for r in range(world)])
    assert sorted(idx.tolist()) == list(range(tx * ty))          # every tile exactly once
    counts = [(owner == r).sum() for r in range(world)]
    if policy == "bands":
        assert max(counts) - min(counts) <= tx                    # whole tile rows
        for r in range(world):                                    # contiguous bands
            rows = np.flatnonzero((owner == r).any(axis=1))
            assert len(rows) == 0 or rows[-1] - rows[0] + 1 == len(rows)
    elif tx * ty >= 64 * world:
        assert max(counts) <= 1.35 * (tx * ty) / world            # balanced


def test_interleave_balances_a_centred_object():
    """The helmet covers the middle of the frame: every rank must get a comparable share of the centre region."""
    from softglrender_b200 import multigpu as M
    for world in (2, 4, 8):
        owner = M.tile_owner_map(1920, 1080, world, "interleave")
        ty, tx = owner.shape
        centre = owner[ty // 4: 3 * ty // 4, tx // 3: 2 * tx // 3]
        counts = np.array([(centre == r).sum() for r in range(world)], float)
        assert counts.min() / counts.max() > 0.7, (world, counts)


def test_pack_unpack_roundtrip_ragged():
    from softglrender_b200 import multigpu as M
    for (w, h), world in (((257, 131), 3), ((64, 48), 2), ((500, 17), 4)):
        img = _image(w, h, 1)
        owner = M.tile_owner_map(w, h, world)
        out = np.zeros_like(img)
        for r in range(world):
            M.unpack_tiles_host(out, M.pack_tiles_host(img, owner, r), owner, r)
        assert np.array_equal(out, img)


def test_shard_units_covers_everything_once():
    from softglrender_b200 import multigpu as M
    for n, world in ((4096, 8), (7, 2), (3, 4), (0, 2)):
        seen = sorted(u for r in range(world) for u in M.shard_units(n, r, world))
        assert seen == list(range(n))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from softglrender_b200 import multigpu as M
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for (w, h), policy in (((320, 200), "interleave"), ((257, 131), "bands"), ((1920, 1080), "interleave")):
            full = _image(w, h, 7)
            g = M.TileGather(w, h, rank, world, policy)
            # each rank "renders" only the tiles it owns; everything else is garbage that must not reach rank 0's frame
            mine = np.full_like(full, 0xAB)
            M.unpack_tiles_host(mine, M.pack_tiles_host(full, g.owner, rank), g.owner, rank)
            got = g.gather_host(mine)
            if rank == 0:
                assert np.array_equal(got, full), (w, h, policy)
            else:
                assert got is None
        # frame / view parallel: unit u -> rank u % world, frames gathered to rank 0 in unit order
        units = 5
        frames = {u: _image(64, 48, 100 + u) for u in range(units)}
        rounds = (units + world - 1) // world
        collected = {}
        for k in range(rounds):
            u = k * world + rank
            frame = frames[u] if u < units else np.zeros((48, 64, 4), np.uint8)
            got = M.gather_frames_host(frame, rank, world)
            if rank == 0:
                for r in range(world):
                    if k * world + r < units:
                        collected[k * world + r] = got[r]
        if rank == 0:
            assert sorted(collected) == list(range(units))
            assert all(np.array_equal(collected[u], frames[u]) for u in range(units))
        dist.barrier()
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_tile_and_frame_gather_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
