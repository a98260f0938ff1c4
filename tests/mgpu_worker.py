"""torchrun worker of the multi-GPU parity checks (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_worker.py

SURVEY 8e parity rule: the N-GPU image must equal the 1-GPU image byte for byte.  Every rank first renders the full
frame alone (its own 1-GPU image), then the same frame tile-sharded; rank 0 receives the other ranks' tiles (a) through
pack -> NCCL gather -> unpack and (b) by direct peer stores from the shading kernel, and compares both with its own
full render.  Also checks the frame-parallel direct-store gather ([world] frames in rank 0's store).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import torch
    import torch.distributed as dist
    import make_golden
    from softglrender_b200 import capi, multigpu as M
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctl = dist.new_group(backend="gloo")
    capi.init(local, rank, world)
    lib = capi.load()
    stream = torch.cuda.Stream()          # the library renders on torch's (non-default) current stream: NCCL ops and
    torch.cuda.set_stream(stream)         # torch copies are then stream-ordered with the kernels, no host waits
    capi.check(lib.sgl_set_stream(C.c_void_p(stream.cuda_stream)))
    work = os.path.join(ROOT, "build", "tests", "mgpu%d" % rank)
    os.makedirs(work, exist_ok=True)
    report = {}
    for name in ("kat_ms4_revz", "kat_1x"):
        trace, _ = make_golden.build_trace(name, work)
        p = capi.Player(trace, work)
        p.setup()
        tex = p.texture_handle("color")
        w_, h_ = C.c_int(), C.c_int()
        capi.check(lib.sgl_texture_level_size(tex, 0, C.byref(w_), C.byref(h_)))
        w, h = w_.value, h_.value
        p.frame(sync=True)
        full = p.readback("color")[0].reshape(h, w, 4).copy()

        # (a) tile-sharded + NCCL gather
        for policy in ("interleave", "bands"):
            g = M.TileGather(w, h, rank, world, policy)
            g.install(lib)
            junk = torch.full((g.max_count * g.tile_bytes,), 0xAB, dtype=torch.uint8, device="cuda")
            for r in range(world):        # forget the full render: every tile must be produced again by its owner
                capi.check(lib.sgl_tiles_unpack(tex, r, junk.data_ptr(), junk.numel()))
            p.frame(sync=False)
            g.gather_device(lib, tex)
            capi.check(lib.sgl_wait_idle())
            got = p.readback("color")[0].reshape(h, w, 4)
            if rank == 0:
                assert np.array_equal(got, full), "%s: NCCL tile gather (%s) differs from the 1-GPU frame" % (name, policy)
                report["%s.nccl.%s" % (name, policy)] = "byte-exact"

        # (b) tile-sharded + direct peer stores
        g = M.TileGather(w, h, rank, world, "interleave")
        g.install(lib)
        store = M.PeerFrameStore(lib, w, h, rank, world, frames_per_slot=1, slots=2, control_group=ctl)
        frames = []
        for f in range(5):
            store.begin_frame(tex, 0)
            p.frame(sync=False)

            def consume(ptr):
                frames.append(M.device_view(ptr, w * h * 4).clone())
            store.end_frame(consume if rank == 0 else None)
        capi.check(lib.sgl_wait_idle())
        torch.cuda.synchronize()
        dist.barrier()
        assert store.timeouts() == 0, "peer wait timed out"
        if rank == 0:
            for f, t in enumerate(frames):
                assert np.array_equal(t.cpu().numpy().reshape(h, w, 4), full), "%s: direct-store tile gather differs (frame %d)" % (name, f)
            report["%s.p2p.tiles" % name] = "byte-exact x%d frames" % len(frames)
        capi.check(lib.sgl_texture_set_mirror(tex, None))

        # (c) frame-parallel + direct peer stores: every rank renders the whole frame into its slot of rank 0's store
        capi.check(lib.sgl_set_tile_owner_map(None, 0, 0))
        # (4 slots, rank 0 collects two frames behind the renderers and drains the tail with flush())
        store2 = M.PeerFrameStore(lib, w, h, rank, world, frames_per_slot=world, slots=4, control_group=ctl, lag=2)
        got = []
        for f in range(7):
            store2.begin_frame(tex, rank)
            p.frame(sync=False)

            def consume2(ptr):
                got.append(M.device_view(ptr, world * w * h * 4).clone())
            store2.end_frame(consume2 if rank == 0 else None)
        store2.flush(consume2 if rank == 0 else None)
        capi.check(lib.sgl_wait_idle())
        torch.cuda.synchronize()
        dist.barrier()
        assert store2.timeouts() == 0
        if rank == 0:
            assert len(got) == 7
            for t in got:
                a = t.cpu().numpy().reshape(world, h, w, 4)
                for r in range(world):
                    assert np.array_equal(a[r], full), "%s: frame-parallel store, slot %d differs" % (name, r)
            report["%s.p2p.frames" % name] = "byte-exact x%d ranks" % world
        capi.check(lib.sgl_texture_set_mirror(tex, None))

        # (d) frame-parallel, copy-engine pushes instead of mirrored stores
        store3 = M.PeerFrameStore(lib, w, h, rank, world, frames_per_slot=world, slots=4, control_group=ctl, lag=2, dma=True)
        got3 = []
        for f in range(6):
            store3.begin_frame(tex, rank)
            p.frame(sync=False)
            store3.end_frame((lambda ptr: got3.append(M.device_view(ptr, world * w * h * 4).clone())) if rank == 0 else None)
        store3.flush((lambda ptr: got3.append(M.device_view(ptr, world * w * h * 4).clone())) if rank == 0 else None)
        capi.check(lib.sgl_wait_idle())
        torch.cuda.synchronize()
        dist.barrier()
        assert store3.timeouts() == 0
        if rank == 0:
            assert len(got3) == 6
            for t in got3:
                a = t.cpu().numpy().reshape(world, h, w, 4)
                for r in range(world):
                    assert np.array_equal(a[r], full), "%s: copy-engine gather, slot %d differs" % (name, r)
            report["%s.dma.frames" % name] = "byte-exact x%d ranks" % world
        p.close()

    # ---- the BASELINE configs tile-sharded over a real process group: config 2 (the shadow map every rank renders whole is
    #      sampled by the floor) and config 3 (FXAA needs a 32-px halo of its input around the owned tiles)
    from softglrender_b200 import workloads
    if workloads.A.find_assets_dir() is not None:
        shared = os.path.join(ROOT, "build", "tests")
        cases = (("config2", lambda: workloads.build_c2(os.path.join(shared, "c2")), None),
                 ("config3", lambda: workloads.build_c3(os.path.join(shared, "c3"), 1920, 1080), 32))
        for name, build, halo in cases:
            if rank == 0:
                build()                      # traces / IBL maps are cached on disk: one rank writes them
            dist.barrier()
            trace, data = build()
            if halo is not None:
                os.environ["SGL_SHARD_HALO"] = str(halo)
            p = capi.Player(trace, data)
            p.setup()
            tex = p.texture_handle("color")
            w_, h_ = C.c_int(), C.c_int()
            capi.check(lib.sgl_texture_level_size(tex, 0, C.byref(w_), C.byref(h_)))
            w, h = w_.value, h_.value
            p.frame(sync=True)
            full = p.readback("color")[0].reshape(h, w, 4).copy()
            for policy in ("interleave", "bands"):
                g = M.TileGather(w, h, rank, world, policy)
                g.install(lib)
                junk = torch.full((g.max_count * g.tile_bytes,), 0xAB, dtype=torch.uint8, device="cuda")
                for r in range(world):
                    capi.check(lib.sgl_tiles_unpack(tex, r, junk.data_ptr(), junk.numel()))
                p.frame(sync=False)
                g.gather_device(lib, tex)
                capi.check(lib.sgl_wait_idle())
                got = p.readback("color")[0].reshape(h, w, 4)
                if rank == 0:
                    assert np.array_equal(got, full), "%s: NCCL tile gather (%s) differs from the 1-GPU frame" % (name, policy)
                    report["%s.nccl.%s" % (name, policy)] = "byte-exact"
            g = M.TileGather(w, h, rank, world, "interleave")
            g.install(lib)
            store = M.PeerFrameStore(lib, w, h, rank, world, frames_per_slot=1, slots=2, control_group=ctl)
            frames = []
            for f in range(4):
                store.begin_frame(tex, 0)
                p.frame(sync=False)
                store.end_frame((lambda ptr: frames.append(M.device_view(ptr, w * h * 4).clone())) if rank == 0 else None)
            capi.check(lib.sgl_wait_idle())
            torch.cuda.synchronize()
            dist.barrier()
            assert store.timeouts() == 0, "peer wait timed out"
            if rank == 0:
                for f, t in enumerate(frames):
                    assert np.array_equal(t.cpu().numpy().reshape(h, w, 4), full), "%s: direct-store tile gather differs (frame %d)" % (name, f)
                report["%s.p2p.tiles" % name] = "byte-exact x%d frames" % len(frames)
            capi.check(lib.sgl_texture_set_mirror(tex, None))
            capi.check(lib.sgl_set_tile_owner_map(None, 0, 0))
            os.environ.pop("SGL_SHARD_HALO", None)
            p.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK " + json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
