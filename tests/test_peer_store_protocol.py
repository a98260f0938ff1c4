"""Host-side protocol of the direct-store gather (multigpu.PeerFrameStore) on CPU: two gloo ranks drive the real class
against a stand-in for the C ABI whose "device memory" is POSIX shared memory and whose stream-ordered flag kernels
execute synchronously (a stricter schedule than the GPU's: whatever blocks a stream here blocks the caller).  Every frame
a rank "renders" is a pattern derived from (rank, frame); rank 0's consume callback checks every collected slot, so a slot
overwritten too early, a missed frame or a deadlock shows up -- under random skew between the ranks."""
import ctypes as C
import os
import random
import socket
import sys
import time
from multiprocessing import shared_memory

import numpy as np
import pytest

from conftest import ROOT


class FakePeerLib:
    """sgl_peer_* / mirror / async read-back subset of libsglcuda.so over shared memory."""

    def __init__(self, tag, rank):
        self.tag, self.rank, self.blocks, self.next = tag, rank, {}, 1
        self.mirror, self.timeouts, self.render = {}, 0, None

    # -- address space: base = k << 40
    def _map(self, shm):
        base = self.next << 40
        self.next += 1
        self.blocks[base] = shm
        return base

    def _view(self, ptr, n):
        base = (ptr >> 40) << 40
        off = ptr - base
        return np.ndarray((n,), np.uint8, self.blocks[base].buf, off)

    def _u32(self, ptr):
        return np.ndarray((1,), np.uint32, self.blocks[(ptr >> 40) << 40].buf, ptr & ((1 << 40) - 1))

    def sgl_last_error(self):
        return b"fake"

    def sgl_peer_alloc(self, nbytes, ptr_ref, handle):
        name = "sglt_%s_%d_%d" % (self.tag, self.rank, self.next)
        shm = shared_memory.SharedMemory(name=name, create=True, size=max(int(nbytes), 256))
        shm.buf[:] = bytes(len(shm.buf))
        ptr_ref._obj.value = self._map(shm)
        raw = name.encode().ljust(64, b"\0")
        C.memmove(handle, raw, 64)
        return 0

    def sgl_peer_open(self, handle, ptr_ref):
        name = bytes(handle).split(b"\0")[0].decode()
        ptr_ref._obj.value = self._map(shared_memory.SharedMemory(name=name))
        return 0

    def sgl_texture_set_mirror(self, tex, ptr):
        self.mirror[tex] = ptr
        return 0

    def sgl_texture_device_ptr(self, *a):
        return 0

    def sgl_peer_signal(self, ptr, value):
        self._u32(ptr)[0] = value
        return 0

    sgl_peer_signal_after_copies = sgl_peer_signal

    def sgl_texture_readback_async(self, tex, layer, level, kind, dst, n):
        self._view(dst, n)[:] = self.render            # copy-engine form: the finished frame is pushed into the slot
        return 0

    def _wait(self, ptr, count, value, timeout_ms):
        t0 = time.time()
        for i in range(count):
            while int(self._u32(ptr + 64 * i)[0]) < value:
                if (time.time() - t0) * 1e3 > timeout_ms:
                    self.timeouts += 1
                    return
                time.sleep(0.0002)

    def sgl_peer_wait(self, ptr, count, value, timeout_ms):
        self._wait(ptr, count, value, timeout_ms)
        return 0

    def sgl_peer_collect(self, ptr, count, value, peers, timeout_ms, side):
        self._wait(ptr, count, value, timeout_ms)
        for i in range(1, count):
            self._u32(peers[i])[0] = value
        return 0

    def sgl_peer_timeouts(self, ref):
        ref._obj.value = self.timeouts
        return 0

    def close(self):
        for shm in self.blocks.values():
            shm.close()


def _pattern(rank, frame, n):
    return ((np.arange(n, dtype=np.uint32) * 7 + rank * 31 + frame * 101) & 0xFF).astype(np.uint8)


def _worker(rank, world, port, tmp, dma, lag, slots):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from softglrender_b200 import multigpu as M
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = FakePeerLib(os.path.basename(tmp), rank)
    created = []
    try:
        w, h, frames = 24, 10, 40
        n = w * h * 4
        store = M.PeerFrameStore(lib, w, h, rank, world, frames_per_slot=world, slots=slots, lag=lag, dma=dma, timeout_ms=20000)
        created = [b for b in lib.blocks.values()]
        rng = random.Random(1234 + rank)
        seen = []

        def consume(ptr):
            c = len(seen)
            a = lib._view(ptr, world * n).reshape(world, n)
            for r in range(world):
                assert np.array_equal(a[r], _pattern(r, c, n)), "slot of rank %d, frame %d was overwritten or never written" % (r, c)
            seen.append(c)
        for f in range(frames):
            if rng.random() < 0.3:
                time.sleep(rng.random() * 0.004)            # skew between the ranks
            store.begin_frame(7, rank)
            lib.render = _pattern(rank, f, n)
            if not dma:                                     # the "shading kernel" stores through the mirror pointer
                lib._view(lib.mirror[7], n)[:] = lib.render
            store.end_frame(consume if rank == 0 else None)
        store.flush(consume if rank == 0 else None)
        dist.barrier()
        assert store.timeouts() == 0
        if rank == 0:
            assert seen == list(range(frames))
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.barrier()
        lib.close()
        for shm in created:
            try:
                if shm.name.startswith("sglt_%s_%d_" % (os.path.basename(tmp), rank)):
                    shm.unlink()
            except Exception:
                pass
        dist.destroy_process_group()


@pytest.mark.parametrize("dma,lag,slots", [(False, 0, 2), (False, 2, 4), (True, 2, 4), (False, 3, 4)])
def test_peer_frame_store_protocol_world2(dma, lag, slots, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), dma, lag, slots), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
