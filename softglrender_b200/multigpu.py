"""Multi-GPU plumbing of RendererCUDA: one process per GPU (torchrun), sort-first screen-tile ownership for a single
large frame, view/frame parallelism for batches, and the gather of finished tiles / frames to rank 0.

The render path itself has no exchange step: geometry and textures are replicated, every rank runs the vertex stage
for all vertices and bins only into the tiles it owns (``sgl_set_tile_owner_map``).  The only collective is the gather
of finished RGBA8 pixels to rank 0, in two interchangeable forms:

* ``TileGather``      NCCL form: owned tiles are packed by a kernel into a dense staging buffer, moved with
                      ``torch.distributed.gather`` (ncclSend/Recv over NVLink), and unpacked on rank 0.
* ``PeerFrameStore``  direct-store form: rank 0 exports its frame store through a CUDA IPC handle, every rank maps it,
                      and the shading kernel writes each finished (resolved) pixel straight into rank 0's memory over
                      NVLink (``sgl_texture_set_mirror``) -- the gather is fused into the producing kernel; stream-ordered
                      system-scope flags replace the collective's synchronisation.

Host-side logic (ownership maps, shard arithmetic, pack/unpack index maths, the gather protocol) is backend-agnostic and
is covered on CPU with world_size-2 ``gloo`` tests (tests/test_multigpu_host.py); the numpy pack/unpack below exists for
those tests and as the specification of the device kernels (``sglTilePackKernel``).
"""
import ctypes as C

import numpy as np

TILE = 16   # SGL_TILE (csrc/sgl_types.h); checked against sgl_tile_size() when the library is used


# ---------------------------------------------------------------------------------------------------- ownership maps
def tiles_of(width, height, tile=TILE):
    return (width + tile - 1) // tile, (height + tile - 1) // tile


def tile_owner_map(width, height, world, policy="interleave", tile=TILE, block=4):
    """owner[ty, tx] in [0, world).

    ``interleave``: blocks of ``block`` x ``block`` tiles dealt round-robin along a diagonal-shifted order, so that every
    rank gets an even share of any screen region (the helmet sits in the middle of the frame);
    ``bands``: contiguous horizontal bands of whole tile rows; ``stripes[:k]``: k horizontal stripes per rank dealt
    round-robin (what a pass with a screen-space halo, e.g. FXAA, wants: cheap halo, balanced load).
    """
    tx, ty = tiles_of(width, height, tile)
    if world < 1 or world > 255:
        raise ValueError("world size %d" % world)
    ys, xs = np.mgrid[0:ty, 0:tx]
    if policy == "interleave":
        owner = ((xs // block) + (ys // block) * (1 + world // 2)) % world
    elif policy == "bands":
        bounds = [(r * ty) // world for r in range(world + 1)]
        owner = np.zeros((ty, tx), np.int64)
        for r in range(world):
            owner[bounds[r]:bounds[r + 1], :] = r
    elif policy.startswith("stripes"):
        # k horizontal stripes per rank ("stripes" = 2, "stripes:k"), dealt round-robin, boundaries spread evenly over the
        # tile rows: contiguous enough that a screen-space halo stays cheap, fine enough that floor, model and sky are
        # spread over all ranks
        per_rank = int(policy.split(":")[1]) if ":" in policy else 2
        n = min(world * per_rank, ty)
        bounds = [(k * ty) // n for k in range(n + 1)]
        owner = np.zeros((ty, tx), np.int64)
        for k in range(n):
            owner[bounds[k]:bounds[k + 1], :] = k % world
    else:
        raise ValueError("unknown tile ownership policy %r" % policy)
    return np.ascontiguousarray(owner.astype(np.uint8))


def owned_tile_indices(owner, rank):
    """Tile indices (ty * tiles_x + tx) owned by ``rank`` in ascending order = the order of the packed staging buffer."""
    return np.flatnonzero(owner.reshape(-1) == rank)


def shard_units(n_units, rank, world):
    """View / frame parallelism: unit u is rendered by rank u mod world."""
    return range(rank, n_units, world)


# ------------------------------------------------------------------------------- host pack / unpack (specification)
def pack_tiles_host(image, owner, rank, tile=TILE):
    """image (H, W, 4) uint8 -> (n_owned, tile, tile, 4); pixels outside the image are zero."""
    h, w = image.shape[:2]
    tx = owner.shape[1]
    idx = owned_tile_indices(owner, rank)
    out = np.zeros((len(idx), tile, tile, 4), np.uint8)
    for k, t in enumerate(idx):
        y0, x0 = (t // tx) * tile, (t % tx) * tile
        blk = image[y0:y0 + tile, x0:x0 + tile]
        out[k, :blk.shape[0], :blk.shape[1]] = blk
    return out


def unpack_tiles_host(image, packed, owner, rank, tile=TILE):
    h, w = image.shape[:2]
    tx = owner.shape[1]
    for k, t in enumerate(owned_tile_indices(owner, rank)):
        y0, x0 = (t // tx) * tile, (t % tx) * tile
        hh, ww = min(tile, h - y0), min(tile, w - x0)
        image[y0:y0 + hh, x0:x0 + ww] = packed[k, :hh, :ww]
    return image


def device_view(ptr, nbytes):
    """uint8 torch view of raw device memory (library buffers, peer mappings); no copy, no ownership."""
    import torch

    class _Dev:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Dev(), device="cuda")


# --------------------------------------------------------------------------------------------------- NCCL-form gather
class TileGather:
    """Gather of owned tiles to rank 0 through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, width, height, rank, world, policy="interleave", tile=TILE, group=None):
        self.width, self.height, self.rank, self.world, self.tile, self.group = width, height, rank, world, tile, group
        self.owner = tile_owner_map(width, height, world, policy, tile)
        self.counts = [int((self.owner == r).sum()) for r in range(world)]
        self.max_count = max(self.counts)
        self.tile_bytes = tile * tile * 4
        self._stage = None
        self._recv = None

    # -- CPU path of the protocol (gloo tests; also documents what the device path does)
    def gather_host(self, image):
        """Every rank passes its (H, W, 4) image (only owned tiles need to be valid); rank 0 returns the assembled
        frame, the others None."""
        import torch
        import torch.distributed as dist
        mine = pack_tiles_host(image, self.owner, self.rank, self.tile)
        stage = torch.zeros((self.max_count, self.tile, self.tile, 4), dtype=torch.uint8)
        stage[:len(mine)] = torch.from_numpy(mine)
        recv = [torch.empty_like(stage) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(stage, recv, dst=0, group=self.group)
        if self.rank != 0:
            return None
        out = np.array(image, copy=True)
        for r in range(1, self.world):
            unpack_tiles_host(out, recv[r].numpy()[:self.counts[r]], self.owner, r, self.tile)
        return out

    # -- device path
    def install(self, lib):
        """Give the library this rank's ownership map (sgl_set_tile_owner_map)."""
        from . import capi
        assert lib.sgl_tile_size() == self.tile
        capi.check(lib.sgl_set_tile_owner_map(self.owner.ctypes.data, self.owner.shape[1], self.owner.shape[0]))

    def gather_device(self, lib, texture):
        """pack (kernel) -> dist.gather (NCCL) -> unpack (kernels, rank 0).  The library must run on torch's current
        stream (sgl_set_stream) so that everything is stream-ordered without host synchronisation."""
        import torch
        import torch.distributed as dist
        from . import capi
        n = self.max_count * self.tile_bytes
        if self._stage is None:
            self._stage = torch.zeros(n, dtype=torch.uint8, device="cuda")
            self._recv = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(self.world)] if self.rank == 0 else None
        cnt = C.c_int()
        capi.check(lib.sgl_tiles_pack(texture, self.rank, self._stage.data_ptr(), n, C.byref(cnt)))
        dist.gather(self._stage, self._recv, dst=0, group=self.group)
        if self.rank == 0:
            for r in range(1, self.world):
                capi.check(lib.sgl_tiles_unpack(texture, r, self._recv[r].data_ptr(), n))


def gather_frames_host(frame, rank, world, group=None):
    """Frame/view-parallel gather (CPU form): rank 0 gets [world] frames."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(frame))
    recv = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, recv, dst=0, group=group)
    return [r.numpy() for r in recv] if rank == 0 else None


# --------------------------------------------------------------------------------------------- direct-store gather
class PeerFrameStore:
    """Rank 0 owns ``slots`` x ``frames_per_slot`` RGBA8 frames + one 64-byte flag per rank; every rank maps the store
    (CUDA IPC) and renders with its colour target mirrored into it, so finished pixels land in rank 0's HBM as the
    shading kernel produces them.  Flow control (all stream-ordered, no host synchronisation):

      producer r, frame f:  wait(consumed[r] >= f - slots + 1)  ->  render (mirror = slot f % slots)  ->  signal(done[r] = f + 1)
      rank 0,     frame f:  wait(done[0..world) >= f + 1)        ->  consume                           ->  signal(consumed[r] = f + 1) for all r
                            (with lag > 0 rank 0 does this for frame f - lag while frame f renders)

    ``consumed`` flags live in each producer's own memory (rank 0 maps them), ``done`` flags in rank 0's, so every spin
    reads local HBM and every signal is one remote store.
    """
    FLAG_STRIDE = 64

    def __init__(self, lib, width, height, rank, world, frames_per_slot, slots=2, control_group=None, timeout_ms=2000, lag=0,
                 dma=False):
        import torch
        import torch.distributed as dist
        from . import capi
        self.lib, self.rank, self.world, self.slots, self.fps, self.timeout_ms = lib, rank, world, slots, frames_per_slot, timeout_ms
        # rank 0 collects frame f - lag while everybody renders frame f: per-frame jitter between the ranks is absorbed
        # instead of pacing all of them on the slowest one each frame (0 <= lag <= slots - 1; flush() drains the tail)
        self.lag = max(0, min(lag, slots - 1))
        self.collected = 0
        # dma=True: instead of mirrored stores from the shading kernel, the finished (resolved) frame is pushed into the
        # slot by the copy engine (sgl_texture_readback_async onto the peer mapping) -- large NVLink packets, no SM work;
        # only for whole frames (frame/view parallel), a tile-sharded frame has no contiguous owned region
        self.dma = dma
        self.texture, self.kind = None, 1
        self._peer_array = None
        self.frame_bytes = width * height * 4
        self.header = self.FLAG_STRIDE * max(world, 1)
        store_bytes = self.header + self.frame_bytes * frames_per_slot * slots
        # every rank: local `consumed` flag; rank 0: the store.  Failures (CUDA IPC not permitted, ...) are agreed on by
        # all ranks before anyone raises, so that nobody is left waiting in a collective.
        self.local_flag = C.c_void_p()
        self.store = C.c_void_p()
        h_flag, h_store = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
        err = None
        try:
            capi.check(lib.sgl_peer_alloc(256, C.byref(self.local_flag), h_flag))
            if rank == 0:
                capi.check(lib.sgl_peer_alloc(store_bytes, C.byref(self.store), h_store))
        except RuntimeError as e:
            err = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, (bytes(h_flag), bytes(h_store), err), group=control_group)
        self._raise_if_any([h[2] for h in handles])
        self.peer_flags = []
        try:
            if rank != 0:
                buf = (C.c_uint8 * 64).from_buffer_copy(handles[0][1])
                capi.check(lib.sgl_peer_open(buf, C.byref(self.store)))
            else:
                for r in range(world):
                    if r == 0:
                        self.peer_flags.append(self.local_flag.value)
                    else:
                        p = C.c_void_p()
                        buf = (C.c_uint8 * 64).from_buffer_copy(handles[r][0])
                        capi.check(lib.sgl_peer_open(buf, C.byref(p)))
                        self.peer_flags.append(p.value)
        except RuntimeError as e:
            err = str(e)
        errs = [None] * world
        dist.all_gather_object(errs, err, group=control_group)
        self._raise_if_any(errs)
        self.frame_no = 0

    @staticmethod
    def _raise_if_any(errs):
        bad = [(r, e) for r, e in enumerate(errs) if e]
        if bad:
            raise RuntimeError("peer frame store unavailable (rank %d: %s)" % bad[0])

    def slot_ptr(self, frame_no, index):
        return self.store.value + self.header + self.frame_bytes * ((frame_no % self.slots) * self.fps + index)

    def begin_frame(self, texture, index):
        """Queue the back-pressure wait and point the texture's mirror at this frame's slot."""
        from . import capi
        f = self.frame_no
        if self.rank != 0 and f >= self.slots:
            capi.check(self.lib.sgl_peer_wait(self.local_flag.value, 1, f - self.slots + 1, self.timeout_ms))
        if self.dma:
            if texture != self.texture:      # resolved colour of a multisample target, else the colour image itself
                ptr, sz = C.c_void_p(), C.c_size_t()
                self.kind = 1 if self.lib.sgl_texture_device_ptr(texture, 0, 0, 1, C.byref(ptr), C.byref(sz)) == 0 else 0
            self.texture, self.index = texture, index
        else:
            capi.check(self.lib.sgl_texture_set_mirror(texture, self.slot_ptr(f, index)))

    def end_frame(self, consume=None):
        """Queue this rank's done-signal; on rank 0 also the wait for every rank, ``consume(ptr_of_slot)`` and the
        consumed-signals."""
        from . import capi
        f = self.frame_no
        if self.dma:
            capi.check(self.lib.sgl_texture_readback_async(self.texture, 0, 0, self.kind, self.slot_ptr(f, self.index), self.frame_bytes))
            capi.check(self.lib.sgl_peer_signal_after_copies(self.store.value + self.FLAG_STRIDE * self.rank, f + 1))
        else:
            capi.check(self.lib.sgl_peer_signal(self.store.value + self.FLAG_STRIDE * self.rank, f + 1))
        self.frame_no = f + 1
        if self.rank == 0 and f - self.lag >= self.collected:
            self._collect(f - self.lag, consume)

    def _collect(self, upto, consume):
        from . import capi
        if self._peer_array is None:
            self._peer_array = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.peer_flags])
        while self.collected <= upto:
            c = self.collected
            if consume is None:
                # nothing to do with the frame on this stream: wait + consumed-signals in ONE kernel on the library's
                # side stream, so rank 0's own rendering is never serialised behind the slowest peer
                capi.check(self.lib.sgl_peer_collect(self.store.value, self.world, c + 1, self._peer_array, self.timeout_ms, 1))
            else:
                capi.check(self.lib.sgl_peer_wait(self.store.value, self.world, c + 1, self.timeout_ms))
                consume(self.slot_ptr(c, 0))
                capi.check(self.lib.sgl_peer_collect(self.store.value, self.world, c + 1, self._peer_array, self.timeout_ms, 0))
            self.collected = c + 1

    def flush(self, consume=None):
        """Rank 0: collect every frame submitted so far (end of a batch / of a timed region)."""
        if self.rank == 0:
            self._collect(self.frame_no - 1, consume)

    def close(self):
        """Unmap / free the store and the flags.  Every rank must be idle (sgl_wait_idle + a barrier) before ANY rank calls
        this: the mappings of the other ranks die with the owner's allocation."""
        lib = self.lib
        if self.rank != 0:
            if self.store:
                lib.sgl_peer_close(self.store)
        else:
            for r, p in enumerate(self.peer_flags):
                if r != 0:
                    lib.sgl_peer_close(C.c_void_p(p))
            if self.store:
                lib.sgl_peer_free(self.store)
        if self.local_flag:
            lib.sgl_peer_free(self.local_flag)
        self.store, self.local_flag, self.peer_flags, self._peer_array = C.c_void_p(), C.c_void_p(), [], None

    def timeouts(self):
        from . import capi
        n = C.c_uint64()
        capi.check(self.lib.sgl_peer_timeouts(C.byref(n)))
        return int(n.value)
