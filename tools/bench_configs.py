#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 1, 3, 4, 5) on one GPU: frames/s (views/s for config 5) with everything
resident, CUDA events on the library's stream; optionally the compiled reference on the host cores beside it.
bench.py stays the contract benchmark (config 2); this tool feeds DESIGN.md / profiles/.

  python tools/bench_configs.py [--cpu] [--only c1,c3,c4,c4big,c5] [--out profiles/r01_configs.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case_table(work, rank, world):
    from softglrender_b200 import workloads
    n_views = 64 * max(world, 1)
    views = list(range(0, 2048, 2048 // n_views))
    my_views = views[rank::world] if world > 1 else views
    return {
        "c1": ("config1: Cube Blinn-Phong 1000x800 no AA + shadow pass", lambda: workloads.build_c1(work), 1, 200, None),
        "c3": ("config3: BoomBox+GlassTable 3840x2160 shadow + blend + FXAA", lambda: workloads.build_c3(work), 1, 100, "stripes"),
        "c4": ("config4 (scaled): 100k-triangle soup 1920x1080, 8 mip-mapped 1024^2 textures", lambda: workloads.build_c4(work), 1, 100, "interleave"),
        "c4big": ("config4 (scaled): 2M-triangle soup 7680x4320, 8 mip-mapped 2048^2 textures",
                  lambda: workloads.build_c4(work, n_tris=2000000, width=7680, height=4320, tex_size=2048), 1, 10, "interleave"),
        "c4full": ("config4 (full size): 10M-triangle soup 7680x4320, 8 mip-mapped Morton 4096^2 textures",
                   lambda: workloads.build_c4(work, n_tris=10000000, width=7680, height=4320, tex_size=4096), 1, 5, "interleave"),
        "c5": ("config5: %d views of AfricanHead 512x512 per step%s" % (len(views), "" if world == 1 else ", dealt round-robin to %d ranks" % world),
               lambda: workloads.build_c5(work, "AfricanHead", my_views), len(views), 20, None),
    }


def measure_case(lib, key, rank, world, gather_mode="nccl", steps=None, cpu=False, keep_trace=False, work=None, control_group=None, as_rank=None):
    """One secondary workload on `world` GPUs (torch.distributed already initialised by the caller when world > 1; the
    library must run on torch's current stream).  c3 / c4*: ONE frame sharded by screen tiles, owned tiles gathered to rank 0
    inside the timed region (strong scaling); c5: views dealt to the ranks.  Returns the result dict (same on every rank)."""
    import time
    from softglrender_b200 import capi, multigpu as M, workloads
    torch = dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
    work = work or os.path.join(ROOT, "build", "bench_configs")
    name, builder, units, default_steps, shard = case_table(work, rank, world)[key]
    steps = steps or default_steps
    if key.startswith("c4"):
        os.environ["SGL_TEXTURE_LAYOUT"] = "2"       # Morton 32x32 texture storage (BASELINE config 4)
    if key == "c3":
        os.environ["SGL_SHARD_HALO"] = "32"          # FXAA input: 32-px halo around owned tiles
    if world > 1 and key != "c5":
        if rank == 0:
            builder()                                 # one rank writes the cached trace
        dist.barrier()
    trace, data = builder()
    p = capi.Player(trace, data)
    p.setup()
    tex = p.texture_handle("color") if key != "c5" else p.texture_handle("color_v0")
    gather = store = None
    if as_rank and world == 1 and shard:
        # one GPU renders the share of rank r of n (no exchange): what one rank of a tile-sharded run does, without n GPUs
        w_, h_ = C.c_int(), C.c_int()
        capi.check(lib.sgl_texture_level_size(tex, 0, C.byref(w_), C.byref(h_)))
        capi.check(lib.sgl_set_rank(as_rank[0], as_rank[1]))
        gather = M.TileGather(w_.value, h_.value, as_rank[0], as_rank[1], shard)
        gather.install(lib)
        gather_mode = "none"
    if world > 1 and shard:
        w_, h_ = C.c_int(), C.c_int()
        capi.check(lib.sgl_texture_level_size(tex, 0, C.byref(w_), C.byref(h_)))
        gather = M.TileGather(w_.value, h_.value, rank, world, shard)
        gather.install(lib)
        if gather_mode == "p2p":
            # owned pixels are stored straight into rank 0's HBM by the kernel that produces them (multigpu.PeerFrameStore);
            # rank 0 collects frame f - 1 while frame f renders
            try:
                store = M.PeerFrameStore(lib, w_.value, h_.value, rank, world, frames_per_slot=1, slots=3, control_group=control_group,
                                         lag=1, timeout_ms=20000)
            except RuntimeError:      # CUDA IPC not permitted here (all ranks agree): NCCL moves the same bytes
                gather_mode = "nccl"

    def step():
        if store is not None:
            store.begin_frame(tex, 0)
        p.frame(sync=False)
        if store is not None:
            store.end_frame()
        elif gather is not None and gather_mode == "nccl":
            gather.gather_device(lib, tex)

    def sync():
        if store is not None:
            store.flush()
        capi.check(lib.sgl_wait_idle())
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    for _ in range(3):
        step()
    sync()
    h0 = time.perf_counter()
    burst = 4 if units > 1 else 8
    for _ in range(burst):
        step()
    host_ms = (time.perf_counter() - h0) * 1e3 / burst
    sync()
    capi.check(lib.sgl_reset_counters())
    ms = C.c_float()
    capi.check(lib.sgl_timer_begin())
    for _ in range(steps):
        step()
    if store is not None:
        store.flush()
    capi.check(lib.sgl_timer_end(ms))
    sync()
    elapsed = ms.value
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    ctr = capi.counters()
    capi.check(lib.sgl_set_profiling(1))
    for _ in range(3):
        step()
    if store is not None:
        store.flush()
    capi.check(lib.sgl_wait_idle())
    kt = capi.kernel_times()
    capi.check(lib.sgl_set_profiling(0))
    sync()
    if store is not None:
        if store.timeouts():
            raise RuntimeError("bench_configs: %d peer waits timed out on rank %d" % (store.timeouts(), rank))
        capi.check(lib.sgl_texture_set_mirror(tex, None))
        if rank != 0:                   # mappings go before the owner's allocation
            store.close()
        dist.barrier()
        if rank == 0:
            store.close()
    if gather is not None:
        capi.check(lib.sgl_set_tile_owner_map(None, 0, 0))
    if as_rank and world == 1:
        capi.check(lib.sgl_set_rank(0, 1))
    p.close()
    os.environ.pop("SGL_TEXTURE_LAYOUT", None)
    os.environ.pop("SGL_SHARD_HALO", None)
    frags = ctr["fragments_shaded"]
    if world > 1:
        t = torch.tensor([frags], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        frags = int(t.item())
    r = {"workload": name, "n_gpus": world, "units_per_s": units * steps / (elapsed / 1e3), "unit": "views/s" if key == "c5" else "frames/s",
         "steps": steps, "ms_per_step": elapsed / steps, "host_submit_ms_per_step": host_ms, "fragments_per_step": frags / steps,
         "gfrag_per_s": frags / (elapsed / 1e3) / 1e9, "primitives_per_step": ctr["primitives_in"] / steps, "bin_entries_per_step": ctr["primitives_binned"] / steps,
         "clip_overflow": ctr["clip_overflow"], "bin_spills": ctr["bin_spills"], "host_us_pass_end_per_step": ctr["host_ns_pass_end"] / 1e3 / steps,
         "host_us_draw_per_step": ctr["host_ns_draw"] / 1e3 / steps, "passes_per_step": ctr["passes"] / steps, "draws_per_step": ctr["draws"] / steps,
         "kernel_ms_per_step": {k: v[1] / 3.0 for k, v in sorted(kt.items())}}
    if as_rank and world == 1 and shard:
        r["parallelism"] = "one GPU rendering the tiles of rank %d of %d (%s), no exchange" % (as_rank[0], as_rank[1], shard)
    if world > 1:
        r["parallelism"] = ("view-parallel, no exchange (every view stays in the HBM of the rank that rendered it)" if key == "c5" else
                            "one frame sharded by screen tiles (%s), geometry replicated, owned tiles gathered to rank 0 by %s inside the timed region"
                            % (shard, {"nccl": "pack -> NCCL gather -> unpack", "p2p": "direct peer stores from the shading kernel into rank 0's HBM (CUDA IPC over NVLink)"}
                                        .get(gather_mode, "nothing (gather off)")))
    if key.startswith("c4"):
        # B_geom = 64 B x vertices + 4 B x indices, B_out = 4 B x W x H; B_tex left out (lower bound)
        b = 64.0 * ctr["vertices_in"] / steps + 4.0 * ctr["indices_in"] / steps + 4.0 * (7680 * 4320 if key != "c4" else 1920 * 1080)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            peak = 6650.0
        r["roofline"] = {"bound": "hbm", "algorithmic_bytes_per_frame_lower_bound": b, "achieved_GBps": b / 1e9 / (elapsed / steps / 1e3),
                         "peak_GBps": peak * max(world, 1), "frac": b / 1e9 / (elapsed / steps / 1e3) / (peak * max(world, 1)),
                         "roofline_ms_per_frame": b / 1e9 / (peak * max(world, 1)) * 1e3}
    if cpu and world == 1 and os.path.exists(workloads.REF_PLAYER) and key not in ("c4big", "c4full"):
        c = workloads.run_player(workloads.REF_PLAYER, trace, data_dir=data, frames=5 if key != "c5" else 2, warmup=1)
        r["cpu_reference"] = {"units_per_s": units * 1000.0 / c["ms_median"], "ms_per_step": c["ms_median"], "cores": os.cpu_count()}
    if world > 1:
        dist.barrier()
    if key in ("c4big", "c4full", "c3") and rank == 0 and not keep_trace:
        os.remove(trace)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time oracle/_ref/ref_player (all host cores)")
    ap.add_argument("--only", default="c1,c3,c4,c4big,c5")
    ap.add_argument("--out", default="")
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p", "none"], help="N > 1, tile-sharded configs: how owned tiles reach rank 0")
    ap.add_argument("--as-rank", default="", help="R/N on one GPU: render only the tiles rank R of N owns (what one rank of a sharded run does)")
    args = ap.parse_args()
    as_rank = tuple(int(x) for x in args.as_rank.split("/")) if args.as_rank else None
    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    from softglrender_b200 import capi
    dist = None
    if world > 1:      # torchrun
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctl = dist.new_group(backend="gloo") if world > 1 else None      # control plane (IPC handles)
    capi.init(local, rank, world)
    lib = capi.load()
    if world > 1:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        capi.check(lib.sgl_set_stream(C.c_void_p(stream.cuda_stream)))
    results = {}
    for key in args.only.split(","):
        r = measure_case(lib, key, rank, world, args.gather, cpu=args.cpu, control_group=ctl, as_rank=as_rank)
        results[key] = r
        if rank == 0:
            print(key, json.dumps(r), flush=True)
    if args.out and rank == 0:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(results, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
