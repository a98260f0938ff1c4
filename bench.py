#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: DamagedHelmet PBR+IBL, 1920x1080 MSAA4x, reversed-Z frames/s (and Gfrag/s).

  python bench.py --gpus N --steps K --warmup W            RendererCUDA (this repo)
  python bench.py --impl reference --gpus N ...            the reference's own RendererSoft on the host cores

A step = one full frame of config 2 (shadow pass 512^2 + main pass 1920x1080 MSAA4x: light, axis, floor, helmet,
skybox), submitted through the Renderer API trace exactly as the Viewer submits it.  `value` = frames/s with all
inputs resident in HBM; `e2e` adds, per frame, the pinned-host upload of the frame's draw records/uniform snapshots
and the read-back of the resolved 1920x1080 RGBA8 image into pinned host memory.
N > 1: one process per GPU (torchrun).  Default = frame-parallel (every rank renders its own frame per step -> weak
scaling); `--mgpu tiles` shards ONE frame by screen-tile ownership (strong scaling).  Either way the finished pixels are
gathered into rank 0's HBM inside the timed region: by direct peer stores from the shading kernel over NVLink (default)
or by an NCCL gather (`--gather nccl`).  Timing is device-side (CUDA events), max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DamagedHelmet PBR 1080p MSAA4x frames/s"
WIDTH, HEIGHT = 1920, 1080


def _env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML polled every 5 ms when
    pynvml is importable (a 70 ms timed region still gets ~10 samples), else `nvidia-smi` every 200 ms."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.rows = []          # (sm_mhz, sm_max_mhz, set of reasons)
        self.stop = threading.Event()
        self.idx = gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = get(self.h)
        masks = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.rows.append((sm, mx, {k for k, m in masks.items() if bits & m}))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        r = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
        c = [x.strip() for x in r.stdout.strip().split(",")]
        if len(c) >= 7 and c[0].replace(".", "").isdigit():
            self.rows.append((int(float(c[0])), int(float(c[1])), {self.NAMES[i] for i in range(4) if c[3 + i].lower().startswith("active")}))

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.005 if self.nvml else 0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=5)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        reasons = sorted(set().union(*[r[2] for r in self.rows])) if self.rows else []
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_reference_fps(trace, data_dir, frames, warmup):
    """The reference's own CPU path on this box's host cores: oracle/_ref/ref_player when it travelled with the
    snapshot (kind "reference"), else the CPU restatement (kind "port")."""
    from softglrender_b200 import workloads
    if os.path.exists(workloads.REF_PLAYER):
        binary, kind, cores = workloads.REF_PLAYER, "reference", os.cpu_count()
    else:
        if not os.path.exists(workloads.ORACLE_PLAYER):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restate"], check=True)
        binary, kind, cores = workloads.ORACLE_PLAYER, "port", 1
    t0 = time.time()
    r = workloads.run_player(binary, trace, data_dir=data_dir, frames=frames, warmup=warmup)
    return {"value": 1000.0 / r["ms_median"], "unit": "frames/s", "cores": cores, "kind": kind, "ms_per_frame": r["ms_median"],
            "sample": "%d timed frames of the same config-2 trace (median), %d warm-up, %.0f s wall" % (frames, warmup, time.time() - t0)}


def bind_to_gpu_numa_node(local_rank):
    """Pins this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated
    (first touch then places the pages on that node).  With N ranks reading 8.3 MB frames back per step, buffers on the
    wrong socket make every device->host copy cross the inter-socket link (round 1: e2e weak scaling 0.52 at N = 8)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "reason": "no NUMA information for %s" % bus}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"node": node, "reason": "no allowed CPU on the GPU's node"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as e:
        return {"node": None, "reason": str(e)[:80]}


def _json_only_stdout():
    """The contract is ONE JSON line on stdout: everything else a library may print there (NCCL's version banner, ...) is
    sent to stderr by pointing fd 1 at fd 2; the returned writer is the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    out = _json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (configs 3 and 4 tile-sharded)")
    ap.add_argument("--mgpu", default="frames", choices=["frames", "tiles"],
                    help="N > 1: frame-parallel (weak scaling, default) or one frame sharded by screen tiles (strong scaling)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "dma", "nccl"],
                    help="N > 1: finished pixels reach rank 0 by direct peer stores from the shading kernel (p2p), by copy-engine "
                         "pushes of whole frames into rank 0's store (dma; frame-parallel only), or by an NCCL gather")
    args = ap.parse_args()
    rank, local_rank, world = _env_rank()
    K, W = args.steps, max(args.warmup, 3)

    from softglrender_b200 import workloads
    work = os.path.join(ROOT, "build", "bench")
    config = {"workload": "config2: DamagedHelmet.gltf PBR+IBL, Room.jpeg equirect skybox, 1920x1080 MSAA4x, reversed-Z, "
                          "floor+axis+light+512^2 shadow pass (Config defaults)",
              "resolution": [WIDTH, HEIGHT], "msaa": 4, "draws_per_frame": 6, "passes_per_frame": 2,
              "l2_policy": "inputs larger than L2: ~190 MB touched per frame (96 MB skybox cube + 66 MB MSAA colour/depth + "
                           "20 MB material textures + 8 MB resolve) vs 126 MB L2; no explicit flush",
              "parallelism": "single GPU"}

    # ---------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the IBL maps are inputs: generate them with whatever renderer is available on this box
        ibl_player = workloads.CUDA_PLAYER if _cuda_ok() else (workloads.REF_PLAYER if os.path.exists(workloads.REF_PLAYER) else None)
        trace, data = workloads.build_c2(work, WIDTH, HEIGHT, ibl_player=ibl_player)
        # one step = one whole frame on all host cores (~0.13 s on 16 cores); --steps / --warmup are honoured as long as
        # the run stays within a few minutes (a 2-frame probe sizes it), else the number of timed frames is cut and said so
        probe = cpu_reference_fps(trace, data, 2, 1)
        budget_s = 240.0
        frames = max(3, min(K, int(budget_s / max(probe["ms_per_frame"] / 1e3, 1e-3)) - W))
        cb = cpu_reference_fps(trace, data, frames, W)
        if frames < K:
            cb["sample"] += "; --steps %d cut to %d frames to keep the reference arm within ~%d s" % (K, frames, int(budget_s))
        config["gather"] = "n/a"
        line = {"metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": frames, "warmup": W,
                "ms_per_step": cb["ms_per_frame"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "bundled glTF asset + synthetic camera (Config defaults)", "config": config, "impl": "reference",
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        out.write(json.dumps(line) + "\n")
        out.flush()
        return 0

    # ---------------------------------------------------------------------------------------------- RendererCUDA arm
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from softglrender_b200 import capi, multigpu as M
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; RendererCUDA has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # also at N = 1: a pinned frame buffer that lands on the other socket makes the 8.3 MB read-back cross the inter-socket link
    host_numa = bind_to_gpu_numa_node(local_rank) if not os.environ.get("BENCH_NO_NUMA") else "off (BENCH_NO_NUMA)"
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctl = dist.new_group(backend="gloo")      # control plane (IPC handles); the data plane is NVLink
    capi.init(local_rank, rank, world)
    lib = capi.load()
    stream = torch.cuda.Stream()                   # the library renders on torch's current stream, so NCCL ops and torch
    torch.cuda.set_stream(stream)                  # copies are stream-ordered with the kernels (no host waits per frame)
    capi.check(lib.sgl_set_stream(C.c_void_p(stream.cuda_stream)))
    if rank == 0:
        trace, data = workloads.build_c2(work, WIDTH, HEIGHT)
    if world > 1:
        dist.barrier()
    trace, data = workloads.build_c2(work, WIDTH, HEIGHT)
    player = capi.Player(trace, data)
    player.setup()
    color_handle = player.texture_handle("color")
    nbytes = WIDTH * HEIGHT * 4
    host_img = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    host_all = torch.empty(nbytes, dtype=torch.uint8).pin_memory() if (world > 1 and rank == 0 and args.mgpu == "tiles") else None

    # ---- N > 1: what is sharded and how finished pixels reach rank 0 (softglrender_b200/multigpu.py)
    tiles = world > 1 and args.mgpu == "tiles"
    store = gather_in = gather_out = tile_gather = None
    gather = "none"
    if world > 1:
        if tiles:
            tile_gather = M.TileGather(WIDTH, HEIGHT, rank, world, "interleave")
            tile_gather.install(lib)
        gather = args.gather
        if gather == "dma" and tiles:
            gather = "p2p"
        if gather in ("p2p", "dma"):
            try:
                store = M.PeerFrameStore(lib, WIDTH, HEIGHT, rank, world, frames_per_slot=1 if tiles else world, slots=4,
                                         control_group=ctl, lag=2, dma=(gather == "dma"), timeout_ms=10000)
            except RuntimeError as e:      # CUDA IPC not permitted on this box (all ranks agree): NCCL moves the same bytes
                gather = "nccl (p2p unavailable: %s)" % str(e)[:80]
        if store is None and not tiles:
            ptr, sz = C.c_void_p(), C.c_size_t()
            capi.check(lib.sgl_texture_device_ptr(color_handle, 0, 0, 1, C.byref(ptr), C.byref(sz)))
            gather_in = M.device_view(ptr.value, nbytes)
            gather_out = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
    config["parallelism"] = ("single GPU" if world == 1 else
                             ("sort-first screen-tile ownership x%d (16x16 tiles, 4x4-tile blocks interleaved), geometry replicated" % world
                              if tiles else "frame-parallel x%d (every rank renders its own frame each step)" % world))
    config["gather"] = ("n/a" if world == 1 else
                        (("copy-engine pushes of finished frames into rank 0's HBM over NVLink (CUDA IPC), stream-ordered system-scope flags"
                          if store.dma else
                          "direct peer stores: the shading kernel writes resolved pixels into rank 0's HBM over NVLink (CUDA IPC), "
                          "stream-ordered system-scope flags") if store is not None else
                         "NCCL gather to rank 0 (%s)" % gather))
    step_no = [0]

    def step(e2e):
        i = step_no[0]
        step_no[0] += 1
        if store is not None:
            store.begin_frame(color_handle, 0 if tiles else rank)
        player.frame(sync=False)
        if store is not None:
            consume = None
            if e2e and tiles and rank == 0:     # the assembled frame leaves rank 0's store for the host
                def consume(ptr):
                    host_all.copy_(M.device_view(ptr, nbytes), non_blocking=True)
                    g_d2h[0] += nbytes
            store.end_frame(consume)
        elif tile_gather is not None:
            tile_gather.gather_device(lib, color_handle)
            if e2e and rank == 0:
                capi.check(lib.sgl_texture_readback_async(color_handle, 0, 0, 1, host_img[i & 1].data_ptr(), nbytes))
        elif world > 1:
            dist.gather(gather_in, gather_out, dst=0)
        if e2e and not tiles:                   # every rank delivers its own frame to (shared) host memory, pipelined
            capi.check(lib.sgl_texture_readback_async(color_handle, 0, 0, 1, host_img[i & 1].data_ptr(), nbytes))

    g_d2h = [0]

    def flush_store(e2e):
        if store is None:
            return
        consume = None
        if e2e and tiles and rank == 0:
            def consume(ptr):
                host_all.copy_(M.device_view(ptr, nbytes), non_blocking=True)
                g_d2h[0] += nbytes
        store.flush(consume)

    def sync_all():
        capi.check(lib.sgl_wait_idle())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # warm-up: W steps as asked, and in any case ~0.5 s of frames (clock ramp of an idle GPU, lazy module loading, first
    # touch of the pinned buffers) -- all untimed.  The extra count is agreed on by all ranks BEFORE it runs (the peer
    # frame store counts frames, so every rank must submit the same number).
    t_warm = time.perf_counter()
    for _ in range(W):
        step(False)
    capi.check(lib.sgl_wait_idle())
    per_step = max((time.perf_counter() - t_warm) / W, 1e-5)
    n_extra = min(int(0.5 / per_step) + 1, 20000)
    if world > 1:
        t = torch.tensor([n_extra], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_extra = int(t.item())
    for _ in range(n_extra):
        step(False)
    flush_store(False)
    sync_all()
    # CPU cost of recording + submitting one frame, measured on a short burst that cannot fill the launch queue
    h0 = time.perf_counter()
    for _ in range(8):
        step(False)
    host_submit_ms = (time.perf_counter() - h0) * 1e3 / 8
    flush_store(False)
    sync_all()

    # ---- timed region 1: device-resident throughput (CUDA events on the library's stream), max over ranks.
    #      `value` comes from the FIRST region of exactly K steps; the region is then repeated (untimed for `value`) until
    #      ~0.4 s of frames have run, so that a short K (the driver's --steps 20 is a 6 ms window) still comes with a spread
    #      and with clock samples taken under the same load.
    def timed_region():
        ms = C_float()
        sync_all()
        capi.check(lib.sgl_timer_begin())
        for _ in range(K):
            step(False)
        flush_store(False)
        capi.check(lib.sgl_timer_end(ms))
        sync_all()
        return _max_over_ranks(ms.value, world)

    capi.check(lib.sgl_reset_counters())
    with ClockSampler(local_rank) as clocks:
        elapsed_ms = timed_region()
        ctr = capi.counters()
        n_rep = int(min(max(400.0 / max(elapsed_ms, 1e-3), 2), 12))
        if world > 1:
            t = torch.tensor([n_rep], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_rep = int(t.item())
        reps = sorted(timed_region() / K for _ in range(n_rep))
    clocks_summary = clocks.summary()
    spread = {"repeats_of_K_steps": n_rep, "ms_per_step_min": reps[0], "ms_per_step_median": reps[len(reps) // 2], "ms_per_step_max": reps[-1]}

    # ---- timed region 2: end to end through the public API with host buffers: per frame the draw records / uniform
    #      snapshots are uploaded from pinned memory and the finished frame is read back into pinned host memory
    #      (pipelined: the copy of frame f overlaps the geometry + visibility work of frame f+1)
    #      The read-back path gets its own untimed warm-up first (staging buffers, the first DMA into the pinned frames, lazily
    #      loaded copy kernel): one-off costs of some 10 ms that a 500-step region would otherwise carry as 30 us per step.
    for _ in range(max(3, min(W, 20))):
        step(True)
    flush_store(True)
    capi.check(lib.sgl_readback_wait())
    sync_all()
    g_d2h[0] = 0
    capi.check(lib.sgl_reset_counters())
    sync_all()
    e2e_dev_ms = C_float()
    t0 = time.perf_counter()
    capi.check(lib.sgl_timer_begin())
    for _ in range(K):
        step(True)
    flush_store(True)
    capi.check(lib.sgl_timer_end(e2e_dev_ms))      # blocks until the main stream has drained (the last copy may still run)
    e2e_submit_s = time.perf_counter() - t0
    sync_all()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world)
    ctr2 = capi.counters()
    h2d = _sum_over_ranks(ctr2["h2d_bytes"], world)
    d2h = _sum_over_ranks(ctr2["d2h_bytes"] + g_d2h[0], world)
    timeouts = store.timeouts() if store is not None else 0
    if timeouts:
        raise SystemExit("bench.py: %d peer waits timed out on rank %d" % (timeouts, rank))

    # ---- per-kernel device times for the roofline block (profiling events on; separate from the timed regions)
    capi.check(lib.sgl_set_profiling(1))
    for _ in range(20):
        step(False)
    flush_store(False)
    capi.check(lib.sgl_wait_idle())
    ktimes = capi.kernel_times()
    capi.check(lib.sgl_set_profiling(0))
    sync_all()

    # ---- strong scaling of the configs that are DEFINED as one large frame sharded by screen tiles (BASELINE configs 3
    #      and 4): measured at every N (N = 1: unsharded) so that the driver's per-N lines give the efficiency
    strong = None
    if not args.no_strong:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
            bc = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(bc)
            capi.check(lib.sgl_texture_set_mirror(color_handle, None))
            capi.check(lib.sgl_set_tile_owner_map(None, 0, 0))
            strong = {}
            for key, n_steps in (("c3", 40), ("c4big", 8)):
                r = bc.measure_case(lib, key, rank, world, "p2p" if store is not None else "nccl", steps=n_steps,
                                    work=os.path.join(ROOT, "build", "bench"), control_group=ctl)
                strong[key] = {k: r[k] for k in ("workload", "n_gpus", "units_per_s", "unit", "steps", "ms_per_step", "gfrag_per_s", "parallelism",
                                                 "roofline", "kernel_ms_per_step") if k in r}
        except Exception as e:   # never lose the headline line to the extra block
            strong = {"error": str(e)[:300]}
    frames_per_step = 1 if tiles or world == 1 else world
    fps = frames_per_step * K / (elapsed_ms / 1000.0)
    frags_per_frame = _sum_over_ranks(ctr["fragments_shaded"], world) / float(K * frames_per_step)
    launches = _sum_over_ranks(ctr["kernel_launches"], world)
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "bundled glTF asset + synthetic camera (Config defaults)", "config": config,
            "gfrag_per_s": fps * frags_per_frame / 1e9, "fragments_per_frame": frags_per_frame,
            "e2e": {"value": frames_per_step * K / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d // K,
                    "d2h_bytes_per_step": d2h // K},
            "gpu_launches": launches, "clocks": clocks_summary, "host_submit_ms_per_step": host_submit_ms, "spread": spread, "host_numa": host_numa,
            # CPU work of the library per step (state snapshots + arena layout + graph launches), without the time the host spends
            # blocked because the GPU is the bottleneck (the burst figure above includes that)
            "host_cpu_ms_per_step": (ctr["host_ns_pass_end"] + ctr["host_ns_draw"] - ctr["host_ns_wait_gpu"]) / 1e6 / K,
            "host_blocked_on_gpu_ms_per_step": ctr["host_ns_wait_gpu"] / 1e6 / K,
            # scheduling evidence (rank 0): visibility kernels that started while the previous frame was still being shaded
            "early_vis_per_step": ctr["early_vis"] / float(K), "early_vis_per_step_e2e": ctr2["early_vis"] / float(K),
            # the e2e region seen from the device (events on the rendering stream) and from the submitting thread: tells a
            # GPU-side slowdown by the copies from a host-side one
            "e2e_device_ms_per_step": e2e_dev_ms.value / K, "e2e_host_blocked_ms_per_step": ctr2["host_ns_wait_gpu"] / 1e6 / K,
            "e2e_wall_ms_until_main_stream_drained_per_step": e2e_submit_s * 1e3 / K}
    if strong is not None:
        line["strong_scaling"] = strong
    if rank == 0:
        line.update(roofline_blocks(ktimes, ctr, K, elapsed_ms / K, frames_per_step, tiles, world))
        line["kernel_ms_per_frame"] = {k: v[1] / 20.0 for k, v in sorted(ktimes.items())}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_reference_fps(trace, data, 40, 2)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"error": str(e)[:200]}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _sum_over_ranks(v, world):
    if world == 1:
        return int(v)
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v)], dtype=torch.int64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def C_float():
    import ctypes
    return ctypes.c_float()


def _max_over_ranks(v, world):
    if world == 1:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _offline_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


def roofline_blocks(ktimes, ctr, K, ms_per_step, frames_per_step, tiles, world):
    """Roofline per SURVEY.md 8d / DESIGN.md section 6.  HBM is the bounding resource of this path (no dense contraction);
    algorithmic bytes = what MUST cross HBM once with perfect on-chip reuse:

      B_geom = 64 B x VAO vertices + 4 B x indices of every draw incl. the shadow pass   (counters of this run)
      B_tex  = unique texel bytes the frame's samplers touch: MEASURED with the touched-sector bitmap of the instrumentation
               build (tools/gpu/texel_touch.py -> profiles/r02_texel_touch_c2.json); upper bound = level-0 bytes of the bound maps
      B_out  = 4 B x W x H resolved colour (+ the 512^2 float shadow map written and read once)
      B_ms   = 0: multisample colour / depth / visibility live on chip in the ideal pipeline

    Per kernel: sglShadeKernel must read B_tex and write the resolved colour; sglVisKernel must read one 64-byte record per
    primitive that reaches rasterisation.  Everything else either kernel moves today (per-sample depth, owners, per-sample
    colour) is pipeline-internal traffic and shows up in `traffic` (ncu dram bytes), not in the algorithmic figure.
    `achieved` = algorithmic bytes / the kernel's average duration measured live with CUDA events in this run.
    The dominant kernel (largest share of the step, wait kernels of the multi-GPU gather excluded) is `roofline`."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy figure: kernels are timed alone between events)"
    except Exception:
        peak, src = 6650.0, "B200_PROFILING.md fallback"
    frames = float(K) * (1 if (tiles or world == 1) else 1)      # counters are per rank: K frames each
    px = WIDTH * HEIGHT
    verts = ctr["vertices_in"] / frames
    idx = ctr["indices_in"] / frames
    prims = ctr["primitives_in"] / frames
    b_geom = 64.0 * verts + 4.0 * idx
    touch = _offline_json("r02_texel_touch_c2.json")
    level0 = 6 * 2000 * 2000 * 4 + 5 * 1024 * 1024 * 4 + 524280 + 24576 + 512 * 512 * 4
    if touch and "unique_texel_bytes_per_frame" in touch:
        b_tex, tex_src = float(touch["unique_texel_bytes_per_frame"]), "measured: touched 32-byte sectors (profiles/r02_texel_touch_c2.json, git %s)" % touch.get("git")
    else:
        b_tex, tex_src = float(min(level0, 16.0 * ctr["fragments_shaded"] / frames + 5 * 1024 * 1024 * 4)), "upper bound: min(level-0 bytes of the bound maps, 16 B x fragments)"
    b_out = 4.0 * px + 2 * 512 * 512 * 4
    ncu = _offline_json("ncu_summary.json") or {}

    def block(name, b_alg, what, avg_ms):
        achieved = b_alg / 1e9 / (avg_ms / 1e3)
        blk = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": b_alg, "algorithmic_bytes_are": what, "traffic": None}
        k = (ncu.get("kernels") or {}).get(name.split("<")[0])
        if k:   # offline: one `ncu --set full` capture of the same command, committed under profiles/
            blk["traffic"] = k.get("dram_bytes_per_launch")
            top = sorted((k.get("stall_per_issue") or {}).items(), key=lambda kv: -kv[1])
            top = [t for t in top if t[0] != "selected"][:2]
            blk["offline_ncu"] = {"source": "profiles/ncu_summary.json (%s)" % ncu.get("source"), "issue_active_pct": k.get("issue_active_pct"),
                                  "warps_active_pct": k.get("warps_active_pct"), "top_stalls_per_issue": dict(top),
                                  "registers_per_thread": k.get("registers_per_thread")}
        return blk

    cands = {k: v for k, v in ktimes.items() if not k.startswith("sglPeer")}
    blocks = []
    for name, (launches, total_ms) in sorted(cands.items(), key=lambda kv: -kv[1][1]):
        avg = total_ms / max(launches, 1)
        if name.startswith("sglShade"):
            blocks.append(block(name, b_tex + 4.0 * px, "unique texel sectors in + resolved colour out", avg))
        elif name.startswith("sglVis"):
            blocks.append(block(name, 64.0 * prims, "one 64-byte record per primitive in (depth / owners stay on chip in the ideal)", avg))
        elif name.startswith("sglRaster"):
            blocks.append(block(name, b_tex + 4.0 * px + 64.0 * prims, "primitive records + unique texel sectors in, resolved colour out", avg))
    frame_b = b_geom + b_tex + b_out
    frame_ms = ms_per_step / float(frames_per_step)
    out = {"roofline": blocks[0] if blocks else None,
           "roofline_kernels": blocks[1:3],
           "roofline_frame": {"bound": "hbm", "algorithmic_bytes_per_frame": frame_b, "b_geom": b_geom, "b_tex": b_tex, "b_out": b_out, "b_ms": 0,
                              "b_tex_source": tex_src, "achieved": frame_b / 1e9 / (frame_ms / 1e3), "peak": peak, "unit": "GB/s",
                              "frac": frame_b / 1e9 / (frame_ms / 1e3) / peak, "peak_source": src,
                              "roofline_ms_per_frame": frame_b / 1e9 / peak * 1e3}}
    if out["roofline"]:
        out["roofline"]["peak_source"] = src
        out["roofline"]["note"] = ("HBM is not what binds this path at 1080p (SURVEY 8d: a frame needs %.0f MB, %.1f us at the measured peak): the "
                                   "kernels are latency / issue bound -- see offline_ncu and profiles/README.md" % (frame_b / 1e6, frame_b / 1e9 / peak * 1e6))
    return out


if __name__ == "__main__":
    sys.exit(main())
