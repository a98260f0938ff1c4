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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time oracle/_ref/ref_player (all host cores)")
    ap.add_argument("--only", default="c1,c3,c4,c4big,c5")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from softglrender_b200 import capi, workloads
    capi.init(0)
    lib = capi.load()
    work = os.path.join(ROOT, "build", "bench_configs")
    views = list(range(0, 2048, 32))
    cases = {
        "c1": ("config1: Cube Blinn-Phong 1000x800 no AA + shadow pass", lambda: workloads.build_c1(work), 1, 200),
        "c3": ("config3: BoomBox+GlassTable 3840x2160 shadow + blend + FXAA", lambda: workloads.build_c3(work), 1, 100),
        "c4": ("config4 (scaled): 100k-triangle soup 1920x1080, 8 mip-mapped 1024^2 textures", lambda: workloads.build_c4(work), 1, 100),
        "c4big": ("config4 (scaled): 2M-triangle soup 7680x4320, 8 mip-mapped 2048^2 textures",
                  lambda: workloads.build_c4(work, n_tris=2000000, width=7680, height=4320, tex_size=2048), 1, 10),
        "c5": ("config5: 64 views of AfricanHead 512x512 per step", lambda: workloads.build_c5(work, "AfricanHead", views), len(views), 20),
    }
    results = {}
    for key in args.only.split(","):
        name, builder, units, steps = cases[key]
        trace, data = builder()
        p = capi.Player(trace, data)
        p.setup()
        for _ in range(3):
            p.frame(sync=False)
        capi.check(lib.sgl_wait_idle())
        import time
        h0 = time.perf_counter()
        burst = 4 if units > 1 else 8
        for _ in range(burst):
            p.frame(sync=False)
        host_ms = (time.perf_counter() - h0) * 1e3 / burst
        capi.check(lib.sgl_wait_idle())
        capi.check(lib.sgl_reset_counters())
        ms = C.c_float()
        capi.check(lib.sgl_timer_begin())
        for _ in range(steps):
            p.frame(sync=False)
        capi.check(lib.sgl_timer_end(ms))
        ctr = capi.counters()
        capi.check(lib.sgl_set_profiling(1))
        for _ in range(3):
            p.frame(sync=False)
        capi.check(lib.sgl_wait_idle())
        kt = capi.kernel_times()
        capi.check(lib.sgl_set_profiling(0))
        p.close()
        r = {"workload": name, "units_per_s": units * steps / (ms.value / 1e3), "unit": "views/s" if key == "c5" else "frames/s",
             "ms_per_step": ms.value / steps, "host_submit_ms_per_step": host_ms, "fragments_per_step": ctr["fragments_shaded"] / steps,
             "gfrag_per_s": ctr["fragments_shaded"] / (ms.value / 1e3) / 1e9, "primitives_per_step": ctr["primitives_in"] / steps,
             "clip_overflow": ctr["clip_overflow"], "host_us_pass_end_per_step": ctr["host_ns_pass_end"] / 1e3 / steps,
             "host_us_draw_per_step": ctr["host_ns_draw"] / 1e3 / steps, "passes_per_step": ctr["passes"] / steps, "draws_per_step": ctr["draws"] / steps, "kernel_ms_per_step": {k: v[1] / 3.0 for k, v in sorted(kt.items())}}
        if args.cpu and os.path.exists(workloads.REF_PLAYER) and key != "c4big":
            c = workloads.run_player(workloads.REF_PLAYER, trace, data_dir=data, frames=5 if key != "c5" else 2, warmup=1)
            r["cpu_reference"] = {"units_per_s": units * 1000.0 / c["ms_median"], "ms_per_step": c["ms_median"], "cores": os.cpu_count()}
        results[key] = r
        print(key, json.dumps(r), flush=True)
        if key in ("c4big", "c3"):
            os.remove(trace)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
