#!/usr/bin/env python
"""Where config 2's frame time goes: kernel times of scene variants (skybox / axis / floor / model removed one at a time)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softglrender_b200 import capi, workloads          # noqa: E402
from softglrender_b200.scene import scenes            # noqa: E402


def main():
    capi.init(0)
    lib = capi.load()
    work = os.path.join(ROOT, "build", "bench")
    ibl_files = workloads.build_ibl(work)              # makes the IBL maps
    ad = workloads.assets_dir()
    variants = {"full": {}, "no_skybox": dict(show_skybox=False), "no_axis": dict(world_axis=False), "no_floor": dict(show_floor=False),
                "no_light_point": dict(show_light=False), "cube_model": dict(_model="Cube"), "no_shadow": dict(shadow_map=False)}
    for name, cfg in variants.items():
        cfg = dict(cfg)
        model = cfg.pop("_model", "DamagedHelmet")
        trace = os.path.join(work, "c2var_%s.sglt" % name)
        scenes.config2_helmet(ad, 1920, 1080, ibl_files=ibl_files, model=model, **cfg).save(trace)
        p = capi.Player(trace, work)
        p.setup()
        for _ in range(5):
            p.frame(sync=False)
        capi.check(lib.sgl_wait_idle())
        ms = C.c_float()
        capi.check(lib.sgl_reset_counters())
        capi.check(lib.sgl_timer_begin())
        for _ in range(100):
            p.frame(sync=False)
        capi.check(lib.sgl_timer_end(ms))
        ctr = capi.counters()
        capi.check(lib.sgl_set_profiling(1))
        for _ in range(10):
            p.frame(sync=False)
        capi.check(lib.sgl_wait_idle())
        kt = capi.kernel_times()
        capi.check(lib.sgl_set_profiling(0))
        p.close()
        os.remove(trace)
        print("%-15s %.3f ms/frame  frags %8d binned %7d | " % (name, ms.value / 100, ctr["fragments_shaded"] / 100, ctr["primitives_binned"] / 100) +
              " ".join("%s=%.0f" % (k.replace("sgl", "").replace("Kernel", ""), v[1] / 10 * 1e3) for k, v in sorted(kt.items())), flush=True)


if __name__ == "__main__":
    main()
