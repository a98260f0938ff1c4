#!/usr/bin/env python
"""Unique texel bytes one frame needs (B_tex of SURVEY 8d's roofline), MEASURED with the touched-sector bitmap of the
instrumentation build:

    python -m softglrender_b200.build --variant touch -DSGL_TOUCH_BITMAP
    SGL_LIB_DIR=$PWD/softglrender_b200/lib_variants/touch python tools/gpu/texel_touch.py [c2|c3|c4] > profiles/r02_texel_touch_c2.json

Every texel load of a sampler marks its 32-byte DRAM sector; after one steady-state frame the set bits are counted per
texture.  bench.py reads the committed JSON for `roofline.algorithmic_bytes` (it is a property of the workload, not of a run)."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softglrender_b200 import capi, workloads          # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    work = os.path.join(ROOT, "build", "bench")
    if which == "c2":
        trace, data = workloads.build_c2(work, 1920, 1080)
        name = "config2 DamagedHelmet 1920x1080 MSAA4x"
    elif which == "c3":
        trace, data = workloads.build_c3(work)
        name = "config3 BoomBox+GlassTable 3840x2160 FXAA"
    else:
        trace, data = workloads.build_c4(work, n_tris=2000000, width=7680, height=4320, tex_size=2048)
        name = "config4 (scaled) 2M-triangle soup 7680x4320"
    capi.init(0)
    lib = capi.load()
    p = capi.Player(trace, data)
    p.setup()
    p.frame(sync=True)
    n = C.c_ulonglong()
    capi.check(lib.sgl_debug_texel_touch(0, 1, C.byref(n)))       # reset after set-up + warm-up frame
    p.frame(sync=True)
    per = {}
    total = 0
    for h in range(1, 512):
        rc = lib.sgl_debug_texel_touch(h, 0, C.byref(n))
        if rc != 0:
            break
        if n.value:
            per[str(h)] = int(n.value)
            total += int(n.value)
    sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, stdout=subprocess.PIPE, text=True).stdout.strip()
    print(json.dumps({"workload": name, "trace": os.path.basename(trace), "unique_texel_bytes_per_frame": total,
                      "sector_bytes": 32, "per_texture_handle": per, "git": sha,
                      "how": "touched-sector bitmap of the -DSGL_TOUCH_BITMAP build, one steady-state frame"}))
    p.close()


if __name__ == "__main__":
    main()
