#!/usr/bin/env python
"""Kernel timeline of config 2 in steady state (stage overlap ON) without and with the per-frame asynchronous read-back:
where the frame grows when the finished image also travels to the host."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["SGL_PROFILE_OVERLAP"] = "1"
import torch                                     # noqa: E402
from softglrender_b200 import capi, workloads   # noqa: E402
capi.init(0)
lib = capi.load()
trace, data = workloads.build_c2(os.path.join(ROOT, "build", "bench"), 1920, 1080)
p = capi.Player(trace, data)
p.setup()
h = p.texture_handle("color")
n = 1920 * 1080 * 4
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
for rb in (0, 1):
    sys.stderr.write("[sgl timeline] ---- read-back %d\n" % rb)
    for i in range(40):
        p.frame(sync=False)
        if rb:
            capi.check(lib.sgl_texture_readback_async(h, 0, 0, 1, host[i & 1].data_ptr(), n))
    capi.check(lib.sgl_set_profiling(1))
    for i in range(4):
        p.frame(sync=False)
        if rb:
            capi.check(lib.sgl_texture_readback_async(h, 0, 0, 1, host[i & 1].data_ptr(), n))
    capi.check(lib.sgl_wait_idle())
    capi.kernel_times()
    capi.check(lib.sgl_set_profiling(0))
