#!/usr/bin/env python
"""Kernel timeline of config 2 in steady state with the geometry / pixel stage overlap ON (SGL_PROFILE_OVERLAP=1)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["SGL_PROFILE_OVERLAP"] = "1"
from softglrender_b200 import capi, workloads   # noqa: E402
capi.init(0)
lib = capi.load()
trace, data = workloads.build_c2(os.path.join(ROOT, "build", "bench"), 1920, 1080)
p = capi.Player(trace, data)
p.setup()
for _ in range(30):
    p.frame(sync=False)
capi.check(lib.sgl_set_profiling(1))
for _ in range(4):
    p.frame(sync=False)
capi.check(lib.sgl_wait_idle())
capi.kernel_times()
