#!/usr/bin/env python
"""Runs the trace player N times on one golden fixture and reports every output that differs from the first run or from
the committed golden (run-to-run determinism of single-frame traces)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np                                       # noqa: E402
import make_golden                                      # noqa: E402
from softglrender_b200 import workloads                 # noqa: E402
from softglrender_b200.scene.trace import read_outputs  # noqa: E402

names = sys.argv[1].split(",")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
work = os.path.join(ROOT, "build", "tests")
os.makedirs(work, exist_ok=True)
for name in names:
    trace, sha = make_golden.build_trace(name, work)
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    first, bad = None, []
    for i in range(n):
        out = os.path.join(work, name + ".loop.out")
        workloads.run_player(workloads.CUDA_PLAYER, trace, out=out, data_dir=work)
        o = read_outputs(out)
        if first is None:
            first = o
            for k in g.files:
                if k == "trace_sha256":
                    continue
                a, b = g[k], o[k]
                if a.dtype == np.uint8:
                    d = np.abs(a.astype(np.int32) - b.astype(np.int32)).max(axis=-1)
                    print(name, k, "vs golden: within1 %.6f max %d" % (float((d <= 1).mean()), int(d.max())))
                else:
                    print(name, k, "vs golden: depth mismatches", int((a.view(np.uint32) != b.view(np.uint32)).sum()))
            continue
        for k in first:
            d = int((first[k].view(np.uint8) != o[k].view(np.uint8)).sum())
            if d:
                bad.append((i, k, d))
    print(name, "runs", n, "differences from the first run:", bad or "none", flush=True)
