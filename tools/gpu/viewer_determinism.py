#!/usr/bin/env python
"""Run-to-run determinism of RendererCUDA behind the reference's own Viewer (integration/_build/viewer_headless): the same
scene rendered N times per scheduling variant; every output must be byte-identical to the first run of the first variant."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np                                      # noqa: E402
from test_viewer_integration import run_viewer          # noqa: E402
from softglrender_b200.scene.trace import read_outputs  # noqa: E402

work = os.path.join(ROOT, "build", "tests", "integ")
os.makedirs(work, exist_ok=True)
cases = {"c2": ("--model", "DamagedHelmet", "--skybox", "Room", "--ibl", "--aa", "msaa", "--reverse-z", "--width", 960, "--height", 540),
         "c1": ("--model", "Cube", "--blinnphong", "--width", 1000, "--height", 800)}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variants = [{}, {"SGL_NO_EARLY_VIS": "1"}, {"SGL_NO_RENAME": "1"}, {"SGL_NO_EARLY_VIS": "1", "SGL_NO_RENAME": "1"}]
for name, args in cases.items():
    first = None
    for env in variants:
        for k in ("SGL_NO_EARLY_VIS", "SGL_NO_RENAME"):
            os.environ.pop(k, None)
        os.environ.update(env)
        bad = {}
        for i in range(n):
            out = os.path.join(work, "det_%s.out" % name)
            run_viewer(work, "cuda", out, *args, "--frames", 3)
            o = read_outputs(out)
            if first is None:
                first = o
                continue
            for tag in first:
                a, b = first[tag], o[tag]
                d = int((a.view(np.uint8) != b.view(np.uint8)).sum()) if a.shape == b.shape else -1
                if d:
                    bad.setdefault(tag, []).append(d)
        print(name, env or "default", "runs", n, "differences from the first run:", bad or "none", flush=True)
