#!/usr/bin/env python
"""Which CTA processed which work item when (visibility kernel): gaps between consecutive items of a CTA."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softglrender_b200 import capi, workloads   # noqa: E402
os.environ["SGL_DEBUG_LAUNCH"] = "1"
capi.init(0)
lib = capi.load()
trace, data = workloads.build_c2(os.path.join(ROOT, "build", "bench"), 1920, 1080)
p = capi.Player(trace, data)
p.setup()
p.frame(sync=True)
nt = 68 * 120 if False else None
tx, ty = C.c_int(), C.c_int()
capi.check(lib.sgl_get_tile_list_sizes(None, 0, C.byref(tx), C.byref(ty)))
n = tx.value * ty.value
capi.check(lib.sgl_debug_tile_times(1, None, 0))
for _ in range(3):
    p.frame(sync=True)
t = np.zeros((2 * n, 2), np.uint64)
capi.check(lib.sgl_debug_tile_times(0, t.ctypes.data, 2 * n))
times, who = t[:n], t[n:]
t0 = times[:, 0].min()
start = (times[:, 0] - t0).astype(np.float64) / 1e3
end = (times[:, 1] - t0).astype(np.float64) / 1e3
cta = who[:, 0].astype(np.int64)
print("CTAs seen:", len(np.unique(cta)), "span %.1f us" % end.max())
gaps, busy = [], []
for c in np.unique(cta)[:2000]:
    m = np.where(cta == c)[0]
    o = m[np.argsort(start[m])]
    s, e = start[o], end[o]
    gaps += list(s[1:] - e[:-1])
    busy.append((e - s).sum())
gaps = np.array(gaps)
print("items per CTA: mean %.1f; busy per CTA mean %.1f us; gap between items: median %.2f mean %.2f p90 %.2f max %.2f us" % (
    n / max(len(np.unique(cta)), 1), np.mean(busy), np.median(gaps), gaps.mean(), np.percentile(gaps, 90), gaps.max()))
c = np.unique(cta)[5]
m = np.where(cta == c)[0]
o = m[np.argsort(start[m])]
print("CTA", c, [(round(float(start[i]), 1), round(float(end[i] - start[i]), 1)) for i in o][:20])
