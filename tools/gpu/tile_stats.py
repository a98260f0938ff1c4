#!/usr/bin/env python
"""Per-tile primitive list lengths of config 2's main pass (load-balance instrumentation)."""
import ctypes as C
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softglrender_b200 import capi, workloads   # noqa: E402
capi.init(0)
lib = capi.load()
trace, data = workloads.build_c2(os.path.join(ROOT, "build", "bench"), 1920, 1080)
p = capi.Player(trace, data)
p.setup()
p.frame(sync=True)
tx, ty = C.c_int(), C.c_int()
capi.check(lib.sgl_get_tile_list_sizes(None, 0, C.byref(tx), C.byref(ty)))
a = np.zeros(tx.value * ty.value, np.uint32)
capi.check(lib.sgl_get_tile_list_sizes(a.ctypes.data, a.size, None, None))
a = a.reshape(ty.value, tx.value)
print("tiles", a.shape, "prepared (heavy) lists: sum", int(a[a != 0xFFFFFFFF].sum()), "; tiles without a prepared list (light, or overflow):", int((a == 0xFFFFFFFF).sum()))
a = np.where(a == 0xFFFFFFFF, 0, a)
v = np.sort(a[a != 0xFFFFFFFF].ravel())[::-1]
print("top 30:", v[:30].tolist())
print("percentiles 50/90/99/99.9:", [int(np.percentile(v, q)) for q in (50, 90, 99, 99.9)])
print("classes >=192/48/12/rest:", int((v >= 192).sum()), int(((v >= 48) & (v < 192)).sum()), int(((v >= 12) & (v < 48)).sum()), int((v < 12).sum()))
ys, xs = np.unravel_index(np.argsort(a.ravel())[::-1][:10], a.shape)
print("heaviest tiles (tx,ty,n):", [(int(x), int(y), int(a[y, x])) for x, y in zip(xs, ys)])
# per-tile duration of the visibility kernel
capi.check(lib.sgl_debug_tile_times(1, None, 0))
for _ in range(3):
    p.frame(sync=True)
t = np.zeros((ty.value * tx.value, 2), np.uint64)
capi.check(lib.sgl_debug_tile_times(0, t.ctypes.data, t.shape[0]))
t0 = t[:, 0].min()
start = (t[:, 0] - t0).astype(np.float64) / 1e3
dur = (t[:, 1] - t[:, 0]).astype(np.float64) / 1e3
end = start + dur
print("vis kernel span %.1f us; tile duration us: median %.1f p90 %.1f p99 %.1f max %.1f" % (end.max(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
order = np.argsort(dur)[::-1][:12]
print("slowest tiles (tx,ty,n,start,dur):", [(int(i % tx.value), int(i // tx.value), int(a.ravel()[i]), round(float(start[i]), 1), round(float(dur[i]), 1)) for i in order])
last = np.argsort(end)[::-1][:8]
print("last to finish (tx,ty,n,start,dur,end):", [(int(i % tx.value), int(i // tx.value), int(a.ravel()[i]), round(float(start[i]), 1), round(float(dur[i]), 1), round(float(end[i]), 1)) for i in last])
for lo, hi in ((0, 3), (3, 12), (12, 48), (48, 192), (192, 10000)):
    m = (a.ravel() >= lo) & (a.ravel() < hi)
    if m.any():
        print("  n in [%d,%d): %d tiles, mean dur %.1f us, sum %.0f us" % (lo, hi, m.sum(), dur[m].mean(), dur[m].sum()))
# visibility-kernel span in steady state (frames submitted back to back, geometry of the next frame overlapping)
capi.check(lib.sgl_debug_tile_times(1, None, 0))
for _ in range(40):
    p.frame(sync=False)
capi.check(lib.sgl_wait_idle())
capi.check(lib.sgl_debug_tile_times(0, t.ctypes.data, t.shape[0]))
print("steady-state vis kernel span: %.1f us" % ((t[:, 1].max() - t[:, 0].min()) / 1e3))
