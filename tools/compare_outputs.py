"""Compare two trace-player output files (reference vs candidate): per-tag parity statistics."""
import sys
import os
import json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softglrender_b200.scene.trace import read_outputs


def compare(ref_path, out_path):
    a, b = read_outputs(ref_path), read_outputs(out_path)
    res = {}
    for k in a:
        if k not in b:
            res[k] = "missing"
            continue
        x, y = a[k], b[k]
        if x.shape != y.shape:
            res[k] = "shape %s vs %s" % (x.shape, y.shape)
            continue
        if x.dtype == np.uint8:
            d = np.abs(x.astype(np.int32) - y.astype(np.int32)).max(axis=-1)
            res[k] = dict(kind="color", shape=list(x.shape), exact=float((d == 0).mean()), within1=float((d <= 1).mean()),
                          max=int(d.max()), n_gt1=int((d > 1).sum()))
        else:
            eq = x.view(np.uint32) == y.view(np.uint32)
            res[k] = dict(kind="depth", shape=list(x.shape), bit_exact=float(eq.mean()), mismatches=int((~eq).sum()),
                          max_abs=float(np.nanmax(np.abs(x - y))) if (~eq).any() else 0.0)
    return res


if __name__ == "__main__":
    r = compare(sys.argv[1], sys.argv[2])
    for k, v in r.items():
        print(k, json.dumps(v))
