"""Unit-level known-answer vectors produced by the REFERENCE (oracle/_ref/ref_kat, built from /root/reference by
oracle/Makefile): image layouts (Base/Buffer.h:141-213), the sampler (SamplerSoft.h:118-373) over filter x wrap x border
x NPOT x offsets x cube, and barycentric / z / 1/w (RendererSoft.cpp:771-797,1021-1056).

    python tests/golden/make_unit_kats.py        # needs oracle/_ref/ref_kat; writes tests/golden/unit_kats.npz

The inputs are seeded here and stored next to the reference's answers, so the tests need neither this script nor the
reference.  `run_kat(binary, ...)` is also what the tests use to push the same inputs through oracle/_build/oracle_kat."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_KAT = os.path.join(ROOT, "oracle", "_ref", "ref_kat")
OUT = os.path.join(HERE, "unit_kats.npz")

LAYOUT_SIZES = [(5, 3), (37, 21), (64, 64), (100, 33)]
N_COORDS = 48


def run_sample(binary, texels, fmt, layers, mips, filt, wrap, border, coords, lod, offsets=None):
    h, w = texels.shape[-3], texels.shape[-2]
    n = len(lod)
    hdr = struct.pack("<10i", w, h, layers, fmt, int(mips), filt, wrap, border, n, 0 if offsets is None else 1)
    offs = np.zeros((n, 2), np.int32) if offsets is None else offsets.astype(np.int32)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(hdr + texels.tobytes() + coords.astype(np.float32).tobytes() + lod.astype(np.float32).tobytes() + offs.tobytes())
        subprocess.run([binary, "sample", fin, fout], check=True)
        return np.fromfile(fout, np.uint32)


def run_bary(binary, tris, samples):
    n_tri, n_per = tris.shape[0], samples.shape[1]
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<2i", n_tri, n_per))
            for t in range(n_tri):
                f.write(tris[t].astype(np.float32).tobytes() + samples[t].astype(np.float32).tobytes())
        subprocess.run([binary, "bary", fin, fout], check=True)
        raw = np.fromfile(fout, np.uint32).reshape(n_tri, n_per, 6)
    return raw[..., 0].astype(np.int32), raw[..., 1:4].copy().view(np.float32), raw[..., 4:6].copy().view(np.float32)


def run_layout(binary, w, h):
    with tempfile.TemporaryDirectory() as d:
        fout = os.path.join(d, "out.bin")
        subprocess.run([binary, "layout", str(w), str(h), fout], check=True)
        raw = np.fromfile(fout, np.uint32)
    n_t = int(raw[0])
    tiled = raw[2:2 + n_t]
    n_m = int(raw[2 + n_t])
    morton = raw[4 + n_t:4 + n_t + n_m]
    return tiled, morton


def sample_inputs():
    """name -> dict(texels, fmt, layers, mips, coords, lod, offsets)"""
    rng = np.random.RandomState(20261017)
    cases = {}
    for name, (w, h) in (("pot16", (16, 16)), ("npot48", (48, 48)), ("odd33x17", (33, 17))):
        texels = rng.randint(0, 256, (1, h, w, 4)).astype(np.uint8)
        coords = rng.uniform(-1.6, 2.6, (N_COORDS, 2)).astype(np.float32)
        coords[:6] = [[0, 0], [1, 1], [0.5, 0.5], [1.0 / w, 1.0 / h], [-1.0 / w, 1 - 0.5 / h], [0.999999, 1e-7]]
        lod = rng.uniform(-0.7, 6.5, N_COORDS).astype(np.float32)
        lod[:4] = [0, 0.5, 1.0, 1.5]
        offsets = rng.randint(-2, 3, (N_COORDS, 2)).astype(np.int32)
        cases[name] = dict(texels=texels, fmt=0, layers=1, mips=True, coords=coords, lod=lod, offsets=offsets)
    # float texture (shadow map: NEAREST + CLAMP_TO_BORDER + texel offsets, BlinnPhongSoft.h:165-189)
    cases["f32_16"] = dict(texels=rng.rand(1, 16, 16).astype(np.float32).view(np.uint8).reshape(1, 16, 16, 4), fmt=1, layers=1, mips=False,
                           coords=rng.uniform(-0.3, 1.3, (N_COORDS, 2)).astype(np.float32), lod=np.zeros(N_COORDS, np.float32),
                           offsets=rng.randint(-1, 2, (N_COORDS, 2)).astype(np.int32))
    # cube map with mips (prefilter map: LINEAR_MIPMAP_LINEAR at lod = roughness * 4, PbrSoft.h:302-303)
    dirs = rng.normal(size=(N_COORDS, 3)).astype(np.float32)
    dirs[:6] = [[1, 1, 1], [-1, 1, 1], [1, -1, -1], [0, 0, 1], [0.5, 0.5, -0.5], [-2, 2, 0.1]]   # ties of the if-chain
    cases["cube8"] = dict(texels=rng.randint(0, 256, (6, 8, 8, 4)).astype(np.uint8), fmt=0, layers=6, mips=True, coords=dirs,
                          lod=rng.uniform(0, 3.5, N_COORDS).astype(np.float32), offsets=None)
    return cases


def sample_combos(name):
    if name == "f32_16":
        return [(f, w, b) for f in (0, 1) for w in (2, 3) for b in (0, 1)]
    if name == "cube8":
        return [(f, 2, 0) for f in (0, 1, 3, 5)]
    return [(f, w, 1 if w == 3 else 0) for f in range(6) for w in range(4)]


def bary_inputs():
    rng = np.random.RandomState(77)
    n_tri, n_per = 48, 40
    tris = np.zeros((n_tri, 3, 4), np.float32)
    tris[..., 0] = rng.uniform(-20, 120, (n_tri, 3))
    tris[..., 1] = rng.uniform(-20, 90, (n_tri, 3))
    tris[..., 2] = rng.uniform(0, 1, (n_tri, 3))
    tris[..., 3] = rng.uniform(0.05, 2.0, (n_tri, 3))
    tris[:8, :, :2] = np.round(tris[:8, :, :2])            # integer vertices: samples land exactly on edges
    tris[8] = tris[9]
    tris[8, 2, :2] = tris[8, 1, :2]                          # degenerate (|u.z| < eps)
    tris[10, :, :2] = tris[10, 0, :2] + rng.uniform(-0.4, 0.4, (3, 2))   # sub-pixel triangle
    samples = np.zeros((n_tri, n_per, 2), np.float32)
    for t in range(n_tri):
        lo, hi = tris[t, :, :2].min(axis=0), tris[t, :, :2].max(axis=0)
        px = np.floor(rng.uniform(lo - 1, hi + 1, (n_per, 2)))
        offs = np.array([[0.5, 0.5], [0.375, 0.875], [0.875, 0.625], [0.125, 0.375], [0.625, 0.125]], np.float32)
        samples[t] = px + offs[rng.randint(0, 5, n_per)]
        if t < 8:   # exact vertex / edge midpoints
            samples[t, 0] = tris[t, 0, :2]
            samples[t, 1] = (tris[t, 0, :2] + tris[t, 1, :2]) * 0.5
            samples[t, 2] = (tris[t, 1, :2] + tris[t, 2, :2]) * 0.5
    return tris, samples


def main():
    if not os.path.exists(REF_KAT):
        sys.exit("oracle/_ref/ref_kat missing: make -C oracle ref (needs /root/reference)")
    out = {}
    for (w, h) in LAYOUT_SIZES:
        t, m = run_layout(REF_KAT, w, h)
        out["layout_tiled_%dx%d" % (w, h)] = t
        out["layout_morton_%dx%d" % (w, h)] = m
    for name, c in sample_inputs().items():
        for k in ("texels", "coords", "lod"):
            out["sample_%s_%s" % (name, k)] = c[k]
        if c["offsets"] is not None:
            out["sample_%s_offsets" % name] = c["offsets"]
        for (f, w, b) in sample_combos(name):
            out["sample_%s_f%d_w%d_b%d" % (name, f, w, b)] = run_sample(REF_KAT, c["texels"], c["fmt"], c["layers"], c["mips"], f, w, b,
                                                                      c["coords"], c["lod"])
            if c["offsets"] is not None:
                out["sample_%s_f%d_w%d_b%d_off" % (name, f, w, b)] = run_sample(REF_KAT, c["texels"], c["fmt"], c["layers"], c["mips"], f, w, b,
                                                                              c["coords"], c["lod"], c["offsets"])
    tris, samples = bary_inputs()
    inside, bc, zw = run_bary(REF_KAT, tris, samples)
    out.update(bary_tris=tris, bary_samples=samples, bary_inside=inside, bary_bc=bc, bary_zw=zw)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %d bytes)" % (OUT, len(out), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
