"""CPU: the unit-level known-answer vectors the REFERENCE produced (tests/golden/unit_kats.npz, generator
tests/golden/make_unit_kats.py -> oracle/_ref/ref_kat) against (a) the CPU restatement (oracle/_build/oracle_kat) and
(b) the numpy statement of the Tiled / Morton index functions (Base/Buffer.h:151-158,185-202)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_unit_kats as K  # noqa: E402

ORACLE_KAT = os.path.join(ROOT, "oracle", "_build", "oracle_kat")


@pytest.fixture(scope="module")
def golden():
    return np.load(K.OUT)


@pytest.fixture(scope="module")
def oracle_kat(oracle_player):
    assert os.path.exists(ORACLE_KAT), "make -C oracle restate builds oracle_kat next to oracle_player"
    return ORACLE_KAT


def layout_index(layout, w, x, y):
    """numpy statement of TiledBuffer / MortonBuffer::convertIndex."""
    x, y = x.astype(np.uint32), y.astype(np.uint32)
    if layout == 1:
        tw = (w + 3) // 4
        return (((y >> 2) * tw + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3)
    tw = (w + 31) // 32
    res = (x & 31) | ((y & 31) << 16)
    res = (res | (res << 4)) & 0x0f0f0f0f
    res = (res | (res << 2)) & 0x33333333
    res = (res | (res << 1)) & 0x55555555
    morton = (res | (res >> 15)) & 0xffff
    return (((y >> 5) * tw + (x >> 5)) << 10) + morton


def expected_storage(layout, w, h):
    ts = 4 if layout == 1 else 32
    iw, ih = (w + ts - 1) // ts * ts, (h + ts - 1) // ts * ts
    out = np.zeros(iw * ih, np.uint32)
    yy, xx = np.mgrid[0:h, 0:w]
    out[layout_index(layout, w, xx.ravel(), yy.ravel())] = (xx.ravel() | (yy.ravel() << 16)).astype(np.uint32)
    return out


@pytest.mark.parametrize("size", K.LAYOUT_SIZES)
def test_layout_formulas_match_reference_buffers(size, golden):
    w, h = size
    assert np.array_equal(expected_storage(1, w, h), golden["layout_tiled_%dx%d" % (w, h)])
    assert np.array_equal(expected_storage(2, w, h), golden["layout_morton_%dx%d" % (w, h)])


@pytest.mark.parametrize("name", sorted(K.sample_inputs()))
def test_oracle_sampler_matches_reference_vectors(name, golden, oracle_kat):
    c = K.sample_inputs()[name]
    assert np.array_equal(c["texels"], golden["sample_%s_texels" % name]), "seeded inputs drifted from the committed vectors"
    for (f, w, b) in K.sample_combos(name):
        got = K.run_sample(oracle_kat, c["texels"], c["fmt"], c["layers"], c["mips"], f, w, b, c["coords"], c["lod"])
        assert np.array_equal(got, golden["sample_%s_f%d_w%d_b%d" % (name, f, w, b)]), (name, f, w, b)
        if c["offsets"] is not None:
            got = K.run_sample(oracle_kat, c["texels"], c["fmt"], c["layers"], c["mips"], f, w, b, c["coords"], c["lod"], c["offsets"])
            assert np.array_equal(got, golden["sample_%s_f%d_w%d_b%d_off" % (name, f, w, b)]), (name, f, w, b, "offsets")


def test_oracle_barycentric_matches_reference_vectors(golden, oracle_kat):
    inside, bc, zw = K.run_bary(oracle_kat, golden["bary_tris"], golden["bary_samples"])
    assert np.array_equal(inside, golden["bary_inside"])
    assert np.array_equal(bc.view(np.uint32), golden["bary_bc"].view(np.uint32))
    assert np.array_equal(zw.view(np.uint32), golden["bary_zw"].view(np.uint32))
    assert 0.1 < golden["bary_inside"].mean() < 0.9          # the vectors exercise both outcomes


def test_reference_reproduces_committed_vectors(golden):
    """Where the compiled reference travelled with the snapshot, the committed vectors are still what it answers."""
    if not os.path.exists(K.REF_KAT):
        pytest.skip("oracle/_ref/ref_kat not built (no reference tree)")
    t, m = K.run_layout(K.REF_KAT, 37, 21)
    assert np.array_equal(t, golden["layout_tiled_37x21"]) and np.array_equal(m, golden["layout_morton_37x21"])
    inside, bc, zw = K.run_bary(K.REF_KAT, golden["bary_tris"], golden["bary_samples"])
    assert np.array_equal(inside, golden["bary_inside"]) and np.array_equal(bc.view(np.uint32), golden["bary_bc"].view(np.uint32))
    c = K.sample_inputs()["npot48"]
    got = K.run_sample(K.REF_KAT, c["texels"], 0, 1, True, 5, 0, 0, c["coords"], c["lod"])
    assert np.array_equal(got, golden["sample_npot48_f5_w0_b0"])
