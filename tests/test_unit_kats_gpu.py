"""GPU: the device functions of the pipeline (sgl_kat_* entry points of the C ABI) against the unit-level known-answer
vectors the REFERENCE produced (tests/golden/unit_kats.npz): sampler over filter x wrap x border x NPOT x offsets x cube x
image layout, the split-phase taps of the straight-line shader paths, barycentric / z / 1/w, and the raw Tiled / Morton
storage (Base/Buffer.h:141-213).  Everything here is bit-exact."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_unit_kats as K  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(K.OUT)


@pytest.fixture(scope="module")
def lib():
    from softglrender_b200 import capi
    capi.init(0)
    return capi.load()


def _texture(lib, texels, fmt, layers, mips, layout):
    from softglrender_b200 import capi
    h, w = texels.shape[1], texels.shape[2]
    desc = capi.SglTextureDesc(width=w, height=h, type=1 if layers == 6 else 0, format=fmt, use_mipmaps=int(mips), multi_sample=0, layout=layout)
    handle = C.c_int()
    capi.check(lib.sgl_texture_create(C.byref(desc), C.byref(handle)))
    for l in range(layers):
        img = np.ascontiguousarray(texels[l])
        capi.check(lib.sgl_texture_upload(handle.value, l, 0, img.ctypes.data))
    if mips:
        capi.check(lib.sgl_texture_gen_mips(handle.value))
    return handle.value


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("name", sorted(K.sample_inputs()))
def test_sampler_matches_reference_vectors(name, layout, golden, lib):
    from softglrender_b200 import capi
    c = K.sample_inputs()[name]
    tex = _texture(lib, c["texels"], c["fmt"], c["layers"], c["mips"], layout)
    coords = np.ascontiguousarray(c["coords"], np.float32)
    lod = np.ascontiguousarray(c["lod"], np.float32)
    n = len(lod)
    checked = 0
    for (f, w, b) in K.sample_combos(name):
        variants = [(None, "")] + ([(np.ascontiguousarray(c["offsets"], np.int32), "_off")] if c["offsets"] is not None else [])
        for offs, suffix in variants:
            want = golden["sample_%s_f%d_w%d_b%d%s" % (name, f, w, b, suffix)]
            out = np.zeros(n, np.uint32)
            capi.check(lib.sgl_kat_sample(tex, f, w, b, coords.ctypes.data, lod.ctypes.data, offs.ctypes.data if offs is not None else None,
                                          n, 0, out.ctypes.data))
            assert np.array_equal(out, want), (name, layout, f, w, b, suffix, int((out != want).sum()))
            checked += 1
            if layout == 0 and c["fmt"] == 0 and f == 1 and w in (0, 2):   # "simple" sampler: the split-phase taps give the same bits
                out2 = np.zeros(n, np.uint32)
                capi.check(lib.sgl_kat_sample(tex, f, w, b, coords.ctypes.data, lod.ctypes.data, offs.ctypes.data if offs is not None else None,
                                              n, 1, out2.ctypes.data))
                assert np.array_equal(out2, want), (name, "split-phase", f, w, suffix)
    assert checked >= 4
    capi.check(lib.sgl_texture_destroy(tex))


def test_barycentric_matches_reference_vectors(golden, lib):
    from softglrender_b200 import capi
    tris, samples = golden["bary_tris"], golden["bary_samples"]
    n = samples.shape[1]
    for t in range(tris.shape[0]):
        tri = np.ascontiguousarray(tris[t], np.float32)
        xy = np.ascontiguousarray(samples[t], np.float32)
        bc, inside, zw = np.zeros((n, 3), np.float32), np.zeros(n, np.int32), np.zeros((n, 2), np.float32)
        capi.check(lib.sgl_kat_barycentric(tri.ctypes.data, xy.ctypes.data, n, bc.ctypes.data, inside.ctypes.data, zw.ctypes.data))
        assert np.array_equal(inside, golden["bary_inside"][t]), t
        assert np.array_equal(bc.view(np.uint32), golden["bary_bc"][t].view(np.uint32)), t
        m = inside.astype(bool)
        assert np.array_equal(zw[m].view(np.uint32), golden["bary_zw"][t][m].view(np.uint32)), t


@pytest.mark.parametrize("size", K.LAYOUT_SIZES)
def test_tiled_and_morton_storage_matches_reference_buffers(size, golden, lib):
    """Upload a linear image, read the RAW device storage back (kind 2) and compare it with what TiledBuffer / MortonBuffer
    hold for the same image -- pins the index functions and the tile padding to the reference, not to ourselves."""
    from softglrender_b200 import capi
    w, h = size
    yy, xx = np.mgrid[0:h, 0:w]
    img = (xx | (yy << 16)).astype(np.uint32)
    for layout, key in ((1, "layout_tiled_%dx%d"), (2, "layout_morton_%dx%d")):
        want = golden[key % (w, h)]
        tex = _texture(lib, img.view(np.uint8).reshape(1, h, w, 4), 0, 1, False, layout)
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        capi.check(lib.sgl_texture_device_ptr(tex, 0, 0, 0, C.byref(ptr), C.byref(nbytes)))
        assert nbytes.value == want.nbytes, (layout, nbytes.value, want.nbytes)
        raw = np.full(want.shape, 0xDEADBEEF, np.uint32)
        capi.check(lib.sgl_texture_readback(tex, 0, 0, 2, raw.ctypes.data, raw.nbytes))
        assert np.array_equal(raw, want), (layout, size)
        back = np.zeros((h, w), np.uint32)                      # and the ordinary read-back undoes the layout
        capi.check(lib.sgl_texture_readback(tex, 0, 0, 0, back.ctypes.data, back.nbytes))
        assert np.array_equal(back, img)
        capi.check(lib.sgl_texture_destroy(tex))
