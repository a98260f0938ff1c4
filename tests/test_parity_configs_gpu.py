"""BASELINE.json configs 3, 4 and 5 as GPU parity cases (config 1 and 2 live in test_parity_gpu.py).  Same bar: depth
bit-exact, colour within 1/255 on >= 99.9 % of pixels; checker = compiled reference (deterministic build) when it
travelled with the snapshot, else the CPU restatement.  At sizes the checker cannot afford, size-independent properties:
determinism, texture-layout invariance, tile-sharded == unsharded."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import compare_outputs

pytestmark = pytest.mark.gpu


def _run(binary, trace, out, data, env=None):
    from softglrender_b200 import workloads
    from softglrender_b200.scene.trace import read_outputs
    workloads.run_player(binary, trace, out=out, data_dir=data, env=env)
    r = read_outputs(out)
    os.remove(out)
    return r


def test_config3_boombox_glasstable_fxaa_4k(checker_player, work_dir):
    """Config 3 at full size: BoomBox + GlassTable, shadow pass, alpha-blended glass (fused ordered tile kernel), FXAA
    pass sampling the rendered colour target, 3840x2160."""
    from softglrender_b200 import workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    trace, data = workloads.build_c3(os.path.join(work_dir, "c3"))
    ref = _run(checker_player, trace, os.path.join(data, "c3.ref.out"), data)
    got = _run(workloads.CUDA_PLAYER, trace, os.path.join(data, "c3.cuda.out"), data)
    rep = compare_outputs(ref, got)
    print("config3 parity:", rep)
    assert set(ref) >= {"color", "depth", "shadow", "color_prefxaa"}
    os.remove(trace)


@pytest.mark.parametrize("msaa", [False, True])
def test_config4_soup_100k_vs_checker(msaa, checker_player, work_dir):
    """Config 4 at a size the CPU checker finishes in seconds: 100k random triangles (0.5-64 px, random winding),
    8 mip-mapped REPEAT textures (LINEAR_MIPMAP_LINEAR), 1920x1080."""
    from softglrender_b200 import workloads
    trace, data = workloads.build_c4(os.path.join(work_dir, "c4"), msaa=msaa)
    ref = _run(checker_player, trace, os.path.join(data, "c4.ref.out"), data)
    got = _run(workloads.CUDA_PLAYER, trace, os.path.join(data, "c4.cuda.out"), data)
    print("config4 parity (msaa=%s):" % msaa, compare_outputs(ref, got))
    # Morton-tiled texture storage (BASELINE config 4 names it) changes addressing only
    morton = _run(workloads.CUDA_PLAYER, trace, os.path.join(data, "c4.morton.out"), data, env={"SGL_TEXTURE_LAYOUT": "2"})
    for k in got:
        assert np.array_equal(got[k], morton[k]), k
    os.remove(trace)


def test_config4_soup_2m_8k_properties(work_dir):
    """Config 4 towards full size (2 M triangles, 7680x4320, 2048^2 Morton textures): too slow for the CPU checker, so the
    size-independent properties: (1) two runs are byte-identical, (2) linear and Morton texture layouts agree,
    (3) a tile-sharded render (one GPU playing both ranks in turn) reproduces the unsharded frame tile by tile,
    (4) no clip-arena overflow."""
    import torch
    from softglrender_b200 import capi, multigpu as M, workloads
    trace, data = workloads.build_c4(os.path.join(work_dir, "c4"), n_tris=2000000, width=7680, height=4320, tex_size=2048)
    capi.init(0)
    lib = capi.load()
    frames = {}
    for layout in (2, 0):
        os.environ["SGL_TEXTURE_LAYOUT"] = str(layout)
        p = capi.Player(trace, data)
        try:
            p.setup()
            p.frame(sync=True)
            buf, (w, h, _, _) = p.readback("color")
            frames[layout] = (buf.reshape(h, w, 4).copy(), p.readback("depth")[0].copy())
            if layout == 0:
                p.frame(sync=True)
                assert np.array_equal(p.readback("color")[0].reshape(h, w, 4), frames[0][0])       # (1)
                assert capi.counters()["clip_overflow"] == 0                                          # (4)
                tex = p.texture_handle("color")
                world = 2
                g = M.TileGather(w, h, 0, world, "interleave")
                g.install(lib)
                n = g.max_count * g.tile_bytes
                stage = torch.zeros(n, dtype=torch.uint8, device="cuda")
                junk = torch.full((n,), 0x5A, dtype=torch.uint8, device="cuda")
                out = np.zeros_like(frames[0][0])
                for r in range(world):
                    capi.check(lib.sgl_set_rank(r, world))
                    for q in range(world):
                        capi.check(lib.sgl_tiles_unpack(tex, q, junk.data_ptr(), n))
                    p.frame(sync=False)
                    cnt = C.c_int()
                    capi.check(lib.sgl_tiles_pack(tex, r, stage.data_ptr(), n, C.byref(cnt)))
                    capi.check(lib.sgl_wait_idle())
                    M.unpack_tiles_host(out, stage.cpu().numpy()[:cnt.value * g.tile_bytes].reshape(-1, M.TILE, M.TILE, 4), g.owner, r)
                assert np.array_equal(out, frames[0][0])                                              # (3)
        finally:
            lib.sgl_set_tile_owner_map(None, 0, 0)
            lib.sgl_set_rank(0, 1)
            p.close()
            os.environ.pop("SGL_TEXTURE_LAYOUT", None)
    assert np.array_equal(frames[0][0], frames[2][0]) and np.array_equal(frames[0][1], frames[2][1])   # (2)
    os.remove(trace)


@pytest.mark.parametrize("model,views", [("AfricanHead", [0, 700, 1400, 2047]), ("Robot", [2048, 2900, 3500, 4095])])
def test_config5_multiview_batch(model, views, checker_player, work_dir):
    """Config 5: views of the 4096-view Fibonacci sphere (AfricanHead = Blinn-Phong OBJ with shadows, Robot = PBR glTF),
    512x512, no AA; every view of the batch is compared."""
    from softglrender_b200 import workloads
    if workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    trace, data = workloads.build_c5(os.path.join(work_dir, "c5"), model, views)
    ref = _run(checker_player, trace, os.path.join(data, "c5.ref.out"), data)
    got = _run(workloads.CUDA_PLAYER, trace, os.path.join(data, "c5.cuda.out"), data)
    rep = compare_outputs(ref, got)
    assert len([k for k in rep if k.startswith("color_v")]) == len(views)
    print("config5 parity:", model, rep)
