"""Regenerates tests/golden/*.npz from the compiled reference (oracle/_ref/ref_player_st, built from /root/reference
by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture = outputs of the REFERENCE RendererSoft (single-worker, deterministic) on a trace that the test-suite
can regenerate byte-identically from softglrender_b200.scene (sha256 of the trace is stored with the fixture).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softglrender_b200 import workloads                     # noqa: E402
from softglrender_b200.scene import synth, scenes           # noqa: E402
from softglrender_b200.scene.trace import read_outputs      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

FIXTURES = {
    # name: (builder, needs_assets)
    "kat_1x": (lambda ad: synth.kat_trace(192, 128, msaa=False, reverse_z=False, seed=7), False),
    "kat_ms4_revz": (lambda ad: synth.kat_trace(192, 128, msaa=True, reverse_z=True, seed=11), False),
    "kat_nomip_odd": (lambda ad: synth.kat_trace(161, 97, msaa=True, reverse_z=False, seed=23, mipmaps=False), False),
    "c1_cube_200x160": (lambda ad: scenes.config1_cube(ad, 200, 160), True),
    # ShaderSkybox with EQUIRECTANGULAR_MAP (atan2 / asin, SkyboxSoft.h:83-98) sampling Room.jpeg directly: pbrIbl off
    "sky_equirect_200x120": (lambda ad: scenes.config2_helmet(ad, width=200, height=120, skybox="Room", model="Cube", pbr_ibl=False), True),
    # degenerate inputs: clear-only pass, empty draw, 1x1 / 3x2 targets, small viewport, load pass with a leading blended
    # draw, zero-area and fully clipped triangles, empty depth-only pass
    "edge_1x": (lambda ad: synth.edge_trace(msaa=False), False),
    "edge_ms4": (lambda ad: synth.edge_trace(msaa=True), False),
}


def build_trace(name, work):
    builder, needs_assets = FIXTURES[name]
    ad = workloads.assets_dir() if needs_assets else None
    w = builder(ad)
    path = os.path.join(work, name + ".sglt")
    w.save(path)
    with open(path, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    return path, sha


def main():
    work = os.path.join(ROOT, "build", "golden")
    os.makedirs(work, exist_ok=True)
    if not os.path.exists(workloads.REF_PLAYER_ST):
        raise SystemExit("oracle/_ref/ref_player_st missing: run `make -C oracle ref` where /root/reference is mounted")
    for name in FIXTURES:
        trace, sha = build_trace(name, work)
        out = os.path.join(work, name + ".ref.out")
        workloads.run_player(workloads.REF_PLAYER_ST, trace, out=out, data_dir=work)
        arrays = read_outputs(out)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), trace_sha256=np.array(sha), **arrays)
        print(name, sha[:12], {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
