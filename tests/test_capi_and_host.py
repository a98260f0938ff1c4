"""CPU-side checks: the C-ABI library loads and exports every declared symbol; host logic of the scene builder."""
import ctypes
import hashlib
import os
import re
import struct

import numpy as np
import pytest

from conftest import ROOT


def test_header_symbols_exported():
    """libsglcuda.so exports every function include/sglcuda.h declares (no compute calls without a GPU)."""
    from softglrender_b200 import capi
    with open(os.path.join(ROOT, "include", "sglcuda.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(sgl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = capi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libsglcuda.so does not export " + name
    assert declared == set(capi.C_ABI_SYMBOLS), declared ^ set(capi.C_ABI_SYMBOLS)


def test_ctypes_mirrors_have_the_sizes_of_the_header(tmp_path):
    """The Python mirrors of the C ABI's structs (capi.py) are as large as the C compiler makes the header's: a field added on
    one side only would shift every later counter silently."""
    import shutil
    import subprocess
    from softglrender_b200 import capi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "sglcuda.h"\nint main(void) { printf("%zu %zu\\n", sizeof(SglCounters), sizeof(SglKernelTime)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    c_counters, c_ktime = (int(x) for x in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split())
    assert ctypes.sizeof(capi.SglCounters) == c_counters
    assert ctypes.sizeof(capi.SglKernelTime) == c_ktime


def test_no_cpu_fallback():
    """Without a CUDA device sgl_init must fail loudly; the product has no CPU path."""
    import torch
    from softglrender_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    rc = lib.sgl_init(0, 0, 1)
    assert rc != 0
    assert b"no CPU fallback" in lib.sgl_last_error()
    with pytest.raises(RuntimeError):
        capi.init(0)


def test_shader_reflection_matches_reference_layout():
    """Uniform offsets / sampler slots / define bits = getUniformsDesc()/getDefines() of the software shaders
    (PbrSoft.h:77-103, BlinnPhongSoft.h:71-96, BasicSoft.h:49-61, SkyboxSoft.h:44-62, FxaaSoft.h:39-53)."""
    from softglrender_b200 import capi
    lib = capi.load()
    PBR, BP, BASIC, SKY, FXAA, PRE = 3, 2, 1, 4, 7, 6
    assert lib.sgl_shader_uniform_offset(PBR, b"UniformsModel") == 0
    assert lib.sgl_shader_uniform_offset(PBR, b"UniformsScene") == 256
    assert lib.sgl_shader_uniform_offset(PBR, b"UniformsMaterial") == 320
    assert lib.sgl_shader_uniform_offset(BASIC, b"UniformsMaterial") == 256
    assert lib.sgl_shader_uniform_offset(BASIC, b"UniformsScene") == -1
    assert lib.sgl_shader_uniform_offset(PRE, b"UniformsPrefilter") == 256
    assert [lib.sgl_shader_sampler_slot(PBR, n) for n in (b"u_albedoMap", b"u_normalMap", b"u_emissiveMap", b"u_aoMap",
                                                           b"u_metalRoughnessMap", b"u_irradianceMap", b"u_prefilterMap")] == list(range(7))
    assert lib.sgl_shader_sampler_slot(PBR, b"u_shadowMap") == -1      # PBR has no shadow sampler
    assert lib.sgl_shader_sampler_slot(BP, b"u_shadowMap") == 4
    assert lib.sgl_shader_sampler_slot(SKY, b"u_cubeMap") == 1
    assert lib.sgl_shader_define_bit(PBR, b"METALROUGHNESS_MAP") == 4
    assert lib.sgl_shader_define_bit(SKY, b"CUBE_MAP") == -1           # ignored name, like ShaderProgramSoft.h:37-45
    assert [lib.sgl_shader_varying_floats(s) for s in (BASIC, BP, PBR, SKY, FXAA)] == [0, 32, 28, 4, 2]
    assert lib.sgl_shader_uniform_size(FXAA) == 8


def test_trace_roundtrip_and_determinism():
    from softglrender_b200.scene import synth
    a = synth.kat_trace(96, 64, msaa=True, seed=3).tobytes()
    b = synth.kat_trace(96, 64, msaa=True, seed=3).tobytes()
    assert a == b and a[:4] == b"SGLT"
    # walk the command stream
    off, n = 8, 0
    while off < len(a):
        op, ln = struct.unpack_from("<II", a, off)
        off += 8 + ln
        n += 1
    assert off == len(a) and n > 50


def test_golden_traces_are_reproducible():
    """The committed fixtures are tied to traces this checkout can regenerate byte-for-byte."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    for name, (builder, needs_assets) in make_golden.FIXTURES.items():
        if needs_assets:
            continue
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        sha = hashlib.sha256(builder(None).tobytes()).hexdigest()
        assert sha == str(g["trace_sha256"]), name


def test_uniform_packing_sizes():
    from softglrender_b200.scene import viewer
    m = np.eye(4, dtype=np.float32)
    assert len(viewer.pack_uniforms_model(True, m, m, np.eye(3), m)) == 256
    assert len(viewer.pack_uniforms_scene((0, 0, 0), (1, 1, 1), (2, 2, 2), (3, 3, 3))) == 64
    b = viewer.pack_uniforms_material(True, False, True, 10.0, 0.5, (1, 2, 3, 4))
    assert len(b) == 48 and struct.unpack_from("<f", b, 12)[0] == 10.0 and struct.unpack_from("<4f", b, 32) == (1, 2, 3, 4)


def test_camera_reversed_z_projection():
    """Camera::projectionMatrix (Camera.cpp:26-44): infinite far plane, reversed-Z maps near -> 1, far -> 0."""
    from softglrender_b200.scene.viewer import Camera
    c = Camera(60.0, 16 / 9, 0.01)
    c.reverse_z = True
    p = c.projection()
    near = p @ np.array([0, 0, -0.01, 1], np.float32)
    far = p @ np.array([0, 0, -1e6, 1], np.float32)
    assert abs(near[2] / near[3] - 1.0) < 1e-6 and abs(far[2] / far[3]) < 1e-6
    c.reverse_z = False
    p = c.projection()
    near = p @ np.array([0, 0, -0.01, 1], np.float32)
    assert abs(near[2] / near[3]) < 1e-5
