"""The drop-in boundary as a compile-time fact: RendererCUDA's host classes build against the UNMODIFIED reference headers
(Render/Renderer.h:24-59 and companions) where the reference tree is mounted, and against this repo's signature-compatible
re-declarations everywhere."""
import os
import subprocess

import pytest

from conftest import ROOT

REF = "/root/reference"
HOST = os.path.join(ROOT, "softglrender_b200", "host")
SRC = os.path.join(HOST, "Render", "CUDA", "RendererCUDA.cpp")


def _syntax_check(includes, std):
    cmd = ["g++", "-std=" + std, "-fsyntax-only", "-Wall", "-Werror=overloaded-virtual", "-I" + os.path.join(ROOT, "include")] + \
          ["-I" + i for i in includes] + [SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]


def test_host_classes_compile_against_reference_headers():
    if not os.path.isdir(os.path.join(REF, "src", "Render")):
        pytest.skip("reference tree not mounted")
    # the reference's own language level (CMakeLists.txt: gnu++11) and include roots; no repo header shadows Render/*.h
    _syntax_check([os.path.join(REF, "src"), os.path.join(REF, "third_party", "glm")], "gnu++11")


def test_host_classes_compile_against_own_declarations():
    _syntax_check([HOST], "c++17")


def test_interface_redeclaration_matches_reference_virtuals():
    """RenderAPI.h must declare the same virtual surface as the reference: a virtual the reference lacks (e.g. an extra
    virtual destructor) would let code compile here that does not compile there."""
    if not os.path.isdir(os.path.join(REF, "src", "Render")):
        pytest.skip("reference tree not mounted")
    import re

    def virtual_dtors(text):
        return sorted(set(re.findall(r"virtual\s+~(\w+)\s*\(", text)))
    ours = virtual_dtors(open(os.path.join(HOST, "Render", "RenderAPI.h")).read())
    theirs = []
    for name in ("Renderer.h", "Texture.h", "Framebuffer.h", "Vertex.h", "Uniform.h", "ShaderProgram.h", "PipelineStates.h", "RenderStates.h"):
        theirs += virtual_dtors(open(os.path.join(REF, "src", "Render", name)).read())
    assert ours == sorted(set(theirs)), (ours, theirs)


def test_viewer_cuda_compiles_headless_and_with_the_gl_present_path(tmp_path):
    """ViewerCUDA against the reference's own Viewer.h: the headless form the integration harness links, and the windowed form
    (-DSGL_VIEWER_CUDA_PRESENT_GL) that uploads the frame into the Viewer's GL texture exactly like ViewerSoftware::swapBuffer
    (ViewerSoftware.h:30-44) -- compiled against the reference's vendored glad headers; there is no display here to run it."""
    if not os.path.isdir(os.path.join(REF, "src", "Viewer")):
        pytest.skip("reference tree not mounted")
    tu = tmp_path / "viewer_cuda_tu.cpp"
    tu.write_text('#include "Viewer/ViewerCUDA.h"\nint viewer_cuda_tu() { return sizeof(SoftGL::View::ViewerCUDA) > 0; }\n')
    inc = [os.path.join(REF, "src"), os.path.join(REF, "src", "Viewer"), os.path.join(REF, "third_party", "glm"),
           os.path.join(REF, "third_party", "glad", "include"), os.path.join(REF, "third_party", "json11"),
           os.path.join(REF, "third_party", "imgui"), os.path.join(REF, "third_party", "assimp", "include"),
           os.path.join(ROOT, "integration", "_build", "assimp", "include"), HOST, os.path.join(ROOT, "include")]
    for defs in ([], ["-DSGL_VIEWER_CUDA_PRESENT_GL"]):
        cmd = ["g++", "-std=gnu++11", "-fsyntax-only"] + defs + ["-I" + i for i in inc if os.path.isdir(i)] + [str(tu)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, (defs, r.stdout[-3000:])
