"""RendererCUDA: a B200-native (sm_100a) implementation of SoftGLRender's software pipeline behind the
reference's abstract ``Renderer`` interface.

Layout of the package (only what the hot path needs):
  csrc/     CUDA kernels + the C ABI (include/sglcuda.h)            -> lib/libsglcuda.so
  host/     C++ RendererCUDA classes mirroring Render/*.h of the reference
  harness/  headless offscreen harness: the Renderer-API trace player  -> lib/sgl_player, lib/libsglhost.so
  scene/    Python scene builder emitting Renderer-API traces (caller side of the path, not accelerated)
  capi.py   ctypes binding of the C ABI and of the harness

There is no CPU fallback: importing :mod:`softglrender_b200.capi` raises if the CUDA library has not been built,
and ``sgl_init`` fails without a CUDA device.
"""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PACKAGE_DIR)
# SGL_LIB_DIR selects an alternative build of the SAME sources (compile-time A/B variants made by build.py --variant)
LIB_DIR = os.environ.get("SGL_LIB_DIR") or os.path.join(PACKAGE_DIR, "lib")

__all__ = ["PACKAGE_DIR", "REPO_ROOT", "LIB_DIR"]
