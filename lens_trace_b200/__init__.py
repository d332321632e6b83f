"""lens_trace_b200 -- B200-native (sm_100a) implementation of lens_trace's ray-scene hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C-ABI (include/lens_trace_b200.h) -> liblt_b200.so
  host/            C++ host surface mirroring the reference's include/lens_trace/ API -> liblenstrace.so
  capi.py          ctypes binding of the C-ABI (what a maintainer's FFI stub would call)
  host.py          ctypes binding of the C++ host classes (Model, AccelerationStructureExplicit, Camera, Renderer*)
  build.py         in-tree build recipe (nvcc for sm_100a, g++)

There is no CPU fallback: without the built CUDA library every entry point raises.
"""
from . import layouts  # noqa: F401

__all__ = ["layouts", "capi", "host", "build"]
