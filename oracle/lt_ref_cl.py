"""ctypes binding of oracle/_ref/libltref_cl.so -- the reference's OWN OpenCL kernel files compiled for
the host CPU through oracle/cl_shim (oracle/build_ref_cl.sh).  TEST INFRASTRUCTURE ONLY: it is the
pin of the hand restatement (lt_oracle.c) for everything that exists in the reference only as
OpenCL C.  The library can only be (re)built where /root/reference exists; the built file travels."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libltref_cl.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ltrefcl_render.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int] * 9
        lib.ltrefcl_struct_sizes.argtypes = [C.c_void_p]
        lib.ltrefcl_struct_sizes.restype = None
        _lib = lib
    return _lib


def struct_sizes():
    out = np.zeros(5, np.int32)
    load().ltrefcl_struct_sizes(out.ctypes.data)
    return out.tolist()


def render(kernel, sb, camera, width, height, depth=3, kernel_mode=0, work_block=None, local=None,
           max_ray_depth=16, threads=0, fill=0.0):
    """One RendererOpenCL::render() of kernel file `kernel` (layouts.KERNEL_* id 1..6) on the CPU.
    work_block defaults to the whole image (THREAD_ORGANIZATION_MODE_MAX_FIT on a device whose maximum
    work-item sizes exceed the image, renderer_opencl.cpp:84-85)."""
    lib = load()
    wb = work_block or (width, height)
    ls = local or (1, 1)
    out = np.full((height, width, depth), fill, dtype=np.float32)
    cam = np.ascontiguousarray(camera)
    rc = lib.ltrefcl_render(kernel, kernel_mode, sb.nodes.ctypes.data, sb.prims.ctypes.data,
                            sb.materials.ctypes.data, sb.lights.ctypes.data, cam.ctypes.data, out.ctypes.data,
                            width, height, depth, wb[0], wb[1], ls[0], ls[1], max_ray_depth, threads)
    if rc != 0:
        raise ValueError("ltrefcl_render rejected its arguments")
    return out
