"""ctypes binding of the CPU oracle (oracle/liblt_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblt_oracle.so")


class _Scene(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("nodeCount", C.c_int64), ("prims", C.c_void_p), ("primCount", C.c_int64),
                ("materials", C.c_void_p), ("materialCount", C.c_int64), ("lights", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodeTests", C.c_uint64), ("triTests", C.c_uint64)]


_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("lt_oracle.c", "lt_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liblt_oracle.so"])
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.lto_render.argtypes = [C.c_int, C.c_int, C.POINTER(_Scene), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        lib.lto_primary_hits.argtypes = [C.c_int, C.POINTER(_Scene), C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        lib.lto_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int64]
        lib.lto_accumulate.restype = None
        lib.lto_random.argtypes = [C.c_float] * 3
        lib.lto_random.restype = C.c_float
        lib.lto_hemisphere.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        lib.lto_hemisphere.restype = None
        lib.lto_set_fp_mode.argtypes = [C.c_int]
        lib.lto_set_fp_mode.restype = None
        lib.lto_cuda_cosf.argtypes = [C.c_float]
        lib.lto_cuda_cosf.restype = C.c_float
        lib.lto_cuda_sinf.argtypes = [C.c_float]
        lib.lto_cuda_sinf.restype = C.c_float
        _lib = lib
    return _lib


def _scene(sb):
    s = _Scene()
    s.nodes, s.nodeCount = sb.nodes.ctypes.data, len(sb.nodes)
    s.prims, s.primCount = sb.prims.ctypes.data, len(sb.prims)
    s.materials, s.materialCount = sb.materials.ctypes.data, len(sb.materials)
    s.lights = sb.lights.ctypes.data
    return s


def render(kernel, sb, camera, width, height, depth=3, kernel_mode=0, max_ray_depth=16, rows=None, threads=1,
           with_stats=False):
    """One launch of reference kernel `kernel` on the CPU.  rows=(begin,end) renders only those rows."""
    lib = load()
    out = np.zeros((height, width, depth), dtype=np.float32)
    st = Stats()
    cam = np.ascontiguousarray(camera)
    r0, r1 = rows if rows else (0, height)
    s = _scene(sb)
    rc = lib.lto_render(kernel, kernel_mode, C.byref(s), cam.ctypes.data, width, height, depth, max_ray_depth, r0, r1,
                        threads, out.ctypes.data, C.byref(st))
    if rc != 0:
        raise ValueError("lto_render rejected its arguments")
    return (out, st) if with_stats else out


def primary_hits(flavour, sb, camera, width, height):
    lib = load()
    ids = np.zeros((height, width), dtype=np.int32)
    hit = np.zeros((height, width), dtype=np.int32)
    tuv = np.zeros((height, width, 3), dtype=np.float32)
    st = Stats()
    cam = np.ascontiguousarray(camera)
    s = _scene(sb)
    lib.lto_primary_hits(flavour, C.byref(s), cam.ctypes.data, width, height, ids.ctypes.data, hit.ctypes.data,
                         tuv.ctypes.data, C.byref(st))
    return ids, hit, tuv, st


def accumulate(acc, sample, frame_count):
    lib = load()
    assert acc.dtype == np.float32 and sample.dtype == np.float32 and acc.size == sample.size
    lib.lto_accumulate(acc.ctypes.data, sample.ctypes.data, frame_count, acc.size)
    return acc


def random(u, v, seed):
    return load().lto_random(u, v, seed)


def hemisphere(u1, u2, up):
    lib = load()
    upa = np.ascontiguousarray(up, dtype=np.float32)
    out = np.zeros(4, np.float32)
    lib.lto_hemisphere(float(u1), float(u2), upa.ctypes.data, out.ctypes.data)
    return out


FP_DEVICE, FP_PLAIN = 0, 1


class fp_mode:
    """`with fp_mode(FP_PLAIN): ...` -- evaluate the restatement as plain C (see lt_oracle.h) inside the block."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        lib = load()
        self.prev = lib.lto_get_fp_mode()
        lib.lto_set_fp_mode(self.mode)

    def __exit__(self, *exc):
        load().lto_set_fp_mode(self.prev)
