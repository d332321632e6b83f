"""Build recipe for the native parts of lens_trace_b200 (in-tree, so the .so files travel with gpurun).

  liblt_b200.so   CUDA kernels + C-ABI (include/lens_trace_b200.h), sm_100a only
  liblenstrace.so host C++ surface mirroring the reference's include/lens_trace/ API; calls the C-ABI

`python -m lens_trace_b200.build` builds both; nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lens_trace_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INC = os.path.join(ROOT, "include")
LIB_B200 = os.path.join(PKG, "liblt_b200.so")
LIB_HOST = os.path.join(PKG, "liblenstrace.so")

NVCC = os.environ.get("LT_NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("LT_CXX", "/usr/bin/g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo", "--fmad=false",
    "-Xcompiler", "-fPIC", "-cudart", "static",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_b200(force=False, verbose=False, ptxas_verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("lt_kernels.cu", "lt_wavefront.cu", "lt_plugin.cu", "lt_bvh.cu", "lt_capi.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [
        os.path.join(INC, "lens_trace_b200.h")]
    if not force and not _newer(LIB_B200, deps):
        return LIB_B200
    cmd = [NVCC] + NVCC_FLAGS + ["-I", INC, "-I", CSRC, "-shared", "-o", LIB_B200] + srcs + ["-ldl"]
    if ptxas_verbose:
        cmd += ["-Xptxas", "-v"]
    _run(cmd, verbose)
    return LIB_B200


def build_host(force=False, verbose=False):
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    hdrs = []
    for d, _, fs in os.walk(os.path.join(INC, "lens_trace")):
        hdrs += [os.path.join(d, f) for f in fs]
    deps = srcs + hdrs + [os.path.join(INC, "lens_trace_b200.h"), LIB_B200]
    if not force and not _newer(LIB_HOST, deps):
        return LIB_HOST
    cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", INC, "-o", LIB_HOST] + srcs + [
        "-L", PKG, "-llt_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd, verbose)
    return LIB_HOST


def build_all(force=False, verbose=False):
    build_b200(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
