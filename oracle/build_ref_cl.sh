#!/bin/bash
# Builds oracle/_ref/libltref_cl.so: the reference's OWN OpenCL kernel files, read where they lie under
# /root/reference, compiled as C++ for the host CPU through oracle/cl_shim/opencl_c_on_cpp.h and run one
# work-item at a time by oracle/cl_shim/ref_cl_host.cpp.  No OpenCL implementation exists in this image or
# on the GPU box, so this is the only way to EXECUTE the text of the lighting / GI kernels; it is what pins
# the hand restatement in lt_oracle.c for everything that exists only as OpenCL C (tests/test_oracle_cl.py).
#
# The kernel text is not edited by hand and not copied into the repository's history: the only rewrites
# are mechanical and listed here, and their outputs go to oracle/_ref/cl/ (git-ignored).
#   1. `(float2)(` `(float3)(` `(float4)(`  ->  `float2(` ...   OpenCL vector literal -> C++ constructor call
#   2. (the two *_depth.inc copies only) `int maxRayDepth = 16;` -> `int maxRayDepth = clshim_max_ray_depth;`
#      so the bounce cap of BASELINE.json's workloads (4) can be run; the plain copies keep 16.
set -euo pipefail
REF=${LT_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/resources/kernels/opencl" ]; then echo "reference not present at $REF; keeping existing $OUT" >&2; exit 0; fi
mkdir -p "$OUT/cl"
lit() { sed -E 's/\((float[234])\)\(/\1(/g' "$1" > "$2"; ! grep -nE '\((float|int|uint)[234]\)\(' "$2"; }
lit "$REF/resources/kernels/opencl/basic.cl"                                       "$OUT/cl/basic.inc"
lit "$REF/resources/kernels/opencl/basic_lighting.cl"                              "$OUT/cl/basic_lighting.inc"
lit "$REF/resources/kernels/opencl/global_illumination.cl"                         "$OUT/cl/global_illumination_resources.inc"
lit "$REF/examples/accumulator/resources/kernels/accumulator.cl"                   "$OUT/cl/accumulator.inc"
lit "$REF/examples/custom_kernel/resources/kernels/custom_opencl.cl"               "$OUT/cl/custom_opencl.inc"
lit "$REF/examples/global_illumination/resources/kernels/global_illumination.cl"   "$OUT/cl/global_illumination_example.inc"
for v in resources example; do
  sed 's/int maxRayDepth = 16;/int maxRayDepth = clshim_max_ray_depth;/' "$OUT/cl/global_illumination_$v.inc" \
    > "$OUT/cl/global_illumination_${v}_depth.inc"
  grep -q 'clshim_max_ray_depth' "$OUT/cl/global_illumination_${v}_depth.inc"
done
# -O2 without -ffast-math; x86-64 baseline has no FMA and -ffp-contract=off forbids contraction anyway.
# -Wno-narrowing: `{0, 0, distance(..) - SHADOW_RAY_EPSILON, 0, 0}` narrows double -> float inside braces,
# which C (and OpenCL C) allow and C++ only diagnoses; the conversion performed is the same.
/usr/bin/g++ -std=gnu++20 -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -w -Wno-narrowing \
  -I"$HERE/cl_shim" -I"$OUT" "$HERE/cl_shim/ref_cl_host.cpp" -o "$OUT/libltref_cl.so" -lpthread
echo "built $OUT/libltref_cl.so"
