#!/bin/bash
# Builds the reference itself (host classes + CUDA launcher) from its own sources, where they lie
# under /root/reference, into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# Nothing is copied into the repo's history.  The reference's cmake build is NOT run: five source
# files are compiled directly (no -O flag, like its CMakeLists.txt, which sets none).
#   -include array : src/model.cpp uses std::array without including <array> (fails on g++ 13)
# Also stages, as run-time inputs of the reference renderer (it compiles its kernel source with
# NVRTC at run time and loads OBJ files by path):
#   oracle/_ref/resources/kernels/cuda/basic.cu      (verbatim)
#   oracle/_ref/resources/kernels/cuda/id_dump_{a,b}.cu : basic.cu with the colour store replaced by
#     the hit record (primitiveIndex/t/u and hitType/v/t) -- the reference's own plugin mechanism
#     is the only way to read its hit ids.
#   oracle/_ref/resources/models/*                   (verbatim)
set -euo pipefail
REF=${LT_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF; keeping existing $OUT" >&2; exit 0; fi
CUDA=${CUDA_HOME:-/usr/local/cuda}
mkdir -p "$OUT/resources/kernels/cuda" "$OUT/resources/models"
/usr/bin/g++ -std=c++17 -fPIC -shared -include array -w \
  -I"$REF/include" -I"$CUDA/include" \
  "$HERE/ref_shim.cpp" "$REF/src/model.cpp" "$REF/src/camera.cpp" "$REF/src/acceleration_structure_explicit.cpp" \
  "$REF/src/resource.cpp" "$REF/src/cuda/renderer_cuda.cpp" \
  -L"$CUDA/lib64" -L"$CUDA/lib64/stubs" -lcuda -lnvrtc -lcudart -Wl,-Bsymbolic \
  -o "$OUT/libltref.so"
K="$REF/resources/kernels/cuda/basic.cu"
install -m 644 "$K" "$OUT/resources/kernels/cuda/basic.cu"
install -m 644 "$REF"/resources/models/* "$OUT/resources/models/"
# hit-record dumps: replace the final colour assignment of shade() (basic.cu:325)
PAT='outputColor = make_float3(material->diffuse\[0\], material->diffuse\[1\], material->diffuse\[2\]);'
sed "s|$PAT|outputColor = make_float3(__int_as_float(rayPayload.primitiveIndex + 1), rayPayload.t, rayPayload.u);|" "$K" \
  > "$OUT/resources/kernels/cuda/id_dump_a.cu"
sed "s|$PAT|outputColor = make_float3(__int_as_float(rayPayload.hitType + 1), rayPayload.v, rayPayload.t);|" "$K" \
  > "$OUT/resources/kernels/cuda/id_dump_b.cu"
grep -q '__int_as_float(rayPayload.primitiveIndex' "$OUT/resources/kernels/cuda/id_dump_a.cu"
grep -q '__int_as_float(rayPayload.hitType' "$OUT/resources/kernels/cuda/id_dump_b.cu"
# the reference's own custom_kernel example, compiled UNCHANGED against this repo's headers and library
# (drop-in check, run by tests/test_gpu_host_api.py on the GPU box)
ROOT="$(cd "$HERE/.." && pwd)"
if [ -f "$ROOT/lens_trace_b200/liblenstrace.so" ]; then
  mkdir -p "$OUT/bin"
  /usr/bin/g++ -I"$ROOT/include" "$REF/examples/custom_kernel/src/main.cpp" -o "$OUT/bin/ref_custom_kernel" \
    -L"$ROOT/lens_trace_b200" -llenstrace -llt_b200 -Wl,-rpath,'$ORIGIN/../../../lens_trace_b200'
  /usr/bin/g++ -DCUDA_ENABLED -DOPENCL_ENABLED -I"$ROOT/include" "$REF/src/main.cpp" -o "$OUT/bin/ref_LensTrace" \
    -L"$ROOT/lens_trace_b200" -llenstrace -llt_b200 -Wl,-rpath,'$ORIGIN/../../../lens_trace_b200'
fi
echo "built $OUT/libltref.so"
