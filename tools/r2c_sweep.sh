# knob sweep of the persistent kernels (every run bounded by `timeout`)
export LT_PROFILE_NOREF=1
run() { timeout 120 python tools/profile_flat.py "$@" >> gpurun_out/r2f_sweep.log 2>&1 || echo "FAILED/TIMEOUT: $* [$(env | grep -E '^LT_(STREAM|PATH|PREFETCH)' | tr '\n' ' ')]" >> gpurun_out/r2f_sweep.log; }
rm -f gpurun_out/r2f_sweep.log
run synth:707
LT_STREAM_REVERSE=1 run synth:707
LT_STREAM_NODE_STEPS=4 LT_ITER_TRI_TESTS=2 run synth:707
LT_STREAM_NODE_STEPS=6 LT_ITER_TRI_TESTS=3 run synth:707
LT_STREAM_NODE_STEPS=8 LT_ITER_TRI_TESTS=4 run synth:707
LT_STREAM_NODE_STEPS=8 LT_ITER_TRI_TESTS=2 run synth:707
LT_STREAM_NODE_STEPS=4 LT_ITER_TRI_TESTS=1 run synth:707
LT_STREAM_REVERSE=1 LT_STREAM_BLOCKS_PER_SM=4 run synth:707
LT_STREAM_REVERSE=1 LT_STREAM_BLOCKS_PER_SM=10 run synth:707
cat gpurun_out/r2f_sweep.log | grep "kernel 0"
