#!/bin/bash
# The GPU suites under alternative schedules / layouts (every knob keeps the output identical).
# Usage (GPU box): bash tools/run_suite_variants.sh > gpurun_out/suite_variants.log
run() { echo "== $*"; env "$@" timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_host_api.py -q 2>&1 | tail -n 3; }
run LT_WF_OVERLAP=1
run LT_WF_OVERLAP=4
run LT_THREADED_MAX_NODES=0
run LT_WF_SHARED_PRIMARY=0
run LT_WF_ALIVE_LIST=0 LT_WF_REFILL_LANES=1 LT_WF_CHUNK=32
run LT_WAVEFRONT_MIN_PATHS=1 LT_WAVEFRONT_MAX_PATHS=40000
run LT_STREAM_MIN_NODES=0
run LT_DOWNLOAD_THREADS=0
run LT_DOWNLOAD_THREADS=1 LT_DOWNLOAD_STAGED_MIN_BYTES=1
