export LT_PROFILE_NOREF=1
run() { timeout 200 python tools/profile_flat.py "$@" >> gpurun_out/r2p_threaded_big.log 2>&1 || echo "FAILED: $*" >> gpurun_out/r2p_threaded_big.log; }
rm -f gpurun_out/r2p_threaded_big.log
run synth:707
LT_THREADED_MAX_NODES=3000000 run synth:707
LT_PROFILE_FLAGS=64 run synth:707
LT_DOWNLOAD_THREADS=8 timeout 200 python -c "
import bench
print('render call ms, 8 copier threads:', bench.this_repo_render_call_ms(1920,1080))" >> gpurun_out/r2p_threaded_big.log 2>&1
LT_DOWNLOAD_THREADS=2 timeout 200 python -c "
import bench
print('render call ms, 2 copier threads:', bench.this_repo_render_call_ms(1920,1080))" >> gpurun_out/r2p_threaded_big.log 2>&1
timeout 200 python -c "
import bench
print('render call ms, 4 copier threads:', bench.this_repo_render_call_ms(1920,1080))" >> gpurun_out/r2p_threaded_big.log 2>&1
cat gpurun_out/r2p_threaded_big.log
