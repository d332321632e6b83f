# does leaving SM room for the other batch's shade kernel pay?  (bounded runs)
rm -f gpurun_out/r2u_overlap.log
for cfg in "8 8 2" "8 7 2" "8 6 2" "8 5 2" "7 7 2" "6 6 2" "8 6 3" "8 6 4" "8 8 3"; do set -- $cfg
  LT_THREADED_BLOCKS_PER_SM=$1 LT_WF_OVERLAP_TRACE_BLOCKS_PER_SM=$2 LT_WF_OVERLAP=$3 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-cull --no-lbvh --no-protocol > gpurun_out/r2u_tmp.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2u_tmp.json'))
print('threaded_bps $1 overlap_bps $2 streams $3: ms_per_step %.3f  serial %.3f' % (d['ms_per_step'], d['roofline']['pipeline_ms_per_step_in_isolation']))" >> gpurun_out/r2u_overlap.log
done
cat gpurun_out/r2u_overlap.log
