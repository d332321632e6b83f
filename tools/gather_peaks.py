#!/usr/bin/env python3
"""Measures the per-lane 32-byte gather ceilings the node-fetch roofline is rated against (lt_debug_gather_peak,
lens_trace_b200/csrc/lt_microbench.cu) and writes profiles/gather_peaks.json.  Run on the GPU box:

    python tools/gather_peaks.py [out.json]

Table sizes: 21 KB (the Cornell tree's eight threaded copies: L1-resident), 112 MB (traversal set of the
1 M-triangle mesh: L2-resident), 0.9 GB (5 M-triangle mesh: HBM).  For each: independent gathers (4 per lane in
flight) and dependent chains (1 per lane: what one ray does), at 8 persistent blocks per SM like k_wf_trace.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_trace_b200 import capi  # noqa: E402

TABLES = [("l1_21KB", 21 * 1024), ("l2_112MB", 112 * 1000 * 1000), ("hbm_0.9GB", 900 * 1000 * 1000)]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "gather_peaks.json")
    ctx = capi.Context(0)
    res = {"what": "32-byte ld.global.nc.v8.f32 per lane at pseudo-random records; GB/s = 32 B x gathers / time "
                   "(best of 3, CUDA events); 148 x blocks_per_sm persistent blocks of 128 threads",
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "sm_count": ctx.stats().sm_count, "tables": {}}
    for name, nbytes in TABLES:
        iters = 4000 if nbytes < 1e6 else 1000
        entry = {"table_bytes": nbytes}
        for label, dep, ilp, bps in (("independent_ilp4", False, 4, 8), ("independent_ilp1", False, 1, 8),
                                     ("dependent_ilp1", True, 1, 8), ("dependent_ilp2", True, 2, 8),
                                     ("independent_ilp4_16blocks", False, 4, 16)):
            gbs, ns = ctx.gather_peak(nbytes, dependent=dep, ilp=ilp, blocks_per_sm=bps, iters=iters)
            entry[label] = {"gbs": gbs, "ns_per_gather_per_lane": ns, "blocks_per_sm": bps}
        entry["peak_gbs"] = max(v["gbs"] for v in entry.values() if isinstance(v, dict))
        res["tables"][name] = entry
        print(name, json.dumps(entry), flush=True)
    ctx.close()
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
