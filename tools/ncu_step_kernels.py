#!/usr/bin/env python3
"""Per-kernel totals of one bench step from an ncu launch list taken with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
(exact kernels only: the <true> template instances are the counting launches).  Writes/updates
profiles/traffic.json -> "kernels_<workload>" (read by bench.py for the per-kernel DRAM rates) and prints a table.

    python tools/ncu_step_kernels.py gpurun_out/launches.csv cornell_gi_1080p_64spp [profiles/r2_step_kernels.txt]
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, workload = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) if len(r) >= 15]
    hdr, rows = rows[0], rows[1:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per_launch = defaultdict(dict)
    names = {}
    for r in rows:
        per_launch[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
        names[r[ii]] = r[ki]
    agg = defaultdict(lambda: {"launches": 0, "ns": 0.0, "dram": 0.0})
    steps = 0
    for lid, m in per_launch.items():
        n = names[lid]
        base = re.sub(r"^void ", "", n).split("(")[0]
        if not base.startswith("k_wf_") and not base.startswith("k_path") and not base.startswith("k_flat"):
            continue
        if re.match(r"k_wf_(trace|primary)<\(?(bool\))?1", base) or base.startswith("k_wf_trace<1") or base.startswith("k_wf_primary<1"):
            continue  # counting instances (LT_FLAG_STATS)
        key = base.split("<")[0]
        a = agg[key]
        a["launches"] += 1
        a["ns"] += m.get("gpu__time_duration.sum", 0.0)
        a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        if key == "k_wf_primary_trace":
            steps += 1
    steps = max(1, steps)
    total_ns = sum(a["ns"] for a in agg.values())
    out = {}
    lines = ["%d steps in the launch list (one k_wf_primary_trace each); per step, per kernel:" % steps,
             "%-22s %9s %10s %8s %14s %10s" % ("kernel", "launches", "ms", "share", "dram MB", "GB/s")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        ms = a["ns"] / steps / 1e6
        out[k] = {"launches_per_step": a["launches"] / steps, "ms_per_step_under_ncu": ms,
                  "share_of_device_time": a["ns"] / total_ns, "dram_bytes_per_step": a["dram"] / steps}
        lines.append("%-22s %9.1f %10.3f %7.1f%% %14.1f %10.1f" % (k, a["launches"] / steps, ms, 100 * a["ns"] / total_ns,
                                                              a["dram"] / steps / 1e6, a["dram"] / max(a["ns"], 1.0)))
    print("\n".join(lines))
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    tj["kernels_" + workload] = out
    tj["kernels_" + workload + "_source"] = os.path.basename(path)
    json.dump(tj, open(tpath, "w"), indent=1)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
