#!/usr/bin/env python
"""Condense an ncu traffic pass (gpu_profile.sh step 2) of one bench step into the JSON bench.py reads for
`roofline.traffic`: DRAM bytes per tcgen05-GEMM launch, GEMM share of the kernel time, time-weighted tensor-pipe activity.
usage: traffic_summary.py <traffic.csv> <out.json> [source note]"""
import csv
import json
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else path
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
per = defaultdict(dict)
names = {}
for r in csv.DictReader(lines):
    per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "") or 0)
    names[r["ID"]] = r["Kernel Name"]
g_bytes = g_us = all_us = tp = 0.0
n_gemm = 0
for i, m in per.items():
    us = m.get("gpu__time_duration.sum", 0.0) / 1e3
    all_us += us
    if "gemm_tc" in names[i]:
        n_gemm += 1
        g_us += us
        g_bytes += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        tp += us * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
json.dump({"source": note, "gemm_launches_per_step": n_gemm, "gemm_dram_bytes_per_step": g_bytes,
           "gemm_dram_bytes_per_launch": g_bytes / max(n_gemm, 1), "gemm_kernel_us_per_step_under_ncu": g_us,
           "all_kernel_us_per_step_under_ncu": all_us,
           "gemm_tensor_pipe_active_pct_time_weighted": tp / max(g_us, 1e-9)}, open(out, "w"), indent=1)
print(open(out).read())
