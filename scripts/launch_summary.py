#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share.
usage: launch_summary.py <launches.csv> [first_kernel_regex] — with a regex, the window is cut to ONE step:
from the first launch matching it to the launch before its next occurrence."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"a8::<unnamed>::|a8::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    if len(sys.argv) > 2:
        pat = re.compile(sys.argv[2])
        hits = [i for i, r in enumerate(rows) if pat.search(r[0])]
        if len(hits) >= 2:
            rows = rows[hits[0]:hits[1]]
    agg = OrderedDict()
    for n, us, _, _ in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms of kernel time\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n[:90]}` | {c} | {us:.0f} | {100 * us / tot:.1f}% |")


if __name__ == "__main__":
    main()
