"""Condense an `ncu --set full` report into one markdown row per captured launch.
usage: python scripts/ncu_full_extract.py <report.ncu-rep> > profiles/rNN_ncu_full.md   (reads it with `ncu -i ... --page raw --csv`)"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, body = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("duration us", "gpu__time_duration.sum", 1e3), ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("DRAM read MB", "dram__bytes_read.sum", None), ("DRAM write MB", "dram__bytes_write.sum", None),
        ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("regs", "launch__registers_per_thread", 1), ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1)]
units = dict(zip(hdr, rows[1]))


def val(r, key, scale):
    if key not in ix:
        return "-"
    v = r[ix[key]].replace(",", "")
    try:
        f = float(v)
    except ValueError:
        return v
    if scale is None:  # bytes in whatever unit ncu chose
        u = units.get(key, "")
        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
        return f"{f:.1f}"
    f *= scale
    return f"{f:.1f}" if f < 1000 and f != int(f) else f"{f:.0f}"


print("| # | kernel | " + " | ".join(c[0] for c in cols) + " |")
print("|---:|---|" + "---:|" * len(cols))
for n, r in enumerate(body):
    name = r[ix["Kernel Name"]]
    name = name.replace("void ", "").replace("a8::", "").replace("<unnamed>::", "").split("(")[0]
    print(f"| {n} | `{name[:70]}` | " + " | ".join(val(r, k, s) for _, k, s in cols) + " |")
