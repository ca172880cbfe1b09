#!/usr/bin/env python
"""Top sampled SASS instructions of one kernel in an `ncu --page source --csv` dump.
usage: ncu_hot.py <src.csv> <kernel-ordinal> [n]"""
import csv, sys
path, which = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lines = open(path).read().split('\n')
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
s = starts[which]; e = starts[which + 1] if which + 1 < len(starts) else len(lines)
print(lines[s][:160])
rows = list(csv.reader(lines[s + 1:e]))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']]) for r in body)
print('total samples', tot, 'instructions', len(body))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][ix['# Samples']]))[:n]:
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{k:5d} {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/max(tot,1):5.1f}%  {r[ix['Source']].strip()[:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
