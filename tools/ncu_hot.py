#!/usr/bin/env python
"""Dev helper: top stall sites of an .ncu-rep source page (SASS view)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; body = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci['# Samples']]) for r in body)
base = int(body[0][0], 16)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ranked = sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:top]
print("total samples", tot)
for r in ranked:
    n = int(r[ci['# Samples']])
    st = sorted(((int(r[ci[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{int(r[0],16)-base:05x} {100*n/tot:5.1f}%  {r[1].strip():60s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  exec={r[ci['Instructions Executed']]}")
