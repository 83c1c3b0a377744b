#!/usr/bin/env python
"""Dev helper: decode scoreboard control bits (write/read barrier index, wait mask, stall) of a kernel's SASS."""
import re, subprocess, sys
so, name = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
f = [x for x in funcs if name in x.split("\n")[0]][0]
lines = f.split("\n")
for i, line in enumerate(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", line)
    if not m: continue
    a = int(m.group(1), 16)
    if not (lo <= a <= hi): continue
    m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
    h = int(m2.group(1), 16)
    stall = (h >> 41) & 0xF; wbar = (h >> 46) & 7; rbar = (h >> 49) & 7; wait = (h >> 52) & 0x3F
    t = m.group(2).strip()
    if wbar != 7 or rbar != 7 or wait:
        print(f"{a:05x} W{wbar if wbar!=7 else '-'} R{rbar if rbar!=7 else '-'} wait={wait:06b} st={stall:2d}  {t}")
