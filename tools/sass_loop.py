#!/usr/bin/env python
"""Dev helper: dump one kernel's SASS from the built .so, find its hottest backward-branch loop (the one
holding LDG.E.128) and print the opcode histogram and pipe-time estimate per loop iteration."""
import re, subprocess, sys, collections
so = sys.argv[1]; name = sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
f = [x for x in funcs if x.split("\n")[0].strip().find(name) >= 0]
if not f: sys.exit("function not found: " + "\n".join(x.split("\n")[0] for x in funcs[1:]))
f = f[0]
ins = []
for line in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
# backward branches
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
best = None
for lo, hi in loops:
    body = [t for a, t in ins if lo <= a <= hi]
    n128 = sum(("LDG.E.128" in t) or ("LDG.E.64" in t) for t in body)
    if n128 >= 8 and (best is None or len(body) < best[2]):
        best = (lo, hi, len(body), body)
lo, hi, n, body = best
print(f"{f.splitlines()[0].strip()}\nloop 0x{lo:x}..0x{hi:x}: {n} instructions, {sum(("LDG.E.128" in t) or ("LDG.E.64" in t) for t in body)} wide LDG")
h = collections.Counter()
for t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    h[t.split()[0].split(".")[0]] += 1
fma2 = h["FFMA2"] + h["FMUL2"] + h["FADD2"]
fma1 = h["FFMA"] + h["FMUL"] + h["FADD"]
print("  ".join(f"{k}:{v}" for k, v in h.most_common()))
print(f"packed fp {fma2}, scalar fp {fma1}, IMAD {h['IMAD']} -> FMA-pipe clks ~{2*fma2 + fma1 + 2*h['IMAD']} per iteration; "
      f"XU-ish (MUFU/F2I/FRND/I2F*) {h['MUFU']+h['F2I']+h['FRND']+h['I2F']+h['I2FP']}; local mem {h['LDL']+h['STL']}")
if len(sys.argv) > 3:
    for a, t in ins:
        if lo <= a <= hi: print(f"{a:05x} {t}")
