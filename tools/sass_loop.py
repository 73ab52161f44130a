"""Opcode histogram of the hottest loop (largest backward branch span) of a kernel in libpil.so."""
import re
import subprocess
import sys
from collections import Counter

so, fun = sys.argv[1], sys.argv[2]
px_per_iter = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, so], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
need = int(sys.argv[4]) if len(sys.argv) > 4 else int(px_per_iter)
best = None
for addr, text in ins:
    m = re.search(r"BRA\s+(?:U?P\d,\s*)?(0x[0-9a-f]+)", text)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr:
            nex = sum(1 for a, t in ins if tgt <= a <= addr and "MUFU.EX2" in t)
            if nex >= need and (best is None or addr - tgt < best[1] - best[0]):
                best = (tgt, addr)
lo, hi = best
body = [t for a, t in ins if lo <= a <= hi]
ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
n = len(body)
print(f"loop {lo:#x}..{hi:#x}: {n} instr, {n/px_per_iter:.1f} per px (at {px_per_iter} px/iter)")
for k, v in ops.most_common():
    print(f"  {v:4d} {v/px_per_iter:5.2f}/px {k}")
