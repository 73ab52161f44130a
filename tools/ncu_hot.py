"""Print the most stall-sampled SASS instructions of a kernel from an .ncu-rep (source page)."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
def toi(v):
    try:
        return int(v)
    except ValueError:
        return 0
data = [(toi(r[isamp]), r[ia].strip(), toi(r[iex]), r) for r in rows[2:] if len(r) > max(isamp, iex, ia)]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for idx, (s, src, ex, r) in enumerate(data):
    if s >= thr:
        extra = " ".join(f"{h.replace('stall_','')}={r[i]}" for i, h in stall_cols if toi(r[i]) * 4 >= s and toi(r[i]) > 0)
        print(f"{idx:5d} {s:6d} {100.0*s/tot:5.1f}% ex={ex:8d}  {src[:70]:70s} {extra}")
