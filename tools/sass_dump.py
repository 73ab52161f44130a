"""Print the SASS of one address range of a kernel in a .so, optionally without the arithmetic (to see the overhead).
    python tools/sass_dump.py <lib.so> <mangled fun> <lo hex> <hi hex> [--no-math]"""
import re
import subprocess
import sys

so, fun, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, so], capture_output=True, text=True).stdout
math = re.compile(r"\b(FFMA2|FADD2|FMUL2|FADD|FFMA|MUFU|FMUL|SHFL)\b")
for line in out.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and lo <= int(m.group(1), 16) <= hi:
        if "--no-math" in sys.argv and math.search(m.group(2)):
            continue
        print(f"{m.group(1)}  {m.group(2)}")
