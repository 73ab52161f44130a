"""List every backward-branch loop of a kernel in a .so that contains activations (MUFU.EX2), with its
instruction count per row of 4 pixels -- an offline proxy for the hot-loop quality of a build.

    python tools/sass_loops.py <lib.so> [mangled function name]
"""
import re
import subprocess
import sys

so = sys.argv[1]
fun = sys.argv[2] if len(sys.argv) > 2 else "_ZN3pil14pil_bwd_kernelILi1EffLb1EEEvNS_7BwdArgsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, so], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{so}: {len(ins)} instructions in {fun[:40]}...")
for a, t in ins:
    m = re.search(r"BRA\s+(?:U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a:
            body = [tt for x, tt in ins if tgt <= x <= a]
            ex = sum(1 for tt in body if "MUFU.EX2" in tt)
            if ex >= 4:
                mov = sum(1 for tt in body if re.match(r"(@!?P\d\s+)?(MOV|IMAD\.MOV)", tt))
                sel = sum(1 for tt in body if "FSEL" in tt or re.match(r"(@!?P\d\s+)?SEL", tt))
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, {ex // 4} rows -> {len(body) / (ex / 4):.1f} instr/row (moves {mov}, selects {sel})")
