"""Offline view of a kernel's hot loops in a built .so (cuobjdump -sass): every backward-branch loop that contains
activations (MUFU.EX2), its instruction count per row of 4 pixels, an opcode histogram and a pipe-cycle estimate.

    python tools/sass_loops.py <lib.so> [mangled function name] [--hist]

Pipe-cycle model (cycles a warp-instruction holds its pipe on one SM sub-partition; measured on B200 with
tools/microbench/pipe_bench.cu):
    FFMA2 / FADD2 / FMUL2 : max(2, number of DISTINCT register-pair operands)   (3 distinct pairs -> 3.0, else 2.0)
    scalar FP32 (FFMA/FADD/FMUL): 1.5 with 3 distinct register operands, else 1
    HFMA2 (bf16x2 / f16x2): 2        MUFU: 8 (own pipe, overlaps the FMA pipe)
Everything else is counted as one issue slot only.
"""
import re
import subprocess
import sys
from collections import Counter

DEFAULT = "_ZN3pil14pil_bwd_kernelILi1EffLb1EEEvNS_7BwdArgsE"


def disasm(so, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, so], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def strip_pred(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t)


def fma_pipe_cycles(t):
    t = strip_pred(t)
    op = t.split()[0]
    base = op.split(".")[0]
    if base in ("FFMA2", "FADD2", "FMUL2"):
        srcs = t.split(",")[1:]
        regs = {re.match(r"\s*-?\|?(R\d+)", s).group(1) for s in srcs if re.match(r"\s*-?\|?R\d+", s)}
        return max(2.0, float(len(regs)))
    if base in ("FFMA", "FADD", "FMUL"):
        srcs = t.split(",")[1:]
        regs = {re.match(r"\s*-?\|?(R\d+)", s).group(1) for s in srcs if re.match(r"\s*-?\|?R\d+", s)}
        return 1.5 if len(regs) >= 3 else 1.0
    if base == "HFMA2":
        return 2.0
    if base in ("IMAD", "FSETP", "FSET", "FMNMX", "FSEL") and base == "IMAD":
        return 1.0
    return 0.0


def loops(ins):
    for a, t in ins:
        m = re.search(r"BRA\s+(?:U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [tt for x, tt in ins if tgt <= x <= a]
            ex = sum(1 for tt in body if "MUFU.EX2" in tt)
            if ex >= 4:
                yield tgt, a, body, ex // 4


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    so = args[0]
    fun = args[1] if len(args) > 1 else DEFAULT
    ins = disasm(so, fun)
    print(f"{so}: {len(ins)} instructions in {fun[:60]}")
    for tgt, a, body, rows in loops(ins):
        mov = sum(1 for t in body if re.match(r"(MOV|IMAD\.MOV)", strip_pred(t)))
        fma = sum(fma_pipe_cycles(t) for t in body)
        mufu = 8.0 * sum(1 for t in body if strip_pred(t).startswith("MUFU"))
        print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, {rows} rows -> {len(body) / rows:.1f} instr/row "
              f"(moves {mov}); model cycles/row: issue {len(body) / rows:.0f}, fma pipe {fma / rows:.0f}, mufu {mufu / rows:.0f}")
        if "--hist" in sys.argv:
            ops = Counter(strip_pred(t).split()[0] for t in body)
            for k, v in ops.most_common():
                print(f"      {v:4d} {v / rows:6.2f}/row {k}")
            p3 = sum(1 for t in body if strip_pred(t).split()[0].split(".")[0] in ("FFMA2", "FADD2", "FMUL2") and fma_pipe_cycles(t) >= 3)
            print(f"      packed ops with 3 distinct register-pair operands: {p3}")


if __name__ == "__main__":
    main()
