"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small tracked text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r01.ncu-rep profiles/r01_ncu_full_cfg3.md [--title "..."]

Reads the report with `ncu -i ... --page raw --csv` and keeps the metrics the roofline discussion uses:
duration, DRAM bytes (the `roofline.traffic` figure of bench.py), DRAM / L2 / SM throughput, occupancy,
issue utilisation and the instruction count.
"""
import csv
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes.sum.per_second", "DRAM throughput"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of hw peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.max", "SM elapsed cycles"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
    ("smsp__warps_active.avg.per_cycle_active", "active warps / scheduler"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles / issued inst"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe instructions %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe instructions %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe instructions %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe instructions %"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--title" else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    lines = [f"# {title}", "", f"source report: `{rep}` (scratch, not tracked); extracted with tools/ncu_summary.py", ""]
    for r in rows[2:]:
        lines.append(f"## {r[ik]}")
        lines.append("")
        lines.append("| metric | ncu name | value | unit |")
        lines.append("|---|---|---|---|")
        for name, label in KEEP:
            if name in hdr:
                i = hdr.index(name)
                lines.append(f"| {label} | `{name}` | {r[i]} | {units[i]} |")
        # warp stall breakdown (top 6)
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            lines.append("")
            lines.append("warp stall reasons (warps stalled per issue-active cycle, top 6): " +
                         ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
