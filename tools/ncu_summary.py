#!/usr/bin/env python3
"""One-line-per-launch summary of an ncu report: `python tools/ncu_summary.py rep.ncu-rep` (needs ncu on PATH)."""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "ms"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_cycles_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_inst_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_inst_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_inst_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[ci["Kernel Name"]]
    grid = r[ci.get("Grid Size", ci.get("launch__grid_size", 0))]
    out = [name[:60], f"grid={grid}"]
    for metric, label in WANT:
        if metric in ci:
            v = r[ci[metric]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            out.append(f"{label}={v}{units[ci[metric]] if label in ('ms', 'dram_rd', 'dram_wr', 'sm_clock') else ''}")
    print("  ".join(out))
