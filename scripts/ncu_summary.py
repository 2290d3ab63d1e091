#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into a small text file for profiles/: headline metrics from the raw
page and the hottest SASS lines from the source page.  Usage: ncu_summary.py <rep> <out.txt> [title]"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit = rows[0], rows[1]
WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct", "gpu__dram_throughput.avg.pct", "sm__pipe_tensor_cycles_active_realtime.avg.pct",
        "sm__inst_executed_pipe_alu_realtime.avg.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__average_warps_issue_stalled"]
lines = [f"# {title}", f"# source: ncu --set full --clock-control none --import-source on  ({rep.split('/')[-1]})", ""]
for r in rows[2:]:
    for h, u, v in zip(hdr, unit, r):
        if any(h == w or h.startswith(w) or ("TriageCompute." + w) in h for w in WANT) and ".per_second" not in h and "peak_sustained " not in h + " ":
            if h.endswith(".max") or h.endswith(".min") or ".max." in h or ".min." in h or ".sum.pct" in h and "tc_wavefronts" not in h:
                continue
            lines.append(f"{h:95s} {u:16s} {v}")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
try:
    h = srows[1]
    iS, iI, iSrc = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Source")
    data = [(int(r[iS] or 0), int(r[iI] or 0), r[iSrc].strip()) for r in srows[2:] if len(r) > iI]
    tot = sum(d[0] for d in data) or 1
    lines.append(f"# hottest SASS lines by warp-stall samples (total {tot})")
    for s, i, t in sorted(data, reverse=True)[:25]:
        lines.append(f"{s:8d} {100.0 * s / tot:5.1f}%  executed {i:10d}  {t[:100]}")
    ops = {}
    for s, i, t in data:
        op = t.split()[1] if t.startswith("@") and len(t.split()) > 1 else (t.split()[0] if t else "")
        ops[op] = ops.get(op, 0) + i
    lines.append("")
    lines.append("# executed warp-instructions by opcode (top 20)")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:20]:
        lines.append(f"{n:12d}  {op}")
except Exception as e:  # source page missing
    lines.append(f"# source page unavailable: {e}")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, len(lines), "lines")
