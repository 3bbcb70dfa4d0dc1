#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + stall/opcode breakdown from the source page.
Usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} [{units[i]}] = {r[i][:100]}")
    top = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:8]
    for v, h in top:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.2f}")
args = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kre:
    args += ["-k", f"regex:{kre}"]
src = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# the source page repeats a header per kernel
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h = b["rows"][0]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    cs, ce = Counter(), Counter()
    for r in b["rows"][1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0].rstrip(";")
        cs[op] += int(r[iN])
        ce[op] += int(r[iE])
    ts, te = sum(cs.values()) or 1, sum(ce.values()) or 1
    print("-" * 100)
    print(b["name"][:90], "instr", te, "samples", ts)
    for op, _ in ce.most_common(14):
        print(f"   {op:10s} exec {ce[op] / te * 100:5.1f}%   samples {cs[op] / ts * 100:5.1f}%")
