"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

    python tools/ncu_summary.py <tag>

Reads gpurun_out/launches_<tag>.csv (gpu__time_duration per launch, cold cache, serialised) and
gpurun_out/prof_<tag>.ncu-rep (one --set full capture) and writes profiles/<tag>_launches.csv,
profiles/<tag>_summary.md.
"""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = []

lpath = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(l for l in open(lpath) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    per = OrderedDict()
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.write("id,kernel,grid,block,gpu_time_ns\n")
        for r in rows[1:]:
            name = r[ki].split("(")[0].replace("void ", "")
            ns = float(r[vi].replace(",", ""))
            f.write(f"{r[0]},{name},{r[gi].replace(',', ' ')},{r[bi].replace(',', ' ')},{ns:.0f}\n")
            per.setdefault(name, []).append(ns)
    total = sum(sum(v) for v in per.values())
    lines.append(f"## Launch list `{tag}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold cache, serialised)\n")
    lines.append("| kernel | launches | mean us | share of listed time |\n|---|---|---|---|")
    for k, v in per.items():
        lines.append(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / total:.2f} % |")
    lines.append("")

rpath = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rpath):
    raw = subprocess.run(["ncu", "-i", rpath, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic"]
    for r in rows[2:]:
        lines.append(f"## Full capture `{tag}`: `{r[hdr.index('Kernel Name')].split('(')[0]}` (ncu --set full --clock-control none)\n")
        lines.append("| metric | value | unit |\n|---|---|---|")
        for w in want:
            if w in hdr:
                lines.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
        lines.append("")
    # top stall reasons / hottest source lines
    src = subprocess.run(["ncu", "-i", rpath, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 6]
    if srows:
        h = srows[0]
        si, ci = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
        ei = h.index("Instructions Executed")
        body = [r for r in srows[1:] if r[ci].isdigit()]
        tot = sum(int(r[ci]) for r in body) or 1
        body.sort(key=lambda r: -int(r[ci]))
        lines.append(f"### Hottest SASS lines by warp-stall samples ({tot} samples)\n")
        lines.append("| samples % | instructions executed | SASS |\n|---|---|---|")
        for r in body[:25]:
            lines.append(f"| {100 * int(r[ci]) / tot:.1f} | {r[ei]} | `{r[si].strip()}` |")
        lines.append("")

open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
