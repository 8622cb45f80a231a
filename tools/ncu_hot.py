"""Summarise an ncu report: headline metrics + the source lines (from -lineinfo) that collect the most warp-stall samples.
usage: python tools/ncu_hot.py report.ncu-rep [top_n]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__cluster_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in want:
        if k in d:
            print(f"  {k}: {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = src.splitlines()
# find header line
hi = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.reader(io.StringIO("\n".join(lines[hi:])))
h = next(rd)
ci = {k: i for i, k in enumerate(h)}
samp = []
for r in rd:
    if len(r) < len(h):
        continue
    try:
        n = int(r[ci["Warp Stall Sampling (All Samples)"]])
    except ValueError:
        continue
    reasons = {k[6:]: int(r[i]) for k, i in ci.items() if k.startswith("stall_") and "Not Issued" not in k and r[i].isdigit() and int(r[i]) > 0}
    samp.append((n, r[ci["Address"]], r[ci["Source"]].strip(), reasons))
tot = sum(s[0] for s in samp)
print(f"total samples {tot}")
for n, addr, s, reasons in sorted(samp, reverse=True)[:top]:
    rs = ", ".join(f"{k}:{v}" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * n / tot:5.1f}%  {addr[-5:]}  {s[:70]:70s}  {rs}")
