#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into a small text summary for profiles/: per launch the duration, pipe
utilisation, DRAM traffic, occupancy limits and stall mix, plus the hottest SASS instructions of the first launch.
usage: ncu_summary.py report.ncu-rep > profiles/summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("# ncu summary of %s (ncu --set full --clock-control none; per-launch values)" % rep.split("/")[-1])
for h in want + stall:
    if h in hdr:
        i = hdr.index(h)
        name = h.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")
        print("%-72s %-12s %s" % (name, units[i], "  ".join(r[i][:40] for r in data)))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h2, tab = None, []
for r in rows:
    if r and r[0] == "Address":
        if h2 is None:
            h2 = r
            continue
        break
    if h2 and len(r) == len(h2):
        tab.append(r)
if h2:
    ix = {h: i for i, h in enumerate(h2)}
    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, "# Samples") for r in tab) or 1.0
    print("\n# hottest SASS instructions of the first launch (share of warp-stall samples)")
    for r in sorted(tab, key=lambda r: -f(r, "# Samples"))[:25]:
        print("%5.1f%%  %s" % (100 * f(r, "# Samples") / tot, r[ix["Source"]][:100]))
    mix = {}
    for r in tab:
        op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
        if op:
            mix[op[0].split(".")[0]] = mix.get(op[0].split(".")[0], 0) + f(r, "Instructions Executed")
    t2 = sum(mix.values()) or 1.0
    print("\n# executed instruction mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / t2) for k, v in sorted(mix.items(), key=lambda x: -x[1])[:12]))
