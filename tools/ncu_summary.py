#!/usr/bin/env python3
"""Summarise an .ncu-rep: key metrics, stall ratios, dynamic opcode mix (reads raw + source pages)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    for n in want:
        if n in h:
            print(f"{n[-64:]:66s} {r[h.index(n)][:60]} {rows[1][h.index(n)]}")
    st = sorted(((float(r[h.index(n)] or 0), n.split("stalled_")[1].split("_per_issue")[0]) for n in stalls), reverse=True)
    print("stalls/issue: " + "  ".join(f"{n}={v:.2f}" for v, n in st[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None; hdr = None; data = collections.OrderedDict()
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        kern = r[1] + f"#{len(data)}"; data[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) == len(hdr):
        data[kern].append(r)
for k, rs in data.items():
    ie = hdr.index("Instructions Executed"); ss = hdr.index("Warp Stall Sampling (All Samples)"); sc = hdr.index("Source")
    agg = collections.Counter(); st = collections.Counter(); tot = tots = 0
    for r in rs:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[sc]); op = m.group(2) if m else "?"
        n = int(r[ie]); s_ = int(r[ss]); agg[op] += n; st[op] += s_; tot += n; tots += s_
    print("-" * 100); print(k[:70], "warp-inst", tot, "samples", tots)
    print("  inst : " + "  ".join(f"{o}={v/tot:.1%}" for o, v in agg.most_common(16)))
    print("  stall: " + "  ".join(f"{o}={v/max(tots,1):.1%}" for o, v in st.most_common(12)))
