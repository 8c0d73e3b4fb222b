#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into text: headline metrics + stall samples / instructions per
source line of predict_sm100.cu.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
SRC = os.path.join(ROOT, "serenade_b200", "csrc", "predict_sm100.cu")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
print(f"# ncu summary of {os.path.basename(rep)}")
for k in keys:
    if k in m:
        print(f"{k:90s} {m[k][0]} {m[k][1]}")

# per-source-line aggregation through nvdisasm line info of the built library
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "serenade_b200", "libvmis_b200.so")], cwd=td,
                   capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("predict_sm100.")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, capture_output=True, text=True).stdout
cur, off2line = None, {}
for l in dis.splitlines():
    mm = re.search(r'//## File "(.*?)", line (\d+)', l)
    if mm:
        cur = int(mm.group(2)) if mm.group(1).endswith("predict_sm100.cu") else None
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if mm:
        off2line[int(mm.group(1), 16)] = cur
srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(srcp)))
h = rows[1]
ia, isamp, iinst = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
base = int(rows[2][ia], 16)
agg, tot = collections.defaultdict(lambda: [0, 0]), [0, 0]
for r in rows[2:]:
    ln = off2line.get(int(r[ia], 16) - base)
    agg[ln][0] += int(r[isamp]); agg[ln][1] += int(r[iinst]); tot[0] += int(r[isamp]); tot[1] += int(r[iinst])
src = open(SRC).read().splitlines()
print(f"\n# stall samples / warp instructions per source line (total samples {tot[0]}, warp instructions {tot[1]})")
for ln, (s, i) in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    text = src[ln - 1].strip()[:90] if ln else "<no line info / inlined library code>"
    print(f"L{ln}: samples {100 * s / tot[0]:5.1f}%  inst {100 * i / tot[1]:5.1f}%  | {text}")
