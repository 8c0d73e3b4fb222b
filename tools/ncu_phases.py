#!/usr/bin/env python
"""Per-phase warp-instruction share and lane efficiency of vmis_predict_kernel from an .ncu-rep.
Usage: tools/ncu_phases.py gpurun_out/prof.ncu-rep <queries in the profiled launch>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, nq = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 65536
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", ROOT + "/serenade_b200/libvmis_b200.so"], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("predict_sm100.")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, capture_output=True, text=True).stdout
# nvdisasm prints inlined call chains: "//## File ..., line N inlined at ..."; the LAST plain File line before an
# instruction is the innermost location; we attribute header code to the enclosing predict_sm100.cu line.
cur, off2line = None, {}
for l in dis.splitlines():
    mm = re.search(r'//## File "(.*?)", line (\d+)(.*)', l)
    if mm:
        if mm.group(1).endswith("predict_sm100.cu"):
            cur = int(mm.group(2))
        else:
            inl = re.search(r'inlined at "(.*?predict_sm100\.cu)", line (\d+)', mm.group(3))
            cur = int(inl.group(2)) if inl else cur
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if mm:
        off2line[int(mm.group(1), 16)] = cur
srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(srcp))); h = rows[1]
ia, iinst, ith, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
iwf, iwfi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
stall_cols = {k: h.index(k) for k in ("stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_mio", "stall_math")}
def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0
base = int(rows[2][ia], 16)
src = open(ROOT + "/serenade_b200/csrc/predict_sm100.cu").read().splitlines()
def find(s):
    if s.startswith("phase"):     # code markers look like "// ------ phase 2a: ..."; the header comment also names phases
        return next(i + 1 for i, l in enumerate(src) if re.search(r"// -{10,} " + re.escape(s), l))
    return next(i + 1 for i, l in enumerate(src) if s in l)
marks = [("helpers (sort/scan/hash)", 1), ("accumulate (phase 2b)", find("// phase 2b: A[item] += w")), ("compact_slots (global table)", find("// compaction of the occupied slots")),
         ("select helpers (u32 net, exact_elem)", find("constexpr int kIdxBits")),
         ("select_exact (phase 3, rare)", find("// phase 3, exact path")),
         ("select_table (phase 3)", find("// phase 3 on the shared table")),
         ("kernel prologue", find("vmis_predict_kernel(const IndexView")),
         ("phase 0", find("phase 0")), ("phase 1 merge", find("phase 1")), ("phase 1b top-k", find("phase 1b")),
         ("neighbours mode", find("if (neighbors_mode) {")), ("phase 2a directory", find("phase 2a")),
         ("phase 2b+3 driver", find("phase 2b + 3")), ("end", find("uint32_t next_pow2"))]
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
stalls = collections.defaultdict(lambda: collections.Counter())
for r in rows[2:]:
    ln = off2line.get(int(r[ia], 16) - base)
    b = "unattributed"
    if isinstance(ln, int):
        for name, start in marks:
            if ln >= start:
                b = name
    agg[b][0] += int(r[iinst]); agg[b][1] += int(r[ith]); agg[b][2] += int(r[isamp]); agg[b][3] += num(r[iwf]); agg[b][4] += num(r[iwfi])
    for k, ci in stall_cols.items():
        stalls[b][k] += num(r[ci])
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print(f"# {os.path.basename(rep)}: {tot / nq:.0f} warp instructions per query")
twf = sum(v[3] for v in agg.values())
for b, (i, t, s, wf, wfi) in sorted(agg.items(), key=lambda x: -x[1][0]):
    top = ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in stalls[b].most_common(3))
    print(f"{b:38s} inst {100 * i / tot:5.1f}%  warp-inst/query {i / nq:7.0f}  lane-eff {t / (32 * max(i,1)):.2f}  stall-samples {100 * s / ts:5.1f}%"
          f"  smem-wavefronts/query {wf / nq:6.0f} (ideal {wfi / nq:5.0f})  [{top}]")
print(f"# shared-memory wavefronts per query: {twf / nq:.0f}")
