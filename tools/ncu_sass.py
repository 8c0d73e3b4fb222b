#!/usr/bin/env python
"""SASS listing of vmis_predict_kernel with per-instruction executed counts (per query) and stall samples from an
.ncu-rep, annotated with the predict_sm100.cu line.  Usage: tools/ncu_sass.py rep n_queries [line_lo line_hi]"""
import csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, nq = sys.argv[1], int(sys.argv[2])
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", ROOT + "/serenade_b200/libvmis_b200.so"], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("predict_sm100.")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, capture_output=True, text=True).stdout
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
ia, isrc, iinst, ith, isamp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
base = int(rows[2][ia], 16)
tot = sum(int(r[iinst]) for r in rows[2:]); ts = sum(int(r[isamp]) for r in rows[2:])
sel_i = sel_s = 0
for r in rows[2:]:
    off = int(r[ia], 16) - base
    ln = off2line.get(off)
    if isinstance(ln, int) and lo <= ln <= hi:
        inst = int(r[iinst]); sel_i += inst; sel_s += int(r[isamp])
        print(f"{off:6x} L{ln:<4d} {inst / nq:8.1f} lanes {int(r[ith]) / max(inst, 1):5.1f} smp {100 * int(r[isamp]) / ts:5.2f}%  {r[isrc].strip()}")
print(f"# selected: {sel_i / nq:.0f} warp-inst/query ({100 * sel_i / tot:.1f}%), samples {100 * sel_s / ts:.1f}%; total {tot / nq:.0f}")
