#!/usr/bin/env python
"""Static SASS instruction count of vmis_predict_kernel per predict_sm100.cu line range (no GPU needed).
Usage: tools/sass_static.py line_lo line_hi [--list] [--lib path]"""
import os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lo, hi = int(sys.argv[1]), int(sys.argv[2])
lib = ROOT + "/serenade_b200/libvmis_b200.so"
if "--lib" in sys.argv: lib = sys.argv[sys.argv.index("--lib") + 1]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("predict_sm100.")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, capture_output=True, text=True).stdout
cur = None; per = collections.Counter(); rows = []
for l in dis.splitlines():
    mm = re.search(r'//## File "(.*?)", line (\d+)(.*)', l)
    if mm:
        if mm.group(1).endswith("predict_sm100.cu"): cur = int(mm.group(2))
        else:
            inl = re.search(r'inlined at "(.*?predict_sm100\.cu)", line (\d+)', mm.group(3)); cur = int(inl.group(2)) if inl else cur
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if mm and isinstance(cur, int) and lo <= cur <= hi:
        per[cur] += 1; rows.append((mm.group(1), cur, mm.group(2)))
if "--list" in sys.argv:
    for a, c, t in rows: print(a, f"L{c}", t)
src = open(ROOT + "/serenade_b200/csrc/predict_sm100.cu").read().splitlines()
for ln in sorted(per): print(f"L{ln}: {per[ln]:4d}  | {src[ln-1].strip()[:110]}")
print("total", sum(per.values()))
