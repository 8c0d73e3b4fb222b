#!/usr/bin/env python
"""Opcode histogram of vmis_predict_kernel in the CURRENT build of libvmis_b200.so (cuobjdump -sass; no GPU needed):
the evidence that the sm_100a path uses TMA bulk copies + mbarriers (UBLKCP, SYNCS), 64-bit shared-memory atomics
(ATOMS.CAS.64) and warp reductions (REDUX), plus the share of control-flow instructions.
Usage: tools/sass_histogram.py > profiles/<round>_sass_vmis_predict_kernel.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "serenade_b200", "libvmis_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
arch = re.search(r"arch = (sm_\w+)", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if "vmis_predict_kernel" not in name:
        continue
    ops = collections.Counter()
    full = collections.Counter()
    for line in b.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
        if m:
            op = m.group(1)
            ops[op.split(".")[0]] += 1
            full[op] += 1
    tot = sum(ops.values())
    print(f"# {os.path.relpath(so, ROOT)}  ({arch.group(1) if arch else '?'})")
    print(f"# kernel {name}: {tot} SASS instructions (static)")
    ctl = sum(ops[o] for o in ("BRA", "BSSY", "BSYNC", "BREAK", "WARPSYNC", "EXIT", "CALL", "RET"))
    print(f"# control flow (BRA/BSSY/BSYNC/BREAK/WARPSYNC/EXIT/CALL/RET): {ctl} = {100 * ctl / tot:.1f} %")
    print("# sm_100a features:")
    for key in ("UBLKCP", "SYNCS", "ATOMS", "ATOMG", "REDUX", "MATCH", "VOTE", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "CCTL", "NANOSLEEP"):
        sub = {k: v for k, v in full.items() if k.split(".")[0] == key}
        if sub:
            print(f"  {key:9s} {sum(sub.values()):5d}   " + ", ".join(f"{k} x{v}" for k, v in sorted(sub.items(), key=lambda x: -x[1])[:6]))
    print("# full histogram (base opcode: count)")
    for op, c in ops.most_common():
        print(f"  {op:12s} {c:6d}  {100 * c / tot:5.1f} %")
