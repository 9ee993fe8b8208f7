#!/usr/bin/env python3
"""Static SASS opcode histogram per kernel of libatde_b200.so (cuobjdump -sass): instruction count, code bytes and the
opcodes that prove what the hardware executes — packed fp32 (FMUL2 / FFMA2), bulk async copies (UBLKCP = cp.async.bulk),
mbarrier transactions (SYNCS), warp shuffles, shared / global / local memory traffic, IEEE division fix-up calls.
usage: sass_histogram.py [libatde_b200.so] > profiles/rN_sass_histogram.txt"""
import re, subprocess, sys
from collections import Counter, OrderedDict
so = sys.argv[1] if len(sys.argv) > 1 else "atracdenc_b200/libatde_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["FMUL2", "FFMA2", "FADD2", "UBLKCP", "SYNCS", "UTMA", "SHFL", "REDUX", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "CALL", "ATOMS", "DADD", "DMUL", "DFMA", "MUFU"]
print(f"{'kernel':48s} {'instrs':>7s} {'KiB':>6s}  " + " ".join(f"{k:>6s}" for k in KEY))
for k, h in hist.items():
    n = sum(h.values())
    fam = Counter()
    for op, c in h.items():
        base = op.split(".")[0]
        for key in KEY:
            if base.startswith(key):
                fam[key] += c
                break
    print(f"{k[-48:]:48s} {n:7d} {n * 16 / 1024:6.1f}  " + " ".join(f"{fam[key]:6d}" for key in KEY))
print()
for k, h in hist.items():
    n = sum(h.values())
    top = ", ".join(f"{op} {c}" for op, c in h.most_common(14))
    print(f"{k}: {n} instructions\n    {top}")
