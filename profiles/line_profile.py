#!/usr/bin/env python3
"""Per-source-line view of one kernel of an .ncu-rep: executed warp-instructions, average active lanes, stall
samples and shared-memory wavefronts, by joining `ncu --page source --csv` (SASS rows) with `nvdisasm -g` line
info of the cubin the report was taken from.  Inlined code is attributed to the innermost source line.
usage: line_profile.py <report.ncu-rep> <kernel name> <cubin> <mangled-substring> [top]"""
import csv, re, subprocess, sys
from collections import defaultdict
rep, kname, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for ln in sass.splitlines():
    if ln.startswith('//---') and '.text.' in ln:
        inside = mangled in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
col = {n: (hdr.index(n) if n in hdr else -1) for n in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed',
                                                         'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal')}
stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
base = int(rows[2][col['Address']], 16)
agg = defaultdict(lambda: [0, 0, 0, 0, 0]); stalls = defaultdict(lambda: defaultdict(int)); tot = [0, 0, 0]
ops = defaultdict(int); allst = defaultdict(int)
for r in rows[2:]:
    if len(r) < len(hdr) or r[col['Address']] == 'Address':
        continue
    off = int(r[col['Address']], 16) - base
    key = line_of.get(off)
    n, t, s = int(r[col['Instructions Executed']]), int(r[col['Thread Instructions Executed']]), int(r[col['# Samples']])
    a = agg[key]
    a[0] += n; a[1] += t; a[2] += s; a[3] += int((r[col['L1 Wavefronts Shared']] if col['L1 Wavefronts Shared'] >= 0 else 0) or 0); a[4] += int((r[col['L1 Wavefronts Shared Ideal']] if col['L1 Wavefronts Shared Ideal'] >= 0 else 0) or 0)
    tot[0] += n; tot[1] += t; tot[2] += s
    for nme, i in stall_cols:
        v = int(r[i] or 0)
        if v:
            stalls[key][nme] += v; allst[nme] += v
    w = r[col['Source']].split()
    ops[w[1] if w[0].startswith('@') else w[0]] += n
print(f"kernel {kname}: {tot[0]} warp-instructions, {tot[1] / max(1, tot[0]):.1f} lanes/instr, {tot[2]} stall samples")
print("all samples by reason:", ", ".join(f"{k[6:]} {100 * v / tot[2]:.0f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
print(f"{'instr':>12} {'%':>5} {'lanes':>5} {'samp%':>6} {'shWave':>10} {'ideal':>10}  line  top stalls")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    st = ", ".join(f"{n[6:]} {v}" for n, v in sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3])
    print(f"{a[0]:12d} {100 * a[0] / tot[0]:5.1f} {a[1] / max(1, a[0]):5.1f} {100 * a[2] / max(1, tot[2]):6.1f} {a[3]:10d} {a[4]:10d}  {k}  {st}")
print('--- by opcode')
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]:
    print(f'{v:14d} {100 * v / tot[0]:5.1f}%  {k}')
