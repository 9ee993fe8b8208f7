#!/usr/bin/env python3
"""Per-source-line table of one kernel from an `ncu --page source --csv` dump joined with `nvdisasm -g -c` of the cubin
the report was taken from: executed warp-instructions, active lanes per instruction, stall samples (with the top stall
reasons) and shared-memory wavefronts (actual / ideal) per CUDA source line.
usage: line_profile.py <ncu_source.csv> <nvdisasm_g.txt> <kernel-mangled-substring> [top] [--by-instr|--by-samples]"""
import csv, re, sys
from collections import defaultdict
src_csv, sass_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 60
order = "instr" if "--by-instr" in sys.argv else "samples"
line_of, cur, inside = {}, None, False
for ln in open(sass_txt):
    if ln.startswith('//---') and '.text.' in ln:
        inside = kern in ln
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
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
base = int(rows[2][col['Address']], 16)
agg = defaultdict(lambda: defaultdict(float))
tot_i = tot_s = 0
reasons = defaultdict(float)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if r[col['Address']] == 'Address':          # a second launch of the kernel follows: the first one is enough
        break
    key = line_of.get(int(r[col['Address']], 16) - base)
    a = agg[key]
    n = float(r[col['Instructions Executed']] or 0)
    a['i'] += n; tot_i += n
    a['t'] += float(r[col['Thread Instructions Executed']] or 0)
    sm = float(r[col['# Samples']] or 0)
    a['s'] += sm; tot_s += sm
    a['w'] += float(r[col['L1 Wavefronts Shared']] or 0)
    a['wi'] += float(r[col['L1 Wavefronts Shared Ideal']] or 0)
    for st in stalls:
        v = float(r[col[st]] or 0)
        a[st] += v; reasons[st] += v
print(f"kernel {kern}: {int(tot_i)} warp-instructions, {sum(a['t'] for a in agg.values()) / max(tot_i, 1):.1f} lanes/instr, {int(tot_s)} stall samples")
print("all samples by reason: " + ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.0f}%" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:8]))
print(f"{'instr':>12} {'%':>5} {'lanes':>5} {'samp%':>6} {'shWave':>10} {'ideal':>10}  line  top stalls")
keyf = (lambda kv: -kv[1]['i']) if order == "instr" else (lambda kv: -kv[1]['s'])
for k, a in sorted(agg.items(), key=keyf)[:top]:
    ts = sorted(((st[6:], a[st]) for st in stalls), key=lambda kv: -kv[1])[:3]
    print(f"{int(a['i']):12d} {100 * a['i'] / tot_i:5.1f} {a['t'] / max(a['i'], 1):5.1f} {100 * a['s'] / max(tot_s, 1):6.1f} {int(a['w']):10d} {int(a['wi']):10d}  {k}  " + ", ".join(f"{n} {int(v)}" for n, v in ts))
