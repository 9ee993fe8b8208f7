#!/usr/bin/env python3
"""Joins an `ncu --page source --csv` SASS dump with `nvdisasm -g` line info and prints executed
warp-instructions per CUDA source line (and per inlined call chain head).
usage: sass_by_line.py <ncu_source.csv> <nvdisasm_g.txt> <kernel-mangled-substring> [top]"""
import csv, re, sys
from collections import defaultdict
src_csv, sass_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: track current line annotation per instruction offset within kernel section
line_of = {}
cur = None; inside = False
for ln in open(sass_txt):
    if ln.startswith('//---') and '.text.' in ln:
        inside = kern in ln
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ai, ii, si = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Source')
base = int(rows[2][ai], 16)
per = defaultdict(int); tot = 0; ops = defaultdict(int)
for r in rows[2:]:
    off = int(r[ai], 16) - base
    n = int(r[ii]); tot += n
    key = line_of.get(off, (None, ''))[0]
    per[key] += n
    ops[r[si].split()[0] if not r[si].strip().startswith('@') else r[si].split()[1]] += n
print('total warp-instructions', tot)
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    print(f'{v:14d} {100*v/tot:5.1f}%  {k}')
print('--- by opcode')
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
    print(f'{v:14d} {100*v/tot:5.1f}%  {k}')
