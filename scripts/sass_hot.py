#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source sass` output: executed warp-instructions and stall
samples per opcode and the hottest instructions.  Usage: python scripts/sass_hot.py src.csv [n_points]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
npts = float(sys.argv[2]) if len(sys.argv) > 2 else None
kernels = []
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        kernels.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
k = kernels[0]
h = k['hdr']
iS, iE, iT, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
ops, samp = Counter(), Counter()
tot = 0
for r in k['rows']:
    src = r[iS].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    e = int(r[iE] or 0)
    ops[op] += e
    samp[op] += int(r[iSamp] or 0)
    tot += e
print(k['name'][:80], 'warp-instructions', tot, 'static', len(k['rows']))
if npts:
    print(f'  per point: {tot * 32 / npts:.1f} thread-instructions')
ts = sum(samp.values())
for op, e in ops.most_common(28):
    extra = f'{e * 32 / npts:7.1f}/pt' if npts else ''
    print(f'  {op:12s} {e:12d} {100 * e / tot:5.1f}%  {extra}   stall-samples {100 * samp[op] / max(ts, 1):5.1f}%')
print('hottest stall sites:')
hot = sorted(k['rows'], key=lambda r: -int(r[iSamp] or 0))[:18]
for r in hot:
    print(f'  {int(r[iSamp]):6d}  {r[iS].strip()[:100]}')
