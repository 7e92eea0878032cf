#!/usr/bin/env python
"""Offline view of the fused kernel's point loop: builds nothing, reads `cuobjdump -sass` of a library and prints the
instruction mix between the loop's MUFU.EX2 (first instruction group of a point) and its back-edge.
Usage: scripts/sass_loop.py <lib.so> [mangled kernel name substring]"""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "k_align_globalILi1ELi6ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
rows, on = [], False
for l in out.splitlines():
    if "Function :" in l:
        on = want in l
        if on and rows:
            break
    m = re.match(r'\s*/\*([0-9a-f]{4})\*/\s+(.*?);', l)
    if on and m:
        rows.append((int(m.group(1), 16), re.sub(r'\s+', ' ', m.group(2)).strip()))
# the point loop: the backward branch whose body contains the 4 LDG.128 taps
best = None
for i, (a, t) in enumerate(rows):
    m = re.search(r'BRA 0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        body = [r for r in rows if lo <= r[0] <= a]
        if sum('LDG.E.LTC256B.128' in r[1] or 'LDG.E.128' in r[1] for r in body) >= 4:
            if best is None or len(body) < len(best):
                best = body
print(f"{want}: {len(rows)} instructions; point loop {len(best)} instructions "
      f"[{best[0][0]:04x}..{best[-1][0]:04x}]")
ops = Counter()
for a, t in best:
    op = t.split()[1] if t.startswith('@') else t.split()[0]
    ops[op.split('.')[0]] += 1
print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common()))
if len(sys.argv) > 3:
    for a, t in best:
        print(f"{a:04x} {t}")
