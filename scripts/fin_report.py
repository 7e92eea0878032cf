#!/usr/bin/env python
"""Condense one tuning visit of the second GN kernel (k_gn_finalize_solve) into profiles/<tag>_finalize.txt:
per-kernel device time from the ncu launch list, and the stall-sample / executed-instruction split of the kernel
between its __syncthreads() phases (from `ncu --set full --import-source on`, SASS source page).
Usage: python scripts/fin_report.py <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1]
out = [f"# {tag}: second kernel of a GN iteration (finalize + damped solve + retraction), bench.py --pairs 64"]
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{tag}.csv")) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
out.append("== ncu launch list (gpu__time_duration, cold cache, serialised)")
for k, v in d.items():
    out.append(f"  {k:45s} n={len(v):2d}  avg {sum(v) / len(v) / 1000:8.2f} us   min {min(v) / 1000:8.2f} us")
for mode in ("gn", "grad"):
    p = f"gpurun_out/bench_{mode}_{tag}.json"
    if os.path.exists(p):
        j = json.loads(open(p).read().strip().splitlines()[-1])
        out.append(f"== bench ({mode}): value {j['value']:.0f}/s  ms_per_step {j['ms_per_step']:.4f}  fused kernel "
                   f"{j['roofline']['kernel_ms']:.4f} ms  -> rest of the step {1000 * (j['ms_per_step'] - j['roofline']['kernel_ms']):.1f} us")
rep = f"gpurun_out/prof_fin_{tag}.ncu-rep"
if os.path.exists(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[1]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    body = [r for r in rows[2:] if len(r) == len(h)]
    tot = sum(int(r[iN] or 0) for r in body)
    out.append(f"== {rows[0][1][:60]}: {tot} stall samples, {len(body)} static instructions; phases split at BAR.SYNC")
    acc = dyn = start = 0
    for i, r in enumerate(body):
        acc += int(r[iN] or 0)
        dyn += int(r[iE] or 0)
        if "BAR.SYNC" in r[iS] or i == len(body) - 1:
            out.append(f"  instr {start:5d}-{i:5d}  samples {acc:5d} ({100 * acc / max(tot, 1):4.1f}%)  warp-instr/CTA {dyn / 64:8.0f}")
            acc = dyn = 0
            start = i + 1
    out.append("  hottest sites:")
    for i in sorted(range(len(body)), key=lambda i: -int(body[i][iN] or 0))[:8]:
        out.append(f"    {body[i][iN]:>5s}  @{i:5d}  {body[i][iS].strip()[:70]}")
open(f"profiles/{tag}_finalize.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
