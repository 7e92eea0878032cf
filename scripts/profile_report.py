#!/usr/bin/env python
"""Turn the raw artefacts of scripts/gpu_check.sh (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/profile_report.py r01e 35143680

writes profiles/<tag>_launches.csv (the ncu gpu__time_duration launch list), profiles/<tag>_gn.txt and
profiles/<tag>_grad.txt (key ncu --set full metrics + executed-instruction mix + stall sites of the fused
kernel) and profiles/<tag>_bench.jsonl (the bench lines of the same visit)."""
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
npts = sys.argv[2] if len(sys.argv) > 2 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
src = os.path.join(root, "gpurun_out")
os.makedirs(out, exist_ok=True)


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


if os.path.exists(f"{src}/launches_{tag}.csv"):
    shutil.copy(f"{src}/launches_{tag}.csv", f"{out}/{tag}_launches.csv")
traffic = {}
for mode in ("gn", "grad"):
    rep = f"{src}/prof_{mode}_{tag}.ncu-rep"
    if not os.path.exists(rep):
        continue
    summ = run([sys.executable, f"{root}/scripts/ncu_summary.py", rep])
    csvp = f"/tmp/{tag}_{mode}_src.csv"
    with open(csvp, "w") as f:
        f.write(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))
    hot = run([sys.executable, f"{root}/scripts/sass_hot.py", csvp] + ([npts] if npts else []))
    with open(f"{out}/{tag}_{mode}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, kernel k_align_global ({mode} mode), bench.py --pairs 64\n")
        f.write(f"# source: gpurun_out/prof_{mode}_{tag}.ncu-rep (not tracked)\n\n")
        f.write(summ + "\n" + hot)
    rd = wr = None
    for line in summ.splitlines():
        if "dram__bytes_read.sum" in line:
            v, unit = line.split()[-2:]
            rd = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
        if "dram__bytes_write.sum" in line:
            v, unit = line.split()[-2:]
            wr = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
    if rd is not None and wr is not None:
        traffic[f"{mode}_bytes_per_launch_64pairs"] = rd + wr
if traffic and "_" not in tag:      # variant visits (<variant>_<tag>) never replace the default library's traffic figure
    traffic["source"] = f"ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, visit {tag}"
    sys.path.insert(0, root)
    import __graft_entry__ as entry
    traffic["lib_hash"] = entry.source_hash()      # bench.py reports the figure only for this exact source tree
    json.dump(traffic, open(f"{out}/traffic.json", "w"), indent=1)
lines = []
for name in sorted(os.listdir(src)):
    if name.startswith("bench_") and name.endswith(f"_{tag}.json"):
        txt = open(os.path.join(src, name)).read().strip()
        if txt:
            lines.append(txt.splitlines()[-1])
if lines:
    open(f"{out}/{tag}_bench.jsonl", "w").write("\n".join(lines) + "\n")
for extra in (f"pytest_{tag}.log", f"smoke_{tag}.log", f"gpu_{tag}.txt"):
    if os.path.exists(f"{src}/{extra}"):
        shutil.copy(f"{src}/{extra}", f"{out}/{tag}_{extra.replace('_' + tag, '')}")
print("wrote", sorted(n for n in os.listdir(out) if n.startswith(tag)))
