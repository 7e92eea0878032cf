#!/bin/bash
# final check of the tree as the driver will run it: GPU tests, smoke, both bench arms
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_$TAG.log 2>&1; tail -3 $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; tail -c 700 $OUT/bench_ref_$TAG.json; echo
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d["e2e"]
print("value %.0f %s | roofline %s | e2e %.0f frac_of_link %.3f | launches %s | clocks %s" % (d["value"], d["unit"], {k: d["roofline"][k] for k in ("frac", "traffic", "achieved", "peak")}, e["value"], e.get("frac_of_link", -1), d["gpu_launches"], d["clocks"]))
PY
