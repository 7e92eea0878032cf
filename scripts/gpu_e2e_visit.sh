#!/bin/bash
# End-to-end tuning visit: parity tests, the full bench line, and the e2e arm at several ingest chunk sizes.
# Usage: gpurun --timeout 600 -- 'bash scripts/gpu_e2e_visit.sh <tag>'
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01}
timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $OUT/pytest_$TAG.log; tail -4 $OUT/pytest_$TAG.log
timeout 400 python bench.py --steps 30 --warmup 5 2> $OUT/bench_gn_$TAG.err > $OUT/bench_gn_$TAG.json; tail -3 $OUT/bench_gn_$TAG.err
for c in 4 8 32; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-chunk $c 2> $OUT/bench_c${c}_$TAG.err > $OUT/bench_c${c}_$TAG.json
done
python - $TAG <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/bench_*_%s.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); e = d["e2e"]
        print(f, "value=%.0f e2e=%.0f f32=%.0f packed=%.0f params=%.0f h2d=%d" % (
            d["value"], e["value"], e["frames_f32"]["value"], e["prepacked"]["value"], e["params_only"]["value"], e["h2d_bytes_per_step"]))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
