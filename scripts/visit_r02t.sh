#!/bin/bash
# visit r02t (1 GPU): column pass of the hole fill in parallel strips; fill tests, sanitizer, c4 figures with holes, ncu of the fill
TAG=${1:-r02t}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fill.py -q -x 2>&1 | tail -5 | tee $OUT/pytest_fill_$TAG.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fill.py -q -x -k "golden_case or batch_equals" > $OUT/sanitizer_fill_$TAG.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" $OUT/sanitizer_fill_$TAG.log | tail -3
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fill.py -q -x -k "golden_case" > $OUT/racecheck_fill_$TAG.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_fill_$TAG.log | tail -3
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err
python - $OUT/bench_c4_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("c4 value %.0f frames/s  per-frame %.0f  hole_fill %s" % (d["value"], d["per_frame_calls"]["value"], json.dumps(d["hole_fill"])))
PY
tail -3 $OUT/bench_c4_$TAG.err
timeout 300 ncu --set full --clock-control none -k regex:"k_fill_rows|k_fill_cols" -c 4 -f -o $OUT/prof_fill_$TAG \
    python -m pytest tests/test_gpu_fill.py -q -x -k vga > $OUT/ncu_fill_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_$TAG.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_lm|k_window|k_ingest" -c 60 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_$TAG.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_gn_$TAG.json 2> $OUT/bench_gn_$TAG.err
tail -c 600 $OUT/bench_gn_$TAG.json
