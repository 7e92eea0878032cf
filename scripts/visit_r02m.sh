#!/bin/bash
# visit r02m (1 GPU): full test suite after the fixes, headline bench (e2e params path fixed), c4 + compaction with the
# vectorised mask passes and the device-side tile table, ncu captures of the default library
TAG=${1:-r02m}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/pytest_$TAG.log 2>&1
grep -E "sign flips|passed|failed|FAILED|Error" $OUT/pytest_$TAG.log | tail -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_gn_$TAG.json 2> $OUT/bench_gn_$TAG.err
python - $OUT/bench_gn_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d["e2e"]; o = d["other_iteration"]
print("GN frac %.3f value %.0f | grad frac %.3f | e2e %.0f (link %.1f GB/s, frac %.2f) target-only %.0f (frac %.2f) | dropin %.0f it/s | cpu %.2f" % (
    d["roofline"]["frac"], d["value"], o["roofline_frac"], e["value"], e["h2d_link_GBps"], e["frac_of_link"],
    e["target_frame_only"]["value"], e["target_frame_only"]["frac_of_link"], d["dropin_single_pair"]["iters_per_s"], d["cpu_baseline"]["value"]))
PY
timeout 300 python bench.py --steps 30 --warmup 5 --mode grad --no-cpu-baseline --no-e2e > $OUT/bench_grad_$TAG.json 2> $OUT/bench_grad_$TAG.err
for w in compaction c4; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_${w}_$TAG.json 2> $OUT/bench_${w}_$TAG.err
  tail -c 900 $OUT/bench_${w}_$TAG.json; echo; tail -3 $OUT/bench_${w}_$TAG.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_lm|k_window|k_ingest" -c 60 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_row_fill|k_row_count|k_row_scan" -c 3 -f -o $OUT/prof_compact_$TAG \
    python bench.py --workload compaction > $OUT/ncu_compact_$TAG.log 2>&1
ls $OUT | wc -l
