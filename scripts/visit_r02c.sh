#!/bin/bash
# visit r02c: full GPU test suite, headline bench (with end-to-end arms), the other BASELINE workloads, fused-ingest variant
TAG=${1:-r02c}
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/super_primitive_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
lscpu | head -20 > $OUT/lscpu_$TAG.txt 2>&1; nvidia-smi topo -m >> $OUT/lscpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/pytest_$TAG.log 2>&1
grep -E "GPU vs float64|GN vs Adam|passed|failed|FAILED|Error" $OUT/pytest_$TAG.log | tail -40
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_gn_$TAG.json 2> $OUT/bench_gn_$TAG.err
python - $OUT/bench_gn_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d["e2e"]; o = d["other_iteration"]
print("GN frac %.3f value %.0f | grad frac %.3f | e2e %.0f (link %.1f GB/s, frac %.2f) target-only %.0f | dropin %.0f it/s | cpu %.2f" % (
    d["roofline"]["frac"], d["value"], o["roofline_frac"], e["value"], e["h2d_link_GBps"], e["frac_of_link"],
    e["target_frame_only"]["value"], d["dropin_single_pair"]["iters_per_s"], d["cpu_baseline"]["value"]))
print("numa", e.get("numa"))
PY
for w in c2levels c3 c4 c5 compaction; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_${w}_$TAG.json 2> $OUT/bench_${w}_$TAG.err
  tail -c 600 $OUT/bench_${w}_$TAG.json; echo; tail -3 $OUT/bench_${w}_$TAG.err
done
# fused source ingest (experiment build): parity tests of the e2e path, then the end-to-end arm
SPB200_LIB=$L/libspb200_fused.so timeout 600 python -m pytest tests/test_gpu_bench_e2e.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -5 > $OUT/pytest_fused_$TAG.log
tail -2 $OUT/pytest_fused_$TAG.log
for lean in 0 1; do
  SPB200_LIB=$L/libspb200_fused.so SPB_E2E_LEAN=$([ $lean = 1 ] && echo 1) timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
      > $OUT/bench_e2e_fused${lean}_$TAG.json 2> $OUT/bench_e2e_fused${lean}_$TAG.err
  python - $OUT/bench_e2e_fused${lean}_$TAG.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d["e2e"]
    print(sys.argv[1], "e2e %.0f link %.1f GB/s frac %.2f target-only %.0f" % (e["value"], e["h2d_link_GBps"], e["frac_of_link"], e["target_frame_only"]["value"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_lm|k_window|k_ingest" -c 60 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 python scripts/profile_dropin.py > $OUT/dropin_profile_$TAG.txt 2>&1; tail -25 $OUT/dropin_profile_$TAG.txt
ls $OUT | wc -l
