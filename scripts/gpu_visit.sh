#!/bin/bash
# One budget-bounded GPU-box visit, most important evidence first (a cut-off visit still leaves the earlier files):
#   1 parity tests (-m gpu) + smoke      2 bench.py (GN, full line) + first-order mode      3 variants: tests + bench
#   4 ncu launch list                    5 ncu --set full of the fused kernel (GN, gradient; best variant if given)
# Usage (build container): gpurun --timeout 900 -- 'bash scripts/gpu_visit.sh <tag> "<variant> ..."'
TAG=${1:-r01}
VARIANTS=$2
OUT=gpurun_out
mkdir -p $OUT
B="--no-cpu-baseline --no-e2e"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -60 > $OUT/pytest_$TAG.log
tail -8 $OUT/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
timeout 400 python bench.py --steps 30 --warmup 5 2> $OUT/bench_gn_$TAG.err | tee $OUT/bench_gn_$TAG.json | cut -c1-300
tail -5 $OUT/bench_gn_$TAG.err
timeout 200 python bench.py --steps 30 --warmup 5 --mode grad $B 2> $OUT/bench_grad_$TAG.err | tee $OUT/bench_grad_$TAG.json | cut -c1-200
summ() {   # one line per bench file
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%s value=%.0f frac=%.3f kernel_ms=%.4f sm=%s" % (f, d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
}
for v in $VARIANTS; do
  LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so
  SPB200_LIB=$LIB timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $OUT/pytest_${v}_$TAG.log
  tail -2 $OUT/pytest_${v}_$TAG.log
  for m in gn grad; do
    SPB200_LIB=$LIB timeout 200 python bench.py --steps 30 --warmup 5 --mode $m $B 2> $OUT/bench_${m}_${v}_$TAG.err > $OUT/bench_${m}_${v}_$TAG.json
  done
  summ $OUT/bench_gn_${v}_$TAG.json $OUT/bench_grad_${v}_$TAG.json
done
summ $OUT/bench_gn_$TAG.json $OUT/bench_grad_$TAG.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_lm|k_window" -c 40 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 $B > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_$TAG \
    python bench.py --steps 3 --warmup 3 $B > $OUT/ncu_gn_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad $B > $OUT/ncu_grad_$TAG.log 2>&1
for v in $VARIANTS; do
  SPB200_LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so timeout 300 ncu --set full --clock-control none --import-source on \
      -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_${v}_$TAG python bench.py --steps 3 --warmup 3 $B > $OUT/ncu_gn_${v}_$TAG.log 2>&1
done
ls -la $OUT | tail -20
