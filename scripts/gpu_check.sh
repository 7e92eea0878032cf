#!/bin/bash
# One GPU-box visit: parity tests, smoke, benchmark (both iteration kinds), ncu launch list + full capture.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag] [noprof] [variants]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -80 > $OUT/pytest_$TAG.log
tail -25 $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 2> $OUT/bench_gn_$TAG.err | tee $OUT/bench_gn_$TAG.json
timeout 300 python bench.py --steps 30 --warmup 5 --mode grad --no-cpu-baseline --no-e2e 2> $OUT/bench_grad_$TAG.err | tee $OUT/bench_grad_$TAG.json
if [ "$3" != "" ]; then
  for v in $3; do
    for m in gn grad; do
      SPB200_LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --mode $m \
         --no-cpu-baseline --no-e2e 2> $OUT/bench_${m}_${v}_$TAG.err | tee $OUT/bench_${m}_${v}_$TAG.json | cut -c1-120,500-900
    done
  done
fi
if [ "$2" != "noprof" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_lm" -c 40 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_$TAG.log 2>&1
ls -la $OUT | tail -12
fi
