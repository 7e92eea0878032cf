#!/bin/bash
# Tuning visit: benchmark build-time variants of the library (libspb200_<v>.so) in both iteration modes.
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_variants.sh <tag> "<v1> <v2> ..."'   ("base" = the default library)
TAG=${1:-var}
OUT=gpurun_out
mkdir -p $OUT
for v in $2; do
  for m in gn grad; do
    if [ "$v" = "base" ]; then LIB=$PWD/super_primitive_b200/csrc/libspb200.so; else LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so; fi
    SPB200_LIB=$LIB timeout 200 python bench.py --steps 30 --warmup 5 --mode $m --no-cpu-baseline --no-e2e \
        2> $OUT/bench_${m}_${v}_$TAG.err > $OUT/bench_${m}_${v}_$TAG.json
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${m}_${v}_$TAG.json").read().strip().splitlines()[-1])
    print("$v $m value=%.0f frac=%.3f kernel_ms=%s clocks=%s" % (d["value"], d["roofline"]["frac"], d["roofline"].get("kernel_ms"), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$v $m FAILED", e)
PY
  done
done
if [ "$3" = "fin" ]; then
  # where does the second kernel of a GN iteration spend its time?
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_lm" -c 24 --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_finalize_solve -s 3 -c 1 -f -o $OUT/prof_fin_$TAG \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_fin_$TAG.log 2>&1
  grep -E "finalize" $OUT/launches_$TAG.csv | tail -3
fi
