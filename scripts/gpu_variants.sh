#!/bin/bash
# Tuning visit: build-time variants of the library (libspb200_<v>.so, scripts/build_variant.sh) against the default one.
# Usage: gpurun --timeout 600 -- 'bash scripts/gpu_variants.sh <tag> "<tested variants>" "<bench-only variants>" "<ncu variants>" [basetests]'
#   tested variants : parity tests (-m gpu) + bench in both iteration modes
#   bench-only      : bench only (diagnostic builds whose results are wrong on purpose)
#   ncu variants    : one ncu --set full capture of the fused GN kernel each
TAG=${1:-var}
OUT=gpurun_out
mkdir -p $OUT
B="--no-cpu-baseline --no-e2e"
summ() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        o = d["other_iteration"]
        print("%s value=%.0f frac=%.3f kernel_ms=%.4f | other frac=%.3f kernel_ms=%.4f | sm=%s" % (
            f, d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], o["roofline_frac"], o["kernel_ms"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
}
if [ "$5" = "basetests" ]; then
  timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $OUT/pytest_$TAG.log; tail -3 $OUT/pytest_$TAG.log
fi
# the GN line carries the first-order iteration as `other_iteration`, so one run per library covers both kernels
timeout 200 python bench.py --steps 30 --warmup 5 $B 2> $OUT/bench_gn_base_$TAG.err > $OUT/bench_gn_base_$TAG.json
summ $OUT/bench_gn_base_$TAG.json
for v in $2; do
  LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so
  SPB200_LIB=$LIB timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $OUT/pytest_${v}_$TAG.log
  tail -2 $OUT/pytest_${v}_$TAG.log
done
for v in $2 $3; do
  LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so
  SPB200_LIB=$LIB timeout 200 python bench.py --steps 30 --warmup 5 $B 2> $OUT/bench_gn_${v}_$TAG.err > $OUT/bench_gn_${v}_$TAG.json
  summ $OUT/bench_gn_${v}_$TAG.json
done
for v in $4; do
  SPB200_LIB=$PWD/super_primitive_b200/csrc/libspb200_$v.so timeout 300 ncu --set full --clock-control none --import-source on \
      -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_${v}_$TAG python bench.py --steps 3 --warmup 3 $B > $OUT/ncu_gn_${v}_$TAG.log 2>&1
done
ls $OUT | tail -5
