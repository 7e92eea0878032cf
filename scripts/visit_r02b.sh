#!/bin/bash
# visit r02b: new headline parity tests, texture-gather microbenchmark, 64-register GN variant, compute-sanitizer
TAG=r02b
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
timeout 900 python -m pytest tests/test_gpu_headline.py -q -m gpu -s 2>&1 | tail -60 > $OUT/pytest_headline_$TAG.log
tail -25 $OUT/pytest_headline_$TAG.log
timeout 120 scripts/micro/tex_gather_bench 2>&1 | tee $OUT/tex_gather_$TAG.txt
summ() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); o = d["other_iteration"]
        print("%-40s GN frac=%.3f kernel=%.4f ms value=%.0f | grad frac=%.3f kernel=%.4f ms" % (
            f.split("/")[-1], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["value"], o["roofline_frac"], o["kernel_ms"]))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
}
timeout 300 python bench.py $B > $OUT/bench_base_$TAG.json 2> $OUT/bench_base_$TAG.err
summ $OUT/bench_base_$TAG.json
for v in o4 w6o4; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py > $OUT/sanitizer_${tool}_$TAG.log 2>&1
  tail -4 $OUT/sanitizer_${tool}_$TAG.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_$TAG.log 2>&1
