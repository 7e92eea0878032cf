#!/bin/bash
# visit r02f: constant-cache context experiment (tests + bench), c4 after the fixed-point render
TAG=r02f
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
summ() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); o = d["other_iteration"]
        print("%-40s GN frac=%.3f kernel=%.4f ms step=%.4f value=%.0f | grad frac=%.3f kernel=%.4f ms step=%.4f" % (
            f.split("/")[-1], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["value"], o["roofline_frac"], o["kernel_ms"], o["ms_per_step"]))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
}
timeout 300 python bench.py $B > $OUT/bench_base_$TAG.json 2> $OUT/bench_base_$TAG.err
summ $OUT/bench_base_$TAG.json
for v in cc ccw6 ccg3; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
done
SPB200_LIB=$L/libspb200_cc.so timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $OUT/pytest_cc_$TAG.log
tail -3 $OUT/pytest_cc_$TAG.log
timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err
tail -c 400 $OUT/bench_c4_$TAG.json; echo
SPB200_LIB=$L/libspb200_cc.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_cc_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_cc_$TAG.log 2>&1
SPB200_LIB=$L/libspb200_cc.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_cc_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_cc_$TAG.log 2>&1
