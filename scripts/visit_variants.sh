#!/bin/bash
# Tuning visit: default library through the parity tests, then bench.py (GN + first-order kernel fractions) for the
# default and every variant library given (built beforehand with scripts/build_variant.sh; they travel with the snapshot).
#   gpurun --timeout 1500 -- 'bash scripts/visit_variants.sh r02a "g1t0 g4t0 ..."'
TAG=${1:-r02a}
VARIANTS=$2
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
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
for v in $VARIANTS; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_$TAG.log 2>&1
ls -la $OUT | tail -8
