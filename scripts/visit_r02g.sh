#!/bin/bash
# visit r02g: cached source colours read straight from global memory (8-byte slots), 256-point tiles
TAG=r02g
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
for v in rd256 rd128 t256; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
  tail -2 $OUT/bench_${v}_$TAG.err
done
SPB200_LIB=$L/libspb200_rd256.so timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $OUT/pytest_rd256_$TAG.log
tail -4 $OUT/pytest_rd256_$TAG.log
for w in c2levels c5 c3; do
  SPB200_LIB=$L/libspb200_rd256.so timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_${w}_rd256_$TAG.json 2> $OUT/bench_${w}_rd256_$TAG.err
  python - $OUT/bench_${w}_rd256_$TAG.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"][:40], "value %.1f %s frac %.3f" % (d["value"], d["unit"], d["roofline"]["frac"]))
    for l in d.get("levels", []): print("   ", l["level"], l["target"], l["iteration"], "frac %.3f" % l["frac"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
SPB200_LIB=$L/libspb200_rd256.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_rd256_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_rd256_$TAG.log 2>&1
SPB200_LIB=$L/libspb200_rd256.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_grad_rd256_$TAG \
    python bench.py --steps 3 --warmup 3 --mode grad --no-cpu-baseline --no-e2e > $OUT/ncu_grad_rd256_$TAG.log 2>&1
