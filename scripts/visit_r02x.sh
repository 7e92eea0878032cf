#!/bin/bash
# visit r02x: SPB_BOXCTA decomposition: A = pipeline without boxes, B = boxes staged but unused, D/E = 4 consumer warps per CTA
# (experiment: the variant libraries need profiles/r02x_boxcta.patch applied -- `git apply profiles/r02x_boxcta.patch` -- and scripts/build_variant.sh; the default tree does not contain the switch)
TAG=r02x
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
for v in bctaF bctaG bctaH; do
export SPB200_LIB=$L/libspb200_$v.so
timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
python - $OUT/bench_${v}_$TAG.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); o = d["other_iteration"]
    print(sys.argv[2], "GN frac=%.3f kernel=%.4f ms | grad frac=%.3f kernel=%.4f ms" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], o["roofline_frac"], o["kernel_ms"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
SPB200_LIB=$L/libspb200_bctaD.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_bctaD_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_bctaD_$TAG.log 2>&1
