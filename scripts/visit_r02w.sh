#!/bin/bash
# visit r02w: SPB_BOXCTA with the producer's box computation moved ahead of its stage wait; ncu with source
# (experiment: the variant libraries need profiles/r02x_boxcta.patch applied -- `git apply profiles/r02x_boxcta.patch` -- and scripts/build_variant.sh; the default tree does not contain the switch)
TAG=r02w
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
export SPB200_LIB=$L/libspb200_bcta.so
timeout 300 python -m pytest tests/test_gpu_gn.py tests/test_gpu_parity.py -q -x 2>&1 | tail -2
timeout 300 python bench.py $B > $OUT/bench_bcta_$TAG.json 2> $OUT/bench_bcta_$TAG.err
python - $OUT/bench_bcta_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); o = d["other_iteration"]
print("GN frac=%.3f kernel=%.4f ms | grad frac=%.3f kernel=%.4f ms" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], o["roofline_frac"], o["kernel_ms"]))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_bcta_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_bcta_$TAG.log 2>&1
