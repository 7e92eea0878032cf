#!/bin/bash
# visit r02q: ncu of the GN kernel, strided default vs CTA-contiguous runs, on the strips headline
# (experiment: the variant libraries need profiles/r02p_cta_runs.patch applied -- `git apply profiles/r02p_cta_runs.patch` -- and scripts/build_variant.sh; the default tree does not contain the switch)
TAG=r02q
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/super_primitive_b200/csrc
timeout 300 ncu --set full --clock-control none -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_base_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_base_$TAG.log 2>&1
SPB200_LIB=$L/libspb200_runs.so timeout 300 ncu --set full --clock-control none -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_runs_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_runs_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep
