#!/bin/bash
# visit r02k: ncu captures of the fused kernel on the other workloads (no source change after r02j)
TAG=r02k
OUT=gpurun_out; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 4 -c 1 -f -o $OUT/prof_c3_$TAG \
    python bench.py --workload c3 --steps 3 --warmup 3 > $OUT/ncu_c3_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 4 -c 1 -f -o $OUT/prof_c5_$TAG \
    python bench.py --workload c5 --units 256 --steps 3 --warmup 3 > $OUT/ncu_c5_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_align|finalize|k_row|k_tile|k_segment|k_depth|k_keypoints" -c 80 --csv \
    --log-file $OUT/launches_c4_$TAG.csv python bench.py --workload c4 --units 8 --steps 1 --warmup 3 > $OUT/ncu_c4_$TAG.log 2>&1
ls -la $OUT | grep $TAG
