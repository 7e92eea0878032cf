#!/bin/bash
# Multi-GPU visit: the driver's launch line for N ranks, both arms.
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 2> $OUT/bench_n${N}_$TAG.err | tee $OUT/bench_n${N}_$TAG.json | cut -c1-400
tail -3 $OUT/bench_n${N}_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 3 2> $OUT/bench_ref_n${N}_$TAG.err | tee $OUT/bench_ref_n${N}_$TAG.json | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> $OUT/bench_n1_$TAG.err | tee $OUT/bench_n1_$TAG.json | cut -c1-200
