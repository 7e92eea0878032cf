#!/bin/bash
# multi-GPU visit (gpurun --gpus N): the driver's launch line for the headline bench + the sharded workloads with their
# bit-equality shard check
N=${1:-2}; TAG=${2:-r02e}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_n${N}_$TAG.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 900 bash -c "$(declare -f run); N=$N; run 29611 --steps 20 --warmup 5" > $OUT/bench_gn_n${N}_$TAG.json 2> $OUT/bench_gn_n${N}_$TAG.err
python - $OUT/bench_gn_n${N}_$TAG.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1]); e = d["e2e"]
    print("N=%d value %.0f e2e %.0f link(min over ranks, concurrent) %.1f GB/s frac %.2f target-only %.0f numa %s gather %.2f ms" % (
        d["n_gpus"], d["value"], e["value"], e["h2d_link_GBps"], e["frac_of_link"], e["target_frame_only"]["value"], e["numa"], d.get("final_gather_ms", 0)))
except Exception as ex:
    print("FAILED", ex)
PY
port=29621
for w in c3 c5 c4; do
  port=$((port+1))
  timeout 900 bash -c "$(declare -f run); N=$N; run $port --workload $w --steps 10 --warmup 3" > $OUT/bench_${w}_n${N}_$TAG.json 2> $OUT/bench_${w}_n${N}_$TAG.err
  python - $OUT/bench_${w}_n${N}_$TAG.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(d["config"]["workload"][:40], "N=%d value %.1f %s frac %.3f shard_check %s" % (d["n_gpus"], d["value"], d["unit"], d["roofline"]["frac"], d.get("shard_check")))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
  tail -2 $OUT/bench_${w}_n${N}_$TAG.err
done
