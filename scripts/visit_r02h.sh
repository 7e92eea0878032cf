#!/bin/bash
TAG=r02h
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
summ() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); o = d["other_iteration"]
        print("%-40s GN frac=%.3f kernel=%.4f ms | grad frac=%.3f kernel=%.4f ms step=%.4f" % (
            f.split("/")[-1], d["roofline"]["frac"], d["roofline"]["kernel_ms"], o["roofline_frac"], o["kernel_ms"], o["ms_per_step"]))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
}
timeout 300 python bench.py $B > $OUT/bench_base_$TAG.json 2> $OUT/bench_base_$TAG.err
summ $OUT/bench_base_$TAG.json
for v in gr3 gr3h; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > $OUT/bench_c3_${v}_$TAG.json 2> $OUT/bench_c3_${v}_$TAG.err
  python -c "
import json,sys
d=json.loads(open('$OUT/bench_c3_${v}_$TAG.json').read().strip().splitlines()[-1]); print('c3 $v', d['value'], d['roofline']['frac'])"
done
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $OUT/pytest_$TAG.log
tail -4 $OUT/pytest_$TAG.log
