#!/bin/bash
# visit r02p: CTA-contiguous tile chunks with per-run (lazy) segment flush vs the strided default
# (experiment: the variant libraries need profiles/r02p_cta_runs.patch applied -- `git apply profiles/r02p_cta_runs.patch` -- and scripts/build_variant.sh; the default tree does not contain the switch)
TAG=r02p
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --steps 30 --warmup 5"
L=$PWD/super_primitive_b200/csrc
summ() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); o = d["other_iteration"]
        print("%-40s GN frac=%.3f kernel=%.4f ms step=%.4f value=%.0f | grad frac=%.3f kernel=%.4f ms step=%.4f" % (
            f.split("/")[-1], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["value"], o["roofline_frac"], o["kernel_ms"], o["ms_per_step"]))
        b = d.get("blob_segments") or {}
        if "gn" in b: print("    blobs: GN frac=%.3f grad frac=%.3f   e2e %.0f" % (b["gn"]["roofline_frac"], b["first_order"]["roofline_frac"], d["e2e"]["value"]))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
}
timeout 300 python bench.py $B > $OUT/bench_base_$TAG.json 2> $OUT/bench_base_$TAG.err
summ $OUT/bench_base_$TAG.json
for v in runs; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
  tail -2 $OUT/bench_${v}_$TAG.err
done


for w in c2levels c5 c3; do
  for v in base runs; do
    if [ $v = base ]; then unset SPB200_LIB; else export SPB200_LIB=$L/libspb200_$v.so; fi
    timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_${w}_${v}_$TAG.json 2> $OUT/bench_${w}_${v}_$TAG.err
    python - $OUT/bench_${w}_${v}_$TAG.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["config"]["workload"][:40], "value %.1f %s frac %.3f" % (d["value"], d["unit"], d["roofline"]["frac"]))
    for l in d.get("levels", []): print("   ", l["level"], l["target"], l["iteration"], "frac %.3f" % l["frac"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
  done
done
