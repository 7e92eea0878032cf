#!/bin/bash
# visit r02r: CTA-contiguous runs with bounded chunks (more CTAs per pair -> fewer pairs in flight -> smaller L2 working set)
# (experiment: the variant libraries need profiles/r02p_cta_runs.patch applied -- `git apply profiles/r02p_cta_runs.patch` -- and scripts/build_variant.sh; the default tree does not contain the switch)
TAG=r02r
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
for v in base runs40 runs64 runs96; do
  if [ $v = base ]; then unset SPB200_LIB; else export SPB200_LIB=$L/libspb200_$v.so; fi
  timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
  for w in c2levels c5; do
    timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_${w}_${v}_$TAG.json 2> $OUT/bench_${w}_${v}_$TAG.err
    python - $OUT/bench_${w}_${v}_$TAG.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ", sys.argv[2], d["config"]["workload"][:30], "value %.1f frac %.3f" % (d["value"], d["roofline"]["frac"]),
          " ".join("%s:%.3f" % (l["iteration"], l["frac"]) for l in d.get("levels", [])))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
  done
done
