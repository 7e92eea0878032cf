#!/bin/bash
# visit r02v: CTA-level producer/consumer pipeline with the round's projected target footprint staged in shared memory (SPB_BOXCTA)
# (experiment: the variant libraries need profiles/r02x_boxcta.patch applied -- `git apply profiles/r02x_boxcta.patch` -- and scripts/build_variant.sh; the default tree does not contain the switch)
TAG=r02v
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
for v in bcta; do
  export SPB200_LIB=$L/libspb200_$v.so
  timeout 300 python -m pytest tests/test_gpu_gn.py tests/test_gpu_parity.py tests/test_gpu_headline.py -q -x 2>&1 | tail -3
  timeout 300 python bench.py $B > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
  summ $OUT/bench_${v}_$TAG.json
  tail -2 $OUT/bench_${v}_$TAG.err
done
unset SPB200_LIB
timeout 300 python -m pytest tests/test_gpu_gn.py tests/test_gpu_parity.py -q -x 2>&1 | tail -2
timeout 300 python bench.py $B > $OUT/bench_base_$TAG.json 2> $OUT/bench_base_$TAG.err
summ $OUT/bench_base_$TAG.json
SPB200_LIB=$L/libspb200_bcta.so timeout 300 ncu --set full --clock-control none -k regex:k_align_global -s 3 -c 1 -f -o $OUT/prof_gn_bcta_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_gn_bcta_$TAG.log 2>&1
