#!/bin/bash
# First GPU visit of the next round: the experiment builds prepared (and only compile-checked) in round 1.
#   build container:  scripts/build_variant.sh cc -DSPB_CTX_CONST=1 ; scripts/build_variant.sh fused -DSPB_INGEST_FUSED=1
#   gpurun --timeout 600 -- 'bash scripts/round2_first_visit.sh r02a'
# cc    : gradient mode of the batched solver reads the per-pair context from a __constant__ array (LSU data pipe 75 % busy,
#         a quarter of it context re-reads) -- parity tests with the variant library, then both iteration modes
# fused : source half of the frame ingest in one kernel straight from the 8-bit frame -- parity tests (incl. the lean
#         e2e case that is skipped on the default library), then the end-to-end arm with SPB_E2E_LEAN=1
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e"
L=$PWD/super_primitive_b200/csrc
[ -f $L/libspb200_cc.so ] || bash scripts/build_variant.sh cc -DSPB_CTX_CONST=1 | tail -1
[ -f $L/libspb200_fused.so ] || bash scripts/build_variant.sh fused -DSPB_INGEST_FUSED=1 | tail -1
timeout 200 python bench.py --steps 30 --warmup 5 $B > $OUT/bench_gn_base_$TAG.json 2> $OUT/bench_gn_base_$TAG.err
for v in cc fused; do
  SPB200_LIB=$L/libspb200_$v.so timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $OUT/pytest_${v}_$TAG.log
  tail -3 $OUT/pytest_${v}_$TAG.log
done
SPB200_LIB=$L/libspb200_cc.so timeout 200 python bench.py --steps 30 --warmup 5 $B > $OUT/bench_gn_cc_$TAG.json 2> $OUT/bench_gn_cc_$TAG.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_e2e_base_$TAG.json 2> $OUT/bench_e2e_base_$TAG.err
SPB200_LIB=$L/libspb200_fused.so SPB_E2E_LEAN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
    > $OUT/bench_e2e_fused_$TAG.json 2> $OUT/bench_e2e_fused_$TAG.err
python - $TAG <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/bench_*_%s.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); o = d["other_iteration"]; e = d.get("e2e") or {}
        print(f, "GN frac=%.3f  first-order frac=%.3f kernel_ms=%.4f  e2e=%s" % (
            d["roofline"]["frac"], o["roofline_frac"], o["kernel_ms"], e.get("value")))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
SPB200_LIB=$L/libspb200_cc.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_global -s 3 -c 1 -f \
    -o $OUT/prof_grad_cc_$TAG python bench.py --steps 3 --warmup 3 --mode grad $B > $OUT/ncu_grad_cc_$TAG.log 2>&1
