#!/bin/bash
# GPU-box session: sweep of XTB_HOT_PAD (placement of the hot handlers relative to the
# instruction-cache lines, csrc/xtb_interp.cuh::xtb_run_fast).  Build the variants first:
#   scripts/build_variants.sh p1 "-DXTB_HOT_PAD=1" ... p7 "-DXTB_HOT_PAD=7"
TAG=${1:-pad}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # sfx, label, bench args...
  local sfx=$1; shift; local label=$1; shift
  XTB_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --quick "$@" > $OUT/bench_${label}.json 2>> $OUT/err.log
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${label}.json"))
    print("${label}: PET/s %.3e frac %.3f kernel_ms %.1f" % (d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"]))
except Exception as e:
    print("${label} FAILED", e)
PY
}
{
for sfx in "" _p1 _p2 _p3 _p4 _p5 _p6 _p7; do
  run "$sfx" thin${sfx}_exact --steps 3 --warmup 1 --turns 10
  run "$sfx" thin${sfx}_fma --steps 3 --warmup 1 --turns 10 --fma
done
} > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt; tail -5 $OUT/err.log
