#!/bin/bash
# GPU-box session: radiation kernels, blocks / SM sweep.
TAG=${1:-t11}
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
for sfx in "" $VARIANTS; do
  run "$sfx" lepq${sfx} --workload lep_quantum --particles 200000 --steps 2 --warmup 1 --turns 2
  run "$sfx" clicq${sfx} --workload clic_dr_quantum --particles 300000 --steps 2 --warmup 1 --turns 2
done
} > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt; tail -5 $OUT/err.log
