#!/bin/bash
# GPU-box session: A/B of the main-coefficient trimming, turn-by-turn monitor leg.
TAG=${1:-t9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # sfx, label, bench args...
  local sfx=$1; shift; local label=$1; shift
  XTB_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --quick "$@" > $OUT/bench_${label}.json 2>> $OUT/err.log
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${label}.json"))
    print("${label}: PET/s %.3e frac %.3f kernel_ms %.1f" % (d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"]), d.get("monitor", ""))
except Exception as e:
    print("${label} FAILED", e)
PY
}
{
for rep in 1 2; do
for sfx in "" $VARIANTS; do
  run "$sfx" lep${sfx}_exact_$rep --workload lep_thick --particles 300000 --steps 2 --warmup 1 --turns 3
done
done
run "" thin_monitor --monitor --particles 1000000 --steps 3 --warmup 1 --turns 10
run "" lep_monitor --workload lep_thick --monitor --particles 300000 --steps 2 --warmup 1 --turns 3
run "" thin_exact --steps 3 --warmup 1 --turns 10
} > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt; tail -5 $OUT/err.log
