#!/bin/bash
# GPU-box session: run-unrolling on/off (thin workloads), heavy default check.
TAG=${1:-s7}
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
for sfx in "" _ur0; do
  run "$sfx" thin${sfx}_exact --steps 3 --warmup 1 --turns 10
  run "$sfx" thin${sfx}_fma --steps 3 --warmup 1 --turns 10 --fma
  run "$sfx" sps${sfx}_exact --workload sps_apertures --particles 1000000 --steps 3 --warmup 1 --turns 10
done
run "" lep_exact --workload lep_thick --particles 300000 --steps 2 --warmup 1 --turns 3
} > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt; tail -5 $OUT/err.log
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|exit" $OUT/pytest_gpu.log | tail -3
