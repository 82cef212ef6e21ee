#!/bin/bash
# GPU-box session: thick inlining/occupancy variants, radiation workloads.
TAG=${1:-s6}
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
for sfx in "" _m1hb3 _m1hb4 _m1hb5 _m1hb6 _m0hb4; do
  run "$sfx" lep${sfx}_exact --workload lep_thick --particles 300000 --steps 2 --warmup 1 --turns 3
done
run "_m1hb4" clicq_m1hb4 --workload clic_dr_quantum --particles 300000 --steps 2 --warmup 1 --turns 3
run "" clicq --workload clic_dr_quantum --particles 300000 --steps 2 --warmup 1 --turns 3
run "_m1hb4" lepq_m1hb4 --workload lep_quantum --particles 300000 --steps 2 --warmup 1 --turns 3
} > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt; tail -5 $OUT/err.log
