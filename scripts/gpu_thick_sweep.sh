#!/bin/bash
# GPU-box session: thick path after the multi-lane body (NPT_HEAVY, blocks/SM sweep), GPU
# parity tests, ncu capture of the thick kernel.
TAG=${1:-t2}
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
  run "$sfx" lep${sfx}_exact --workload lep_thick --particles 300000 --steps 2 --warmup 1 --turns 3
done
run "" lep_fma --workload lep_thick --particles 300000 --steps 2 --warmup 1 --turns 3 --fma
run "" thin_exact --steps 3 --warmup 1 --turns 10
run "" sps_exact --workload sps_apertures --particles 1000000 --steps 3 --warmup 1 --turns 10
} > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt; tail -5 $OUT/err.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|exit" $OUT/pytest_gpu.log | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_lep -f python bench.py --workload lep_thick --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_lep.log 2>&1
ls -la $OUT
