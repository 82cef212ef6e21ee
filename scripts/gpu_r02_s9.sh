#!/bin/bash
# Round-2 session 9: throughput against the number of turns per launch (same beam).
TAG=${1:-r02s9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  local label=$1; shift
  timeout 900 python bench.py --quick --no-cpu-baseline "$@" > $OUT/q_${label}.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/q_${label}.json'))
    print('${label}', '%.4e PET/s'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'ms/step %.1f'%d['ms_per_step'], d.get('beam'))
except Exception as e:
    print('${label} FAILED', e)
PY
}
for t in 5 10 25 50 100 200 400; do
  run turns$t --turns $t --steps 2 --warmup 1
done
run turns100_fma --turns 100 --steps 2 --warmup 1 --fma
run turns25_fma --turns 25 --steps 2 --warmup 1 --fma
run turns100_sps --workload sps_apertures --turns 100 --steps 2 --warmup 1
run turns20_sps --workload sps_apertures --turns 20 --steps 2 --warmup 1
run turns100_c25 --turns 100 --steps 2 --warmup 1 --compact-every 25
