#!/bin/bash
# Round-2 session 8: long launches of the dynamic-aperture workload (losses, compaction), small
# beams with the block-size rule, full GPU suite with the bit-exact LEP ring.
TAG=${1:-r02s8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|bit-identical" $OUT/pytest_gpu.log | tail -30
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
run da_200x5 --turns 200 --steps 5 --warmup 0
run da_1000_nocompact --turns 1000 --steps 1 --warmup 0
run da_1000_compact100 --turns 1000 --steps 1 --warmup 0 --compact-every 100
run da_1000_compact25 --turns 1000 --steps 1 --warmup 0 --compact-every 25
run toy_thick_1e4 --workload toy_ring --particles 10000 --turns 100 --steps 2 --warmup 1
run toy_thin_1e4 --workload toy_ring_thin --particles 10000 --turns 1000 --steps 3 --warmup 3
