#!/bin/bash
# Round-2 session 2: GPU suite with the slice tests; particles-per-thread sweep for small beams;
# A/B of the proxy fence and of the hot-loop code placement (same box).
TAG=${1:-r02s2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $OUT/pytest_gpu.log | tail -30
q() {  # label, env..., -- bench args
  local label=$1; shift
  env "$@" timeout 300 python bench.py --quick --no-cpu-baseline --steps 4 --warmup 3 $BARGS > $OUT/q_${label}.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/q_${label}.json'))
    print('${label}', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e:
    print('${label} FAILED', e)
PY
}
for n in 125000 250000 500000 1000000; do
  for npt in 1 2 3; do
    BARGS="--particles $n" q n${n}_npt${npt} XTB_NPT_FORCE=$npt
  done
done
for sfx in "" _nofence _pad1 _pad2 _pad4; do
  BARGS="--particles 1000000" q lib${sfx}_exact XTB_LIB_SUFFIX=$sfx
  BARGS="--particles 1000000 --fma" q lib${sfx}_fma XTB_LIB_SUFFIX=$sfx
done
BARGS="--particles 1000000" q again_exact XTB_LIB_SUFFIX=
