#!/bin/bash
# Round-2 session 4: particles per thread x block size for small beams (one box).
TAG=${1:-r02s4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
q() {
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
for wl in hllhc_da sps_apertures; do
for n in 62500 125000 250000 375000 500000; do
  for npt in 1 2 3; do
    for t in 128 64 32; do
      BARGS="--particles $n --workload $wl" q ${wl}_n${n}_npt${npt}_t${t} XTB_NPT_FORCE=$npt XTB_THREADS_FORCE=$t
    done
  done
done
done
