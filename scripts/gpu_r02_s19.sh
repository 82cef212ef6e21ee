#!/bin/bash
# session 19: A/B of the integer pre-filter in the drift end of the heavy run loop
TAG=${1:-r02s19}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_radiation.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for rep in 1 2; do
for v in "" _nopf; do
  for wl in lep_thick clic_dr_mean lep_mean clic_dr_qkick; do
    XTB_LIB_SUFFIX=$v timeout 400 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 2 --particles 300000 --no-cpu-baseline > $OUT/bench_${wl}${v}_$rep.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${wl}${v}_$rep.json')); print('$wl$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e: print('$wl$v FAILED', e)
PY
  done
done
done
