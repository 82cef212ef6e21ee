#!/bin/bash
# session 23: three lanes per thread in the heavy run loops (lanes pinned in thread-local memory)
TAG=${1:-r02s23}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in "" _h3l; do
  for wl in lep_thick lep_mean clic_dr_mean; do
    XTB_LIB_SUFFIX=$v timeout 400 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 2 --particles 600000 --no-cpu-baseline > $OUT/bench_${wl}${v}_$rep.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${wl}${v}_$rep.json')); print('$wl$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e: print('$wl$v FAILED', e)
PY
  done
done
done
XTB_LIB_SUFFIX=_h3l timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "thick or lep or losses" > $OUT/pytest_h3l.log 2>&1; tail -3 $OUT/pytest_h3l.log
tail -3 $OUT/bench.err
