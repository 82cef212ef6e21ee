#!/bin/bash
# session 18: A/B of the specialised sextupole / octupole handler, code-placement pads
TAG=${1:-r02s18}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in "" _nopn _pad1 _pad2 _pad4; do
  XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline > $OUT/q${v}_$rep.json 2>> $OUT/bench.err
  XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline --workload sps_apertures --particles 2000000 > $OUT/sps${v}_$rep.json 2>> $OUT/bench.err
  python - <<PY
import json
for w in ('q','sps'):
    try:
        d=json.load(open('$OUT/%s${v}_$rep.json'%w)); print(w+'$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
    except Exception as e: print(w+'$v FAILED', e)
PY
done
done
