#!/bin/bash
# session 26: code placement of the hot function with the aperture pre-filter in (pads 0-4)
TAG=${1:-r02s26}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in "" _noap _pad1 _pad2 _pad3 _pad5; do
  XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline > $OUT/q${v}_$rep.json 2>> $OUT/bench.err
  python - <<PY
import json
for w in ('q',):
    try:
        d=json.load(open('$OUT/%s${v}_$rep.json'%w)); print(w+'$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
    except Exception as e: print(w+'$v FAILED', e)
PY
done
done
for v in _pad1 _pad2 _pad3 _pad5; do
  XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline --workload sps_apertures --particles 2000000 > $OUT/sps${v}.json 2>> $OUT/bench.err
  python -c "import json; d=json.load(open('$OUT/sps${v}.json')); print('sps$v', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])"
done
