#!/bin/bash
# session 27: the aperture pre-filter as a separate instantiation of the hot function, against
# the build without it (hllhc must not move, SPS keeps the gain); GPU tests
TAG=${1:-r02s27}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in "" _old; do
  XTB_LIB_ABI_OVERRIDE=6 XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline > $OUT/q${v}_$rep.json 2>> $OUT/bench.err
  XTB_LIB_ABI_OVERRIDE=6 XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline --workload sps_apertures --particles 2000000 > $OUT/sps${v}_$rep.json 2>> $OUT/bench.err
  python - <<PY
import json
for w in ('q','sps'):
    try:
        d=json.load(open('$OUT/%s${v}_$rep.json'%w)); print(w+'$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
    except Exception as e: print(w+'$v FAILED', e)
PY
done
done
tail -2 $OUT/bench.err
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
