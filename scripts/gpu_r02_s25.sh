#!/bin/bash
# session 25: A/B of the integer pre-filter of the fast aperture ops (SPS, 2e6 particles); hllhc check
TAG=${1:-r02s25}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_restated.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for rep in 1 2; do
for v in "" _noap; do
  XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline --workload sps_apertures --particles 2000000 > $OUT/sps${v}_$rep.json 2>> $OUT/bench.err
  XTB_LIB_SUFFIX=$v timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline > $OUT/q${v}_$rep.json 2>> $OUT/bench.err
  python - <<PY
import json
for w in ('sps','q'):
    try:
        d=json.load(open('$OUT/%s${v}_$rep.json'%w)); print(w+'$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'], d.get('beam'))
    except Exception as e: print(w+'$v FAILED', e)
PY
done
done
tail -3 $OUT/bench.err
