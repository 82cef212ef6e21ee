#!/bin/bash
# session 20: A/B of the radiating thin kick (Horner sums reused, compile-time orders, 1/length
# folded by the host, one decode for all lanes) against the previous commit; hllhc check
TAG=${1:-r02s20}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_radiation.py tests/test_quantum_kick.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for rep in 1 2; do
for v in "" _old; do
  for wl in clic_dr_mean clic_dr_qkick lep_mean lep_thick; do
    XTB_LIB_ABI_OVERRIDE=5 XTB_LIB_SUFFIX=$v timeout 400 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 2 --particles 300000 --no-cpu-baseline > $OUT/bench_${wl}${v}_$rep.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${wl}${v}_$rep.json')); print('$wl$v', '$rep', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e: print('$wl$v FAILED', e)
PY
  done
done
done
XTB_LIB_SUFFIX= timeout 300 python bench.py --quick --steps 3 --warmup 3 --turns 100 --no-cpu-baseline > $OUT/q.json 2>> $OUT/bench.err
python -c "import json; d=json.load(open('$OUT/q.json')); print('hllhc', '%.4e'%d['value'], d['roofline']['frac'])"
tail -3 $OUT/bench.err
