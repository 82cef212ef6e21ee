#!/bin/bash
# session 15: lanes of the heavy run loops in registers (copy-in / copy-out around out-of-line maps);
# occupancy variants; by-function profile of the LEP thick kernel
TAG=${1:-r02s15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_radiation.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for v in "" _h3 _h5; do
  for wl in lep_thick clic_dr_mean lep_mean clic_dr_quantum; do
    XTB_LIB_SUFFIX=$v timeout 400 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 2 --particles 300000 > $OUT/bench_${wl}$v.json 2>> $OUT/bench.err
    python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${wl}$v.json')); print('$wl$v', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e: print('$wl$v FAILED', e)
PY
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_lep -f python bench.py --workload lep_thick --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_lep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_clic_mean -f python bench.py --workload clic_dr_mean --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_clic_mean.log 2>&1
