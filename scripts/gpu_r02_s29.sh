#!/bin/bash
# session 29: Chebyshev series of the photon spectrum unrolled with constant-bank operands, A/B
TAG=${1:-r02s29}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "" _loop; do
  for wl in clic_dr_quantum lep_quantum; do
    XTB_LIB_SUFFIX=$v timeout 200 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 1 --particles 300000 --no-cpu-baseline > $OUT/bench_${wl}${v}.json 2>> $OUT/bench.err
    python -c "import json; d=json.load(open('$OUT/bench_${wl}${v}.json')); print('$wl$v', '%.4e'%d['value'])"
  done
done
timeout 100 python -m pytest tests/test_gpu_radiation.py -m gpu -q -x -k "quantum" > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
