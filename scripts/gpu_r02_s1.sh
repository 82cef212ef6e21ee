#!/bin/bash
# Round-2 session 1: the whole GPU suite (new both-tier row tests included), bench with the
# new launch planner vs the legacy single grid (A/B), small shards (strong scaling), baseline ncu.
TAG=${1:-r02s1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/host.txt; nvidia-smi -L >> $OUT/host.txt
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $OUT/pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
for n in 1000000 909312 500000 250000 125000; do
  for shape in planned legacy; do
    XTB_LAUNCH_SHAPE=$shape timeout 300 python bench.py --quick --steps 4 --warmup 3 --particles $n \
        > $OUT/q_${n}_${shape}.json 2>> $OUT/bench.err
    python - <<PY
import json
d=json.load(open('$OUT/q_${n}_${shape}.json'))
print('$n $shape', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'launches', d['gpu_launches'])
PY
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --turns 2 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/smoke.log; cat $OUT/bench.json
