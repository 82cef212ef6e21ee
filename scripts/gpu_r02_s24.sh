#!/bin/bash
# session 24: why are small blocks slow for a small shard? 125 000 particles, 3 particles per
# thread in blocks of 32 / 64 threads against the chosen shape (1 particle per thread, 128)
TAG=${1:-r02s24}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "0 0" "3 32" "3 64" "2 64" "2 32"; do
  set -- $cfg
  XTB_NPT_FORCE=$1 XTB_THREADS_FORCE=$2 timeout 300 python bench.py --quick --particles 125000 --steps 3 --warmup 3 --turns 100 --no-cpu-baseline > $OUT/q_$1_$2.json 2>> $OUT/bench.err
  python -c "import json; d=json.load(open('$OUT/q_$1_$2.json')); print('npt $1 threads $2', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])"
done
for cfg in "0 0" "3 32"; do
  set -- $cfg
  XTB_NPT_FORCE=$1 XTB_THREADS_FORCE=$2 timeout 600 ncu --set full --clock-control none -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_$1_$2 -f python bench.py --quick --particles 125000 --steps 1 --warmup 1 --turns 3 --no-cpu-baseline > $OUT/ncu_$1_$2.log 2>&1
done
du -sh $OUT
