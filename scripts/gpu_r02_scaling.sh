#!/bin/bash
# Round-2 scaling session (gpurun --gpus 8): configs[1] as quoted -- 10^6 particles IN TOTAL
# (strong scaling) at N = 1, 2, 4, 8, and 10^6 per GPU (weak) at N = 8; the driver's launch line.
TAG=${1:-r02scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
run() {  # label, n, args...
  local label=$1; local n=$2; shift 2
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $OUT/$label.json 2> $OUT/$label.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n \
        bench.py --gpus $n "$@" > $OUT/$label.json 2> $OUT/$label.err
  fi
  echo "exit $?" >> $OUT/$label.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/$label.json'))
    print('$label', 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'frac(rank0) %.4f'%d['roofline']['frac'], d['scaling'], d['config']['particles_total'])
except Exception as e:
    print('$label FAILED', e)
PY
}
NMAX=$(nvidia-smi -L | wc -l)
if [ $NMAX -ge 8 ]; then
  # (N = 1, 2 were measured on smaller allocations: profiles/r02_bench_exact.json, *_n2.json)
  for n in 4 8; do
    run strong_n$n $n --scaling strong --particles 1000000 --steps 3 --warmup 3 --no-cpu-baseline
  done
  run weak_n8 8 --steps 3 --warmup 3 --no-cpu-baseline
else
  run strong_n$NMAX $NMAX --scaling strong --particles 1000000 --steps 3 --warmup 3 --no-cpu-baseline
  run weak_n$NMAX $NMAX --steps 3 --warmup 3 --cpu-seconds 6
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29599 \
    bench.py --impl reference --gpus $NMAX --steps 1 --warmup 1 --cpu-seconds 4 > $OUT/ref_n$NMAX.json 2> $OUT/ref_n$NMAX.err
cut -c1-400 $OUT/ref_n$NMAX.json
