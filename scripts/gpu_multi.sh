#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the driver's launch line for N ranks, both arms.
N=${1:-2}; TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --cpu-seconds 4 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "exit $?" >> $OUT/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 --cpu-seconds 6 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "exit $?" >> $OUT/bench_ref_n$N.err
cat $OUT/bench_n$N.json | cut -c1-600; tail -3 $OUT/bench_n$N.err; cat $OUT/bench_ref_n$N.json | cut -c1-300; tail -2 $OUT/bench_ref_n$N.err
