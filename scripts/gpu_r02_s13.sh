#!/bin/bash
# session 13: GPU test suite, headline bench, long launches with losses inside
TAG=${1:-r02s13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; cat $OUT/bench_default.json
timeout 900 python scripts/diag_lost2.py > $OUT/diag_lost.log 2>&1
cat $OUT/diag_lost.log
