#!/bin/bash
# GPU-box session: parity tests + smoke + default bench line.
# Usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_tests.sh [tag]'
TAG=${1:-r01t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/host.txt; nvidia-smi -L >> $OUT/host.txt
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
grep -E "passed|failed|error|exit" $OUT/pytest_gpu.log | tail -5; tail -2 $OUT/smoke.log; cat $OUT/bench.json
