#!/bin/bash
# session 16: GPU suite with the backtracking and loss-location-refinement tests
TAG=${1:-r02s16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -15 $OUT/pytest_gpu.log
