#!/bin/bash
TAG=${1:-r02s10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python scripts/diag_lost2.py > $OUT/diag_lost.log 2>&1
cat $OUT/diag_lost.log
