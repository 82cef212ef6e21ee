#!/bin/bash
# Round-2 session 5: GPU suite (bit-exact thin rings), bench, by-function profile after the RF rework.
TAG=${1:-r02s5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|bit-identical" $OUT/pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
cat $OUT/bench.json
for wl in sps_apertures lep_thick; do
  timeout 400 python bench.py --workload $wl --quick --steps 3 --warmup 2 --turns 10 --particles 1000000 > $OUT/bench_$wl.json 2>> $OUT/bench.err
  cat $OUT/bench_$wl.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
