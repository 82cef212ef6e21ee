#!/bin/bash
# session 28: by-function profile of the quantum radiation kernel (CLIC-DR, photon by photon)
TAG=${1:-r02s28}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_clic_quantum -f python bench.py --workload clic_dr_quantum --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu.log 2>&1
du -sh $OUT; tail -2 $OUT/ncu.log
