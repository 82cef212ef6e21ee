#!/bin/bash
# Second half of the evidence session (gpurun merges at most 64 MiB back per call): full ncu
# captures of the thick and of the radiation kernel, one bench line per BASELINE config.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_lep -f python bench.py --workload lep_thick --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_lep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_clic_mean -f python bench.py --workload clic_dr_mean --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_clic_mean.log 2>&1
du -sh $OUT
bash scripts/gpu_r02_configs.sh $TAG
