#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list, ncu full capture of the tracking kernel.
# Usage (here): gpurun --timeout 1800 -- 'bash scripts/gpu_session.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
nproc > $OUT/host.txt; nvidia-smi -L >> $OUT/host.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err
timeout 300 python bench.py --fma --no-cpu-baseline > $OUT/bench_fma.json 2>> $OUT/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err
for wl in sps_apertures lep_thick; do
  timeout 400 python bench.py --workload $wl --quick --steps 3 --warmup 1 --turns 5 --particles 500000 > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
for wl in lep_quantum clic_dr_quantum lep_mean clic_dr_mean lep_qkick clic_dr_qkick; do
  timeout 400 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 2 --particles 300000 > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --turns 2 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
# (gpurun merges at most 64 MiB back: one capture with source here, the thick / radiation
# kernels in scripts/gpu_session_b.sh)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track--fma -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline --fma > $OUT/ncu_full--fma.log 2>&1
du -sh $OUT
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json
for f in $OUT/bench_*.json; do python - <<PY
import json
try:
    d=json.load(open('$f')); print('$f'.split('/')[-1], '%.4e'%d['value'], d.get('roofline',{}).get('frac'))
except Exception as e: print('$f', 'FAILED', e)
PY
done
