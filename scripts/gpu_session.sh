#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list, ncu full capture of the tracking kernel.
# Usage (here): gpurun --timeout 1800 -- 'bash scripts/gpu_session.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
nproc > $OUT/host.txt; nvidia-smi -L >> $OUT/host.txt
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
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
for mode in "" "--fma"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track${mode} -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline $mode > $OUT/ncu_full${mode}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_lep -f python bench.py --workload lep_thick --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_lep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_clic_mean -f python bench.py --workload clic_dr_mean --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_clic_mean.log 2>&1
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json
for f in $OUT/bench_*.json; do python - <<PY
import json
try:
    d=json.load(open('$f')); print('$f'.split('/')[-1], '%.4e'%d['value'], d.get('roofline',{}).get('frac'))
except Exception as e: print('$f', 'FAILED', e)
PY
done
