#!/bin/bash
# One bench line per BASELINE.json config at its stated per-GPU size (single GPU), --quick lines.
TAG=${1:-r02cfg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # label, args...
  local label=$1; shift
  timeout 900 python bench.py --quick --no-cpu-baseline "$@" > $OUT/cfg_${label}.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/cfg_${label}.json'))
    m=d.get('monitor',{})
    print('${label}', '%.4e PET/s'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'ms/step %.1f'%d['ms_per_step'], ('monitor %.1f GB/s'%m['GB_per_s_of_kernel_time']) if m else '')
except Exception as e:
    print('${label} FAILED', e)
PY
}
run cfg1_toy_ring --workload toy_ring --particles 10000 --turns 1000 --steps 3 --warmup 3
run cfg1_toy_ring_thin --workload toy_ring_thin --particles 10000 --turns 1000 --steps 3 --warmup 3
run cfg2_hllhc_1000turns --workload hllhc_da --particles 1000000 --turns 1000 --steps 1 --warmup 3
run cfg2_hllhc_125k --workload hllhc_da --particles 125000 --turns 1000 --steps 1 --warmup 3
run cfg3_sps_4M --workload sps_apertures --particles 4000000 --turns 100 --steps 2 --warmup 3
run cfg4_clic_quantum_1M --workload clic_dr_quantum --particles 1000000 --turns 3 --steps 1 --warmup 3
run cfg4_clic_mean_1M --workload clic_dr_mean --particles 1000000 --turns 5 --steps 1 --warmup 3
run cfg4_clic_qkick_1M --workload clic_dr_qkick --particles 1000000 --turns 5 --steps 1 --warmup 3
run cfg4_lep_qkick_1M --workload lep_qkick --particles 1000000 --turns 2 --steps 1 --warmup 3
run cfg4_lep_quantum_1M --workload lep_quantum --particles 1000000 --turns 1 --steps 1 --warmup 3
run cfg4_lep_mean_1M --workload lep_mean --particles 1000000 --turns 2 --steps 1 --warmup 3
run cfg5_lep_thick_1M --workload lep_thick --particles 1000000 --turns 5 --steps 2 --warmup 3
run cfg5_lep_thick_1M_monitor --workload lep_thick --particles 1000000 --turns 5 --steps 2 --warmup 3 --monitor
run cfg5_toy_monitor --workload toy_ring_thin --particles 1000000 --turns 20 --steps 2 --warmup 3 --monitor
