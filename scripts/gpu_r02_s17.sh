#!/bin/bash
# session 17: GPU tests of the quantum-kick model (synthetic tables) and the BeamStatsMonitor
TAG=${1:-r02s17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_quantum_kick.py tests/test_beam_stats_monitor.py tests/test_radiation_monitors_hostsim.py -m gpu -q > $OUT/pytest.log 2>&1; tail -25 $OUT/pytest.log
