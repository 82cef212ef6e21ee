"""The reference arm of bench.py (`--impl reference`: the reference's own C on the host cores)
runs without a GPU and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'libxt_ref_omp.so')),
                    reason='oracle/_ref not built (needs /root/reference)')
def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0', '--cpu-seconds', '1'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ll for ll in out.stdout.splitlines() if ll.strip()]
    assert len(lines) == 1, lines            # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'particle-element-turns/s'
    assert d['unit'] == 'particle-element-turns/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['dtype'] == 'f64' and d['vs_baseline'] is None
    assert d['value'] > 1e6
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e2e = d['e2e']
    assert e2e['value'] == d['value'] and e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'hllhc' in d['config']['workload']
