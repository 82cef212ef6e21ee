"""Diagnostic 2 (GPU box): the slow stretch (turns 800-1000 of the DA beam) under variations."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import xtrack_b200 as xb
from xtrack_b200 import _cabi

line = bench.load_line('hllhc_14')
n = 1_000_000
ic = bench.initial_conditions('hllhc_da', line, n, 0)
ref = line.particle_ref
p = xb.Particles(p0c=float(ref.get('p0c')[0]), mass0=ref.mass0, q0=ref.q0, _device='cuda:0', **ic)
line.build_tracker(_device='cuda:0')


def timed(pp, turns, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); line.track(pp, num_turns=turns, **kw); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / turns

timed(p.copy(), 5)
line.track(p, num_turns=800)
torch.cuda.synchronize()
print('state at turn 800: lost', int((p.get('state') <= 0).sum()))

def report(label, q, ms):
    st = q.get('state')
    print(f'{label:34s} {ms:7.3f} ms/turn  lost {int((st <= 0).sum())}', flush=True)

q = p.copy(); report('turns 800-1000 in ONE launch', q, timed(q, 200))
q = p.copy(); ms = [timed(q, 20) for _ in range(10)]; print('  in launches of 20 turns:', ['%.2f' % v for v in ms])
q = p.copy(); report('turns 800-1800 in ONE launch', q, timed(q, 1000))
