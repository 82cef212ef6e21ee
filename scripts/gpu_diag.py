"""GPU diagnostic: where does the kernel deviate from the oracle (bitwise)?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pp in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, pp)
import numpy as np, torch
import xtrack_b200 as xb, common

def cmp(line, p_host, turns, exact, label, **kw):
    ref = common.oracle_track(line, p_host, turns, **{k: v for k, v in kw.items() if k in ('num_ele_track',)})
    p = p_host.copy(_device='cuda:0')
    line.build_tracker(_device='cuda:0', exact_arithmetic=exact)
    if 'num_ele_track' in kw:
        line.track(p, num_elements=kw['num_ele_track'])
    else:
        line.track(p, num_turns=turns)
    got = common.by_id(p)
    dev = common.max_rel_dev(got, ref, fields=common.ALL_F64)
    nb = {f: int(np.sum(got[f] != ref[f])) for f in common.ALL_F64}
    print(f'{label:40s} exact={exact} maxrel={max(dev.values()):.2e} worst={max(dev, key=dev.get)} nbitdiff={nb}', flush=True)

# single elements
sig = common.SIGMAS['toy']
ref_p = xb.Particles(p0c=1.2e9)
def one(els, label, n=2000):
    line = xb.Line(elements=els); line.particle_ref = ref_p
    p = common.gaussian_particles(line, n, 1, sig)
    cmp(line, p, 1, True, label)
import math
one([xb.Drift(length=1.3)], 'drift')
one([xb.Multipole(knl=[0, 0.3])], 'mult order1')
one([xb.Multipole(knl=[0.1, 0.3, 2.0, 30.], ksl=[0, 0.1, 1.0])], 'mult order3')
one([xb.Multipole(knl=[0.7], hxl=0.7, length=0.5)], 'mult_h')
one([xb.Multipole(knl=[0.7, 0.2], hxl=0.7, length=0.5)], 'mult_h b1')
one([xb.Cavity(voltage=1e5, frequency=1e7, lag=180.)], 'cavity lag180')
one([xb.Cavity(voltage=1e5, frequency=1e7, lag=30.)], 'cavity lag30')
one([xb.Drift(length=1.0), xb.Multipole(knl=[0, 0.3]), xb.Drift(length=2.0)], 'd-m-d')
one([xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4)], 'dipedge lin')
one([xb.SRotation(angle=20.)], 'srot')
one([xb.LimitEllipse(a=0.05, b=0.03)], 'ellipse')
one([xb.Quadrupole(length=0.5, k1=0.3)], 'quad')
one([xb.Bend(length=1.5, angle=0.1, k0='from_h')], 'bend')
one([xb.Sextupole(length=0.5, k2=3.)], 'sext')

for name in ('toy', 'hllhc_14', 'sps'):
    line = common.toy_ring(thin=True) if name == 'toy' else common.load_line(name)
    p = common.gaussian_particles(line, 2000, 11, common.SIGMAS[name])
    for ne in (1, 2, 10, 100, 1000, len(line)):
        if ne <= len(line):
            cmp(line, p, 1, True, f'{name} first {ne} elements', num_ele_track=ne)
    for t in (1, 3, 10):
        cmp(line, p, t, True, f'{name} {t} turns')
        cmp(line, p, t, False, f'{name} {t} turns')
