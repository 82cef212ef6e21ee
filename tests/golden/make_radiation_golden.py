"""Generates tests/golden/clic_dr_quantum_stats.json: beam moments of the CLIC-DR stand-in under
quantum synchrotron radiation, tracked by the REFERENCE's own C code (oracle/_ref, OpenMP build
with radiation) -- run HERE, where /root/reference exists; the GPU box reads the fixture.

512 electrons, Gaussian start (tests/common.SIGMAS['clic_dr']), generator seeds 1..512 through
the reference's `Particles_initialize_rand_gen`; moments after 100, 200 and 300 turns (0.4
longitudinal damping times: the mean energy has moved to the synchronous phase, the energy
spread is being rebuilt by the quantum excitation, the betatron amplitudes are damping).

Usage:  python tests/golden/make_radiation_golden.py        (~10 min on 8 cores)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
for pp in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, pp)

import common                     # noqa: E402
import ref_oracle as ro           # noqa: E402

N, CHECKPOINTS, SEED = 512, (100, 200, 300), 1
FIELDS = ('x', 'px', 'y', 'py', 'zeta', 'delta')


def synchronous_zeta(line):
    """The ring has one cavity (harmonic 2852, 4.5 MV, phase pi) and loses U0 ~ 3.9 MeV per
    turn; nothing tapers the magnets or re-phases the cavity here (`compensate_radiation_
    energy_loss` is outside the contract), so the beam is started at the synchronous phase:
    V sin(k zeta_s) = U0, with U0 from one turn of the mean model on the reference particle."""
    import xtrack_b200 as xb
    ref = line.particle_ref
    p0c = float(ref.get('p0c')[0])
    one = xb.Particles(p0c=p0c, mass0=ref.mass0, q0=ref.q0)
    line.configure_radiation(model='mean')
    els = [ee for ee in line.elements if type(ee).__name__ != 'Cavity']
    res = common.oracle_track(xb.Line(elements=els), one, 1, variant='synrad')
    u0 = -float(res['ptau'][0]) * p0c                 # eV lost in one turn, no RF
    cav = [ee for ee in line.elements if type(ee).__name__ == 'Cavity'][0]
    k = 2 * np.pi * cav.harmonic / line.get_length()
    return float(np.arcsin(u0 / cav.voltage) / k), u0


def initial_beam(line):
    zeta_s, u0 = synchronous_zeta(line)
    line.configure_radiation(model='quantum')
    p = common.gaussian_particles(line, N, SEED, common.SIGMAS['clic_dr'])
    p.zeta = p.get('zeta') + zeta_s
    return p, zeta_s, u0


def main():
    line = common.load_line('clic_dr')
    line.configure_radiation(model='quantum')
    p, zeta_s, u0 = initial_beam(line)
    print('U0 = %.4e eV, synchronous zeta = %.6f m' % (u0, zeta_s), flush=True)
    common.seed_rng_host(p, np.arange(1, N + 1, dtype=np.uint32))
    hp = ro.HostParticles.from_particles(p)
    re_ = ro.RefElements(line.elements)
    ro.load('synrad_omp').xt_ref_set_num_threads(int(os.environ.get('XTB_GOLDEN_THREADS', 8)))
    out = {'n_particles': N, 'seed': SEED, 'lattice': 'clic_dr', 'model': 'quantum',
           'zeta_offset': zeta_s, 'u0_eV': u0,
           'source': 'oracle/_ref synrad_omp (reference C headers)', 'turns': {}}
    done = 0
    for tt in CHECKPOINTS:
        ro.track_line(hp, re_, num_turns=tt - done, ele_start=0, num_ele_track=len(line),
                      flag_end_turn_actions=1, flag_reset_s_at_end_turn=1,
                      line_length=line.get_length(), variant='synrad_omp')
        done = tt
        res = hp.sorted_by_id()
        alive = res['state'] > 0
        out['turns'][str(tt)] = {
            'n_alive': int(alive.sum()),
            'mean': {ff: float(res[ff][alive].mean()) for ff in FIELDS},
            'std': {ff: float(res[ff][alive].std()) for ff in FIELDS}}
        print(tt, out['turns'][str(tt)], flush=True)
    with open(os.path.join(HERE, 'clic_dr_quantum_stats.json'), 'w') as fid:
        json.dump(out, fid, indent=1)


if __name__ == '__main__':
    main()
