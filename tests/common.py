"""Shared helpers of the test-suite (fixtures, oracle driver, comparisons)."""
import os

import numpy as np

import xtrack_b200 as xb
import ref_oracle as ro

HERE = os.path.dirname(os.path.abspath(__file__))
LATTICES = os.path.join(HERE, 'golden', 'lattices')

COORDS = ('x', 'px', 'y', 'py', 'zeta', 'delta')
ALL_F64 = ('x', 'px', 'y', 'py', 'zeta', 'delta', 's', 'ptau', 'rpp', 'rvv')
INT_FIELDS = ('state', 'at_turn', 'at_element', 'particle_id')

# beam sizes used for the synthetic Gaussian beams
SIGMAS = {
    'hllhc_14': dict(x=2e-4, px=3e-6, y=2e-4, py=3e-6, zeta=5e-2, delta=1e-4),
    'sps': dict(x=2e-3, px=5e-5, y=1e-3, py=3e-5, zeta=0.2, delta=1e-3),
    'lep': dict(x=2e-4, px=2e-6, y=5e-5, py=1e-6, zeta=5e-3, delta=3e-4),
    'clic_dr': dict(x=1e-4, px=2e-5, y=2e-5, py=4e-6, zeta=2e-3, delta=1e-3),
    'toy': dict(x=1e-3, px=1e-4, y=1e-3, py=1e-4, zeta=5e-2, delta=1e-4),
}


def load_fixture(name):
    import gzip
    import json
    with gzip.open(os.path.join(LATTICES, name + '.json.gz'), 'rt') as fid:
        return json.load(fid)


def load_line(name, **kwargs):
    dd = load_fixture(name)
    line = xb.Line.from_dict(dd, replace_unsupported=True, **kwargs)
    if line.particle_ref is None and 'particle' in dd:
        line.particle_ref = xb.Particles.from_dict(dd['particle'])
    return line


def gaussian_particles(line, n, seed, sig, device='cpu', capacity=None, scale=1.0, **extra):
    rng = np.random.default_rng(seed)
    ref = line.particle_ref
    kw = {kk: rng.normal(0, sig[kk] * scale, n) for kk in COORDS}
    kw.update(extra)
    return xb.Particles(p0c=float(ref.get('p0c')[0]), mass0=ref.mass0, q0=ref.q0,
                        _device=device, _capacity=capacity, **kw)


def toy_ring(thin=False):
    """The 16-element ring of the reference's examples/toy_ring/000_toy_ring.py:13-36
    (4 FODO-like cells: quadrupoles, 1 m drifts, sector bends of pi/2... scaled so
    that the ring closes), or a thin variant with multipoles and one cavity."""
    import math
    n_bends = 4
    if not thin:
        els, names = [], []
        for ii in range(4):
            els += [xb.Quadrupole(length=0.3, k1=0.1 if ii % 2 == 0 else -0.7),
                    xb.Drift(length=1.0),
                    xb.Bend(length=3.0, angle=2 * math.pi / n_bends, k0='from_h', model='full',
                            edge_entry_active=0, edge_exit_active=0),
                    xb.Drift(length=1.0)]
        line = xb.Line(elements=els)
    else:
        els = []
        for ii in range(8):
            els += [xb.Drift(length=1.0),
                    xb.Multipole(knl=[0, 0.3 if ii % 2 == 0 else -0.3]),
                    xb.Drift(length=1.0),
                    xb.Multipole(knl=[2 * math.pi / 8], hxl=2 * math.pi / 8, length=0.5)]
        els.append(xb.Cavity(voltage=1e5, frequency=1e7, lag=180.))
        line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=1.2e9, mass0=xb.PROTON_MASS_EV)
    return line


_OMP_VARIANT = {'serial': 'omp', 'noise': 'noise_omp', 'synrad': 'synrad_omp',
                'synrad_noise': 'synrad_noise_omp'}


def oracle_track(line, particles, num_turns, *, ele_start=0, num_ele_track=None,
                 flag_end_turn_actions=True, monitor=None, flag_monitor=0, variant='serial',
                 track_flags=0, parallel=False):
    """Runs the reference-header oracle on a copy of `particles`; returns the fields
    ordered by particle_id.  `parallel`: the OpenMP build of the same variant -- particles
    are independent, so surviving particles come out bit-identical to the serial build.
    Lost ones need not: the reference's OpenMP context skips lost particles in every
    per-particle block (headers/track.h:43), including the EXIT transformation of a
    misaligned element, so a particle lost inside a shifted aperture keeps the element-frame
    coordinates there, while the serial and the GPU contexts (no state test, track.h:20-31,
    51-55) transform it back.  The product follows the GPU context; whenever the run lost
    particles the result is therefore recomputed with the serial build."""
    if parallel:
        out = oracle_track(line, particles, num_turns, ele_start=ele_start,
                           num_ele_track=num_ele_track,
                           flag_end_turn_actions=flag_end_turn_actions, monitor=monitor,
                           flag_monitor=flag_monitor, variant=_OMP_VARIANT.get(variant, variant),
                           track_flags=track_flags)
        if (out['state'] > 0).all():
            return out
        if monitor is not None:
            for vv in monitor.arrays.values():
                vv[:] = 0
    hp = ro.HostParticles.from_particles(particles)
    re = ro.RefElements(line.elements)
    ro.track_line(hp, re, num_turns=num_turns, ele_start=ele_start,
                  num_ele_track=len(line) if num_ele_track is None else num_ele_track,
                  flag_end_turn_actions=flag_end_turn_actions,
                  flag_reset_s_at_end_turn=line.reset_s_at_end_turn, line_length=line.get_length(),
                  global_xy_limit=line.config['XTRACK_GLOBAL_XY_LIMIT'], monitor=monitor,
                  flag_monitor=flag_monitor, variant=variant, track_flags=track_flags)
    return hp.sorted_by_id()


def by_id(particles):
    """Fields of an xb.Particles ordered by particle_id (allocated slots only)."""
    st = particles.get('state')
    alloc = np.where(st > xb.LAST_INVALID_STATE)[0]
    order = alloc[np.argsort(particles.get('particle_id')[alloc], kind='stable')]
    return {nn: particles.get(nn)[order] for nn, _ in xb.Particles.per_particle_vars}


def max_rel_dev(got, ref, fields=COORDS, mask=None):
    """max over particles of |got-ref| / scale, scale = max(|ref|) of that coordinate over
    the beam (the "1e-12 relative" of BASELINE.json is read against the beam size)."""
    out = {}
    for ff in fields:
        a, b = got[ff], ref[ff]
        if mask is not None:
            a, b = a[mask], b[mask]
        scale = np.max(np.abs(b)) if len(b) and np.max(np.abs(b)) > 0 else 1.0
        out[ff] = float(np.max(np.abs(a - b)) / scale) if len(b) else 0.0
    return out


def libm_yardstick(line, particles, num_turns, ref=None, fields=COORDS, variant='serial', **kw):
    """The reference's OWN sensitivity to the libm it links: max_rel_dev between the clean
    oracle and the oracle rebuilt with +-1 ulp noise on every transcendental result
    (oracle variant `noise`, shim/xobjects/headers/ulp_noise.h).  Lattices with per-particle
    sin/cos/sinh/atan2... calls (cavities, RF-multipoles, thick magnets) amplify a last-bit
    difference of those calls over the turns; two correct libms (glibc here, CUDA's on the
    device) differ at exactly that level, so this is the floor below which agreement with
    the CPU reference cannot be demanded of ANY device implementation."""
    if ref is None:
        ref = oracle_track(line, particles, num_turns, variant=variant, **kw)
    noisy = oracle_track(line, particles, num_turns,
                         variant={'serial': 'noise', 'synrad': 'synrad_noise'}[variant], **kw)
    return max_rel_dev(noisy, ref, fields=fields, mask=ref['state'] > 0)


# Parity bar (BASELINE.json north_star): 1e-12 relative over the first 10 turns.  It is
# applied as written wherever the reference itself is reproducible to that level; where
# the reference's result moves by more than that under +-1 ulp of libm noise, the bar is
# that yardstick times a small factor (see libm_yardstick and DESIGN.md "Parity").
RTOL = 1e-12
EXACT_FACTOR = 3.0      # EXACT kernel (no FMA contraction): device libm vs glibc only
FMA_FACTOR = 8.0        # FMA kernel: every mul+add pair rounds once instead of twice


def assert_parity(got, ref, yard, exact, mask=None, fields=COORDS, label='', floor=None):
    dev = max_rel_dev(got, ref, fields=fields, mask=mask)
    factor = EXACT_FACTOR if exact else FMA_FACTOR
    if floor is None:
        floor = RTOL if exact else 2 * RTOL
    # the yardstick is one noise realisation: take it per plane group (the coordinates of
    # a group are coupled by the optics), not per single coordinate
    groups = (('x', 'px', 'y', 'py'), ('zeta', 'delta'))
    gy = {}
    for gg in groups:
        vv = max([yard.get(ff, 0.0) for ff in gg])
        for ff in gg:
            gy[ff] = vv
    bad = {ff: (dev[ff], max(floor, factor * gy.get(ff, 0.0))) for ff in fields
           if dev[ff] > max(floor, factor * gy.get(ff, 0.0))}
    print(label, 'exact' if exact else 'fma', 'dev', {k: '%.1e' % v for k, v in dev.items()},
          'yardstick', {k: '%.1e' % v for k, v in yard.items()})
    assert not bad, f'{label}: deviation above tolerance (got, allowed): {bad}'
    return dev


def seed_rng_host(particles, seeds):
    """Seeds the per-particle generator of HOST particles with the reference's own
    `Particles_initialize_rand_gen` (oracle build of particles/rng_src/particles_rng.h:12-28;
    the product seeds on the device with `xtb_rng_init`)."""
    hp = ro.HostParticles.from_particles(particles)
    assert np.array_equal(hp.arrays['particle_id'], particles.get('particle_id')), \
        'seed before any loss (slot order = id order)'
    ro.init_rand_gen(hp, seeds)
    for nn in ro.U32_VARS:
        setattr(particles, nn, hp.arrays[nn])
    return particles


def beam_monitor_ring(kind, n=400, turns=6):
    """SPS ring (losses on the way) with a BeamPositionMonitor / BeamSizeMonitor in the middle:
    4 slots per turn (sampling_frequency = 4 frev) so that zeta sorts the beam into several
    slots, id range restricted.  Returns (line, reference line for the oracle, its monitor,
    our monitor, host particles)."""
    line = load_line('sps')
    cls = getattr(xb, kind)
    frev = 299792458.0 / line.get_length()
    kw = dict(particle_id_range=(15, n - 20), start_at_turn=1, stop_at_turn=turns - 1,
              frev=frev, sampling_frequency=4 * frev)
    if kind == 'BeamProfileMonitor':
        kw.update(nx=24, x_range=0.03, ny=16, y_range=(-0.008, 0.012))
    els = list(line.elements)
    mon = cls(**kw)
    els.insert(len(els) // 2, mon)
    line2 = xb.Line(elements=els)
    line2.particle_ref = line.particle_ref
    mon_ref = cls(**kw)
    if kind == 'BeamProfileMonitor':
        mon_ref._host = {'counts_x': np.zeros(mon_ref.sample_size * mon_ref.nx),
                         'counts_y': np.zeros(mon_ref.sample_size * mon_ref.ny)}
    else:
        mon_ref._host = np.zeros((5, mon_ref.n_slots))
    els_ref = list(els)
    els_ref[len(line.elements) // 2] = mon_ref
    sig = dict(SIGMAS['sps'])
    sig['zeta'] = 0.25 * line.get_length() / 4          # spread over neighbouring slots
    p_host = gaussian_particles(line2, n, 3, sig, scale=5.0)
    return line2, els_ref, mon_ref, mon, p_host


def lep_with_apertures(every=40, half_x=4e-3, half_y=1.5e-3):
    """The LEP thick lattice with a LimitRect after every `every`-th element: losses between
    and inside runs of thick magnets (exercises the loss bookkeeping of the thick run loop)."""
    line = load_line('lep')
    els = []
    for ii, el in enumerate(line.elements):
        els.append(el)
        if ii % every == every - 1:
            els.append(xb.LimitRect(min_x=-half_x, max_x=half_x, min_y=-half_y, max_y=half_y))
    line2 = xb.Line(elements=els)
    line2.particle_ref = line.particle_ref
    return line2

