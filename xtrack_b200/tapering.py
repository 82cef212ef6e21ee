"""Tapering: magnet strengths and cavity phases that compensate the synchrotron-radiation
energy loss around the ring (`xtrack/tapering.py:9-159`,
`Line.compensate_radiation_energy_loss`).

Same procedure and result as the reference: the closed orbit without radiation, the cavities
put on crest at zero frequency, the voltage shared among them raised until a particle on that
orbit loses no net energy in a turn while every radiating magnet is scaled to the particle's
momentum THERE (the reference's `XS_FLAG_SR_TAPER` passes), then `delta_taper` of every
element set to the momentum deviation found at it and the cavity phases set so that the
original voltages give the synchronous particle exactly that energy.

One difference in HOW: under `XS_FLAG_SR_TAPER` the reference's kernel scales the strengths
with the momentum of the particle it is tracking (`track_magnet.h:448-470`: a single-particle
mode).  Here every element constant is folded by the host lowering, so the scaling is a fixed
point instead: each pass is lowered with the momentum deviations the PREVIOUS pass recorded
element by element (`turn_by_turn_monitor='ONE_TURN_EBE'`), and the loop runs until the energy
balance AND the deviations used equal the deviations found -- at which point the pass is the
reference's self-consistent pass.  The coupling is weak (second order in the energy sawtooth):
the extra condition costs a few iterations.  Runs on the tracker's device, one particle.
"""
import numpy as np

from .particles import Particles

CLIGHT = 299792458.0


def _say(verbose, *args):
    if verbose:
        print(*args)


def _one_particle(line, coords, delta):
    ref = line.particle_ref
    return Particles(p0c=float(ref.get('p0c')[0]), mass0=ref.mass0, q0=ref.q0,
                     x=coords[0], px=coords[1], y=coords[2], py=coords[3], zeta=0., delta=delta,
                     _device=line.tracker.device)


def find_closed_orbit_4d(line, delta0=0.0, tol=1e-13, max_iter=30, step=1e-6):
    """The transverse closed orbit at momentum deviation `delta0` (Newton iteration on the
    one-turn map with a finite-difference Jacobian: five particles per step).  Returns the
    array (x, px, y, py)."""
    ref = line.particle_ref
    vv = np.zeros(4)
    for _ in range(max_iter):
        pts = np.tile(vv, (5, 1))
        for kk in range(4):
            pts[kk + 1, kk] += step
        pp = Particles(p0c=float(ref.get('p0c')[0]), mass0=ref.mass0, q0=ref.q0, x=pts[:, 0],
                       px=pts[:, 1], y=pts[:, 2], py=pts[:, 3], zeta=0., delta=delta0,
                       _device=line.tracker.device)
        line.track(pp)
        if np.any(pp.get('state') <= 0):
            raise RuntimeError('closed-orbit search: test particle lost')
        order = np.argsort(pp.get('particle_id'))
        out = np.stack([pp.get(nn)[order] for nn in ('x', 'px', 'y', 'py')], axis=1)
        ff = out[0] - vv
        if np.max(np.abs(ff)) < tol:
            return vv
        jac = (out[1:] - out[0]).T / step                  # d(out) / d(in)
        vv = vv - np.linalg.solve(jac - np.eye(4), ff)
    raise RuntimeError('closed-orbit search did not converge')


def compensate_radiation_energy_loss(line, delta0='zero_mean', rtol_eneloss=1e-12, max_iter=100,
                                     verbose=True, co_search_at=None, **kwargs):
    """See the module docstring.  `delta0='zero_mean'`: the momentum deviation at the start is
    chosen so that its average around the ring is zero (second round, as the reference);
    a number: the deviation at the start of the line.  `record_iterations=True` keeps the
    element-by-element records of all passes in `line._tapering_iterations`."""
    if line.tracker is None:
        raise ValueError('the line needs a tracker (line.build_tracker)')
    if line.particle_ref is None:
        raise ValueError('Particle reference is not set')
    if abs(line.particle_ref.q0) != 1:
        raise ValueError('Only |q0| = 1 is supported (for now)')
    if co_search_at is not None:
        raise NotImplementedError('co_search_at')
    names = list(line.element_names)
    if len(set(names)) != len(names):
        raise ValueError('Line must not contain repeated elements to use '
                         '`compensate_radiation_energy_loss(...)`. ')
    elements = [line.element_dict[nn] for nn in names]
    for ee in elements:
        if 'SliceCavity' in type(ee).__name__:
            raise ValueError(f"Element type '{type(ee).__name__}' is not supported for radiation "
                             'energy loss compensation.')
    record_iterations = bool(kwargs.pop('record_iterations', False))
    if record_iterations:
        line._tapering_iterations = []
    if line.config.get('XTRACK_MULTIPOLE_NO_SYNRAD', True):
        raise ValueError('radiation is off: call line.configure_radiation(model="mean") first')

    _say(verbose, 'Compensating energy loss.')
    # closed orbit without radiation and without cavity kicks (4d)
    saved_flags = dict(line.track_flags)
    line.config['XTRACK_MULTIPOLE_NO_SYNRAD'] = True
    line.track_flags['XS_FLAG_KILL_CAVITY_KICK'] = True
    try:
        co = find_closed_orbit_4d(line, delta0=0.0)
    finally:
        line.track_flags.clear()
        line.track_flags.update(saved_flags)
        line.config['XTRACK_MULTIPOLE_NO_SYNRAD'] = False

    tapered = [ee for ee in elements if hasattr(ee, 'delta_taper')]
    tapered_idx = [ii for ii, ee in enumerate(elements) if hasattr(ee, 'delta_taper')]
    cavities = [ee for ee in elements if type(ee).__name__ == 'Cavity']
    cav_idx = [ii for ii, ee in enumerate(elements) if type(ee).__name__ == 'Cavity']
    if not cavities:
        raise ValueError('no cavity in the line')

    def one_turn(delta_start):
        pp = _one_particle(line, co, delta_start)
        line.track(pp, turn_by_turn_monitor='ONE_TURN_EBE')
        mon = line.record_last_track
        if record_iterations:
            line._tapering_iterations.append(mon)
        ptau, delta = mon.get('ptau')[0], mon.get('delta')[0]
        eloss = -(ptau[-1] - ptau[0]) * float(pp.get('p0c')[0])
        return pp, mon, eloss, delta

    p_test, mon, eloss, _ = one_turn(tapered[0].delta_taper if tapered else 0.0)
    energy0 = float(np.sqrt(p_test.get('p0c')[0] ** 2 + p_test.mass0 ** 2))
    if p_test.get('state')[0] > 0 and abs(eloss) < energy0 * rtol_eneloss:
        _say(verbose, '  - No compensation needed')
        return

    beta0 = float(p_test.get('beta0')[0])
    v0 = np.array([cc.voltage for cc in cavities])
    f0 = np.array([cc.frequency for cc in cavities])
    h0 = np.array([cc.harmonic for cc in cavities])
    lag_zero = np.array([cc.lag for cc in cavities])
    phase_zero = np.array([cc.phase for cc in cavities])
    for cc in cavities:
        cc.lag_taper = 0.0
    f0_all = f0 + h0 / (line.get_length() / beta0 / CLIGHT)
    eneloss_partitioning = v0 / v0.sum()

    # all cavities on crest and at zero frequency
    for cc, lz, pz in zip(cavities, lag_zero, phase_zero):
        cc.phase_taper = np.pi / 2 - np.deg2rad(lz) - pz
        cc.voltage = 0.0
        cc.frequency = 0.0
        cc.harmonic = 0.0

    _say(verbose, 'Share energy loss among cavities (repeat until energy loss is zero)')
    num_rounds, delta_start = (2, 0.0) if delta0 == 'zero_mean' else (1, float(delta0))
    delta_used = None
    delta = ss = None
    s_elements = np.concatenate([[0.], np.cumsum([ee.length if ee.isthick_now else 0.
                                                  for ee in line.elements])])
    for rnd in range(num_rounds):
        i_iter = 0
        while True:
            if rnd == 1 and delta0 == 'zero_mean':
                delta_ave = np.trapezoid(delta, ss) / ss[-1]
                delta_start -= delta_ave
            # the strengths of this pass: scaled to the momentum the last pass found at each
            # element (the start value everywhere in the very first pass)
            if delta_used is None:
                delta_used = np.full(len(elements), delta_start)
            for ee, ii in zip(tapered, tapered_idx):
                ee.delta_taper = float(delta_used[ii])
            p_test, mon, eloss, delta_found = one_turn(delta_start)
            if p_test.get('state')[0] <= 0:
                raise RuntimeError('tapering: the test particle was lost')
            _say(verbose, f'Energy loss: {eloss:_.3f} eV             ')
            mismatch = float(np.max(np.abs(delta_found[:-1] - delta_used)))
            delta = delta_found[:-1]
            ss = s_elements[:-1]
            delta_used = delta.copy()
            if abs(eloss) < energy0 * rtol_eneloss and mismatch < 1e-13:
                break
            if abs(eloss) >= energy0 * rtol_eneloss:
                for cc, part in zip(cavities, eneloss_partitioning):
                    cc.voltage = cc.voltage + eloss * part
            i_iter += 1
            if i_iter > max_iter:
                raise RuntimeError('Maximum number of iterations reached')
    delta_all = mon.get('delta')[0]
    delta_taper_full = 0.5 * (delta_all[:-1] + delta_all[1:])     # (last point: end of the line)

    _say(verbose, '  - Set delta_taper')
    for ee, ii in zip(tapered, tapered_idx):
        ee.delta_taper = float(delta_taper_full[ii])

    _say(verbose, '  - Restore cavity voltage and frequency. Set cavity lag')
    v_synchronous = np.array([cc.voltage for cc in cavities])
    zeta_at_cav = mon.get('zeta')[0][:-1][cav_idx]
    active = np.abs(v0) > 0
    v_ratio = np.zeros_like(v0)
    v_ratio[active] = v_synchronous[active] / v0[active]
    if not np.all(np.abs(v_ratio[active]) < 1):
        raise RuntimeError('the cavity voltage cannot make up for the energy loss')
    inst_phase = np.arcsin(v_ratio)
    total_phase = inst_phase - (2 * np.pi) * f0_all * zeta_at_cav / beta0 / CLIGHT
    total_phase = np.pi - total_phase            # above transition
    phase_taper = total_phase - np.deg2rad(lag_zero) - phase_zero
    phase_taper[~active] = 0
    # (not in the reference: what the procedure found, for whoever wants to look)
    line._tapering_info = dict(closed_orbit_4d=co, delta_start=float(delta_start),
                               energy_loss_per_turn=float(v_synchronous.sum()),
                               residual_energy_loss=float(eloss))
    for cc, vv, ff, hh, pt in zip(cavities, v0, f0, h0, phase_taper):
        cc.voltage = float(vv)
        cc.frequency = float(ff)
        cc.harmonic = float(hh)
        cc.lag_taper = 0.0
        cc.phase_taper = float(pt)
