"""CPU tier, SURVEY.md §8(a) rows R, S, Mon2 and §8(b): synchrotron radiation (mean and
quantum models with the per-particle Tausworthe generator), the LastTurnsMonitor, and
the C-ABI surface -- the device code through its HOST build (tests/hostsim, test
infrastructure) against the reference-header oracle compiled WITH radiation
(oracle variant `synrad`: no -DXTRACK_MULTIPOLE_NO_SYNRAD).  Same libm, same IEEE
operation order, same random stream -> the expected result is bit-identity.
The real kernels are covered by tests/test_gpu_radiation.py (-m gpu).
"""
import ctypes as ct
import os
import re

import numpy as np
import pytest

import xtrack_b200 as xb
import common
import hostsim

QE = 1.602176634e-19
CLIGHT = 299792458.0
EPSILON_0 = 8.8541878128e-12
HBAR = 1.054571817e-34


def _track(line, p_host, num_turns, **kw):
    p = p_host.copy()
    hostsim.build_hostsim_tracker(line)
    line.track(p, num_turns=num_turns, **kw)
    return p


@pytest.mark.parametrize('model', ['mean', 'quantum'])
@pytest.mark.parametrize('name', ['clic_dr', 'lep'])
def test_radiation_ten_turns_bit_identical(name, model):
    """BASELINE config #4 stand-ins (thin CLIC-DR, thick LEP): 10 turns with
    configure_radiation(model) (line.py:4744-4837).  The quantum model draws from the
    particle's own generator state (base_rng.h:23-42), which travels with the particle: the
    photon sequence, and with it every coordinate and the final generator state, must be
    reproduced exactly."""
    line = common.load_line(name)
    line.configure_radiation(model=model)
    n = 48
    p_host = common.gaussian_particles(line, n, 5, common.SIGMAS[name])
    if model == 'quantum':
        common.seed_rng_host(p_host, np.arange(1, n + 1, dtype=np.uint32) * 7919)
    ref = common.oracle_track(line, p_host, 10, variant='synrad')
    got = common.by_id(_track(line, p_host, 10))
    for ff in common.ALL_F64 + ('state', 'at_turn', 'at_element') + xb.particles.U32_VARS:
        assert np.array_equal(got[ff], ref[ff]), ff
    # radiation did happen: the beam lost energy
    assert ref['delta'].mean() < -1e-3
    if model == 'quantum':
        assert not np.array_equal(ref['_rng_s1'], p_host.get('_rng_s1'))


def test_radiation_off_is_the_no_synrad_build():
    """configure_radiation(None) on a line that had it on: back to the radiation-free
    program (XTRACK_MULTIPOLE_NO_SYNRAD), identical to the `serial` oracle."""
    line = common.load_line('clic_dr')
    line.configure_radiation(model='mean')
    line.configure_radiation(model=None)
    p_host = common.gaussian_particles(line, 30, 6, common.SIGMAS['clic_dr'])
    ref = common.oracle_track(line, p_host, 3)
    got = common.by_id(_track(line, p_host, 3))
    for ff in common.ALL_F64:
        assert np.array_equal(got[ff], ref[ff]), ff


def _bend_setup(thick, flag):
    """The single 2 T, 1 m dipole of the reference's tests/test_radiation.py:27-72."""
    L_bend, B_T = 1.0, 2.0
    p0c = 5e9
    P0_J = p0c / CLIGHT * QE
    theta = B_T * QE / P0_J * L_bend
    if thick:
        el = xb.Bend(length=L_bend, angle=theta, k0='from_h', radiation_flag=flag)
    else:
        el = xb.Multipole(knl=[theta], length=L_bend, hxl=theta, radiation_flag=flag)
    line = xb.Line(elements=[el])
    line.particle_ref = xb.Particles(p0c=p0c, mass0=xb.ELECTRON_MASS_EV)
    line.config['XTRACK_MULTIPOLE_NO_SYNRAD'] = False
    return line, theta, L_bend


@pytest.mark.parametrize('thick', [False, True], ids=['thin', 'thick'])
def test_single_bend_energy_loss(thick):
    """tests/test_radiation.py:27-118 of the reference: the mean model reproduces the
    classical radiated power (4e-5), the quantum model has the same mean (5e-3); and both
    agree with the reference's own C to the bit."""
    n = 20000
    kw = dict(p0c=5e9, x=np.zeros(n), px=1e-4, py=-1e-4, mass0=xb.ELECTRON_MASS_EV)
    res = {}
    for flag in (1, 2):
        line, theta, L_bend = _bend_setup(thick, flag)
        p_host = xb.Particles(**kw)
        common.seed_rng_host(p_host, np.arange(1, n + 1, dtype=np.uint32) * 104729)
        ref = common.oracle_track(line, p_host, 1, variant='synrad')
        line._extra_config['_needs_rng'] = False        # seeded above, on the host
        got = common.by_id(_track(line, p_host, 1))
        for ff in common.ALL_F64 + ('state',) + xb.particles.U32_VARS:
            assert np.array_equal(got[ff], ref[ff]), (flag, ff)
        res[flag] = got
    gamma0 = float(p_host.get('gamma0')[0])
    gamma = gamma0      # delta = 0
    rho_0 = L_bend / theta
    mass0_kg = xb.ELECTRON_MASS_EV * QE / CLIGHT ** 2
    r0 = QE ** 2 / (4 * np.pi * EPSILON_0 * mass0_kg * CLIGHT ** 2)
    Ps = (2 * r0 * CLIGHT * mass0_kg * CLIGHT ** 2 * gamma0 ** 2 * gamma ** 2) / (3 * rho_0 ** 2)
    dE_eV = -Ps * (L_bend / CLIGHT) / QE
    dE_trk = res[1]['ptau'][0] * 5e9
    np.testing.assert_allclose(dE_trk, dE_eV, rtol=4e-5, atol=0)
    np.testing.assert_allclose(np.mean(res[2]['delta']), res[1]['delta'][0], rtol=2e-2, atol=0)


def test_unseeded_generator_kills_particles():
    """A generator state of all zeros is an error: RandomUniform_generate kills the particle
    with state -20 (random_src/uniform.h:34-53, RNG_ERR_SEEDS_NOT_SET; kill semantics of
    local_particle_custom_api.h:248-256).  The reference's photon loop would then spin for
    ever in RandomExponential_generate (exponential.h:19-25 retries while the draw is 0), so
    there is no oracle run here: the device code leaves the loop on the error and the
    particle ends as the reference's kill leaves it.  (Line.track seeds before tracking, as
    tracker.py:1364-1365 does, so this only guards a caller who bypasses it.)"""
    line, _, _ = _bend_setup(False, 2)
    p_host = xb.Particles(p0c=5e9, x=np.zeros(10), px=1e-4, mass0=xb.ELECTRON_MASS_EV)
    line._extra_config['_needs_rng'] = False
    got = common.by_id(_track(line, p_host, 1))
    assert np.all(got['state'] == -20)
    assert np.all(got['at_element'] == 0) and np.all(got['at_turn'] == 0)
    for ff in ('x', 'px', 'y', 'py', 'zeta'):
        assert np.all(got[ff] == 1e30), ff
    assert np.all(got['delta'] == -1.0)


def test_last_turns_monitor_golden():
    """The reference's tests/test_monitor.py:195-231, golden arrays included."""
    particles = xb.Particles(p0c=6.5e12, x=[1, 2, 3, 4, 5, 6])
    monitor = xb.LastTurnsMonitor(n_last_turns=5, particle_id_range=(1, 5))
    line = xb.Line(elements=[monitor])
    hostsim.build_hostsim_tracker(line)
    for turn in range(10):
        line.track(particles, num_turns=1)
        x = particles.get('x').copy()
        x += np.array([1, -1, 2, -2, 3, -3.])
        particles.x = x
        st = particles.get('state').copy()
        if turn == 2:
            st[1] = 0
        if turn == 4:
            st[2] = 0
        if turn == 6:
            st[3] = 0
        particles.state = st
    assert np.all(monitor.particle_id == np.array([[0, 0, 1, 1, 1], [2] * 5, [3] * 5, [4] * 5]))
    assert np.all(monitor.at_turn == np.array([np.clip(n - np.arange(4, -1, -1), 0, None)
                                               for n in (2, 4, 6, 9)]))
    assert np.all(monitor.x == np.array([[0, 0, 2, 1, 0], [3, 5, 7, 9, 11], [0, -2, -4, -6, -8],
                                         [20, 23, 26, 29, 32]]))


def test_last_turns_monitor_in_ring_vs_oracle():
    """LastTurnsMonitor inside a lossy ring (every_n_turns = 2): ring-buffer content equal
    to the reference's own monitor code (monitors/last_turns_monitor.h:16-55)."""
    import ref_oracle as ro
    line = common.load_line('sps')
    n = 300
    els = list(line.elements)
    mon = xb.LastTurnsMonitor(n_last_turns=3, particle_id_range=(10, 250), every_n_turns=2)
    els.insert(len(els) // 2, mon)
    line2 = xb.Line(elements=els)
    line2.particle_ref = line.particle_ref
    p_host = common.gaussian_particles(line2, n, 3, common.SIGMAS['sps'], scale=6.0)
    # oracle: its own host-side ring buffer
    mon_ref = xb.LastTurnsMonitor(n_last_turns=3, particle_id_range=(10, 250), every_n_turns=2)
    mon_ref._host = {nn: np.zeros(mon.num_particles * (1 if nn == 'lost_at_offset' else 3),
                                  dtype=np.uint32 if nn in ('lost_at_offset', 'particle_id', 'at_turn')
                                  else np.float32)
                     for nn in ('lost_at_offset', 'particle_id', 'at_turn', 'x', 'px', 'y', 'py',
                                'delta', 'zeta')}
    els_ref = list(els)
    els_ref[len(line.elements) // 2] = mon_ref
    hp = ro.HostParticles.from_particles(p_host)
    ro.track_line(hp, ro.RefElements(els_ref), num_turns=9, ele_start=0, num_ele_track=len(els),
                  flag_end_turn_actions=True, flag_reset_s_at_end_turn=True,
                  line_length=line2.get_length(), global_xy_limit=1.0)
    ref = hp.sorted_by_id()
    assert 5 < (ref['state'] <= 0).sum() < n
    got = common.by_id(_track(line2, p_host, 9))
    assert np.array_equal(got['state'], ref['state'])
    dd = mon.allocate()
    for nn in ('lost_at_offset', 'particle_id', 'at_turn'):
        assert np.array_equal(dd[nn].numpy().view(np.uint32), mon_ref._host[nn]), nn
    for nn in ('x', 'px', 'y', 'py', 'delta', 'zeta'):
        assert np.array_equal(dd[nn].numpy(), mon_ref._host[nn]), nn


@pytest.mark.parametrize('kind', ['BeamPositionMonitor', 'BeamSizeMonitor'])
def test_beam_monitors_vs_oracle(kind):
    """monitors/beam_position_monitor.h:16-58, beam_size_monitor.h:16-64 against the
    reference's own code: counts identical, sums equal up to the order of the additions
    (the reference adds with atomics; a few ulp of the sum)."""
    import ref_oracle as ro
    turns = 6
    line2, els_ref, mon_ref, mon, p_host = common.beam_monitor_ring(kind, turns=turns)
    hp = ro.HostParticles.from_particles(p_host)
    ro.track_line(hp, ro.RefElements(els_ref), num_turns=turns, ele_start=0,
                  num_ele_track=len(els_ref), flag_end_turn_actions=True,
                  flag_reset_s_at_end_turn=True, line_length=line2.get_length(),
                  global_xy_limit=1.0)
    ref = hp.sorted_by_id()
    got = common.by_id(_track(line2, p_host, turns))
    assert np.array_equal(got['state'], ref['state'])
    assert mon_ref._host[0].sum() > 100 and (mon_ref._host[0] > 0).sum() >= 6
    assert np.array_equal(mon.count, mon_ref._host[0])
    for ii, nn in enumerate(mon.properties):
        np.testing.assert_allclose(getattr(mon, nn), mon_ref._host[ii], rtol=1e-13, atol=1e-18,
                                   err_msg=nn)
    filled = mon.count > 0
    np.testing.assert_allclose(mon.x_mean[filled], (mon_ref._host[1] / mon_ref._host[0])[filled],
                               rtol=1e-12)
    if kind == 'BeamSizeMonitor':
        assert np.all(mon.x_std[filled & (mon.count > 1)] > 0)
    # dict round trip keeps parameters and record
    mon2 = type(mon).from_dict(mon.to_dict())
    assert mon2.n_slots == mon.n_slots and np.array_equal(mon2.count, mon.count)


def test_beam_profile_monitor_vs_oracle():
    """monitors/beam_profile_monitor.h:15-80 against the reference's own code: the histograms
    (integer counts) are identical."""
    import ref_oracle as ro
    turns = 6
    line2, els_ref, mon_ref, mon, p_host = common.beam_monitor_ring('BeamProfileMonitor', n=400,
                                                                    turns=turns)
    hp = ro.HostParticles.from_particles(p_host)
    ro.track_line(hp, ro.RefElements(els_ref), num_turns=turns, ele_start=0,
                  num_ele_track=len(els_ref), flag_end_turn_actions=True,
                  flag_reset_s_at_end_turn=True, line_length=line2.get_length(),
                  global_xy_limit=1.0)
    ref = hp.sorted_by_id()
    got = common.by_id(_track(line2, p_host, turns))
    assert np.array_equal(got['state'], ref['state'])
    assert mon_ref._host['counts_x'].sum() > 100 and (mon_ref._host['counts_x'] > 0).sum() > 20
    assert np.array_equal(mon.counts_x, mon_ref._host['counts_x'])
    assert np.array_equal(mon.counts_y, mon_ref._host['counts_y'])
    assert mon.x_intensity.shape == (mon.sample_size, 24) and mon.y_intensity.shape[1] == 16
    assert len(mon.x_edges) == 25 and abs(mon.x_grid[0] - (-0.015 + 0.03 / 48)) < 1e-15
    mon2 = type(mon).from_dict(mon.to_dict())
    assert np.array_equal(mon2.counts_x, mon.counts_x) and mon2.dy == mon.dy


def test_cabi_library_exports_every_declared_symbol():
    """§8(b): libxtb200.so loads without a GPU and exports every function that
    include/xtb200.h declares (no compute call is made here)."""
    from xtrack_b200 import _cabi, build
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
    text = open(os.path.join(root, 'include', 'xtb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = sorted(set(re.findall(r'\b(xtb_[a-z0-9_]+)\s*\(', text)))
    assert len(declared) >= 10, declared
    build.build()
    lib = ct.CDLL(build.LIB)
    missing = [nn for nn in declared if not hasattr(lib, nn)]
    assert not missing, missing
    # the Python binding refers to nothing that the header does not declare
    src = open(os.path.join(root, 'xtrack_b200', '_cabi.py')).read()
    used = set(re.findall(r'\blib\.(xtb_[a-z0-9_]+)', src))
    assert used <= set(declared), used - set(declared)
    lib.xtb_last_error_string.restype = ct.c_char_p
    assert isinstance(lib.xtb_last_error_string(), bytes)


def test_no_cpu_fallback():
    """The product refuses to track on the host: no CPU path exists (north star)."""
    line = common.toy_ring(thin=True)
    p = common.gaussian_particles(line, 10, 1, common.SIGMAS['toy'])
    with pytest.raises(Exception):
        line.build_tracker(_device='cpu')
        line.track(p, num_turns=1)


def test_photon_spectrum_moments_from_energy_loss():
    """tests/test_radiation.py:119-158 checks the first two moments of the emitted photon
    spectrum through the photon log (E_ave = 8 sqrt(3)/45 E_crit to 1e-2, the standard
    deviation from <E^2> = 11/27 E_crit^2).  There is no photon log here (out of scope), but the
    energy loss of a particle is a compound Poisson sum: mean = lambda <u>, variance =
    lambda <u^2>, lambda = 5/(2 sqrt 3) alpha gamma theta.  Both moments of the spectrum follow
    from the beam's energy-loss statistics."""
    n = 40000
    line, theta, L_bend = _bend_setup(False, 2)
    p0c = 5e9
    p = xb.Particles(p0c=p0c, x=np.zeros(n), mass0=xb.ELECTRON_MASS_EV)
    common.seed_rng_host(p, (np.arange(1, n + 1, dtype=np.uint64) * 7919 % (1 << 32)).astype(np.uint32))
    line._extra_config['_needs_rng'] = False
    got = common.by_id(_track(line, p, 1))
    dE = -got['ptau'] * p0c                          # eV lost per particle
    gamma = float(got['gamma0'][0])
    hbar = 1.054571817e-34
    mass0_kg = xb.ELECTRON_MASS_EV * QE / CLIGHT ** 2
    B_T = 2.0
    E_crit_eV = 3 * QE * hbar * gamma ** 2 * B_T / (2 * mass0_kg) / QE
    lam = 5 / (2 * np.sqrt(3)) * 0.0072973525693 * gamma * theta
    u_mean = np.mean(dE) / lam
    u_sq = np.var(dE) / lam
    np.testing.assert_allclose(u_mean, 8 * np.sqrt(3) / 45 * E_crit_eV, rtol=1e-2)
    np.testing.assert_allclose(u_sq, 11 / 27 * E_crit_eV ** 2, rtol=4e-2)
