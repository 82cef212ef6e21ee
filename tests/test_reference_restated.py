"""Restatements of the reference's own self-contained tests for the tracking path (SURVEY.md
§8c), run on the host build of the device code (CPU tier) and, marked `gpu`, on the B200.
Same inputs, same assertions and tolerances as the cited reference tests."""
import numpy as np
import pytest
import torch

import xtrack_b200 as xb
import common


def _build(line, on_gpu):
    if on_gpu:
        line.build_tracker(_device='cuda:0')
        return 'cuda:0'
    import hostsim
    hostsim.build_hostsim_tracker(line)
    return 'cpu'


BACKENDS = [pytest.param(False, id='hostsim'), pytest.param(True, id='gpu', marks=pytest.mark.gpu)]


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_aperture_turn_ele_and_monitor(on_gpu):
    """tests/test_aperture_turn_ele_and_monitor.py:15-117: 10 000 drift slices, particles fly
    out through the global aperture; loss s / turn / element and the turn-by-turn monitor."""
    n_part = 100
    tot_length, n_slices, n_turns = 2., 10000, 3
    line = xb.Line(elements=n_slices * [xb.Drift(length=tot_length / n_slices)],
                   element_names=[f'drift{ii}' for ii in range(n_slices)])
    dev = _build(line, on_gpu)
    particles = xb.Particles(p0c=6500e9, x=np.zeros(n_part), px=np.linspace(-1, 1, n_part),
                             y=np.zeros(n_part), py=np.linspace(-2, 2, n_part), _device=dev)
    line.track(particles, num_turns=n_turns, turn_by_turn_monitor=True)
    part_id = particles.get('particle_id')
    part_px, part_py = particles.get('px'), particles.get('py')
    part_s, part_at_turn = particles.get('s'), particles.get('at_turn')
    part_at_element = particles.get('at_element')
    s_tot = tot_length * n_turns
    lim = line.config['XTRACK_GLOBAL_XY_LIMIT']
    s_expected = []
    for ii in range(n_part):
        sx = np.abs(lim / part_px[ii]) if np.abs(part_px[ii]) * s_tot > lim else s_tot
        sy = np.abs(lim / part_py[ii]) if np.abs(part_py[ii] * s_tot) > lim else s_tot
        s_expected.append(min(sx, sy))
    s_expected = np.array(s_expected)
    at_turn_expected = np.int_(np.clip(np.floor(s_expected / tot_length), 0, n_turns))
    at_element_expected = np.floor((s_expected - tot_length * at_turn_expected)
                                   / (tot_length / n_slices))
    at_element_expected = np.int_(np.clip(at_element_expected, 0, n_slices - 1))
    np.testing.assert_allclose(part_s + at_turn_expected * line.get_length(), s_expected, atol=1e-3)
    np.testing.assert_allclose(at_turn_expected, part_at_turn)
    np.testing.assert_allclose(at_element_expected, part_at_element, atol=1.1)
    mon = line.record_last_track
    m = {ff: mon.get(ff) for ff in ('at_turn', 's', 'x', 'y', 'px', 'py')}
    for ii in range(n_part):
        iidd = part_id[ii]
        for tt in range(n_turns):
            if tt <= part_at_turn[ii]:
                assert m['at_turn'][iidd, tt] == tt
                assert np.isclose(m['s'][iidd, tt], 0., atol=1e-14)
                assert np.isclose(m['x'][iidd, tt], tt * tot_length * part_px[ii], atol=1e-14)
                assert np.isclose(m['y'][iidd, tt], tt * tot_length * part_py[ii], atol=1e-14)
                assert np.isclose(m['px'][iidd, tt], part_px[ii], atol=1e-14)
                assert np.isclose(m['py'][iidd, tt], part_py[ii], atol=1e-14)
            else:
                for ff in m:
                    assert m[ff][iidd, tt] == 0


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_global_aperture_after_static_thick_elements(on_gpu):
    """tests/test_apertures.py:16-44: the global aperture check follows the statically thick
    classes (Drift, Quadrupole) and not a Multipole that happens to be thick."""
    line = xb.Line(elements=[xb.Drift(length=2), xb.Quadrupole(length=2, k1=0),
                             xb.Multipole(length=2, isthick=True)])
    dev = _build(line, on_gpu)
    for ele_start, expected in ((1, -1), (2, 1), (0, -1)):
        p = xb.Particles(p0c=7e12, px=0.6, _device=dev)
        line.track(p, ele_start=ele_start, num_elements=1)
        assert p.get('state')[0] == expected, (ele_start, p.get('state'))


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_drift_against_closed_form(on_gpu):
    """tests/test_elements.py:497-522 (there against ducktrack's Drift, whose map is the
    closed form below: ducktrack/elements.py Drift.track), 1e-14."""
    kw = dict(p0c=25.92e9, x=1e-3, px=1e-5, y=-2e-3, py=-1.5e-5, delta=1e-2, zeta=1.)
    line = xb.Line(elements=[xb.Drift(length=10.)])
    dev = _build(line, on_gpu)
    p = xb.Particles(_device=dev, **kw)
    p0 = xb.Particles(**kw)
    line.track(p)
    rpp, rvv = p0.get('rpp')[0], p0.get('rvv')[0]
    xp, yp = kw['px'] * rpp, kw['py'] * rpp
    np.testing.assert_allclose(p.get('x')[0], kw['x'] + xp * 10., rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(p.get('y')[0], kw['y'] + yp * 10., rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(p.get('zeta')[0],
                               kw['zeta'] + 10. * (1. - 1. / rvv * (1 + (xp ** 2 + yp ** 2) / 2)),
                               rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_cavity(on_gpu):
    """tests/test_elements.py:823-846: a cavity at the crest adds its voltage to the energy;
    delta, rpp, rvv follow, tau is unchanged."""
    line = xb.Line(elements=[xb.Cavity(frequency=0, phase=np.pi / 2, voltage=30)])
    dev = _build(line, on_gpu)
    part = xb.Particles(p0c=1e9, delta=[0, 1e-2], zeta=[0, 0.2], _device=dev)
    part0 = part.copy(_device='cpu')
    line.track(part)
    part = part.copy(_device='cpu')
    e, e0 = part.energy.numpy(), part0.energy.numpy()
    np.testing.assert_allclose(e, e0 + 30, atol=5e-7, rtol=0)
    Pc = np.sqrt(e ** 2 - part.mass0 ** 2)
    delta = Pc / part.get('p0c') - 1
    beta = Pc / e
    np.testing.assert_allclose(part.get('delta'), delta, atol=1e-14, rtol=0)
    np.testing.assert_allclose(part.get('rpp'), 1 / (1 + delta), atol=1e-14, rtol=0)
    np.testing.assert_allclose(part.get('rvv'), beta / part.get('beta0'), atol=1e-14, rtol=0)
    np.testing.assert_allclose(part.get('zeta') / part.get('beta0'),
                               part0.get('zeta') / part0.get('beta0'), atol=1e-14, rtol=0)
    np.testing.assert_allclose((part.get('ptau') - part0.get('ptau')) * part0.get('p0c'), 30,
                               atol=1e-9, rtol=0)


def test_particles_monitor_semantics():
    """tests/test_monitor.py:71-192 on the hllhc_14 stand-in (the reference's own fixture json
    is a missing blob): implicit monitor, explicit monitor, dict round trip, frames, monitors
    installed in the line."""
    import hostsim
    line0 = common.load_line('hllhc_14')
    num_particles = 50
    particles0 = common.gaussian_particles(line0, num_particles, 5, common.SIGMAS['hllhc_14'])
    line = line0
    hostsim.build_hostsim_tracker(line)

    particles = particles0.copy()
    num_turns = 30
    line.track(particles, num_turns=num_turns, turn_by_turn_monitor=True)
    mon = line.record_last_track
    assert mon.x.shape == (50, 30)
    assert np.all(mon.at_turn[3, :] == np.arange(0, num_turns))
    assert np.all(mon.particle_id[:, 3] == np.arange(0, num_particles))
    assert np.all(mon.at_element[:, :] == 0)
    assert np.all(mon.pzeta[:, 0] == particles0.get('ptau') / particles0.get('beta0'))

    monitor = xb.ParticlesMonitor(start_at_turn=5, stop_at_turn=15, num_particles=num_particles)
    particles = particles0.copy()
    line.track(particles, num_turns=num_turns, turn_by_turn_monitor=monitor)
    assert monitor.x.shape == (50, 10)
    assert np.all(monitor.at_turn[3, :] == np.arange(5, 15))
    assert np.all(monitor.particle_id[:, 3] == np.arange(0, num_particles))
    assert np.all(monitor.at_element[:, :] == 0)
    assert np.all(monitor.pzeta[:, 0] == mon.pzeta[:, 5])

    dct = monitor.to_dict()
    assert 'data' not in dct
    monitor2 = xb.ParticlesMonitor.from_dict(dct)
    assert monitor2.x.shape == (50, 10)
    assert np.all(monitor2.x == 0) and np.all(monitor2.at_turn == 0) and np.all(monitor2.particle_id == 0)
    particles = particles0.copy()
    line.track(particles, num_turns=num_turns, turn_by_turn_monitor=monitor2)
    assert np.all(monitor2.at_turn[3, :] == np.arange(5, 15))
    assert np.all(monitor2.pzeta[:, 0] == mon.pzeta[:, 5])

    multi = xb.ParticlesMonitor(start_at_turn=5, stop_at_turn=10, n_repetitions=3,
                                repetition_period=20, num_particles=num_particles)
    particles = particles0.copy()
    line.track(particles, num_turns=100, turn_by_turn_monitor=multi)
    assert multi.x.shape == (3, 50, 5)
    assert np.all(multi.at_turn[1, 3, :] == np.arange(25, 30))
    assert np.all(multi.particle_id[2, :, 3] == np.arange(0, num_particles))
    assert np.all(multi.at_element[:, :, :] == 0)
    assert np.all(multi.pzeta[0, :, 0] == mon.pzeta[:, 5])

    # monitors installed in the line
    mon_a = xb.ParticlesMonitor(start_at_turn=5, stop_at_turn=15, num_particles=num_particles)
    mon_b = xb.ParticlesMonitor(start_at_turn=5, stop_at_turn=15, num_particles=num_particles)
    els, names = list(line0.elements), list(line0.element_names)
    ia, ib = 1200, 7000
    els.insert(ib, mon_b);  names.insert(ib, 'mymon_b')
    els.insert(ia, mon_a);  names.insert(ia, 'mymon_a')
    line_w = xb.Line(elements=els, element_names=names)
    line_w.particle_ref = line0.particle_ref
    hostsim.build_hostsim_tracker(line_w)
    particles = particles0.copy()
    line_w.track(particles, num_turns=20)
    for mm, nn in ((mon_a, 'mymon_a'), (mon_b, 'mymon_b')):
        assert mm.x.shape == (50, 10)
        assert np.all(mm.at_turn[3, :] == np.arange(5, 15))
        assert np.all(mm.particle_id[:, 3] == np.arange(0, num_particles))
        assert np.all(mm.at_element[:, :] == line_w.element_names.index(nn))


def test_standalone_element_track():
    """`element.track(particles)` as the reference's tests use it (tests/test_elements.py:
    497-522: `drift.track(particles)`; base_element.py:455-480): the element's map alone, no
    turn bookkeeping, no global aperture check, `at_element` untouched unless asked."""
    import hostsim
    kw = dict(p0c=25.92e9, x=[1e-3, 0.5], px=[1e-5, 0.3], y=-2e-3, py=-1.5e-5, delta=1e-2, zeta=1.)
    p = xb.Particles(**kw)
    drift = xb.Drift(length=10.)
    drift.track(p, _tracker_class=hostsim.HostSimTracker)
    rpp = xb.Particles(**kw).get('rpp')
    np.testing.assert_allclose(p.get('x'), np.array(kw['x']) + np.array(kw['px']) * rpp * 10.,
                               rtol=1e-14, atol=1e-14)
    assert np.all(p.get('state') == 1)              # x = 3.5 m: no global aperture check here
    assert np.all(p.get('at_element') == 0) and np.all(p.get('at_turn') == 0)
    assert np.all(p.get('s') == 10.)
    drift.length = 5.
    drift.track(p, increment_at_element=True, _tracker_class=hostsim.HostSimTracker)
    assert np.all(p.get('s') == 15.) and np.all(p.get('at_element') == 1)
    quad = xb.Multipole(knl=[0, 1e-2])
    px0 = p.get('px').copy()
    quad.track(p, _tracker_class=hostsim.HostSimTracker)
    np.testing.assert_allclose(p.get('px'), px0 - 1e-2 * p.get('x'), rtol=1e-14)
