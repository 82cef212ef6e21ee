"""CPU tier: the lattice lowering (xtrack_b200/lowering.py) and the device op
semantics, exercised through the HOST build of the device headers
(tests/hostsim -- test infrastructure, never used by the product) against the
reference-header oracle.  The host build rounds like the EXACT kernel variant and
links the same libm as the oracle, so the expected result is bit-identity.
The real CUDA kernel is covered by tests/test_gpu_parity.py (-m gpu).
"""
import numpy as np
import pytest

import xtrack_b200 as xb
import common
import hostsim


def _track(line, p_host, num_turns, **kw):
    p = p_host.copy()
    hostsim.build_hostsim_tracker(line)
    line.track(p, num_turns=num_turns, **kw)
    return p


def _assert_identical(got, ref, fields=common.ALL_F64 + ('state', 'at_turn', 'at_element')):
    for ff in fields:
        assert np.array_equal(got[ff], ref[ff]), ff


@pytest.mark.parametrize('name', ['hllhc_14', 'sps', 'clic_dr', 'lep'])
def test_ten_turns_bit_identical(name):
    hostsim.trig_stats(reset=True)
    line = common.load_line(name)
    p_host = common.gaussian_particles(line, 60, 11, common.SIGMAS[name])
    ref = common.oracle_track(line, p_host, 10 if name != 'lep' else 4)
    got = common.by_id(_track(line, p_host, 10 if name != 'lep' else 4))
    _assert_identical(got, ref)
    lookups, misses = hostsim.trig_stats()
    assert misses == 0, (lookups, misses)
    if name == "lep":       # 1696 bends x 32 polar drifts per PAIR of particles and turn, all tabulated
        assert lookups > 30 * 4 * 1696 * 30


def test_ducktrack_golden_full_rings():
    """The reference's own pin of this path (tests/test_full_rings.py:24-118): 10 turns of
    the fixture particle vs ducktrack, tolerances of the reference test."""
    import json
    import os
    with open(os.path.join(common.HERE, 'golden', 'ducktrack_full_rings.json')) as fid:
        golden = json.load(fid)
    for name in ('hllhc_14', 'sps'):
        dd = common.load_fixture(name)
        line = xb.Line.from_dict(dd, replace_unsupported=True)
        line.reset_s_at_end_turn = False
        p = xb.Particles.from_dict(dd['particle'])
        got = common.by_id(_track(line, p, 10))
        gg = golden[name]
        for ff in ('x', 'px', 'y', 'py', 'zeta', 'delta', 's'):
            np.testing.assert_allclose(got[ff][0], gg[ff], rtol=gg['rtol'], atol=gg['atol'],
                                       err_msg=f'{name} {ff}')


@pytest.mark.parametrize('thin', [True, False])
def test_toy_ring(thin):
    line = common.toy_ring(thin=thin)
    p_host = common.gaussian_particles(line, 200, 1, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 20)
    _assert_identical(common.by_id(_track(line, p_host, 20)), ref)


def test_losses_sps_apertures():
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 400, 3, common.SIGMAS['sps'], scale=6.0)
    ref = common.oracle_track(line, p_host, 5)
    assert 5 < (ref['state'] <= 0).sum() < 395
    _assert_identical(common.by_id(_track(line, p_host, 5)), ref)


def test_global_limit_and_partial_turns():
    line = common.load_line('hllhc_14')
    line.config['XTRACK_GLOBAL_XY_LIMIT'] = 2e-3
    p_host = common.gaussian_particles(line, 300, 5, common.SIGMAS['hllhc_14'], scale=3.0)
    ref = common.oracle_track(line, p_host, 2)
    assert (ref['state'] == -1).sum() > 3
    _assert_identical(common.by_id(_track(line, p_host, 2)), ref)

    line.config['XTRACK_GLOBAL_XY_LIMIT'] = 1.0
    p_host = common.gaussian_particles(line, 50, 6, common.SIGMAS['hllhc_14'])
    p = p_host.copy()
    hostsim.build_hostsim_tracker(line)
    line.track(p, ele_start=5000, num_elements=len(line) + 300)
    hp = common.ro.HostParticles.from_particles(p_host)
    re = common.ro.RefElements(line.elements)
    kw = dict(flag_reset_s_at_end_turn=1, line_length=line.get_length())
    common.ro.track_line(hp, re, num_turns=1, ele_start=5000, num_ele_track=len(line) - 5000,
                         flag_end_turn_actions=1, **kw)
    common.ro.track_line(hp, re, num_turns=1, ele_start=0, num_ele_track=5300,
                         flag_end_turn_actions=0, **kw)
    _assert_identical(common.by_id(p), hp.sorted_by_id())


def test_lost_particle_does_not_keep_tripping_the_hot_loop():
    """A lane whose particle was lost goes on executing ops from the axis; inside a crossing
    bump that is a large-amplitude orbit which left the global limit within a turn and then
    made xtb_run_fast return at EVERY drift (27 000 returns for 24 lost particles in this
    case; on the GPU a 200-turn launch with 37 losses ran 30 % slower).  Such lanes are put
    back on the axis: about one extra return per excursion."""
    line = common.load_line('hllhc_14')
    p_host = common.gaussian_particles(line, 60, 5, common.SIGMAS['hllhc_14'], scale=12.0)
    ref = common.oracle_track(line, p_host, 6)
    n_lost = int((ref['state'] <= 0).sum())
    assert 10 < n_lost < 50
    hostsim.stop_counts(reset=True)
    got = common.by_id(_track(line, p_host, 6))
    stops = hostsim.stop_counts()
    _assert_identical(got, ref)
    assert stops['global_prefix'] + stops['global_main'] <= 3 * n_lost, stops


def test_turn_by_turn_monitor():
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 40, 8, common.SIGMAS['sps'], scale=5.0)
    mon_ref = common.ro.HostMonitor(0, 4, 0, 40)
    common.oracle_track(line, p_host, 4, monitor=mon_ref, flag_monitor=1)
    p = p_host.copy()
    hostsim.build_hostsim_tracker(line)
    line.track(p, num_turns=4, turn_by_turn_monitor=True)
    mon = line.record_last_track
    assert mon.x.shape == (40, 4)
    for ff, _ in xb.Particles.per_particle_vars:
        assert np.array_equal(mon.get(ff), mon_ref.field(ff)), ff


def test_misaligned_elements():
    """Shifts / tilts on thin and thick, straight and curved elements
    (track_misalignments.h through the transformations wrapper)."""
    mis = dict(shift_x=1e-3, shift_y=-2e-3, shift_s=5e-3, rot_s_rad=0.02, rot_x_rad=1e-3,
               rot_y_rad=-2e-3, rot_s_rad_no_frame=0.01, rot_shift_anchor=0.2)
    els = [xb.Multipole(knl=[0, 0.1, 2.0], ksl=[0, 0.05], **mis),
           xb.Drift(length=1.0),
           xb.Quadrupole(length=0.5, k1=0.3, **mis),
           xb.LimitRect(min_x=-0.05, max_x=0.05, min_y=-0.05, max_y=0.05, shift_x=0.01),
           xb.Bend(length=1.5, angle=0.1, k0='from_h', edge_entry_angle=0.02, edge_exit_angle=0.03,
                   edge_entry_fint=0.5, edge_entry_hgap=0.02, **mis),
           xb.Sextupole(length=0.3, k2=5., rot_s_rad=0.3),
           xb.Cavity(voltage=1e5, frequency=4e8, lag=30., shift_x=2e-3, rot_s_rad=0.1),
           xb.LimitEllipse(a=0.05, b=0.03, rot_s_rad=0.2, shift_y=1e-3),
           xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='straight-body', **mis),
           xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4, shift_x=1e-3)]
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=2e9)
    p_host = common.gaussian_particles(line, 100, 2, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    _assert_identical(common.by_id(_track(line, p_host, 1)), ref)


@pytest.mark.parametrize('model', ['adaptive', 'full', 'bend-kick-bend', 'rot-kick-rot',
                                   'mat-kick-mat', 'drift-kick-drift-exact',
                                   'drift-kick-drift-expanded', 'rot-kick-rot-low-order',
                                   'rot-kick-rot-high-order'])
@pytest.mark.parametrize('integrator', ['adaptive', 'teapot', 'yoshida4', 'uniform'])
def test_bend_models_and_integrators(model, integrator):
    """All body models x integrators of track_magnet.h on a combined-function bend with
    full edges, and on a straight quadrupole-like magnet where the model exists."""
    els = [xb.Bend(length=2.0, angle=0.15, k0=0.08, k1=0.02, k2=0.5, knl=[0, 0, 0.1, 2.0],
                   ksl=[0, 1e-3], model=model, integrator=integrator, num_multipole_kicks=5,
                   edge_entry_model='full', edge_exit_model='full', edge_entry_angle=0.03,
                   edge_exit_angle=0.04, edge_entry_fint=0.5, edge_exit_fint=0.4,
                   edge_entry_hgap=0.02, edge_exit_hgap=0.02),
           xb.Bend(length=1.0, angle=0.0, k0=0.0, k1=0.1, model=model, integrator=integrator,
                   num_multipole_kicks=3, edge_entry_model='dipole-only',
                   edge_exit_model='suppressed')]
    if model not in ('bend-kick-bend', 'rot-kick-rot'):
        els.append(xb.Quadrupole(length=0.7, k1=0.2, k1s=0.01, model=model, integrator=integrator,
                                 num_multipole_kicks=4, edge_entry_active=1, edge_exit_active=1))
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=3e9)
    p_host = common.gaussian_particles(line, 50, 4, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    hostsim.trig_stats(reset=True)
    _assert_identical(common.by_id(_track(line, p_host, 1)), ref)
    # every sin/cos of element constants came from the table the lowering computed
    lookups, misses = hostsim.trig_stats()
    assert misses == 0, (lookups, misses)
    if model in ('rot-kick-rot', 'bend-kick-bend', 'full', 'adaptive', 'rot-kick-rot-low-order',
                 'rot-kick-rot-high-order'):
        assert lookups > 0


def test_rbend_models():
    els = []
    for rm in ('adaptive', 'curved-body', 'straight-body'):
        for diff in (0.0, 0.01):
            els.append(xb.RBend(length_straight=1.5, angle=0.12, k0='from_h', rbend_model=rm,
                                rbend_angle_diff=diff, k1=0.05, edge_entry_model='full',
                                edge_exit_model='linear', edge_entry_fint=0.5, edge_entry_hgap=0.02,
                                rbend_shift=1e-3))
            els.append(xb.Drift(length=0.2))
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=3e9)
    p_host = common.gaussian_particles(line, 50, 4, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    _assert_identical(common.by_id(_track(line, p_host, 1)), ref)


def test_polygon_drift_exact_rfmultipole_thick_multipole():
    els = [xb.LimitPolygon(x_vertices=[-0.03, 0.03, 0.04, 0.0, -0.04],
                           y_vertices=[-0.02, -0.02, 0.02, 0.035, 0.02]),
           xb.DriftExact(length=2.0), xb.Drift(length=1.0, model='exact'),
           xb.RFMultipole(voltage=1e4, frequency=4e8, lag=10., knl=[1e-3, 1e-2], ksl=[0, 2e-2],
                          pn=[10., 20.], ps=[0., 30.]),
           xb.Multipole(knl=[0.01, 0.2, 1.0], hxl=0.01, length=0.4, isthick=True,
                        num_multipole_kicks=3),
           xb.SRotation(angle=20.), xb.XYShift(dx=1e-3, dy=-1e-3),
           xb.Octupole(length=0.3, k3=100., k3s=20.)]
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=3e9)
    p_host = common.gaussian_particles(line, 300, 4, common.SIGMAS['toy'], scale=20.)
    ref = common.oracle_track(line, p_host, 1)
    assert 10 < (ref['state'] == 0).sum() < 290
    _assert_identical(common.by_id(_track(line, p_host, 1)), ref)


def test_freeze_longitudinal():
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 30, 9, common.SIGMAS['sps'])
    p = _track(line, p_host, 2, freeze_longitudinal=True)
    for ff in ('zeta', 'delta', 'ptau', 'rpp', 'rvv', 's'):
        assert np.array_equal(p.get(ff), p_host.get(ff)), ff
    assert np.all(p.get('at_turn') == 2)
    assert not np.array_equal(p.get('x'), p_host.get('x'))


def test_nonuniform_s_and_mixed_species():
    """The hot loop's fast state (chi == 1, one common s carried once per thread) is a
    per-block decision taken at launch; beams that do not qualify -- particles sitting at
    different s, ions of several charge states -- take the general code path.  Both must
    reproduce the reference, losses included (a lost particle's s is frozen)."""
    line = common.load_line('sps')
    n = 240
    rng = np.random.default_rng(5)
    for case in ('uniform', 's', 'chi'):
        extra = {}
        if case == 's':
            extra['s'] = rng.uniform(0, 3, n)
        if case == 'chi':
            extra['chi'] = np.where(np.arange(n) % 7 == 0, 1.01, 1.0)
            extra['charge_ratio'] = np.ones(n)
        p_host = common.gaussian_particles(line, n, 3, common.SIGMAS['sps'], scale=6.0, **extra)
        ref = common.oracle_track(line, p_host, 6)
        assert 5 < (ref['state'] <= 0).sum() < n
        got = common.by_id(_track(line, p_host, 6))
        _assert_identical(got, ref)


def test_zero_coefficient_specialisations():
    """Ops that leave out the reference's operations on literal-zero coefficients
    (csrc/xtb_ops.h: MULTP1 / MULTPN / MULTH0N; trimmed RF-multipole orders, zero-voltage
    RF elements) against the reference, which computes them all.  np.array_equal treats the
    only admissible difference, the sign of an exact zero, as equal."""
    D = xb.Drift
    els = [D(length=0.7), xb.Multipole(knl=[0, 0.3]),                       # MULTP1 (+prefix)
           xb.Multipole(knl=[0, 0, 2.0]),                                   # MULTPN order 2, no prefix
           D(length=0.4), xb.Multipole(knl=[0, 0, 0, 30.0, 0, 0]),          # order 3, trailing zeros
           D(length=0.4), xb.Multipole(knl=[0, 0, 0, 0, 4e3]),              # order 4
           D(length=0.3), xb.Multipole(knl=[0, 0.2], ksl=[0, 0.1]),         # general order 1
           D(length=0.3), xb.Multipole(ksl=[0, 0.1]),                       # skew only: general
           D(length=0.3), xb.Multipole(knl=[0.0], ksl=[0.0]),               # all zero
           D(length=0.5), xb.Multipole(knl=[0.02], hxl=0.02, length=0.5),   # MULTH0N
           D(length=0.5), xb.Multipole(knl=[0.02], ksl=[1e-3], hxl=0.02, length=0.5),   # MULTH0
           D(length=0.2), xb.RFMultipole(voltage=0., frequency=4e8, knl=[0, 0, 0, 0, 0, 0],
                                         ksl=[0, 0, 0, 0, 0, 0]),           # no strength at all
           D(length=0.2), xb.RFMultipole(voltage=0., frequency=4e8, knl=[0, 0, 0, 0, 0, 0],
                                         ksl=[-4.8e-7, 0, 0, 0, 0, 0]),     # crab placeholder
           D(length=0.2), xb.RFMultipole(voltage=0., frequency=4e8, knl=[4.8e-7, 0, 0, 0, 0, 0],
                                         ksl=[0, 0, 0, 0, 0, 0], pn=[90., 0, 0, 0, 0, 0]),
           D(length=0.2), xb.RFMultipole(voltage=2e3, frequency=4e8, knl=[0, 0, 1e-1, 0],
                                         ksl=[1e-4, 0, 0, 0], pn=[0, 0, 20., 0], ps=[5., 0, 0, 0]),
           D(length=0.2), xb.Cavity(voltage=0., frequency=4e8, lag=30.),
           D(length=0.2), xb.Cavity(voltage=1e5, frequency=4e8, lag=30.)]
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=3e9)
    p_host = common.gaussian_particles(line, 300, 4, common.SIGMAS['toy'], scale=5.)
    # particles exactly on axis / on one axis: the zero-sign cases
    x = p_host.get('x').copy();  y = p_host.get('y').copy()
    px = p_host.get('px').copy();  py = p_host.get('py').copy()
    x[:10] = 0.;  y[5:20] = 0.;  px[:3] = 0.;  py[:8] = 0.
    p_host.x = x;  p_host.y = y;  p_host.px = px;  p_host.py = py
    ref = common.oracle_track(line, p_host, 3)
    got = common.by_id(_track(line, p_host, 3))
    _assert_identical(got, ref)


def test_oracle_openmp_builds_equal_serial():
    """The OpenMP builds of the oracle (used by the GPU tests and as the CPU baseline of
    bench.py) give the serial build's results particle by particle: bit-identical for
    survivors, identical loss records for lost ones (whose coordinates may stay in the frame
    of a misaligned aperture in the reference's OpenMP context, see common.oracle_track)."""
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 500, 3, common.SIGMAS['sps'], scale=6.0)
    for variant in ('serial', 'noise'):
        a = common.oracle_track(line, p_host, 4, variant=variant)
        b = common.oracle_track(line, p_host, 4, variant=common._OMP_VARIANT[variant])
        assert 5 < (a['state'] <= 0).sum() < 495
        alive = a['state'] > 0
        for ff in common.ALL_F64:
            assert np.array_equal(a[ff][alive], b[ff][alive]), ff
        for ff in ('state', 'at_turn', 'at_element'):
            assert np.array_equal(a[ff], b[ff]), ff
        # with losses, parallel=True falls back to the serial build
        _assert_identical(common.oracle_track(line, p_host, 4, variant=variant, parallel=True), a)


def test_rotation_and_translation_elements():
    """`Rotation` / `Translation` (they supersede the deprecated
    SRotation / XYShift; elements_src/rotation.h:13-60, translation.h:13-26), every rotation
    order, zero angles skipped."""
    els = []
    for seq in ('yxs', 'xys', 'sxy', 'syx'):
        els += [xb.Drift(length=0.5),
                xb.Rotation(rot_s_rad=0.02, rot_x_rad=-3e-3, rot_y_rad=2e-3, seq=seq),
                xb.Translation(shift_x=1e-4, shift_y=-2e-4)]
    els += [xb.Rotation(rot_s_rad=0.3), xb.Rotation(rot_y_rad=1e-3), xb.Rotation(),
            xb.Multipole(knl=[0, 0.1])]
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=3e9)
    p_host = common.gaussian_particles(line, 100, 4, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 2)
    _assert_identical(common.by_id(_track(line, p_host, 2)), ref)
    # round trip through the dictionary form
    line2 = xb.Line.from_dict(line.to_dict()) if hasattr(line, 'to_dict') else line
    line2.particle_ref = line.particle_ref
    _assert_identical(common.by_id(_track(line2, p_host, 2)), ref)


@pytest.mark.parametrize('name', ['hllhc_14', 'sps'])
def test_optimize_for_tracking(name):
    """`Line.optimize_for_tracking` (line.py:4951-5026): markers, inactive multipoles and
    zero-length drifts removed, consecutive drifts / multipoles merged, redundant apertures
    dropped.  The optimised line is tracked bit-identically to the reference tracking the SAME
    optimised element list, and agrees with the unoptimised line to rounding (merged drift
    lengths round once instead of twice)."""
    line = common.load_line(name)
    n0 = len(line)
    p_host = common.gaussian_particles(line, 40, 11, common.SIGMAS[name])
    ref_plain = common.oracle_track(line, p_host, 5)
    length0 = line.get_length()
    line.optimize_for_tracking()
    n1 = len(line)
    assert n1 < n0
    cls = {type(ee).__name__ for ee in line.elements}
    assert 'Marker' not in cls
    names = [type(ee).__name__ for ee in line.elements]
    assert not any(a == b == 'Drift' for a, b in zip(names, names[1:]))
    assert abs(line.get_length() - length0) < 1e-9 * length0
    ref_opt = common.oracle_track(line, p_host, 5)
    got = common.by_id(_track(line, p_host, 5))
    _assert_identical(got, ref_opt, fields=common.ALL_F64 + ('state', 'at_turn'))
    dev = common.max_rel_dev(got, ref_plain)
    assert max(dev.values()) < 1e-8, dev
    print(name, n0, '->', n1, 'elements; dev vs unoptimised', max(dev.values()))


def test_losses_in_thick_lattice_bit_identical():
    """LEP thick lattice + apertures, wide beam: particles are lost all along the ring while
    their thread neighbours go on (thick run loop, two particles per thread): loss records
    and every coordinate identical to the reference, odd particle count included."""
    line = common.lep_with_apertures()
    p_host = common.gaussian_particles(line, 61, 17, common.SIGMAS['lep'], scale=4.0)
    ref = common.oracle_track(line, p_host, 3)
    n_lost = int((ref['state'] <= 0).sum())
    assert 8 < n_lost < 55, n_lost
    got = common.by_id(_track(line, p_host, 3))
    _assert_identical(got, ref)



@pytest.mark.parametrize('ele_start,ele_stop', [(0, None), (5, 3), (2, 9)])
def test_with_progress_batches_equal_one_call(ele_start, ele_stop):
    """`Line.track(..., with_progress=N)` (tracker.py:313-381): batches of N turns -- partial
    first / last turns, one monitor over all batches -- give the result of the single call."""
    line = common.toy_ring(thin=True)
    p_host = common.gaussian_particles(line, 40, 21, common.SIGMAS['toy'])
    hostsim.build_hostsim_tracker(line)
    pa, pb = p_host.copy(), p_host.copy()
    line.track(pa, num_turns=11, ele_start=ele_start, ele_stop=ele_stop, turn_by_turn_monitor=True)
    mon_a = line.record_last_track
    line.track(pb, num_turns=11, ele_start=ele_start, ele_stop=ele_stop, turn_by_turn_monitor=True,
               with_progress=4)
    mon_b = line.record_last_track
    ga, gb = common.by_id(pa), common.by_id(pb)
    for ff in common.ALL_F64 + ('state', 'at_turn', 'at_element'):
        assert np.array_equal(ga[ff], gb[ff]), ff
    for ff in ('x', 'px', 'zeta', 'at_turn', 'at_element'):
        assert np.array_equal(mon_a.get(ff), mon_b.get(ff)), ff
    with pytest.raises(ValueError):
        line.track(pb, with_progress=True)


def test_host_build_of_device_sin_cos_gives_libm_bits():
    """csrc/xtb_libm.cuh (glibc's sin / cos algorithm restated, FMA placement of the library's
    x86-64 FMA build) against the installed libm: scripts/glibc/check_libm.c, 2 x 10^6 arguments
    per branch here (10^9 in the session log, profiles/r02_history.md)."""
    import os
    import subprocess
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
    if ' fma ' not in open('/proc/cpuinfo').read():
        pytest.skip('host without FMA: its libm runs the non-FMA build of sin / cos')
    exe = os.path.join(root, 'tests', 'hostsim', '_build', 'check_libm')
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(['g++', '-O2', '-mfma', '-ffp-contract=off', '-fopenmp', '-x', 'c++',
                    os.path.join(root, 'scripts', 'glibc', 'check_libm.c'), '-o', exe, '-lm'],
                   check=True)
    out = subprocess.run([exe, '2000000'], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    # the table is what the generator writes (mathematical values + the library's 18 deviations)
    gen = subprocess.run(['python', os.path.join(root, 'scripts', 'glibc', 'gen_sincostab.py'),
                          '--check'], capture_output=True, text=True)
    assert gen.returncode == 0, gen.stdout + gen.stderr


def test_thick_fixture_derivation_physics_pins():
    """The oracle's element structs for the thick rings are filled from this repository's own
    reading of the JSON (RBend h / length from `length_straight` and `angle`, `k0 = 'from_h'`,
    enumerations by name ...): a misreading would move oracle and product together.  These pins
    do not depend on that reading being right twice -- they are physics the ring must satisfy:
      * the bending angles of a ring add up to 2 pi (LEP, and the reference's lattice-design ring);
      * with k0 == h in every bend, the on-momentum particle started on the design orbit stays
        on it (any mis-derived strength, length or edge angle kicks it off by millimetres);
      * the curved length of an RBend is angle / h with h = 2 sin(angle / 2) / length_straight;
      * a ring cut into thick slices tracks like the unsliced ring (integrator accuracy)."""
    import math
    for name in ('lep', 'ring'):
        line = common.load_line(name)
        bends = [ee for ee in line.elements if type(ee).__name__ in ('Bend', 'RBend')]
        assert abs(sum(ee.angle for ee in bends) / (2 * math.pi) - 1) < 1e-8, name
        for ee in bends:
            assert ee.k0 == ee.h or name == 'ring'
            if type(ee).__name__ == 'RBend' and ee.angle != 0:
                assert abs(ee.h - 2 * math.sin(ee.angle / 2) / ee.length_straight) < 1e-15
                assert abs(ee.length * ee.h - ee.angle) < 1e-15
        ref = line.particle_ref
        p = xb.Particles(p0c=float(ref.get('p0c')[0]), mass0=ref.mass0, q0=ref.q0,
                         x=[0., 1e-4], px=[0., 0.], y=[0., 5e-5], delta=[0., 0.])
        got = common.by_id(_track(line, p, 1))
        assert abs(got['x'][0]) < 2e-7 and abs(got['y'][0]) < 1e-9 and abs(got['px'][0]) < 2e-8, \
            (name, got['x'][0], got['px'][0])
        assert abs(got['x'][1]) < 5e-3          # a betatron oscillation, not an escape
    whole = common.load_line('ring')
    sliced = common.load_line('ring_sliced')
    p_host = common.gaussian_particles(whole, 20, 1, common.SIGMAS['toy'], scale=0.3)
    a, b = common.by_id(_track(whole, p_host, 2)), common.by_id(_track(sliced, p_host, 2))
    dev = common.max_rel_dev(a, b, fields=('x', 'px', 'y', 'py'))
    assert max(dev.values()) < 1e-6, dev
