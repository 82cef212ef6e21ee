"""SURVEY.md §8(a) rows W, A, D, M-edge, T, Mon, F and §8(f1), each against the reference
oracle on BOTH tiers: the host build of the device code (CPU tier, expected: bit identity --
same libm as the oracle) and, marked `gpu`, the CUDA kernel through the C-ABI.

On the B200 the EXACT kernel variant reproduces the reference bit for bit wherever the
element evaluates no libm function on the device (`bitwise=True` below: +, *, /, sqrt with
the element trigonometry folded by the host); the other cases (per-particle atan / asin /
tan / sin / cos of the fringe, wedge and RF maps) are held to 1e-13 of the beam size after
one passage -- a last-bit difference of the two libms, not amplified by any optics yet.
"""
import numpy as np
import pytest

import xtrack_b200 as xb
import common
import ref_oracle as ro

BACKENDS = [pytest.param(False, id='hostsim'), pytest.param(True, id='gpu', marks=pytest.mark.gpu)]
FIELDS = common.ALL_F64
INTS = ('state', 'at_turn', 'at_element')


def _build(line, on_gpu, **kw):
    if on_gpu:
        line.build_tracker(_device='cuda:0', exact_arithmetic=True, **kw)
        return 'cuda:0'
    import hostsim
    line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker, **kw)
    return 'cpu'


def _track(line, p_host, on_gpu, num_turns=1, **kw):
    dev = _build(line, on_gpu)
    p = p_host.copy(_device=dev)
    line.track(p, num_turns=num_turns, **kw)
    return common.by_id(p)


def _compare(got, ref, bitwise, label='', rtol=1e-13, fields=FIELDS, ints=INTS):
    for ff in ints:
        assert np.array_equal(got[ff], ref[ff]), (label, ff)
    worst = 0.0
    for ff in fields:
        if bitwise:
            assert np.array_equal(got[ff], ref[ff]), (label, ff, np.max(np.abs(got[ff] - ref[ff])))
        else:
            ok = np.isfinite(ref[ff])
            scale = max(float(np.max(np.abs(ref[ff][ok]))) if ok.any() else 0.0, 1e-300)
            dev = float(np.max(np.abs(got[ff][ok] - ref[ff][ok]))) / scale
            worst = max(worst, dev)
            assert dev <= rtol, (label, ff, dev)
    return worst


def _line(els, p0c=2e9, **kw):
    line = xb.Line(elements=els)
    line.particle_ref = xb.Particles(p0c=p0c, **kw)
    return line


MIS = dict(shift_x=1e-3, shift_y=-2e-3, shift_s=5e-3, rot_s_rad=0.02, rot_x_rad=1e-3,
           rot_y_rad=-2e-3, rot_s_rad_no_frame=0.01, rot_shift_anchor=0.2)


# ---- W: misalignment wrapper -------------------------------------------------------------
@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_misaligned_straight_and_curved_elements(on_gpu):
    """Shifts, tilts, x / y rotations and anchors on thin and thick, straight and curved
    elements (track_misalignments.h:38-113 straight, :116-378 rigid-matrix algebra of a curved
    parent; wrapper track_local_particle_with_transformations.h:99-204).  Every element here
    is libm-free on the device: bit identity on the B200 too."""
    els = [xb.Multipole(knl=[0, 0.1, 2.0], ksl=[0, 0.05], **MIS),
           xb.Drift(length=1.0),
           xb.LimitRect(min_x=-0.05, max_x=0.05, min_y=-0.05, max_y=0.05, shift_x=0.01,
                        rot_s_rad=0.1),
           xb.Bend(length=1.5, angle=0.1, k0='from_h', edge_entry_angle=0.02, edge_exit_angle=0.03,
                   edge_entry_fint=0.5, edge_entry_hgap=0.02, **MIS),           # curved parent
           xb.Bend(length=1.0, angle=0.05, k0='from_h', rot_x_rad=2e-3, rot_y_rad=1e-3),
           xb.Sextupole(length=0.3, k2=5., rot_s_rad=0.3, shift_y=1e-3),
           xb.LimitEllipse(a=0.05, b=0.03, rot_s_rad=0.2, shift_y=1e-3),
           xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='straight-body', **MIS),
           xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='curved-body', **MIS),
           xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4, shift_x=1e-3, rot_s_rad=0.05),
           xb.LimitPolygon(x_vertices=[-0.04, 0.04, 0.05, 0.0, -0.05],
                           y_vertices=[-0.03, -0.03, 0.03, 0.045, 0.03], shift_x=2e-3, rot_s_rad=0.3),
           xb.Octupole(length=0.2, k3=50., shift_x=-1e-3, rot_shift_anchor=0.1, rot_y_rad=1e-3)]
    line = _line(els)
    p_host = common.gaussian_particles(line, 333, 2, common.SIGMAS['toy'], scale=8.)
    ref = common.oracle_track(line, p_host, 1)
    assert 5 < (ref['state'] <= 0).sum() < 300
    _compare(_track(line, p_host, on_gpu), ref, True, 'misaligned')


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_misaligned_elements_with_device_libm(on_gpu):
    """Misaligned elements whose map calls libm per particle (quadrupole matrix: sin / cos /
    sinh / cosh; cavity: sin)."""
    els = [xb.Quadrupole(length=0.5, k1=0.3, **MIS), xb.Drift(length=0.5),
           xb.Quadrupole(length=0.5, k1=-0.3, k1s=0.01, rot_s_rad=0.7853981633974483),
           xb.Cavity(voltage=1e5, frequency=4e8, lag=30., shift_x=2e-3, rot_s_rad=0.1),
           xb.Cavity(length=0.4, voltage=2e5, frequency=4e8, lag=150., **MIS)]
    line = _line(els)
    p_host = common.gaussian_particles(line, 300, 2, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    _compare(_track(line, p_host, on_gpu), ref, not on_gpu, 'misaligned libm')


# ---- A, D, T: polygon, exact drifts, thick multipole, frame elements -------------------------
@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_polygon_exact_drifts_thick_multipole_frame_elements(on_gpu):
    """LimitPolygon (limitpolygon.h:63-88), DriftExact / Drift(model='exact')
    (track_drift.h:26-40), a thick Multipole (multipole.h:33 `isthick`), SRotation, XYShift,
    Rotation / Translation in every order (rotation.h:13-60, translation.h:13-26)."""
    els = [xb.LimitPolygon(x_vertices=[-0.03, 0.03, 0.04, 0.0, -0.04],
                           y_vertices=[-0.02, -0.02, 0.02, 0.035, 0.02]),
           xb.DriftExact(length=2.0), xb.Drift(length=1.0, model='exact'),
           xb.Multipole(knl=[0.01, 0.2, 1.0], hxl=0.01, length=0.4, isthick=True,
                        num_multipole_kicks=3),
           xb.SRotation(angle=20.), xb.XYShift(dx=1e-3, dy=-1e-3),
           xb.Octupole(length=0.3, k3=100., k3s=20.)]
    for seq in ('yxs', 'xys', 'sxy', 'syx'):
        els += [xb.Drift(length=0.5),
                xb.Rotation(rot_s_rad=0.02, rot_x_rad=-3e-3, rot_y_rad=2e-3, seq=seq),
                xb.Translation(shift_x=1e-4, shift_y=-2e-4)]
    els += [xb.Rotation(rot_s_rad=0.3), xb.Rotation(rot_y_rad=1e-3), xb.Rotation(),
            xb.Multipole(knl=[0, 0.1])]
    line = _line(els, p0c=3e9)
    p_host = common.gaussian_particles(line, 301, 4, common.SIGMAS['toy'], scale=20.)
    ref = common.oracle_track(line, p_host, 2)
    assert 10 < (ref['state'] == 0).sum() < 290
    _compare(_track(line, p_host, on_gpu, num_turns=2), ref, True, 'polygon/exact/thick mult')


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_exact_drifts_line_config(on_gpu):
    """`line.config.XTRACK_USE_EXACT_DRIFTS` (drift.h:22): every Drift takes the exact map."""
    line = common.load_line('sps')
    line.config['XTRACK_USE_EXACT_DRIFTS'] = True
    els = [ee for ee in line.elements if type(ee).__name__ != 'Cavity'][:2000]   # libm-free stretch
    line2 = _line(els, p0c=float(line.particle_ref.get('p0c')[0]))
    line2.config['XTRACK_USE_EXACT_DRIFTS'] = True
    p_host = common.gaussian_particles(line2, 200, 3, common.SIGMAS['sps'], scale=3.0)
    # the oracle has no such switch: hand it the same line with DriftExact elements
    els_ref = [xb.DriftExact(length=ee.length) if type(ee).__name__ == 'Drift' else ee for ee in els]
    ref = common.oracle_track(_line(els_ref, p0c=float(line.particle_ref.get('p0c')[0])), p_host, 1)
    plain = common.oracle_track(_line(els, p0c=float(line.particle_ref.get('p0c')[0])), p_host, 1)
    assert not np.array_equal(ref['zeta'], plain['zeta'])
    _compare(_track(line2, p_host, on_gpu), ref, True, 'exact drifts')


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_rf_multipole(on_gpu):
    els = [xb.Drift(length=0.3),
           xb.RFMultipole(voltage=1e4, frequency=4e8, lag=10., knl=[1e-3, 1e-2], ksl=[0, 2e-2],
                          pn=[10., 20.], ps=[0., 30.]),
           xb.RFMultipole(voltage=2e3, frequency=4e8, knl=[0, 0, 1e-1, 0], ksl=[1e-4, 0, 0, 0],
                          pn=[0, 0, 20., 0], ps=[5., 0, 0, 0], shift_x=1e-3)]
    line = _line(els, p0c=3e9)
    p_host = common.gaussian_particles(line, 300, 4, common.SIGMAS['toy'], scale=3.)
    ref = common.oracle_track(line, p_host, 1)
    _compare(_track(line, p_host, on_gpu), ref, not on_gpu, 'rfmultipole')


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_crab_cavity(on_gpu):
    """CrabCavity (crab_cavity.h -> track_rf.h:116-156): thin and thick
    (drift-kick-drift with several kicks), misaligned, zero voltage, and sliced."""
    import slicing_helper as sh
    els = [xb.Drift(length=0.3),
           xb.CrabCavity(crab_voltage=3e6, frequency=4e8, phase=0.3),
           xb.CrabCavity(length=0.6, crab_voltage=-2e6, frequency=4e8, lag=20., num_kicks=3,
                         integrator='teapot', shift_x=1e-3, rot_s_rad=0.4),
           xb.CrabCavity(length=0.4, crab_voltage=0., frequency=4e8),
           xb.CrabCavity(length=0.5, crab_voltage=1e6, frequency=8e8, model='drift-kick-drift-exact',
                         integrator='yoshida4', num_kicks=2)]
    line = _line(els, p0c=7e12 / 100)
    p_host = common.gaussian_particles(line, 200, 4, common.SIGMAS['toy'], scale=3.)
    ref = common.oracle_track(line, p_host, 2)
    assert not np.array_equal(ref['px'], p_host.get('px'))
    _compare(_track(line, p_host, on_gpu, num_turns=2), ref, True, 'crab cavity')
    for mode in ('thin', 'thick'):
        sl = sh.slice_line(line, n=3, mode=mode, only=('CrabCavity',))
        assert any('SliceCrabCavity' in type(ee).__name__ for ee in sl.elements)
        ref = common.oracle_track(sl, p_host, 1)
        _compare(_track(sl, p_host, on_gpu), ref, True, 'crab cavity slices ' + mode)


# ---- M-edge: full / dipole-only edges, fringes, wedge ------------------------------------------
@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('model', ['rot-kick-rot', 'bend-kick-bend', 'mat-kick-mat',
                                   'drift-kick-drift-exact'])
def test_full_and_dipole_only_edges(model, on_gpu):
    """track_magnet_edge.h:17-167 models 1 (full: dipole fringe, multipole fringe, wedge, quad
    wedge) and 2 (dipole-only), -1 (suppressed); quadrupole fringes; the DipoleEdge element's
    full model (track_dipole_edge_nonlinear.h:12-44)."""
    els = [xb.Bend(length=2.0, angle=0.15, k0=0.08, k1=0.02, k2=0.5, knl=[0, 0, 0.1, 2.0],
                   ksl=[0, 1e-3], model=model, num_multipole_kicks=5,
                   edge_entry_model='full', edge_exit_model='full', edge_entry_angle=0.03,
                   edge_exit_angle=0.04, edge_entry_fint=0.5, edge_exit_fint=0.4,
                   edge_entry_hgap=0.02, edge_exit_hgap=0.02),
           xb.Bend(length=1.0, angle=0.0, k0=0.0, k1=0.1, model=model, num_multipole_kicks=3,
                   edge_entry_model='dipole-only', edge_exit_model='suppressed'),
           xb.Bend(length=1.0, angle=0.04, k0='from_h', model=model,
                   edge_entry_model='dipole-only', edge_exit_model='dipole-only',
                   edge_entry_angle=0.02, edge_exit_angle=0.02, edge_entry_fint=0.5,
                   edge_exit_fint=0.5, edge_entry_hgap=0.02, edge_exit_hgap=0.02),
           xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4, model='full', side='entry'),
           xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4, model='full', side='exit'),
           xb.DipoleEdge(k=0.05, e1=0.0, hgap=0.02, fint=0.4, model='full', side='entry')]
    if model not in ('bend-kick-bend', 'rot-kick-rot'):
        els.append(xb.Quadrupole(length=0.7, k1=0.2, k1s=0.01, model=model, num_multipole_kicks=4,
                                 edge_entry_active=1, edge_exit_active=1))
    line = _line(els, p0c=3e9)
    p_host = common.gaussian_particles(line, 201, 4, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    _compare(_track(line, p_host, on_gpu), ref, not on_gpu, 'edges ' + model)


# ---- f1: optimize_for_tracking -----------------------------------------------------------------
@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_optimize_for_tracking_line(on_gpu):
    """`Line.optimize_for_tracking` (line.py:4951-5026) on the SPS stand-in: the optimised line
    is tracked like the reference tracks the same optimised element list."""
    line = common.load_line('sps')
    n0 = len(line)
    line.optimize_for_tracking()
    assert len(line) < n0
    p_host = common.gaussian_particles(line, 300, 11, common.SIGMAS['sps'], scale=5.0)
    ref = common.oracle_track(line, p_host, 5)
    assert 3 < (ref['state'] <= 0).sum() < 297
    got = _track(line, p_host, on_gpu, num_turns=5)
    if not on_gpu:
        _compare(got, ref, True, 'optimised sps')
    else:
        yard = common.libm_yardstick(line, p_host, 5, ref=ref)
        for ff in INTS:
            assert np.array_equal(got[ff], ref[ff]), ff
        common.assert_parity(got, ref, yard, True, mask=ref['state'] > 0, label='optimised sps')


# ---- Mon: element-by-element and multi-frame monitors ----------------------------------------
@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_one_turn_ebe_monitor(on_gpu):
    """`turn_by_turn_monitor='ONE_TURN_EBE'` (tracker.py:1462-1467; particles_monitor.h:29-31:
    the record index is at_element): one record per element + the end of the line, rows of
    lost particles stay zero behind the loss (flag_monitor 2, plain program)."""
    line = common.load_line('sps')
    els = list(line.elements)[:1500]
    line2 = _line(els, p0c=float(line.particle_ref.get('p0c')[0]))
    n = 60
    p_host = common.gaussian_particles(line2, n, 8, common.SIGMAS['sps'], scale=7.0)
    mon_ref = ro.HostMonitor(0, len(els) + 1, 0, n, ebe_mode=1)
    ref = common.oracle_track(line2, p_host, 1, monitor=mon_ref, flag_monitor=2)
    assert 3 < (ref['state'] <= 0).sum() < n - 3
    dev = _build(line2, on_gpu)
    p = p_host.copy(_device=dev)
    line2.track(p, turn_by_turn_monitor='ONE_TURN_EBE')
    mon = line2.record_last_track
    assert mon.x.shape == (n, len(els) + 1)
    for ff, _ in xb.Particles.per_particle_vars:
        assert np.array_equal(mon.get(ff), mon_ref.field(ff)), ff
    _compare(common.by_id(p), ref, True, 'ebe')


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_multi_frame_monitor_in_line(on_gpu):
    """A ParticlesMonitor element in the line with `n_repetitions` frames
    (particles_monitor.h:49-71), an id range, and one turn-by-turn monitor on top."""
    line = common.toy_ring(thin=True)
    els = list(line.elements)
    kw = dict(start_at_turn=2, stop_at_turn=5, n_repetitions=3, repetition_period=6,
              particle_id_range=(5, 40))
    mon = xb.ParticlesMonitor(**kw)
    els.insert(7, mon)
    line2 = _line(els, p0c=1.2e9)
    n = 50
    p_host = common.gaussian_particles(line2, n, 8, common.SIGMAS['toy'])
    mon_ref = ro.HostMonitor(2, 5, 5, 40, n_repetitions=3, repetition_period=6)
    mon_ref_el = xb.ParticlesMonitor(**kw)
    els_ref = list(els)
    els_ref[7] = _OracleMonitorElement(mon_ref)
    tbt_ref = ro.HostMonitor(0, 20, 0, n)
    hp = ro.HostParticles.from_particles(p_host)
    cm, keep = ro.make_monitor_struct(tbt_ref)
    re_ = ro.RefElements(els_ref)
    ro.track_line(hp, re_, num_turns=20, ele_start=0, num_ele_track=len(els_ref),
                  flag_end_turn_actions=True, flag_reset_s_at_end_turn=True,
                  line_length=line2.get_length(), monitor=tbt_ref, flag_monitor=1)
    dev = _build(line2, on_gpu)
    p = p_host.copy(_device=dev)
    line2.track(p, num_turns=20, turn_by_turn_monitor=True)
    assert mon.x.shape == (3, 35, 3)
    bitwise = not on_gpu           # the toy ring has a cavity: device sin
    for ff in ('x', 'px', 'zeta', 'delta', 'at_turn', 'at_element', 'particle_id', 'state'):
        a, b = mon.get(ff), mon_ref.field(ff)
        if bitwise or a.dtype.kind in 'iu':
            assert np.array_equal(a, b), ff
        else:
            assert np.max(np.abs(a - b)) <= 1e-12 * max(np.max(np.abs(b)), 1e-300), ff
        a, b = line2.record_last_track.get(ff), tbt_ref.field(ff)
        if bitwise or a.dtype.kind in 'iu':
            assert np.array_equal(a, b), ff
    assert np.any(mon.get('x')[2] != 0)


class _OracleMonitorElement:
    """Hands a `ref_oracle.HostMonitor` to `RefElements` as an in-line ParticlesMonitor."""

    def __init__(self, host_monitor):
        self.host_monitor = host_monitor


def _make_monitor_passthrough(self, mon):
    cm, keep = ro.make_monitor_struct(mon.host_monitor)
    self._keep += [cm, keep]
    import ctypes as ct
    return ct.addressof(cm)


_orig_make = ro.RefElements._make


def _make(self, el):
    if isinstance(el, _OracleMonitorElement):
        return _make_monitor_passthrough(self, el), 1000
    return _orig_make(self, el)


ro.RefElements._make = _make


# ---- F: track flags ----------------------------------------------------------------------------
@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('flag', ['XS_FLAG_KILL_CAVITY_KICK', 'XS_FLAG_IGNORE_GLOBAL_APERTURE',
                                  'XS_FLAG_IGNORE_LOCAL_APERTURE'])
def test_track_flags(flag, on_gpu):
    """track_flags.py:5-12: bit 2 suppresses the cavity kick (track_rf.h:352), bit 3 the global
    aperture check (local_particle_custom_api.h:262-289), bit 4 the local apertures
    (limitrect.h / limitellipse.h / limitpolygon.h)."""
    bit = {'XS_FLAG_KILL_CAVITY_KICK': 2, 'XS_FLAG_IGNORE_GLOBAL_APERTURE': 3,
           'XS_FLAG_IGNORE_LOCAL_APERTURE': 4}[flag]
    line = common.load_line('sps')
    line.config['XTRACK_GLOBAL_XY_LIMIT'] = 0.04
    p_host = common.gaussian_particles(line, 300, 3, common.SIGMAS['sps'], scale=6.0)
    plain = common.oracle_track(line, p_host, 3)
    ref = common.oracle_track(line, p_host, 3, track_flags=1 << bit)
    if bit == 2:
        assert not np.array_equal(plain['delta'], ref['delta'])
        alive = ref['state'] > 0
        assert np.array_equal(ref['delta'][alive], p_host.get('delta')[alive])   # no energy change
    else:
        code = -1 if bit == 3 else 0
        assert (plain['state'] == code).sum() > 3 and (ref['state'] == code).sum() == 0
    line.track_flags[flag] = True
    got = _track(line, p_host, on_gpu, num_turns=3)
    # (with the cavity kick suppressed the ring is libm-free: bit identity on the GPU too)
    if bit == 2 or not on_gpu:
        _compare(got, ref, True, flag)
    else:
        yard = common.libm_yardstick(line, p_host, 3, ref=ref, track_flags=1 << bit)
        for ff in INTS:
            assert np.array_equal(got[ff], ref[ff]), ff
        common.assert_parity(got, ref, yard, True, mask=ref['state'] > 0, label=flag)
    line.track_flags[flag] = False
    _compare(_track(line, p_host, on_gpu, num_turns=3), plain, not on_gpu, 'flags off',
             rtol=1e-10)


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_unsupported_flags_are_rejected(on_gpu):
    line = common.toy_ring(thin=True)
    dev = _build(line, on_gpu)
    p = common.gaussian_particles(line, 4, 1, common.SIGMAS['toy'], device=dev)
    with pytest.raises(NotImplementedError):
        line.track(p, backtrack=True, turn_by_turn_monitor=True)
    if on_gpu:
        line.track_flags['XS_FLAG_SR_TAPER'] = True
        with pytest.raises(Exception):
            line.track(p)


# ---- R: unseeded generator ----------------------------------------------------------------------
@pytest.mark.gpu
def test_unseeded_generator_kills_particles_gpu():
    """random_src/uniform.h:34-53: an all-zero generator state kills the particle with state
    -20 (see tests/test_radiation_monitors_hostsim.py for the host tier)."""
    import math
    line = xb.Line(elements=[xb.Bend(length=1.0, angle=0.02, k0='from_h')])
    line.particle_ref = xb.Particles(p0c=5e9, mass0=xb.ELECTRON_MASS_EV)
    line.configure_radiation(model='quantum')
    line._extra_config['_needs_rng'] = False
    p_host = xb.Particles(p0c=5e9, x=np.zeros(100), px=1e-4, mass0=xb.ELECTRON_MASS_EV)
    got = _track(line, p_host, True)
    assert np.all(got['state'] == -20)
    assert np.all(got['at_element'] == 0) and np.all(got['at_turn'] == 0)
    for ff in ('x', 'px', 'y', 'py', 'zeta'):
        assert np.all(got[ff] == 1e30), ff
    assert np.all(got['delta'] == -1.0)
