"""`line.track(..., backtrack=True | 'force')` (XS_FLAG_BACKTRACK; tracker.py:626-646, 702-731,
1222-1235; the per-class inverse maps: drift.h:16-19, track_magnet.h:413-433, track_magnet_edge.h:
70-86, track_rf.h:364-373, dipoleedge.h:37-60, srotation.h:19, xyshift.h:18, rotation.h:25-33,
translation.h:18, track_misalignments.h:59-74,97-112,225-240,362-377, drift_slice_*.h) against
the reference's C run with the flag set, on both tiers: the host build of the device code
(bit identity) and, marked `gpu`, the CUDA kernel.  The product lowers the inverse lattice on
the host (lowering.lower_line(backtrack=True)) and runs it forwards.
"""
import numpy as np
import pytest

import xtrack_b200 as xb
import common
from test_rows_both_tiers import BACKENDS, MIS, _build, _compare, _line

BACKTRACK = 1        # track_flags.py:6


def _back(line, p_host, on_gpu, **kw):
    dev = _build(line, on_gpu)
    p = p_host.copy(_device=dev)
    line.track(p, backtrack=kw.pop('backtrack', 'force'), **kw)
    return common.by_id(p)


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_backtrack_part_of_a_thin_ring(on_gpu):
    """800 elements of the SPS line forwards, then backwards: the reference's result to the
    bit, at_element counted down, and the start recovered to rounding."""
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 60, 3, common.SIGMAS['sps'])
    dev = _build(line, on_gpu)
    p = p_host.copy(_device=dev)
    line.track(p, ele_start=100, ele_stop=900)
    fwd = p.copy(_device='cpu')
    ref = common.oracle_track(line, fwd, 1, ele_start=100, num_ele_track=800,
                              flag_end_turn_actions=False, track_flags=BACKTRACK)
    line.track(p, ele_start=100, ele_stop=900, backtrack=True)
    got = common.by_id(p)
    _compare(got, ref, True, 'sps backtrack')
    assert np.all(got['at_element'] == 0)
    start = common.by_id(p_host)
    for ff, tol in (('x', 1e-14), ('px', 1e-16), ('y', 1e-14), ('py', 1e-16), ('zeta', 1e-14), ('s', 1e-12)):
        assert np.max(np.abs(got[ff] - start[ff])) < tol, ff


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_backtrack_whole_turns(on_gpu):
    """Two complete turns backwards: increment_at_turn_backtrack first (at_turn - 1, at_element
    = len(line), s = line length), then the line from its end (tracker.py:630-636)."""
    line = common.load_line('sps')
    els = [ee for ee in line.elements if type(ee).__name__ != 'Cavity']
    line = _line(els, p0c=float(line.particle_ref.get('p0c')[0]))
    p_host = common.gaussian_particles(line, 40, 5, common.SIGMAS['sps'])
    ref = common.oracle_track(line, p_host, 2, track_flags=BACKTRACK)
    got = _back(line, p_host, on_gpu, num_turns=2, backtrack=True)
    _compare(got, ref, True, 'two turns back')
    assert np.all(got['at_turn'] == -2) and np.all(got['at_element'] == 0)


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_backtrack_element_zoo(on_gpu):
    """Every class with a backtrack branch that is free of per-particle libm: misaligned thin
    and thick magnets (entry / exit transformations swapped and inverted), bends with linear
    edges (swapped, r21 / r43 negated), rbends of both body models, frame elements, apertures,
    exact drifts, a thick multipole."""
    els = [xb.Drift(length=0.7), xb.Multipole(knl=[1e-3, 0.1, 2.0], ksl=[0, 0.05], **MIS),
           xb.DriftExact(length=1.0),
           xb.Bend(length=1.5, angle=0.1, k0='from_h', edge_entry_angle=0.02, edge_exit_angle=0.03,
                   edge_entry_fint=0.5, edge_entry_hgap=0.02, **MIS),
           xb.Bend(length=1.0, angle=0.05, k0='from_h', k1=0.02, model='bend-kick-bend',
                   rot_x_rad=2e-3, rot_y_rad=1e-3),
           xb.Sextupole(length=0.3, k2=5., rot_s_rad=0.3, shift_y=1e-3),
           xb.LimitEllipse(a=0.05, b=0.03, rot_s_rad=0.2, shift_y=1e-3),
           xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='straight-body', **MIS),
           xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='curved-body',
                    edge_entry_angle=0.01, edge_exit_angle=-0.02),
           xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4, shift_x=1e-3, rot_s_rad=0.05),
           xb.SRotation(angle=20.), xb.XYShift(dx=1e-3, dy=-1e-3),
           xb.Rotation(rot_s_rad=0.02, rot_x_rad=-3e-3, rot_y_rad=2e-3, seq='xys'),
           xb.Translation(shift_x=1e-4, shift_y=-2e-4),
           xb.Multipole(knl=[0.01, 0.2, 1.0], hxl=0.01, length=0.4, isthick=True,
                        num_multipole_kicks=3),
           xb.LimitRect(min_x=-0.05, max_x=0.05, min_y=-0.05, max_y=0.05, shift_x=0.01, rot_s_rad=0.1),
           xb.Octupole(length=0.2, k3=50., shift_x=-1e-3, rot_shift_anchor=0.1, rot_y_rad=1e-3),
           xb.Marker()]
    line = _line(els)
    assert line._is_backtrackable
    p_host = common.gaussian_particles(line, 200, 2, common.SIGMAS['toy'], scale=3.)
    # (up to, not including, the last element: a stretch that reaches the end of the line ends
    # the turn, tracker.py:1340-1370)
    ref = common.oracle_track(line, p_host, 1, num_ele_track=len(els) - 1,
                              flag_end_turn_actions=False, track_flags=BACKTRACK)
    got = _back(line, p_host, on_gpu, ele_start=0, ele_stop=len(els) - 1, backtrack=True)
    _compare(got, ref, True, 'zoo')
    # and a stretch in the middle, with losses on the apertures
    p_wide = common.gaussian_particles(line, 300, 7, common.SIGMAS['toy'], scale=25.)
    ref = common.oracle_track(line, p_wide, 1, ele_start=3, num_ele_track=12,
                              flag_end_turn_actions=False, track_flags=BACKTRACK)
    assert 5 < (ref['state'] <= 0).sum() < 295
    _compare(_back(line, p_wide, on_gpu, ele_start=3, ele_stop=15), ref, True, 'zoo, losses')


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_backtrack_with_device_libm_and_full_edges(on_gpu):
    """Quadrupole matrices, thin and thick cavities, an RF multipole, a crab cavity (voltages
    and lengths negated, track_rf.h:364-373); the full edge model cannot be backtracked: the
    reference kills the particle with state -32 (track_magnet_edge.h:83-87, dipoleedge.h:55-60)."""
    els = [xb.Quadrupole(length=0.5, k1=0.3, **MIS), xb.Drift(length=0.5),
           xb.Cavity(voltage=1e5, frequency=4e8, lag=30., shift_x=2e-3, rot_s_rad=0.1),
           xb.Cavity(length=0.4, voltage=2e5, frequency=4e8, lag=150.),
           xb.RFMultipole(voltage=1e4, frequency=4e8, lag=10., knl=[1e-4, 1e-2], pn=[20., 40.]),
           xb.CrabCavity(crab_voltage=1e5, frequency=4e8, lag=10.),
           xb.Quadrupole(length=0.5, k1=-0.3, k1s=0.01), xb.Marker()]
    line = _line(els)
    p_host = common.gaussian_particles(line, 150, 2, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1, num_ele_track=len(els) - 1,
                              flag_end_turn_actions=False, track_flags=BACKTRACK)
    _compare(_back(line, p_host, on_gpu, ele_start=0, ele_stop=len(els) - 1), ref, not on_gpu, 'libm')

    for el in (xb.Bend(length=1.0, angle=0.05, k0='from_h', edge_entry_model='full',
                       edge_exit_model='linear'),
               xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4, model='full')):
        line = _line([xb.Drift(length=1.0), el, xb.Drift(length=1.0)])
        ref = common.oracle_track(line, p_host, 1, num_ele_track=2, flag_end_turn_actions=False,
                                  track_flags=BACKTRACK)
        assert np.all(ref['state'] == -32)
        got = _back(line, p_host, on_gpu, ele_start=0, ele_stop=2)
        assert np.array_equal(got['state'], ref['state'])
        assert np.array_equal(got['at_element'], ref['at_element'])


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_backtrack_sliced_ring(on_gpu):
    """Thin / thick / drift / edge slices of the sliced toy ring (drift_slice_*.h,
    the parents' calls with the weights and lengths negated)."""
    line = common.load_line('ring_sliced')
    n = len(line.element_names) - 1
    p_host = common.gaussian_particles(line, 100, 11, common.SIGMAS['toy'], scale=0.3)
    ref = common.oracle_track(line, p_host, 1, ele_start=0, num_ele_track=n,
                              flag_end_turn_actions=False, track_flags=BACKTRACK)
    _compare(_back(line, p_host, on_gpu, ele_start=0, ele_stop=n), ref, not on_gpu, 'slices',
             rtol=1e-12)


def test_backtrack_refused_for_lines_without_inverse():
    import hostsim
    line = _line([xb.Drift(length=1.0), xb.LastTurnsMonitor(n_last_turns=2, num_particles=5)])
    line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker)
    p = common.gaussian_particles(line, 5, 1, common.SIGMAS['toy'])
    assert not line._is_backtrackable
    with pytest.raises(ValueError, match='not backtrackable'):
        line.track(p, backtrack=True)
    line.track(p, ele_start=0, ele_stop=1, backtrack='force')      # the caller insists
    assert np.allclose(p.get('s'), -1.0)


@pytest.mark.gpu
def test_full_size_round_trip_on_gpu():
    """BASELINE.json configs[1] at its full per-GPU size, through a size-independent property:
    10^6 particles two turns forwards then two turns backwards through hllhc_14 are where they
    started (the inverse lattice undoes the lattice: every lowered inverse map, the reversed
    order and the turn bookkeeping at once)."""
    line = common.load_line('hllhc_14')
    n, nr = 1_000_000, 1000
    r = np.linspace(0, 2e-3, nr + 1)[1:]
    th = np.linspace(0, np.pi / 2, n // nr)
    rr, tt = np.meshgrid(r, th, indexing='ij')
    ref_p = line.particle_ref
    p_host = xb.Particles(x=(rr * np.cos(tt)).ravel(), y=(rr * np.sin(tt)).ravel(),
                          delta=np.full(n, 2.7e-4), p0c=float(ref_p.get('p0c')[0]),
                          mass0=ref_p.mass0, q0=ref_p.q0)
    line.build_tracker(_device='cuda:0', exact_arithmetic=True)
    p = p_host.copy(_device='cuda:0')
    line.track(p, num_turns=2)
    mid = common.by_id(p)
    assert np.all(mid['state'] > 0) and np.all(mid['at_turn'] == 2)
    assert np.max(np.abs(mid['x'] - p_host.get('x'))) > 1e-4          # it did move
    line.track(p, num_turns=2, backtrack=True)
    end = common.by_id(p)
    assert np.all(end['state'] > 0) and np.all(end['at_turn'] == 0) and np.all(end['at_element'] == 0)
    for ff, tol in (('x', 1e-12), ('px', 1e-14), ('y', 1e-12), ('py', 1e-14), ('zeta', 1e-11),
                    ('delta', 1e-15)):
        assert np.max(np.abs(end[ff] - p_host.get(ff))) < tol, (ff, np.max(np.abs(end[ff] - p_host.get(ff))))
