"""Slice elements (SURVEY §8a row E-wrap: ThinSlice* / ThickSlice* / DriftSlice* /
ThinSlice*Entry / Exit, `_parent` resolution, `weight`, `slice_offset` in the misalignment
wrapper, state -42, Replica) against the reference's own generated slice wrappers
(elements_src/{thin,thick,drift}_slice_*.h through the oracle), on both tiers."""
import numpy as np
import pytest

import xtrack_b200 as xb
import common
import slicing_helper as sh
from test_rows_both_tiers import BACKENDS, _track, _compare, _line, MIS, INTS


def _els_for_slicing(mis=None):
    mis = mis or {}
    return [xb.Drift(length=0.4),
            xb.Quadrupole(length=0.6, k1=0.3, k1s=0.01, knl=[0, 0, 0.2], edge_entry_active=1,
                          edge_exit_active=1, **mis),
            xb.Drift(length=0.3),
            xb.Sextupole(length=0.4, k2=3.0, k2s=0.3, num_multipole_kicks=3, **mis),
            xb.Octupole(length=0.3, k3=200., knl=[1e-4], **mis),
            xb.Bend(length=1.5, angle=0.1, k0='from_h', k1=0.02, knl=[0, 0, 0.5],
                    edge_entry_angle=0.02, edge_exit_angle=0.03, edge_entry_fint=0.5,
                    edge_entry_hgap=0.02, edge_exit_fint=0.5, edge_exit_hgap=0.02, **mis),
            xb.Bend(length=1.0, angle=0.05, k0=0.051, model='bend-kick-bend',
                    edge_entry_model='full', edge_exit_model='full', num_multipole_kicks=4),
            xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='straight-body',
                     rbend_angle_diff=0.01, k1=0.05, **mis),
            xb.RBend(length_straight=1.2, angle=0.08, k0='from_h', rbend_model='curved-body',
                     edge_entry_fint=0.4, edge_entry_hgap=0.02),
            xb.Multipole(knl=[0.01, 0.2, 1.0], ksl=[0, 0.05], hxl=0.01, length=0.4, isthick=True,
                         num_multipole_kicks=3, **mis),
            xb.Cavity(length=0.5, voltage=2e5, frequency=4e8, lag=150., **mis),
            xb.Drift(length=0.2)]


@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('mode,scheme,n', [('thin', 'teapot', 4), ('thin', 'uniform', 1),
                                          ('thick', 'uniform', 3), ('thin', 'teapot', 2)])
def test_sliced_elements_vs_reference_slices(mode, scheme, n, on_gpu):
    """Every sliceable class, thin (teapot / uniform) and thick slicing, with edge slices."""
    line = sh.slice_line(_line(_els_for_slicing()), n=n, mode=mode, scheme=scheme)
    names = {type(ee).__name__ for ee in line.elements}
    assert {'ThinSliceBendEntry', 'ThinSliceRBendExit', 'ThinSliceQuadrupoleEntry'} <= names
    assert (('ThinSliceQuadrupole' in names and 'DriftSliceRBend' in names) if mode == 'thin'
            else 'ThickSliceCavity' in names)
    p_host = common.gaussian_particles(line, 201, 3, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 2)
    # thin slices are libm-free apart from the cavity's sine and the full-model fringes
    _compare(_track(line, p_host, on_gpu, num_turns=2), ref, not on_gpu, f'{mode} {scheme} {n}',
             rtol=1e-12)


@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('mode', ['thin', 'thick'])
def test_slices_of_misaligned_parents(mode, on_gpu):
    """The parent's shifts and tilts act on every slice, the anchor moved by `slice_offset`
    (track_local_particle_with_transformations.h:143-146); curved parents: rigid-matrix path
    per slice with angle * weight."""
    mis = dict(shift_x=1e-3, shift_y=-2e-3, shift_s=5e-3, rot_s_rad=0.02, rot_shift_anchor=0.2)
    els = _els_for_slicing(mis)
    if mode == 'thick':      # x / y rotations: allowed for thick slices (and edge slices)
        els[5] = xb.Bend(length=1.5, angle=0.1, k0='from_h', **MIS)
        els[1] = xb.Quadrupole(length=0.6, k1=0.3, **MIS)
    line = sh.slice_line(_line(els), n=3, mode=mode)
    p_host = common.gaussian_particles(line, 150, 3, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    assert (ref['state'] > 0).all()
    _compare(_track(line, p_host, on_gpu), ref, not on_gpu, 'misaligned slices ' + mode,
             rtol=1e-12)


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_thin_slice_of_rotated_curved_parent_is_invalid(on_gpu):
    """A thin slice of a curved parent with an x / y rotation cannot be transformed: the
    particle is flagged with XT_INVALID_THIN_SLICE_TRANSFORM = -42 and the slice is not
    tracked (track_local_particle_with_transformations.h:132-139)."""
    els = [xb.Drift(length=0.5), xb.Bend(length=1.0, angle=0.05, k0='from_h', rot_x_rad=1e-3),
           xb.Drift(length=0.5)]
    line = sh.slice_line(_line(els), n=2, mode='thin')
    p_host = common.gaussian_particles(line, 64, 3, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 1)
    assert np.all(ref['state'] == -42)
    got = _track(line, p_host, on_gpu)
    _compare(got, ref, True, 'state -42')
    # the edge slice in front of it and the drift slice were tracked, the thin slice was not
    assert np.all(got['at_element'] == 3)


@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('name', ['ring', 'ring_sliced'])
def test_reference_ring_with_replicas_and_slices(name, on_gpu):
    """examples/lattice_design/ring.json of the reference: a line of `Replica`s of thick
    elements, and the same line with the elements replaced by the slices the file holds for
    them (tests/golden/make_lattices.py).  Sliced and unsliced rings agree to the accuracy of
    the integrator."""
    line = common.load_line(name)
    classes = {type(ee).__name__ for ee in line.elements}
    if name == 'ring_sliced':
        assert {'ThickSliceBend', 'DriftSlice', 'ThinSliceBendEntry', 'ThickSliceQuadrupole',
                'ThickSliceSextupole'} <= classes
    p_host = common.gaussian_particles(line, 200, 1, common.SIGMAS['toy'], scale=0.3)
    ref = common.oracle_track(line, p_host, 3, parallel=True)
    got = _track(line, p_host, on_gpu, num_turns=3)
    if not on_gpu:
        _compare(got, ref, True, name)
    else:
        yard = common.libm_yardstick(line, p_host, 3, ref=ref, parallel=True)
        for ff in INTS:
            assert np.array_equal(got[ff], ref[ff]), ff
        common.assert_parity(got, ref, yard, True, label=name)
    if name == 'ring_sliced':
        whole = common.oracle_track(common.load_line('ring'), p_host, 3, parallel=True)
        dev = common.max_rel_dev(ref, whole, fields=('x', 'px', 'y', 'py'))
        assert max(dev.values()) < 1e-6, dev


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_sliced_lep(on_gpu):
    """The LEP thick lattice, every magnet cut into 4 thin teapot slices + edge slices
    (9 230 -> ~40 000 elements), and into 2 thick slices."""
    line0 = common.load_line('lep')
    for mode, n, turns in (('thin', 4, 2), ('thick', 2, 1)):
        line = sh.slice_line(line0, n=n, mode=mode)
        assert len(line) > (3 if mode == "thin" else 2) * len(line0)
        p_host = common.gaussian_particles(line, 100, 11, common.SIGMAS['lep'])
        ref = common.oracle_track(line, p_host, turns, parallel=True)
        got = _track(line, p_host, on_gpu, num_turns=turns)
        if not on_gpu:
            _compare(got, ref, True, 'lep ' + mode)
        else:
            yard = common.libm_yardstick(line, p_host, turns, ref=ref, parallel=True)
            for ff in INTS:
                assert np.array_equal(got[ff], ref[ff]), ff
            common.assert_parity(got, ref, yard, True, label='lep sliced ' + mode)


def test_slice_host_model():
    line = sh.slice_line(_line(_els_for_slicing()), n=4, mode='thin')
    assert abs(line.get_length() - _line(_els_for_slicing()).get_length()) < 1e-12
    # the dictionary form round-trips: parents are found by name, replicas are followed
    dd = {'elements': {}, 'element_names': list(line.element_names)}
    for nn, ee in line.element_dict.items():
        if isinstance(ee, xb.elements._Slice):
            dd['elements'][nn] = {'__class__': type(ee).__name__, 'parent_name': ee.parent_name,
                                  'weight': ee.weight, 'slice_offset': ee.slice_offset}
    parents = {ee.parent_name for ee in line.element_dict.values()
               if isinstance(ee, xb.elements._Slice)}
    with pytest.raises(KeyError):
        xb.Line.from_dict(dict(dd, elements={k: v for k, v in dd['elements'].items()}))
    rep = xb.elements.Replica(parent_name='a')
    assert rep.resolve({'a': xb.elements.Replica('b'), 'b': 7}) == 7
    with pytest.raises(RecursionError):
        rep.resolve({'a': xb.elements.Replica('b'), 'b': xb.elements.Replica('a')})
    assert len(parents) == 9
