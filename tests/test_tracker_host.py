"""Host logic of `Tracker` (no GPU): the launch plan of a `track()` call and the
invalidation of the lowered lattice when an element is edited after `build_tracker`."""
import numpy as np
import pytest

import xtrack_b200 as xb
from xtrack_b200.tracker import split_turns
import common
import hostsim


def _walk(n_line, ele_start, plan):
    """(elements traversed, end-of-turn actions performed, final position) of a plan."""
    pos, done, ends = ele_start, 0, 0
    done += plan.head
    pos += plan.head
    if plan.head_ends_turn:
        assert pos == n_line
        ends += 1
    if pos == n_line:
        pos = 0
    if plan.full_turns:
        assert pos == 0
        done += plan.full_turns * n_line
        ends += plan.full_turns if plan.full_turns_end_turn else 0
    if plan.tail:
        assert pos == 0
        done += plan.tail
        pos = plan.tail
    return done, ends, pos


# (n_line, ele_start, kwargs) -> (head, full_turns, tail, head_ends_turn, monitor_turns)
# expected values worked out from the semantics of xtrack's Line.track arguments
# (tests/test_tracker.py:153 of the reference exercises the same combinations)
TABLE = [
    ((10, 0, dict()), (10, 0, 0, True, 1)),
    ((10, 0, dict(num_turns=5)), (10, 4, 0, True, 5)),
    ((10, 3, dict(num_turns=5)), (7, 4, 0, True, 5)),
    ((10, 3, dict(ele_stop=8)), (5, 0, 0, False, 1)),
    ((10, 3, dict(ele_stop=8, num_turns=3)), (7, 1, 8, True, 3)),
    ((10, 8, dict(ele_stop=3)), (2, 0, 3, True, 2)),            # stop lies in the next turn
    ((10, 8, dict(ele_stop=3, num_turns=2)), (2, 1, 3, True, 3)),
    ((10, 3, dict(ele_stop=3)), (7, 0, 3, True, 2)),            # a full turn from the middle
    ((10, 0, dict(ele_stop=10, num_turns=2)), (10, 0, 10, True, 2)),   # last stretch = whole line
    ((10, 0, dict(ele_stop=0)), (10, 0, 0, True, 1)),
    ((10, 2, dict(num_elements=5)), (5, 0, 0, False, 1)),
    ((10, 2, dict(num_elements=8)), (8, 0, 0, True, 1)),
    ((10, 2, dict(num_elements=9)), (8, 0, 1, True, 2)),
    ((10, 2, dict(num_elements=38)), (8, 3, 0, True, 4)),
    ((10, 2, dict(num_elements=41)), (8, 3, 3, True, 5)),
    ((10, 0, dict(num_elements=0)), (0, 0, 0, False, 1)),
    ((10, 10, dict(num_elements=3)), (0, 0, 3, True, 2)),
]


@pytest.mark.parametrize('args,expected', TABLE)
def test_split_turns_table(args, expected):
    n_line, ele_start, kw = args
    plan = split_turns(n_line, ele_start, **kw)
    got = (plan.head, plan.full_turns, plan.tail, plan.head_ends_turn, plan.monitor_turns)
    assert got == expected, plan
    assert plan.full_turns_end_turn is True
    # the plan covers exactly the requested stretch
    done, ends, pos = _walk(n_line, ele_start, plan)
    if 'num_elements' in kw:
        assert done == kw['num_elements']
        assert pos == (ele_start + kw['num_elements']) % n_line or plan.tail == 0
    else:
        turns = kw.get('num_turns', 1)
        stop = kw.get('ele_stop')
        if stop is None:
            assert done == turns * n_line - ele_start and ends == turns
        else:
            extra = 1 if stop <= ele_start else 0
            assert done == (turns - 1 + extra) * n_line + stop - ele_start
    plan = split_turns(n_line, ele_start, skip_end_turn_actions=True, **kw)
    assert not plan.head_ends_turn and not plan.full_turns_end_turn


def test_split_turns_rejects_conflicting_arguments():
    with pytest.raises(ValueError):
        split_turns(10, 0, num_elements=3, ele_stop=4)
    with pytest.raises(ValueError):
        split_turns(10, 0, num_elements=3, num_turns=2)
    with pytest.raises(ValueError):
        split_turns(10, 0, num_turns=0)


def _final(line, p_host, turns=2):
    p = p_host.copy()
    line.track(p, num_turns=turns)
    return common.by_id(p)


def test_element_edits_after_build_take_effect():
    """In the reference the elements are views into the tracker's buffer: an edit after
    `build_tracker` is seen by the next `track()`.  Here the tracker lowers again."""
    line = common.toy_ring(thin=True)
    p_host = common.gaussian_particles(line, 30, 1, common.SIGMAS['toy'])
    hostsim.build_hostsim_tracker(line)
    tracker = line.tracker
    a = _final(line, p_host)
    lattice_a = tracker._lattice
    b = _final(line, p_host)
    assert tracker._lattice is lattice_a            # nothing changed: no new lowering
    assert np.array_equal(a['px'], b['px'])

    line[1].knl[1] = 0.45                           # item assignment into a coefficient array
    c = _final(line, p_host)
    assert tracker._lattice is not lattice_a
    assert not np.array_equal(a['px'], c['px'])
    ref = common.oracle_track(line, p_host, 2)
    for ff in ('x', 'px', 'y', 'py', 'zeta', 'delta'):
        assert np.array_equal(c[ff], ref[ff]), ff

    line.elements[-1].voltage = 3e5                 # scalar field (cavity voltage scan)
    d = _final(line, p_host)
    assert not np.array_equal(c['delta'], d['delta'])
    ref = common.oracle_track(line, p_host, 2)
    assert np.array_equal(d['delta'], ref['delta'])

    line.element_names = line.element_names[:-1]    # the sequence itself
    e = _final(line, p_host)
    assert np.array_equal(e['delta'], p_host.get('delta'))      # no cavity left
    assert np.all(e['at_turn'] == 2)


def test_bend_and_edge_setters_follow_the_reference():
    b = xb.Bend(length=2.0, angle=0.1, k0='from_h')
    assert b.h == 0.05 and b.k0 == 0.05
    b.angle = 0.2
    assert b.h == 0.1 and b.k0 == 0.1
    b.k0 = 0.3
    assert b.k0 == 0.3 and not b.k0_from_h
    b.length = 4.0
    assert b.h == 0.05 and b.k0 == 0.3
    r = xb.RBend(length_straight=2.0, angle=0.1, k0='from_h')
    h0 = r.h
    r.length_straight = 4.0
    assert abs(r.h - h0 / 2) < 1e-15 and r.k0 == r.h
    e = xb.DipoleEdge(k=0.05, e1=0.03, hgap=0.02, fint=0.4)
    r21 = e.r21
    e.k = 0.1
    assert abs(e.r21 - 2 * r21) < 1e-18


def test_implicit_rebuild_keeps_build_options():
    line = common.toy_ring(thin=True)
    line.build_tracker(_device='cpu', exact_arithmetic=False, fuse=False,
                       _tracker_class=hostsim.HostSimTracker)
    line.configure_radiation(model='mean')          # invalidates the tracker
    assert line.tracker is None
    p = common.gaussian_particles(line, 4, 1, common.SIGMAS['toy'])
    line.track(p, num_turns=1)
    assert isinstance(line.tracker, hostsim.HostSimTracker)
    assert line.tracker.exact_arithmetic is False and line.tracker.fuse is False


def test_particles_device_is_normalised():
    import torch
    from xtrack_b200.particles import normalise_device
    assert normalise_device('cpu') == torch.device('cpu')
    assert normalise_device('cuda:1') == torch.device('cuda', 1)
    if torch.cuda.is_available():
        assert normalise_device('cuda').index is not None


@pytest.mark.parametrize('name', ['hllhc_14', 'sps', 'lep', 'clic_dr', 'ring', 'ring_sliced'])
def test_line_to_dict_round_trip(name, tmp_path):
    """`Line.to_dict / to_json` (line.py:775, 899) -> `from_dict / from_json`: the reloaded
    line lowers to the same op stream, word for word; slices find their parents again."""
    from xtrack_b200 import lowering
    line = common.load_line(name)
    if name == 'lep':
        line.configure_radiation(model='mean')
    path = tmp_path / (name + '.json.gz')
    line.to_json(path)
    line2 = xb.Line.from_json(path, replace_unsupported=True)
    assert line2.element_names == line.element_names
    assert line2.config == line.config
    assert line2._extra_config['_radiation_model'] == line._extra_config['_radiation_model']
    synrad = name == 'lep'
    pa = lowering.lower_line(line.elements, synrad=synrad)
    pb = lowering.lower_line(line2.elements, synrad=synrad)
    for fused in (False, True):
        wa, oa = pa.finish(fused=fused)
        wb, ob = pb.finish(fused=fused)
        assert np.array_equal(wa, wb) and np.array_equal(oa, ob)
    for ff in ('p0c', 'beta0', 'gamma0'):
        assert np.array_equal(line2.particle_ref.get(ff), line.particle_ref.get(ff))
    assert line2.particle_ref.mass0 == line.particle_ref.mass0
    line3 = line.copy()
    assert np.array_equal(lowering.lower_line(line3.elements, synrad=synrad).finish()[0],
                          pa.finish()[0])
