"""GPU parity tests (-m gpu): the CUDA kernel, called through the C-ABI, against the
reference-header oracle on identical inputs.

Bar (BASELINE.json north_star): FP64 coordinates within 1e-12 relative over the first
10 turns; loss turn / element / state bit-exact where coordinates agree.  The EXACT
kernel variant (no FMA contraction) is additionally expected to reproduce the oracle
to the last bit wherever no libm call is involved.
"""
import numpy as np
import pytest
import torch

import xtrack_b200 as xb
import common

pytestmark = pytest.mark.gpu

RTOL = 1e-12      # relative to the beam size of each coordinate (common.max_rel_dev)


def _track_gpu(line, p_host, num_turns, exact, **kw):
    p = p_host.copy(_device='cuda:0')
    line.build_tracker(_device='cuda:0', exact_arithmetic=exact)
    line.track(p, num_turns=num_turns, **kw)
    torch.cuda.synchronize()
    return p


@pytest.mark.parametrize('exact', [True, False], ids=['exact', 'fma'])
@pytest.mark.parametrize('name', ['hllhc_14', 'sps', 'clic_dr', 'lep'])
def test_ten_turns_vs_oracle(name, exact):
    line = common.load_line(name)
    n = 3000 if name != 'lep' else 1000
    p_host = common.gaussian_particles(line, n, 11, common.SIGMAS[name])
    ref = common.oracle_track(line, p_host, 10)
    got = common.by_id(_track_gpu(line, p_host, 10, exact))
    assert np.array_equal(got['particle_id'], ref['particle_id'])
    alive = ref['state'] > 0
    dev = common.max_rel_dev(got, ref, mask=alive)
    print(name, 'exact' if exact else 'fma', dev,
          'bit-identical fraction x:', float(np.mean(got['x'] == ref['x'])))
    assert max(dev.values()) < RTOL, dev
    for ff in ('state', 'at_turn', 'at_element'):
        assert np.array_equal(got[ff], ref[ff]), ff
    np.testing.assert_allclose(got['s'], ref['s'], rtol=1e-13, atol=1e-9)


@pytest.mark.parametrize('thin', [True, False], ids=['thin', 'thick'])
def test_toy_ring(thin):
    line = common.toy_ring(thin=thin)
    p_host = common.gaussian_particles(line, 10000, 1, common.SIGMAS['toy'])
    ref = common.oracle_track(line, p_host, 10)
    got = common.by_id(_track_gpu(line, p_host, 10, exact=False))
    dev = common.max_rel_dev(got, ref, mask=ref['state'] > 0)
    assert max(dev.values()) < RTOL, dev
    for ff in ('state', 'at_turn', 'at_element'):
        assert np.array_equal(got[ff], ref[ff]), ff


def test_losses_sps_apertures():
    """SPS stand-in with 1101 LimitRect + 597 LimitEllipse: a wide beam loses particles
    on the local apertures; state / at_turn / at_element must be identical."""
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 20000, 3, common.SIGMAS['sps'], scale=6.0)
    ref = common.oracle_track(line, p_host, 20)
    n_lost = int((ref['state'] <= 0).sum())
    assert 200 < n_lost < 19000, n_lost
    for exact in (True, False):
        got = common.by_id(_track_gpu(line, p_host, 20, exact))
        same = (np.array_equal(got['state'], ref['state']), np.mean(got['at_turn'] == ref['at_turn']),
                np.mean(got['at_element'] == ref['at_element']))
        print('exact' if exact else 'fma', 'lost', n_lost, same)
        frac = np.mean((got['state'] == ref['state']) & (got['at_turn'] == ref['at_turn'])
                       & (got['at_element'] == ref['at_element']))
        assert frac >= 0.9999, frac
        lost = ref['state'] <= 0
        ok = lost & (got['at_element'] == ref['at_element']) & (got['at_turn'] == ref['at_turn'])
        # coordinates of lost particles are frozen where they were lost
        dev = common.max_rel_dev(got, ref, mask=ok)
        assert max(dev.values()) < RTOL, dev


def test_global_aperture_and_partial_turns():
    """Aperture-less lattice: losses on the global x/y limit (state -1); element ranges."""
    line = common.load_line('hllhc_14')
    line.config['XTRACK_GLOBAL_XY_LIMIT'] = 2e-3
    p_host = common.gaussian_particles(line, 4000, 5, common.SIGMAS['hllhc_14'], scale=3.0)
    ref = common.oracle_track(line, p_host, 3)
    assert (ref['state'] == -1).sum() > 50
    got = common.by_id(_track_gpu(line, p_host, 3, True))
    for ff in ('state', 'at_turn', 'at_element'):
        assert np.array_equal(got[ff], ref[ff]), ff
    # partial tracking: ele_start / num_elements (tests/test_tracker.py:153 of the reference)
    line.config['XTRACK_GLOBAL_XY_LIMIT'] = 1.0
    p_host = common.gaussian_particles(line, 500, 6, common.SIGMAS['hllhc_14'])
    p = p_host.copy(_device='cuda:0')
    line.build_tracker(_device='cuda:0', exact_arithmetic=True)
    line.track(p, ele_start=5000, num_elements=len(line) + 300)
    got = common.by_id(p)
    hp = common.ro.HostParticles.from_particles(p_host)
    re = common.ro.RefElements(line.elements)
    kw = dict(flag_reset_s_at_end_turn=1, line_length=line.get_length())
    common.ro.track_line(hp, re, num_turns=1, ele_start=5000, num_ele_track=len(line) - 5000,
                         flag_end_turn_actions=1, **kw)
    common.ro.track_line(hp, re, num_turns=1, ele_start=0, num_ele_track=5300,
                         flag_end_turn_actions=0, **kw)
    ref = hp.sorted_by_id()
    dev = common.max_rel_dev(got, ref)
    assert max(dev.values()) < RTOL, dev
    assert np.array_equal(got['at_element'], ref['at_element'])
    assert np.array_equal(got['at_turn'], ref['at_turn'])


def test_turn_by_turn_monitor():
    line = common.load_line('sps')
    p_host = common.gaussian_particles(line, 300, 8, common.SIGMAS['sps'], scale=5.0)
    mon_ref = common.ro.HostMonitor(0, 6, 0, 300)
    ref = common.oracle_track(line, p_host, 6, monitor=mon_ref, flag_monitor=1)
    p = p_host.copy(_device='cuda:0')
    line.build_tracker(_device='cuda:0', exact_arithmetic=True)
    line.track(p, num_turns=6, turn_by_turn_monitor=True)
    mon = line.record_last_track
    assert mon.x.shape == (300, 6)
    for ff in ('x', 'px', 'y', 'py', 'zeta', 'delta', 'ptau', 'rvv', 'rpp', 's', 'p0c', 'beta0',
               'gamma0', 'chi', 'weight'):
        a, b = mon.get(ff), mon_ref.field(ff)
        scale = max(np.max(np.abs(b)), 1e-300)
        assert np.max(np.abs(a - b)) / scale < RTOL, ff
    for ff in ('state', 'at_turn', 'at_element', 'particle_id', 'parent_particle_id'):
        assert np.array_equal(mon.get(ff), mon_ref.field(ff)), ff


def test_freeze_longitudinal():
    line = common.load_line('hllhc_14')
    p_host = common.gaussian_particles(line, 300, 9, common.SIGMAS['hllhc_14'])
    z0, d0 = p_host.get('zeta').copy(), p_host.get('delta').copy()
    p = _track_gpu(line, p_host, 3, True, freeze_longitudinal=True)
    assert np.array_equal(p.get('zeta'), z0)
    assert np.array_equal(p.get('delta'), d0)
    assert np.all(p.get('at_turn') == 3)
    assert not np.array_equal(p.get('x'), p_host.get('x'))
