"""BeamStatsMonitor (monitors/beam_stats_monitor.h:11-419,
beam_stats_monitor/beam_stats_monitor.py:120-1165) against the reference's own kernel code
compiled from where it lies: all four modes (whole beam, bunch slots, slices of bunches,
slices of a full turn for a coasting beam), weights, particle-id range, turn stride, species
sums, profiles.  Bin populations (sums of weights) and profiles must be identical; the moment
sums are equal up to the order of the additions (the reference adds with atomics).  Host
build of the device code and, marked `gpu`, the CUDA kernel.
"""
import numpy as np
import pytest

import xtrack_b200 as xb
from xtrack_b200.monitors import BSM_RAW_FIELDS
import common
import ref_oracle as ro
from test_rows_both_tiers import BACKENDS, _build

ALL_STATS = (['num_particles', 'sum_charge_ratio', 'sum_mass_ratio', 'mean_mass_ratio']
             + [f'mean_{cc}' for cc in ('x', 'px', 'y', 'py', 'zeta', 'delta', 'pzeta')]
             + [f'sigma_{cc}' for cc in ('x', 'y', 'zeta', 'delta')]
             + ['cov_x_px', 'cov_px_y', 'cov_y_pzeta', 'cov_zeta_delta', 'cov_x_delta', 'cov_py_zeta',
                'gemitt_x_projected', 'nemitt_y_projected', 'gemitt_zeta_projected'])

MODES = {
    'beam': dict(),
    'bunch': dict(bunch_spacing_zeta=2.0, filling_scheme=[1, 0, 1, 1, 1], selected_slots=[3, 0, 2]),
    'slice': dict(bunch_spacing_zeta=2.0, filled_slots=[0, 1, 2, 3], selected_slots=[2, 0],
                  zeta_range=(-0.6, 0.9), num_slices=5),
    'coasting': dict(coasting=True, num_slices=7),
}


def _ring_with_monitor(mode, n=600, turns=7):
    line = common.load_line('sps')
    kw = dict(start_at_turn=1, stop_at_turn=turns, every_n_turns=2, stats=ALL_STATS,
              particle_id_range=(10, n - 30),
              profiles={'x': dict(range=(-0.02, 0.03), num_bins=12),
                        'zeta': dict(range=(-7., 2.), num_bins=9)}, **MODES[mode])
    els = list(line.elements)
    mon = xb.BeamStatsMonitor(**kw)
    els.insert(len(els) // 3, mon)
    line2 = xb.Line(elements=els)
    line2.particle_ref = line.particle_ref
    mon_ref = xb.BeamStatsMonitor(**kw)
    total = sum(mon_ref._flat_size * cfg['num_bins'] for cfg in mon_ref._profile_config.values())
    mon_ref._host = {'moments': np.zeros((len(BSM_RAW_FIELDS), max(mon_ref._flat_size, 1))),
                     'touched': np.zeros(max(mon_ref._num_records, 1), dtype=np.int64),
                     'profile_counts': np.zeros(max(total, 1))}
    els_ref = list(els)
    els_ref[len(line.elements) // 3] = mon_ref
    sig = dict(common.SIGMAS['sps'])
    p_host = common.gaussian_particles(line2, n, 3, sig, scale=4.0)
    rng = np.random.default_rng(11)
    zeta = p_host.get('zeta')
    if mode in ('bunch', 'slice'):
        zeta = zeta * 40 - 2.0 * rng.integers(0, 5, n)         # bunches two metres apart
    elif mode == 'coasting':
        zeta = rng.uniform(-0.5, 0.5, n) * line.get_length()
    p_host.zeta = zeta
    p_host.weight = rng.uniform(0.5, 2.0, n)
    return line2, els_ref, mon_ref, mon, p_host


@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('mode', list(MODES))
def test_beam_stats_monitor_vs_reference_code(mode, on_gpu):
    turns = 7
    line2, els_ref, mon_ref, mon, p_host = _ring_with_monitor(mode, turns=turns)
    hp = ro.HostParticles.from_particles(p_host)
    ro.track_line(hp, ro.RefElements(els_ref), num_turns=turns, ele_start=0,
                  num_ele_track=len(els_ref), flag_end_turn_actions=True,
                  flag_reset_s_at_end_turn=True, line_length=line2.get_length(), global_xy_limit=1.0)
    ref = hp.sorted_by_id()
    dev = _build(line2, on_gpu)
    p = p_host.copy(_device=dev)
    line2.track(p, num_turns=turns)
    got = common.by_id(p)
    assert np.array_equal(got['state'], ref['state'])
    assert 20 < (ref['state'] <= 0).sum() < 500          # losses on the way: populations change

    hh = mon_ref._host
    flat = mon._flat_size
    assert np.array_equal(mon.touched_records, hh['touched'][:mon._num_records])
    assert hh['touched'].sum() == 3
    n_ref = hh['moments'][0][:flat]
    assert (n_ref > 0).sum() >= {'beam': 3, 'bunch': 6, 'slice': 10, 'coasting': 12}[mode]
    for ii, ff in enumerate(BSM_RAW_FIELDS):
        if ff not in mon._needed_fields:
            assert not hh['moments'][ii].any(), ff       # the reference did not touch it either
            continue
        scale = np.max(np.abs(hh['moments'][ii])) + 1e-300
        np.testing.assert_allclose(mon._raw(ff), hh['moments'][ii][:flat], rtol=0,
                                   atol=(2e-13 if not on_gpu else 1e-12) * scale, err_msg=ff)
    prof_ref, off = hh['profile_counts'], 0
    for cc, cfg in mon._profile_config.items():
        nn = flat * cfg['num_bins']
        arr = prof_ref[off:off + nn].reshape((*mon._data_shape, cfg['num_bins']))
        if mon.coasting:
            arr = arr[:, 0, :, :]
        np.testing.assert_allclose(mon.profiles[cc], arr, rtol=1e-13, atol=1e-13, err_msg=cc)
        assert arr.sum() > (10 if cc == 'x' else 0)
        off += nn

    # statistics, axes and selectors (beam_stats_monitor.py:832-905)
    level = mon.default_level
    assert level == {'beam': 'beam', 'bunch': 'bunch', 'slice': 'slice', 'coasting': 'slice'}[mode]
    npart = mon.get('num_particles')
    shape = {'beam': (3,), 'bunch': (3, 3), 'slice': (3, 2, 5), 'coasting': (3, 7)}[mode]
    assert npart.shape == shape and mon.num_particles.shape == shape
    filled = npart > 0
    sx = hh['moments'][BSM_RAW_FIELDS.index('sum_x')][:flat].reshape(mon._data_shape)
    nref = n_ref.reshape(mon._data_shape)
    if mon.coasting:
        sx, nref = sx[:, 0, :], nref[:, 0, :]
    np.testing.assert_allclose(mon.mean_x[filled], (sx / np.where(nref > 0, nref, 1))[filled], rtol=1e-9, atol=1e-15)
    assert np.all(mon.sigma_x[npart > 3 * p_host.get('weight').max()] > 0)
    beam = mon.get('num_particles', level='beam')
    assert beam.shape == (3,) and np.allclose(beam, npart.reshape(3, -1).sum(axis=1))
    assert mon.get('sigma_y', level='beam', turn=3).shape == ()
    assert list(mon.turns) == [1, 3, 5]
    assert np.all(mon.get('gemitt_x_projected', level='beam') > 0)
    bg = float(p_host.get('beta0')[0] * p_host.get('gamma0')[0])
    np.testing.assert_allclose(mon.get('nemitt_y_projected', level='beam'),
                               bg * np.sqrt(np.maximum(
                                   mon.get('cov_y_pzeta', level='beam') * 0
                                   + _det(mon, 'y', 'py'), 0)), rtol=1e-9)
    if mode == 'bunch':
        assert list(mon.selected_slots) == [3, 0, 2] and list(mon.filled_slots) == [0, 2, 3, 4]
        assert np.array_equal(mon.get('num_particles', slot=0), npart[:, 1])
        assert mon.get('mean_x', slot=[2, 3], turn=5).shape == (2,)
        with pytest.raises(ValueError, match='not recorded'):
            mon.get('mean_x', slot=1)
    if mode == 'slice':
        assert mon.get('mean_zeta', level='bunch').shape == (3, 2)
        assert mon.get('mean_zeta', slot=2, slice_index=-1).shape == (3,)
        assert mon.zeta_centers.shape == (2, 5) and abs(mon.zeta_centers[0, 0] - (-0.45 - 4.0)) < 1e-12
    if mode == 'coasting':
        assert mon.coasting and mon.available_levels == ('beam', 'slice')
        with pytest.raises(ValueError, match='coasting'):
            mon.get('mean_x', slot=0)
    # configuration round trip
    mon2 = xb.BeamStatsMonitor.from_dict(mon.to_dict())
    assert mon2._data_shape == mon._data_shape and mon2.stats == mon.stats
    assert mon2._mode == mon._mode and np.array_equal(mon2._slot_to_selected, mon._slot_to_selected)


def _det(mon, cc, pp):
    return (mon.get(f'sigma_{cc}', level='beam') ** 2
            * (mon.get('cov_px_y', level='beam') * 0 + _var(mon, pp))
            - _cov(mon, cc, pp) ** 2)


def _moms(mon):
    return mon._moments_at_level('beam')


def _var(mon, cc):
    return mon._cov(cc, cc, _moms(mon))


def _cov(mon, c1, c2):
    return mon._cov(c1, c2, _moms(mon))


def test_constructor_checks():
    with pytest.raises(ValueError, match='provided together'):
        xb.BeamStatsMonitor(num_slices=3)
    with pytest.raises(ValueError, match='coasting'):
        xb.BeamStatsMonitor(coasting=True, num_slices=3, num_bunches=2)
    with pytest.raises(ValueError, match='bunch_spacing_zeta'):
        xb.BeamStatsMonitor(num_bunches=2)
    with pytest.raises(ValueError, match='unfilled'):
        xb.BeamStatsMonitor(filled_slots=[0, 2], selected_slots=[1], bunch_spacing_zeta=1.)
    with pytest.raises(ValueError, match='Unsupported statistic'):
        xb.BeamStatsMonitor(stats=['betx'])
    mon = xb.BeamStatsMonitor(start_at_turn=4)
    assert mon.stop_at_turn == 5 and mon.stats == ('num_particles', 'mean_x', 'mean_y', 'sigma_x', 'sigma_y')
    assert mon._needed_fields == {'num_particles', 'sum_beta0_gamma0', 'sum_x', 'sum_y', 'sum_x_x', 'sum_y_y'}
    mon.allocate()
    mon.start_new_frame(10)
    assert list(mon.turns) == [10] and int(mon._store['desc'][0]) == 10
