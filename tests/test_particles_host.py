"""Host-side Particles helpers kept from xtrack's API (xtrack/particles/particles.py:1002-1130,
1280-1330, 1450-1560, 1600-1900): consistent energy updates, derived quantities, subsets."""
import numpy as np
import pytest

import xtrack_b200 as xb


def _p(n=6):
    return xb.Particles(p0c=7e12, x=np.linspace(-1e-3, 1e-3, n), px=1e-5, py=-2e-5,
                        delta=np.linspace(-1e-3, 1e-3, n), zeta=0.01)


def test_update_delta_and_ptau_are_consistent():
    p = _p()
    ref = xb.Particles(p0c=7e12, delta=np.linspace(2e-4, 5e-4, 6))
    new = np.linspace(2e-4, 5e-4, 6)
    new_with_nan = new.copy()
    new_with_nan[2] = np.nan
    st = p.get('state').copy()
    st[4] = 0
    p.state = st
    old = {ff: p.get(ff).copy() for ff in ('delta', 'ptau', 'rvv', 'rpp')}
    p.update_delta(new_with_nan)
    for ff in ('delta', 'ptau', 'rvv', 'rpp'):
        got = p.get(ff)
        for ii in range(6):
            if ii in (2, 4):       # NaN entry / lost particle: untouched
                assert got[ii] == old[ff][ii], (ff, ii)
            else:
                assert got[ii] == ref.get(ff)[ii], (ff, ii)
    # ptau -> delta round trip
    q = _p()
    q.update_ptau(ref.get('ptau'))
    np.testing.assert_allclose(q.get('delta'), ref.get('delta'), rtol=1e-12, atol=1e-18)
    np.testing.assert_array_equal(q.get('ptau'), ref.get('ptau'))


def test_update_p0c_and_derived_quantities():
    p = _p()
    p.update_p0c(6.5e12)
    ref = xb.Particles(p0c=6.5e12)
    assert np.all(p.get('p0c') == 6.5e12)
    np.testing.assert_array_equal(p.get('beta0'), np.full(6, ref.get('beta0')[0]))
    np.testing.assert_array_equal(p.get('gamma0'), np.full(6, ref.get('gamma0')[0]))
    e = p.energy.numpy()
    np.testing.assert_allclose(e, np.sqrt(6.5e12 ** 2 + p.mass0 ** 2) + p.get('ptau') * 6.5e12,
                               rtol=1e-15)
    np.testing.assert_allclose(p.rigidity0.numpy(), 6.5e12 / 299792458.0, rtol=1e-15)
    kps = p.kin_ps.numpy()
    np.testing.assert_allclose(kps, np.sqrt((1 + p.get('delta')) ** 2 - 1e-10 - 4e-10), rtol=1e-15)
    np.testing.assert_allclose(p.kin_xprime.numpy(), 1e-5 / kps, rtol=1e-15)


def test_filter_merge_add_particles():
    p = _p(8)
    sub = p.filter(p.get('x') > 0)
    assert sub._capacity == 4 and np.all(sub.get('x') > 0)
    assert np.array_equal(sub.get('particle_id'), np.array([4, 5, 6, 7]))
    big = xb.Particles(p0c=7e12, x=np.arange(3) * 1e-3, _capacity=10)
    assert big.remove_unused_space()._capacity == 3
    m = xb.Particles.merge([p, big])
    assert m._capacity == 11
    ids = m.get('particle_id')
    assert len(set(ids.tolist())) == 11 and np.array_equal(ids[:8], np.arange(8))
    assert np.array_equal(m.get('x')[8:], np.arange(3) * 1e-3)
    st = big.get('state').copy()
    st[1] = 0
    big.state = st
    p.add_particles(big)                       # lost particles are not added
    assert p._capacity == 10 and np.sum(p.get('state') > 0) == 10
    with pytest.raises(ValueError):
        xb.Particles.merge([p, xb.Particles(p0c=7e12, mass0=xb.ELECTRON_MASS_EV)])


def test_pandas_round_trip():
    p = _p(5)
    df = p.to_pandas()
    assert len(df) == 5 and 'x' in df and 'particle_id' in df
    q = xb.Particles.from_pandas(df)
    for ff in ('x', 'px', 'delta', 'ptau', 'rvv', 'rpp', 'zeta', 'particle_id', 'state', 'p0c'):
        assert np.array_equal(q.get(ff), p.get(ff)), ff
    assert q.mass0 == p.mass0 and q.q0 == p.q0
