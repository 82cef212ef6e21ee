"""The counter-based production generator (north star: "a counter-based RNG for radiation";
csrc/xtb_rng.cuh Philox4x32-10): known answers, an independent numpy restatement, stream
properties that the design relies on (independence of slot / sharding / launch splitting), and
radiation statistics equal to the reference generator's within their statistical error."""
import ctypes as ct

import numpy as np
import pytest

import xtrack_b200 as xb
import common
import hostsim

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xffffffff


def philox_numpy(k0, k1, c0, c1, c2=0, c3=0):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
    SC'11), written from the paper's round function, plain Python integers."""
    c = [c0, c1, c2, c3]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & MASK, (p0 >> 32) ^ c[3] ^ k1, p0 & MASK]
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c


def _host_block(k0, k1, c0, c1):
    out = (ct.c_uint32 * 4)()
    hostsim.load().xtb_hostsim_philox(k0, k1, c0, c1, out)
    return list(out)


def test_known_answers_and_host_build():
    # Random123 known-answer vectors for philox4x32-10 (counter words 2, 3 are zero here only
    # in the first case; the others pin the numpy restatement, which takes all four)
    assert philox_numpy(0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox_numpy(MASK, MASK, MASK, MASK, MASK, MASK) == [0x408f276d, 0x41c83b0e,
                                                                 0xa20bc7c6, 0x6d5451fd]
    assert philox_numpy(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e,
                        0x03707344) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    rng = np.random.default_rng(1)
    for _ in range(200):
        k0, k1, c0, c1 = (int(v) for v in rng.integers(0, 2 ** 32, 4))
        assert _host_block(k0, k1, c0, c1) == philox_numpy(k0, k1, c0, c1)


@pytest.mark.gpu
def test_device_blocks_equal_numpy():
    from xtrack_b200 import _cabi
    got = _cabi.eval_philox(0x12345678, 0x9abcdef0, 0xfffffff0, 7, 64)
    for ii in range(64):
        assert list(got[ii]) == philox_numpy(0x12345678, 0x9abcdef0, (0xfffffff0 + ii) & MASK, 7)
    assert list(_cabi.eval_philox(0, 0, 0, 0, 1)[0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c,
                                                         0x9b00dbd8]


def _bend_line():
    line = xb.Line(elements=[xb.Bend(length=1.0, angle=0.02, k0='from_h',
                                     edge_entry_active=0, edge_exit_active=0)])
    line.particle_ref = xb.Particles(p0c=5e9, mass0=xb.ELECTRON_MASS_EV)
    line.configure_radiation(model='quantum')
    return line


BACKENDS = [pytest.param(False, id='hostsim'), pytest.param(True, id='gpu', marks=pytest.mark.gpu)]


def _track(line, p, on_gpu, num_turns=1, **kw):
    if on_gpu:
        line.build_tracker(_device='cuda:0', **kw)
        q = p.copy(_device='cuda:0')
    else:
        line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker, **kw)
        q = p.copy()
    line.track(q, num_turns=num_turns)
    return q


@pytest.mark.parametrize('on_gpu', BACKENDS)
def test_philox_streams_and_statistics(on_gpu):
    """One 2 T-class bend, quantum model (tests/test_radiation.py:27-118 of the reference):
      * the draw counter advances, the key stays; the first draws are the numpy values;
      * same key -> same stream whatever the slot, the beam size or the launch splitting;
      * the mean energy loss equals the reference generator's within the statistical error,
        and the analytic classical loss within the quantum-model tolerance of the reference
        test (5e-3)."""
    n = 20000 if on_gpu else 3000
    line = _bend_line()
    p0 = xb.Particles(p0c=5e9, mass0=xb.ELECTRON_MASS_EV, x=np.zeros(n))
    seeds = np.arange(1, n + 1, dtype=np.uint32) * 7 + 1
    pa = p0.copy()
    pa._init_random_number_generator(seeds=seeds, mode='philox')
    assert np.array_equal(pa.get('_rng_s1'), seeds) and np.all(pa.get('_rng_s3') == 0)
    qa = _track(line, pa, on_gpu, num_turns=3)
    assert np.array_equal(qa.get('_rng_s1'), seeds)
    assert np.array_equal(qa.get('_rng_s2'), np.arange(n, dtype=np.uint32))
    draws = qa.get('_rng_s3').astype(np.int64)
    assert draws.min() >= 3 and draws.max() < 20000 and np.all(qa.get('_rng_s4') == 0)
    # launch splitting: 3 turns == 1 + 2 turns (the counter lives in the particle SoA)
    qb = _track(line, pa, on_gpu, num_turns=1)
    line.track(qb, num_turns=2)
    for ff in ('delta', 'px', '_rng_s3'):
        assert np.array_equal(qa.get(ff), qb.get(ff)), ff
    # slot / beam-size independence: a sub-beam in reversed slot order, same keys
    sel = np.arange(0, n, 7)[::-1].copy()
    pc = xb.Particles(p0c=5e9, mass0=xb.ELECTRON_MASS_EV, x=np.zeros(len(sel)), particle_id=sel)
    pc._init_random_number_generator(seeds=seeds[sel], mode='philox')
    qc = common.by_id(_track(line, pc, on_gpu, num_turns=3))
    ga = common.by_id(qa)
    for ff in ('delta', 'px', 'zeta', '_rng_s3'):
        assert np.array_equal(qc[ff], ga[ff][np.sort(sel)]), ff
    # statistics against the reference's generator on the same beam
    pt = p0.copy()
    if on_gpu:
        pt = pt.copy(_device='cuda:0')
        pt._init_random_number_generator(seeds=seeds, mode='tausworthe')
        pt = pt.copy(_device='cpu')
    else:
        common.seed_rng_host(pt, seeds)
    qt = _track(line, pt, on_gpu, num_turns=3)
    da, dt = qa.get('delta'), qt.get('delta')
    err = np.sqrt(da.var() / n + dt.var() / n)
    assert abs(da.mean() - dt.mean()) < 5 * err, (da.mean(), dt.mean(), err)
    assert abs(da.std() / dt.std() - 1) < 6 / np.sqrt(n) + 0.02
    mean_line = _bend_line()
    mean_line.configure_radiation(model='mean')
    qm = _track(mean_line, p0, on_gpu, num_turns=3)
    np.testing.assert_allclose(da.mean(), qm.get('delta')[0], rtol=5 * err / abs(da.mean()) + 5e-3)


def test_tracker_seeds_in_its_own_mode():
    line = _bend_line()
    p = xb.Particles(p0c=5e9, mass0=xb.ELECTRON_MASS_EV, x=np.zeros(50))
    q = _track(line, p, False, rng='philox')
    assert q._rng_mode == 'philox' and np.all(q.get('_rng_s3') > 0) and np.all(q.get('_rng_s4') == 0)
    with pytest.raises(ValueError):
        line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker, rng='mt19937')
