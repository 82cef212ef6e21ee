"""`configure_radiation(model='quantum-kick')` (radiation_flag 3; synrad_spectrum.h:220-459,
track_magnet_radiation.h:261): the photon COUNT of a slice is drawn, the total energy radiated
comes from tabulated inverse CDFs of the sum of N photon energies.  The reference's code for
this model, compiled from where it lies, runs on the same table blob as the product (oracle
shim: the generated header's arrays are run-time pointers), so the host build of the device
code must reproduce it bit for bit, generator state included; the B200 is held to the
last-bit differences of `log` / `exp`.  A synthetic table set exercises every branch (left
tail, centre, right tail, direct tables, power-of-two chunks); the shipped tables
(xtrack_b200/data, made by the reference's generator) are checked for their physics.
"""
import numpy as np
import pytest

import xtrack_b200 as xb
from xtrack_b200 import synrad_tables
import common
import ref_oracle as ro
from test_rows_both_tiers import BACKENDS, _build


def synthetic_blob(n_tail=40, n_center=61, tail_max=9.8e-2):
    u_tail = np.logspace(-15, np.log10(tail_max), n_tail)
    left_u = np.concatenate(([0.0], u_tail))
    right_v = left_u.copy()
    center_u = np.linspace(tail_max, 1 - tail_max, n_center)
    tables = {}
    for nn in synrad_tables.TABLE_COUNTS:
        # a smooth increasing quantile function whose scale grows with N (as the true one)
        q = lambda u: nn * (0.32 * (-np.log1p(-np.clip(u, 0, 1 - 1e-16))) ** 0.9 + 1e-6 * u ** (1 / 3.))
        x_left = np.maximum(q(left_u), 1e-30)
        x_center = q(center_u)
        x_right = q(1 - right_v)
        tables[nn] = np.log(np.concatenate([x_left, x_center, x_right]))
    return synrad_tables.make_blob(left_u, center_u, right_v, tables, tail_max)


def _ring(name):
    line = common.load_line(name)
    line.configure_radiation(model='quantum-kick')
    return line


@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('name', ['clic_dr', 'lep'])
def test_quantum_kick_vs_reference_code(name, on_gpu):
    line = _ring(name)
    blob = synthetic_blob()
    line.synrad_tables = blob
    n = 48
    p_host = common.gaussian_particles(line, n, 5, common.SIGMAS[name])
    common.seed_rng_host(p_host, np.arange(1, n + 1, dtype=np.uint32) * 7919)
    ro.set_synrad_tables(blob, 'synrad')
    try:
        ref = common.oracle_track(line, p_host, 5, variant='synrad')
    finally:
        ro.set_synrad_tables(None, 'synrad')
    dev = _build(line, on_gpu)
    p = p_host.copy(_device=dev)
    line.track(p, num_turns=5)
    got = common.by_id(p)
    assert ref['delta'].mean() < -5e-4            # it radiated
    assert not np.array_equal(ref['_rng_s1'], p_host.get('_rng_s1'))
    for ff in ('state', 'at_turn', 'at_element'):
        assert np.array_equal(got[ff], ref[ff]), ff
    if not on_gpu:
        for ff in common.ALL_F64 + xb.particles.U32_VARS:
            assert np.array_equal(got[ff], ref[ff]), ff
    else:
        # same random stream for every particle, coordinates to the libm's last bits
        same = np.ones(n, dtype=bool)
        for ff in xb.particles.U32_VARS:
            same &= got[ff] == ref[ff]
        assert same.mean() >= 0.95
        for ff in ('x', 'px', 'y', 'py', 'zeta', 'delta'):
            scale = np.max(np.abs(ref[ff]))
            assert np.max(np.abs(got[ff][same] - ref[ff][same])) <= 1e-9 * scale, ff


def test_missing_tables_fail_loudly(monkeypatch):
    import hostsim
    line = _ring('clic_dr')
    monkeypatch.setattr(synrad_tables, 'DATA_FILE', '/nonexistent/tables.npz')
    monkeypatch.setattr(synrad_tables, '_blob', None)
    with pytest.raises(FileNotFoundError, match='quantum-kick'):
        line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker)


def test_shipped_tables_are_the_sum_of_n_photon_energies():
    """xtrack_b200/data/synrad_total_energy_tables.npz (made by the reference's generator,
    scripts/make_synrad_tables.py): quantile functions of X_N = x_1 + ... + x_N, x = E / E_c
    distributed as the synchrotron-radiation photon spectrum, for which <x> = 8 / (15 sqrt 3)
    and <x^2> = 11 / 27."""
    blob = synrad_tables.load_blob()
    n_l, n_c, n_r = (int(v) for v in blob[:3])
    assert (n_l, n_c, n_r) == (2296, 1601, 2296) and blob[4] == 32
    assert len(blob) == 8 + (n_l + n_c + n_r) * (1 + len(synrad_tables.TABLE_COUNTS))
    left_u, center_u, right_v = blob[8:8 + n_l], blob[8 + n_l:8 + n_l + n_c], blob[8 + n_l + n_c:8 + n_l + n_c + n_r]
    assert left_u[0] == 0 and np.all(np.diff(left_u) > 0) and np.all(np.diff(center_u) > 0)
    assert abs(left_u[-1] - blob[3]) < 1e-15 and abs(center_u[0] - blob[3]) < 1e-15
    size = n_l + n_c + n_r
    mean1, var1 = 8 / (15 * np.sqrt(3)), 11 / 27 - (8 / (15 * np.sqrt(3))) ** 2
    for kk, nn in enumerate(synrad_tables.TABLE_COUNTS):
        tt = blob[8 + size * (1 + kk):8 + size * (2 + kk)]
        ql, qc, qr = np.exp(tt[:n_l]), np.exp(tt[n_l:n_l + n_c]), np.exp(tt[n_l + n_c:])
        assert np.all(np.diff(ql) >= 0) and np.all(np.diff(qc) > 0) and np.all(np.diff(qr) <= 0), nn
        assert ql[-1] <= qc[0] * (1 + 1e-9) and qc[-1] <= qr[-1] * (1 + 1e-9), nn
        m1 = np.trapezoid(ql, left_u) + np.trapezoid(qc, center_u) + np.trapezoid(qr, right_v)
        m2 = (np.trapezoid(ql ** 2, left_u) + np.trapezoid(qc ** 2, center_u)
              + np.trapezoid(qr ** 2, right_v))
        assert abs(m1 / (nn * mean1) - 1) < 2e-3, (nn, m1)
        assert abs((m2 - m1 ** 2) / (nn * var1) - 1) < 2e-2, (nn, m2 - m1 ** 2)


def test_quantum_kick_energy_loss_statistics_vs_photon_by_photon():
    """The same ring under `quantum` (photon by photon) and `quantum-kick` (shipped tables):
    mean and spread of the energy lost in two turns agree within the statistics of the sample."""
    import hostsim
    n = 500
    out = {}
    for model in ('quantum', 'quantum-kick'):
        line = common.load_line('clic_dr')
        line.configure_radiation(model=model)
        p = common.gaussian_particles(line, n, 5, common.SIGMAS['clic_dr'], scale=0.1)
        common.seed_rng_host(p, np.arange(1, n + 1, dtype=np.uint32) * 104729)
        e0 = p.get('ptau').copy()
        line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker)
        line.track(p, num_turns=2)
        assert np.all(p.get('state') > 0)
        out[model] = (p.get('ptau') - e0) * p.get('p0c')          # eV
    mq, mk = out['quantum'].mean(), out['quantum-kick'].mean()
    sq, sk = out['quantum'].std(), out['quantum-kick'].std()
    assert mq < -5e6                                   # ~ 4 MeV per turn
    assert abs(mk / mq - 1) < 4 * sq / abs(mq) / np.sqrt(n) + 5e-3, (mk, mq)
    assert abs(sk / sq - 1) < 0.15, (sk, sq)
