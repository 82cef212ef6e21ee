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
