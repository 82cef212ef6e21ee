"""Inverse-CDF tables of the `quantum-kick` radiation model (radiation_flag 3).

The reference ships them as a generated C header (`xtrack/headers/synrad_total_energy_tables.h`,
made by `xtrack/headers/_generate_synrad_total_energy_tables.py`; the header itself is one of
the large blobs missing from the reference checkout).  Here they are DATA:
`data/synrad_total_energy_tables.npz`, produced by `scripts/make_synrad_tables.py`, which runs
the reference's own generator and stores what it computes; the lattice uploads them to the
device as one blob (layout: include/xtb200.h, `xtb_lattice_set_synrad_tables`).
"""
import os

import numpy as np

DATA_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data',
                         'synrad_total_energy_tables.npz')
DIRECT_TABLE_MAX = 32
POWER_TABLES = (64, 128, 256)
TABLE_COUNTS = tuple(range(1, DIRECT_TABLE_MAX + 1)) + POWER_TABLES

_blob = None


def make_blob(left_u, center_u, right_v, log_tables, tail_probability_max):
    """`log_tables[N]`: log of the quantiles of the sum of N photon energies on the
    concatenated (left u | centre u | right v) probability grid, N in TABLE_COUNTS."""
    left_u, center_u, right_v = (np.asarray(v, dtype=np.float64) for v in (left_u, center_u, right_v))
    size = len(left_u) + len(center_u) + len(right_v)
    parts = [np.array([len(left_u), len(center_u), len(right_v), tail_probability_max,
                       DIRECT_TABLE_MAX, 0., 0., 0.]), left_u, center_u, right_v]
    for nn in TABLE_COUNTS:
        tt = np.asarray(log_tables[nn], dtype=np.float64)
        if tt.shape != (size,):
            raise ValueError(f'table {nn}: {tt.shape} entries, the grids have {size}')
        parts.append(tt)
    return np.ascontiguousarray(np.concatenate(parts))


def load_blob(path=None):
    """The blob of the shipped tables (cached)."""
    global _blob
    if path is None and _blob is not None:
        return _blob
    ff = path or DATA_FILE
    if not os.path.exists(ff):
        raise FileNotFoundError(
            f'{ff} is missing: the quantum-kick radiation model needs the inverse-CDF tables '
            '(scripts/make_synrad_tables.py generates them with the reference\'s generator)')
    dd = np.load(ff)
    blob = make_blob(dd['left_u'], dd['center_u'], dd['right_v'],
                     {nn: dd[f'log_table_{nn}'] for nn in TABLE_COUNTS},
                     float(dd['tail_probability_max']))
    if path is None:
        _blob = blob
    return blob
