#!/usr/bin/env python
"""Makes xtrack_b200/data/synrad_total_energy_tables.npz: the inverse-CDF tables of the
`quantum-kick` radiation model.

The reference generates them offline with xtrack/headers/_generate_synrad_total_energy_tables.py
into a C header that is not part of the reference checkout (.MISSING_LARGE_BLOBS).  This script
RUNS THAT GENERATOR where it lies (nothing of it is copied; it is executed with its output path
pointed at a scratch directory -- about half an hour on one core), reads the numbers out of the
header it writes and stores them as float64 arrays.  Needs /root/reference (or
XTB_REFERENCE_ROOT) and scipy.

    python scripts/make_synrad_tables.py [--header already_generated.h]
"""
import argparse
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xtrack_b200 import synrad_tables      # noqa: E402

REF = os.environ.get('XTB_REFERENCE_ROOT', '/root/reference')
GENERATOR = os.path.join(REF, 'xtrack', 'headers', '_generate_synrad_total_energy_tables.py')


def run_generator(workdir):
    scope = {'__name__': 'synrad_table_generator', '__file__': os.path.join(workdir, '_generate.py')}
    exec(compile(open(GENERATOR).read(), GENERATOR, 'exec'), scope)
    scope['main']()
    return os.path.join(workdir, 'synrad_total_energy_tables.h')


def parse_header(path):
    text = open(path).read()
    out = {}
    for mm in re.finditer(r'(synrad_total_energy_\w+)\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}', text):
        out[mm.group(1)] = np.array([float(v) for v in mm.group(2).replace('\n', ' ').split(',')
                                     if v.strip()])
    defs = dict(re.findall(r'#define\s+(XTRACK_SYNRAD_TOTAL_ENERGY_\w+)\s+(\S+)', text))
    return out, defs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--header', help='a header the generator already wrote')
    args = ap.parse_args()
    header = args.header or run_generator(tempfile.mkdtemp(prefix='synrad_tables_'))
    arrays, defs = parse_header(header)
    left_u = arrays['synrad_total_energy_left_u_grid']
    center_u = arrays['synrad_total_energy_center_u_grid']
    right_v = arrays['synrad_total_energy_right_v_grid']
    assert len(left_u) == int(defs['XTRACK_SYNRAD_TOTAL_ENERGY_LEFT_SIZE'])
    assert len(center_u) == int(defs['XTRACK_SYNRAD_TOTAL_ENERGY_CENTER_SIZE'])
    assert len(right_v) == int(defs['XTRACK_SYNRAD_TOTAL_ENERGY_RIGHT_SIZE'])
    assert int(defs['XTRACK_SYNRAD_TOTAL_ENERGY_DIRECT_TABLE_MAX']) == synrad_tables.DIRECT_TABLE_MAX
    tail_max = float(defs['XTRACK_SYNRAD_TOTAL_ENERGY_TAIL_PROBABILITY_MAX'])
    size = len(left_u) + len(center_u) + len(right_v)
    save = dict(left_u=left_u, center_u=center_u, right_v=right_v,
                tail_probability_max=np.float64(tail_max))
    for nn in synrad_tables.TABLE_COUNTS:
        tt = arrays[f'synrad_total_energy_log_table_{nn}']
        assert tt.shape == (size,), (nn, tt.shape)
        save[f'log_table_{nn}'] = tt
    os.makedirs(os.path.dirname(synrad_tables.DATA_FILE), exist_ok=True)
    np.savez_compressed(synrad_tables.DATA_FILE, **save)
    print(synrad_tables.DATA_FILE, os.path.getsize(synrad_tables.DATA_FILE), 'bytes;',
          len(synrad_tables.TABLE_COUNTS), 'tables of', size, 'quantiles')


if __name__ == '__main__':
    main()
