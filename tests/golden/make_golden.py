"""Generates the golden fixtures under tests/golden/ (run HERE, where
/root/reference exists; the fixtures and this script are committed, the GPU box
only reads the fixtures).

Sources of truth, in order:
  1. `ducktrack` -- the reference's own independent numpy implementation, used
     by the reference's tests/test_full_rings.py:24-118 to pin 10-turn results.
     Imported from /root/reference with two stub modules (`xobjects.JEncoder`,
     empty `xtrack`) because xobjects is not installed.
  2. the reference's own C headers compiled through the oracle shim
     (oracle/_ref/libxt_ref_serial.so).

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def import_ducktrack():
    xo = types.ModuleType('xobjects')

    class JEncoder(json.JSONEncoder):
        def default(self, obj):
            if isinstance(obj, np.ndarray):
                return obj.tolist()
            return json.JSONEncoder.default(self, obj)
    xo.JEncoder = JEncoder
    sys.modules['xobjects'] = xo
    sys.modules['xtrack'] = types.ModuleType('xtrack')
    sys.path.insert(0, REF)
    import ducktrack
    sys.path.remove(REF)
    return ducktrack


def strip_unsupported(line_dct, keep):
    """Replace element classes outside `keep` by zero-length drifts (both codes
    then see the same lattice)."""
    out = dict(line_dct)
    els = {}
    for nn, ee in line_dct['elements'].items():
        if ee['__class__'] in keep:
            els[nn] = ee
        else:
            els[nn] = {'__class__': 'Drift', 'length': 0.0}
    out['elements'] = els
    return out


def ducktrack_full_rings():
    dtk = import_ducktrack()
    import xtrack_b200 as xb
    out = {}
    cases = {
        'hllhc_14': ('test_data/hllhc_14/line_and_particle.json', (1e-9, 3e-11)),
        'sps': ('test_data/sps_w_spacecharge/line_no_spacecharge_and_particle.json',
                (2e-8, 7e-9)),
    }
    keep = set(xb.elements.ELEMENT_CLASSES)
    for name, (fname, tol) in cases.items():
        with open(os.path.join(REF, fname)) as fid:
            dd = json.load(fid)
        ldct = strip_unsupported(dd['line'], keep)
        used = set(ldct['element_names'])
        ldct_dtk = dict(ldct)
        # ducktrack wants a list of elements in line order
        ldct_dtk['elements'] = [ldct['elements'][nn] for nn in ldct['element_names']]
        testline = dtk.TestLine.from_dict(ldct_dtk)
        part = dtk.TestParticles.from_dict(dd['particle']).copy()
        for _ in range(10):
            testline.track(part)
        out[name] = {
            'source': f'ducktrack 10 turns, {fname} (tests/test_full_rings.py:24-118)',
            'rtol': tol[0], 'atol': tol[1],
            **{vv: float(np.atleast_1d(getattr(part, vv))[0])
               for vv in ('x', 'px', 'y', 'py', 'zeta', 'delta', 's')}}
        print(name, out[name])
    with open(os.path.join(HERE, 'ducktrack_full_rings.json'), 'w') as fid:
        json.dump(out, fid, indent=1)


if __name__ == '__main__':
    ducktrack_full_rings()
