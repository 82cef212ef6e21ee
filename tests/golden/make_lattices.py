"""Generates the lattice fixtures under tests/golden/lattices/ from the reference's
own test data (run HERE, where /root/reference exists; the GPU box reads only the
committed fixtures).  Each fixture is the `elements` / `element_names` /
`particle_ref` part of the reference JSON, unchanged except:

  * elements not referenced by `element_names`, and the xdeps sections
    (`_var_manager`, `_var_management_data`, ...), are dropped;
  * lep: the cavity voltages, which the reference file drives through the knob
    `vrfc231` (deferred expression "((1.0 * vars['vrfc231']) * 1000000.0)",
    all knobs stored as 0), are resolved with vrfc231 = 12.65 [MV], the value the
    reference's own scripts set (examples/spin_lep/002a_monte_carlo_polarization.py:14).

Usage:  python tests/golden/make_lattices.py
"""
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
OUT = os.path.join(HERE, 'lattices')

SOURCES = {
    'hllhc_14': 'test_data/hllhc_14/line_and_particle.json',
    'sps': 'test_data/sps_w_spacecharge/line_no_spacecharge_and_particle.json',
    'lep': 'test_data/lep/lep.json',
    'clic_dr': 'test_data/clic_dr/line.json',
}


def slim(name, dd):
    line = dd['line'] if 'line' in dd else dd
    used = set(line['element_names'])
    els = line['elements']
    out = {'__class__': 'Line',
           'elements': {k: v for k, v in els.items() if k in used},
           'element_names': line['element_names'],
           'particle_ref': line.get('particle_ref'),
           'source': SOURCES[name]}
    if 'particle' in dd:
        out['particle'] = dd['particle']
    if name == 'lep':
        n = 0
        for target, expr in line['_var_manager']:
            if target.endswith('.voltage') and "vars['vrfc231']" in expr:
                elname = target[len("element_refs['"):target.index("']")]
                out['elements'][elname]['voltage'] = (1.0 * 12.65) * 1000000.0
                n += 1
        out['resolved_knobs'] = {'vrfc231': 12.65, 'n_cavities': n}
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, rel in SOURCES.items():
        with open(os.path.join(REF, rel)) as fid:
            dd = json.load(fid)
        data = json.dumps(slim(name, dd), separators=(',', ':')).encode()
        with open(os.path.join(OUT, name + '.json.gz'), 'wb') as fid:
            fid.write(gzip.compress(data, 9, mtime=0))
        print(name, len(data), '->', os.path.getsize(os.path.join(OUT, name + '.json.gz')))


if __name__ == '__main__':
    main()
