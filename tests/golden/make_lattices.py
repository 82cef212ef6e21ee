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


def ring_fixtures():
    """examples/lattice_design/ring.json: a line of `Replica`s of thick elements, whose element
    dictionary also holds the slices that `cut_at_s` made of every element of the first arc
    cells (ThickSliceBend / ThinSliceBendEntry / Exit / ThickSliceQuadrupole / ThickSliceSextupole
    / DriftSlice: `<name>..entry_map`, `<name>..0`, ..., `<name>..exit_map`).
      ring         the line as stored;
      ring_sliced  the same line with every element that has a complete set of slices in the
                   dictionary replaced by them (weights add up to one)."""
    with open(os.path.join(REF, 'examples/lattice_design/ring.json')) as fid:
        dd = json.load(fid)
    els, names = dd['elements'], dd['element_names']

    def closure(used):
        used = set(used)
        todo = list(used)
        while todo:
            ed = els[todo.pop()]
            pn = ed.get('parent_name')
            if pn is not None and pn not in used:
                used.add(pn)
                todo.append(pn)
        return used

    def pack(nn, src):
        used = closure(nn)
        return {'__class__': 'Line', 'elements': {k: v for k, v in els.items() if k in used},
                'element_names': nn, 'particle_ref': dd['particle_ref'], 'source': src}

    out = {'ring': pack(names, 'examples/lattice_design/ring.json')}
    by_prefix = {}
    for kk in els:
        if '..' in kk:
            by_prefix.setdefault(kk.rsplit('..', 1)[0], []).append(kk)
    sliced, n_sliced = [], 0
    for nn in names:
        parts = by_prefix.get(nn)
        if not parts:
            sliced.append(nn)
            continue
        body = sorted((kk for kk in parts if kk.rsplit('..', 1)[1].isdigit()),
                      key=lambda kk: int(kk.rsplit('..', 1)[1]))
        wsum = sum(els[kk].get('weight', 0.0) for kk in body)
        if abs(wsum - 1.0) > 1e-9:
            sliced.append(nn)
            continue
        seq = ([nn + '..entry_map'] if nn + '..entry_map' in els else []) + body + (
            [nn + '..exit_map'] if nn + '..exit_map' in els else [])
        sliced += seq
        n_sliced += 1
    out['ring_sliced'] = pack(sliced, 'examples/lattice_design/ring.json (elements replaced by '
                                      'their slices from the same file)')
    out['ring_sliced']['n_elements_sliced'] = n_sliced
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, data in ring_fixtures().items():
        raw = json.dumps(data, separators=(',', ':')).encode()
        with open(os.path.join(OUT, name + '.json.gz'), 'wb') as fid:
            fid.write(gzip.compress(raw, 9, mtime=0))
        print(name, len(data['element_names']), 'elements', len(raw), '->',
              os.path.getsize(os.path.join(OUT, name + '.json.gz')))
    for name, rel in SOURCES.items():
        with open(os.path.join(REF, rel)) as fid:
            dd = json.load(fid)
        data = json.dumps(slim(name, dd), separators=(',', ':')).encode()
        with open(os.path.join(OUT, name + '.json.gz'), 'wb') as fid:
            fid.write(gzip.compress(data, 9, mtime=0))
        print(name, len(data), '->', os.path.getsize(os.path.join(OUT, name + '.json.gz')))


if __name__ == '__main__':
    main()
