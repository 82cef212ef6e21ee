"""TEST INFRASTRUCTURE: builds sliced versions of a line for the slice-element tests, laying
the slices out as the reference's slicer does (xtrack/slicing.py:342-470: `<name>..entry_map`,
drift slices `drift_<name>..i` and thin slices `<name>..i` in teapot / uniform positions, or
thick slices, `<name>..exit_map`; `slice_offset` = distance from the parent's entrance).
The slicing algorithm itself is outside the hot-path contract: the product
only has to TRACK such lines; this helper exists so that the tests have some."""
import xtrack_b200 as xb
from xtrack_b200 import elements as _el

_SLICEABLE = ('Bend', 'RBend', 'Quadrupole', 'Sextupole', 'Octupole', 'Multipole', 'Cavity',
              'CrabCavity')
_WITH_EDGES = ('Bend', 'RBend', 'Quadrupole', 'Sextupole', 'Octupole')


def teapot_drift_weights(n):
    if n == 1:
        return [0.5, 0.5]
    edge = 1. / (2 * (1 + n))
    middle = n / (n ** 2 - 1)
    return [edge] + [middle] * (n - 1) + [edge]


def uniform_drift_weights(n):
    return [1. / (n + 1)] * (n + 1)


def slice_line(line, n=4, mode='thin', scheme='teapot', only=None):
    """New Line in which every sliceable thick element (class in `only`, default all) is
    replaced by its slices; the parents stay in `element_dict` (not in `element_names`)."""
    dd = dict(line.element_dict)
    names = []
    for nn in line.element_names:
        el = dd[nn]
        if isinstance(el, _el.Replica):
            el = el.resolve(dd)
        cname = type(el).__name__
        thick = cname in _SLICEABLE and (cname != 'Multipole' or el._isthick_field > 0)
        if not thick or (only is not None and cname not in only):
            names.append(nn)
            continue
        pname = nn if not isinstance(dd[nn], _el.Replica) else dd[nn].resolve(dd, get_name=True)
        seq = []

        def add(key, cls_name, **kw):
            dd[key] = _el.SLICE_CLASSES[cls_name](parent_name=pname, _parent=el, **kw)
            seq.append(key)

        if cname in _WITH_EDGES:
            add(f'{nn}..entry_map', f'ThinSlice{cname}Entry', slice_offset=0.0)
        offset = 0.0
        if mode == 'thin':
            dw = teapot_drift_weights(n) if scheme == 'teapot' else uniform_drift_weights(n)
            for ii, ww in enumerate(dw):
                add(f'drift_{nn}..{ii}', f'DriftSlice{cname}', weight=ww, slice_offset=offset)
                offset += el.length * ww
                if ii < n:
                    add(f'{nn}..{ii}', f'ThinSlice{cname}', weight=1. / n, slice_offset=offset)
        else:
            for ii in range(n):
                add(f'{nn}..{ii}', f'ThickSlice{cname}', weight=1. / n, slice_offset=offset)
                offset += el.length * (1. / n)
        if cname in _WITH_EDGES:
            add(f'{nn}..exit_map', f'ThinSlice{cname}Exit', slice_offset=el.length)
        names += seq
    out = xb.Line(elements=dd, element_names=names, particle_ref=line.particle_ref)
    out.config.update(line.config)
    out._extra_config.update(line._extra_config)
    return out
