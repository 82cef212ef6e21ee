"""Line clean-up passes of `Line.optimize_for_tracking`.

Host-side restatement of the reference's passes, in the reference's order
(xtrack/line.py:4951-5026):

  remove_markers                 line.py:5136-5166
  remove_inactive_multipoles     line.py:5169-5216
  merge_consecutive_multipoles   line.py:5613-5683
  remove_redundant_apertures     line.py:5317-5404
  remove_zero_length_drifts      line.py:5219-5261
  merge_consecutive_drifts       line.py:5264-5314

The last two steps of the reference, `use_simple_bends` / `use_simple_quadrupoles`
(line.py:5407-5444: swap plain multipoles for `SimpleThinBend` / `SimpleThinQuadrupole`), need
no element swap here: the lattice lowering already emits the specialised ops for multipoles
with a single non-zero coefficient (csrc/xtb_ops.h "Zero-coefficient specialisation"), with or
without this method.  `_replace_with_equivalent_elements` (slices -> stand-alone elements)
has no counterpart: the loader resolves slices when the line is read.

Every pass edits `line.element_names` / `line.element_dict` in place and invalidates the
tracker; merged elements are new objects (the originals stay untouched in `element_dict`).
"""
import numpy as np

from . import elements as _el


def _is_drift(ee):
    return type(ee).__name__.startswith('Drift')


def _is_aperture(ee):
    return type(ee).__name__.startswith('Limit')


def _keep_list(keep):
    if keep is None:
        return []
    if isinstance(keep, str):
        return [keep]
    return list(keep)


def _total_knl_ksl(ee):
    """get_total_knl_ksl of a Multipole, beam_elements/_common.py:369-403 (minimum length 4;
    relative strengths scaled by the main strength)."""
    nn = max(4, len(ee.knl), len(ee.ksl), len(ee.knl_rel), len(ee.ksl_rel))
    knl = np.zeros(nn)
    ksl = np.zeros(nn)
    knl[:len(ee.knl)] += np.asarray(ee.knl, dtype=float)
    ksl[:len(ee.ksl)] += np.asarray(ee.ksl, dtype=float)
    if len(ee.knl_rel) or len(ee.ksl_rel):
        ms = ee.main_strength
        knl[:len(ee.knl_rel)] += ms * np.asarray(ee.knl_rel, dtype=float)
        ksl[:len(ee.ksl_rel)] += ms * np.asarray(ee.ksl_rel, dtype=float)
    return knl, ksl


def _trim_common_trailing_zeros(knl, ksl):
    last_nonzero = 0
    for ii, vv in enumerate(knl):
        if vv != 0:
            last_nonzero = ii
    for ii, vv in enumerate(ksl):
        if vv != 0:
            last_nonzero = max(last_nonzero, ii)
    return knl[:last_nonzero + 1], ksl[:last_nonzero + 1]


def _elements_equal(e1, e2):
    """`_apertures_equal` (line.py:7792-7803): same class, same stored fields."""
    if type(e1) is not type(e2):
        return False
    d1, d2 = vars(e1), vars(e2)
    if d1.keys() != d2.keys():
        return False
    for kk in d1:
        if not np.array_equal(np.asarray(d1[kk]), np.asarray(d2[kk])):
            return False
    return True


def remove_markers(line, keep=None):
    keep = _keep_list(keep)
    line.element_names = [nn for nn in line.element_names
                          if not (isinstance(line.element_dict[nn], _el.Marker)
                                  and type(line.element_dict[nn]) is _el.Marker
                                  and nn not in keep)]
    line._invalidate()
    return line


def remove_inactive_multipoles(line, keep=None):
    keep = _keep_list(keep)
    names = []
    for nn in line.element_names:
        ee = line.element_dict[nn]
        if (isinstance(ee, _el.Multipole) and nn not in keep
                and not (ee.isthick_now and ee.length != 0)):
            knl, ksl = _total_knl_ksl(ee)
            aux = [ee.hxl, ee.rot_x_rad, ee.rot_y_rad, *knl, *ksl]
            if np.sum(np.abs(np.array(aux))) == 0.0:
                continue
        names.append(nn)
    line.element_names = names
    line._invalidate()
    return line


def remove_zero_length_drifts(line, keep=None):
    keep = _keep_list(keep)
    line.element_names = [nn for nn in line.element_names
                          if not (_is_drift(line.element_dict[nn]) and nn not in keep
                                  and line.element_dict[nn].length == 0.0)]
    line._invalidate()
    return line


def merge_consecutive_drifts(line, keep=None):
    """Consecutive drifts become one drift of the summed length (a NEW element: the summed
    length is rounded once, as in the reference, so the merged line is not bit-identical to
    the unmerged one; it is bit-identical to the reference tracking the merged line)."""
    keep = _keep_list(keep)
    names = []
    for ii, nn in enumerate(line.element_names):
        ee = line.element_dict[nn]
        if ii > 0 and _is_drift(ee) and nn not in keep:
            prev_nn = names[-1]
            prev_ee = line.element_dict[prev_nn]
            if _is_drift(prev_ee) and prev_nn not in keep and type(prev_ee) is type(ee):
                if not prev_nn.startswith('_merged_'):       # do not touch the original
                    new_nn = f'_merged_{len(names) - 1}_{prev_nn}'   # (names may repeat)
                    kw = {'length': prev_ee.length}
                    if hasattr(prev_ee, 'model'):
                        kw['model'] = prev_ee.model
                    line.element_dict[new_nn] = type(prev_ee)(**kw)
                    names[-1] = new_nn
                    prev_ee = line.element_dict[new_nn]
                prev_ee.length += ee.length
                continue
        names.append(nn)
    line.element_names = names
    line._invalidate()
    return line


def merge_consecutive_multipoles(line, keep=None):
    keep = _keep_list(keep)
    names = []
    for nn in line.element_names:
        ee = line.element_dict[nn]
        if names and isinstance(ee, _el.Multipole) and nn not in keep and not ee.isthick_now:
            prev_nn = names[-1]
            prev_ee = line.element_dict[prev_nn]
            if (isinstance(prev_ee, _el.Multipole) and not prev_ee.isthick_now
                    and prev_ee.hxl == ee.hxl == 0
                    and not (ee.rot_x_rad != 0 or ee.rot_y_rad != 0)
                    and not (prev_ee.rot_x_rad != 0 or prev_ee.rot_y_rad != 0)
                    and prev_nn not in keep):
                prev_knl, prev_ksl = _total_knl_ksl(prev_ee)
                ee_knl, ee_ksl = _total_knl_ksl(ee)
                oo = max(len(prev_knl), len(prev_ksl), len(ee_knl), len(ee_ksl))
                knl = np.zeros(oo)
                ksl = np.zeros(oo)
                knl[:len(prev_knl)] += prev_knl
                knl[:len(ee_knl)] += ee_knl
                ksl[:len(prev_ksl)] += prev_ksl
                ksl[:len(ee_ksl)] += ee_ksl
                knl, ksl = _trim_common_trailing_zeros(knl, ksl)
                newee = _el.Multipole(knl=knl, ksl=ksl, hxl=prev_ee.hxl, length=prev_ee.length,
                                      radiation_flag=prev_ee.radiation_flag)
                new_nn = prev_nn + '_' + nn
                if new_nn in line.element_dict:                  # (names may repeat)
                    new_nn = f'{new_nn}_at{len(names) - 1}'
                line.element_dict[new_nn] = newee
                names[-1] = new_nn
                continue
        names.append(nn)
    line.element_names = names
    line._invalidate()
    return line


def remove_redundant_apertures(line, keep=None, drifts_that_need_aperture=()):
    """Of three or more equal apertures separated only by drifts and markers, the middle
    ones are removed (line.py:5317-5404)."""
    keep = _keep_list(keep)
    aper_to_remove = []
    aper_0 = aper_m1 = aper_m2 = None
    for nn in line.element_names:
        ee = line.element_dict[nn]
        if _is_aperture(ee):
            aper_m2 = aper_m1
            aper_m1 = aper_0
            aper_0 = nn
        elif ((not isinstance(ee, _el.Marker) and not _is_drift(ee))
              or nn in drifts_that_need_aperture):
            aper_0 = aper_m1 = aper_m2 = None
        if (aper_m2 is not None
                and _elements_equal(line.element_dict[aper_0], line.element_dict[aper_m1])
                and _elements_equal(line.element_dict[aper_m1], line.element_dict[aper_m2])):
            if aper_m1 not in keep:
                aper_to_remove.append(aper_m1)
                aper_m1 = aper_m2
                aper_m2 = None
    names = list(line.element_names)
    for name in aper_to_remove:
        names.remove(name)
    line.element_names = names
    line._invalidate()
    return line


def optimize_for_tracking(line, keep_markers=False, verbose=False):
    """Line.optimize_for_tracking, line.py:4951-5026 (no collective elements on this path)."""
    n0 = len(line.element_names)
    if keep_markers is False:
        remove_markers(line)
    elif keep_markers is not True:
        remove_markers(line, keep=keep_markers)
    remove_inactive_multipoles(line)
    merge_consecutive_multipoles(line)
    remove_redundant_apertures(line)
    remove_zero_length_drifts(line)
    merge_consecutive_drifts(line)
    if verbose:
        print(f'optimize_for_tracking: {n0} -> {len(line.element_names)} elements')
    return line
