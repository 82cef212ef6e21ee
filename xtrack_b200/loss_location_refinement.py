"""Loss-location refinement: where, between two aperture markers, a
lost particle actually hit the vacuum chamber.

Same procedure and public names as `xtrack.LossLocationRefinement`
(loss_location_refinement/loss_location_refinement.py:29-658 of the reference): for every
aperture that stopped particles, the chamber between it and the previous aperture is modelled
by apertures interpolated every `ds` metres -- copies when the two are identical and no
shift / rotation lies between them (:161-172), otherwise convex polygons interpolated between
the two aperture shapes as the beam sees them through the transformations around them
(:174-183, 418-463, 570-658) -- the lost particles are BACKTRACKED to the previous aperture
(:349-353) and tracked again through the refined stretch (:355); `s`, the coordinates and the
energy variables of the particle at the refined loss point replace the coarse ones (:364-380).

What differs from the reference is where it runs: the reference asserts a CPU context
(:76-78); here the backtracking, the aperture characterisation (n_theta x ~300 probe
particles per aperture) and the re-tracking are launches of the CUDA kernel on the tracker's
device.  Interpolated apertures that fall inside a thick element cut it into thick slices of
that element (the reference's `Line.insert`), with the element's edges kept at the ends.
"""
import logging

import numpy as np

from . import elements as _el
from .particles import Particles

logger = logging.getLogger(__name__)


# ---- predicates of line.py:7734-7761 ------------------------------------------------------
def _resolve(ee, line):
    return ee.resolve(line.element_dict) if isinstance(ee, _el.Replica) else ee


def _is_aperture(ee, line):
    return type(_resolve(ee, line)).__name__.startswith('Limit')


def _is_thick(ee, line):
    return bool(getattr(_resolve(ee, line), 'isthick', False))


def _allow_loss_refinement(ee, line):
    return bool(getattr(_resolve(ee, line), 'allow_loss_refinement', False))


def _has_backtrack(ee, line):
    return bool(getattr(_resolve(ee, line), 'has_backtrack', False))


def _skip_in_loss_location_refinement(ee, line):
    return bool(getattr(_resolve(ee, line), 'skip_in_loss_location_refinement', False))


def _element_s_locations(line):
    """s at the entry of every element (tracker_data `element_s_locations`)."""
    ss, out = 0.0, []
    for ee in line.elements:
        out.append(ss)
        if ee.isthick_now:
            ss += ee.length
    return np.array(out)


class _preserve_track_flags:
    """line.py `_preserve_track_flags`: the flags of the line as they were, afterwards."""

    def __init__(self, line):
        self.line = line

    def __enter__(self):
        self.saved = dict(self.line.track_flags)

    def __exit__(self, *exc):
        self.line.track_flags.clear()
        self.line.track_flags.update(self.saved)


class LossLocationRefinement:
    """Refines the location of the lost particles within a line.

    Parameters as in the reference (:29-60): `n_theta` angles and radial step `dr` of the
    aperture characterisation, `r_max` a radius larger than every aperture, `ds` the spacing
    of the interpolated apertures, `save_refine_lines` keeps the refined stretches
    (`refine_lines[i_aperture]`), `allowed_backtrack_types` element classes to backtrack
    through although they do not declare `allow_loss_refinement`."""

    def __init__(self, line, backtrack_line=None, n_theta=None, r_max=None, dr=None, ds=None,
                 save_refine_lines=False, allowed_backtrack_types=()):
        if backtrack_line is not None:
            raise ValueError('Backtracking line not supported anymore!')
        if line.tracker is None:
            raise ValueError('the line needs a tracker (line.build_tracker)')
        self.line = line
        self._original_line = line
        self.i_apertures, self.apertures = find_apertures(line)
        self.save_refine_lines = save_refine_lines
        if save_refine_lines:
            self.refine_lines = {}
        self.n_theta = n_theta
        self.r_max = r_max
        self.dr = dr
        self.ds = ds
        self.allowed_backtrack_types = tuple(allowed_backtrack_types)

    def refine_loss_location(self, particles, i_apertures=None, with_progress=True):
        """Refines, in place, the lost particles of `particles` (state 0) that stopped at the
        apertures `i_apertures` (all apertures of the line by default)."""
        if i_apertures is None:
            i_apertures = self.i_apertures
        line = self.line
        state = particles.get('state')
        at_element = particles.get('at_element')
        for i_ap in i_apertures:
            if not np.any((at_element == i_ap) & (state == 0)):
                continue
            if self.i_apertures.index(i_ap) == 0:
                logger.warning('Unable to handle the first aperture in the line')
                continue
            i_aper_1 = i_ap
            i_aper_0 = self.i_apertures[self.i_apertures.index(i_ap) - 1]
            for ii in range(i_aper_0, i_aper_1):
                if _skip_in_loss_location_refinement(line[ii], line):
                    return
            s0, s1, _ = generate_interp_aperture_locations(line, i_aper_0, i_aper_1, self.ds)
            assert s1 >= s0
            if s1 - s0 <= self.ds:
                continue
            if (not check_for_active_shifts_and_rotations(line, i_aper_0, i_aper_1)
                    and apertures_are_identical(line[i_aper_0], line[i_aper_1], line)):
                interp_line, i_end_thin_0, i_start_thin_1, s0, s1 = interp_aperture_replicate(
                    line, i_aper_0, i_aper_1, self.ds)
            else:
                interp_line, i_end_thin_0, i_start_thin_1, s0, s1 = interp_aperture_using_polygons(
                    line, i_aper_0, i_aper_1, self.n_theta, self.r_max, self.dr, self.ds)
            interp_line._original_line = self._original_line
            refine_loss_location_single_aperture(
                particles, i_aper_1, i_end_thin_0, line, interp_line, inplace=True,
                allowed_backtrack_types=self.allowed_backtrack_types)
            if self.save_refine_lines:
                interp_line.i_start_thin_0 = i_end_thin_0
                interp_line.i_start_thin_1 = i_start_thin_1
                interp_line.s0 = s0
                interp_line.s1 = s1
                self.refine_lines[i_ap] = interp_line


# ---- what lies between two apertures (:211-281) --------------------------------------------
_FRAME_FIELDS = {'SRotation': ('angle',), 'Rotation': ('rot_s_rad', 'rot_x_rad', 'rot_y_rad'),
                 'Translation': ('shift_x', 'shift_y'), 'XYShift': ('dx', 'dy')}


def check_for_active_shifts_and_rotations(line, i_aper_0, i_aper_1):
    """Is there a frame element that actually moves the frame between the two apertures?"""
    for ii in range(i_aper_0, i_aper_1):
        ee = _resolve(line[ii], line)
        for ff in _FRAME_FIELDS.get(type(ee).__name__, ()):
            if abs(getattr(ee, ff)) > 1e-15:
                return True
    return False


def fields_equal(a, b, atol=1e-15):
    """Equality of two element fields to `atol` (scalars, arrays, sequences of them)."""
    if a is b:
        return True
    if type(a) is not type(b):
        return False
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(fields_equal(x, y, atol) for x, y in zip(a, b))
    if isinstance(a, np.ndarray) or np.isscalar(a):
        aa, bb = np.asarray(a), np.asarray(b)
        if aa.dtype.kind in 'fiub' and bb.dtype.kind in 'fiub':
            return aa.shape == bb.shape and bool(np.all(np.abs(aa - bb) <= atol))
        return aa.shape == bb.shape and bool(np.all(aa == bb))
    return a == b


def apertures_are_identical(aper1, aper2, line):
    aper1, aper2 = _resolve(aper1, line), _resolve(aper2, line)
    if type(aper1) is not type(aper2):
        return False
    d1, d2 = aper1.to_dict(), aper2.to_dict()
    return set(d1) == set(d2) and all(fields_equal(d1[kk], d2[kk]) for kk in d1)


def find_apertures(line):
    i_apertures, apertures = [], []
    for ii, ee in enumerate(line.elements):
        if _is_aperture(ee, line):
            i_apertures.append(ii)
            apertures.append(ee)
    return i_apertures, apertures


def find_adjacent_thick(line, i_element, direction):
    """Index of the nearest thick element from `i_element` on, upstream or downstream."""
    assert direction in ('upstream', 'downstream')
    increment = -1 if direction == 'upstream' else 1
    ii = i_element
    while not _is_thick(line[ii], line):
        ii += increment
        if ii < 0 or ii >= len(line):
            raise ValueError('no thick element next to the aperture')
    return ii


def generate_interp_aperture_locations(line, i_aper_0, i_aper_1, ds):
    s_el = _element_s_locations(line)
    s0, s1 = s_el[i_aper_0], s_el[i_aper_1]
    assert s1 >= s0
    n_segments = int(np.ceil((s1 - s0) / ds))
    if n_segments <= 1:
        s_vect = np.array([])
    else:
        s_vect = np.linspace(s0, s1, n_segments + 1)[1:-1]
    return s0, s1, s_vect


# ---- the refinement proper (:297-381) -------------------------------------------------------
def refine_loss_location_single_aperture(particles, i_aper_1, i_end_thin_0, line, interp_line,
                                         inplace=True, allowed_backtrack_types=()):
    state = particles.get('state')
    mask_part = (state == 0) & (particles.get('at_element') == i_aper_1)
    take = lambda nn: particles.get(nn)[mask_part]
    part_refine = Particles(
        p0c=take('p0c'), mass0=particles.mass0, q0=particles.q0, x=take('x'), px=take('px'),
        y=take('y'), py=take('py'), zeta=take('zeta'), delta=take('delta'), s=take('s'),
        chi=take('chi'), charge_ratio=take('charge_ratio'), _device=particles.device)

    i_start = i_end_thin_0 + 1
    i_stop = i_aper_1
    original = getattr(interp_line, '_original_line', line)
    for nn in original.element_names[i_start:i_stop]:
        ee = original.element_dict[nn]
        can_backtrack = True
        if not _has_backtrack(ee, line):
            can_backtrack = False
        elif not _allow_loss_refinement(ee, line):
            can_backtrack = isinstance(_resolve(ee, line), tuple(allowed_backtrack_types))
        if not can_backtrack:
            if _skip_in_loss_location_refinement(ee, line):
                return 'skipped'
            raise TypeError(f'Cannot backtrack through element {nn} of type '
                            f'{_resolve(ee, line).__class__.__name__}')

    with _preserve_track_flags(line):
        line.track_flags['XS_FLAG_IGNORE_GLOBAL_APERTURE'] = True
        line.track(part_refine, ele_start=i_start, ele_stop=i_stop, backtrack='force')

    # through the stretch with the extra apertures.  (A small fraction is not lost again: they
    # are at the edge, and end at the end of the stretch, which is where they belong.)
    interp_line.track(part_refine)
    st = part_refine.get('state')
    if np.any(st < 0):
        raise RuntimeError(f'Particles are lost with error codes: {st[st < 0]}')

    if inplace:
        order = np.argsort(part_refine.get('particle_id'), kind='stable')
        for nn in ('x', 'px', 'y', 'py', 'zeta', 's', 'delta', 'ptau', 'rvv', 'rpp', 'p0c',
                   'gamma0', 'beta0'):
            cur = particles.get(nn)
            cur[mask_part] = part_refine.get(nn)[order]
            setattr(particles, nn, cur)
    return part_refine


def interp_aperture_replicate(line, i_aper_0, i_aper_1, ds, mode='end'):
    i_start_thin_1 = find_adjacent_thick(line, i_aper_1, 'upstream') + 1
    i_end_thin_0 = find_adjacent_thick(line, i_aper_0, 'downstream') - 1
    s0, s1, s_vect = generate_interp_aperture_locations(line, i_aper_0, i_aper_1, ds)
    if mode not in ('end', 'start'):
        raise ValueError(f'Invalid mode: {mode}')
    aper_to_copy = _resolve(line[i_aper_1 if mode == 'end' else i_aper_0], line)
    interp_line = build_interp_line(
        s0=s0, s1=s1, s_interp=s_vect, aper_0=aper_to_copy.copy(), aper_1=aper_to_copy.copy(),
        aper_interp=[aper_to_copy.copy() for _ in s_vect], line=line,
        i_start_thin_0=i_end_thin_0, i_start_thin_1=i_start_thin_1)
    return interp_line, i_end_thin_0, i_start_thin_1, s0, s1


def _convex_hull_in_order(x, y):
    from scipy.spatial import ConvexHull
    hull = ConvexHull(np.array([x, y]).T)
    i_hull = np.sort(hull.vertices)
    return x[i_hull], y[i_hull]


def interp_aperture_using_polygons(line, i_aper_0, i_aper_1, n_theta, r_max, dr, ds):
    polygon_1, i_start_thin_1 = characterize_aperture(line, i_aper_1, n_theta, r_max, dr,
                                                      coming_from='upstream')
    polygon_0, i_end_thin_0 = characterize_aperture(line, i_aper_0, n_theta, r_max, dr,
                                                    coming_from='downstream')
    s0, s1, s_vect = generate_interp_aperture_locations(line, i_aper_0, i_aper_1, ds)
    delta_s = s1 - s0
    interp_polygons = []
    for ss in s_vect:
        x_nc = (polygon_1.x_vertices * (ss - s0) / delta_s + polygon_0.x_vertices * (s1 - ss) / delta_s)
        y_nc = (polygon_1.y_vertices * (ss - s0) / delta_s + polygon_0.y_vertices * (s1 - ss) / delta_s)
        x_hull, y_hull = _convex_hull_in_order(x_nc, y_nc)
        interp_polygons.append(_el.LimitPolygon(x_vertices=x_hull, y_vertices=y_hull))
    interp_line = build_interp_line(
        s0=s0, s1=s1, s_interp=s_vect, aper_0=polygon_0, aper_1=polygon_1,
        aper_interp=interp_polygons, line=line, i_start_thin_0=i_end_thin_0,
        i_start_thin_1=i_start_thin_1)
    return interp_line, i_end_thin_0, i_start_thin_1, s0, s1


# ---- the refined stretch: elements between the apertures + interpolated apertures (:490-531) --
def _cut_thick_element(name, ee, cuts):
    """`ee` (thick, length L) cut at the distances `cuts` (0 < c < L) from its entry: the
    pieces as (name, element) pairs -- drifts as shorter drifts, magnets and RF elements as
    thick slices of the element between its entry and exit edge slices (what the reference's
    `Line.insert` makes of a thick element)."""
    length = ee.length
    bounds = [0.0, *cuts, length]
    cname = type(ee).__name__
    pieces = []
    if cname in ('Drift', 'DriftExact'):
        for kk in range(len(bounds) - 1):
            new = ee.copy()
            new.length = bounds[kk + 1] - bounds[kk]
            pieces.append((f'{name}..{kk}', new))
        return pieces
    if isinstance(ee, _el._Slice):
        if ee._slice_kind not in ('thick', 'drift'):
            raise NotImplementedError(f'cannot cut a {cname}')
        par = ee.parent
        for kk in range(len(bounds) - 1):
            ww = ee.weight * (bounds[kk + 1] - bounds[kk]) / length
            pieces.append((f'{name}..{kk}', type(ee)(
                parent_name=ee.parent_name, _parent=par, weight=ww,
                slice_offset=ee.slice_offset + bounds[kk], radiation_flag=ee.radiation_flag,
                delta_taper=ee.delta_taper)))
        return pieces
    thick_cls = _el.SLICE_CLASSES.get('ThickSlice' + cname)
    if thick_cls is None:
        raise NotImplementedError(f'cannot place an interpolated aperture inside a {cname}')
    entry_cls = _el.SLICE_CLASSES.get(f'ThinSlice{cname}Entry')
    exit_cls = _el.SLICE_CLASSES.get(f'ThinSlice{cname}Exit')
    if entry_cls is not None:
        pieces.append((f'{name}..entry_map', entry_cls(parent_name=name, _parent=ee)))
    for kk in range(len(bounds) - 1):
        pieces.append((f'{name}..{kk}', thick_cls(
            parent_name=name, _parent=ee, weight=(bounds[kk + 1] - bounds[kk]) / length,
            slice_offset=bounds[kk])))
    if exit_cls is not None:
        pieces.append((f'{name}..exit_map', exit_cls(parent_name=name, _parent=ee,
                                                    slice_offset=length)))
    return pieces


def build_interp_line(s0, s1, s_interp, aper_0, aper_1, aper_interp, line, i_start_thin_0,
                      i_start_thin_1, tol=1e-10):
    """The elements i_start_thin_0 + 1 .. i_start_thin_1 - 1 of `line` with `aper_0` in front,
    `aper_1` behind and `aper_interp[k]` at s = `s_interp[k]` (thick elements cut there)."""
    from .line import Line
    names_in = line.element_names[i_start_thin_0 + 1:i_start_thin_1]
    elements = {}
    names = []
    counter = [0]

    def add_aperture(aper):
        nn = f'_interp_aper_{counter[0]}'
        while nn in line.element_dict or nn in elements:
            counter[0] += 1
            nn = f'_interp_aper_{counter[0]}'
        counter[0] += 1
        elements[nn] = aper
        names.append(nn)

    add_aperture(aper_0)
    pending = list(zip([float(ss) - s0 for ss in s_interp], aper_interp))    # (at, aperture)
    pos = 0.0
    for nn in names_in:
        ee = _resolve(line.element_dict[nn], line)
        ll = ee.length if ee.isthick_now else 0.0
        while pending and pending[0][0] <= pos + tol:                        # in front of it
            add_aperture(pending.pop(0)[1])
        inside = []
        while pending and pending[0][0] < pos + ll - tol:
            inside.append(pending.pop(0))
        if not inside:
            elements[nn] = ee
            if isinstance(ee, _el._Slice) and ee.parent_name is not None:
                elements.setdefault(ee.parent_name, ee.parent)
            names.append(nn)
        else:
            pieces = _cut_thick_element(nn, ee, [at - pos for at, _ in inside])
            if not isinstance(ee, _el._Slice) and type(ee).__name__ not in ('Drift', 'DriftExact'):
                elements[nn] = ee                    # the parent of the slices (not in the sequence)
            elif isinstance(ee, _el._Slice) and ee.parent_name is not None:
                elements.setdefault(ee.parent_name, ee.parent)
            apers = [aa for _, aa in inside]
            for pname, piece in pieces:
                elements[pname] = piece
                names.append(pname)
                is_body = not pname.endswith(('..entry_map', '..exit_map'))
                if is_body and apers and not pname.endswith(f'..{len(inside)}'):
                    add_aperture(apers.pop(0))
        pos += ll
    for _, aper in pending:                                                   # (at the very end)
        add_aperture(aper)
    add_aperture(aper_1)

    interp_line = Line(elements=elements, element_names=names, particle_ref=line.particle_ref)
    interp_line.config.update(line.config)
    interp_line._extra_config.update({kk: vv for kk, vv in line._extra_config.items()
                                      if kk in ('_radiation_model', '_needs_rng')})
    interp_line.build_tracker(_device=line.tracker.device, **line._tracker_kwargs)
    interp_line.reset_s_at_end_turn = False
    interp_line.track_flags['XS_FLAG_IGNORE_GLOBAL_APERTURE'] = True
    return interp_line


# ---- the shape of an aperture as the beam sees it (:570-658) --------------------------------
def polygon_impact_from_origin(x_vertices, y_vertices, theta):
    """Where the rays from the origin at the angles `theta` leave the convex polygon
    (LimitPolygon.impact_point_and_normal with x_in = y_in = 0 in the reference)."""
    xv = np.asarray(x_vertices, dtype=float)
    yv = np.asarray(y_vertices, dtype=float)
    x1, y1 = np.roll(xv, -1), np.roll(yv, -1)
    ex, ey = x1 - xv, y1 - yv                               # edges
    dx, dy = np.cos(theta)[:, None], np.sin(theta)[:, None]
    den = dx * ey[None, :] - dy * ex[None, :]
    with np.errstate(divide='ignore', invalid='ignore'):
        tt = (xv[None, :] * ey[None, :] - yv[None, :] * ex[None, :]) / den      # along the ray
        uu = (xv[None, :] * dy - yv[None, :] * dx) / den                        # along the edge
    ok = (np.abs(den) > 0) & (tt > 0) & (uu >= -1e-12) & (uu <= 1 + 1e-12)
    tt = np.where(ok, tt, np.inf)
    t_hit = tt.min(axis=1)
    if not np.all(np.isfinite(t_hit)):
        raise ValueError('the origin is not inside the aperture polygon')
    return t_hit * dx[:, 0], t_hit * dy[:, 0]


def _first_stopped_radius(line, theta, r_from, r_grid, track_kw):
    """Probe particles at the radii `r_from[j] + r_grid[i]` along every angle `theta[j]`, tracked
    through the aperture: per angle, the index of the first grid radius that is stopped and
    the probes' coordinates (arrays [angle, radius])."""
    rr = r_from[:, None] + r_grid[None, :]
    xx, yy = rr * np.cos(theta)[:, None], rr * np.sin(theta)[:, None]
    logger.info(f'aperture scan: {xx.size} probe particles')
    probes = Particles(p0c=1, x=xx.ravel().copy(), y=yy.ravel().copy(), _device=line.tracker.device)
    with _preserve_track_flags(line):
        line.track_flags['XS_FLAG_IGNORE_GLOBAL_APERTURE'] = True
        line.track(probes, **track_kw)
    by_id = np.argsort(probes.get('particle_id'), kind='stable')
    passed = probes.get('state')[by_id].reshape(rr.shape) > 0
    return np.argmin(passed, axis=1), xx, yy


def characterize_aperture(line, i_aperture, n_theta, r_max, dr, coming_from='upstream'):
    """The aperture `i_aperture` together with the thin transformations around it, as the beam
    sees it: a convex polygon with a vertex at each of `n_theta` angles.  Probe particles on a
    polar grid go through the thin elements between the adjacent thick element and the aperture
    (backwards for `coming_from='downstream'`) -- a coarse radial scan up to `r_max` in a
    hundred steps, then one with step `dr` over the two coarse steps around the radius where each
    angle was stopped (:570-658 of the reference's module: same grids, same polygon)."""
    if coming_from == 'upstream':
        first = find_adjacent_thick(line, i_aperture, 'upstream') + 1
        track_kw = dict(ele_start=first, ele_stop=i_aperture + 1, backtrack=False)
        index_start_thin = first
    elif coming_from == 'downstream':
        last = find_adjacent_thick(line, i_aperture, 'downstream')
        for ii in range(i_aperture, min(last + 1, len(line))):
            if not _has_backtrack(line[ii], line):
                raise TypeError(f'Cannot backtrack through element {line.element_names[ii]}')
        track_kw = dict(ele_start=i_aperture, ele_stop=last, backtrack='force')
        index_start_thin = last - 1
    else:
        raise ValueError(f'Invalid direction: {coming_from}')

    theta = np.linspace(0, 2 * np.pi, n_theta + 1)[:-1]
    coarse_step = r_max / 100.
    coarse = np.arange(0, r_max, coarse_step)
    i_stop, _, _ = _first_stopped_radius(line, theta, np.zeros(n_theta), coarse, track_kw)
    inner = coarse[i_stop - 1]                       # last coarse radius that passed
    fine = np.arange(0, 2 * coarse_step, dr)
    i_stop, xx, yy = _first_stopped_radius(line, theta, inner, fine, track_kw)
    rows = np.arange(n_theta)
    x_hull, y_hull = _convex_hull_in_order(xx[rows, i_stop], yy[rows, i_stop])
    # the hull has no vertex at most angles: put one at every requested angle
    xv, yv = polygon_impact_from_origin(x_hull, y_hull, theta)
    return _el.LimitPolygon(x_vertices=xv, y_vertices=yv), index_start_thin
