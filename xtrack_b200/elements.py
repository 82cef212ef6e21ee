"""Host-side beam-element model for the tracking hot path.

Each class mirrors the *fields and constructor semantics* of the xtrack element
of the same name (reference `_xofields` tables and `__init__` logic), so that a
line stored as xtrack JSON can be loaded with plain `json` and handed either to
the lowering pass (`xtrack_b200.lowering`, product) or to the reference-header
oracle (`oracle/`, tests only).  No tracking physics lives here.

Reference field tables / constructors followed:
  Drift            xtrack/beam_elements/drift.py            (length, model)
  Marker           xtrack/beam_elements/marker.py
  Multipole        xtrack/beam_elements/multipole.py:80-97,143-176
  Quadrupole/...   xtrack/beam_elements/quadrupole.py:65-83, sextupole.py, octupole.py
  Bend / RBend     xtrack/beam_elements/_common.py:559-743, rbend.py:91-213
  Cavity           xtrack/beam_elements/cavity.py:63-122
  DipoleEdge       xtrack/beam_elements/dipole_edge.py:40-134
  SRotation        xtrack/beam_elements/s_rotation.py:30-82
  XYShift          xtrack/beam_elements/xy_shift.py
  LimitRect/Ellipse/Polygon  xtrack/beam_elements/limit_*.py
  misalignment fields        xtrack/base_element.py:274-282
  enumerations               xtrack/beam_elements/_common.py:20-83
"""
import math

import numpy as np

DEFAULT_MULTIPOLE_ORDER = 5          # _common.py:18
UNLIMITED = 1e10                     # _aperture_common.py:6

MODEL_DRIFT = {'adaptive': 0, 'expanded': 1, 'exact': 2}
MODEL_CURVED = {'adaptive': 0, 'full': 1, 'bend-kick-bend': 2, 'rot-kick-rot': 3,
                'mat-kick-mat': 4, 'drift-kick-drift-exact': 5,
                'drift-kick-drift-expanded': 6, 'rot-kick-rot-low-order': 7,
                'rot-kick-rot-high-order': 8, 'expanded': 4}
MODEL_STRAIGHT = {k: v for k, v in MODEL_CURVED.items()
                  if v not in (2, 3) or k == 'expanded'}
MODEL_RF = {k: v for k, v in MODEL_STRAIGHT.items() if v not in (1, 4)}
INTEGRATOR = {'adaptive': 0, 'teapot': 1, 'yoshida4': 2, 'uniform': 3}
EDGE_MODEL = {'suppressed': -1, 'linear': 0, 'full': 1, 'dipole-only': 2}
RBEND_MODEL = {'adaptive': 0, 'curved-body': 1, 'straight-body': 2}

MISALIGN_FIELDS = ('shift_x', 'shift_y', 'shift_s', 'rot_s_rad', 'rot_x_rad',
                   'rot_y_rad', 'rot_s_rad_no_frame', 'rot_shift_anchor')

# keys that xtrack's JSON carries but that have no effect on tracking
_IGNORED_KEYS = {'__class__', 'name_associated_aperture', 'prototype', 'extra',
                 '_isthick', 'hyl', 'auto_to_numpy', 'flag_auto_to_numpy'}


def _enum(value, table, what):
    if value is None:
        return 0
    if isinstance(value, str):
        try:
            return table[value]
        except KeyError:
            raise ValueError(f'Invalid {what}: {value}')
    return int(value)


def _inv_factorial(n):
    """`1.0 / factorial(n, exact=True)` as in _common.py:366,544."""
    return 1.0 / math.factorial(int(n))


def _f(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        v = np.asarray(v).reshape(-1)[0]
    return float(v)


def _prepare_multipolar_params(order, **arrays):
    """Pads coefficient arrays to a common length (reference _common.py:481-546)."""
    order = order or 0
    lengths = [len(a) if a is not None else 0 for a in arrays.values()]
    target_len = max(order + 1, *lengths)
    out = {}
    for name, arr in arrays.items():
        new = np.zeros(target_len, dtype=np.float64)
        if arr is not None:
            new[:len(arr)] = np.asarray(arr, dtype=np.float64)
        out[name] = new
    out['order'] = target_len - 1
    out['inv_factorial_order'] = _inv_factorial(target_len - 1)
    return out


def _rel_arrays(kwargs):
    """knl_rel / ksl_rel default to [0] and are padded to equal length
    (reference _common.py:548-557)."""
    knl_rel = list(kwargs.pop('knl_rel', [0]))
    ksl_rel = list(kwargs.pop('ksl_rel', [0]))
    n = max(len(knl_rel), len(ksl_rel))
    knl_rel += [0] * (n - len(knl_rel))
    ksl_rel += [0] * (n - len(ksl_rel))
    return (np.asarray(knl_rel, dtype=np.float64),
            np.asarray(ksl_rel, dtype=np.float64))


# Every write to an element field bumps this counter; a `Tracker` remembers the value its
# lattice was lowered at and lowers again when it has moved (in the reference the elements are
# views into the tracker's buffer, so an edit -- a cavity voltage scan, `el.knl[1] = ...` --
# takes effect at the next `track()`; here the op stream is a frozen copy of the values).
_MUTATIONS = [0]
_UNTRACKED_ATTRS = frozenset(('_data', '_device', '_host', '_standalone_line'))


def mutation_count():
    return _MUTATIONS[0]


class _TrackedArray(np.ndarray):
    """Coefficient array of an element (knl, ksl, ...): item assignment and in-place
    arithmetic count as element mutations."""

    def __setitem__(self, key, value):
        _MUTATIONS[0] += 1
        np.ndarray.__setitem__(self, key, value)

    def _inplace(name):
        def op(self, other):
            _MUTATIONS[0] += 1
            return getattr(np.ndarray, name)(self, other)
        op.__name__ = name
        return op

    __iadd__ = _inplace('__iadd__')
    __isub__ = _inplace('__isub__')
    __imul__ = _inplace('__imul__')
    __itruediv__ = _inplace('__itruediv__')
    del _inplace


class BeamElement:
    """Base class: class-level flags have the meaning of base_element.py:410-419."""
    isthick = False              # *static* thickness (drives the global-aperture check)
    allow_rot_and_shift = True
    behaves_like_drift = False
    has_backtrack = False
    allow_loss_refinement = False      # backtracking through it while refining a loss location
    needs_rng = False
    iscollective = False

    # monitors opt out: their parameters reach the kernel at launch / `set_inline_monitors`
    # time, and `track(turn_by_turn_monitor=True)` creates one per call
    _mutation_tracked = True

    def __setattr__(self, name, value):
        if self._mutation_tracked and name not in _UNTRACKED_ATTRS:
            _MUTATIONS[0] += 1
            if type(value) is np.ndarray and value.dtype == np.float64:
                value = value.view(_TrackedArray)
        object.__setattr__(self, name, value)

    def _init_misalign(self, kwargs):
        if self.allow_rot_and_shift:
            for nn in MISALIGN_FIELDS:
                setattr(self, nn, float(kwargs.pop(nn, 0.0)))

    def _finish(self, kwargs):
        for kk in list(kwargs):
            if kk in _IGNORED_KEYS:
                kwargs.pop(kk)
        if kwargs:
            raise NameError(f'{type(self).__name__}: invalid argument(s) {sorted(kwargs)}')

    @property
    def has_misalignment(self):
        """`rot_shift_active` of track_local_particle_with_transformations.h:189-196."""
        if not self.allow_rot_and_shift:
            return False
        return any(getattr(self, nn) != 0.0 for nn in MISALIGN_FIELDS[:7])

    @classmethod
    def from_dict(cls, dct):
        dct = dict(dct)
        dct.pop('__class__', None)
        return cls(**dct)

    # names of the constructor arguments that `to_dict` stores (per class)
    _dict_fields = ()

    def to_dict(self):
        """Dictionary form read back by `from_dict` (and by xtrack's own `from_dict`: the keys
        are the reference's field names, base_element.py `to_dict`).  Misalignment fields are
        stored when they are set."""
        out = {'__class__': type(self).__name__}
        for nn in self._dict_fields:
            vv = getattr(self, nn)
            if isinstance(vv, np.ndarray):
                vv = [float(v) for v in vv]
            table = self._enum_table(nn)
            if table is not None:       # enumerations are stored by name, as xtrack does
                vv = next(kk for kk, ii in table.items() if ii == vv)
            out[nn] = vv
        if self.allow_rot_and_shift:
            for nn in MISALIGN_FIELDS:
                if getattr(self, nn) != 0.0:
                    out[nn] = getattr(self, nn)
        return out

    def _enum_table(self, field):
        if field == 'model':
            return getattr(self, '_model_table', None)
        return {'integrator': INTEGRATOR, 'edge_entry_model': EDGE_MODEL,
                'edge_exit_model': EDGE_MODEL, 'rbend_model': RBEND_MODEL}.get(field)

    def copy(self):
        return type(self).from_dict(self.to_dict())

    def track(self, particles=None, increment_at_element=False, _tracker_class=None):
        """Stand-alone tracking of this element (base_element.py:455-480): the element's map on
        every active particle, no end-of-turn action and -- as in the reference's per-element
        kernel -- no global aperture check; `at_element` moves only if asked to."""
        if particles is None:
            raise RuntimeError('Please provide particles to track!')
        from .line import Line
        line = self.__dict__.get('_standalone_line')
        if line is None:
            line = Line(elements=[self])
            line.config['XTRACK_GLOBAL_XY_LIMIT'] = 1e300
            if type(self).needs_rng or getattr(self, 'radiation_flag', 0):
                line.config['XTRACK_MULTIPOLE_NO_SYNRAD'] = not bool(getattr(self, 'radiation_flag', 0))
                line._extra_config['_needs_rng'] = True
            object.__setattr__(self, '_standalone_line', line)
        # (re)lower every time: the element's fields may have changed since the last call
        line.build_tracker(_device=particles.device,
                           **({'_tracker_class': _tracker_class} if _tracker_class else {}))
        at_element = particles.get('at_element').copy()
        state_before = particles.get('state').copy()
        line.track(particles, ele_start=0, num_elements=1, _force_no_end_turn_actions=True)
        if not increment_at_element:
            new = particles.get('at_element').copy()
            moved = (state_before > 0) & (particles.get('state') > 0)
            new[moved] = at_element[moved]
            particles.at_element = new

    def get_length(self):
        return float(getattr(self, 'length', 0.0)) if self.isthick_now else 0.0

    @property
    def isthick_now(self):
        return bool(self.isthick)

    def __repr__(self):
        return f'{type(self).__name__}(...)'


class Marker(BeamElement):
    allow_loss_refinement = True
    allow_rot_and_shift = False
    behaves_like_drift = True
    has_backtrack = True

    def __init__(self, **kwargs):
        kwargs.pop('_dummy', None)
        self._finish(kwargs)


class Drift(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('length', 'model')
    _model_table = MODEL_DRIFT
    isthick = True
    allow_rot_and_shift = False
    behaves_like_drift = True
    has_backtrack = True

    def __init__(self, length=0.0, model=None, **kwargs):
        self.length = float(length)
        self.model = _enum(model, MODEL_DRIFT, 'model')
        self._finish(kwargs)


class DriftExact(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('length',)
    isthick = True
    allow_rot_and_shift = False
    behaves_like_drift = True
    has_backtrack = True

    def __init__(self, length=0.0, **kwargs):
        self.length = float(length)
        self._finish(kwargs)


class _Magnet(BeamElement):
    """Common handling of knl/ksl/order/model/integrator (`_HasKnlKsl.__init__`,
    _common.py:419-460)."""
    has_backtrack = True
    _model_table = MODEL_STRAIGHT
    _default_order = DEFAULT_MULTIPOLE_ORDER

    def _init_knl_ksl(self, kwargs, default_order=True):
        order = kwargs.pop('order', None)
        knl = kwargs.pop('knl', None)
        ksl = kwargs.pop('ksl', None)
        if default_order:
            order = order or self._default_order
        pp = _prepare_multipolar_params(order, knl=knl, ksl=ksl)
        kwargs.pop('inv_factorial_order', None)
        self.knl = pp['knl']
        self.ksl = pp['ksl']
        self.order = pp['order']
        self.inv_factorial_order = pp['inv_factorial_order']
        self.knl_rel, self.ksl_rel = _rel_arrays(kwargs)
        self.model = _enum(kwargs.pop('model', None), self._model_table, 'model')
        self.integrator = _enum(kwargs.pop('integrator', None), INTEGRATOR, 'integrator')

    def _init_common_scalars(self, kwargs):
        self.num_multipole_kicks = int(kwargs.pop('num_multipole_kicks', 0))
        self.radiation_flag = int(kwargs.pop('radiation_flag', 0))
        self.delta_taper = float(kwargs.pop('delta_taper', 0.0))


class Multipole(_Magnet):
    _dict_fields = ('order', 'knl', 'ksl', 'knl_rel', 'ksl_rel', 'model', 'integrator', 'num_multipole_kicks', 'radiation_flag', 'delta_taper', 'length', 'hxl', 'main_order', 'main_is_skew', 'isthick')

    @property
    def isthick(self):          # the dynamic `isthick` field (multipole.py:181)
        return bool(self._isthick_field > 0)

    def __init__(self, **kwargs):
        if 'bal' in kwargs:
            raise ValueError('`bal` not supported anymore')
        if 'hyl' in kwargs:
            assert _f(kwargs['hyl']) == 0.0, 'hyl is not supported anymore'
        self._init_knl_ksl(kwargs, default_order=False)
        self.length = float(kwargs.pop('length', 0.0))
        self.hxl = float(kwargs.pop('hxl', 0.0))
        self.main_order = int(kwargs.pop('main_order', 0))
        self.main_is_skew = int(bool(kwargs.pop('main_is_skew', 0)))
        self._isthick_field = int(bool(kwargs.pop('isthick', 0)))
        self._init_common_scalars(kwargs)
        self._init_misalign(kwargs)
        self._finish(kwargs)

    @property
    def isthick_now(self):
        return self._isthick_field > 0

    @property
    def main_strength(self):
        return (self.ksl if self.main_is_skew else self.knl)[self.main_order]


class _StraightMagnet(_Magnet):
    allow_loss_refinement = True
    isthick = True
    _main = None     # ('k1', 'k1s') ...

    def __init__(self, **kwargs):
        self._init_knl_ksl(kwargs)
        kn, ks = self._main
        setattr(self, kn, float(kwargs.pop(kn, 0.0)))
        setattr(self, ks, float(kwargs.pop(ks, 0.0)))
        self.length = float(kwargs.pop('length', 0.0))
        self.main_is_skew = int(bool(kwargs.pop('main_is_skew', 0)))
        self.edge_entry_active = int(bool(kwargs.pop('edge_entry_active', 0)))
        self.edge_exit_active = int(bool(kwargs.pop('edge_exit_active', 0)))
        self._init_common_scalars(kwargs)
        self._init_misalign(kwargs)
        self._finish(kwargs)


class Quadrupole(_StraightMagnet):
    _dict_fields = ('order', 'knl', 'ksl', 'knl_rel', 'ksl_rel', 'model', 'integrator', 'num_multipole_kicks', 'radiation_flag', 'delta_taper', 'k1', 'k1s', 'length', 'main_is_skew', 'edge_entry_active', 'edge_exit_active')
    _main = ('k1', 'k1s')


class Sextupole(_StraightMagnet):
    _dict_fields = ('order', 'knl', 'ksl', 'knl_rel', 'ksl_rel', 'model', 'integrator', 'num_multipole_kicks', 'radiation_flag', 'delta_taper', 'k2', 'k2s', 'length', 'main_is_skew', 'edge_entry_active', 'edge_exit_active')
    _main = ('k2', 'k2s')


class Octupole(_StraightMagnet):
    _dict_fields = ('order', 'knl', 'ksl', 'knl_rel', 'ksl_rel', 'model', 'integrator', 'num_multipole_kicks', 'radiation_flag', 'delta_taper', 'k3', 'k3s', 'length', 'main_is_skew', 'edge_entry_active', 'edge_exit_active')
    _main = ('k3', 'k3s')


class _BendCommon(_Magnet):
    allow_loss_refinement = True
    isthick = True
    _model_table = MODEL_CURVED

    def _init_bend_fields(self, kwargs):
        self.k1 = float(kwargs.pop('k1', 0.0))
        self.k2 = float(kwargs.pop('k2', 0.0))
        self.edge_entry_active = int(bool(kwargs.pop('edge_entry_active', 1)))
        self.edge_exit_active = int(bool(kwargs.pop('edge_exit_active', 1)))
        for nn in ('edge_entry_angle', 'edge_exit_angle', 'edge_entry_angle_fdown',
                   'edge_exit_angle_fdown', 'edge_entry_fint', 'edge_exit_fint',
                   'edge_entry_hgap', 'edge_exit_hgap'):
            setattr(self, nn, float(kwargs.pop(nn, 0.0)))
        self._init_common_scalars(kwargs)
        self._init_misalign(kwargs)
        # raw defaults before properties fire
        self._k0 = 0.0
        self._k0_from_h = 1
        self._h = 0.0
        self._angle = 0.0
        self._length = 0.0
        self.edge_entry_model = 0
        self.edge_exit_model = 0

    # -- properties that the reference triggers in a fixed order ------------
    @property
    def h(self):
        return self._h

    @property
    def angle(self):
        return self._angle

    @property
    def length(self):
        return self._length

    @property
    def k0(self):
        return self._k0

    @k0.setter
    def k0(self, value):
        self._set_k0(value)

    @property
    def k0_from_h(self):
        return bool(self._k0_from_h)

    @k0_from_h.setter
    def k0_from_h(self, value):
        self._set_k0_from_h(value)

    def _set_k0(self, value):
        # _common.py:660-670
        if isinstance(value, str):
            if value != 'from_h':
                raise ValueError("k0 can only be set to 'from_h' as a string")
            self._set_k0_from_h(True)
        else:
            self._set_k0_from_h(False)
            self._k0 = float(value)

    def to_dict(self):
        out = super().to_dict()
        if self._k0_from_h:
            out['k0_from_h'] = True
        else:
            out['k0'] = self._k0
        return out

    def _set_k0_from_h(self, value):
        # _common.py:676-682
        if value:
            self._k0 = self._h
        elif self._k0_from_h:
            self._k0 = 0.0
        self._k0_from_h = int(bool(value))


class Bend(_BendCommon):
    _dict_fields = ('order', 'knl', 'ksl', 'knl_rel', 'ksl_rel', 'model', 'integrator', 'num_multipole_kicks', 'radiation_flag', 'delta_taper', 'length', 'angle', 'k1', 'k2', 'edge_entry_active', 'edge_exit_active', 'edge_entry_model', 'edge_exit_model', 'edge_entry_angle', 'edge_exit_angle', 'edge_entry_angle_fdown', 'edge_exit_angle_fdown', 'edge_entry_fint', 'edge_exit_fint', 'edge_entry_hgap', 'edge_exit_hgap')

    def __init__(self, **kwargs):
        if 'h' in kwargs:
            # backward compatibility of from_dict (_common.py:733-738)
            if 'angle' not in kwargs:
                kwargs['angle'] = kwargs['h'] * kwargs['length']
            kwargs.pop('h')
        if kwargs.get('k0_from_h', False) and 'k0' not in kwargs:
            kwargs['k0'] = 'from_h'
            kwargs.pop('k0_from_h')
        for nn in ('edge_entry_model', 'edge_exit_model'):
            if '_' + nn in kwargs:          # the xofield's own name, as `to_dict` may store it
                kwargs.setdefault(nn, kwargs['_' + nn])
                kwargs.pop('_' + nn)
        props = [(nn, kwargs.pop(nn)) for nn in
                 ('length', 'angle', 'k0_from_h', 'edge_entry_model',
                  'edge_exit_model', 'k0') if nn in kwargs]
        self._init_knl_ksl(kwargs)
        self._init_bend_fields(kwargs)
        self._finish(kwargs)
        for nn, val in props:
            if nn in ('length', 'angle', 'k0_from_h', 'k0'):
                setattr(self, nn, val)
            else:
                setattr(self, nn, _enum(val, EDGE_MODEL, nn))

    def _set_length(self, val):
        # _common.py:634-643
        self._length = float(val)
        self._h = self._angle / self._length if self._length != 0 else 0.0
        if self._k0_from_h:
            self._k0 = self._h

    def _set_angle(self, val):
        # _common.py:622-628
        self._angle = float(val)
        if self._length != 0:
            self._h = self._angle / self._length
            if self._k0_from_h:
                self._k0 = self._h

    length = property(lambda self: self._length, _set_length)
    angle = property(lambda self: self._angle, _set_angle)


class RBend(_BendCommon):
    _dict_fields = ('order', 'knl', 'ksl', 'knl_rel', 'ksl_rel', 'model', 'integrator', 'num_multipole_kicks', 'radiation_flag', 'delta_taper', 'length_straight', 'angle', 'rbend_angle_diff', 'rbend_model', 'rbend_compensate_sagitta', 'rbend_shift', 'k1', 'k2', 'edge_entry_active', 'edge_exit_active', 'edge_entry_model', 'edge_exit_model', 'edge_entry_angle', 'edge_exit_angle', 'edge_entry_angle_fdown', 'edge_exit_angle_fdown', 'edge_entry_fint', 'edge_exit_fint', 'edge_entry_hgap', 'edge_exit_hgap')

    def __init__(self, **kwargs):
        if 'h' in kwargs:
            raise ValueError('Setting `h` directly is not allowed.')
        if 'length' in kwargs:
            assert 'length_straight' in kwargs
            kwargs.pop('length')
        if kwargs.get('k0_from_h', False) and 'k0' not in kwargs:
            kwargs['k0'] = 'from_h'
            kwargs.pop('k0_from_h')
        for nn in ('edge_entry_model', 'edge_exit_model'):
            if '_' + nn in kwargs:
                kwargs.setdefault(nn, kwargs['_' + nn])
                kwargs.pop('_' + nn)
        props = [(nn, kwargs.pop(nn)) for nn in
                 ('length_straight', 'angle', 'k0_from_h', 'edge_entry_model',
                  'edge_exit_model', 'rbend_angle_diff', 'rbend_model', 'k0')
                 if nn in kwargs]
        self._init_knl_ksl(kwargs)
        self._init_bend_fields(kwargs)
        self._length_straight = 0.0
        self.rbend_model = 0
        self.rbend_compensate_sagitta = int(bool(kwargs.pop('rbend_compensate_sagitta', 1)))
        self.rbend_shift = float(kwargs.pop('rbend_shift', 0.0))
        self._rbend_angle_diff = 0.0
        self._finish(kwargs)
        for nn, val in props:
            if nn in ('length_straight', 'angle', 'rbend_angle_diff', 'k0_from_h', 'k0'):
                setattr(self, nn, val)
            elif nn == 'rbend_model':
                self.rbend_model = _enum(val, RBEND_MODEL, nn)
            else:
                setattr(self, nn, _enum(val, EDGE_MODEL, nn))

    def _set_geometry(name):
        def setter(self, val):
            setattr(self, name, float(val))
            self._update_rbend_h_length_k0()
        return setter

    length_straight = property(lambda self: self._length_straight, _set_geometry('_length_straight'))
    rbend_angle_diff = property(lambda self: self._rbend_angle_diff, _set_geometry('_rbend_angle_diff'))
    angle = property(lambda self: self._angle, _set_geometry('_angle'))
    del _set_geometry

    def _update_rbend_h_length_k0(self):
        # rbend.py:192-213
        angle = self._angle
        ls = self.length_straight
        diff = self.rbend_angle_diff
        theta_in = 0.5 * angle - diff / 2
        theta_out = 0.5 * angle + diff / 2
        if abs(angle) < 1e-10:
            length = ls
            h = 0.0
        elif abs(ls) < 1e-10:
            length = 0.0
            h = 0.0
        else:
            h = (math.sin(theta_in) + math.sin(theta_out)) / ls
            length = angle / h
        self._h = h
        self._length = length
        if self._k0_from_h:
            self._k0 = self._h


class Cavity(BeamElement):
    allow_loss_refinement = True
    _model_table = MODEL_RF
    _dict_fields = ('length', 'voltage', 'frequency', 'lag', 'phase', 'harmonic', 'lag_taper', 'phase_taper', 'absolute_time', 'num_kicks', 'model', 'integrator')
    isthick = True
    has_backtrack = True

    def __init__(self, **kwargs):
        self.model = _enum(kwargs.pop('model', None), MODEL_RF, 'model')
        self.integrator = _enum(kwargs.pop('integrator', None), INTEGRATOR, 'integrator')
        for nn in ('length', 'voltage', 'frequency', 'lag', 'phase', 'harmonic',
                   'lag_taper', 'phase_taper'):
            setattr(self, nn, float(kwargs.pop(nn, 0.0)))
        self.absolute_time = int(kwargs.pop('absolute_time', 0))
        self.num_kicks = int(kwargs.pop('num_kicks', 0))
        self._init_misalign(kwargs)
        self._finish(kwargs)


class CrabCavity(BeamElement):
    """beam_elements/crab_cavity.py:51-63, elements_src/crab_cavity.h: an RF dipole kick
    (track_rf.h:116-156, `transverse_voltage`); `lag` in degrees (deprecated there), `phase` in
    radians, both added."""
    allow_loss_refinement = True
    isthick = True
    has_backtrack = True
    _model_table = MODEL_RF
    _dict_fields = ('length', 'crab_voltage', 'frequency', 'lag', 'phase', 'lag_taper',
                    'phase_taper', 'absolute_time', 'num_kicks', 'model', 'integrator')

    def __init__(self, **kwargs):
        self.model = _enum(kwargs.pop('model', None), MODEL_RF, 'model')
        self.integrator = _enum(kwargs.pop('integrator', None), INTEGRATOR, 'integrator')
        for nn in ('length', 'crab_voltage', 'frequency', 'lag', 'phase', 'lag_taper',
                   'phase_taper'):
            setattr(self, nn, float(kwargs.pop(nn, 0.0)))
        self.absolute_time = int(kwargs.pop('absolute_time', 0))
        self.num_kicks = int(kwargs.pop('num_kicks', 0))
        self._init_misalign(kwargs)
        self._finish(kwargs)


class RFMultipole(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('order', 'knl', 'ksl', 'pn', 'ps', 'phase_n', 'phase_s', 'voltage', 'frequency', 'lag', 'phase', 'absolute_time')
    """rf_multipole.py:51-65; constructor `_HasKnlKsl.__init__` with the phase
    arrays (pn/ps in degrees, phase_n/phase_s in radians)."""
    has_backtrack = True

    def __init__(self, **kwargs):
        order = kwargs.pop('order', None)
        arrays = {nn: kwargs.pop(nn, None) for nn in
                  ('knl', 'ksl', 'pn', 'ps', 'phase_n', 'phase_s')}
        order = order or DEFAULT_MULTIPOLE_ORDER
        pp = _prepare_multipolar_params(order, **arrays)
        kwargs.pop('inv_factorial_order', None)
        for nn in arrays:
            setattr(self, nn, pp[nn])
        self.order = pp['order']
        self.inv_factorial_order = pp['inv_factorial_order']
        for nn in ('voltage', 'frequency', 'lag', 'phase'):
            setattr(self, nn, float(kwargs.pop(nn, 0.0)))
        self.absolute_time = int(kwargs.pop('absolute_time', 0))
        self._init_misalign(kwargs)
        self._finish(kwargs)


class DipoleEdge(BeamElement):
    _model_table = {'linear': 0, 'full': 1, 'suppressed': -1}
    _dict_fields = ('k', 'e1', 'e1_fd', 'hgap', 'fint', 'model', 'side', 'delta_taper')
    has_backtrack = True

    def __init__(self, k=None, e1=None, e1_fd=None, hgap=None, fint=None,
                 model=None, side=None, **kwargs):
        if 'h' in kwargs:
            assert k is None
            k = kwargs.pop('h')
        self.k = float(k or 0.0)
        self.e1 = float(e1 or 0.0)
        self.e1_fd = float(e1_fd or 0.0)
        self.hgap = float(hgap or 0.0)
        self.fint = float(fint or 0.0)
        self.model = _enum(model, self._model_table, 'model')
        self.side = _enum(side, {'entry': 0, 'exit': 1}, 'side')
        self.delta_taper = float(kwargs.pop('delta_taper', 0.0))
        kwargs.pop('r21', None)
        kwargs.pop('r43', None)
        self._init_misalign(kwargs)
        self._finish(kwargs)

    def _r21_r43(self):
        # dipole_edge.py:127-134 (numpy scalar math in the reference); evaluated on access so
        # that edits of k / e1 / e1_fd / hgap / fint are followed, as the reference's setters do
        corr = np.float64(2.0) * self.k * self.hgap * self.fint
        r21 = self.k * np.tan(self.e1)
        e1_v = self.e1 + self.e1_fd
        temp = corr / np.cos(e1_v) * (np.float64(1) + np.sin(e1_v) * np.sin(e1_v))
        r43 = -self.k * np.tan(e1_v - temp)
        return float(r21), float(r43)

    @property
    def r21(self):
        return self._r21_r43()[0]

    @property
    def r43(self):
        return self._r21_r43()[1]


class SRotation(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('cos_z', 'sin_z')
    allow_rot_and_shift = False
    has_backtrack = True

    def __init__(self, angle=None, cos_z=None, sin_z=None, **kwargs):
        # s_rotation.py:53-82 (angle in degrees)
        if angle is None and (cos_z is not None or sin_z is not None):
            if cos_z is None or sin_z is None:
                raise ValueError('At least two of (cos, sin, tan) must be given')
            anglerad = math.atan2(sin_z, cos_z)
        elif angle is not None:
            anglerad = angle / 180 * np.pi
        else:
            anglerad = 0.0
        self.cos_z = float(np.cos(anglerad)) if cos_z is None else float(cos_z)
        self.sin_z = float(np.sin(anglerad)) if sin_z is None else float(sin_z)
        self._finish(kwargs)


class XYShift(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('dx', 'dy')
    allow_rot_and_shift = False
    has_backtrack = True

    def __init__(self, dx=0.0, dy=0.0, **kwargs):
        self.dx = float(dx)
        self.dy = float(dy)
        self._finish(kwargs)


class Translation(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('shift_x', 'shift_y')
    """beam_elements/translation.py:15-40, elements_src/translation.h:13-26 (supersedes the
    deprecated XYShift)."""
    allow_rot_and_shift = False
    has_backtrack = True

    def __init__(self, shift_x=0.0, shift_y=0.0, **kwargs):
        self.shift_x = float(shift_x)
        self.shift_y = float(shift_y)
        self._finish(kwargs)


class Rotation(BeamElement):
    allow_loss_refinement = True
    _dict_fields = ('rot_s_rad', 'rot_x_rad', 'rot_y_rad', 'seq')
    """beam_elements/rotation.py:14-100, elements_src/rotation.h:13-60: up to three frame
    rotations about x, y, s in the order `seq` (default 'yxs'); zero angles are skipped."""
    allow_rot_and_shift = False
    has_backtrack = True
    _AXIS = {'x': 0, 'y': 1, 's': 2}

    def __init__(self, rot_s_rad=0.0, rot_x_rad=0.0, rot_y_rad=0.0, seq='yxs', **kwargs):
        self.rot_s_rad = float(rot_s_rad)
        self.rot_x_rad = float(rot_x_rad)
        self.rot_y_rad = float(rot_y_rad)
        if sorted(seq) != ['s', 'x', 'y']:
            raise ValueError("seq must be a permutation of 'x', 'y', 's'")
        self.seq = seq
        self._finish(kwargs)

    @property
    def _first_rot(self):
        return self._AXIS[self.seq[0]]

    @property
    def _second_rot(self):
        return self._AXIS[self.seq[1]]

    @property
    def _third_rot(self):
        return self._AXIS[self.seq[2]]


class LimitRect(BeamElement):
    _dict_fields = ('min_x', 'max_x', 'min_y', 'max_y')
    has_backtrack = True

    def __init__(self, min_x=-UNLIMITED, max_x=UNLIMITED, min_y=-UNLIMITED,
                 max_y=UNLIMITED, **kwargs):
        self.min_x, self.max_x = float(min_x), float(max_x)
        self.min_y, self.max_y = float(min_y), float(max_y)
        self._init_misalign(kwargs)
        self._finish(kwargs)


class LimitEllipse(BeamElement):
    _dict_fields = ('a_squ', 'b_squ', 'a_b_squ')
    has_backtrack = True

    def __init__(self, a=None, b=None, a_squ=None, b_squ=None, **kwargs):
        # limit_ellipse.py:50-70
        if a is None and a_squ is None:
            a = UNLIMITED
        if b is None and b_squ is None:
            b = UNLIMITED
        if a is not None:
            a_squ = a * a
        if b is not None:
            b_squ = b * b
        a_b_squ = kwargs.pop('a_b_squ', a_squ * b_squ)
        if not (a_squ > 0.0 and b_squ > 0.0):
            raise ValueError('a_squ and b_squ have to be positive definite')
        self.a_squ, self.b_squ, self.a_b_squ = float(a_squ), float(b_squ), float(a_b_squ)
        self._init_misalign(kwargs)
        self._finish(kwargs)


class LimitPolygon(BeamElement):
    _dict_fields = ('x_vertices', 'y_vertices')
    has_backtrack = True

    def __init__(self, x_vertices=None, y_vertices=None, **kwargs):
        assert len(x_vertices) == len(y_vertices)
        self.x_vertices = np.asarray(x_vertices, dtype=np.float64)
        self.y_vertices = np.asarray(y_vertices, dtype=np.float64)
        for nn in ('x_normal', 'y_normal', 'resc_fac', 'svg'):
            kwargs.pop(nn, None)
        self._init_misalign(kwargs)
        self._finish(kwargs)


class _Placeholder(Marker):
    """Element classes outside the hot-path contract (out-of-scope
    list).  They are loaded as markers only when `Line.from_dict` is called with
    `replace_unsupported=True`; their count is reported by the loader."""

    def __init__(self, **kwargs):
        pass


class Replica:
    """base_element.py:619-656: a place in the line that stands for another element (by
    name); `Line.elements` hands out the element at the end of the chain."""

    def __init__(self, parent_name):
        self.parent_name = parent_name

    def __repr__(self):
        return f'Replica(parent_name="{self.parent_name}")'

    def to_dict(self):
        return {'__class__': 'Replica', 'parent_name': self.parent_name}

    @classmethod
    def from_dict(cls, dct):
        return cls(parent_name=dct['parent_name'])

    def resolve(self, element_container, get_name=False):
        target = self.parent_name
        visited = {target}
        while isinstance(element_container[target], Replica):
            target = element_container[target].parent_name
            if target in visited:
                raise RecursionError(f'Resolving replica of `{self.parent_name}` leads to a '
                                     'circular reference: check the correctness of your line.')
            visited.add(target)
        return target if get_name else element_container[target]


# ---- slices (beam_elements/slice_base.py:7-14, slice_elements_{thin,thick,drift,edge}.py) ----
ID_RADIATION_FROM_PARENT = 10


class _Slice(BeamElement):
    """A slice of a thick parent element: holds `weight` (fraction of the parent),
    `slice_offset`, its own `radiation_flag` (10 = the parent's) and `delta_taper`; every
    other parameter, the misalignment included, is the parent's (`_parent`, resolved by
    name when the line is assembled: tracker_data.py:160-172)."""
    has_backtrack = True
    allow_rot_and_shift = False          # no fields of its own ...
    rot_and_shift_from_parent = True     # ... the parent's apply (not for drift slices)
    _slice_kind = None                   # 'thin' | 'thick' | 'drift' | 'entry' | 'exit'
    _parent_class = None

    def __init__(self, parent_name=None, _parent=None, weight=0.0, slice_offset=0.0,
                 radiation_flag=ID_RADIATION_FROM_PARENT, delta_taper=0.0, **kwargs):
        self.parent_name = parent_name
        self._parent = _parent
        self.weight = float(weight)
        self.slice_offset = float(slice_offset)
        self.radiation_flag = int(radiation_flag)
        self.delta_taper = float(delta_taper)
        self._finish(kwargs)

    _dict_fields = ('parent_name', 'weight', 'slice_offset', 'radiation_flag', 'delta_taper')

    @property
    def parent(self):
        if self._parent is None:
            raise RuntimeError(f'{type(self).__name__}: parent `{self.parent_name}` not resolved '
                               '(slices are resolved when they are part of a Line)')
        return self._parent

    @property
    def length(self):
        if self._slice_kind in ('entry', 'exit') or not self.isthick:
            return 0.0
        par = self.parent
        if (type(par).__name__ == 'RBend' and self._slice_kind == 'drift'
                and par.rbend_model == 2):
            # drift_slice_rbend.h: a straight-body drift advances s by the curved length
            return par.length * self.weight
        return par.length * self.weight

    @property
    def has_misalignment(self):
        return self.rot_and_shift_from_parent and self.parent.has_misalignment


def _make_slice_classes():
    out = {}
    parents = {'Bend': Bend, 'RBend': RBend, 'Quadrupole': Quadrupole, 'Sextupole': Sextupole,
               'Octupole': Octupole, 'Multipole': Multipole, 'Cavity': Cavity,
               'CrabCavity': CrabCavity}
    for pname, pcls in parents.items():
        kinds = [('ThinSlice' + pname, 'thin', False, True),
                 ('ThickSlice' + pname, 'thick', True, True),
                 ('DriftSlice' + pname, 'drift', True, False)]
        if pname not in ('Multipole', 'Cavity', 'CrabCavity'):
            kinds += [('ThinSlice' + pname + 'Entry', 'entry', False, True),
                      ('ThinSlice' + pname + 'Exit', 'exit', False, True)]
        for cname, kind, thick, from_parent in kinds:
            out[cname] = type(cname, (_Slice,), dict(
                isthick=thick, _slice_kind=kind, _parent_class=pcls,
                rot_and_shift_from_parent=from_parent,
                behaves_like_drift=(kind == 'drift'),
                allow_loss_refinement=kind in ('thick', 'drift'),
                __doc__=f'{kind} slice of a {pname} (generated wrapper '
                        f'elements_src/{kind if kind in ("thin", "thick", "drift") else "thin"}'
                        f'_slice_*.h of the reference)'))
    for cname, pcls in (('DriftSlice', Drift), ('DriftExactSlice', DriftExact)):
        out[cname] = type(cname, (_Slice,), dict(
            isthick=True, _slice_kind='drift', _parent_class=pcls,
            rot_and_shift_from_parent=False, behaves_like_drift=True,
            allow_loss_refinement=True))
    return out


SLICE_CLASSES = _make_slice_classes()
globals().update(SLICE_CLASSES)

ELEMENT_CLASSES = {cls.__name__: cls for cls in (
    Marker, Drift, DriftExact, Multipole, Quadrupole, Sextupole, Octupole, Bend,
    RBend, Cavity, CrabCavity, RFMultipole, DipoleEdge, SRotation, XYShift, Rotation, Translation, LimitRect, LimitEllipse,
    LimitPolygon, *SLICE_CLASSES.values())}
