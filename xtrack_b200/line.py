"""`Line`: element sequence + the `track()` entry point kept from xtrack.

Reference API mirrored (xtrack/line.py):
  Line.from_dict / from_json  :384-503   (plain `json`; deferred expressions ignored)
  Line.build_tracker          :1537
  Line.track                  :1902-1990 (signature and argument meaning)
  config defaults             :284-293   (XTRACK_GLOBAL_XY_LIMIT=1.0, reset_s_at_end_turn=True,
                                          XTRACK_MULTIPOLE_NO_SYNRAD=True)
  get_length                  :3447-3457
  configure_radiation         :4744-4837
  freeze_longitudinal         :4446-4508

Everything else of `xt.Line` (optics, matching, editing, knobs) is outside the
hot-path scope.
"""
import json

import numpy as np

from . import elements as _el
from .monitors import (ParticlesMonitor, LastTurnsMonitor, BeamPositionMonitor,
                       BeamSizeMonitor, BeamProfileMonitor, BeamStatsMonitor)
from .particles import Particles

_MONITOR_CLASSES = {'ParticlesMonitor': ParticlesMonitor,
                    'LastTurnsMonitor': LastTurnsMonitor,
                    'BeamPositionMonitor': BeamPositionMonitor,
                    'BeamSizeMonitor': BeamSizeMonitor,
                    'BeamProfileMonitor': BeamProfileMonitor,
                    'BeamStatsMonitor': BeamStatsMonitor}


class Line:

    def __init__(self, elements=(), element_names=None, particle_ref=None):
        if isinstance(elements, dict):
            self.element_dict = dict(elements)
            if element_names is None:
                raise ValueError('`element_names` must be provided if `elements` is a dict.')
            self.element_names = list(element_names)
        else:
            if element_names is None:
                element_names = [f'e{ii}' for ii in range(len(elements))]
            self.element_dict = dict(zip(element_names, elements))
            self.element_names = list(element_names)
        self.particle_ref = particle_ref
        self.config = {
            'XTRACK_MULTIPOLE_NO_SYNRAD': True,
            'XTRACK_GLOBAL_XY_LIMIT': 1.0,
        }
        self._extra_config = {
            'skip_end_turn_actions': False,
            'reset_s_at_end_turn': True,
            '_radiation_model': None,
            '_needs_rng': False,
        }
        self.track_flags = {'XS_FLAG_BACKTRACK': False,
                            'XS_FLAG_KILL_CAVITY_KICK': False,
                            'XS_FLAG_IGNORE_GLOBAL_APERTURE': False,
                            'XS_FLAG_IGNORE_LOCAL_APERTURE': False,
                            'XS_FLAG_SR_TAPER': False,
                            'XS_FLAG_SR_KICK_SAME_AS_FIRST': False}
        self.tracker = None
        self._tracker_kwargs = {}
        self.record_last_track = None
        self.time_last_track = None
        self.unsupported_replaced = {}

    # -- construction ------------------------------------------------------
    @classmethod
    def from_dict(cls, dct, replace_unsupported=False):
        """Loads the `elements` / `element_names` / `particle_ref` / `config`
        sections of an xtrack line dictionary.  The xdeps `_var_manager`
        section (deferred expressions) is ignored: element values are taken as
        stored.  Classes outside the hot-path contract raise.  `replace_unsupported`
        is an ALLOW-LIST of class-name patterns (fnmatch) that may be turned into
        markers instead -- `True` stands for `('BeamBeam*',)`, the collective
        beam-beam lenses of xfields that sit in the LHC fixtures as placeholders;
        what was replaced is counted in `line.unsupported_replaced`.  Anything
        else outside the contract still raises: physics is never dropped silently."""
        import fnmatch
        if replace_unsupported is True:
            replace_unsupported = ('BeamBeam*',)
        elif not replace_unsupported:
            replace_unsupported = ()
        elif isinstance(replace_unsupported, str):
            replace_unsupported = (replace_unsupported,)
        if dct.get('__class__', 'Line') != 'Line':
            raise ValueError(f"Expected __class__ to be 'Line', got {dct['__class__']!r}")
        eld = dct['elements']
        names = dct.get('element_names', None)
        if isinstance(eld, list):
            assert names is not None and len(names) == len(eld)
            eld = dict(zip(names, eld))
        replaced = {}
        used = set(names) if names is not None else set(eld)
        # ... plus what the line refers to by name: parents of slices, targets of replicas
        todo = [nn for nn in used if 'parent_name' in eld.get(nn, {})]
        while todo:
            pn = eld[todo.pop()]['parent_name']
            if pn not in used:
                if pn not in eld:
                    raise KeyError(f'parent element `{pn}` is not in the dictionary')
                used.add(pn)
                if 'parent_name' in eld[pn]:
                    todo.append(pn)
        elements = {}
        for nn, ed in eld.items():
            if nn not in used:
                continue
            cname = ed['__class__']
            if cname == 'Replica':
                elements[nn] = _el.Replica.from_dict(ed)
            elif cname in _el.ELEMENT_CLASSES:
                elements[nn] = _el.ELEMENT_CLASSES[cname].from_dict(ed)
            elif cname in _MONITOR_CLASSES:
                elements[nn] = _MONITOR_CLASSES[cname].from_dict(ed)
            elif any(fnmatch.fnmatchcase(cname, pat) for pat in replace_unsupported):
                replaced[cname] = replaced.get(cname, 0) + 1
                elements[nn] = _el.Marker()
            else:
                raise NotImplementedError(
                    f'element class {cname} ({nn}) is outside the hot-path contract '
                    f'(replace_unsupported allows only {tuple(replace_unsupported)})')
        pref = None
        if dct.get('particle_ref') is not None:
            pref = Particles.from_dict(dct['particle_ref'])
        self = cls(elements=elements, element_names=names or list(elements),
                   particle_ref=pref)
        self._resolve_parents()
        self.unsupported_replaced = replaced
        if 'config' in dct:
            for kk in ('XTRACK_MULTIPOLE_NO_SYNRAD', 'XTRACK_GLOBAL_XY_LIMIT',
                       'XTRACK_USE_EXACT_DRIFTS'):
                if kk in dct['config']:
                    self.config[kk] = dct['config'][kk]
        if '_extra_config' in dct:
            for kk in ('skip_end_turn_actions', 'reset_s_at_end_turn', '_radiation_model',
                       '_needs_rng'):
                if kk in dct['_extra_config']:
                    self._extra_config[kk] = dct['_extra_config'][kk]
        return self

    @classmethod
    def from_json(cls, path, **kwargs):
        if str(path).endswith('.gz'):
            import gzip
            with gzip.open(path, 'rt') as fid:
                dct = json.load(fid)
        else:
            with open(path) as fid:
                dct = json.load(fid)
        if 'line' in dct and 'elements' not in dct:
            dct = dct['line']
        return cls.from_dict(dct, **kwargs)

    def to_dict(self, include_var_management=False):
        """line.py:775-838: `elements` (every entry of the element dictionary that the line
        uses, parents of slices and targets of replicas included), `element_names`,
        `particle_ref`, `config`, `_extra_config`.  There is no xdeps variable manager here,
        so nothing else is written."""
        used = set(self.element_names)
        todo = list(used)
        while todo:
            pn = getattr(self.element_dict[todo.pop()], 'parent_name', None)
            if pn is not None and pn not in used:
                used.add(pn)
                todo.append(pn)
        out = {'__class__': 'Line',
               'elements': {nn: ee.to_dict() for nn, ee in self.element_dict.items() if nn in used},
               'element_names': list(self.element_names),
               'config': dict(self.config),
               '_extra_config': dict(self._extra_config)}
        if self.particle_ref is not None:
            out['particle_ref'] = self.particle_ref.to_dict()
        return out

    def to_json(self, path, indent=1, **kwargs):
        """line.py:899-931 (numpy arrays are written as lists; `.gz` paths are compressed)."""
        class _Encoder(json.JSONEncoder):
            def default(self, obj):
                if isinstance(obj, np.ndarray):
                    return obj.tolist()
                if isinstance(obj, np.generic):
                    return obj.item()
                return json.JSONEncoder.default(self, obj)
        dct = self.to_dict(**kwargs)
        if str(path).endswith('.gz'):
            import gzip
            with gzip.open(path, 'wt') as fid:
                json.dump(dct, fid, cls=_Encoder, indent=indent)
        else:
            with open(path, 'w') as fid:
                json.dump(dct, fid, cls=_Encoder, indent=indent)

    def copy(self):
        new = Line.from_dict(self.to_dict(), replace_unsupported=True)
        new.track_flags = dict(self.track_flags)
        return new

    # -- introspection -----------------------------------------------------
    @property
    def elements(self):
        """The elements in line order; a `Replica` is followed to the element it stands for
        (base_element.py:636-653)."""
        dd = self.element_dict
        return tuple(ee.resolve(dd) if isinstance(ee, _el.Replica) else ee
                     for ee in (dd[nn] for nn in self.element_names))

    def _resolve_parents(self):
        """Slices find their parent element by name (tracker_data.py:160-172)."""
        dd = self.element_dict
        for ee in dd.values():
            if isinstance(ee, _el._Slice) and ee.parent_name is not None:
                par = dd[ee.parent_name]
                if isinstance(par, _el.Replica):
                    par = par.resolve(dd)
                if not isinstance(par, ee._parent_class):
                    raise TypeError(f'{type(ee).__name__}: parent `{ee.parent_name}` is a '
                                    f'{type(par).__name__}')
                if ee._parent is not par:
                    ee._parent = par

    def __len__(self):
        return len(self.element_names)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.element_dict[key]
        return self.element_dict[self.element_names[key]]

    def get_length(self):
        ll = 0
        for ee in self.elements:
            if ee.isthick_now:
                ll += ee.length
        return ll

    @property
    def skip_end_turn_actions(self):
        return self._extra_config['skip_end_turn_actions']

    @skip_end_turn_actions.setter
    def skip_end_turn_actions(self, value):
        self._extra_config['skip_end_turn_actions'] = bool(value)

    @property
    def reset_s_at_end_turn(self):
        return self._extra_config['reset_s_at_end_turn']

    @reset_s_at_end_turn.setter
    def reset_s_at_end_turn(self, value):
        self._extra_config['reset_s_at_end_turn'] = bool(value)

    @property
    def _needs_rng(self):
        return self._extra_config['_needs_rng']

    @property
    def _is_backtrackable(self):
        """tracker_data.py: every element of the line states an inverse map"""
        return all(getattr(ee, 'has_backtrack', False) for ee in self.elements)

    def get_flags_register(self):
        """track_flags.py:5-12,42-50"""
        bits = {'XS_FLAG_BACKTRACK': 0, 'XS_FLAG_KILL_CAVITY_KICK': 2,
                'XS_FLAG_IGNORE_GLOBAL_APERTURE': 3, 'XS_FLAG_IGNORE_LOCAL_APERTURE': 4,
                'XS_FLAG_SR_TAPER': 5, 'XS_FLAG_SR_KICK_SAME_AS_FIRST': 6}
        reg = 0
        for nn, bb in bits.items():
            if self.track_flags.get(nn, False):
                reg |= (1 << bb)
        return reg

    # -- configuration -----------------------------------------------------
    def configure_radiation(self, model=None):
        """line.py:4744-4837 (model_beamstrahlung / bhabha / spin are out of scope)."""
        table = {None: 0, 'mean': 1, 'quantum': 2, 'quantum-kick': 3}
        if model not in table:
            raise ValueError(f'Invalid radiation model: {model}')
        flag = table[model]
        self._extra_config['_radiation_model'] = model
        for ee in self.element_dict.values():
            if hasattr(ee, 'radiation_flag'):
                ee.radiation_flag = flag
        self._extra_config['_needs_rng'] = model in ('quantum', 'quantum-kick')
        self.config['XTRACK_MULTIPOLE_NO_SYNRAD'] = (model is None)
        self._invalidate()

    def compensate_radiation_energy_loss(self, delta0='zero_mean', rtol_eneloss=1e-12,
                                         max_iter=100, verbose=True, **kwargs):
        """line.py `compensate_radiation_energy_loss` -> tapering.py:9-159 (see tapering.py)."""
        from . import tapering
        return tapering.compensate_radiation_energy_loss(
            self, delta0=delta0, rtol_eneloss=rtol_eneloss, max_iter=max_iter, verbose=verbose,
            **kwargs)

    def _invalidate(self):
        self.tracker = None

    # -- clean-up passes (xtrack/line.py:4951-5683; see optimize.py) --------
    def optimize_for_tracking(self, compile=True, verbose=False, keep_markers=False):
        from . import optimize
        dev = self.tracker.device if self.tracker is not None else None
        optimize.optimize_for_tracking(self, keep_markers=keep_markers, verbose=verbose)
        if dev is not None and compile and dev.type == 'cuda':
            self.build_tracker(_device=dev, **self._tracker_kwargs)
        return self

    def remove_markers(self, inplace=True, keep=None):
        from . import optimize
        return optimize.remove_markers(self, keep=keep)

    def remove_inactive_multipoles(self, inplace=True, keep=None):
        from . import optimize
        return optimize.remove_inactive_multipoles(self, keep=keep)

    def remove_zero_length_drifts(self, inplace=True, keep=None):
        from . import optimize
        return optimize.remove_zero_length_drifts(self, keep=keep)

    def merge_consecutive_drifts(self, inplace=True, keep=None):
        from . import optimize
        return optimize.merge_consecutive_drifts(self, keep=keep)

    def merge_consecutive_multipoles(self, inplace=True, keep=None):
        from . import optimize
        return optimize.merge_consecutive_multipoles(self, keep=keep)

    def remove_redundant_apertures(self, inplace=True, keep=None, drifts_that_need_aperture=()):
        from . import optimize
        return optimize.remove_redundant_apertures(
            self, keep=keep, drifts_that_need_aperture=drifts_that_need_aperture)

    # -- tracking ----------------------------------------------------------
    def build_tracker(self, _device=None, **kwargs):
        """Lowers the lattice and uploads it to the GPU (replaces
        `Tracker.__init__` + JIT compile, tracker.py:38-147)."""
        from .tracker import Tracker
        # remembered for the implicit rebuilds (device change, configure_radiation,
        # optimize_for_tracking): they keep the user's choices
        self._tracker_kwargs = dict(kwargs)
        self._resolve_parents()
        tracker_class = kwargs.pop('_tracker_class', Tracker)
        self.tracker = tracker_class(self, device=_device, **kwargs)
        return self.tracker

    def _track_in_batches(self, particles, with_progress, *, ele_start, ele_stop, num_elements,
                          num_turns, turn_by_turn_monitor, freeze_longitudinal, time):
        """`with_progress` of xtrack (tracker.py:313-381): the turns are tracked in batches of
        `with_progress` turns (100 for True) -- first batch from `ele_start` to the end of the
        line, middle batches whole turns, last batch down to `ele_stop` -- with one monitor
        spanning all of them.  (The progress bar itself is a host nicety: one line per batch on
        stderr when `XTRACK_B200_PROGRESS` is set.)"""
        import os
        import sys
        if num_turns is None:
            raise ValueError('Tracking with progress indicator is only possible over more than '
                             'one turn.')
        batch_size = 100 if with_progress is True else int(with_progress)
        n_el = len(self.element_names)
        e0 = ele_start or 0
        if isinstance(e0, str):
            e0 = self.element_names.index(e0)
        e1 = ele_stop
        if isinstance(e1, str):
            e1 = self.element_names.index(e1)
        if e1 is None:
            e1 = n_el
        if e0 >= e1:
            num_turns += 1           # the incomplete turn needs its own slot (and monitor space)
        if turn_by_turn_monitor is True:
            _, turn_by_turn_monitor = self.tracker._get_monitor(particles, True, num_turns)
        total_time = 0.0
        num_turns_orig = num_turns - 1 if e0 >= e1 else num_turns
        for ii in range(0, num_turns, batch_size):
            kw = dict(ele_start=ele_start, ele_stop=ele_stop, num_elements=num_elements,
                      num_turns=num_turns_orig)
            first, last = ii == 0, ii + batch_size >= num_turns
            if first and last:
                pass                                     # the only batch: track as normal
            elif first:
                kw.update(ele_stop=None, num_turns=batch_size)
            elif last:
                kw.update(ele_start=None, num_turns=num_turns % batch_size or batch_size)
            else:
                kw.update(ele_start=None, ele_stop=None, num_turns=batch_size)
            self.tracker.track(particles, turn_by_turn_monitor=turn_by_turn_monitor,
                               freeze_longitudinal=freeze_longitudinal, time=time, **kw)
            if time and self.time_last_track is not None:
                total_time += self.time_last_track
            if os.environ.get('XTRACK_B200_PROGRESS'):
                print(f'Tracking: {min(ii + batch_size, num_turns)}/{num_turns} turns',
                      file=sys.stderr)
        if time:
            self.time_last_track = total_time

    def track(self, particles, ele_start=0, ele_stop=None, num_elements=None,
              num_turns=None, turn_by_turn_monitor=None,
              multi_element_monitor_at=None, freeze_longitudinal=False,
              time=False, with_progress=False, **kwargs):
        """Same arguments as xtrack `Line.track` (line.py:1902-1914).  Particles
        are updated in place; `line.record_last_track` holds the monitor and
        `line.time_last_track` the device time when `time=True`."""
        if multi_element_monitor_at is not None:
            raise NotImplementedError('MultiElementMonitor is CPU-only in the reference '
                                      'and outside the hot-path contract')
        backtrack = kwargs.get('backtrack', False)
        if backtrack is not False and with_progress:
            raise NotImplementedError('with_progress while backtracking')
        if self.tracker is None or self.tracker.device != particles.device:
            self.build_tracker(_device=particles.device, **self._tracker_kwargs)
        if with_progress:
            return self._track_in_batches(
                particles, with_progress, ele_start=ele_start, ele_stop=ele_stop,
                num_elements=num_elements, num_turns=num_turns,
                turn_by_turn_monitor=turn_by_turn_monitor,
                freeze_longitudinal=freeze_longitudinal, time=time)
        return self.tracker.track(
            particles, ele_start=ele_start, ele_stop=ele_stop,
            num_elements=num_elements, num_turns=num_turns,
            turn_by_turn_monitor=turn_by_turn_monitor,
            freeze_longitudinal=freeze_longitudinal, time=time,
            _force_no_end_turn_actions=kwargs.get('_force_no_end_turn_actions', False),
            backtrack=backtrack)
