"""ORACLE driver (test infrastructure; only tests/, smoke() and bench.py's CPU
legs may import this).

Runs the reference's own tracking code -- the unmodified physics headers of
/root/reference compiled by `build_ref.py` into `oracle/_ref/libxt_ref_*.so` --
on host numpy arrays.  The element structs it fills are the ones `gen_shim.py`
declares from `element_specs.py`.

Particle ordering: the CPU contexts of the reference permute particles (lost
ones are swapped to the end, local_particle_custom_api.h:108-164); results are
therefore returned sorted by `particle_id` when `restore_order=True`.
"""
import ctypes as ct
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import build_ref                                   # noqa: E402
from element_specs import SPECS, TYPE_ID, all_fields   # noqa: E402

LAST_INVALID_STATE = -999999999

F64_VARS = ('p0c', 'gamma0', 'beta0', 's', 'zeta', 'x', 'y', 'px', 'py', 'ptau',
            'delta', 'rpp', 'rvv', 'chi', 'charge_ratio', 'weight', 'ax', 'ay',
            'spin_x', 'spin_y', 'spin_z', 'anomalous_magnetic_moment')
I64_VARS = ('pdg_id', 'particle_id', 'at_element', 'at_turn', 'state',
            'parent_particle_id')
U32_VARS = ('_rng_s1', '_rng_s2', '_rng_s3', '_rng_s4')
ALL_VARS = F64_VARS + I64_VARS + U32_VARS
_NP = {**{n: np.float64 for n in F64_VARS}, **{n: np.int64 for n in I64_VARS},
       **{n: np.uint32 for n in U32_VARS}}
_CT = {np.float64: ct.c_double, np.int64: ct.c_int64, np.uint32: ct.c_uint32}


class ParticlesDataC(ct.Structure):
    _fields_ = ([(n, ct.c_int64) for n in ('_capacity', '_num_active_particles',
                                           '_num_lost_particles',
                                           'start_tracking_at_element')]
                + [(n, ct.c_double) for n in ('q0', 'mass0', 't_sim')]
                + [(n, ct.POINTER(_CT[_NP[n]])) for n in ALL_VARS])


class ParticlesMonitorC(ct.Structure):
    _fields_ = ([(n, ct.c_int64) for n in (
        'start_at_turn', 'stop_at_turn', 'part_id_start', 'part_id_end', 'ebe_mode',
        'n_records', 'n_repetitions', 'repetition_period', 'flag_auto_to_numpy')]
        + [('data', ct.POINTER(ParticlesDataC))])


class LastTurnsDataC(ct.Structure):
    _fields_ = ([(n, ct.POINTER(ct.c_uint32)) for n in ('lost_at_offset', 'particle_id', 'at_turn')]
                + [(n, ct.POINTER(ct.c_float)) for n in ('x', 'px', 'y', 'py', 'delta', 'zeta')])


class LastTurnsMonitorC(ct.Structure):
    _fields_ = ([(n, ct.c_int64) for n in ('particle_id_start', 'num_particles',
                                           'n_last_turns', 'every_n_turns')]
                + [('data', ct.POINTER(LastTurnsDataC))])


class BeamRecordC(ct.Structure):
    _fields_ = ([('n', ct.c_int64)]
                + [(n, ct.POINTER(ct.c_double)) for n in ('count', 'x_sum', 'y_sum', 'x2_sum', 'y2_sum')])


class BeamMonitorC(ct.Structure):
    _fields_ = ([(n, ct.c_int64) for n in ('particle_id_start', 'num_particles', 'start_at_turn',
                                           'stop_at_turn')]
                + [('frev', ct.c_double), ('sampling_frequency', ct.c_double),
                   ('data', ct.POINTER(BeamRecordC))])


class BeamProfileRecordC(ct.Structure):
    _fields_ = [('n_x', ct.c_int64), ('n_y', ct.c_int64),
                ('counts_x', ct.POINTER(ct.c_double)), ('counts_y', ct.POINTER(ct.c_double))]


class BeamProfileMonitorC(ct.Structure):
    _fields_ = ([(n, ct.c_int64) for n in ('particle_id_start', 'num_particles', 'start_at_turn',
                                           'stop_at_turn', 'nx', 'ny', 'sample_size')]
                + [(n, ct.c_double) for n in ('frev', 'sampling_frequency', 'x_min', 'dx', 'y_min', 'dy')]
                + [('data', ct.POINTER(BeamProfileRecordC))])


class BeamStatsMonitorC(ct.Structure):
    _fields_ = ([(n, ct.c_int64) for n in ('start_at_turn', 'stop_at_turn', 'every_n_turns', '_mode',
                                           '_num_records', '_num_selected_slots', '_num_slices',
                                           '_particle_id_start', '_particle_id_stop')]
                + [(n, ct.c_double) for n in ('_z_min_edge', '_dzeta', '_bunch_spacing_zeta')]
                + [('n_slot_to_selected', ct.c_int64),
                   ('_slot_to_selected', ct.POINTER(ct.c_int64)),
                   ('_selected_slots', ct.POINTER(ct.c_int64)),
                   ('field', ct.POINTER(ct.c_double) * 38), ('len_field', ct.c_int64 * 38),
                   ('touched', ct.POINTER(ct.c_int64)),
                   ('n_profiles', ct.c_int64), ('len_counts', ct.c_int64),
                   ('counts', ct.POINTER(ct.c_double)),
                   ('offsets', ct.POINTER(ct.c_int64)), ('num_bins', ct.POINTER(ct.c_int64)),
                   ('coord_id', ct.POINTER(ct.c_int64)),
                   ('pmin', ct.POINTER(ct.c_double)), ('bin_width', ct.POINTER(ct.c_double))])


_STRUCTS = {}


def _struct_for(name):
    if name not in _STRUCTS:
        ff = [('_parent', ct.c_void_p)] if 'parent' in SPECS[name] else []
        for fn, kind in all_fields(name):
            if kind == 'arr':
                ff += [(fn, ct.POINTER(ct.c_double)), (fn + '__len', ct.c_int64)]
            else:
                ff.append((fn, ct.c_double if kind == 'f64' else ct.c_int64))
        _STRUCTS[name] = type(name + 'DataC', (ct.Structure,), {'_fields_': ff})
    return _STRUCTS[name]


_LIBS = {}


def available():
    return os.path.exists(build_ref.lib_path('serial')) or build_ref.available()


def load(variant='serial'):
    if variant not in _LIBS:
        path = build_ref.lib_path(variant)
        if build_ref.available():
            build_ref.build([variant])
        if not os.path.exists(path):
            raise RuntimeError(f'{path} missing and /root/reference not available to build it')
        lib = ct.CDLL(path)
        lib.xt_ref_track_line.restype = None
        lib.xt_ref_track_line.argtypes = [
            ct.POINTER(ParticlesDataC), ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_int64),
            ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int,
            ct.c_double, ct.c_void_p, ct.c_uint64]
        lib.xt_ref_set_global_xy_limit.argtypes = [ct.c_double]
        lib.xt_ref_init_rand_gen.argtypes = [ct.POINTER(ParticlesDataC),
                                             ct.POINTER(ct.c_uint32), ct.c_int]
        lib.xt_ref_num_threads.restype = ct.c_int
        lib.xt_ref_set_num_threads.argtypes = [ct.c_int]
        lib.xt_ref_set_synrad_tables.argtypes = [ct.c_void_p]
        _LIBS[variant] = lib
    return _LIBS[variant]


class HostParticles:
    """Plain numpy SoA + the C view of it."""

    def __init__(self, fields, q0=1.0, mass0=938272088.16, t_sim=0.0):
        n = len(fields['x'])
        self.arrays = {nn: np.ascontiguousarray(fields[nn], dtype=_NP[nn]).copy()
                       for nn in ALL_VARS}
        self.q0, self.mass0, self.t_sim = q0, mass0, t_sim
        self.capacity = n
        self.c = ParticlesDataC()
        self.c._capacity = n
        self.c.start_tracking_at_element = -1
        self.c.q0, self.c.mass0, self.c.t_sim = q0, mass0, t_sim
        for nn in ALL_VARS:
            setattr(self.c, nn, self.arrays[nn].ctypes.data_as(ct.POINTER(_CT[_NP[nn]])))
        self.reorganize()

    @classmethod
    def from_particles(cls, p):
        """From an `xtrack_b200.Particles` (any device)."""
        return cls({nn: p.get(nn) for nn in ALL_VARS}, q0=p.q0, mass0=p.mass0,
                   t_sim=p.t_sim)

    def reorganize(self):
        st = self.arrays['state']
        act = st > 0
        lost = (st < 1) & (st > LAST_INVALID_STATE)
        na, nl = int(act.sum()), int(lost.sum())
        if not act[:na].all():
            for nn in ALL_VARS:
                if nn.startswith('_rng'):
                    continue
                vv = self.arrays[nn]
                va, vl = vv[act].copy(), vv[lost].copy()
                vv[:na] = va
                vv[na:na + nl] = vl
                vv[na + nl:] = LAST_INVALID_STATE
        self.c._num_active_particles = na
        self.c._num_lost_particles = nl

    def sorted_by_id(self):
        """dict of arrays ordered by particle_id (allocated particles only)."""
        st = self.arrays['state']
        alloc = np.where(st > LAST_INVALID_STATE)[0]
        order = alloc[np.argsort(self.arrays['particle_id'][alloc], kind='stable')]
        return {nn: self.arrays[nn][order].copy() for nn in ALL_VARS}


def _attr(el, name, fn):
    spec = SPECS[name]
    an = spec.get('attr', {}).get(fn, fn)
    if hasattr(el, an):
        return getattr(el, an)
    if fn in spec.get('defaults', {}):
        return spec['defaults'][fn]
    if fn == '_dummy':
        return 0
    raise AttributeError(f'{name}: no attribute {an}')


class RefElements:
    """ElementRefData: array of pointers to element structs + type ids
    (tracker_data.py:237-255)."""

    def __init__(self, elements):
        self._keep = []
        n = len(elements)
        self.ptrs = (ct.c_void_p * n)()
        self.type_ids = (ct.c_int64 * n)()
        self._cache = {}
        for ii, el in enumerate(elements):
            self.ptrs[ii], self.type_ids[ii] = self._get(el)

    def _get(self, el):
        key = id(el)
        if key not in self._cache:
            self._cache[key] = self._make(el)
        return self._cache[key]

    def _make(self, el):
        name = type(el).__name__
        if name == 'ParticlesMonitor':
            return self._make_monitor(el), 1000
        if name == 'LastTurnsMonitor':
            return self._make_last_turns(el), 1001
        if name in ('BeamPositionMonitor', 'BeamSizeMonitor'):
            return self._make_beam_monitor(el), 1002 if name == 'BeamPositionMonitor' else 1003
        if name == 'BeamProfileMonitor':
            return self._make_beam_profile(el), 1004
        if name == 'BeamStatsMonitor':
            return self._make_beam_stats(el), 1005
        if name not in SPECS:
            raise NotImplementedError(f'oracle: element class {name} not supported')
        st = _struct_for(name)()
        if 'parent' in SPECS[name]:
            # (the parent's struct is shared by all its slices, as in the reference's buffer)
            assert type(el.parent).__name__ == SPECS[name]['parent']
            st._parent = self._get(el.parent)[0]
        for fn, kind in all_fields(name):
            vv = _attr(el, name, fn)
            if kind == 'arr':
                arr = np.ascontiguousarray(vv, dtype=np.float64)
                self._keep.append(arr)
                setattr(st, fn, arr.ctypes.data_as(ct.POINTER(ct.c_double)))
                setattr(st, fn + '__len', len(arr))
            elif kind == 'f64':
                setattr(st, fn, float(vv))
            else:
                setattr(st, fn, int(vv))
        self._keep.append(st)
        return ct.addressof(st), TYPE_ID[name]

    def _make_monitor(self, mon):
        cm, keep = make_monitor_struct(mon)
        self._keep += [cm, keep]
        return ct.addressof(cm)

    def _make_beam_stats(self, mon):
        """The oracle accumulates into `mon._host`: dict with 'moments' float64 [38, flat]
        (rows in the reference's field order; rows the monitor does not keep have length 0
        for the reference's code), 'touched' int64 [n_records], 'profile_counts' float64."""
        from xtrack_b200.monitors import BSM_RAW_FIELDS, _BSM_COORDS
        hh = mon._host
        cm = BeamStatsMonitorC()
        for nn in ('start_at_turn', 'stop_at_turn', 'every_n_turns', '_mode', '_num_records',
                   '_num_selected_slots', '_num_slices', '_particle_id_start', '_particle_id_stop',
                   '_z_min_edge', '_dzeta', '_bunch_spacing_zeta'):
            setattr(cm, nn, getattr(mon, nn))
        i64 = lambda arr: np.ascontiguousarray(arr, dtype=np.int64)
        s2s, sel = i64(mon._slot_to_selected), i64(mon._selected_slots if len(mon._selected_slots) else [0])
        cm.n_slot_to_selected = len(mon._slot_to_selected)
        cm._slot_to_selected = s2s.ctypes.data_as(ct.POINTER(ct.c_int64))
        cm._selected_slots = sel.ctypes.data_as(ct.POINTER(ct.c_int64))
        for ii, ff in enumerate(BSM_RAW_FIELDS):
            cm.field[ii] = hh['moments'][ii].ctypes.data_as(ct.POINTER(ct.c_double))
            cm.len_field[ii] = hh['moments'].shape[1] if ff in mon._needed_fields else 0
        cm.touched = hh['touched'].ctypes.data_as(ct.POINTER(ct.c_int64))
        cfgs = list(mon._profile_config.items())
        offs, total = [], 0
        for _, cfg in cfgs:
            offs.append(total)
            total += mon._flat_size * cfg['num_bins']
        offsets, nbins = i64(offs or [0]), i64([c['num_bins'] for _, c in cfgs] or [0])
        cid = i64([_BSM_COORDS.index(cc) for cc, _ in cfgs] or [0])
        pmin = np.ascontiguousarray([c['range'][0] for _, c in cfgs] or [0.], dtype=np.float64)
        bw = np.ascontiguousarray([(c['range'][1] - c['range'][0]) / c['num_bins'] for _, c in cfgs]
                                  or [1.], dtype=np.float64)
        cm.n_profiles = len(cfgs)
        cm.len_counts = total
        cm.counts = hh['profile_counts'].ctypes.data_as(ct.POINTER(ct.c_double))
        cm.offsets = offsets.ctypes.data_as(ct.POINTER(ct.c_int64))
        cm.num_bins = nbins.ctypes.data_as(ct.POINTER(ct.c_int64))
        cm.coord_id = cid.ctypes.data_as(ct.POINTER(ct.c_int64))
        cm.pmin = pmin.ctypes.data_as(ct.POINTER(ct.c_double))
        cm.bin_width = bw.ctypes.data_as(ct.POINTER(ct.c_double))
        self._keep += [cm, s2s, sel, offsets, nbins, cid, pmin, bw]
        return ct.addressof(cm)

    def _make_beam_monitor(self, mon):
        """The oracle accumulates into `mon._host`: float64 [5, n_slots] (rows count, x_sum,
        y_sum, x2_sum, y2_sum; the position monitor uses the first three)."""
        rec = BeamRecordC()
        rec.n = mon.n_slots
        for ii, nn in enumerate(('count', 'x_sum', 'y_sum', 'x2_sum', 'y2_sum')):
            setattr(rec, nn, mon._host[ii].ctypes.data_as(ct.POINTER(ct.c_double)))
        cm = BeamMonitorC()
        for nn in ('particle_id_start', 'num_particles', 'start_at_turn', 'stop_at_turn',
                   'frev', 'sampling_frequency'):
            setattr(cm, nn, getattr(mon, nn))
        cm.data = ct.pointer(rec)
        self._keep += [rec, cm]
        return ct.addressof(cm)

    def _make_beam_profile(self, mon):
        """The oracle counts into `mon._host = {'counts_x': ..., 'counts_y': ...}` (float64)."""
        rec = BeamProfileRecordC()
        rec.n_x, rec.n_y = len(mon._host['counts_x']), len(mon._host['counts_y'])
        rec.counts_x = mon._host['counts_x'].ctypes.data_as(ct.POINTER(ct.c_double))
        rec.counts_y = mon._host['counts_y'].ctypes.data_as(ct.POINTER(ct.c_double))
        cm = BeamProfileMonitorC()
        for nn, _ in BeamProfileMonitorC._fields_[:-1]:
            setattr(cm, nn, getattr(mon, nn))
        cm.data = ct.pointer(rec)
        self._keep += [rec, cm]
        return ct.addressof(cm)

    def _make_last_turns(self, mon):
        d = LastTurnsDataC()
        for nn in ('lost_at_offset', 'particle_id', 'at_turn'):
            setattr(d, nn, mon._host[nn].ctypes.data_as(ct.POINTER(ct.c_uint32)))
        for nn in ('x', 'px', 'y', 'py', 'delta', 'zeta'):
            setattr(d, nn, mon._host[nn].ctypes.data_as(ct.POINTER(ct.c_float)))
        cm = LastTurnsMonitorC()
        cm.particle_id_start = mon.particle_id_start
        cm.num_particles = mon.num_particles
        cm.n_last_turns = mon.n_last_turns
        cm.every_n_turns = mon.every_n_turns
        cm.data = ct.pointer(d)
        self._keep += [d, cm]
        return ct.addressof(cm)


class HostMonitor:
    """Host ParticlesMonitor record store for the oracle (zero-initialised,
    monitors/particles_monitor.py:78-104)."""

    def __init__(self, start_at_turn, stop_at_turn, part_id_start, part_id_end,
                 ebe_mode=0, n_repetitions=1, repetition_period=-1):
        self.start_at_turn, self.stop_at_turn = int(start_at_turn), int(stop_at_turn)
        self.part_id_start, self.part_id_end = int(part_id_start), int(part_id_end)
        self.ebe_mode = int(ebe_mode)
        self.n_repetitions, self.repetition_period = int(n_repetitions), int(repetition_period)
        n_turns = self.stop_at_turn - self.start_at_turn
        self.n_records = n_turns * (self.part_id_end - self.part_id_start) * self.n_repetitions
        self.arrays = {nn: np.zeros(self.n_records, dtype=_NP[nn]) for nn in ALL_VARS}

    def field(self, name):
        n_cols = self.stop_at_turn - self.start_at_turn
        if self.n_repetitions == 1:
            return self.arrays[name].reshape(-1, n_cols)
        return self.arrays[name].reshape(self.n_repetitions, -1, n_cols)


def make_monitor_struct(mon):
    data = ParticlesDataC()
    data._capacity = mon.n_records
    for nn in ALL_VARS:
        setattr(data, nn, mon.arrays[nn].ctypes.data_as(ct.POINTER(_CT[_NP[nn]])))
    cm = ParticlesMonitorC()
    for nn in ('start_at_turn', 'stop_at_turn', 'part_id_start', 'part_id_end',
               'ebe_mode', 'n_records', 'n_repetitions', 'repetition_period'):
        setattr(cm, nn, getattr(mon, nn))
    cm.data = ct.pointer(data)
    return cm, data


def track_line(hp, ref_elements, *, num_turns, ele_start, num_ele_track,
               flag_end_turn_actions, flag_reset_s_at_end_turn, flag_monitor=0,
               num_ele_line=None, line_length=0.0, monitor=None, track_flags=0,
               global_xy_limit=1.0, variant='serial'):
    """One `track_line` launch (argument meaning of tracker.py:546-564)."""
    lib = load(variant)
    lib.xt_ref_set_global_xy_limit(float(global_xy_limit))
    mon_ptr, keep = None, None
    if monitor is not None:
        cm, keep = make_monitor_struct(monitor)
        mon_ptr = ct.addressof(cm)
    lib.xt_ref_track_line(
        ct.byref(hp.c), ref_elements.ptrs, ref_elements.type_ids,
        int(num_turns), int(ele_start), int(num_ele_track),
        int(bool(flag_end_turn_actions)), int(bool(flag_reset_s_at_end_turn)),
        int(flag_monitor), int(num_ele_line or len(ref_elements.ptrs)),
        float(line_length), mon_ptr, int(track_flags))


_synrad_tables_keep = {}


def set_synrad_tables(blob, variant='synrad'):
    """Hands the quantum-kick tables (xtrack_b200.synrad_tables blob) to the reference code."""
    lib = load(variant)
    if blob is None:
        _synrad_tables_keep.pop(variant, None)
        lib.xt_ref_set_synrad_tables(None)
        return
    arr = np.ascontiguousarray(blob, dtype=np.float64)
    _synrad_tables_keep[variant] = arr
    lib.xt_ref_set_synrad_tables(arr.ctypes.data_as(ct.c_void_p))


def init_rand_gen(hp, seeds, variant='serial'):
    lib = load(variant)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    lib.xt_ref_init_rand_gen(ct.byref(hp.c), seeds.ctypes.data_as(ct.POINTER(ct.c_uint32)),
                             len(seeds))
