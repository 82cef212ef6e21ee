"""Turn-by-turn monitors kept from xtrack's API.

ParticlesMonitor  -- xtrack/monitors/particles_monitor.py:12-142,180-260 and
                     particles_monitor.h:13-77: one full 32-field particle
                     record per (particle, turn); record index
                     `n_turns*(particle_id - part_id_start) + (at_turn - start_at_turn)`.
LastTurnsMonitor  -- xtrack/monitors/last_turns_monitor.py:18-117 and
                     last_turns_monitor.h:16-55: rolling FP32 buffer of the
                     last N recorded turns of every particle.

Record storage lives on the tracking device as torch tensors in exactly the
reference layout (`[field][particle_row][turn_col]`, zero-initialised), so that
`monitor.x` etc. have the reference's shapes and content.
"""
import numpy as np
import torch

from .elements import BeamElement
from .particles import PER_PARTICLE_VARS, U32_VARS, _TORCH_DTYPE


class ParticlesMonitor(BeamElement):
    _mutation_tracked = False
    allow_rot_and_shift = False
    behaves_like_drift = True
    has_backtrack = True
    allow_loss_refinement = True

    def __init__(self, start_at_turn=None, stop_at_turn=None, n_repetitions=None,
                 repetition_period=None, num_particles=None, particle_id_range=None,
                 ebe_mode=0, _device='cpu', **kwargs):
        if particle_id_range is not None:
            assert num_particles is None
            part_id_start, part_id_end = particle_id_range
        else:
            assert num_particles is not None
            part_id_start, part_id_end = 0, num_particles
        assert part_id_end - part_id_start >= 0
        if repetition_period is not None:
            assert n_repetitions is not None
        if n_repetitions is not None:
            assert repetition_period is not None
        if repetition_period is None:
            repetition_period = -1
            n_repetitions = 1
        self.start_at_turn = int(start_at_turn)
        self.stop_at_turn = int(stop_at_turn)
        self.part_id_start = int(part_id_start)
        self.part_id_end = int(part_id_end)
        self.ebe_mode = int(ebe_mode)
        self.n_repetitions = int(n_repetitions)
        self.repetition_period = int(repetition_period)
        n_turns = self.stop_at_turn - self.start_at_turn
        self.n_records = n_turns * (self.part_id_end - self.part_id_start) * self.n_repetitions
        self._device = torch.device(_device)
        self._data = None
        self._finish(kwargs)

    @classmethod
    def from_dict(cls, dct):
        # particles_monitor.py:113-142
        dct = dict(dct)
        for kk in ('__class__', 'data', 'n_records', 'auto_to_numpy'):
            dct.pop(kk, None)
        dct.setdefault('start_at_turn', 0)
        ps, pe = dct.pop('part_id_start', None), dct.pop('part_id_end', None)
        if 'particle_id_range' not in dct and pe is not None:
            dct['particle_id_range'] = (ps or 0, pe)
        if dct.get('repetition_period', None) == -1:
            dct.pop('repetition_period')
        if 'repetition_period' not in dct and dct.get('n_repetitions', None) == 1:
            dct.pop('n_repetitions')
        return cls(**dct)

    def to_dict(self):
        """Parameters only, no `data` (particles_monitor.py:106-112)."""
        return {'__class__': 'ParticlesMonitor', 'start_at_turn': self.start_at_turn,
                'stop_at_turn': self.stop_at_turn, 'part_id_start': self.part_id_start,
                'part_id_end': self.part_id_end, 'ebe_mode': self.ebe_mode,
                'n_repetitions': self.n_repetitions, 'repetition_period': self.repetition_period,
                'n_records': self.n_records, 'auto_to_numpy': True}

    # -- storage -------------------------------------------------------------
    def allocate(self, device=None):
        if device is not None:
            self._device = torch.device(device)
        if self._data is None or self._data[PER_PARTICLE_VARS[0][0]].device != self._device:
            self._data = {nn: torch.zeros(self.n_records, dtype=_TORCH_DTYPE[dt],
                                          device=self._device)
                          for nn, dt in PER_PARTICLE_VARS}
        return self._data

    def field_pointers(self):
        data = self.allocate()
        return [int(data[nn].data_ptr()) for nn, _ in PER_PARTICLE_VARS]

    def _shape(self):
        n_cols = self.stop_at_turn - self.start_at_turn
        if self.n_repetitions == 1:
            return (self.n_records // n_cols, n_cols)
        return (self.n_repetitions, self.n_records // n_cols // self.n_repetitions, n_cols)

    def get(self, name):
        """Field as numpy, shaped `(n_particles, n_turns)` (or with a leading
        frame axis), particles_monitor.py:153-171."""
        vv = self.allocate()[name].cpu().numpy()
        if name in U32_VARS:
            vv = vv.view(np.uint32)
        return vv.reshape(self._shape())

    def __getattr__(self, name):
        if name.startswith('_') and name not in U32_VARS:
            raise AttributeError(name)
        if name in dict(PER_PARTICLE_VARS):
            return self.get(name)
        if name == 'pzeta':
            return self.get('ptau') / self.get('beta0')
        raise AttributeError(name)


class LastTurnsMonitor(BeamElement):
    _mutation_tracked = False
    allow_rot_and_shift = False
    behaves_like_drift = True
    allow_loss_refinement = True

    properties = ('particle_id', 'at_turn', 'x', 'px', 'y', 'py', 'delta', 'zeta')

    def __init__(self, *, n_last_turns=None, num_particles=None, particle_id_range=None,
                 every_n_turns=1, _device='cpu', **kwargs):
        if num_particles is not None and particle_id_range is None:
            particle_id_start = 0
        elif particle_id_range is not None and num_particles is None:
            particle_id_start = particle_id_range[0]
            num_particles = particle_id_range[1] - particle_id_range[0]
        else:
            raise ValueError('Exactly one of `num_particles` or `particle_id_range` '
                             'parameters must be specified')
        self.particle_id_start = int(particle_id_start)
        self.num_particles = int(num_particles)
        self.n_last_turns = int(n_last_turns)
        self.every_n_turns = int(every_n_turns)
        self._device = torch.device(_device)
        self._data = None
        self._finish(kwargs)

    @classmethod
    def from_dict(cls, dct):
        dct = dict(dct)
        for kk in ('__class__', 'data'):
            dct.pop(kk, None)
        ps = dct.pop('particle_id_start', None)
        if ps is not None:
            dct['particle_id_range'] = (ps, ps + dct.pop('num_particles'))
        return cls(**dct)

    def to_dict(self):
        """Parameters only, no `data` (last_turns_monitor.py:18-44)."""
        return {'__class__': 'LastTurnsMonitor', 'particle_id_start': self.particle_id_start,
                'num_particles': self.num_particles, 'n_last_turns': self.n_last_turns,
                'every_n_turns': self.every_n_turns}

    def allocate(self, device=None):
        if device is not None:
            self._device = torch.device(device)
        if self._data is None or self._data['x'].device != self._device:
            size = self.num_particles * self.n_last_turns
            dd = {'lost_at_offset': torch.zeros(self.num_particles, dtype=torch.int32,
                                                device=self._device)}
            for nn in ('particle_id', 'at_turn'):
                dd[nn] = torch.zeros(size, dtype=torch.int32, device=self._device)
            for nn in ('x', 'px', 'y', 'py', 'delta', 'zeta'):
                dd[nn] = torch.zeros(size, dtype=torch.float32, device=self._device)
            self._data = dd
        return self._data

    def field_pointers(self):
        dd = self.allocate()
        return [int(dd[nn].data_ptr()) for nn in
                ('lost_at_offset', 'particle_id', 'at_turn', 'x', 'px', 'y', 'py',
                 'delta', 'zeta')]

    def __getattr__(self, attr):
        if attr in LastTurnsMonitor.properties:
            # un-roll the ring buffer (last_turns_monitor.py:108-117)
            dd = self.allocate()
            val = dd[attr].cpu().numpy()
            if attr in ('particle_id', 'at_turn'):
                val = val.view(np.uint32)
            val = np.reshape(val, (self.num_particles, self.n_last_turns))
            off = dd['lost_at_offset'].cpu().numpy().view(np.uint32).astype(np.int64) + 1
            r, c = np.ogrid[:val.shape[0], :val.shape[1]]
            c = (c + off[:, np.newaxis]) % val.shape[1]
            return val[r, c]
        raise AttributeError(attr)


class _BeamSlotMonitor(BeamElement):
    _mutation_tracked = False
    """Common part of BeamPositionMonitor / BeamSizeMonitor: per time slot
    `i = round(sampling_frequency * ((at_turn - start_at_turn) / frev - zeta / beta0 / c0))`
    the count and the sums of x, y (, x^2, y^2) of the particles crossing the monitor
    (xtrack/monitors/beam_position_monitor.py:25-137, beam_size_monitor.py:27-149).
    The record lives on the tracking device: one float64 tensor `[n_sums, n_slots]`."""
    allow_rot_and_shift = False
    behaves_like_drift = True
    allow_loss_refinement = True
    properties = ()

    def __init__(self, *, particle_id_range=None, particle_id_start=None, num_particles=None,
                 start_at_turn=None, stop_at_turn=None, frev=None, sampling_frequency=None,
                 _device='cpu', **kwargs):
        if particle_id_range is None:
            if particle_id_start is None:
                particle_id_start = 0
            if num_particles is None:
                num_particles = -1
        elif particle_id_start is None and num_particles is None:
            particle_id_start = particle_id_range[0]
            num_particles = particle_id_range[1] - particle_id_range[0]
        else:
            raise ValueError('Parameter `particle_id_range` must not be used together with '
                             '`num_particles` and/or `particle_id_start`')
        self.particle_id_start = int(particle_id_start)
        self.num_particles = int(num_particles)
        self.start_at_turn = int(0 if start_at_turn is None else start_at_turn)
        self.stop_at_turn = int(0 if stop_at_turn is None else stop_at_turn)
        self.frev = float(1 if frev is None else frev)
        self.sampling_frequency = float(1 if sampling_frequency is None else sampling_frequency)
        self.n_slots = int(round((self.stop_at_turn - self.start_at_turn)
                                 * self.sampling_frequency / self.frev))
        self._device = torch.device(_device)
        self._data = None
        data = kwargs.pop('data', None)
        self._finish(kwargs)
        if data is not None:
            self.allocate()
            for ii, prop in enumerate(self.properties):
                self._data[ii] = torch.as_tensor(np.asarray(data[prop], dtype=np.float64))

    @classmethod
    def from_dict(cls, dct):
        dct = dict(dct)
        for kk in ('__class__', '_index'):
            dct.pop(kk, None)
        return cls(**dct)

    def to_dict(self):
        return {'__class__': type(self).__name__, 'particle_id_start': self.particle_id_start,
                'num_particles': self.num_particles, 'start_at_turn': self.start_at_turn,
                'stop_at_turn': self.stop_at_turn, 'frev': self.frev,
                'sampling_frequency': self.sampling_frequency,
                'data': {pp: getattr(self, pp).tolist() for pp in self.properties}}

    def allocate(self, device=None):
        if device is not None:
            self._device = torch.device(device)
        if self._data is None:
            self._data = torch.zeros((len(self.properties), max(self.n_slots, 1)),
                                     dtype=torch.float64, device=self._device)
        elif self._data.device != self._device:
            self._data = self._data.to(self._device)
        return self._data

    def __getattr__(self, attr):
        if attr.startswith('_'):
            raise AttributeError(attr)
        props = type(self).properties
        if attr in props:
            return self.allocate()[props.index(attr), :self.n_slots].cpu().numpy()
        if attr in ('x_mean', 'y_mean', 'x_cen', 'y_cen', 'x_centroid', 'y_centroid'):
            with np.errstate(invalid='ignore', divide='ignore'):   # NaN for empty slots
                return getattr(self, attr[0] + '_sum') / self.count
        if attr in ('x_var', 'y_var') and 'x2_sum' in props:
            with np.errstate(invalid='ignore', divide='ignore'):
                return (getattr(self, attr[0] + '2_sum') / self.count
                        - getattr(self, attr[0] + '_mean') ** 2)
        if attr in ('x_std', 'y_std') and 'x2_sum' in props:
            return getattr(self, attr[0] + '_var') ** 0.5
        raise AttributeError(attr)


class BeamPositionMonitor(_BeamSlotMonitor):
    properties = ('count', 'x_sum', 'y_sum')


class BeamSizeMonitor(_BeamSlotMonitor):
    properties = ('count', 'x_sum', 'y_sum', 'x2_sum', 'y2_sum')


class BeamProfileMonitor(BeamElement):
    _mutation_tracked = False
    """Transverse profiles per time sample (xtrack/monitors/beam_profile_monitor.py:20-207,
    beam_profile_monitor.h:15-80): `x_intensity` / `y_intensity` of shape (sample_size, n).
    The counts live on the tracking device as two float64 tensors, like the reference's record."""
    allow_rot_and_shift = False
    behaves_like_drift = True
    allow_loss_refinement = True
    properties = ('counts_x', 'counts_y')

    def __init__(self, *, particle_id_range=None, particle_id_start=None, num_particles=None,
                 start_at_turn=None, stop_at_turn=None, frev=None, sampling_frequency=None,
                 nx=None, x_range=None, ny=None, y_range=None, n=None, range=None,
                 _device='cpu', **kwargs):
        if particle_id_range is None:
            if particle_id_start is None:
                particle_id_start = 0
            if num_particles is None:
                num_particles = -1
        elif particle_id_start is None and num_particles is None:
            particle_id_start = particle_id_range[0]
            num_particles = particle_id_range[1] - particle_id_range[0]
        else:
            raise ValueError('Parameter `particle_id_range` must not be used together with '
                             '`num_particles` and/or `particle_id_start`')
        self.particle_id_start = int(particle_id_start)
        self.num_particles = int(num_particles)
        self.start_at_turn = int(0 if start_at_turn is None else start_at_turn)
        self.stop_at_turn = int(0 if stop_at_turn is None else stop_at_turn)
        self.frev = float(1 if frev is None else frev)
        self.sampling_frequency = float(1 if sampling_frequency is None else sampling_frequency)
        if 'x_min' in kwargs:           # from_dict: the stored raster
            self.nx, self.x_min, self.dx = int(nx), float(kwargs.pop('x_min')), float(kwargs.pop('dx'))
            self.ny, self.y_min, self.dy = int(ny), float(kwargs.pop('y_min')), float(kwargs.pop('dy'))
        else:
            if nx is None:
                nx = n or 128
            if x_range is None:
                if range is None:
                    raise ValueError('Either `x_range` or `range` must be provided')
                x_range = range
            if np.isscalar(x_range):
                x_range = (-x_range / 2, x_range / 2)
            if ny is None:
                ny = n or 128
            if y_range is None:
                if range is None:
                    raise ValueError('Either `y_range` or `range` must be provided')
                y_range = range
            if np.isscalar(y_range):
                y_range = (-y_range / 2, y_range / 2)
            self.nx, self.x_min, self.dx = int(nx), float(x_range[0]), (x_range[1] - x_range[0]) / nx
            self.ny, self.y_min, self.dy = int(ny), float(y_range[0]), (y_range[1] - y_range[0]) / ny
        self.sample_size = int(round((self.stop_at_turn - self.start_at_turn)
                                     * self.sampling_frequency / self.frev))
        kwargs.pop('sample_size', None)
        self._device = torch.device(_device)
        self._data = None
        data = kwargs.pop('data', None)
        self._finish(kwargs)
        if data is not None:
            dd = self.allocate()
            dd['counts_x'][:self.sample_size * self.nx] = torch.as_tensor(
                np.asarray(data['counts_x'], dtype=np.float64))
            dd['counts_y'][:self.sample_size * self.ny] = torch.as_tensor(
                np.asarray(data['counts_y'], dtype=np.float64))

    @classmethod
    def from_dict(cls, dct):
        dct = dict(dct)
        for kk in ('__class__', '_index'):
            dct.pop(kk, None)
        return cls(**dct)

    def to_dict(self):
        return {'__class__': 'BeamProfileMonitor', 'particle_id_start': self.particle_id_start,
                'num_particles': self.num_particles, 'start_at_turn': self.start_at_turn,
                'stop_at_turn': self.stop_at_turn, 'frev': self.frev,
                'sampling_frequency': self.sampling_frequency, 'nx': self.nx, 'x_min': self.x_min,
                'dx': self.dx, 'ny': self.ny, 'y_min': self.y_min, 'dy': self.dy,
                'sample_size': self.sample_size,
                'data': {'counts_x': self.counts_x.tolist(), 'counts_y': self.counts_y.tolist()}}

    def allocate(self, device=None):
        if device is not None:
            self._device = torch.device(device)
        if self._data is None:
            self._data = {nn: torch.zeros(max(self.sample_size * nb, 1), dtype=torch.float64,
                                          device=self._device)
                          for nn, nb in (('counts_x', self.nx), ('counts_y', self.ny))}
        elif self._data['counts_x'].device != self._device:
            self._data = {nn: vv.to(self._device) for nn, vv in self._data.items()}
        return self._data

    @property
    def counts_x(self):
        return self.allocate()['counts_x'][:self.sample_size * self.nx].cpu().numpy()

    @property
    def counts_y(self):
        return self.allocate()['counts_y'][:self.sample_size * self.ny].cpu().numpy()

    @property
    def x_edges(self):
        return self.x_min + self.dx * np.arange(self.nx + 1)

    @property
    def x_grid(self):
        return self.x_edges[1:] - self.dx / 2

    @property
    def x_intensity(self):
        return self.counts_x.reshape((-1, self.nx))

    @property
    def y_edges(self):
        return self.y_min + self.dy * np.arange(self.ny + 1)

    @property
    def y_grid(self):
        return self.y_edges[1:] - self.dy / 2

    @property
    def y_intensity(self):
        return self.counts_y.reshape((-1, self.ny))


# ---- BeamStatsMonitor (xtrack/monitors/beam_stats_monitor/beam_stats_monitor.py:120-1330, ----
# ---- kernel part monitors/beam_stats_monitor.h:11-419) -------------------------------------
_BSM_COORDS = ('x', 'px', 'y', 'py', 'zeta', 'delta', 'pzeta')
_BSM_FIRST = _BSM_COORDS + ('charge_ratio', 'mass_ratio')
_BSM_SECOND = tuple(f'{a}_{b}' for a, b in (
    ('x', 'x'), ('x', 'px'), ('x', 'y'), ('x', 'py'), ('x', 'zeta'), ('x', 'delta'), ('x', 'pzeta'),
    ('px', 'px'), ('px', 'y'), ('px', 'py'), ('px', 'zeta'), ('px', 'delta'), ('px', 'pzeta'),
    ('y', 'y'), ('y', 'py'), ('y', 'zeta'), ('y', 'delta'), ('y', 'pzeta'),
    ('py', 'py'), ('py', 'zeta'), ('py', 'delta'), ('py', 'pzeta'),
    ('zeta', 'zeta'), ('zeta', 'delta'), ('zeta', 'pzeta'), ('delta', 'delta'), ('pzeta', 'pzeta')))
# the record arrays, in the order the device descriptor lists their addresses
BSM_RAW_FIELDS = (('num_particles', 'sum_beta0_gamma0') + tuple(f'sum_{cc}' for cc in _BSM_FIRST)
                  + tuple(f'sum_{mm}' for mm in _BSM_SECOND))
_BSM_PLANES = {'x': ('x', 'px'), 'y': ('y', 'py'), 'zeta': ('zeta', 'pzeta')}
_BSM_DEFAULT_STATS = ('num_particles', 'mean_x', 'mean_y', 'sigma_x', 'sigma_y')
BSM_DESC_HEADER = 17          # descriptor words in front of the field addresses (csrc/xtb_interp.cuh)


def _bsm_pair(rest):
    """'x_px' -> ('x', 'px'); coordinate names never contain an underscore."""
    aa, bb = rest.split('_')
    if aa not in _BSM_COORDS or bb not in _BSM_COORDS:
        raise ValueError(f'Unknown coordinate pair `{rest}`')
    return aa, bb


def _bsm_moment_name(c1, c2):
    if f'{c1}_{c2}' in _BSM_SECOND:
        return f'{c1}_{c2}'
    if f'{c2}_{c1}' in _BSM_SECOND:
        return f'{c2}_{c1}'
    raise ValueError(f'No second moment of `{c1}`, `{c2}`')


def _bsm_moments_for_stat(name):
    if name == 'num_particles':
        return ()
    if name in ('sum_charge_ratio', 'sum_mass_ratio'):
        return (name[4:],)
    kind, _, rest = name.partition('_')
    if kind == 'mean' and rest in _BSM_FIRST:
        return (rest,)
    if kind == 'sigma' and rest in _BSM_COORDS:
        return (rest, _bsm_moment_name(rest, rest))
    if kind == 'cov':
        c1, c2 = _bsm_pair(rest)
        return (c1, c2, _bsm_moment_name(c1, c2))
    if kind in ('gemitt', 'nemitt') and rest.endswith('_projected'):
        plane = rest[:-len('_projected')]
        if plane in _BSM_PLANES:
            cc, pp = _BSM_PLANES[plane]
            return (cc, pp, _bsm_moment_name(cc, cc), _bsm_moment_name(pp, pp),
                    _bsm_moment_name(cc, pp))
    raise ValueError(f'Unsupported statistic `{name}` (the covariance-optics statistics of the '
                     'reference are host-side analysis outside the tracking path)')


class BeamStatsMonitor(BeamElement):
    """Weighted beam statistics per logged turn — for the whole beam, per bunch slot, per
    longitudinal slice of each bunch, or per slice of a full turn (coasting beam) — accumulated
    IN THE KERNEL as primitive moments (sum of weights, weighted sums of the coordinates and of
    their products): turn-by-turn beam sizes and emittances without 240-byte particle records.

    Constructor, modes, public statistics and `get(stat, level=, turn=, slot=, slice_index=)`
    follow xtrack's BeamStatsMonitor (beam_stats_monitor.py:120-330, 832-905); statistics:
    `num_particles`, `sum_charge_ratio`, `sum_mass_ratio`, `mean_<c>`, `sigma_<c>`,
    `cov_<c1>_<c2>`, `gemitt_<plane>_projected`, `nemitt_<plane>_projected`, optional weighted
    `profiles`.  Not provided: the covariance-optics statistics and the HDF5 output (host-side
    analysis).  The moments live on the tracking device; lanes of a warp that fall into the
    same bin are summed in the warp before ONE lane adds to memory (`beam_stats_record`)."""
    _mutation_tracked = False
    allow_rot_and_shift = False
    behaves_like_drift = True
    allow_loss_refinement = True

    def __init__(self, *, start_at_turn=0, stop_at_turn=None, every_n_turns=1, zeta_range=None,
                 num_slices=None, bunch_spacing_zeta=None, num_bunches=None, filling_scheme=None,
                 filled_slots=None, selected_slots=None, coasting=False, particle_id_range=None,
                 stats=None, profiles=None, _device='cpu', **kwargs):
        coasting = bool(coasting)
        if coasting and num_slices is None:
            raise ValueError('`num_slices` must be provided in coasting mode')
        if coasting and zeta_range is not None:
            raise ValueError('`zeta_range` cannot be used in coasting mode')
        slice_mode = (zeta_range is not None or num_slices is not None) and not coasting
        if (not coasting) and (zeta_range is None) != (num_slices is None):
            raise ValueError('`zeta_range` and `num_slices` must be provided together')
        if stop_at_turn is None:
            stop_at_turn = start_at_turn + 1
        if every_n_turns <= 0:
            raise ValueError('`every_n_turns` must be positive')
        if particle_id_range is None:
            pid_start, pid_stop = -1, -1
        else:
            pid_start, pid_stop = (int(v) for v in particle_id_range)
            if pid_start < 0 or pid_stop < pid_start:
                raise ValueError('`particle_id_range` must be (start, stop) with 0 <= start <= stop')
        stats = tuple(dict.fromkeys(_BSM_DEFAULT_STATS if stats is None else stats))
        any_bunch_input = (filled_slots is not None or filling_scheme is not None
                           or selected_slots is not None or bunch_spacing_zeta is not None
                           or num_bunches is not None)
        bunch_mode = (not slice_mode) and (not coasting) and any_bunch_input
        if coasting and any_bunch_input:
            raise ValueError('Bunched-beam filling inputs cannot be used in coasting mode')

        if coasting:
            filled = np.array([0], dtype=np.int64)
            selected = np.array([0], dtype=np.int64)
        elif slice_mode or bunch_mode:
            given = [nn for nn, vv in (('`num_bunches`', num_bunches), ('`filled_slots`', filled_slots),
                                       ('`filling_scheme`', filling_scheme)) if vv is not None]
            if len(given) > 1:
                raise ValueError('Only one of `num_bunches`, `filled_slots`, and `filling_scheme` '
                                 'can be provided')
            if filling_scheme is not None:
                filled = np.nonzero(np.asarray(filling_scheme, dtype=np.int64))[0].astype(np.int64)
            elif filled_slots is not None:
                filled = np.asarray(filled_slots, dtype=np.int64)
            elif num_bunches is not None:
                filled = np.arange(int(num_bunches), dtype=np.int64)
            else:
                filled = np.array([0], dtype=np.int64)
            if len(filled) == 0:
                raise ValueError('At least one filled slot is required')
            if len(np.unique(filled)) != len(filled):
                raise ValueError('`filled_slots` cannot contain duplicates')
            selected = (filled.copy() if selected_slots is None
                        else np.asarray(selected_slots, dtype=np.int64))
            if len(np.unique(selected)) != len(selected):
                raise ValueError('`selected_slots` cannot contain duplicates')
            missing = [int(ss) for ss in selected if ss not in set(filled.tolist())]
            if missing:
                raise ValueError(f'`selected_slots` contains unfilled slots: {missing}')
            if len(selected) > 1 and bunch_spacing_zeta is None:
                raise ValueError('`bunch_spacing_zeta` must be provided when more than one slot '
                                 'is selected')
            if bunch_spacing_zeta is None and (len(filled) != 1 or filled[0] != 0
                                               or len(selected) != 1 or selected[0] != 0):
                raise ValueError('`bunch_spacing_zeta` is required unless the monitor uses only '
                                 'physical slot 0')
            if bunch_spacing_zeta is not None and float(bunch_spacing_zeta) <= 0:
                raise ValueError('`bunch_spacing_zeta` must be positive')
        else:
            filled = np.array([], dtype=np.int64)
            selected = np.array([], dtype=np.int64)
        if len(selected) and int(selected.min()) < 0:
            raise ValueError('Slot numbers must be non-negative')
        slot_to_selected = np.full(int(selected.max()) + 1 if len(selected) else 0, -1, dtype=np.int64)
        for ii, ss in enumerate(selected):
            slot_to_selected[int(ss)] = ii

        self.start_at_turn, self.stop_at_turn = int(start_at_turn), int(stop_at_turn)
        self.every_n_turns = int(every_n_turns)
        self._num_records = len(self.turns)
        self._bunch_spacing_zeta = 0.0 if bunch_spacing_zeta is None else float(bunch_spacing_zeta)
        if coasting:
            self._mode, self._num_selected_slots, self._num_slices = 3, 1, int(num_slices)
            self._z_min_edge, self._dzeta = 0.0, 1.0 / int(num_slices)
            self._available_levels, self._default_level = ('beam', 'slice'), 'slice'
        elif slice_mode:
            self._mode, self._num_selected_slots, self._num_slices = 2, len(selected), int(num_slices)
            self._z_min_edge = float(zeta_range[0])
            self._dzeta = (float(zeta_range[1]) - float(zeta_range[0])) / int(num_slices)
            self._available_levels, self._default_level = ('beam', 'bunch', 'slice'), 'slice'
        elif bunch_mode:
            self._mode, self._num_selected_slots, self._num_slices = 1, len(selected), 1
            self._z_min_edge = -0.5 * self._bunch_spacing_zeta
            self._dzeta = 0.0
            self._available_levels, self._default_level = ('beam', 'bunch'), 'bunch'
        else:
            self._mode, self._num_selected_slots, self._num_slices = 0, 0, 0
            self._z_min_edge = self._dzeta = 0.0
            self._available_levels, self._default_level = ('beam',), 'beam'
        if (slice_mode or coasting) and self._num_slices <= 0:
            raise ValueError('`num_slices` must be positive')
        if slice_mode and self._dzeta <= 0:
            raise ValueError('`zeta_range` must be increasing')
        self._data_shape = {0: (self._num_records,),
                            1: (self._num_records, self._num_selected_slots)}.get(
            self._mode, (self._num_records, self._num_selected_slots, self._num_slices))
        self._particle_id_start, self._particle_id_stop = pid_start, pid_stop
        self._selected_slots, self._filled_slots = selected, filled
        self._slot_to_selected = slot_to_selected
        self._stats_names = stats
        moments = set()
        for nn in stats:
            moments.update(_bsm_moments_for_stat(nn))
        self._moment_names = ('num_particles',) + tuple(sorted(moments))
        self._needed_fields = {'num_particles', 'sum_beta0_gamma0'} | {f'sum_{mm}' for mm in moments}
        self._profile_config = {}
        for coord, cfg in (profiles or {}).items():
            if coord not in _BSM_COORDS:
                raise ValueError(f'Unknown profile coordinate `{coord}`')
            lo, hi = (float(v) for v in cfg['range'])
            nb = int(cfg['num_bins'])
            if not (hi > lo) or nb <= 0:
                raise ValueError(f'Profile `{coord}`: `range` must be increasing, `num_bins` positive')
            self._profile_config[coord] = dict(range=(lo, hi), num_bins=nb)
        self._device = torch.device(_device)
        self._store = None
        kwargs.pop('output_file', None)
        self._finish(kwargs)

    # -- storage on the tracking device ---------------------------------------------------
    @property
    def _flat_size(self):
        return int(np.prod(self._data_shape, dtype=np.int64))

    def allocate(self, device=None):
        """The moments (one float64 tensor, a row per recorded field), the touched-record flags,
        the profile counts and the descriptor the kernel reads (layout: `beam_stats_record`,
        csrc/xtb_interp.cuh).  Returns the descriptor tensor."""
        if device is not None:
            self._device = torch.device(device)
        if self._store is not None and self._store['moments'].device == self._device:
            return self._store['desc']
        old = self._store
        dev = self._device
        rows = [ff for ff in BSM_RAW_FIELDS if ff in self._needed_fields]
        st = dict(rows=rows,
                  moments=torch.zeros((len(rows), max(self._flat_size, 1)), dtype=torch.float64, device=dev),
                  touched=torch.zeros(max(self._num_records, 1), dtype=torch.int64, device=dev))
        offsets, total = [], 0
        for cfg in self._profile_config.values():
            offsets.append(total)
            total += self._flat_size * cfg['num_bins']
        st['profile_counts'] = torch.zeros(max(total, 1), dtype=torch.float64, device=dev)
        if old is not None:
            for kk in ('moments', 'touched', 'profile_counts'):
                st[kk].copy_(old[kk].to(dev))
        n_s2s = len(self._slot_to_selected)
        words = np.zeros(BSM_DESC_HEADER + len(BSM_RAW_FIELDS) + n_s2s + 5 * len(offsets), dtype=np.int64)
        fwords = words.view(np.float64)
        words[0:7] = (self.start_at_turn, self.stop_at_turn, self.every_n_turns, self._mode,
                      self._num_records, self._num_selected_slots, self._num_slices)
        fwords[7:10] = (self._z_min_edge, self._dzeta, self._bunch_spacing_zeta)
        words[10:14] = (self._particle_id_start, self._particle_id_stop, n_s2s,
                        int(self._selected_slots[0]) if len(self._selected_slots) else 0)
        words[14] = st['touched'].data_ptr()
        words[15] = len(offsets)
        words[16] = st['profile_counts'].data_ptr()
        for ii, ff in enumerate(BSM_RAW_FIELDS):
            words[BSM_DESC_HEADER + ii] = (st['moments'][rows.index(ff)].data_ptr()
                                           if ff in self._needed_fields else 0)
        pos = BSM_DESC_HEADER + len(BSM_RAW_FIELDS)
        words[pos:pos + n_s2s] = self._slot_to_selected
        pos += n_s2s
        for off, (coord, cfg) in zip(offsets, self._profile_config.items()):
            words[pos:pos + 3] = (off, cfg['num_bins'], _BSM_COORDS.index(coord))
            fwords[pos + 3] = cfg['range'][0]
            fwords[pos + 4] = (cfg['range'][1] - cfg['range'][0]) / cfg['num_bins']
            pos += 5
        st['desc'] = torch.from_numpy(words).to(dev)
        self._store = st
        return st['desc']

    def _raw(self, field):
        st = self._store if self._store is not None else (self.allocate(), self._store)[1]
        if field not in st['rows']:
            return np.zeros(0)
        return st['moments'][st['rows'].index(field), :self._flat_size].cpu().numpy()

    # -- configuration as the reference exposes it -----------------------------------------
    @property
    def stats(self):
        return self._stats_names

    @property
    def turns(self):
        return np.arange(self.start_at_turn, self.stop_at_turn, self.every_n_turns, dtype=np.int64)

    @property
    def selected_slots(self):
        return self._selected_slots.copy()

    @property
    def filled_slots(self):
        return self._filled_slots.copy()

    @property
    def coasting(self):
        return self._mode == 3

    @property
    def particle_id_range(self):
        return None if self._particle_id_start < 0 else (self._particle_id_start, self._particle_id_stop)

    @property
    def available_levels(self):
        return self._available_levels

    @property
    def default_level(self):
        return self._default_level

    @property
    def zeta_centers(self):
        if 'slice' not in self._available_levels or self.coasting:
            return None
        base = self._z_min_edge + (np.arange(self._num_slices) + 0.5) * self._dzeta
        return base[None, :] - self._selected_slots[:, None] * self._bunch_spacing_zeta

    @property
    def touched_records(self):
        self.allocate()
        return self._store['touched'][:self._num_records].cpu().numpy()

    @property
    def profile_coordinates(self):
        return tuple(self._profile_config)

    @property
    def profile_bin_edges(self):
        return {cc: cfg['range'][0] + (cfg['range'][1] - cfg['range'][0]) / cfg['num_bins']
                * np.arange(cfg['num_bins'] + 1) for cc, cfg in self._profile_config.items()}

    @property
    def profiles(self):
        self.allocate()
        counts = self._store['profile_counts'].cpu().numpy()
        out, off = {}, 0
        for cc, cfg in self._profile_config.items():
            nn = self._flat_size * cfg['num_bins']
            arr = counts[off:off + nn].reshape((*self._data_shape, cfg['num_bins']))
            out[cc] = arr[:, 0, :, :] if self.coasting else arr
            off += nn
        return out

    def to_dict(self, **kwargs):
        out = {'__class__': type(self).__name__, 'start_at_turn': self.start_at_turn,
               'stop_at_turn': self.stop_at_turn, 'every_n_turns': self.every_n_turns,
               'stats': list(self._stats_names)}
        if self.particle_id_range is not None:
            out['particle_id_range'] = self.particle_id_range
        if self.coasting:
            out['coasting'] = True
            out['num_slices'] = self._num_slices
        if 'slice' in self._available_levels and not self.coasting:
            out['zeta_range'] = (self._z_min_edge, self._z_min_edge + self._dzeta * self._num_slices)
            out['num_slices'] = self._num_slices
        if 'bunch' in self._available_levels and not self.coasting:
            out['filled_slots'] = self._filled_slots.tolist()
            out['selected_slots'] = self._selected_slots.tolist()
            if self._bunch_spacing_zeta > 0:
                out['bunch_spacing_zeta'] = self._bunch_spacing_zeta
        if self._profile_config:
            out['profiles'] = {cc: dict(cfg) for cc, cfg in self._profile_config.items()}
        return out

    @classmethod
    def from_dict(cls, dct):
        dct = {kk: vv for kk, vv in dct.items() if kk not in ('__class__', '_index')}
        return cls(**dct)

    def __getattr__(self, attr):
        if attr.startswith('_'):
            raise AttributeError(attr)
        if attr in self.__dict__.get('_stats_names', ()):
            return self.get(attr)
        raise AttributeError(attr)

    # -- statistics from the primitive moments (beam_stats_monitor.py:832-1165) -------------
    def _moments_at_level(self, level):
        mm = {}
        for name in self._moment_names + ('sum_beta0_gamma0',):
            field = name if name in ('num_particles', 'sum_beta0_gamma0') else f'sum_{name}'
            mm[name] = self._raw(field).reshape(self._data_shape)
        if level == self._default_level:
            return mm
        out = {}
        for name, value in mm.items():
            if self._default_level == 'slice':
                out[name] = np.sum(value, axis=2) if level == 'bunch' else np.sum(value, axis=(1, 2))
            elif self._default_level == 'bunch' and level == 'beam':
                out[name] = np.sum(value, axis=1)
            else:
                out[name] = value
        return out

    @staticmethod
    def _mean(coord, mm):
        ww = mm['num_particles']
        out = np.zeros_like(ww, dtype=float)
        np.divide(mm[coord], ww, out=out, where=ww > 0)
        return out

    @classmethod
    def _cov(cls, c1, c2, mm):
        ww = mm['num_particles']
        out = np.zeros_like(ww, dtype=float)
        np.divide(mm[_bsm_moment_name(c1, c2)], ww, out=out, where=ww > 0)
        out -= cls._mean(c1, mm) * cls._mean(c2, mm)
        return out

    def _stat(self, name, mm):
        if name == 'num_particles':
            return mm['num_particles']
        if name in ('sum_charge_ratio', 'sum_mass_ratio'):
            return mm[name[4:]]
        kind, _, rest = name.partition('_')
        if kind == 'mean':
            return self._mean(rest, mm)
        if kind == 'sigma':
            return np.sqrt(np.maximum(self._cov(rest, rest, mm), 0))
        if kind == 'cov':
            return self._cov(*_bsm_pair(rest), mm)
        cc, pp = _BSM_PLANES[rest[:-len('_projected')]]
        det = self._cov(cc, cc, mm) * self._cov(pp, pp, mm) - self._cov(cc, pp, mm) ** 2
        out = np.sqrt(np.maximum(det, 0))
        if kind == 'nemitt':
            ww = mm['num_particles']
            bg = np.zeros_like(ww, dtype=float)
            np.divide(mm['sum_beta0_gamma0'], ww, out=bg, where=ww > 0)
            out = out * bg
        return out

    def _validated_level(self, level, slot, slice_index):
        if level is None:
            level = self._default_level
        elif level not in self._available_levels:
            raise ValueError(f'`level` must be one of {self._available_levels}, got {level!r}')
        if level == 'beam' and slot is not None:
            raise ValueError('`slot` cannot be used with level="beam"')
        if level == 'beam' and slice_index is not None:
            raise ValueError('`slice_index` cannot be used with level="beam"')
        if level != 'beam' and self.coasting and slot is not None:
            raise ValueError('`slot` cannot be used in coasting mode')
        if level == 'bunch' and slice_index is not None:
            raise ValueError('`slice_index` cannot be used with level="bunch"')
        return level

    @staticmethod
    def _indices_of(values, value, name):
        arr = np.asarray(value)
        idx = []
        for item in arr.reshape(-1):
            if int(item) != item:
                raise ValueError(f'`{name}` must be an integer')
            hit = np.nonzero(values == int(item))[0]
            if len(hit) == 0:
                raise ValueError(f'`{name}`={int(item)} is not recorded')
            idx.append(int(hit[0]))
        return np.array(idx, dtype=np.int64), arr.ndim == 0

    def record_index(self, turn):
        return int(self._indices_of(self.turns, turn, 'turn')[0][0])

    def slot_index(self, slot):
        return int(self._indices_of(self._selected_slots, slot, 'slot')[0][0])

    def get(self, stat, *, level=None, turn=None, slot=None, slice_index=None, keepdims=False):
        """A recorded statistic; axes (logged turn[, selected slot[, slice]]) of the requested
        aggregation `level`; scalar selectors drop their axis unless `keepdims`."""
        if stat not in self._stats_names:
            raise ValueError(f'Statistic `{stat}` is not recorded')
        level = self._validated_level(level, slot, slice_index)
        out = self._stat(stat, self._moments_at_level(level))
        if self.coasting and level == 'slice':
            out = out[:, 0, :]
        squeeze = []
        if turn is not None:
            idx, scalar = self._indices_of(self.turns, turn, 'turn')
            out = np.take(out, idx, axis=0)
            if scalar:
                squeeze.append(0)
        if level in ('bunch', 'slice') and not self.coasting and slot is not None:
            idx, scalar = self._indices_of(self._selected_slots, slot, 'slot')
            out = np.take(out, idx, axis=1)
            if scalar:
                squeeze.append(1)
        if level == 'slice' and slice_index is not None:
            if int(slice_index) != slice_index:
                raise ValueError('`slice_index` must be an integer')
            ii = int(slice_index) + (self._num_slices if slice_index < 0 else 0)
            if not 0 <= ii < self._num_slices:
                raise ValueError(f'`slice_index`={ii} is outside the recorded slice range')
            axis = 1 if self.coasting else 2
            out = np.take(out, [ii], axis=axis)
            squeeze.append(axis)
        if not keepdims:
            for axis in sorted(squeeze, reverse=True):
                out = np.squeeze(out, axis=axis)
        return out

    def start_new_frame(self, start_at_turn):
        """Clears the data and moves the same-size frame of logged turns (:1010-1027)."""
        if self.coasting:
            raise ValueError('`start_new_frame` cannot be used in coasting mode')
        self.start_at_turn = int(start_at_turn)
        self.stop_at_turn = self.start_at_turn + self._num_records * self.every_n_turns
        if self._store is not None:
            for kk in ('moments', 'touched', 'profile_counts'):
                self._store[kk].zero_()
            self._store['desc'][0:2] = torch.tensor([self.start_at_turn, self.stop_at_turn],
                                                    dtype=torch.int64, device=self._device)
