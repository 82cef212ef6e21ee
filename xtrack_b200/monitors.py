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
