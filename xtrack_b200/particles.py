"""`Particles`: the SoA beam container kept from xtrack's API.

Layout and semantics follow xtrack/particles/particles.py:
  * field list and dtypes           :26-83   (22 f64, 6 i64, 4 u32 per particle)
  * constructor / derived variables :453-728, 1805-1832, 1900-2060
  * unused slots hold -999999999    :24, 626-631
  * reorganize / sort / id range    :1198-1259, 1148, 1340-1348
  * rng seeding                     :1395-1418 (done on the GPU through the C-ABI)

Storage: one torch tensor per field (float64 / int64 / int32-viewed-as-uint32),
resident on the device given by `_device` ('cuda:N' for tracking).  The C-ABI
receives the raw device pointers (`xtb_particles_t`), i.e. the tensors *are*
the kernel's buffers -- no staging copies.
"""
import numpy as np
import torch

LAST_INVALID_STATE = -999999999
PROTON_MASS_EV = 938272088.16
ELECTRON_MASS_EV = 510998.95
CLIGHT = 299792458.0

SIZE_VARS = ('_capacity', '_num_active_particles', '_num_lost_particles',
             'start_tracking_at_element')
SCALAR_VARS = ('q0', 'mass0', 't_sim')

F64_VARS = ('p0c', 'gamma0', 'beta0', 's', 'zeta', 'x', 'y', 'px', 'py',
            'ptau', 'delta', 'rpp', 'rvv',
            'chi', 'charge_ratio', 'weight', 'ax', 'ay',
            'spin_x', 'spin_y', 'spin_z', 'anomalous_magnetic_moment')
I64_VARS = ('pdg_id', 'particle_id', 'at_element', 'at_turn', 'state',
            'parent_particle_id')
U32_VARS = ('_rng_s1', '_rng_s2', '_rng_s3', '_rng_s4')

# (name, numpy dtype) in the reference order; index = slot in xtb_particles_t.field
PER_PARTICLE_VARS = tuple(
    [(nn, np.float64) for nn in F64_VARS]
    + [(nn, np.int64) for nn in I64_VARS]
    + [(nn, np.uint32) for nn in U32_VARS])
FIELD_INDEX = {nn: ii for ii, (nn, _) in enumerate(PER_PARTICLE_VARS)}
BYTES_PER_PARTICLE = 22 * 8 + 6 * 8 + 4 * 4   # 240

_TORCH_DTYPE = {np.float64: torch.float64, np.int64: torch.int64,
                np.uint32: torch.int32}


def normalise_device(device):
    """torch.device with an explicit index for CUDA ('cuda' -> 'cuda:<current>'), so that the
    devices of particles, trackers and monitors compare equal when they are the same GPU."""
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    return device


def _as_array(v, n, dtype):
    if np.isscalar(v) or (hasattr(v, '__len__') and len(v) == 1):
        return np.full(n, np.asarray(v).reshape(-1)[0], dtype=dtype)
    out = np.asarray(v, dtype=dtype)
    if out.shape != (n,):
        raise ValueError('All per particle vars have to be of the same length.')
    return out


class Particles:

    per_particle_vars = PER_PARTICLE_VARS

    def __init__(self, _capacity=None, _device='cpu', **kwargs):
        accepted = set(nn for nn, _ in PER_PARTICLE_VARS) | set(SCALAR_VARS) | {
            'start_tracking_at_element', 'energy0', 'tau', 'pzeta', 'mass_ratio',
            'kinetic_energy0', 'rigidity0', 'name', '_num_active_particles',
            '_num_lost_particles'}
        if set(kwargs) - accepted:
            raise NameError(f'Invalid argument(s) provided: {set(kwargs) - accepted}')

        per_part_inputs = [nn for nn, _ in PER_PARTICLE_VARS] + [
            'energy0', 'kinetic_energy0', 'rigidity0', 'tau', 'pzeta', 'mass_ratio']
        n = 1
        for nn in per_part_inputs:
            if nn not in kwargs or kwargs[nn] is None:
                continue
            vv = kwargs[nn]
            if np.isscalar(vv) or len(vv) == 1:
                continue
            if len(vv) != n and n > 1:
                raise ValueError('All per particle vars have to be of the same length.')
            n = len(vv)
        if _capacity is None:
            _capacity = n
        if _capacity <= 0:
            raise ValueError('Explicitly provided `_capacity` has to be greater than zero.')
        if _capacity < n:
            raise ValueError(f'Capacity ({_capacity}) has to be greater or equal to the '
                             f'number of particles ({n}).')

        self._capacity = int(_capacity)
        self._num_active_particles = -1
        self._num_lost_particles = -1
        self.q0 = float(kwargs.get('q0', 1.0))
        self.mass0 = float(kwargs.get('mass0', PROTON_MASS_EV))
        self.t_sim = float(kwargs.get('t_sim', 0.0))
        self.start_tracking_at_element = int(kwargs.get('start_tracking_at_element', -1))
        self.name = kwargs.get('name', None)

        h = {}   # host staging (numpy), first n entries valid
        get = lambda nn, default=None: kwargs.get(nn, default)
        arr = lambda nn, default, dt=np.float64: _as_array(get(nn, default), n, dt)

        h['state'] = arr('state', 1, np.int64)
        h['particle_id'] = arr('particle_id', np.arange(n) if n > 1 else 0, np.int64)
        h['parent_particle_id'] = _as_array(
            get('parent_particle_id', h['particle_id']), n, np.int64)
        for nn, dflt in (('s', 0), ('weight', 1), ('x', 0), ('y', 0), ('px', 0),
                         ('py', 0), ('ax', 0), ('ay', 0),
                         ('anomalous_magnetic_moment', 0), ('spin_x', 0),
                         ('spin_y', 0), ('spin_z', 0)):
            h[nn] = arr(nn, dflt)
        h['at_turn'] = arr('at_turn', 0, np.int64)
        h['at_element'] = arr('at_element', 0, np.int64)
        h['pdg_id'] = arr('pdg_id', 0, np.int64)

        # chi / charge_ratio / mass_ratio (particles.py:2030-2060)
        chi, cr, mr = get('chi'), get('charge_ratio'), get('mass_ratio')
        nargs = sum(v is not None for v in (chi, cr, mr))
        if nargs == 0:
            h['chi'] = np.ones(n)
            h['charge_ratio'] = np.ones(n)
        elif nargs == 1:
            raise ValueError('Two of `chi`, `charge_ratio` and `mass_ratio` must be provided.')
        else:
            if chi is None:
                cr, mr = arr('charge_ratio', 1), arr('mass_ratio', 1)
                chi = cr / mr
            elif cr is None:
                chi, mr = arr('chi', 1), arr('mass_ratio', 1)
                cr = chi * mr
            h['chi'] = _as_array(chi, n, np.float64)
            h['charge_ratio'] = _as_array(cr, n, np.float64)
        mass_ratio = h['charge_ratio'] / h['chi']

        # reference momentum (particles.py:1900-1954)
        m0 = self.mass0
        p0c, energy0, gamma0, beta0 = (get('p0c'), get('energy0'), get('gamma0'),
                                       get('beta0'))
        kin0, rig0 = get('kinetic_energy0'), get('rigidity0')
        if all(v is None for v in (p0c, energy0, gamma0, beta0, kin0, rig0)):
            p0c = 1e9
        if p0c is not None:
            _p0c = arr('p0c', p0c)
            _e0 = np.sqrt(_p0c ** 2 + m0 ** 2)
            _beta0 = _p0c / _e0
            _gamma0 = _e0 / m0
        elif energy0 is not None:
            _e0 = arr('energy0', None)
            _p0c = np.sqrt(_e0 ** 2 - m0 ** 2)
            _beta0 = _p0c / _e0
            _gamma0 = _e0 / m0
        elif gamma0 is not None:
            _gamma0 = arr('gamma0', None)
            _beta0 = np.sqrt(1 - 1 / _gamma0 ** 2)
            _p0c = m0 * _gamma0 * _beta0
        elif beta0 is not None:
            _beta0 = arr('beta0', None)
            _gamma0 = 1 / np.sqrt(1 - _beta0 ** 2)
            _p0c = m0 * _gamma0 * _beta0
        elif kin0 is not None:
            _e0 = arr('kinetic_energy0', None) + m0
            _p0c = np.sqrt(_e0 ** 2 - m0 ** 2)
            _beta0 = _p0c / _e0
            _gamma0 = _e0 / m0
        else:
            _p0c = arr('rigidity0', None) * abs(self.q0) * CLIGHT
            _e0 = np.sqrt(_p0c ** 2 + m0 ** 2)
            _beta0 = _p0c / _e0
            _gamma0 = _e0 / m0
        # given values prevail over computed ones (`_setattr_if_consistent`)
        h['p0c'] = arr('p0c', None) if get('p0c') is not None else _p0c
        h['gamma0'] = arr('gamma0', None) if get('gamma0') is not None else _gamma0
        h['beta0'] = arr('beta0', None) if get('beta0') is not None else _beta0
        b0 = h['beta0']

        # energy deviations (particles.py:1956-2008)
        delta, ptau, pzeta = get('delta'), get('ptau'), get('pzeta')
        if all(v is None for v in (delta, ptau, pzeta)):
            if get('rpp') is not None or get('rvv') is not None:
                raise ValueError('Setting `delta` and `ptau` by only giving `_rpp` '
                                 'and `_rvv` is not supported.')
            if np.any(mass_ratio != 1.0):
                raise ValueError('Need to provide `delta` or `ptau` with non-default '
                                 'mass ratios.')
            delta = 0.0
        if delta is not None:
            _delta = arr('delta', delta)
            _ptau = np.sqrt(_delta ** 2 + 2 * _delta + 1 / b0 ** 2) - 1 / b0
        elif ptau is not None:
            _ptau = arr('ptau', None)
            _delta = np.sqrt(_ptau ** 2 + 2 * _ptau / b0 + 1) - 1
        else:
            _ptau = arr('pzeta', None) * b0
            _delta = np.sqrt(_ptau ** 2 + 2 * _ptau / b0 + 1) - 1
        h['delta'] = arr('delta', None) if get('delta') is not None else _delta
        h['ptau'] = arr('ptau', None) if get('ptau') is not None else _ptau
        d = h['delta']
        delta_beta0 = d * b0
        ptau_beta0 = np.sqrt(delta_beta0 ** 2 + 2 * delta_beta0 * b0 + 1) - 1
        h['rvv'] = (arr('rvv', None) if get('rvv') is not None
                    else (1 + d) / (1 + ptau_beta0))
        h['rpp'] = arr('rpp', None) if get('rpp') is not None else 1 / (1 + d)

        # zeta (particles.py:2010-2028)
        if get('zeta') is not None:
            h['zeta'] = arr('zeta', None)
        elif get('tau') is not None:
            h['zeta'] = b0 * arr('tau', None)
        else:
            h['zeta'] = np.zeros(n)

        for nn in U32_VARS:
            h[nn] = arr(nn, 0, np.uint32)

        self._device = normalise_device(_device)
        self._fields = {}
        for nn, dt in PER_PARTICLE_VARS:
            full = np.full(self._capacity, 0 if nn.startswith('_rng') else LAST_INVALID_STATE,
                           dtype=dt)
            full[:n] = h[nn]
            if dt is np.uint32:
                full = full.view(np.int32)
            self._fields[nn] = torch.from_numpy(full).to(self._device)

        if self._device.type == 'cpu':
            self.reorganize()

    # -- access ------------------------------------------------------------
    def __getattr__(self, name):
        ff = self.__dict__.get('_fields')
        if ff is not None and name in ff:
            return ff[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        ff = self.__dict__.get('_fields')
        if ff is not None and name in ff:
            tt = ff[name]
            if torch.is_tensor(value):
                tt.copy_(value.to(tt.dtype))
            else:
                vv = np.asarray(value)
                if vv.dtype == np.uint32:
                    vv = vv.view(np.int32)
                tt.copy_(torch.as_tensor(vv).to(tt.dtype).expand_as(tt))
            return
        object.__setattr__(self, name, value)

    def get(self, name):
        """Field as a numpy array (uint32 for the rng state)."""
        vv = self._fields[name].cpu().numpy()
        if name in U32_VARS:
            vv = vv.view(np.uint32)
        return vv

    @property
    def device(self):
        return self._device

    @property
    def pzeta(self):
        return self.ptau / self.beta0

    @property
    def energy0(self):
        return torch.sqrt(self.p0c ** 2 + self.mass0 ** 2)

    # -- consistent updates of the energy variables (particles.py:1450-1560, 1956-2008) --------
    def _masked_set(self, name, new, mask):
        cur = self.get(name).astype(np.float64)
        cur[mask] = new[mask]
        setattr(self, name, cur)

    def _update_energy_deviations_host(self, *, delta=None, ptau=None):
        b0 = self.get('beta0')
        st = self.get('state')
        given = delta if delta is not None else ptau
        given = np.broadcast_to(np.asarray(given, dtype=np.float64), b0.shape)
        mask = (~np.isnan(given)) & (st > 0)
        with np.errstate(invalid='ignore'):
            if delta is not None:
                _delta = given
                _ptau = np.sqrt(_delta ** 2 + 2 * _delta + 1 / b0 ** 2) - 1 / b0
            else:
                _ptau = given
                _delta = np.sqrt(_ptau ** 2 + 2 * _ptau / b0 + 1) - 1
            delta_beta0 = _delta * b0
            ptau_beta0 = np.sqrt(delta_beta0 ** 2 + 2 * delta_beta0 * b0 + 1) - 1
            new_rvv = (1 + _delta) / (1 + ptau_beta0)
            new_rpp = 1 / (1 + _delta)
        self._masked_set('delta', _delta, mask)
        self._masked_set('ptau', _ptau, mask)
        self._masked_set('rvv', new_rvv, mask)
        self._masked_set('rpp', new_rpp, mask)

    def update_delta(self, new_delta_value):
        """`delta` of the active particles; `ptau`, `rvv`, `rpp` follow (NaN entries are left
        alone), particles.py:1450-1465."""
        self._update_energy_deviations_host(delta=new_delta_value)

    def update_ptau(self, new_ptau):
        """particles.py:1486-1501."""
        self._update_energy_deviations_host(ptau=new_ptau)

    def update_p0c(self, new_p0c):
        """`p0c` of the active particles; `gamma0`, `beta0` follow (particles.py:1522-1537,
        `_update_refs`); the energy deviations are left as they are, as in the reference."""
        new_p0c = np.broadcast_to(np.asarray(new_p0c, dtype=np.float64), self.get('p0c').shape)
        mask = (~np.isnan(new_p0c)) & (self.get('state') > 0)
        m0 = self.mass0
        with np.errstate(invalid='ignore'):
            e0 = np.sqrt(new_p0c ** 2 + m0 ** 2)
            beta0 = new_p0c / e0
            gamma0 = e0 / m0
        self._masked_set('p0c', new_p0c, mask)
        self._masked_set('beta0', beta0, mask)
        self._masked_set('gamma0', gamma0, mask)

    # -- derived quantities (read only; particles.py:1600-1900) ---------------------------
    @property
    def mass_ratio(self):
        return self.charge_ratio / self.chi

    @property
    def energy(self):
        return (self.energy0 + self.ptau * self.p0c) * self.mass_ratio

    @property
    def kinetic_energy0(self):
        return self.energy0 - self.mass0

    @property
    def rigidity0(self):
        return self.p0c / (abs(self.q0) * CLIGHT)

    @property
    def kin_px(self):
        return self.px - self.ax

    @property
    def kin_py(self):
        return self.py - self.ay

    @property
    def kin_ps(self):
        return torch.sqrt((1 + self.delta) ** 2 - self.kin_px ** 2 - self.kin_py ** 2)

    @property
    def kin_xprime(self):
        return self.kin_px / self.kin_ps

    @property
    def kin_yprime(self):
        return self.kin_py / self.kin_ps

    # -- subsets / unions (particles.py:1002-1130, 1280-1330) ------------------------------
    def filter(self, mask):
        """New Particles with the slots where `mask` is true (same device)."""
        mask = torch.as_tensor(np.asarray(mask.cpu() if torch.is_tensor(mask) else mask),
                               dtype=torch.bool)
        idx = torch.nonzero(mask, as_tuple=False).flatten().to(self._device)
        out = object.__new__(Particles)
        for kk, vv in self.__dict__.items():
            if kk != '_fields':
                object.__setattr__(out, kk, vv)
        object.__setattr__(out, '_fields', {nn: tt[idx].clone() for nn, tt in self._fields.items()})
        object.__setattr__(out, '_capacity', int(idx.numel()))
        return out

    def remove_unused_space(self):
        """particles.py:1280-1288."""
        return self.filter(self.get('state') > LAST_INVALID_STATE)

    @classmethod
    def merge(cls, lst, _device=None):
        """One Particles object out of several (same reference charge and mass); particle ids
        are made unique as in the reference (particles.py:1002-1088): an object whose ids
        collide with the ones before it is shifted behind them."""
        first = lst[0]
        dev = normalise_device(_device) if _device is not None else first._device
        for pp in lst[1:]:
            if pp.q0 != first.q0 or pp.mass0 != first.mass0:
                raise ValueError('Cannot merge particles with different q0 / mass0')
        parts = [pp.remove_unused_space() for pp in lst]
        out = parts[0].filter(np.ones(parts[0]._capacity, dtype=bool))
        fields = {nn: [pp._fields[nn].to(dev) for pp in parts] for nn in out._fields}
        next_id = 0
        for ii, pp in enumerate(parts):
            ids = fields['particle_id'][ii]
            if ii > 0 and ids.numel() and int(ids.min()) < next_id:
                shift = next_id - int(ids.min())
                fields['particle_id'][ii] = ids + shift
                par = fields['parent_particle_id'][ii]
                fields['parent_particle_id'][ii] = par + shift
                ids = fields['particle_id'][ii]
            if ids.numel():
                next_id = max(next_id, int(ids.max()) + 1)
        object.__setattr__(out, '_fields', {nn: torch.cat(vv) for nn, vv in fields.items()})
        object.__setattr__(out, '_capacity', int(out._fields['x'].numel()))
        object.__setattr__(out, '_device', dev)
        return out

    def add_particles(self, part, keep_lost=False):
        """particles.py:1291-1330 (returns nothing: this object grows)."""
        other = part if keep_lost else part.filter(part.get('state') > 0)
        merged = Particles.merge([self, other], _device=self._device)
        object.__setattr__(self, '_fields', merged._fields)
        object.__setattr__(self, '_capacity', merged._capacity)

    def to(self, device):
        """Move the SoA to `device` (host<->device copy of all 32 fields)."""
        new = object.__new__(Particles)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k != '_fields'})
        device = normalise_device(device)
        new.__dict__['_device'] = device
        new.__dict__['_fields'] = {nn: tt.to(device) for nn, tt in self._fields.items()}
        return new

    def copy(self, _device=None):
        new = object.__new__(Particles)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k != '_fields'})
        dev = self._device if _device is None else normalise_device(_device)
        new.__dict__['_device'] = dev
        new.__dict__['_fields'] = {nn: tt.detach().clone().to(dev)
                                   for nn, tt in self._fields.items()}
        return new

    # -- bookkeeping -------------------------------------------------------
    def reorganize(self):
        """Active first, then lost, then unallocated (particles.py:1198-1259;
        the rng state does not travel, as in the reference)."""
        state = self._fields['state']
        mask_active = state > 0
        mask_lost = (state < 1) & (state > LAST_INVALID_STATE)
        n_active = int(mask_active.sum())
        n_lost = int(mask_lost.sum())
        if not bool(mask_active[:n_active].all()):
            for nn, dt in PER_PARTICLE_VARS:
                if nn.startswith('_rng'):
                    continue
                vv = self._fields[nn]
                va, vl = vv[mask_active].clone(), vv[mask_lost].clone()
                vv[:n_active] = va
                vv[n_active:n_active + n_lost] = vl
                vv[n_active + n_lost:] = LAST_INVALID_STATE
        self._num_active_particles = n_active
        self._num_lost_particles = n_lost
        return n_active, n_lost

    def sort(self, by='particle_id', interleave_lost_particles=False):
        """Order allocated particles by `by` (particles.py:1148-1196); all fields,
        including the rng state, move together."""
        state = self._fields['state']
        key = self._fields[by]
        allocated = state > LAST_INVALID_STATE
        n_alloc = int(allocated.sum())
        if interleave_lost_particles:
            big = torch.iinfo(torch.int64).max if key.dtype == torch.int64 else float('inf')
            kk = torch.where(allocated, key, torch.full_like(key, big))
            order = torch.argsort(kk, stable=True)
        else:
            grp = torch.where(state > 0, 0, torch.where(allocated, 1, 2))
            o1 = torch.argsort(key, stable=True)
            order = o1[torch.argsort(grp[o1], stable=True)]
        for nn in self._fields:
            self._fields[nn].copy_(self._fields[nn][order].clone())
        self._num_active_particles = int((state > 0).sum())
        self._num_lost_particles = n_alloc - self._num_active_particles

    def get_active_particle_id_range(self):
        state = self._fields['state']
        ids = self._fields['particle_id'][state > 0]
        return int(ids.min()), int(ids.max()) + 1

    def _has_valid_rng_state(self):
        # particles.py:1381-1393
        state = self._fields['state']
        ok = torch.ones_like(state, dtype=torch.bool)
        zero = torch.ones_like(state, dtype=torch.bool)
        for nn in U32_VARS:
            zero &= self._fields[nn] == 0
        return not bool((zero & (state > 0)).any())

    def _init_random_number_generator(self, seeds=None, mode='tausworthe'):
        """Seeds the per-particle generator.
        mode 'tausworthe': the reference's generator, seeded on the GPU with `xtb_rng_init`
        (particles.py:1395-1418 + rng_src/particles_rng.h:12-28) -- the parity mode;
        mode 'philox': the counter-based production generator (csrc/xtb_rng.cuh): the state
        words become key = (seed, particle_id) and a zero draw counter."""
        from . import _cabi
        if mode not in ('tausworthe', 'philox'):
            raise ValueError(f'unknown generator mode {mode!r}')
        if seeds is None:
            seeds = np.random.randint(low=1, high=4e9, size=self._capacity,
                                      dtype=np.uint32)
        else:
            assert len(seeds) == self._capacity
            seeds = np.asarray(seeds, dtype=np.uint32)
        if mode == 'philox':
            if np.any(seeds == 0):
                raise ValueError('seeds must not be zero')
            self._rng_s1 = seeds
            self._rng_s2 = (self.get('particle_id') & 0xffffffff).astype(np.uint32)
            self._rng_s3 = np.zeros(self._capacity, dtype=np.uint32)
            self._rng_s4 = np.zeros(self._capacity, dtype=np.uint32)
        else:
            _cabi.rng_init(self, seeds)
        object.__setattr__(self, '_rng_mode', mode)

    # -- (de)serialisation (particles.py:734-852) ----------------------------
    def to_dict(self):
        out = {nn: getattr(self, nn) for nn in SCALAR_VARS}
        out['start_tracking_at_element'] = self.start_tracking_at_element
        for nn, _ in PER_PARTICLE_VARS:
            out[nn] = self.get(nn)
        return out

    def to_pandas(self):
        """One row per particle slot, scalars repeated (particles.py:913-954)."""
        import pandas as pd
        return pd.DataFrame(self.to_dict())

    @classmethod
    def from_pandas(cls, df, load_rng_state=True, _device='cpu'):
        """particles.py:883-911."""
        dct = df.to_dict(orient='list')
        for nn in list(SCALAR_VARS) + ['start_tracking_at_element']:
            if nn in dct and not np.isscalar(dct[nn]):
                dct[nn] = dct[nn][0]
        return cls.from_dict(dct, load_rng_state=load_rng_state, _device=_device)

    @classmethod
    def from_dict(cls, dct, load_rng_state=True, _device='cpu', _capacity=None):
        dct = {k: v for k, v in dct.items()
               if k not in ('__class__', 'name', '_capacity', '_num_active_particles',
                            '_num_lost_particles')}
        state = np.atleast_1d(np.asarray(dct.get('state', 1)))
        if state.size > 1:
            keep = state > LAST_INVALID_STATE
            for kk, vv in list(dct.items()):
                if hasattr(vv, '__len__') and len(vv) == state.size:
                    dct[kk] = np.asarray(vv)[keep]
        if not load_rng_state:
            for nn in U32_VARS:
                dct.pop(nn, None)
        for kk in ('energy0', 'mass_ratio', 'tau', 'pzeta', 'kinetic_energy0', 'rigidity0'):
            dct.pop(kk, None)
        return cls(_capacity=_capacity, _device=_device, **dct)

    def field_pointers(self):
        """Device addresses in the `xtb_particles_t.field[32]` order."""
        return [int(self._fields[nn].data_ptr()) for nn, _ in PER_PARTICLE_VARS]
