"""`Tracker`: host orchestration of a `Line.track` call.

Mirrors the non-collective path of the reference tracker:
  Tracker.__init__                  xtrack/tracker.py:38-147   -> lattice lowering + upload
  Tracker._track_no_collective      xtrack/tracker.py:1200-1443 (turn splitting :1270-1333,
                                    end-of-turn flags :1335-1342, monitor :1344-1352,
                                    rng seeding :1364-1365, <= 3 kernel launches :1372-1436)
  Tracker._get_monitor              xtrack/tracker.py:1445-1480
The kernel launches go through the C-ABI (`_cabi.Lattice.track` -> `xtb_track`).
"""
import numpy as np
import torch

from . import _cabi
from . import lowering
from .elements import mutation_count
from .monitors import ParticlesMonitor
from .particles import normalise_device


class TurnPlan:
    """How one `track()` call is cut into kernel launches: `head` elements from `ele_start`
    (one launch), `full_turns` whole turns (one launch), `tail` elements from the start of
    the line (one launch)."""
    __slots__ = ('head', 'full_turns', 'tail', 'head_ends_turn', 'full_turns_end_turn',
                 'monitor_turns')

    def __repr__(self):
        return 'TurnPlan(' + ', '.join(f'{k}={getattr(self, k)}' for k in self.__slots__) + ')'


def split_turns(n_line, ele_start, *, ele_stop=None, num_elements=None, num_turns=None,
                skip_end_turn_actions=False):
    """Launch plan of a `track()` call; same results as the branch tree of the reference
    (xtrack/tracker.py:1270-1342), expressed on the absolute position `end` = number of
    elements from the start of turn 0 at which tracking stops."""
    if num_elements is not None:
        if num_elements < 0:
            raise ValueError('num_elements must not be negative')
        if ele_stop is not None:
            raise ValueError('Cannot use both num_elements and ele_stop!')
        if num_turns is not None:
            raise ValueError('Cannot use both num_elements and num_turns!')
        end = ele_start + num_elements
        stops_inside_first_turn = end <= n_line
        last_is_partial = True            # a remainder of `end` is an unfinished turn
    else:
        turns = 1 if num_turns is None else int(num_turns)
        if turns <= 0:
            raise ValueError('num_turns must be positive')
        if ele_stop is None:
            end = turns * n_line
        else:
            if not 0 <= ele_stop <= n_line:
                raise ValueError('ele_stop outside the line')
            if ele_stop <= ele_start:
                turns += 1                # the stop lies in the following turn
            end = (turns - 1) * n_line + ele_stop
        stops_inside_first_turn = (turns == 1)
        last_is_partial = ele_stop is not None
    plan = TurnPlan()
    if stops_inside_first_turn:
        plan.head, plan.full_turns, plan.tail = end - ele_start, 0, 0
    else:
        plan.head = n_line - ele_start
        rest = end - n_line               # elements after the first (partial) turn
        if last_is_partial and n_line > 0:
            # the stretch up to `end` of the last turn runs without end-of-turn actions,
            # even when it happens to cover the whole line (ele_stop == len(line))
            plan.tail = rest % n_line if num_elements is not None else ele_stop
            plan.full_turns = (rest - plan.tail) // n_line
        else:
            plan.tail = 0
            plan.full_turns = rest // n_line if n_line > 0 else 0
    assert plan.head >= 0
    plan.head_ends_turn = (not skip_end_turn_actions) and (ele_start + plan.head == n_line)
    plan.full_turns_end_turn = not skip_end_turn_actions
    plan.monitor_turns = plan.full_turns + (2 if plan.tail > 0 else 1)
    return plan


class Tracker:

    def __init__(self, line, device=None, exact_arithmetic=True, compact_every=None, fuse=True,
                 rng='tausworthe'):
        self.line = line
        self.device = normalise_device('cuda' if device is None else device)
        # exact_arithmetic=True (default): kernel built without FMA contraction, rounds like
        # the reference's CPU build (bit-identical wherever no libm call is involved);
        # False: FMA-contracted build, ~1e-11 relative from the reference after 10 LHC turns.
        self.exact_arithmetic = bool(exact_arithmetic)
        self.compact_every = compact_every
        # fuse=True: full-turn launches run the FUSED program (drift-prefixed fast ops,
        # csrc/xtb_ops.h); False: everything runs the PLAIN one-element-per-op program
        self.fuse = bool(fuse)
        # generator the tracker seeds unseeded particles with: 'tausworthe' (the reference's,
        # parity mode) or 'philox' (counter-based, production); particles seeded by the caller
        # carry their own mode (Particles._init_random_number_generator)
        if rng not in ('tausworthe', 'philox'):
            raise ValueError(f'unknown generator {rng!r}')
        self.rng = rng
        self.num_elements = len(line.element_names)
        self._config_key = None
        self._lattice = None
        self._back_lattice = None       # the inverse maps in reverse order, built on first use
        self._back_key = None
        self.program = None
        self._ensure_lattice()

    # -- lattice ------------------------------------------------------------
    def _current_config_key(self):
        # compile-time configuration of the reference kernel + everything the lowered op
        # stream is a frozen copy of: the element sequence and the element field values
        # (elements.mutation_count moves with every field write)
        return (not self.line.config.get('XTRACK_MULTIPOLE_NO_SYNRAD', True),
                bool(self.line.config.get('XTRACK_USE_EXACT_DRIFTS', False)),
                hash(tuple(self.line.element_names)), mutation_count())

    def _ensure_lattice(self):
        """One lowered lattice per distinct compile-time config, like the
        reference's kernel cache keyed by the config hash (tracker.py:1564-1596)."""
        key = self._current_config_key()
        if self._lattice is not None and key == self._config_key:
            return
        synrad, exact_drifts = key[:2]
        self.num_elements = len(self.line.element_names)
        prog = lowering.lower_line(self.line.elements, synrad=synrad, exact_drifts=exact_drifts,
                                   device=self.device)
        fused = prog.finish(fused=True) if self.fuse else (None, None)
        plain = prog.finish(fused=False)
        self.line_length = self.line.get_length()
        if self._lattice is not None:
            self._lattice.close()
        self._lattice = self._make_lattice(fused, plain)
        for mm in prog.monitors + prog.last_turns_monitors:
            mm.allocate(self.device)
        if prog.monitors or prog.last_turns_monitors:
            self._lattice.set_inline_monitors(prog.monitors, prog.last_turns_monitors)
        if synrad and self.line._extra_config.get('_radiation_model') == 'quantum-kick':
            self._lattice.set_synrad_tables(self._synrad_tables())
        self.program = prog
        # (allocating the in-line monitors above may have written element fields)
        self._config_key = key[:3] + (mutation_count(),)

    def _make_lattice(self, fused, plain):
        # the C-ABI handle; raises unless `self.device` is a CUDA device (no CPU fallback)
        return _cabi.Lattice(fused, plain, self.line_length, self.device)

    def _synrad_tables(self):
        # (a line may carry its own tables: `line.synrad_tables = synrad_tables.make_blob(...)`)
        blob = getattr(self.line, 'synrad_tables', None)
        if blob is None:
            from . import synrad_tables
            blob = synrad_tables.load_blob()
        return blob

    def _ensure_back_lattice(self):
        """The lattice XS_FLAG_BACKTRACK stands for: every element lowered as its inverse map
        (lowering.lower_line(backtrack=True)), in reverse order -- the kernel then runs it
        forwards like any other program."""
        key = self._current_config_key()
        if self._back_lattice is not None and key == self._back_key:
            return
        prog = lowering.lower_line(list(reversed(self.line.elements)), synrad=key[0],
                                   exact_drifts=key[1], device=self.device, backtrack=True)
        fused = prog.finish(fused=True) if self.fuse else (None, None)
        plain = prog.finish(fused=False)
        if self._back_lattice is not None:
            self._back_lattice.close()
        self._back_lattice = self._make_lattice(fused, plain)
        if prog.monitors or prog.last_turns_monitors:
            for mm in prog.monitors + prog.last_turns_monitors:
                mm.allocate(self.device)
            self._back_lattice.set_inline_monitors(prog.monitors, prog.last_turns_monitors)
        self._back_key = key[:3] + (mutation_count(),)

    # -- monitor ------------------------------------------------------------
    def _get_monitor(self, particles, turn_by_turn_monitor, num_turns):
        if turn_by_turn_monitor is None or turn_by_turn_monitor is False:
            return 0, None
        if turn_by_turn_monitor is True:
            monitor = ParticlesMonitor(
                _device=particles.device, start_at_turn=0, stop_at_turn=num_turns,
                particle_id_range=particles.get_active_particle_id_range())
            return 1, monitor
        if isinstance(turn_by_turn_monitor, str) and turn_by_turn_monitor == 'ONE_TURN_EBE':
            monitor = ParticlesMonitor(
                _device=particles.device, start_at_turn=0, stop_at_turn=self.num_elements + 1,
                particle_id_range=particles.get_active_particle_id_range(), ebe_mode=1)
            return 2, monitor
        if isinstance(turn_by_turn_monitor, ParticlesMonitor):
            return (2 if turn_by_turn_monitor.ebe_mode == 1 else 1), turn_by_turn_monitor
        raise ValueError('Please provide a valid monitor object')

    # -- tracking -----------------------------------------------------------
    def track(self, particles, ele_start=0, ele_stop=None, num_elements=None, num_turns=None,
              turn_by_turn_monitor=None, freeze_longitudinal=False, time=False,
              _force_no_end_turn_actions=False, backtrack=False):
        line = self.line
        if particles.device != self.device:
            raise ValueError(f'particles are on {particles.device}, tracker on {self.device}')
        self._ensure_lattice()
        if backtrack is False and line.track_flags.get('XS_FLAG_BACKTRACK', False):
            backtrack = 'force'         # the flag set by hand: what the reference's kernel reads
        if backtrack is not False:      # tracker.py:1222-1235
            if isinstance(backtrack, str):
                assert backtrack == 'force'
            elif not line._is_backtrackable:
                raise ValueError('This line is not backtrackable.')
            if turn_by_turn_monitor not in (None, False):
                raise NotImplementedError('turn-by-turn monitor while backtracking')
            self._ensure_back_lattice()

        # start position (tracker.py:1252-1266)
        if particles.start_tracking_at_element >= 0:
            if ele_start != 0:
                raise ValueError('The argument ele_start is used, but '
                                 'particles.start_tracking_at_element is set as well. '
                                 'Please use only one of those methods.')
            ele_start = particles.start_tracking_at_element
            particles.start_tracking_at_element = -1
        if isinstance(ele_start, str):
            ele_start = line.element_names.index(ele_start)
        if ele_start is None:
            ele_start = 0
        assert ele_start >= 0
        assert ele_start <= self.num_elements

        if isinstance(ele_stop, str):
            ele_stop = line.element_names.index(ele_stop)
        plan = split_turns(self.num_elements, ele_start, ele_stop=ele_stop,
                           num_elements=num_elements, num_turns=num_turns,
                           skip_end_turn_actions=(line.skip_end_turn_actions
                                                  or _force_no_end_turn_actions))

        flag_monitor, monitor = self._get_monitor(particles, turn_by_turn_monitor, plan.monitor_turns)
        if monitor is not None:
            monitor.allocate(self.device)

        if line._needs_rng and not particles._has_valid_rng_state():
            particles._init_random_number_generator(mode=self.rng)

        variant = 0
        if self.exact_arithmetic:
            variant |= _cabi.VARIANT_EXACT
        if not line.config.get('XTRACK_MULTIPOLE_NO_SYNRAD', True):
            variant |= _cabi.VARIANT_SYNRAD
        if freeze_longitudinal:
            variant |= _cabi.VARIANT_FREEZE_LONG
        if getattr(particles, '_rng_mode', 'tausworthe') == 'philox':
            variant |= _cabi.VARIANT_PHILOX
        common = dict(flag_reset_s_at_end_turn=line.reset_s_at_end_turn,
                      flag_monitor=flag_monitor, monitor=monitor,
                      # (backtracking is resolved here, in the lowering: the kernel never sees it)
                      track_flags=line.get_flags_register() & ~1,
                      global_xy_limit=float(line.config.get('XTRACK_GLOBAL_XY_LIMIT', 1.0)),
                      variant_flags=variant)

        if time:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(torch.cuda.current_stream(self.device))

        # at most three launches, as the reference issues them (tracker.py:1372-1436)
        if backtrack is not False:
            self._backtrack_pass(particles, ele_start, plan.head, plan.head_ends_turn, common)
            for _ in range(plan.full_turns):
                self._backtrack_pass(particles, 0, self.num_elements, plan.full_turns_end_turn,
                                     common)
            if plan.tail > 0:
                self._backtrack_pass(particles, 0, plan.tail, False, common)
        else:
            self._lattice.track(particles, num_turns=1, ele_start=ele_start,
                                num_ele_track=plan.head,
                                flag_end_turn_actions=plan.head_ends_turn, **common)
            if plan.full_turns > 0:
                self._track_middle(particles, plan.full_turns, plan.full_turns_end_turn, common)
            if plan.tail > 0:
                self._lattice.track(particles, num_turns=1, ele_start=0, num_ele_track=plan.tail,
                                    flag_end_turn_actions=False, **common)

        if time:
            ev1.record(torch.cuda.current_stream(self.device))
            ev1.synchronize()
            line.time_last_track = ev0.elapsed_time(ev1) * 1e-3
        else:
            line.time_last_track = None
        line.record_last_track = monitor
        self.record_last_track = monitor

    def _backtrack_pass(self, particles, ele_start, num_ele_track, end_turn_actions, common):
        """One pass of `track_line` under XS_FLAG_BACKTRACK (tracker.py:626-646, 702-731):
        the turn bookkeeping FIRST (increment_at_turn_backtrack,
        local_particle_custom_api.h:88-101: at_turn - 1, at_element = len(line), s = line
        length), then the elements ele_start .. ele_start + num_ele_track - 1 from the last
        to the first, each as its inverse, at_element counted DOWN.  The kernel runs the
        reversed inverse lattice forwards; it counts at_element up, mirrored here."""
        ff = particles._fields
        if end_turn_actions:
            act = ff['state'] > 0
            ff['at_turn'][act] -= 1
            ff['at_element'][act] = self.num_elements
            if self.line.reset_s_at_end_turn:
                ff['s'][act] = self.line_length
        if num_ele_track <= 0:
            return
        at0 = ff['at_element'].clone()
        self._back_lattice.track(
            particles, num_turns=1,
            ele_start=self.num_elements - (ele_start + num_ele_track),
            num_ele_track=num_ele_track, flag_end_turn_actions=False, **common)
        ff['at_element'].copy_(2 * at0 - ff['at_element'])

    def _track_middle(self, particles, num_turns, flag_end_turn_actions, common):
        """Full turns.  With `compact_every=N` the turns are issued in chunks of N
        and the surviving particles are compacted into dense warps between chunks
        (GPU stream compaction in place of the CPU contexts' reorganize,
        particles.py:1198-1259); the original slot order is restored at the end."""
        nn = self.compact_every
        if not nn or nn >= num_turns:
            self._lattice.track(particles, num_turns=num_turns, ele_start=0,
                                num_ele_track=self.num_elements,
                                flag_end_turn_actions=flag_end_turn_actions, **common)
            return
        perm_total = None
        done = 0
        while done < num_turns:
            chunk = min(nn, num_turns - done)
            self._lattice.track(particles, num_turns=chunk, ele_start=0,
                                num_ele_track=self.num_elements,
                                flag_end_turn_actions=flag_end_turn_actions, **common)
            done += chunk
            if done < num_turns:
                perm, n_active, _ = _cabi.compact(particles)
                perm_total = perm if perm_total is None else perm_total[perm]
                if n_active == 0:
                    break
        if perm_total is not None:
            inv = torch.empty_like(perm_total)
            inv[perm_total] = torch.arange(len(perm_total), device=perm_total.device)
            for nn_, tt in particles._fields.items():
                tt.copy_(tt[inv])
