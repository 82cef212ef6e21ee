"""TEST INFRASTRUCTURE ONLY: host (g++) build of the device physics headers, used by
the CPU test tier to check the lattice lowering and op semantics against the
reference oracle where no GPU exists.  Nothing under xtrack_b200/ imports this."""
import ctypes as ct
import os
import subprocess

import numpy as np
import torch

import xtrack_b200 as xb
from xtrack_b200 import _cabi
from xtrack_b200.tracker import Tracker

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
LIB = os.path.join(HERE, '_build', 'libxtb_hostsim.so')
_lib = None
# explicit fma() calls of the device code (xtb_thick.cuh::div_by) as one instruction where the
# host has it (no contraction of a*b+c happens either way: -ffp-contract=off)
try:
    _FMA_FLAG = ['-mfma'] if ' fma ' in open('/proc/cpuinfo').read() else []
except OSError:
    _FMA_FLAG = []


def _sources():
    csrc = os.path.join(ROOT, 'xtrack_b200', 'csrc')
    return [os.path.join(HERE, 'xtb_hostsim.cpp')] + [
        os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(('.cuh', '.h'))]


def load():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        if (not os.path.exists(LIB)
                or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in _sources())):
            subprocess.run(['g++', '-O2', '-ffp-contract=off', *_FMA_FLAG, '-std=c++17', '-fPIC', '-shared',
                            os.path.join(HERE, 'xtb_hostsim.cpp'), '-o', LIB, '-lm'], check=True)
        _lib = ct.CDLL(LIB)
        _lib.xtb_hostsim_track.restype = ct.c_int
        _lib.xtb_hostsim_track.argtypes = [
            ct.c_void_p, ct.c_void_p, ct.POINTER(_cabi.XtbParticles), ct.c_int64, ct.c_int32,
            ct.c_int32, ct.c_int32, ct.c_int32, ct.c_int32, ct.POINTER(_cabi.XtbMonitor),
            ct.c_uint64, ct.c_double, ct.c_uint32, ct.c_double, ct.c_void_p, ct.c_void_p,
            ct.c_int32, ct.c_int32]
    return _lib


class HostSimLattice:
    """Mirrors the program selection of `xtb_track` (csrc/xtb_api.cu): the FUSED program
    when the element range falls on its op boundaries, else the PLAIN one."""
    npt = 3          # particle slots carried together, as in the thin CUDA kernels
    npt_heavy = 2    # ... and in the thick ones

    def __init__(self, fused, plain, line_length):
        self.plain = (np.ascontiguousarray(plain[0], dtype=np.uint64),
                      np.ascontiguousarray(plain[1], dtype=np.uint32))
        self.fused = None
        if fused is not None and fused[0] is not None:
            self.fused = (np.ascontiguousarray(fused[0], dtype=np.uint64),
                          np.ascontiguousarray(fused[1], dtype=np.uint32))
        self.line_length = float(line_length)
        self.programs_used = []
        self._mons = None
        self._ltms = None

    def set_inline_monitors(self, monitors, last_turns):
        self._mons = (_cabi.XtbMonitor * max(1, len(monitors)))(
            *[_cabi.monitor_struct(m) for m in monitors])
        self._ltms = (_cabi.XtbLastTurnsMonitor * max(1, len(last_turns)))(
            *[_cabi.last_turns_struct(m) for m in last_turns])

    def set_synrad_tables(self, blob):
        self._synrad_tables = np.ascontiguousarray(blob, dtype=np.float64)

    def track(self, particles, *, num_turns, ele_start, num_ele_track, flag_end_turn_actions,
              flag_reset_s_at_end_turn, flag_monitor=0, monitor=None, track_flags=0,
              global_xy_limit=1.0, variant_flags=0, stream=None):
        tt = getattr(self, '_synrad_tables', None)
        load().xtb_hostsim_set_synrad_tables(ct.c_void_p(tt.ctypes.data if tt is not None else None))
        pst = _cabi.particles_struct(particles)
        mst = ct.byref(_cabi.monitor_struct(monitor)) if monitor is not None else None
        na = 0xffffffff
        prog = self.plain
        if (self.fused is not None and flag_monitor != 2
                and not (variant_flags & _cabi.VARIANT_PLAIN_PROGRAM)
                and self.fused[1][ele_start] != na
                and self.fused[1][ele_start + num_ele_track] != na):
            prog = self.fused
        self.programs_used.append('fused' if prog is self.fused else 'plain')
        rc = load().xtb_hostsim_track(
            prog[0].ctypes.data, prog[1].ctypes.data, ct.byref(pst), int(num_turns),
            int(ele_start), int(num_ele_track), int(bool(flag_end_turn_actions)),
            int(bool(flag_reset_s_at_end_turn)), int(flag_monitor), mst, int(track_flags),
            float(global_xy_limit), int(variant_flags), self.line_length,
            ct.cast(self._mons, ct.c_void_p) if self._mons is not None else None,
            ct.cast(self._ltms, ct.c_void_p) if self._ltms is not None else None, self.npt | (self.npt_heavy << 8), 0)
        assert rc == 0

    def close(self):
        pass


class HostSimTracker(Tracker):
    """`Tracker` whose launches go to the host build of the device code."""

    def __init__(self, line, device=None, **kwargs):
        super().__init__(line, device='cpu', **kwargs)

    def _make_lattice(self, fused, plain):
        return HostSimLattice(fused, plain, self.line_length)


def build_hostsim_tracker(line):
    line.tracker = HostSimTracker(line)
    return line.tracker


def trig_stats(reset=True):
    """(lookups, misses) of the host-tabulated element trigonometry since the last reset."""
    lib = load()
    a, b = ct.c_longlong(0), ct.c_longlong(0)
    lib.xtb_hostsim_trig_stats(ct.byref(a), ct.byref(b), int(reset))
    return a.value, b.value


STOP_NAMES = ('end', 'slow', 'global_prefix', 'global_main', 'rect', 'ellipse')


def stop_counts(reset=True):
    """How often the hot loop (xtb_run_fast) came back, by reason, since the last reset."""
    lib = load()
    out = (ct.c_ulonglong * 8)()
    lib.xtb_hostsim_stop_counts(out, int(reset))
    return dict(zip(STOP_NAMES, list(out)))
