"""ctypes binding of `libxtb200.so` (include/xtb200.h).

The only way from the Python host into the CUDA code.  There is no fallback:
if the library is missing it is built with nvcc; if that fails, or no CUDA
device is usable, the calls raise.
"""
import ctypes as ct
import os

import numpy as np
import torch

from . import build as _build

NUM_FIELDS = 32

VARIANT_EXACT = 1
VARIANT_SYNRAD = 2
VARIANT_FREEZE_LONG = 4
VARIANT_PLAIN_PROGRAM = 8
VARIANT_PHILOX = 16


class XtbParticles(ct.Structure):
    _fields_ = [('capacity', ct.c_int64), ('q0', ct.c_double), ('mass0', ct.c_double),
                ('t_sim', ct.c_double), ('field', ct.c_void_p * NUM_FIELDS)]


class XtbMonitor(ct.Structure):
    _fields_ = [(nn, ct.c_int64) for nn in (
        'start_at_turn', 'stop_at_turn', 'part_id_start', 'part_id_end', 'ebe_mode',
        'n_repetitions', 'repetition_period')] + [('field', ct.c_void_p * NUM_FIELDS)]


class XtbLastTurnsMonitor(ct.Structure):
    _fields_ = [(nn, ct.c_int64) for nn in (
        'particle_id_start', 'num_particles', 'n_last_turns', 'every_n_turns')] + [
        ('field', ct.c_void_p * 9)]


class XtbStats(ct.Structure):
    _fields_ = [('n_alive', ct.c_int64), ('n_lost', ct.c_int64), ('sum', ct.c_double * 6),
                ('sum2', ct.c_double * 21)]


class XtbError(RuntimeError):
    pass


_lib = None


def lib_path():
    return _build.LIB


def load():
    """Loads (building first if needed) the shared library; raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    # an existing library is used as it is (it was built by `__graft_entry__.build()` /
    # `python -m xtrack_b200.build`; file times do not survive the copy to a GPU box)
    if not os.path.exists(_build.LIB):
        try:
            _build.build()
        except Exception as err:
            raise XtbError(f'libxtb200.so is missing and could not be built: {err}')
    lib = ct.CDLL(_build.LIB)
    lib.xtb_last_error_string.restype = ct.c_char_p
    lib.xtb_version.restype = ct.c_char_p
    lib.xtb_launch_count.restype = ct.c_int64
    # a stale library must not interpret a newer op stream (the op set is mirrored by hand
    # between lowering.py and csrc/xtb_ops.h)
    from . import lowering
    try:
        lib.xtb_ops_abi_version.restype = ct.c_int
        got = int(lib.xtb_ops_abi_version())
    except AttributeError:
        got = None
    # (XTB_LIB_ABI_OVERRIDE: A/B sessions that load an OLDER build as a tuning variant beside the
    # product library, when the format change between the two is known not to matter to it)
    if got != lowering.OPS_ABI_VERSION and not (
            _build.SUFFIX and os.environ.get('XTB_LIB_ABI_OVERRIDE') == str(got)):
        raise XtbError(f'{_build.LIB} interprets op-stream format {got}, the host lowering '
                       f'emits {lowering.OPS_ABI_VERSION}: rebuild with `python -m xtrack_b200.build -f`')
    lib.xtb_lattice_create.argtypes = [ct.c_void_p, ct.c_size_t, ct.c_void_p,
                                       ct.c_void_p, ct.c_size_t, ct.c_void_p, ct.c_size_t,
                                       ct.c_double, ct.c_int, ct.POINTER(ct.c_void_p)]
    lib.xtb_lattice_destroy.argtypes = [ct.c_void_p]
    lib.xtb_lattice_set_inline_monitors.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_size_t,
                                                    ct.c_void_p, ct.c_size_t]
    lib.xtb_lattice_set_synrad_tables.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_size_t]
    lib.xtb_track.argtypes = [ct.c_void_p, ct.POINTER(XtbParticles), ct.c_int64, ct.c_int32,
                              ct.c_int32, ct.c_int32, ct.c_int32, ct.c_int32,
                              ct.POINTER(XtbMonitor), ct.c_uint64, ct.c_double, ct.c_uint32,
                              ct.c_void_p]
    lib.xtb_rng_init.argtypes = [ct.POINTER(XtbParticles), ct.c_void_p, ct.c_int64, ct.c_int,
                                 ct.c_void_p]
    lib.xtb_reduce_stats.argtypes = [ct.POINTER(XtbParticles), ct.c_void_p, ct.c_int,
                                     ct.c_void_p]
    lib.xtb_loss_histogram.argtypes = [ct.POINTER(XtbParticles), ct.c_void_p, ct.c_int64,
                                       ct.c_int, ct.c_void_p]
    lib.xtb_compact_scratch_bytes.argtypes = [ct.c_int64]
    lib.xtb_compact_scratch_bytes.restype = ct.c_size_t
    lib.xtb_compact.argtypes = [ct.POINTER(XtbParticles), ct.c_void_p, ct.c_void_p, ct.c_void_p,
                                ct.c_int, ct.c_void_p]
    lib.xtb_measure_dfma_peak.argtypes = [ct.c_int, ct.c_double, ct.POINTER(ct.c_double)]
    lib.xtb_selftest_math.argtypes = [ct.c_int, ct.c_int64, ct.c_uint64, ct.c_int,
                                      ct.POINTER(ct.c_uint64)]
    lib.xtb_eval_philox.argtypes = [ct.c_int, ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_uint32,
                                    ct.c_int64, ct.c_void_p]
    lib.xtb_eval_libm.argtypes = [ct.c_int, ct.c_void_p, ct.c_int64, ct.c_void_p]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise XtbError(f'libxtb200 error {rc}: {load().xtb_last_error_string().decode()}')


def _require_cuda(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise XtbError('xtrack_b200 tracks on CUDA devices only (no CPU fallback); '
                       f'particles are on {device}')
    if not torch.cuda.is_available():
        raise XtbError('no CUDA device available')
    return device.index if device.index is not None else torch.cuda.current_device()


def particles_struct(p):
    st = XtbParticles()
    st.capacity = p._capacity
    st.q0, st.mass0, st.t_sim = p.q0, p.mass0, p.t_sim
    for ii, ptr in enumerate(p.field_pointers()):
        st.field[ii] = ptr
    return st


def monitor_struct(mon):
    st = XtbMonitor()
    for nn in ('start_at_turn', 'stop_at_turn', 'part_id_start', 'part_id_end', 'ebe_mode',
               'n_repetitions', 'repetition_period'):
        setattr(st, nn, getattr(mon, nn))
    for ii, ptr in enumerate(mon.field_pointers()):
        st.field[ii] = ptr
    return st


def last_turns_struct(mon):
    st = XtbLastTurnsMonitor()
    for nn in ('particle_id_start', 'num_particles', 'n_last_turns', 'every_n_turns'):
        setattr(st, nn, getattr(mon, nn))
    for ii, ptr in enumerate(mon.field_pointers()):
        st.field[ii] = ptr
    return st


class Lattice:
    """Handle of a lowered lattice resident on one GPU."""

    def __init__(self, fused, plain, line_length, device):
        """`fused`, `plain`: (words, elem_offset) of the two programs of the line
        (lowering.Program.finish); `fused` may be (None, None)."""
        self.device_index = _require_cuda(device)
        lib = load()
        pw = np.ascontiguousarray(plain[0], dtype=np.uint64)
        po = np.ascontiguousarray(plain[1], dtype=np.uint32)
        self.n_elements = len(po) - 1
        if fused is not None and fused[0] is not None:
            fw = np.ascontiguousarray(fused[0], dtype=np.uint64)
            fo = np.ascontiguousarray(fused[1], dtype=np.uint32)
            fargs = (fw.ctypes.data, len(fw), fo.ctypes.data)
        else:
            fargs = (None, 0, None)
        hh = ct.c_void_p()
        _check(lib.xtb_lattice_create(*fargs, pw.ctypes.data, len(pw), po.ctypes.data,
                                      self.n_elements, float(line_length), self.device_index,
                                      ct.byref(hh)))
        self.handle = hh

    def set_inline_monitors(self, monitors, last_turns):
        lib = load()
        mm = (XtbMonitor * max(1, len(monitors)))(*[monitor_struct(m) for m in monitors])
        ll = (XtbLastTurnsMonitor * max(1, len(last_turns)))(
            *[last_turns_struct(m) for m in last_turns])
        _check(lib.xtb_lattice_set_inline_monitors(self.handle, mm, len(monitors), ll,
                                                   len(last_turns)))

    def set_synrad_tables(self, blob):
        """Inverse-CDF tables of the quantum-kick radiation model (synrad_tables.make_blob)."""
        blob = np.ascontiguousarray(blob, dtype=np.float64)
        _check(load().xtb_lattice_set_synrad_tables(self.handle, blob.ctypes.data, len(blob)))

    def track(self, particles, *, num_turns, ele_start, num_ele_track, flag_end_turn_actions,
              flag_reset_s_at_end_turn, flag_monitor=0, monitor=None, track_flags=0,
              global_xy_limit=1.0, variant_flags=0, stream=None):
        lib = load()
        pst = particles_struct(particles)
        mst = ct.byref(monitor_struct(monitor)) if monitor is not None else None
        if stream is None:
            stream = torch.cuda.current_stream(self.device_index).cuda_stream
        _check(lib.xtb_track(self.handle, ct.byref(pst), int(num_turns), int(ele_start),
                             int(num_ele_track), int(bool(flag_end_turn_actions)),
                             int(bool(flag_reset_s_at_end_turn)), int(flag_monitor), mst,
                             int(track_flags), float(global_xy_limit), int(variant_flags),
                             ct.c_void_p(stream)))

    def close(self):
        if getattr(self, 'handle', None):
            load().xtb_lattice_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_init(particles, seeds):
    dev = _require_cuda(particles.device)
    seeds_dev = torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.uint32).view(np.int32)).to(
        particles.device)
    pst = particles_struct(particles)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _check(load().xtb_rng_init(ct.byref(pst), seeds_dev.data_ptr(), len(seeds), dev,
                               ct.c_void_p(stream)))
    torch.cuda.current_stream(dev).synchronize()     # seeds_dev must outlive the kernel


def reduce_stats(particles):
    """Per-GPU partial sums as a float64 tensor [n_alive, n_lost, sum[6], sum2[21]] on the
    particles' device (ready for `torch.distributed.all_reduce`)."""
    dev = _require_cuda(particles.device)
    raw = torch.zeros(ct.sizeof(XtbStats) // 8, dtype=torch.int64, device=particles.device)
    pst = particles_struct(particles)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _check(load().xtb_reduce_stats(ct.byref(pst), raw.data_ptr(), dev, ct.c_void_p(stream)))
    out = torch.empty(29, dtype=torch.float64, device=particles.device)
    out[:2] = raw[:2].to(torch.float64)
    out[2:] = raw[2:].view(torch.float64)
    return out


def loss_histogram(particles, n_elements):
    dev = _require_cuda(particles.device)
    hist = torch.zeros(n_elements + 1, dtype=torch.int64, device=particles.device)
    pst = particles_struct(particles)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _check(load().xtb_loss_histogram(ct.byref(pst), hist.data_ptr(), int(n_elements), dev,
                                     ct.c_void_p(stream)))
    return hist


def compact(particles):
    """Stable partition of the slots into [active | lost | unallocated] on the GPU.
    Returns (perm, n_active, n_lost); perm[dst] = src slot."""
    dev = _require_cuda(particles.device)
    n = particles._capacity
    perm = torch.empty(n, dtype=torch.int64, device=particles.device)
    counts = torch.zeros(2, dtype=torch.int64, device=particles.device)
    scratch = torch.empty(load().xtb_compact_scratch_bytes(n), dtype=torch.uint8,
                          device=particles.device)
    pst = particles_struct(particles)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _check(load().xtb_compact(ct.byref(pst), perm.data_ptr(), counts.data_ptr(),
                              scratch.data_ptr(), dev, ct.c_void_p(stream)))
    cc = counts.cpu()
    return perm, int(cc[0]), int(cc[1])


def measure_dfma_peak(device=0, seconds=1.0):
    """(sustained, burst) FP64 FMA FLOP/s of `device`."""
    out = (ct.c_double * 2)()
    _check(load().xtb_measure_dfma_peak(int(device), float(seconds), out))
    return float(out[0]), float(out[1])


def selftest_math(device=0, n_samples=1 << 28, seed=1, exponent_range=30):
    """Mismatches (rcp, sqrt, div) of csrc/xtb_math.cuh against the IEEE operators."""
    out = (ct.c_uint64 * 3)()
    _check(load().xtb_selftest_math(int(device), int(n_samples), int(seed), int(exponent_range), out))
    return tuple(int(v) for v in out)


def eval_philox(k0, k1, c0, c1, n, device=0):
    """n blocks of Philox4x32-10 from the device: counter (c0 + i, c1, 0, 0), key (k0, k1)."""
    out = np.empty((n, 4), dtype=np.uint32)
    _check(load().xtb_eval_philox(int(device), int(k0), int(k1), int(c0), int(c1), int(n),
                                  out.ctypes.data))
    return out


def eval_libm(x, device=0):
    """dict of sin, cos, exp, expm1, sinh, cosh of the host array `x` by the device functions of
    csrc/xtb_libm.cuh."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty((6, len(x)), dtype=np.float64)
    _check(load().xtb_eval_libm(int(device), x.ctypes.data, len(x), out.ctypes.data))
    return dict(zip(('sin', 'cos', 'exp', 'expm1', 'sinh', 'cosh'), out))


def launch_count():
    return int(load().xtb_launch_count())
