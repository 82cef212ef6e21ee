"""Lattice lowering: xtrack elements -> the op stream ("program") that the CUDA
tracking kernel interprets (format: csrc/xtb_ops.h).

This replaces, on the host and once per `build_tracker`, everything the
reference re-derives per element call inside its kernel from element-constant
data: parent/slice weights, default model/integrator selection, automatic
kick counts, `configure_tracking_model`, coefficient scaling with factorials,
edge coefficients, rbend geometry and the misalignment frame algebra.  All of
it is evaluated with plain IEEE doubles in the reference's operation order
(Python floats and `math` = the same libm the CPU reference links), so the
parameters the GPU receives are bit-identical to what the reference computes.

Reference routines restated here (xtrack/beam_elements/elements_src/):
  track_magnet.h:289-648            track_magnet_particles   -> `_lower_magnet`
  track_magnet_configure.h:7-194    configure_tracking_model -> `_configure_tracking_model`
  track_magnet_kick.h:149-228       kick_is_inactive / Horner scaling -> `_horner_coeffs`
  track_magnet_edge.h:17-167        edge model selection     -> `_lower_edge`
  track_dipole_edge_linear.h:12-27  linear edge coefficients
  track_rf.h:317-453                track_rf_particles       -> `_lower_rf`
  track_misalignments.h:38-357      misalignment algebra     -> `_misalign_*`
  headers/track_local_particle_with_transformations.h:99-204 wrapper -> `_with_transformations`
  drift.h, multipole.h, quadrupole.h, sextupole.h, octupole.h, bend.h, rbend.h,
  cavity.h, rfmultipole.h, dipoleedge.h, srotation.h, xyshift.h, limit*.h (wrappers)
"""
import math

import numpy as np

# ---- opcodes / flags: mirror of csrc/xtb_ops.h ----------------------------
OPS_ABI_VERSION = 7          # XTB_OPS_ABI_VERSION: _cabi.load() checks the library against it
F_START, F_END, F_GLOBAL, F_DRIFT = 0x01, 0x02, 0x04, 0x08

# fast set (fused program only): one whole element per op
FOP_NOP = 0
FOP_MULT0 = 1
FOP_MULT1 = 2
FOP_MULTN = 3          # aux = order >= 2
FOP_MULTPN = 4         # aux = order >= 2, only the top normal coefficient non-zero
FOP_MULTH0 = 5
FOP_EDGE = 6
FOP_RECT = 7
FOP_ELLIPSE = 8
FOP_FDRIFT = 9
FOP_MULTP1 = 10        # plain normal quadrupole kick
FOP_MULTH0N = 11       # MULTH0 with cs_0 == 0
FOP_MULTH1N = 12       # order 1 with curvature + k1*h term, normal components only

OPBIT_DRIFT = 16
GENERIC_FIRST = 32

# generic set
OP_NOP = 32
OP_DRIFT = 33
OP_DRIFT_EXACT = 34
OP_MULT = 35
OP_MULT_H = 36
OP_CAVITY = 37
OP_RFMULT = 38
OP_EDGE_LIN = 39
OP_SROT = 40
OP_XYSHIFT = 41
OP_SSHIFT = 42
OP_YROT = 43
OP_XROT = 44
OP_LIMIT_RECT = 45
OP_LIMIT_ELLIPSE = 46
OP_LIMIT_POLYGON = 47
OP_MONITOR = 48
OP_LAST_TURNS = 49
OP_KILL = 50
OP_SET_STATE = 51
OP_ADD_S_ZETA = 52
OP_ADD_X = 53
OP_BEAM_MON = 54
OP_BEAM_PROFILE = 55
OP_CRAB = 56
OP_BEAM_STATS = 57
# heavy set
OP_MAGNET_BODY = 64
OP_MAGNET_EDGE = 65
OP_DIPEDGE_NL = 66
HEAVY_FIRST = 64

NOT_ADDRESSABLE = 0xffffffff
MASK64 = (1 << 64) - 1
BODY_NK_SHIFT = 15      # OP_MAGNET_BODY aux: num_kicks field (csrc/xtb_thick.cuh::body_par)
TILE_WORDS = 1024

ONE_OVER_FACT = [1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664,
                 0.008333333333333333, 0.001388888888888889, 0.0001984126984126984,
                 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
                 2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10,
                 1.1470745597729725e-11, 7.647163731819816e-13, 4.779477332387385e-14,
                 2.8114572543455206e-15, 1.5619206968586225e-16, 8.22063524662433e-18]


def one_over_factorial(n):
    """headers/factorial.h:5-37"""
    if n < 0:
        return 0.
    if n < 20:
        return ONE_OVER_FACT[n]
    return 1. / math.gamma(float(n + 1))


# Algorithmic flop table per op (SURVEY.md §8d convention: add/mul = 1, FMA = 2,
# div = 1, sqrt = 1, libm call = 1; work on all-zero coefficient sets is not
# credited).  Single source for bench.py's roofline.achieved.
def flops_drift():
    return 17


def flops_mult(order):
    return 8 * order + 4


FLOPS = {
    OP_NOP: 0, OP_DRIFT: 17, OP_DRIFT_EXACT: 21, OP_CAVITY: 25, OP_EDGE_LIN: 6,
    OP_SROT: 12, OP_XYSHIFT: 2, OP_SSHIFT: 23, OP_YROT: 30, OP_XROT: 30,
    OP_LIMIT_RECT: 0, OP_LIMIT_ELLIPSE: 5, OP_MONITOR: 0, OP_LAST_TURNS: 0,
    OP_BEAM_STATS: 80, OP_KILL: 0, OP_SET_STATE: 0, OP_ADD_S_ZETA: 2, OP_ADD_X: 1,
}


class Program:
    """Accumulates the ops of each element; `finish(fused)` returns the uint64 word
    array + element offsets of the FUSED or the PLAIN program (csrc/xtb_ops.h)."""

    def __init__(self):
        self.elements = []      # per element: (ops, static_thick); op = [opcode, aux, params]
        self._cur = []
        self.flops = 0.0        # algorithmic flop per particle-turn (sum over elements)
        self.n_transc = 0.0
        self.op_hist = {}
        self.has_heavy = False
        self.monitors = []      # in-line ParticlesMonitor objects (index = aux)
        self.last_turns_monitors = []
        self.beam_monitors = []     # BeamPosition / BeamSize monitors (records kept alive here)

    def op(self, opcode, params=(), aux=0, flops=None, transc=0):
        self._cur.append([int(opcode), int(aux), [float(p) if not isinstance(p, _RawWord)
                                                   else p for p in params]])
        if flops is None:
            flops = FLOPS.get(opcode, 0)
        self.flops += flops
        self.n_transc += transc
        self.op_hist[opcode] = self.op_hist.get(opcode, 0) + 1
        if opcode >= HEAVY_FIRST:
            self.has_heavy = True

    def _merge_linear_edges(self):
        """[EDGE_LIN] MAGNET_BODY [EDGE_LIN] of one element (a bend with linear edges) becomes
        ONE body op that applies the edge kicks itself (csrc/xtb_thick.cuh::magnet_body_n):
        one out-of-line op per bend instead of three.  Same arithmetic, same order."""
        ops = self._cur
        codes = [o[0] for o in ops]
        if OP_MAGNET_BODY not in codes or len(ops) > 3:
            return
        ib = codes.index(OP_MAGNET_BODY)
        before, after = ops[:ib], ops[ib + 1:]
        if (len(before) > 1 or len(after) > 1 or not (before or after)
                or any(o[0] != OP_EDGE_LIN for o in before + after)):
            return
        body = ops[ib]
        e_in = before[0][2][:2] if before else [0.0, 0.0]
        e_out = after[0][2][:2] if after else [0.0, 0.0]
        body[1] |= ((1 if before else 0) << 13) | ((1 if after else 0) << 14)
        body[2] = body[2] + list(e_in) + list(e_out)
        self._cur = [body]
        self.op_hist[OP_EDGE_LIN] -= len(before) + len(after)

    def end_element(self, static_thick):
        self._merge_linear_edges()
        if not self._cur:
            self._cur.append([OP_NOP, 0, []])
            self.op_hist[OP_NOP] = self.op_hist.get(OP_NOP, 0) + 1
        self.elements.append((self._cur, bool(static_thick)))
        self._cur = []

    # -- emission -------------------------------------------------------------
    @staticmethod
    def _emit(words, opcode, flags, aux, params, prefix_length=None):
        params = list(params)
        if len(params) & 1:
            params.append(0.0)
        nw = 2 + len(params)
        if nw > TILE_WORDS:
            raise ValueError(f'op of {nw} words > tile of {TILE_WORDS}')
        if prefix_length is not None:
            flags |= F_DRIFT
            if opcode < GENERIC_FIRST:
                opcode |= OPBIT_DRIFT       # fast ops: the prefixed form is its own opcode
        hdr = opcode | (flags << 8) | (nw << 16) | ((aux & 0xffffffff) << 32)
        words.append(np.uint64(hdr))
        words.append(np.float64(0.0 if prefix_length is None else prefix_length).view(np.uint64))
        for pp in params:
            if isinstance(pp, _RawWord):
                words.append(np.uint64(pp.value))
            else:
                words.append(np.float64(pp).view(np.uint64))

    @staticmethod
    def _fast_form(ops, static_thick):
        """(fast opcode, params) of a single-op element that has a fast equivalent."""
        if len(ops) != 1 or static_thick:
            return None
        opcode, aux, params = ops[0]
        if opcode == OP_NOP:
            return FOP_NOP, [], 0
        if opcode == OP_MULT and aux >= 1 and all(v == 0.0 for v in params[1:2 * (aux + 1)]):
            # plain normal magnet: every coefficient but the top normal one is a literal zero
            # (csrc/xtb_ops.h "Zero-coefficient specialisation")
            if aux == 1:
                return FOP_MULTP1, [params[0], 0.0], 0
            return FOP_MULTPN, [params[0], 0.0], aux
        if opcode == OP_MULT and aux <= 1:
            return FOP_MULT0 + aux, params, 0
        if opcode == OP_MULT and 2 * (aux + 1) + 2 <= 64:
            return FOP_MULTN, params, aux
        if opcode == OP_MULT_H and (aux & 0xff) == 0 and not ((aux >> 8) & 1):
            if params[5] == 0.0:
                return FOP_MULTH0N, [params[0], params[1], params[4], 0.0], 0
            return FOP_MULTH0, [params[0], params[1], params[4], params[5]], 0
        if (opcode == OP_MULT_H and (aux & 0xff) == 1 and ((aux >> 8) & 1)
                and params[5] == 0.0 and params[7] == 0.0):
            # [hl, B0, B1, 0, cn_1, cs_1, cn_0, cs_0] -> [hl, B0, B1, cn_1, cn_0, 0]
            return FOP_MULTH1N, [params[0], params[1], params[2], params[4], params[6], 0.0], 0
        if opcode == OP_EDGE_LIN:
            return FOP_EDGE, params[:2], 0
        if opcode == OP_LIMIT_RECT:
            min_x, max_x, min_y, max_y = params[:4]
            tx = min(-min_x, max_x) if (min_x < 0 < max_x) else 0.0
            ty = min(-min_y, max_y) if (min_y < 0 < max_y) else 0.0
            return FOP_RECT, params[:4], _inside_for_sure(tx, ty)
        if opcode == OP_LIMIT_ELLIPSE:
            a_squ, b_squ = params[:2]
            # |x| < 0.7 a and |y| < 0.7 b: x^2 / a^2 + y^2 / b^2 < 0.98
            return FOP_ELLIPSE, params[:3], _inside_for_sure(0.7 * math.sqrt(a_squ), 0.7 * math.sqrt(b_squ))
        return None

    def finish(self, fused=False):
        words = []
        offsets = []
        if not fused:
            for ops, static_thick in self.elements:
                offsets.append(len(words))
                for ii, (opcode, aux, params) in enumerate(ops):
                    flags = (F_START if ii == 0 else 0)
                    if ii == len(ops) - 1:
                        flags |= F_END | (F_GLOBAL if static_thick else 0)
                    self._emit(words, opcode, flags, aux, params)
                if len(words) - offsets[-1] > TILE_WORDS:
                    raise ValueError('element larger than a program tile')
        else:
            pending = None          # (length, element index) of a Drift waiting for its host op
            for ie, (ops, static_thick) in enumerate(self.elements):
                is_drift = (len(ops) == 1 and ops[0][0] == OP_DRIFT and static_thick)
                if is_drift and pending is None:
                    pending = (ops[0][2][0], ie)
                    offsets.append(None)            # filled when the host op is emitted
                    continue
                start = len(words)
                plen = None
                if pending is not None:
                    plen = pending[0]
                    offsets[pending[1]] = start
                    offsets.append(NOT_ADDRESSABLE)
                    pending = None
                else:
                    offsets.append(start)
                fast = ((FOP_FDRIFT, [ops[0][2][0]], 0) if is_drift
                        else self._fast_form(ops, static_thick))
                if fast is not None:
                    self._emit(words, fast[0], 0, fast[2], fast[1], prefix_length=plen)
                else:
                    for ii, (opcode, aux, params) in enumerate(ops):
                        flags = (F_START if ii == 0 else 0)
                        if ii == len(ops) - 1:
                            flags |= F_END | (F_GLOBAL if static_thick else 0)
                        self._emit(words, opcode, flags, aux, params,
                                   prefix_length=plen if ii == 0 else None)
                if len(words) - start > TILE_WORDS:
                    raise ValueError('element larger than a program tile')
            if pending is not None:
                offsets[pending[1]] = len(words)
                self._emit(words, FOP_FDRIFT, 0, 0, [pending[0]])
        offsets.append(len(words))
        return (np.array(words, dtype=np.uint64), np.array(offsets, dtype=np.uint32))


def _inside_for_sure(tx, ty):
    """aux word of the fast aperture ops: a particle with |x| < tx and |y| < ty is inside the
    aperture for sure.  The kernel compares the HIGH 32-bit words of |x|, |y| (integers: no
    FP64 instruction) with the high 16 bits of those of tx, ty rounded DOWN (packed x: bits
    31-16, y: bits 15-0); 0 = no such box, every particle takes the exact test."""
    def hi16(tt):
        if not (tt > 0.0) or not math.isfinite(tt):
            return 0
        return int(np.float64(tt).view(np.uint64) >> np.uint64(48)) & 0x7fff
    return (hi16(tx) << 16) | hi16(ty)


class _RawWord:
    """A parameter word that is an integer bit pattern, not a double."""

    def __init__(self, value):
        self.value = int(value) & 0xffffffffffffffff


# ---------------------------------------------------------------------------
# helpers restating element-constant arithmetic of the reference
# ---------------------------------------------------------------------------
def _horner_coeffs(knl, ksl, order, inv_factorial_order, factor, trim=True):
    """Scaled coefficients in Horner order (highest first), as pairs.
    track_magnet_kick.h:183-228: `chi * knl[index] * factor * inv_factorial`
    with `inv_factorial *= index` going down (chi is applied on the device).
    Leading all-zero orders are dropped (their contribution is an exact zero)."""
    if order < 0 or knl is None or ksl is None:
        return []
    inv_factorial = float(inv_factorial_order)
    index = int(order)
    cn = [0.0] * (order + 1)
    cs = [0.0] * (order + 1)
    cn[index] = float(knl[index]) * factor * inv_factorial
    cs[index] = float(ksl[index]) * factor * inv_factorial
    while index > 0:
        inv_factorial *= index
        index -= 1
        cn[index] = (float(knl[index]) * factor) * inv_factorial
        cs[index] = (float(ksl[index]) * factor) * inv_factorial
    top = order
    while trim and top > 0 and cn[top] == 0.0 and cs[top] == 0.0:
        top -= 1
    out = []
    for ii in range(top, -1, -1):
        out += [cn[ii], cs[ii]]
    return out


def _all_zero(c):
    return all(v == 0.0 for v in c)


def _kick_is_inactive(order, knl, ksl, k0, k1, k2, k3, k0s, k1s, k2s, k3s, h):
    """track_magnet_kick.h:149-180"""
    for v in (h, k0, k1, k2, k3, k0s, k1s, k2s, k3s):
        if v != 0:
            return False
    for index in range(order, -1, -1):
        if knl[index] != 0 or ksl[index] != 0:
            return False
    return True


def _configure_tracking_model(model, k0, k1, h, ks):
    """track_magnet_configure.h:7-194 -> dict"""
    if model == 1:
        model = 3
    h_is_zero = abs(h) < 1e-8
    if model == 2:
        drift_model = 5 if h_is_zero else 4
    elif model == 3:
        drift_model = 1 if h_is_zero else 7
    elif model == 4:
        drift_model = 3
    elif model == 5:
        drift_model = 1
    elif model == 6:
        drift_model = 0
    elif model == -1:
        drift_model = -1
    elif model == -2:
        drift_model = 6
    elif model == 7:
        drift_model = 1 if h_is_zero else 2
    elif model == 8:
        drift_model = 1 if h_is_zero else 8
    else:
        drift_model = 99999999
    c = dict(k0_drift=0.0, k1_drift=0.0, h_drift=0.0, ks_drift=0.0, k0_kick=0.0,
             k1_kick=0.0, h_kick=0.0, k0_h_correction=0.0, k1_h_correction=0.0,
             kick_rot_frame=0, drift_model=drift_model)
    if drift_model in (-1, 0, 1):
        c.update(k0_kick=k0, k1_kick=k1, h_kick=h, k0_h_correction=k0,
                 k1_h_correction=k1, kick_rot_frame=1)
    elif drift_model == 2:
        c.update(h_drift=h, k0_kick=k0, k1_kick=k1, h_kick=h, k0_h_correction=k0,
                 k1_h_correction=k1)
    elif drift_model == 3:
        c.update(k0_drift=k0, k1_drift=k1, h_drift=h, h_kick=h, k1_h_correction=k1)
    elif drift_model == 4:
        c.update(k0_drift=k0, h_drift=h, k1_kick=k1, h_kick=h, k1_h_correction=k1)
    elif drift_model == 5:
        c.update(k0_drift=k0, k1_kick=k1)
    elif drift_model == 6:
        c.update(ks_drift=ks, k0_kick=k0, k1_kick=k1, h_kick=h, k0_h_correction=k0,
                 k1_h_correction=k1, kick_rot_frame=1)
    elif drift_model in (7, 8):
        c.update(k0_drift=k0, h_drift=h, k1_kick=k1, h_kick=h, k0_h_correction=k0,
                 k1_h_correction=k1)
    return c


def _linear_edge_coefficients(k, e1, e1_fd, hgap, fint):
    """track_dipole_edge_linear.h:12-27"""
    corr = 2.0 * k * hgap * fint
    r21 = k * math.tan(e1)
    e1_v = e1 + e1_fd
    sin_e1_v = math.sin(e1_v)
    temp = corr / math.cos(e1_v) * (1.0 + sin_e1_v * sin_e1_v)
    r43 = -k * math.tan(e1_v - temp)
    return r21, r43


# -- misalignment (track_misalignments.h) -------------------------------------
def _emit_y_rotate(prog, theta):
    if theta != 0.0:
        prog.op(OP_YROT, [math.sin(theta), math.cos(theta), math.tan(theta)], transc=0)


def _emit_x_rotate(prog, phi):
    if phi != 0.0:
        prog.op(OP_XROT, [-math.sin(phi), math.cos(phi), -math.tan(phi)])


def _emit_s_rotate(prog, psi):
    if psi != 0.0:
        prog.op(OP_SROT, [math.sin(psi), math.cos(psi)])


def _emit_s_shift(prog, ds):
    if ds != 0.0:
        prog.op(OP_SSHIFT, [ds])


def _emit_misalign_steps(prog, steps, backtrack):
    """The elementary transformations of one misalignment function in order; when
    backtracking, the inverse ones in reverse order (track_misalignments.h:59-74, 97-112,
    225-240, 362-377)."""
    if backtrack:
        steps = [(kind, *[-v for v in args]) for kind, *args in reversed(steps)]
    for kind, *args in steps:
        if kind == 'xy':
            prog.op(OP_XYSHIFT, list(args))
        elif kind == 's':
            _emit_s_shift(prog, *args)
        elif kind == 'yrot':
            _emit_y_rotate(prog, *args)
        elif kind == 'xrot':
            _emit_x_rotate(prog, *args)
        else:
            _emit_s_rotate(prog, *args)


def _misalign_entry_straight(prog, dx, dy, ds, theta, phi, psi_no_frame, anchor, length,
                             psi_with_frame, backtrack=False):
    """track_misalignments.h:38-75"""
    mis_x = dx - anchor * math.cos(phi) * math.sin(theta)
    mis_y = dy - anchor * math.sin(phi)
    mis_s = ds - anchor * (math.cos(phi) * math.cos(theta) - 1)
    _emit_misalign_steps(prog, [('xy', mis_x, mis_y), ('s', mis_s), ('yrot', theta),
                                ('xrot', phi), ('srot', psi_no_frame),
                                ('srot', psi_with_frame)], backtrack)


def _misalign_exit_straight(prog, dx, dy, ds, theta, phi, psi_no_frame, anchor, length,
                            psi_with_frame, backtrack=False):
    """track_misalignments.h:78-113"""
    neg_part_length = anchor - length
    mis_x = neg_part_length * math.cos(phi) * math.sin(theta) - dx
    mis_y = neg_part_length * math.sin(phi) - dy
    mis_s = neg_part_length * (math.cos(phi) * math.cos(theta) - 1) - ds
    _emit_misalign_steps(prog, [('srot', -psi_with_frame), ('srot', -psi_no_frame),
                                ('xrot', -phi), ('yrot', -theta), ('s', mis_s),
                                ('xy', mis_x, mis_y)], backtrack)


def _mat_mul(a, b):
    """matrix_multiply_4x4, track_misalignments.h:381-389"""
    r = [[0.0] * 4 for _ in range(4)]
    for i in range(4):
        for j in range(4):
            r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] + a[i][3] * b[3][j]
    return r


def _rigid_inverse(m):
    """matrix_rigid_affine_inverse, track_misalignments.h:392-420"""
    inv = [[0.0] * 4 for _ in range(4)]
    for i in range(3):
        for j in range(3):
            inv[i][j] = m[j][i]
    inv[0][3] = -m[0][0] * m[0][3] - m[1][0] * m[1][3] - m[2][0] * m[2][3]
    inv[1][3] = -m[0][1] * m[0][3] - m[1][1] * m[1][3] - m[2][1] * m[2][3]
    inv[2][3] = -m[0][2] * m[0][3] - m[1][2] * m[1][3] - m[2][2] * m[2][3]
    inv[3] = [0.0, 0.0, 0.0, 1.0]
    return inv


def _misalignment_matrix(dx, dy, ds, theta, phi, psi):
    s_phi, c_phi = math.sin(phi), math.cos(phi)
    s_theta, c_theta = math.sin(theta), math.cos(theta)
    s_psi, c_psi = math.sin(psi), math.cos(psi)
    return [
        [-s_phi * s_psi * s_theta + c_psi * c_theta,
         -c_psi * s_phi * s_theta - c_theta * s_psi, c_phi * s_theta, dx],
        [c_phi * s_psi, c_phi * c_psi, s_phi, dy],
        [-c_theta * s_phi * s_psi - c_psi * s_theta,
         -c_psi * c_theta * s_phi + s_psi * s_theta, c_phi * c_theta, ds],
        [0.0, 0.0, 0.0, 1.0]]


def _misalign_entry_curved(prog, dx, dy, ds, theta, phi, psi_no_frame, anchor, length,
                           angle, h, psi_with_frame, backtrack=False):
    """track_misalignments.h:116-241"""
    if angle == 0.0 and (length != 0.0 or h == 0.0):
        return _misalign_entry_straight(prog, dx, dy, ds, theta, phi, psi_no_frame, anchor,
                                        length, psi_with_frame, backtrack)
    mm = _misalignment_matrix(dx, dy, ds, theta, phi, psi_no_frame)
    if length != 0.0:
        h = angle / length
    part_angle = anchor * h
    cpa, spa = math.cos(part_angle), math.sin(part_angle)
    cps, sps = math.cos(psi_with_frame), math.sin(psi_with_frame)
    first = [
        [(cpa - 1) * (cps * cps) + 1, (cpa - 1) * cps * sps, -cps * spa, (cpa - 1) * cps / h],
        [(cpa - 1) * cps * sps, (cpa - 1) * (sps * sps) + 1, -spa * sps, (cpa - 1) * sps / h],
        [cps * spa, spa * sps, cpa, spa / h],
        [0.0, 0.0, 0.0, 1.0]]
    inv_first = _rigid_inverse(first)
    me = _mat_mul(_mat_mul(first, mm), inv_first)
    mis_x, mis_y, mis_s = me[0][3], me[1][3], me[2][3]
    rot_theta = math.atan2(me[0][2], me[2][2])
    rot_phi = math.atan2(me[1][2], math.sqrt(me[1][0] * me[1][0] + me[1][1] * me[1][1]))
    rot_psi = math.atan2(me[1][0], me[1][1])
    _emit_misalign_steps(prog, [('xy', mis_x, mis_y), ('s', mis_s), ('yrot', rot_theta),
                                ('xrot', rot_phi), ('srot', rot_psi),
                                ('srot', psi_with_frame)], backtrack)


def _misalign_exit_curved(prog, dx, dy, ds, theta, phi, psi_no_frame, anchor, length,
                          angle, h, psi_with_frame, backtrack=False):
    """track_misalignments.h:244-378"""
    if angle == 0.0 and (length != 0.0 or h == 0.0):
        return _misalign_exit_straight(prog, dx, dy, ds, theta, phi, psi_no_frame, anchor,
                                       length, psi_with_frame, backtrack)
    mm = _misalignment_matrix(dx, dy, ds, theta, phi, psi_no_frame)
    inv_mm = _rigid_inverse(mm)
    if length != 0.0:
        h = angle / length
    part_angle = angle - h * anchor
    spa, cpa = math.sin(part_angle), math.cos(part_angle)
    st, ct = math.sin(psi_with_frame), math.cos(psi_with_frame)
    second = [
        [(cpa - 1) * (ct * ct) + 1, (cpa - 1) * ct * st, ct * -spa, (cpa - 1) * ct / h],
        [(cpa - 1) * ct * st, (cpa - 1) * (st * st) + 1, -spa * st, (cpa - 1) * st / h],
        [ct * spa, spa * st, cpa, spa / h],
        [0.0, 0.0, 0.0, 1.0]]
    inv_second = _rigid_inverse(second)
    re = _mat_mul(_mat_mul(inv_second, inv_mm), second)
    mis_x, mis_y, mis_s = re[0][3], re[1][3], re[2][3]
    rot_theta = math.atan2(re[0][2], re[2][2])
    rot_phi = math.atan2(re[1][2], math.sqrt(re[1][0] * re[1][0] + re[1][1] * re[1][1]))
    rot_psi = math.atan2(re[1][0], re[1][1])
    _emit_misalign_steps(prog, [('srot', -psi_with_frame), ('xy', mis_x, mis_y), ('s', mis_s),
                                ('yrot', rot_theta), ('xrot', rot_phi), ('srot', rot_psi)],
                         backtrack)


def _with_transformations(prog, el, body, *, length=0.0, curved=False, weight=1.0,
                          backtrack=False):
    """headers/track_local_particle_with_transformations.h:174-204 (+ :99-162); when
    backtracking the exit transformation comes first, each one inverted (:141-161)"""
    if not el.allow_rot_and_shift or not el.has_misalignment:
        body()
        return
    args = [el.shift_x, el.shift_y, el.shift_s, el.rot_y_rad, el.rot_x_rad,
            el.rot_s_rad_no_frame, el.rot_shift_anchor, length * weight]
    if curved:
        args = args + [el.angle * weight, el.h, el.rot_s_rad]
        entry, exit_ = _misalign_entry_curved, _misalign_exit_curved
    else:
        args = args + [el.rot_s_rad]
        entry, exit_ = _misalign_entry_straight, _misalign_exit_straight
    if backtrack:
        exit_(prog, *args, backtrack=True)
        body()
        entry(prog, *args, backtrack=True)
    else:
        entry(prog, *args)
        body()
        exit_(prog, *args)


# ---------------------------------------------------------------------------
# magnets (track_magnet.h:289-648)
# ---------------------------------------------------------------------------
def _emit_thin_multipole(prog, coeffs, hl, b0, b1):
    order = len(coeffs) // 2 - 1
    if hl == 0.0 and b0 == 0.0 and b1 == 0.0:
        if _all_zero(coeffs):
            prog.op(OP_NOP)
        else:
            prog.op(OP_MULT, coeffs, aux=order, flops=flops_mult(order))
    else:
        has_b1 = 1 if b1 != 0.0 else 0
        prog.op(OP_MULT_H, [hl, b0, b1, 0.0] + coeffs, aux=order | (has_b1 << 8),
                flops=flops_mult(order) + 14 + (8 if has_b1 else 0))


def _lower_edge(prog, synrad, *, model, is_exit, half_gap, knorm, kskew, knl, ksl,
                factor_knl_ksl, kl_order, knl_rel, ksl_rel, factor_knl_ksl_rel, order_rel,
                length, face_angle, face_angle_feed_down, fringe_integral,
                factor_for_backtrack=1.0):
    """track_magnet_edge.h:17-167 (no solenoid)"""
    k0 = 0.0
    k0 += knorm[0]
    if abs(length) > 1e-10 and kl_order > -1:
        k0 += factor_knl_ksl * knl[0] / length
    if abs(length) > 1e-10 and order_rel > -1 and knl_rel is not None:
        k0 += factor_knl_ksl_rel * knl_rel[0] / length
    if model == 0:
        r21, r43 = _linear_edge_coefficients(k0, face_angle, face_angle_feed_down, half_gap,
                                             fringe_integral)
        r21 = r21 * factor_for_backtrack
        r43 = r43 * factor_for_backtrack
        prog.op(OP_EDGE_LIN, [r21, r43])
    elif model in (1, 2):
        if factor_for_backtrack < 0:
            prog.op(OP_SET_STATE, aux=-32)       # the full edge model cannot be backtracked
        should_rotate = 0
        sin_, cos_, tan_ = 0.0, 1.0, 0.0
        if abs(face_angle) > 10e-10:
            should_rotate = 1
            sin_, cos_, tan_ = math.sin(face_angle), math.cos(face_angle), math.tan(face_angle)
        if is_exit:
            k0 = -k0
        nkl = kl_order + 1
        params = [k0, sin_, cos_, tan_, fringe_integral, half_gap, face_angle,
                  length / factor_knl_ksl, *[float(v) for v in knorm[:4]],
                  *[float(v) for v in kskew[:4]],
                  *[float(v) for v in knl[:nkl]], *[float(v) for v in ksl[:nkl]]]
        aux = (1 if is_exit else 0) | (model << 1) | (should_rotate << 3) | (nkl << 4)
        prog.op(OP_MAGNET_EDGE, params, aux=aux, flops=300, transc=6)
    # model 3 / -1: nothing besides the (zero) ax, ay reset


def _lower_magnet(prog, cfg, *, weight, length, order, inv_factorial_order, knl, ksl,
                  knl_rel, ksl_rel, rel_ref_strength, num_multipole_kicks, model,
                  default_model, integrator, default_integrator, radiation_flag,
                  delta_taper, h, hxl, k0, k1, k2, k3, k0s, k1s, k2s, k3s,
                  rbend_model=-1, rbend_compensate_sagitta=0, rbend_shift=0.0,
                  rbend_angle_diff=0.0, length_straight=0.0, body_active=1,
                  edge_entry_active=0, edge_exit_active=0, edge_entry_model=0,
                  edge_exit_model=0, edge_entry_angle=0.0, edge_exit_angle=0.0,
                  edge_entry_angle_fdown=0.0, edge_exit_angle_fdown=0.0,
                  edge_entry_fint=0.0, edge_exit_fint=0.0, edge_entry_hgap=0.0,
                  edge_exit_hgap=0.0, radiation_flag_parent=0):
    synrad = cfg['synrad']
    order_rel = len(knl_rel) - 1
    inv_factorial_order_rel = one_over_factorial(order_rel)
    factor_knl_ksl = 1.0
    theta_in = theta_out = 0.0
    cos_theta_in, sin_theta_in, cos_theta_out, sin_theta_out = 1.0, 0.0, 1.0, 0.0
    length_curved = 0.0
    x0_mid = x0_in = x0_out = 0.0

    if rbend_model == 0:
        rbend_model = 1
    angle = h * length
    if rbend_model == 1:
        edge_entry_angle += (angle - rbend_angle_diff) / 2.0
        edge_exit_angle += (angle + rbend_angle_diff) / 2.0
    elif rbend_model == 2:
        theta_in = (angle - rbend_angle_diff) / 2.0
        if abs(theta_in) > 1e-10:
            sin_theta_in, cos_theta_in = math.sin(theta_in), math.cos(theta_in)
        theta_out = (angle + rbend_angle_diff) / 2.0
        if abs(theta_out) > 1e-10:
            sin_theta_out, cos_theta_out = math.sin(theta_out), math.cos(theta_out)
        length_curved = length
        length = length_straight
        x0_mid -= rbend_shift
        if rbend_compensate_sagitta and abs(angle) > 1e-10:
            cos_rbha = math.cos(angle / 2.)
            x0_mid += 0.5 / h * (1 - cos_rbha)
        x0_in = x0_mid
        x0_out = x0_mid
        if abs(angle) > 1e-10:
            px0_in = math.sin(theta_in)
            px0_mid = px0_in - h * length_straight / 2
            sqrt_mid = math.sqrt(1 - px0_mid * px0_mid)
            x0_in -= 1 / h * (sqrt_mid - cos_theta_in)
            x0_out += 1 / h * (cos_theta_out - sqrt_mid)
        h = 0.0
        edge_entry_angle_fdown += theta_in
        edge_exit_angle_fdown += theta_out

    if cfg.get('backtrack'):          # track_magnet.h:413-433
        core_length = -length * weight
        core_length_curved = -length_curved * weight
        factor_knl_ksl_body = -factor_knl_ksl * weight
        factor_knl_ksl_edge = factor_knl_ksl
        factor_backtrack_edge = -1.
        hxl = -hxl
        edge_entry_active, edge_exit_active = edge_exit_active, edge_entry_active
        edge_entry_model, edge_exit_model = edge_exit_model, edge_entry_model
        edge_entry_angle, edge_exit_angle = edge_exit_angle, edge_entry_angle
        edge_entry_angle_fdown, edge_exit_angle_fdown = edge_exit_angle_fdown, edge_entry_angle_fdown
        edge_entry_fint, edge_exit_fint = edge_exit_fint, edge_entry_fint
        edge_entry_hgap, edge_exit_hgap = edge_exit_hgap, edge_entry_hgap
        theta_in, theta_out = -theta_out, -theta_in
        cos_theta_in, cos_theta_out = cos_theta_out, cos_theta_in
        sin_theta_in, sin_theta_out = -sin_theta_out, -sin_theta_in
        x0_in, x0_out = x0_out, x0_in
    else:
        core_length = length * weight
        core_length_curved = length_curved * weight
        factor_knl_ksl_body = factor_knl_ksl * weight
        factor_knl_ksl_edge = factor_knl_ksl
        factor_backtrack_edge = 1.

    if synrad:
        if radiation_flag == 10:
            radiation_flag = radiation_flag_parent
        if radiation_flag:
            factor_knl_ksl_body *= (1. + delta_taper)
            factor_knl_ksl_edge *= (1. + delta_taper)
            k0 *= (1 + delta_taper); k1 *= (1 + delta_taper)
            k2 *= (1 + delta_taper); k3 *= (1 + delta_taper)
            k0s *= (1 + delta_taper); k1s *= (1 + delta_taper)
            k2s *= (1 + delta_taper); k3s *= (1 + delta_taper)
    else:
        radiation_flag = 0

    knorm = [k0, k1, k2, k3]
    kskew = [k0s, k1s, k2s, k3s]
    edge_common = dict(knorm=knorm, kskew=kskew, knl=knl, ksl=ksl,
                       factor_knl_ksl=factor_knl_ksl_edge, kl_order=order,
                       knl_rel=knl_rel, ksl_rel=ksl_rel,
                       factor_knl_ksl_rel=factor_knl_ksl_edge * rel_ref_strength,
                       order_rel=order_rel, length=length,
                       factor_for_backtrack=factor_backtrack_edge)

    if edge_entry_active:
        if rbend_model == 2:
            prog.op(OP_YROT, [-sin_theta_in, cos_theta_in, -sin_theta_in / cos_theta_in])
            prog.op(OP_ADD_X, [x0_in])
        _lower_edge(prog, synrad, model=edge_entry_model, is_exit=0, half_gap=edge_entry_hgap,
                    face_angle=edge_entry_angle, face_angle_feed_down=edge_entry_angle_fdown,
                    fringe_integral=edge_entry_fint, **edge_common)

    if body_active:
        if integrator == 0:
            integrator = default_integrator
        if model == 0:
            model = default_model
        if model == -1:
            integrator = 3
            num_multipole_kicks = 1
        if weight != 1.0 and num_multipole_kicks > 0:
            num_multipole_kicks = int(math.ceil(num_multipole_kicks * weight))
        if num_multipole_kicks == 0:
            if not _kick_is_inactive(order, knl, ksl, k0, k1, k2, k3, k0s, k1s, k2s, k3s, h):
                if abs(h) < 1e-8:
                    num_multipole_kicks = 1
                else:
                    b_circum = 2 * 3.14159 / abs(h)
                    num_multipole_kicks = int(abs(core_length) / b_circum / 0.5e-3)
                    if num_multipole_kicks < 1:
                        num_multipole_kicks = 1
        c = _configure_tracking_model(model, k0, k1, h, 0.0)
        if c['drift_model'] in (6, 99999999):
            raise NotImplementedError('solenoid / invalid magnet model outside the contract')

        coeffs = _horner_coeffs(knl, ksl, order, inv_factorial_order, factor_knl_ksl_body)
        coeffs_rel = _horner_coeffs(knl_rel, ksl_rel, order_rel, inv_factorial_order_rel,
                                    factor_knl_ksl_body * rel_ref_strength)
        kmain_n = [c['k0_kick'], c['k1_kick'], k2, k3]
        kmain_s = [k0s, k1s, k2s, k3s]
        knl_main = [v * core_length for v in kmain_n]
        ksl_main = [v * core_length for v in kmain_s]
        coeffs_main = _horner_coeffs(knl_main, ksl_main, 3, 1. / (3 * 2), 1.0, trim=False)

        # curvature terms of track_magnet_kick.h:98-142 (element constants)
        htot = c['h_kick']
        if core_length != 0:
            htot += hxl / core_length
        k0l_mult = 0.0
        if order >= 0:
            k0l_mult = knl[0] * factor_knl_ksl_body
        if order_rel >= 0:
            k0l_mult += knl_rel[0] * factor_knl_ksl_body * rel_ref_strength
        a0 = c['k0_h_correction'] * core_length + k0l_mult
        k1l_mult = 0.0
        if order >= 1:
            k1l_mult = knl[1] * factor_knl_ksl_body
        if order_rel >= 1:
            k1l_mult += knl_rel[1] * factor_knl_ksl_body * rel_ref_strength
        a1 = c['k1_h_correction'] * core_length + k1l_mult

        drift_only = (num_multipole_kicks == 0 and c['k0_kick'] == 0 and c['k1_kick'] == 0
                      and c['h_kick'] == 0)
        thin_fast = (model == -1 and not (synrad and radiation_flag and core_length > 0)
                     and _all_zero(coeffs_rel) and _all_zero(coeffs_main) and not synrad)
        if thin_fast:
            # kick-only element, kick_weight = 1: the fast thin-multipole ops
            kw = 1.0
            if c['kick_rot_frame']:
                hl = c['h_kick'] * core_length * kw + hxl * kw
            else:
                hl = 0.0
            b0 = ((-1.0 * a0) * kw) * htot
            b1 = ((htot * 1.0) * a1) * kw
            if not coeffs:
                coeffs = [0.0, 0.0]
            _emit_thin_multipole(prog, coeffs, hl, b0, b1)
        elif drift_only and not (synrad and radiation_flag) and c['drift_model'] in (0, 1, -1):
            if c['drift_model'] == 0 and core_length != 0.0:
                prog.op(OP_DRIFT, [core_length])
            elif c['drift_model'] == 1 and core_length != 0.0:
                prog.op(OP_DRIFT_EXACT, [core_length])
        else:
            flags = ((integrator & 3) | ((c['drift_model'] + 1) << 2)
                     | ((1 if c['kick_rot_frame'] else 0) << 6)
                     | ((0 if _all_zero(coeffs) else 1) << 7)
                     | ((0 if _all_zero(coeffs_rel) else 1) << 8)
                     | ((0 if _all_zero(coeffs_main) else 1) << 9)
                     | ((radiation_flag & 3) << 10)
                     | ((1 if drift_only else 0) << 12))
            if num_multipole_kicks >= (1 << 17):
                raise ValueError('num_multipole_kicks too large')
            aux = flags | (num_multipole_kicks << BODY_NK_SHIFT)
            if not coeffs:
                coeffs = [0.0, 0.0]
            if not coeffs_rel:
                coeffs_rel = [0.0, 0.0]
            ou = len(coeffs) // 2 - 1
            orl = len(coeffs_rel) // 2 - 1
            trig, n_inner = _trig_table(c['drift_model'], integrator, num_multipole_kicks,
                                        drift_only, core_length, c['h_drift'])
            # (a kick-only body has no drift map: the slot of k0_drift carries RN(1 / length),
            # which the radiating thin kick divides by -- xtb_thick.cuh::thin_rad_kick_run)
            inv_length = (1.0 / core_length) if (c['drift_model'] == -1 and core_length != 0.0) else None
            params = [core_length, c['k0_drift'] if inv_length is None else inv_length,
                      c['k1_drift'], c['h_drift'], c['h_kick'], hxl,
                      a0, a1, htot, _RawWord(ou | (orl << 32)),
                      c['k0_drift'] + c['k0_kick'], c['k1_drift'] + c['k1_kick'], k2, k3,
                      k0s, k1s, k2s, k3s,
                      *coeffs_main, *coeffs, *coeffs_rel,
                      _RawWord((len(trig) // TRIG_STRIDE) | (n_inner << 32)),
                      (1 / c['h_drift']) if trig else 0.0, *trig]
            prog.op(OP_MAGNET_BODY, params, aux=aux,
                    flops=_body_flops(c['drift_model'], integrator, num_multipole_kicks,
                                      drift_only, ou, coeffs_main),
                    transc=_body_transc(c['drift_model'], integrator, num_multipole_kicks,
                                        drift_only))
        if rbend_model == 2 and model >= 0:
            ds = core_length_curved - core_length
            prog.op(OP_ADD_S_ZETA, [ds])

    if edge_exit_active:
        _lower_edge(prog, synrad, model=edge_exit_model, is_exit=1, half_gap=edge_exit_hgap,
                    face_angle=edge_exit_angle, face_angle_feed_down=edge_exit_angle_fdown,
                    fringe_integral=edge_exit_fint, **edge_common)
        if rbend_model == 2:
            prog.op(OP_ADD_X, [-x0_out])
            prog.op(OP_YROT, [-sin_theta_out, cos_theta_out, -sin_theta_out / cos_theta_out])


_YOSHIDA_D = (3.922568052387799819591407413100e-01, 5.100434119184584780271052295575e-01,
              -4.710533854097565531482416645304e-01, 6.875316825251809316199569366290e-02)


def _body_drift_lengths(integrator, n_kicks, drift_only, length):
    """The distinct lengths the integrator loops of csrc/xtb_thick.cuh::magnet_body
    (track_magnet.h:181-276) pass to the drift map, computed with the same operations."""
    if drift_only:
        return [length]
    if n_kicks <= 0:
        return []
    if integrator == 1:
        edge_w, inside_w = 0.5, 0.0
        if n_kicks > 1:
            edge_w = 1. / (2 * (1 + n_kicks))
            inside_w = float(n_kicks) / (float(n_kicks * n_kicks) - 1)
        out = [edge_w * length]
        if n_kicks > 1:
            out.append(inside_w * length)
        return out
    if integrator == 3:
        drift_weight = 1. / n_kicks
        return [0.5 * drift_weight * length]
    n_slices = n_kicks // 7 + (1 if n_kicks % 7 else 0)
    slice_length = length / n_slices
    return [slice_length * d for d in _YOSHIDA_D]


_Y4_NESTED_D = (0.6756035959798289, -0.17560359597982889)
TRIG_STRIDE = 6


def _trig_table(drift_model, integrator, n_kicks, drift_only, length, h):
    """cos(h*s), sin(h*s), sin(h*s/2) and RN(1/cos(h*s)) for every length s that the polar
    drift / curved exact bend of this body op will be called with (track_magnet_drift.h:45-87,
    272-345, 521-550): element constants that the reference evaluates per particle, evaluated
    here with the C library's libm (math.*), the one the reference's CPU build links.
    Returns (flat list of [s, cos, sin, sin_half, 1/cos, 0] entries, n_inner): entry
    `outer class * n_inner + inner class`, the outer classes being the distinct sub-step
    lengths of the integrator in `_body_drift_lengths` order and the inner ones the distinct
    fractions of the nested Yoshida bends (csrc/xtb_thick.cuh::magnet_drift_n indexes the
    same way)."""
    if drift_model not in (2, 4, 7, 8) or h == 0.0:
        return [], 1
    if drift_model in (2, 4):
        inner = (None,)
    elif drift_model == 7:
        inner = _Y4_NESTED_D
    else:
        inner = _YOSHIDA_D
    out = []
    for dl in _body_drift_lengths(integrator, n_kicks, drift_only, length):
        for d in inner:
            s = dl if d is None else d * dl
            ca = math.cos(h * s)
            out += [s, ca, math.sin(h * s), math.sin(0.5 * h * s),
                    (1.0 / ca) if ca != 0.0 else 0.0, 0.0]
    return out, len(inner)


_DRIFT_FLOPS = {-1: 0, 0: 17, 1: 21, 2: 45, 3: 95, 4: 60, 5: 40, 7: 4 * 45 + 9, 8: 8 * 45 + 21}
_DRIFT_TRANSC = {-1: 0, 0: 0, 1: 0, 2: 3, 3: 4, 4: 4, 5: 2, 7: 12, 8: 24}


def _body_steps(integrator, n_kicks, drift_only):
    """(#drift calls, #kick calls) of the integrator loops, track_magnet.h:181-276"""
    if drift_only:
        return 1, 0
    if integrator == 1:
        return n_kicks + 1, n_kicks
    if integrator == 3:
        return 2 * n_kicks, n_kicks
    n_slices = n_kicks // 7 + (1 if n_kicks % 7 else 0)
    return 8 * n_slices, 7 * n_slices


def _body_flops(drift_model, integrator, n_kicks, drift_only, order_user, coeffs_main):
    nd, nk = _body_steps(integrator, n_kicks, drift_only)
    kick = flops_mult(order_user) + (0 if _all_zero(coeffs_main) else flops_mult(3)) + 10
    return nd * _DRIFT_FLOPS.get(drift_model, 0) + nk * kick


def _body_transc(drift_model, integrator, n_kicks, drift_only):
    nd, _ = _body_steps(integrator, n_kicks, drift_only)
    return nd * _DRIFT_TRANSC.get(drift_model, 0)


# ---------------------------------------------------------------------------
# RF (track_rf.h:317-453)
# ---------------------------------------------------------------------------
def _lower_rf(prog, cfg, *, weight, length, voltage, frequency, harmonic, lag, phase,
              absolute_time, order, knl, ksl, pn, ps, phase_n, phase_s, num_kicks, model,
              default_model, integrator, default_integrator, lag_taper, phase_taper,
              transverse_voltage=0.0, transverse_lag=0.0, transverse_phase=0.0):
    if cfg['synrad']:
        lag += lag_taper
        phase += phase_taper
    if cfg.get('backtrack'):          # track_rf.h:364-373 (the RF elements have no edges here)
        body_length = -length
        factor_knl_ksl_body = -1.0
        voltage = -voltage
        transverse_voltage = -transverse_voltage
    else:
        body_length = length
        factor_knl_ksl_body = 1.0
    if integrator == 0:
        integrator = default_integrator
    if model == 0:
        model = default_model
    if model == -1:
        integrator = 3
        num_kicks = 1
    if weight != 1.0 and num_kicks > 0:
        num_kicks = int(math.ceil(num_kicks * weight))
    if num_kicks == 0:
        num_kicks = 1
    c = _configure_tracking_model(model, 0., 0., 0., 0.)
    drift_model = c['drift_model']
    if drift_model == 5:
        drift_model = 1        # straight exact bend with k0 = 0 is the exact drift
    if drift_model not in (-1, 0, 1):
        raise NotImplementedError(f'RF element model {model} outside the contract')
    ll = body_length * weight
    vv = voltage * weight
    tv = transverse_voltage * weight
    ff = factor_knl_ksl_body * weight

    def drift(dl):
        if drift_model == -1 or dl == 0.0:
            return
        prog.op(OP_DRIFT if drift_model == 0 else OP_DRIFT_EXACT, [dl])

    def kick(kw):
        if cfg.get('_crab'):
            # CrabCavity (track_rf.h:116-156): no longitudinal voltage (its energy kick is an
            # exact zero), the transverse kick needs the particle's p0c: evaluated on the device
            prog.op(OP_CRAB, [tv * kw, frequency, transverse_lag, transverse_phase],
                    aux=int(absolute_time), flops=40, transc=2)
            return
        if order >= 0:
            # track_rf.h:65-115.  `bal = factor_knl_ksl * knl[kk] / factorial` is element
            # constant: folded here with the reference's operations (factorial accumulated
            # as a double product).  Orders above the highest non-zero coefficient add
            # cos*(0*z) - cos*(0*z) = +-0 to the kicks: exact identities, dropped (a crab
            # cavity placeholder with order 5 and one or no strength set costs 24 sin/cos
            # per particle in the reference for nothing).  order_eff = -1: only the energy
            # bookkeeping of LocalParticle_add_to_energy remains.
            fk = ff * kw
            bal = []
            factorial = 1.0
            for kk in range(order + 1):
                if kk > 0:
                    factorial *= kk
                bal.append((fk * float(knl[kk]) / factorial, fk * float(ksl[kk]) / factorial))
            order_eff = order
            while order_eff >= 0 and bal[order_eff][0] == 0.0 and bal[order_eff][1] == 0.0:
                order_eff -= 1
            params = [vv * kw, frequency, lag, phase, 0.0]
            for kk in range(order_eff + 1):
                params += [bal[kk][0], bal[kk][1], float(pn[kk]), float(ps[kk]),
                           float(phase_n[kk]), float(phase_s[kk])]
            prog.op(OP_RFMULT, params, aux=order_eff, flops=25 + 30 * (order_eff + 1),
                    transc=(1 if vv * kw != 0.0 else 0) + 4 * (order_eff + 1))
        else:
            prog.op(OP_CAVITY, [vv * kw, frequency, harmonic, lag, phase, 0.0, 0.0],
                    aux=int(absolute_time), transc=1)

    if integrator == 1:
        kick_weight = 1. / num_kicks
        edge_drift_weight = 0.5
        inside_drift_weight = 0.0
        if num_kicks > 1:
            edge_drift_weight = 1. / (2 * (1 + num_kicks))
            inside_drift_weight = float(num_kicks) / (float(num_kicks * num_kicks) - 1)
        drift(edge_drift_weight * ll)
        for _ in range(num_kicks - 1):
            kick(kick_weight)
            drift(inside_drift_weight * ll)
        kick(kick_weight)
        drift(edge_drift_weight * ll)
    elif integrator == 3:
        kick_weight = 1. / num_kicks
        drift_weight = kick_weight
        for _ in range(num_kicks):
            drift(0.5 * drift_weight * ll)
            kick(kick_weight)
            drift(0.5 * drift_weight * ll)
    elif integrator == 2:
        num_slices = num_kicks // 7 + (1 if num_kicks % 7 != 0 else 0)
        slice_length = ll / num_slices
        kick_weight = 1. / num_slices
        d = [3.922568052387799819591407413100e-01, 5.100434119184584780271052295575e-01,
             -4.710533854097565531482416645304e-01, 6.875316825251809316199569366290e-02]
        k = [7.845136104775599639182814826199e-01, 2.355732133593569921359289764951e-01,
             -1.177679984178870098432412305556e+00, 1.315186320683906284756403692882e+00]
        seq = [0, 1, 2, 3, 3, 2, 1, 0]
        for _ in range(num_slices):
            for ii, dd in enumerate(seq):
                drift(slice_length * d[dd])
                if ii < 7:
                    kick(kick_weight * k[[0, 1, 2, 3, 2, 1, 0][ii]])


# ---------------------------------------------------------------------------
# per-class lowering
# ---------------------------------------------------------------------------
def _magnet_common_kwargs(el):
    return dict(order=el.order, inv_factorial_order=el.inv_factorial_order,
                knl=[float(v) for v in el.knl], ksl=[float(v) for v in el.ksl],
                knl_rel=[float(v) for v in el.knl_rel], ksl_rel=[float(v) for v in el.ksl_rel],
                num_multipole_kicks=el.num_multipole_kicks, integrator=el.integrator,
                radiation_flag=el.radiation_flag, delta_taper=el.delta_taper)


_MAGNET_CLASSES = ('Multipole', 'Quadrupole', 'Sextupole', 'Octupole', 'Bend', 'RBend')
_EDGE_KEYS = ('edge_entry_active', 'edge_exit_active', 'edge_entry_model', 'edge_exit_model',
              'edge_entry_angle', 'edge_exit_angle', 'edge_entry_angle_fdown',
              'edge_exit_angle_fdown', 'edge_entry_fint', 'edge_exit_fint', 'edge_entry_hgap',
              'edge_exit_hgap')


def _magnet_call(el):
    """Arguments of the `track_magnet_particles` call that the element's wrapper header makes
    (multipole.h:16-75, quadrupole.h:15-75, sextupole.h, octupole.h, bend.h:15-76,
    rbend.h:15-75), as keywords of `_lower_magnet` (weight excluded)."""
    name = type(el).__name__
    if name == 'Multipole':
        thick = el._isthick_field > 0
        return dict(length=el.length, **_magnet_common_kwargs(el),
                    rel_ref_strength=float(el.main_strength),
                    model=(el.model if thick else -1), default_model=6, default_integrator=3,
                    h=0., hxl=el.hxl, k0=0., k1=0., k2=0., k3=0., k0s=0., k1s=0., k2s=0., k3s=0.)
    if name in ('Quadrupole', 'Sextupole', 'Octupole'):
        kk = dict(k1=0., k2=0., k3=0., k1s=0., k2s=0., k3s=0.)
        kn, ks = el._main
        kk[kn], kk[ks] = getattr(el, kn), getattr(el, ks)
        main = getattr(el, ks) if el.main_is_skew else getattr(el, kn)
        return dict(length=el.length, **_magnet_common_kwargs(el),
                    rel_ref_strength=el.length * main, model=el.model,
                    default_model=4 if name == 'Quadrupole' else 6, default_integrator=3,
                    h=0., hxl=0., k0=0., k0s=0., **kk,
                    edge_entry_active=el.edge_entry_active, edge_exit_active=el.edge_exit_active,
                    edge_entry_model=1, edge_exit_model=1)
    assert name in ('Bend', 'RBend')
    extra = {}
    if name == 'RBend':
        extra = dict(rbend_model=el.rbend_model,
                     rbend_compensate_sagitta=el.rbend_compensate_sagitta,
                     rbend_shift=el.rbend_shift, rbend_angle_diff=el.rbend_angle_diff,
                     length_straight=el.length_straight)
    return dict(length=el.length, **_magnet_common_kwargs(el),
                rel_ref_strength=el.k0 * el.length, model=el.model, default_model=3,
                default_integrator=2, h=el.h, hxl=0., k0=el.k0, k1=el.k1, k2=el.k2, k3=0.,
                k0s=0., k1s=0., k2s=0., k3s=0.,
                edge_entry_active=el.edge_entry_active, edge_exit_active=el.edge_exit_active,
                edge_entry_model=el.edge_entry_model, edge_exit_model=el.edge_exit_model,
                edge_entry_angle=el.edge_entry_angle, edge_exit_angle=el.edge_exit_angle,
                edge_entry_angle_fdown=el.edge_entry_angle_fdown,
                edge_exit_angle_fdown=el.edge_exit_angle_fdown,
                edge_entry_fint=el.edge_entry_fint, edge_exit_fint=el.edge_exit_fint,
                edge_entry_hgap=el.edge_entry_hgap, edge_exit_hgap=el.edge_exit_hgap, **extra)


def _cavity_call(el):
    """Arguments of the `track_rf_particles` call of cavity.h:12-48 (weight excluded)."""
    return dict(length=el.length, voltage=el.voltage, frequency=el.frequency,
                harmonic=el.harmonic, lag=el.lag, phase=el.phase,
                absolute_time=el.absolute_time, order=-1, knl=None, ksl=None, pn=None, ps=None,
                phase_n=None, phase_s=None, num_kicks=el.num_kicks, model=el.model,
                default_model=6, integrator=el.integrator, default_integrator=3,
                lag_taper=el.lag_taper, phase_taper=el.phase_taper)


def _crab_call(el):
    """Arguments of the `track_rf_particles` call of crab_cavity.h (weight excluded)."""
    return dict(length=el.length, voltage=0., frequency=el.frequency, harmonic=0., lag=0., phase=0.,
                transverse_voltage=el.crab_voltage, transverse_lag=el.lag,
                transverse_phase=el.phase, absolute_time=el.absolute_time, order=-1, knl=None,
                ksl=None, pn=None, ps=None, phase_n=None, phase_s=None, num_kicks=el.num_kicks,
                model=el.model, default_model=6, integrator=el.integrator, default_integrator=3,
                lag_taper=el.lag_taper, phase_taper=el.phase_taper)


def _lower_slice(prog, el, cfg):
    """Slices of thick elements (slice_elements_{thin,thick,drift,edge}.py).  Their C wrappers
    are GENERATED from the parent's wrapper (elements_src/_generate_slice_elements_c_code.py):
    the parent's call with
      thick  weight = slice weight, radiation_flag / delta_taper of the slice, edges off;
      thin   as thick, and kick only: model -1, uniform integrator, one kick;
      entry  (exit)  body off, the parent's entry (exit) edge only, weight unused;
      drift  an expanded drift (exact under XTRACK_USE_EXACT_DRIFTS) of weight * length
             (drift_slice_*.h; the straight-body RBend: exact drift of the straight length and
             the path-length difference to the curved frame, drift_slice_rbend.h).
    The misalignment is the parent's, the anchor moved by the slice's offset
    (track_local_particle_with_transformations.h:99-162)."""
    par = el.parent
    pname = type(par).__name__
    kind = el._slice_kind
    back = bool(cfg.get('backtrack'))
    if kind == 'drift':
        ll = el.weight * par.length
        if back:                        # drift_slice_*.h: the lengths with the other sign
            ll = -ll
        if pname == 'DriftExact':
            prog.op(OP_DRIFT_EXACT, [ll])
        elif pname == 'RBend' and par.rbend_model == 2:
            ls = par.length_straight * el.weight
            ds_corr = (par.length - par.length_straight) * el.weight
            if back:
                ls *= -1
                ds_corr *= -1
            prog.op(OP_DRIFT_EXACT, [ls])
            prog.op(OP_ADD_S_ZETA, [ds_corr])
        else:
            model = 2 if cfg.get('exact_drifts') else 1
            if pname == 'Drift' and not cfg.get('exact_drifts'):
                model = par.model or 1
            prog.op(OP_DRIFT if model == 1 else OP_DRIFT_EXACT, [ll])
        return True

    if pname in ('Cavity', 'CrabCavity'):
        kw = _cavity_call(par) if pname == 'Cavity' else _crab_call(par)
        rf_cfg = cfg if pname == 'Cavity' else dict(cfg, _crab=True)
        if kind == 'thin':
            kw.update(num_kicks=1, model=-1, integrator=3)

        def body():
            _lower_rf(prog, rf_cfg, weight=el.weight, **kw)
    else:
        kw = _magnet_call(par)
        kw.update(radiation_flag=el.radiation_flag, radiation_flag_parent=par.radiation_flag,
                  delta_taper=el.delta_taper)
        weight = el.weight
        if kind in ('thin', 'thick'):
            for kk in _EDGE_KEYS:
                if kk in kw:
                    kw[kk] = 0
            if kind == 'thin':
                kw.update(num_multipole_kicks=1, model=-1, integrator=3)
        else:
            off = 'edge_exit' if kind == 'entry' else 'edge_entry'
            for kk in _EDGE_KEYS:
                if kk in kw and kk.startswith(off):
                    kw[kk] = 0
            kw.update(num_multipole_kicks=0, model=0, integrator=0, body_active=0)
            weight = 0.0

        def body():
            _lower_magnet(prog, cfg, weight=weight, **kw)

    if not (el.rot_and_shift_from_parent and par.has_misalignment):
        body()
        return bool(el.isthick)
    curved = pname in ('Bend', 'RBend')
    if kind == 'thin' and curved and (par.rot_x_rad != 0.0 or par.rot_y_rad != 0.0
                                       or par.rot_s_rad_no_frame != 0.0):
        # XT_INVALID_THIN_SLICE_TRANSFORM: the element is not tracked, the particle is lost
        prog.op(OP_SET_STATE, aux=-42)
        return bool(el.isthick)
    length = par.length if el.isthick else 0.0
    args = [par.shift_x, par.shift_y, par.shift_s, par.rot_y_rad, par.rot_x_rad,
            par.rot_s_rad_no_frame, par.rot_shift_anchor - el.slice_offset, length * el.weight]
    if curved:
        args = args + [par.angle * el.weight, par.h, par.rot_s_rad]
        entry, exit_ = _misalign_entry_curved, _misalign_exit_curved
    else:
        args = args + [par.rot_s_rad]
        entry, exit_ = _misalign_entry_straight, _misalign_exit_straight
    if back:
        exit_(prog, *args, backtrack=True)
        body()
        entry(prog, *args, backtrack=True)
    else:
        entry(prog, *args)
        body()
        exit_(prog, *args)
    return bool(el.isthick)


def lower_element(prog, el, cfg):
    """Appends the ops of one element; returns True if the class is statically
    thick (global aperture check after it, tracker.py:681-689)."""
    name = type(el).__name__
    back = bool(cfg.get('backtrack'))      # XS_FLAG_BACKTRACK: every class states its inverse
    sign = -1.0 if back else 1.0

    if name in ('Marker', '_Placeholder'):
        return False

    if name == 'Drift':                 # elements_src/drift.h:16-19
        model = 2 if cfg.get('exact_drifts') else (el.model or 1)
        length = -el.length if back else el.length
        if model == 1:
            prog.op(OP_DRIFT, [length])
        elif model == 2:
            prog.op(OP_DRIFT_EXACT, [length])
        return True

    if name == 'DriftExact':            # elements_src/drift_exact.h:16-19
        prog.op(OP_DRIFT_EXACT, [-el.length if back else el.length])
        return True

    if name in _MAGNET_CLASSES:
        kw = _magnet_call(el)
        thick_len = el.length if (name != 'Multipole' or el._isthick_field > 0) else 0.0
        _with_transformations(prog, el, lambda: _lower_magnet(prog, cfg, weight=1., **kw),
                              length=thick_len, curved=name in ('Bend', 'RBend'), backtrack=back)
        return name != 'Multipole'

    if name == 'Cavity':
        kw = _cavity_call(el)
        _with_transformations(prog, el, lambda: _lower_rf(prog, cfg, weight=1., **kw),
                              length=el.length, backtrack=back)
        return True

    if getattr(el, '_slice_kind', None) is not None:
        return _lower_slice(prog, el, cfg)

    if name == 'CrabCavity':
        kw = _crab_call(el)
        _with_transformations(prog, el, lambda: _lower_rf(prog, dict(cfg, _crab=True), weight=1., **kw),
                              length=el.length, backtrack=back)
        return True

    if name == 'RFMultipole':
        def body():
            _lower_rf(prog, cfg, weight=1., length=0., voltage=el.voltage,
                      frequency=el.frequency, harmonic=0., lag=el.lag, phase=el.phase,
                      absolute_time=0, order=el.order, knl=el.knl, ksl=el.ksl, pn=el.pn,
                      ps=el.ps, phase_n=el.phase_n, phase_s=el.phase_s, num_kicks=1, model=-1,
                      default_model=0, integrator=0, default_integrator=0, lag_taper=0.,
                      phase_taper=0.)
        _with_transformations(prog, el, body, length=0.0, backtrack=back)
        return False

    if name == 'DipoleEdge':            # elements_src/dipoleedge.h:27-75
        def body():
            delta_taper = el.delta_taper if cfg['synrad'] else 0.0
            if el.model == 0:
                r21 = el.r21 * (1 + delta_taper)
                r43 = el.r43 * (1 + delta_taper)
                if back:
                    r21, r43 = -r21, -r43
                prog.op(OP_EDGE_LIN, [r21, r43])
            elif el.model == 1 and back:
                prog.op(OP_SET_STATE, aux=-32)
            elif el.model == 1:
                if abs(el.e1) < 10e-10:
                    sct = [-999.0, -999.0, -999.0]
                else:
                    sct = [math.sin(el.e1), math.cos(el.e1), math.tan(el.e1)]
                prog.op(OP_DIPEDGE_NL, [el.k, el.e1, el.fint, el.hgap, *sct], aux=el.side,
                        flops=300, transc=6)
        _with_transformations(prog, el, body, backtrack=back)
        return False

    if name == 'SRotation':             # elements_src/srotation.h:14-27
        prog.op(OP_SROT, [-el.sin_z if back else el.sin_z, el.cos_z])
        return False

    if name == 'XYShift':               # elements_src/xyshift.h:13-28
        prog.op(OP_XYSHIFT, [-el.dx, -el.dy] if back else [el.dx, el.dy])
        return False

    if name == 'Translation':           # elements_src/translation.h:13-26
        prog.op(OP_XYSHIFT, [-el.shift_x, -el.shift_y] if back else [el.shift_x, el.shift_y])
        return False

    if name == 'Rotation':              # elements_src/rotation.h:13-60
        order = (el._first_rot, el._second_rot, el._third_rot)
        if back:                        # opposite angles, opposite order (:25-33)
            order = order[::-1]
        for axis in order:
            if axis == 0 and el.rot_x_rad != 0.0:
                aa = sign * el.rot_x_rad
                prog.op(OP_XROT, [math.sin(aa), math.cos(aa), math.tan(aa)])
            elif axis == 1 and el.rot_y_rad != 0.0:
                aa = sign * el.rot_y_rad
                prog.op(OP_YROT, [math.sin(aa), math.cos(aa), math.tan(aa)])
            elif axis == 2 and el.rot_s_rad != 0.0:
                aa = sign * el.rot_s_rad
                prog.op(OP_SROT, [math.sin(aa), math.cos(aa)])
        return False

    if name == 'LimitRect':
        _with_transformations(prog, el, lambda: prog.op(
            OP_LIMIT_RECT, [el.min_x, el.max_x, el.min_y, el.max_y]), backtrack=back)
        return False

    if name == 'LimitEllipse':
        _with_transformations(prog, el, lambda: prog.op(
            OP_LIMIT_ELLIPSE, [el.a_squ, el.b_squ, el.a_b_squ]), backtrack=back)
        return False

    if name == 'LimitPolygon':
        nv = len(el.x_vertices)
        _with_transformations(prog, el, lambda: prog.op(
            OP_LIMIT_POLYGON, [*el.x_vertices, *el.y_vertices], aux=nv, flops=7 * nv), backtrack=back)
        return False

    if name == 'ParticlesMonitor':
        prog.monitors.append(el)
        prog.op(OP_MONITOR, aux=len(prog.monitors) - 1)
        return False

    if name == 'LastTurnsMonitor':
        prog.last_turns_monitors.append(el)
        prog.op(OP_LAST_TURNS, aux=len(prog.last_turns_monitors) - 1)
        return False

    if name in ('BeamPositionMonitor', 'BeamSizeMonitor'):
        # the op carries the monitor's parameters and the device address of its record
        rec = el.allocate(cfg.get('device'))
        stop = el.particle_id_start + el.num_particles
        prog.op(OP_BEAM_MON, [_RawWord(el.start_at_turn & MASK64),
                              _RawWord(el.particle_id_start & MASK64), _RawWord(stop & MASK64),
                              el.frev, el.sampling_frequency, _RawWord(el.n_slots),
                              _RawWord(rec.data_ptr()), 0.0],
                aux=len(el.properties), flops=10)
        prog.beam_monitors.append(el)
        return False

    if name == 'BeamProfileMonitor':
        dd = el.allocate(cfg.get('device'))
        stop = el.particle_id_start + el.num_particles
        prog.op(OP_BEAM_PROFILE,
                [_RawWord(el.start_at_turn & MASK64), _RawWord(el.particle_id_start & MASK64),
                 _RawWord(stop & MASK64), el.frev, el.sampling_frequency, _RawWord(el.sample_size),
                 _RawWord(el.nx), el.x_min, el.dx, _RawWord(el.ny), el.y_min, el.dy,
                 _RawWord(dd['counts_x'].data_ptr()), _RawWord(dd['counts_y'].data_ptr())],
                flops=10)
        prog.beam_monitors.append(el)
        return False

    if name == 'BeamStatsMonitor':
        # the op carries the device address of the monitor's descriptor (monitors.py)
        desc = el.allocate(cfg.get('device'))
        prog.op(OP_BEAM_STATS, [_RawWord(desc.data_ptr()), 0.0])
        prog.beam_monitors.append(el)
        return False

    raise NotImplementedError(f'element class {name} is outside the hot-path contract')


def lower_line(elements, *, synrad=False, exact_drifts=False, device=None, backtrack=False):
    """Lowers a sequence of host elements.  Returns the `Program`.  `device`: where the
    records of in-line beam monitors live (their address goes into the program).
    `backtrack`: every element as its inverse map (what the reference's classes do under
    XS_FLAG_BACKTRACK); the caller passes the elements in REVERSE order (tracker.py:628-646)."""
    cfg = dict(synrad=bool(synrad), exact_drifts=bool(exact_drifts), device=device,
               backtrack=bool(backtrack))
    prog = Program()
    cache = {}
    for el in elements:
        static_thick = lower_element(prog, el, cfg)
        prog.end_element(static_thick)
    return prog
