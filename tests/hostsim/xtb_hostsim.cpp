// TEST INFRASTRUCTURE ONLY -- never loaded by the product (xtrack_b200/).
//
// Host build (g++) of the device physics headers xtrack_b200/csrc/xtb_thin.cuh and
// xtb_thick.cuh plus a serial re-statement of the interpreter loop of
// xtb_kernel.cuh, so that the LOWERING (xtrack_b200/lowering.py) and the op
// semantics can be checked against the reference oracle in the CPU test tier
// (`pytest -m "not gpu"`), where no GPU exists.  Built with -ffp-contract=off:
// it rounds like the EXACT kernel variant.  The GPU tests check the real kernel.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __global__
static inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
using std::max;
using std::min;

#include "../../xtrack_b200/csrc/xtb_state.cuh"
#include "../../xtrack_b200/csrc/xtb_thin.cuh"
#include "../../xtrack_b200/csrc/xtb_thick.cuh"

struct HostMon { const xtb_monitor_t* m; };

static void monitor_store(const xtb_monitor_t& m, const int64_t at, const PState& P, const PSlot& G) {
    for (int f = 0; f < XTB_N_F64; ++f) ((double*) m.field[f])[at] = G.ld(f);
    ((double*) m.field[F_S])[at] = P.s;  ((double*) m.field[F_ZETA])[at] = P.zeta;
    ((double*) m.field[F_X])[at] = P.x;  ((double*) m.field[F_Y])[at] = P.y;
    ((double*) m.field[F_PX])[at] = P.px;  ((double*) m.field[F_PY])[at] = P.py;
    ((double*) m.field[F_DELTA])[at] = P.delta;  ((double*) m.field[F_RPP])[at] = P.rpp;
    ((double*) m.field[F_RVV])[at] = P.rvv;  ((double*) m.field[F_CHI])[at] = P.chi;
    for (int f = XTB_N_F64; f < XTB_N_F64 + XTB_N_I64; ++f) ((int64_t*) m.field[f])[at] = G.ldi(f);
    ((int64_t*) m.field[F_AT_ELEMENT])[at] = P.at_element;
    ((int64_t*) m.field[F_AT_TURN])[at] = P.at_turn;
    ((int64_t*) m.field[F_STATE])[at] = P.state;
    for (int f = F_RNG_S1; f <= F_RNG_S4; ++f) ((uint32_t*) m.field[f])[at] = G.ldu(f);
}

static void monitor_record(const xtb_monitor_t& m, const PState& P, const PSlot& G) {
    const int64_t n_turns_record = m.stop_at_turn - m.start_at_turn;
    const int64_t at_turn = m.ebe_mode ? (int64_t) P.at_element : P.at_turn;
    const int64_t pid = G.ldi(F_PARTICLE_ID);
    if (m.n_repetitions == 1) {
        if (at_turn >= m.start_at_turn && at_turn < m.stop_at_turn && pid < m.part_id_end
            && pid >= m.part_id_start)
            monitor_store(m, n_turns_record * (pid - m.part_id_start) + at_turn - m.start_at_turn, P, G);
    } else if (m.n_repetitions > 1) {
        if (at_turn < m.start_at_turn) return;
        const int64_t i_frame = (at_turn - m.start_at_turn) / m.repetition_period;
        if (i_frame < m.n_repetitions && at_turn >= m.start_at_turn + i_frame * m.repetition_period
            && at_turn < m.stop_at_turn + i_frame * m.repetition_period && pid < m.part_id_end
            && pid >= m.part_id_start)
            monitor_store(m, n_turns_record * (m.part_id_end - m.part_id_start) * i_frame
                                 + n_turns_record * (pid - m.part_id_start)
                                 + (at_turn - i_frame * m.repetition_period) - m.start_at_turn, P, G);
    }
}

static void last_turns_record(const xtb_last_turns_monitor_t& m, const PState& P, const PSlot& G) {
    const int64_t pid = G.ldi(F_PARTICLE_ID), at_turn = P.at_turn;
    if (at_turn >= 0 && at_turn % m.every_n_turns == 0 && m.particle_id_start <= pid
        && pid < m.particle_id_start + m.num_particles) {
        const int64_t offset = (at_turn / m.every_n_turns) % m.n_last_turns;
        const int64_t ip = pid - m.particle_id_start, slot = m.n_last_turns * ip + offset;
        ((uint32_t*) m.field[0])[ip] = (uint32_t) offset;
        ((uint32_t*) m.field[1])[slot] = (uint32_t) pid;
        ((uint32_t*) m.field[2])[slot] = (uint32_t) at_turn;
        ((float*) m.field[3])[slot] = (float) P.x;  ((float*) m.field[4])[slot] = (float) P.px;
        ((float*) m.field[5])[slot] = (float) P.y;  ((float*) m.field[6])[slot] = (float) P.py;
        ((float*) m.field[7])[slot] = (float) P.delta;  ((float*) m.field[8])[slot] = (float) P.zeta;
    }
}

template <bool SYNRAD, bool FRZ>
static void run(const XtbTrackArgs& a) {
    for (int64_t slot = 0; slot < a.part.capacity; ++slot) {
        const PSlot G{&a.part, slot};
        PState P;
        P.state = (int32_t) G.ldi(F_STATE);
        bool live = P.state > 0;
        if (!live) continue;
        pstate_load(P, G);
        for (int turn = 0; turn < a.num_turns && live; ++turn) {
            if (a.flag_monitor == 1) monitor_record(a.mon, P, G);
            const uint64_t* pc = a.prog + a.pc_start;
            const uint64_t* pend = a.prog + a.pc_stop;
            while (pc < pend && live) {
                const uint64_t hw = pc[0];
                const uint32_t hx = (uint32_t) hw;
                const uint32_t op = hx & 0xffu;
                const int32_t aux = (int32_t) (hw >> 32);
                const double* q = reinterpret_cast<const double*>(pc + 1);
                if (a.flag_monitor == 2 && (hx & (XTB_F_START << 8))) monitor_record(a.mon, P, G);
                switch (op) {
                case XTB_OP_NOP: break;
                case XTB_OP_DRIFT: drift_expanded<FRZ>(P, q[0]); break;
                case XTB_OP_DRIFT_EXACT: drift_exact<FRZ>(P, q[0]); break;
                case XTB_OP_MULT: mult_kick(P, q, aux); break;
                case XTB_OP_MULT_H: mult_kick_h<FRZ>(P, q, q + 4, aux & 0xff, (aux >> 8) & 1); break;
                case XTB_OP_CAVITY: cavity_kick<FRZ>(P, G, a, q[0], q[1], q[2], q[3], q[4], aux); break;
                case XTB_OP_RFMULT: rfmult_kick<FRZ>(P, G, a, q, aux); break;
                case XTB_OP_EDGE_LIN: edge_linear(P, q[0], q[1]); break;
                case XTB_OP_SROT: srotation(P, q[0], q[1]); break;
                case XTB_OP_XYSHIFT: P.x += -q[0];  P.y += -q[1]; break;
                case XTB_OP_SSHIFT:
                    drift_exact<FRZ>(P, q[0]);
                    if (!FRZ) { P.zeta += -q[0];  P.s += -q[0]; }
                    break;
                case XTB_OP_YROT: yrotation<FRZ>(P, G, q[0], q[1], q[2]); break;
                case XTB_OP_XROT: xrotation<FRZ>(P, G, q[0], q[1], q[2]); break;
                case XTB_OP_LIMIT_RECT:
                    if (!a.ignore_local
                        && !((P.x >= q[0]) && (P.x <= q[1]) && (P.y >= q[2]) && (P.y <= q[3])))
                        P.state = 0;
                    break;
                case XTB_OP_LIMIT_ELLIPSE:
                    if (!a.ignore_local && !(P.x * P.x * q[1] + P.y * P.y * q[0] <= q[2])) P.state = 0;
                    break;
                case XTB_OP_LIMIT_POLYGON:
                    if (!a.ignore_local && !polygon_contains(P.x, P.y, q, q + aux, aux)) P.state = 0;
                    break;
                case XTB_OP_MONITOR: monitor_record(a.inline_mon[aux], P, G); break;
                case XTB_OP_LAST_TURNS: last_turns_record(a.inline_ltm[aux], P, G); break;
                case XTB_OP_KILL: kill_particle<FRZ>(P, G, aux); break;
                case XTB_OP_SET_STATE: P.state = aux; break;
                case XTB_OP_ADD_S_ZETA: if (!FRZ) { P.s += q[0];  P.zeta += q[0]; } break;
                case XTB_OP_ADD_X: P.x += q[0]; break;
                default: heavy_op<SYNRAD, FRZ>(op, aux, q, P, G, a); break;
                }
                if ((hx & (XTB_F_GLOBAL << 8)) && !a.ignore_global) global_aperture_check(P, a.global_xy_limit);
                if (hx & (XTB_F_END << 8)) {
                    if (P.state > 0) P.at_element += 1;
                    else live = false;
                }
                pc += (hx >> 16);
            }
            if (a.flag_monitor == 2 && live) monitor_record(a.mon, P, G);
            if (a.flag_end_turn_actions > 0 && live) {
                P.at_turn += 1;
                P.at_element = 0;
                if (a.flag_reset_s > 0 && !FRZ) P.s = 0.;
            }
        }
        pstate_store(P, G);
    }
}

extern "C" void xtb_hostsim_track(const uint64_t* words, const uint32_t* elem_offset,
                                  const xtb_particles_t* p, int64_t num_turns, int32_t ele_start,
                                  int32_t num_ele_track, int32_t flag_end_turn_actions,
                                  int32_t flag_reset_s, int32_t flag_monitor, const xtb_monitor_t* mon,
                                  uint64_t track_flags, double global_xy_limit, uint32_t variant,
                                  double line_length, const xtb_monitor_t* inline_mon,
                                  const xtb_last_turns_monitor_t* inline_ltm) {
    XtbTrackArgs a;
    std::memset(&a, 0, sizeof(a));
    a.prog = words;
    a.part = *p;
    if (mon) a.mon = *mon;
    a.inline_mon = inline_mon;
    a.inline_ltm = inline_ltm;
    a.pc_start = elem_offset[ele_start];
    a.pc_stop = elem_offset[ele_start + num_ele_track];
    a.num_turns = (int32_t) num_turns;
    a.flag_end_turn_actions = flag_end_turn_actions;
    a.flag_reset_s = flag_reset_s;
    a.flag_monitor = flag_monitor;
    a.ignore_global = (int32_t) ((track_flags >> XTB_FLAG_IGNORE_GLOBAL_APERTURE) & 1);
    a.ignore_local = (int32_t) ((track_flags >> XTB_FLAG_IGNORE_LOCAL_APERTURE) & 1);
    a.kill_cavity_kick = (int32_t) ((track_flags >> XTB_FLAG_KILL_CAVITY_KICK) & 1);
    a.line_length = line_length;
    a.global_xy_limit = global_xy_limit;
    const bool synrad = variant & XTB_VARIANT_SYNRAD, frz = variant & XTB_VARIANT_FREEZE_LONG;
    if (synrad && frz) run<true, true>(a);
    else if (synrad) run<true, false>(a);
    else if (frz) run<false, true>(a);
    else run<false, false>(a);
}
