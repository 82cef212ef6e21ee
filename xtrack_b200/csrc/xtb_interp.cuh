// xtb_interp.cuh -- the op interpreter: monitors, rarely used ops, and the loop that
// executes a range of program words on the particles one thread carries.
//
// Shared verbatim by the CUDA kernel (xtb_kernel.cuh) and by the host build of
// the device code that the CPU test tier uses (tests/hostsim), so that the op
// semantics checked against the oracle without a GPU are the ones the kernel runs.
//
// Element loop semantics restated from the generated `track_line` kernel,
// xtrack/tracker.py:646-711: per element  track -> [global aperture check for
// statically thick classes :681-689] -> is-active check :702-705 (a lost particle
// stops; at_element stays on the element where it was lost) -> at_element++ :707-711.
#pragma once
#include "xtb_state.cuh"
#include "xtb_thin.cuh"
#ifdef XTB_WITH_HEAVY
#include "xtb_thick.cuh"
#endif

// ParticlesMonitor record: LocalParticle_to_Particles(part, data, store_at, 0)
// of monitors/particles_monitor.h:13-77 -- all 32 per-particle fields.
static __device__ __noinline__ void monitor_store(const xtb_monitor_t& m, const int64_t at,
                                                  const PState& P, const PSlot& G) {
    auto D = [&](int f) { return reinterpret_cast<double*>(m.field[f]) + at; };
    auto I = [&](int f) { return reinterpret_cast<int64_t*>(m.field[f]) + at; };
    auto U = [&](int f) { return reinterpret_cast<uint32_t*>(m.field[f]) + at; };
    *D(F_P0C) = G.ld(F_P0C);  *D(F_GAMMA0) = G.ld(F_GAMMA0);  *D(F_BETA0) = G.ld(F_BETA0);
    *D(F_S) = P.s;  *D(F_ZETA) = P.zeta;  *D(F_X) = P.x;  *D(F_Y) = P.y;
    *D(F_PX) = P.px;  *D(F_PY) = P.py;  *D(F_PTAU) = G.ld(F_PTAU);  *D(F_DELTA) = P.delta;
    *D(F_RPP) = P.rpp;  *D(F_RVV) = P.rvv;  *D(F_CHI) = P.chi;
    *D(F_CHARGE_RATIO) = G.ld(F_CHARGE_RATIO);  *D(F_WEIGHT) = G.ld(F_WEIGHT);
    *D(F_AX) = G.ld(F_AX);  *D(F_AY) = G.ld(F_AY);  *D(F_SPIN_X) = G.ld(F_SPIN_X);
    *D(F_SPIN_Y) = G.ld(F_SPIN_Y);  *D(F_SPIN_Z) = G.ld(F_SPIN_Z);  *D(F_ANOM) = G.ld(F_ANOM);
    *I(F_PDG_ID) = G.ldi(F_PDG_ID);  *I(F_PARTICLE_ID) = G.ldi(F_PARTICLE_ID);
    *I(F_AT_ELEMENT) = (int64_t) P.at_element;  *I(F_AT_TURN) = P.at_turn;
    *I(F_STATE) = (int64_t) P.state;  *I(F_PARENT_ID) = G.ldi(F_PARENT_ID);
    *U(F_RNG_S1) = G.ldu(F_RNG_S1);  *U(F_RNG_S2) = G.ldu(F_RNG_S2);
    *U(F_RNG_S3) = G.ldu(F_RNG_S3);  *U(F_RNG_S4) = G.ldu(F_RNG_S4);
}

// monitors/particles_monitor.h:13-77 (P.at_element must hold the CURRENT element index)
static __device__ __noinline__ void monitor_record(const xtb_monitor_t& m, const PState& P,
                                                   const PSlot& G) {
    const int64_t n_turns_record = m.stop_at_turn - m.start_at_turn;
    const int64_t at_turn = m.ebe_mode ? (int64_t) P.at_element : P.at_turn;
    const int64_t particle_id = G.ldi(F_PARTICLE_ID);
    if (m.n_repetitions == 1) {
        if (at_turn >= m.start_at_turn && at_turn < m.stop_at_turn
            && particle_id < m.part_id_end && particle_id >= m.part_id_start) {
            monitor_store(m, n_turns_record * (particle_id - m.part_id_start) + at_turn - m.start_at_turn,
                          P, G);
        }
    } else if (m.n_repetitions > 1) {
        if (at_turn < m.start_at_turn) return;
        const int64_t i_frame = (at_turn - m.start_at_turn) / m.repetition_period;
        if (i_frame < m.n_repetitions && at_turn >= m.start_at_turn + i_frame * m.repetition_period
            && at_turn < m.stop_at_turn + i_frame * m.repetition_period
            && particle_id < m.part_id_end && particle_id >= m.part_id_start) {
            monitor_store(m,
                          n_turns_record * (m.part_id_end - m.part_id_start) * i_frame
                              + n_turns_record * (particle_id - m.part_id_start)
                              + (at_turn - i_frame * m.repetition_period) - m.start_at_turn,
                          P, G);
        }
    }
}

// LastTurnsMonitor_track_local_particle, monitors/last_turns_monitor.h:16-55
static __device__ __noinline__ void last_turns_record(const xtb_last_turns_monitor_t& m,
                                                      const PState& P, const PSlot& G) {
    const int64_t particle_id = G.ldi(F_PARTICLE_ID);
    const int64_t at_turn = P.at_turn;
    const int64_t stop = m.particle_id_start + m.num_particles;
    if (at_turn >= 0 && at_turn % m.every_n_turns == 0 && m.particle_id_start <= particle_id
        && particle_id < stop) {
        const int64_t offset = (at_turn / m.every_n_turns) % m.n_last_turns;
        const int64_t ip = particle_id - m.particle_id_start;
        const int64_t slot = m.n_last_turns * ip + offset;
        reinterpret_cast<uint32_t*>(m.field[0])[ip] = (uint32_t) offset;
        reinterpret_cast<uint32_t*>(m.field[1])[slot] = (uint32_t) particle_id;
        reinterpret_cast<uint32_t*>(m.field[2])[slot] = (uint32_t) at_turn;
        reinterpret_cast<float*>(m.field[3])[slot] = (float) P.x;
        reinterpret_cast<float*>(m.field[4])[slot] = (float) P.px;
        reinterpret_cast<float*>(m.field[5])[slot] = (float) P.y;
        reinterpret_cast<float*>(m.field[6])[slot] = (float) P.py;
        reinterpret_cast<float*>(m.field[7])[slot] = (float) P.delta;
        reinterpret_cast<float*>(m.field[8])[slot] = (float) P.zeta;
    }
}

// Generic thin ops, kept out of line so that they do not weigh on the hot loop's
// register allocation.  `P.at_element` holds the current element index here.
template <bool FRZ>
static __device__ __noinline__ void generic_op(const uint32_t op, const int32_t aux,
                                               const double* __restrict__ q, PState& P,
                                               const PSlot& G, const XtbTrackArgs& a,
                                               const bool live) {
    switch (op) {
    case XTB_OP_DRIFT:
        drift_expanded<FRZ>(P, q[0]);
        break;
    case XTB_OP_DRIFT_EXACT:
        drift_exact<FRZ>(P, q[0]);
        break;
    case XTB_OP_MULT:
        mult_kick(P, q, aux);
        break;
    case XTB_OP_MULT_H:
        mult_kick_h<FRZ>(P, q, q + 4, aux & 0xff, (aux >> 8) & 1);
        break;
    case XTB_OP_CAVITY:
        cavity_kick<FRZ>(P, G, a, q[0], q[1], q[2], q[3], q[4], aux);
        break;
    case XTB_OP_RFMULT:
        rfmult_kick<FRZ>(P, G, a, q, aux);
        break;
    case XTB_OP_EDGE_LIN:
        edge_linear(P, q[0], q[1]);
        break;
    case XTB_OP_SROT:
        srotation(P, q[0], q[1]);
        break;
    case XTB_OP_XYSHIFT:
        P.x += -q[0];
        P.y += -q[1];
        break;
    case XTB_OP_SSHIFT:
        drift_exact<FRZ>(P, q[0]);
        if (!FRZ) { P.zeta += -q[0];  P.s += -q[0]; }
        break;
    case XTB_OP_YROT:
        yrotation<FRZ>(P, G, q[0], q[1], q[2]);
        break;
    case XTB_OP_XROT:
        xrotation<FRZ>(P, G, q[0], q[1], q[2]);
        break;
    case XTB_OP_LIMIT_RECT:
        if (!a.ignore_local) {
            const bool in = (P.x >= q[0]) && (P.x <= q[1]) && (P.y >= q[2]) && (P.y <= q[3]);
            if (!in) P.state = 0;
        }
        break;
    case XTB_OP_LIMIT_ELLIPSE:
        if (!a.ignore_local) {
            const double temp = P.x * P.x * q[1] + P.y * P.y * q[0];
            if (!(temp <= q[2])) P.state = 0;
        }
        break;
    case XTB_OP_LIMIT_POLYGON:
        if (!a.ignore_local && !polygon_contains(P.x, P.y, q, q + aux, aux)) P.state = 0;
        break;
    case XTB_OP_MONITOR:
        if (live) monitor_record(a.inline_mon[aux], P, G);
        break;
    case XTB_OP_LAST_TURNS:
        if (live) last_turns_record(a.inline_ltm[aux], P, G);
        break;
    case XTB_OP_KILL:
        kill_particle<FRZ>(P, G, aux);
        break;
    case XTB_OP_SET_STATE:
        P.state = aux;
        break;
    case XTB_OP_ADD_S_ZETA:
        if (!FRZ) { P.s += q[0];  P.zeta += q[0]; }
        break;
    case XTB_OP_ADD_X:
        P.x += q[0];
        break;
    default:
        break;
    }
}

// A lane without a particle to track executes the ops on this benign on-axis state
// (no per-op predication in the hot loop); nothing of it is ever written back.
template <class S>
__device__ __forceinline__ void pstate_benign(S& P) {
    P.x = P.px = P.y = P.py = P.zeta = P.delta = P.s = 0.;
    P.rpp = P.rv0v = P.chi = 1.;
    P.state = 1;
}
__device__ __forceinline__ void pstate_benign(PState& P) {
    P.x = P.px = P.y = P.py = P.zeta = P.delta = P.s = 0.;
    P.rpp = P.rvv = P.rv0v = P.chi = 1.;
    P.state = 1;
    P.at_turn = 0;
    P.at_element = 0;
}

// ---- cold paths, out of line: their register needs must not weigh on the hot loop ----
// one generic / heavy op on one lane; returns false when the particle was lost and stored
template <bool HEAVY, bool SYNRAD, bool FRZ, class S>
static __device__ __noinline__ bool xtb_slow_op(S& Pk, const PSlot Gk, const XtbPass ps,
                                                const uint32_t eidx, const uint32_t h,
                                                const int32_t aux, const double* __restrict__ q,
                                                const XtbTrackArgs& a) {
    const uint32_t op = h & 0xffu;
    bool alive = true;
    PState T = pstate_full(Pk, Gk, ps, eidx);
    if ((a.flag_monitor == 2) && (h & (XTB_F_START << 8))) monitor_record(a.mon, T, Gk);
#ifdef XTB_WITH_HEAVY
    if (HEAVY && op >= XTB_HEAVY_FIRST) heavy_op<SYNRAD, FRZ>(op, aux, q, T, Gk, a);
    else
#endif
        generic_op<FRZ>(op, aux, q, T, Gk, a, true);
    if ((h & (XTB_F_GLOBAL << 8)) && !a.ignore_global) global_aperture_check(T, a.global_xy_limit);
    if ((h & (XTB_F_END << 8)) && T.state <= 0) {
        // tracker.py:702-711: a lost particle stops here, at_element stays on the
        // element where it was lost
        pstate_store(T, Gk);
        alive = false;
        T.state = 1;
    }
    pstate_back(Pk, T, Gk);
    return alive;
}

struct __align__(16) xtb_w128 { uint64_t x, y; };
struct __align__(16) xtb_d2 { double x, y; };

#ifndef XTB_GLOBAL_FILTER
#define XTB_GLOBAL_FILTER 2
#endif
#ifndef XTB_UNLIKELY
#define XTB_UNLIKELY(c) __builtin_expect(!!(c), 0)
#endif

// The particles one thread carries, as they sit in (thread-local) memory between tiles.
template <int NPT, class S>
struct XtbLanes {
    S P[NPT];
    PSlot G[NPT];
    bool live[NPT];      // lane k still tracks a real, active particle
    uint32_t eidx;       // elements completed so far in this pass (identical for all threads)
};

// Executes the ops of one tile, from word `off` of `tb` up to the XTB_OP_END sentinel,
// on the NPT particles of this thread.
//
// Deliberately NOT inlined into the kernel: the function loads the lanes from `lb` into
// registers, runs the whole tile on registers and writes them back, so that the register
// allocation of the hot loop is not disturbed by the variables of the turn / tile loops
// around it (with it inlined, ptxas spilled the loop's own offset and header words).
// The round trip through local memory costs ~50 instructions per tile of ~10^4.
//
// A particle found lost at the end of an element is written back to the caller's SoA at
// once (pstate_store) and its lane goes on as a dead lane: its registers keep evolving
// but are never stored again.  The fast path never writes `state`: the loss tests branch
// to cold code that stores the particle and clears live[k] -- nothing else, so that no
// register of the hot loop is redefined on a cold path.  Dead lanes are reset to the
// benign state at the end of the tile.
//
// Hot-loop shape: the header of the NEXT op and the first four parameter words of THIS
// op are loaded before any arithmetic (their shared-memory latency is covered by it);
// the drift prefix is one bit test, the main op one switch; the loss test is one
// not-taken branch; no per-op bounds test (the tile ends at a sentinel op).
template <int NPT, bool HEAVY, bool SYNRAD, bool FRZ, bool CHI1, class S>
static __device__ __noinline__ void xtb_run_tile(const uint64_t* __restrict__ tb, uint32_t off,
                                                 XtbLanes<NPT, S>* __restrict__ lb,
                                                 const XtbPass ps, const XtbTrackArgs& a) {
    S P[NPT];
    PSlot G[NPT];
    bool live[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        P[k] = lb->P[k];
        G[k] = lb->G[k];
        live[k] = lb->live[k];
    }
    uint32_t eidx = lb->eidx;
    const double lim = a.global_xy_limit;
    // high word of the limit for the integer pre-filter (0: always take the exact test)
    const uint32_t lim_hi = (lim > 0.) ? (uint32_t) __double2hiint(lim) : 0u;

    // lane k lost in the current element with state code `code`
    auto retire = [&](const int k, const int32_t code) {
        PState T = pstate_full(P[k], G[k], ps, eidx);
        T.state = code;
        pstate_store(T, G[k]);
        live[k] = false;
    };
    // global_aperture_check (local_particle_custom_api.h:262-289) + is-active check.
    // The hot path only needs "is any lane outside?"; XTB_GLOBAL_FILTER selects how:
    //   0  the reference's four comparisons per particle (FP64 pipe: 4 DSETP)
    //   1  |x| <= lim && |y| <= lim (2 DSETP; same truth table, NaN -> outside)
    //   2  integer pre-filter on the high words: |x| > lim implies hi32(|x|) >= hi32(lim),
    //      so lanes whose high words are all below hi32(lim) are inside for sure; no FP64
    //      instruction at all, the exact test runs on the cold path only.
    auto global_check = [&]() {
#if XTB_GLOBAL_FILTER == 2
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            m = max(m, (uint32_t) __double2hiint(P[k].x) & 0x7fffffffu);
            m = max(m, (uint32_t) __double2hiint(P[k].y) & 0x7fffffffu);
        }
        const bool any = m >= lim_hi;
#else
        bool any = false;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
#if XTB_GLOBAL_FILTER == 1
            any = any | !((fabs(P[k].x) <= lim) && (fabs(P[k].y) <= lim));
#else
            any = any | outside_global(P[k], lim);
#endif
        }
#endif
        if (XTB_UNLIKELY(any)) {
            if (!a.ignore_global) {
#pragma unroll
                for (int k = 0; k < NPT; ++k)
                    if (outside_global(P[k], lim) && live[k]) retire(k, -1);
            }
        }
    };
    // the Drift element in front of an op: track, global check, loss check, at_element++
    auto prefix = [&](const double L) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(P[k], L);
        global_check();
        eidx += 1;
    };

    xtb_w128 hw = *reinterpret_cast<const xtb_w128*>(tb + off);
    for (;;) {
        const uint32_t h = (uint32_t) hw.x;
        const uint32_t op = h & 0xffu;
        const double L = __longlong_as_double((long long) hw.y);
        const uint64_t* __restrict__ cur = tb + off;
        off += (h >> 16);
        // first four parameter words of this op and the header of the next one
        const xtb_d2 c0 = *reinterpret_cast<const xtb_d2*>(cur + 2);
        const xtb_d2 c1 = *reinterpret_cast<const xtb_d2*>(cur + 4);
        hw = *reinterpret_cast<const xtb_w128*>(tb + off);

        if (op < XTB_OP_END) {
            // ---------------- fast ops ----------------
            if (op & XTB_OPBIT_DRIFT) prefix(L);
            switch (op & (XTB_OPBIT_DRIFT - 1)) {
            case XTB_OP_MULT0: {
                const double c[2] = {c0.x, c0.y};
#pragma unroll
                for (int k = 0; k < NPT; ++k) mult_kick_c<0, CHI1>(P[k], c);
                break;
            }
            case XTB_OP_MULT1: {
                const double c[4] = {c0.x, c0.y, c1.x, c1.y};
#pragma unroll
                for (int k = 0; k < NPT; ++k) mult_kick_c<1, CHI1>(P[k], c);
                break;
            }
            case XTB_OP_MULTN: {
                // order >= 2: Horner loop, one coefficient pair per step from shared memory
                // (same arithmetic as mult_kick_c, not unrolled: small register footprint)
                const uint32_t order = (uint32_t) (cur[0] >> 32);
                double dpx[NPT], dpy[NPT];
#pragma unroll
                for (int k = 0; k < NPT; ++k) {
                    dpx[k] = CHI1 ? c0.x : P[k].chi * c0.x;
                    dpy[k] = CHI1 ? c0.y : P[k].chi * c0.y;
                }
                xtb_d2 cc = c1;
                for (uint32_t i = 1; i <= order; ++i) {
                    const xtb_d2 cn = *reinterpret_cast<const xtb_d2*>(cur + 4 + 2 * i);
#pragma unroll
                    for (int k = 0; k < NPT; ++k) {
                        const double zre = dpx[k] * P[k].x - dpy[k] * P[k].y;
                        const double zim = dpx[k] * P[k].y + dpy[k] * P[k].x;
                        dpx[k] = (CHI1 ? cc.x : P[k].chi * cc.x) + zre;
                        dpy[k] = (CHI1 ? cc.y : P[k].chi * cc.y) + zim;
                    }
                    cc = cn;
                }
#pragma unroll
                for (int k = 0; k < NPT; ++k) {
                    P[k].px += -dpx[k];
                    P[k].py += dpy[k];
                }
                break;
            }
            case XTB_OP_MULTH0: {
#pragma unroll
                for (int k = 0; k < NPT; ++k) mult_kick_h0<FRZ, CHI1>(P[k], c0.x, c0.y, c1.x, c1.y);
                break;
            }
            case XTB_OP_EDGE: {
#pragma unroll
                for (int k = 0; k < NPT; ++k) edge_linear_c<CHI1>(P[k], c0.x, c0.y);
                break;
            }
            case XTB_OP_RECT: {
                bool any = false;
#pragma unroll
                for (int k = 0; k < NPT; ++k)
                    any = any | !((P[k].x >= c0.x) && (P[k].x <= c0.y) && (P[k].y >= c1.x)
                                  && (P[k].y <= c1.y));
                if (XTB_UNLIKELY(any)) {
                    if (!a.ignore_local) {
#pragma unroll
                        for (int k = 0; k < NPT; ++k)
                            if (!((P[k].x >= c0.x) && (P[k].x <= c0.y) && (P[k].y >= c1.x)
                                  && (P[k].y <= c1.y)) && live[k])
                                retire(k, 0);
                    }
                }
                break;
            }
            case XTB_OP_ELLIPSE: {
                bool any = false;
#pragma unroll
                for (int k = 0; k < NPT; ++k)
                    any = any | !(P[k].x * P[k].x * c0.y + P[k].y * P[k].y * c0.x <= c1.x);
                if (XTB_UNLIKELY(any)) {
                    if (!a.ignore_local) {
#pragma unroll
                        for (int k = 0; k < NPT; ++k)
                            if (!(P[k].x * P[k].x * c0.y + P[k].y * P[k].y * c0.x <= c1.x) && live[k])
                                retire(k, 0);
                    }
                }
                break;
            }
            case XTB_OP_FDRIFT: {
#pragma unroll
                for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(P[k], c0.x);
                global_check();
                break;
            }
            default:      // XTB_OP_NOP
                break;
            }
            eidx += 1;
        } else {
            // ---------------- sentinel, generic and heavy ops ----------------
            if (op == XTB_OP_END) break;
            if (h & (XTB_F_DRIFT << 8)) prefix(L);
            const int32_t aux = (int32_t) (cur[0] >> 32);
            const double* __restrict__ q = reinterpret_cast<const double*>(cur + 2);
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                if (!live[k]) continue;      // these bodies touch the caller's SoA
                S Pc = P[k];                 // (copy: P itself must never have its address taken)
                live[k] = xtb_slow_op<HEAVY, SYNRAD, FRZ>(Pc, G[k], ps, eidx, h, aux, q, a);
                P[k] = Pc;
            }
            if (h & (XTB_F_END << 8)) eidx += 1;
        }
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        if (!live[k]) pstate_benign(P[k]);      // lanes that died in this tile
        lb->P[k] = P[k];
        lb->live[k] = live[k];
    }
    lb->eidx = eidx;
}

// End of a pass over the element range: increment_at_turn (local_particle_custom_api.h:
// 76-84) expressed on the block-uniform counters, s reset on the lanes.
template <int NPT, bool FRZ, class S>
__device__ __forceinline__ void xtb_end_pass(S (&P)[NPT], XtbPass& ps, const uint32_t eidx,
                                             const XtbTrackArgs& a) {
    if (a.flag_end_turn_actions > 0) {
        ps.turn_inc += 1;
        ps.el_off = 0;
        ps.el_reset = 1;
        if (a.flag_reset_s > 0 && !FRZ) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) P[k].s = 0.;
        }
    } else {
        ps.el_off += eidx;
    }
}
