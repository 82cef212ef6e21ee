// xtb_kernel.cuh -- the fused turn x element tracking kernel for sm_100a.
//
// Replaces the generated `track_line` kernel of xtrack/tracker.py:545-749.
// Design:
//   * one thread = one particle slot; coordinates live in FP64 registers over
//     the whole launch (PState), SoA traffic only at entry / exit / loss /
//     monitor records;
//   * the lowered lattice is streamed through shared memory in tiles with
//     1-D bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP), double
//     buffered; all lanes read the same op -> shared-memory broadcast;
//   * a lost particle is written back the moment it is lost and its lanes
//     continue on a benign on-axis state (no per-op predication); a warp whose
//     lanes are all done skips whole tiles (warp vote);
//   * template switches select the kernel variant: HEAVY (thick-magnet ops
//     compiled in), SYNRAD, FRZ (freeze_longitudinal).  The translation unit is
//     compiled twice, with and without FMA contraction (xtb_kernel_inst.cu).
#pragma once
#include <cuda_runtime.h>
#include "xtb_state.cuh"
#include "xtb_thin.cuh"
#ifdef XTB_WITH_HEAVY
#include "xtb_thick.cuh"
#endif

#define XTB_THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XTB_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra XTB_WAIT_DONE;\n"
        "bra XTB_WAIT_LOOP;\n"
        "XTB_WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

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

static __device__ __noinline__ void monitor_record(const xtb_monitor_t& m, const PState& P, const PSlot& G) {
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
static __device__ __noinline__ void last_turns_record(const xtb_last_turns_monitor_t& m, const PState& P,
                                               const PSlot& G) {
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

// Rare ops, kept out of line so that they do not weigh on the hot loop's
// register allocation.
template <bool FRZ>
__device__ __noinline__ void rare_op(const uint32_t op, const int32_t aux,
                                     const double* __restrict__ q, PState& P, const PSlot& G,
                                     const XtbTrackArgs& a, const bool live) {
    switch (op) {
    case XTB_OP_CAVITY:
        cavity_kick<FRZ>(P, G, a, q[0], q[1], q[2], q[3], q[4], aux);
        break;
    case XTB_OP_RFMULT:
        rfmult_kick<FRZ>(P, G, a, q, aux);
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

// EXACT only distinguishes the symbols of the two builds of this translation unit
// (template instantiations are COMDAT: identical names would be merged at link time).
template <bool HEAVY, bool SYNRAD, bool FRZ, bool EXACT>
__global__ void __launch_bounds__(XTB_THREADS, HEAVY ? 1 : 3)
xtb_track_kernel(const __grid_constant__ XtbTrackArgs a) {
    __shared__ __align__(128) uint64_t tile[XTB_NUM_BUF][XTB_TILE_WORDS];
    __shared__ __align__(8) uint64_t full_bar[XTB_NUM_BUF];

    const int64_t slot = (int64_t) blockIdx.x * XTB_THREADS + threadIdx.x;
    const PSlot G{&a.part, slot};
    PState P;
    bool live = false;          // this lane still tracks a real particle
    if (slot < a.part.capacity) {
        P.state = (int32_t) G.ldi(F_STATE);
        live = P.state > 0;     // check_is_active (GPU), local_particle_custom_api.h:188
    } else {
        P.state = 0;
    }
    if (live) {
        pstate_load(P, G);
    } else {
        P.x = P.px = P.y = P.py = P.zeta = P.delta = P.s = 0.;
        P.rpp = P.rvv = P.rv0v = P.chi = 1.;
        P.at_turn = 0;  P.at_element = 0;
    }
    if (!__syncthreads_or(live)) return;     // nothing to track in this block

    const int n_tiles = a.tile_last - a.tile_first + 1;
    const bool resident = (n_tiles == 1);     // whole range fits one tile: load once
    const int64_t total_steps = resident ? 1 : (int64_t) a.num_turns * n_tiles;

    auto issue = [&](int64_t step) {          // one elected thread
        const int k = a.tile_first + (int) (step % n_tiles);
        const int b = (int) (step % XTB_NUM_BUF);
        const uint32_t w0 = a.tile_off[k], w1 = a.tile_off[k + 1];
        const uint32_t bytes = (w1 - w0) * 8u;
        mbar_expect_tx(&full_bar[b], bytes);
        bulk_g2s(&tile[b][0], a.prog + w0, bytes, &full_bar[b]);
    };

    if (threadIdx.x == 0) {
        for (int b = 0; b < XTB_NUM_BUF; ++b) mbar_init(&full_bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int64_t st = 0; st < XTB_NUM_BUF && st < total_steps; ++st) issue(st);
    }

    const double lim = a.global_xy_limit;
    int64_t step = 0;
    for (int turn = 0; turn < a.num_turns; ++turn) {
        // every particle of this block lost: drain the copies in flight and stop
        if (turn > 0 && !__syncthreads_or(live)) {
            if (!resident) {
                for (int64_t st = step; st < step + XTB_NUM_BUF && st < total_steps; ++st) {
                    mbar_wait(&full_bar[st % XTB_NUM_BUF], (uint32_t) ((st / XTB_NUM_BUF) & 1));
                }
            }
            break;
        }
        // xtrack/tracker.py:639-641: the turn-by-turn record precedes the turn
        if (a.flag_monitor == 1 && live) { const PState T = P;  monitor_record(a.mon, T, G); }

        for (int j = 0; j < n_tiles; ++j, ++step) {
            const int k = a.tile_first + j;
            const int b = resident ? 0 : (int) (step % XTB_NUM_BUF);
            if (!resident || turn == 0) {
                mbar_wait(&full_bar[b], (uint32_t) ((step / XTB_NUM_BUF) & 1));
            }
            if (__any_sync(0xffffffffu, live)) {
                const uint32_t w0 = a.tile_off[k], w1 = a.tile_off[k + 1];
                const uint32_t lo = max(w0, a.pc_start), hi = min(w1, a.pc_stop);
                const uint64_t* __restrict__ pc = &tile[b][0] + (lo - w0);
                const uint64_t* const pend = &tile[b][0] + (hi - w0);
                while (pc < pend) {
                    // header + first parameter in one 16-byte broadcast load
                    const ulonglong2 hw = *reinterpret_cast<const ulonglong2*>(pc);
                    const uint2 h = make_uint2((uint32_t) hw.x, (uint32_t) (hw.x >> 32));
                    const uint32_t op = h.x & 0xffu;
                    const int32_t aux = (int32_t) h.y;
                    const double* __restrict__ q = reinterpret_cast<const double*>(pc + 1);
                    const double q0 = __longlong_as_double((long long) hw.y);
                    if (a.flag_monitor == 2 && (h.x & (XTB_F_START << 8)) && live) {
                        const PState T = P;  monitor_record(a.mon, T, G);
                    }
                    if (op == XTB_OP_DRIFT) {
                        drift_expanded<FRZ>(P, q0);
                    } else if (op == XTB_OP_MULT) {
                        mult_kick(P, q, aux);
                    } else if (op == XTB_OP_MULT_H) {
                        mult_kick_h<FRZ>(P, q, q + 4, aux & 0xff, (aux >> 8) & 1);
                    } else if (op == XTB_OP_NOP) {
                    } else if (op == XTB_OP_EDGE_LIN) {
                        edge_linear(P, q[0], q[1]);
                    } else if (op == XTB_OP_LIMIT_RECT) {
                        if (!a.ignore_local) {
                            const bool in = (P.x >= q[0]) && (P.x <= q[1]) && (P.y >= q[2]) && (P.y <= q[3]);
                            if (!in) P.state = 0;
                        }
                    } else if (op == XTB_OP_LIMIT_ELLIPSE) {
                        if (!a.ignore_local) {
                            const double temp = P.x * P.x * q[1] + P.y * P.y * q[0];
                            if (!(temp <= q[2])) P.state = 0;
                        }
                    } else if (op == XTB_OP_SROT) {
                        srotation(P, q[0], q[1]);
                    } else if (op == XTB_OP_XYSHIFT) {
                        P.x += -q[0];
                        P.y += -q[1];
                    } else if (op == XTB_OP_DRIFT_EXACT) {
                        drift_exact<FRZ>(P, q[0]);
                    }
#ifdef XTB_WITH_HEAVY
                    else if (HEAVY && op >= XTB_HEAVY_FIRST) {
                        PState T = P;  heavy_op<SYNRAD, FRZ>(op, aux, q, T, G, a);  P = T;
                    }
#endif
                    else {
                        PState T = P;  rare_op<FRZ>(op, aux, q, T, G, a, live);  P = T;
                    }
                    if (h.x & (XTB_F_GLOBAL << 8)) {
                        if (!a.ignore_global) global_aperture_check(P, lim);
                    }
                    if (h.x & (XTB_F_END << 8)) {
                        // tracker.py:702-711: a lost particle stops here, at_element
                        // stays on the element where it was lost
                        if (live) {
                            if (P.state > 0) {
                                P.at_element += 1;
                            } else {
                                pstate_store(P, G);
                                live = false;
                                P.x = P.px = P.y = P.py = P.zeta = P.delta = 0.;
                                P.rpp = P.rvv = P.rv0v = 1.;
                            }
                        }
                    }
                    pc += (h.x >> 16);
                }
            }
            if (!resident) {
                __syncthreads();       // every warp is done with buffer b
                if (threadIdx.x == 0 && step + XTB_NUM_BUF < total_steps) issue(step + XTB_NUM_BUF);
            }
        }
        if (a.flag_monitor == 2 && live) { const PState T = P;  monitor_record(a.mon, T, G); }
        // increment_at_turn, local_particle_custom_api.h:76-84
        if (a.flag_end_turn_actions > 0 && live) {
            P.at_turn += 1;
            P.at_element = 0;
            if (a.flag_reset_s > 0 && !FRZ) P.s = 0.;
        }
    }
    if (live) pstate_store(P, G);
}
