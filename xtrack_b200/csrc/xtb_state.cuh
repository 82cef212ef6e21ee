// xtb_state.cuh -- per-thread particle state and kernel argument block.
//
// One thread carries one particle slot through the whole lattice and turn loop
// (same mapping as the reference GPU contexts, xtrack/tracker.py:596), but the
// coordinates that elements touch live in FP64 registers for the entire launch;
// the caller's SoA (xtrack/particles/particles.py:49-83) is read at entry and
// written at exit, at loss, and by monitor records only.  Rarely used fields
// (ptau, beta0, p0c, charge_ratio, ...) stay in the SoA and are accessed
// in place (coalesced: thread i <-> slot i) by the few ops that need them.
#pragma once
#include <stdint.h>
#include "../../include/xtb200.h"
#include "xtb_ops.h"

// indices into xtb_particles_t.field[] (order of particles.py:49-83)
enum XtbField {
    F_P0C = 0, F_GAMMA0, F_BETA0, F_S, F_ZETA, F_X, F_Y, F_PX, F_PY, F_PTAU, F_DELTA,
    F_RPP, F_RVV, F_CHI, F_CHARGE_RATIO, F_WEIGHT, F_AX, F_AY, F_SPIN_X, F_SPIN_Y,
    F_SPIN_Z, F_ANOM, F_PDG_ID, F_PARTICLE_ID, F_AT_ELEMENT, F_AT_TURN, F_STATE,
    F_PARENT_ID, F_RNG_S1, F_RNG_S2, F_RNG_S3, F_RNG_S4
};
#define XTB_N_F64 22
#define XTB_N_I64 6

struct XtbTrackArgs {
    const uint64_t* prog;        // device program (words)
    const uint32_t* tile_off;    // [n_tiles+1] word offsets of the tile boundaries
    const xtb_monitor_t* inline_mon;             // device tables for OP_MONITOR / OP_LAST_TURNS
    const xtb_last_turns_monitor_t* inline_ltm;
    xtb_particles_t part;
    xtb_monitor_t mon;           // turn-by-turn monitor (flag_monitor != 0)
    uint32_t pc_start, pc_stop;  // word range to execute each turn
    int32_t tile_first, tile_last;   // tiles covering [pc_start, pc_stop)
    int32_t num_turns;
    int32_t flag_end_turn_actions, flag_reset_s, flag_monitor;
    int32_t ignore_global, ignore_local, kill_cavity_kick;
    double line_length;
    double global_xy_limit;
};

struct PState {
    double x, px, y, py, zeta, delta, rpp, rvv, rv0v, chi, s;
    int64_t at_turn;
    int32_t at_element;
    int32_t state;
};

// In-place access to the caller's SoA for one slot.
struct PSlot {
    const xtb_particles_t* p;
    int64_t i;
    __device__ __forceinline__ double ld(int f) const {
        return reinterpret_cast<const double*>(p->field[f])[i];
    }
    __device__ __forceinline__ void st(int f, double v) const {
        reinterpret_cast<double*>(p->field[f])[i] = v;
    }
    __device__ __forceinline__ int64_t ldi(int f) const {
        return reinterpret_cast<const int64_t*>(p->field[f])[i];
    }
    __device__ __forceinline__ void sti(int f, int64_t v) const {
        reinterpret_cast<int64_t*>(p->field[f])[i] = v;
    }
    __device__ __forceinline__ uint32_t ldu(int f) const {
        return reinterpret_cast<const uint32_t*>(p->field[f])[i];
    }
    __device__ __forceinline__ void stu(int f, uint32_t v) const {
        reinterpret_cast<uint32_t*>(p->field[f])[i] = v;
    }
};

__device__ __forceinline__ void pstate_load(PState& P, const PSlot& G) {
    P.x = G.ld(F_X);  P.px = G.ld(F_PX);  P.y = G.ld(F_Y);  P.py = G.ld(F_PY);
    P.zeta = G.ld(F_ZETA);  P.delta = G.ld(F_DELTA);  P.rpp = G.ld(F_RPP);
    P.rvv = G.ld(F_RVV);  P.chi = G.ld(F_CHI);  P.s = G.ld(F_S);
    P.rv0v = 1. / P.rvv;
    P.at_turn = G.ldi(F_AT_TURN);
    P.at_element = (int32_t) G.ldi(F_AT_ELEMENT);
}

// Write the register-resident fields back (exit, loss, monitor snapshots read
// the rest from the SoA, where write-through keeps them current).
__device__ __forceinline__ void pstate_store(const PState& P, const PSlot& G) {
    G.st(F_X, P.x);  G.st(F_PX, P.px);  G.st(F_Y, P.y);  G.st(F_PY, P.py);
    G.st(F_ZETA, P.zeta);  G.st(F_DELTA, P.delta);  G.st(F_RPP, P.rpp);
    G.st(F_RVV, P.rvv);  G.st(F_S, P.s);
    G.sti(F_AT_TURN, P.at_turn);
    G.sti(F_AT_ELEMENT, (int64_t) P.at_element);
    G.sti(F_STATE, (int64_t) P.state);
}
