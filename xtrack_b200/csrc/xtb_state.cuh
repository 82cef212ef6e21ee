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
    const uint64_t* prog;        // device program image: tiles, each closed by an XTB_OP_END op
    const uint32_t* tile_off;    // [n_tiles+1] word offsets of the tile images
    const xtb_monitor_t* inline_mon;             // device tables for OP_MONITOR / OP_LAST_TURNS
    const xtb_last_turns_monitor_t* inline_ltm;
    xtb_particles_t part;
    xtb_monitor_t mon;           // turn-by-turn monitor (flag_monitor != 0)
    uint32_t pc_start, pc_stop;  // word range (image offsets) to execute each turn
    int32_t tile_first, tile_last;   // tiles covering [pc_start, pc_stop)
    int32_t num_turns;
    uint32_t num_ele_track;          // elements per pass
    int32_t flag_end_turn_actions, flag_reset_s, flag_monitor;
    int32_t ignore_global, ignore_local, kill_cavity_kick;
    int32_t rng_philox;          // the particles' generator state is (key, counter) of Philox4x32-10
    int32_t aperture_prefilter;  // the program holds fast aperture ops (RECT / ELLIPSE)
    double line_length;
    double global_xy_limit;
    const double* synrad_tables; // quantum-kick model: inverse-CDF tables (xtb_thick.cuh::QkTables)
    // particle slots [slot_begin, slot_end) of the caller's SoA handled by this grid (a
    // track call may be split into several grids, see xtb_kernel_inst.cu::launch)
    int64_t slot_begin, slot_end;
};

// Full per-particle state: what the generic / thick-magnet op bodies work on.
struct PState {
    double x, px, y, py, zeta, delta, rpp, rvv, rv0v, chi, s;
    int64_t at_turn;
    int32_t at_element;
    int32_t state;
};

// Slim state of the thin kernels' hot loop: only what the fast ops touch.  rvv stays in
// the caller's SoA (written through whenever the energy changes, like ptau); at_turn and
// at_element are reconstructed from the SoA's entry values plus block-uniform counters
// (XtbPass) whenever a full PState is needed (loss, monitor record, generic op, exit).
struct PHot {
    double x, px, y, py, zeta, delta, rpp, rv0v, chi, s;
    int32_t state;
};

// Block-uniform bookkeeping of the turn loop (identical for all threads of a launch).
struct XtbPass {
    int32_t turn_inc;     // end-of-turn increments of at_turn so far in this launch
    uint32_t el_off;      // elements completed in earlier passes since the last at_element reset
    int32_t el_reset;     // at_element was reset to 0 by an end-of-turn action of this launch
};

// The rarely used fields of one particle (energy bookkeeping of cavities / radiation, loss
// bookkeeping), read from the caller's SoA ONCE at kernel entry and kept in thread-local
// memory (L1-resident) for the launch: the out-of-line op bodies then never wait for HBM.
// ptau and rvv are written back at exit and at loss (pstate_store).
struct PCold {
    double beta0, gamma0, p0c, charge_ratio, ptau, rvv;
    int64_t at_turn0;        // at_turn / at_element at kernel entry
    int32_t at_element0;
    // radiation kernels: the particle's generator state (4 x u32 of the SoA), so that the
    // thousands of photon-emission calls per turn do not each wait for an HBM round trip;
    // written back by pstate_store (exit, loss)
    int32_t rng_cached;
    uint32_t rng[4];
};

// cold fields of a lane without a particle: the lane-parallel RF maps compute on every lane
__device__ __forceinline__ void pcold_benign(PCold& c) {
    c.beta0 = c.gamma0 = c.p0c = c.charge_ratio = c.rvv = 1.;
    c.ptau = 0.;
    c.at_turn0 = 0;
    c.at_element0 = 0;
    c.rng_cached = 0;
}

// Access to one particle slot: the cached cold fields through `c`, everything else in
// place in the caller's SoA.  The field index is a compile-time constant at every call
// site, so the selection below folds away.
struct PSlot {
    const xtb_particles_t* p;
    uint32_t i;              // slot index (xtb_track rejects capacities >= 2^31)
    PCold* c;
    __device__ __forceinline__ double ldg(int f) const {
        return reinterpret_cast<const double*>(p->field[f])[i];
    }
    __device__ __forceinline__ void stg(int f, double v) const {
        reinterpret_cast<double*>(p->field[f])[i] = v;
    }
    __device__ __forceinline__ int64_t ldgi(int f) const {
        return reinterpret_cast<const int64_t*>(p->field[f])[i];
    }
    __device__ __forceinline__ double ld(int f) const {
        switch (f) {
        case F_BETA0: return c->beta0;
        case F_GAMMA0: return c->gamma0;
        case F_P0C: return c->p0c;
        case F_CHARGE_RATIO: return c->charge_ratio;
        case F_PTAU: return c->ptau;
        case F_RVV: return c->rvv;
        default: return ldg(f);
        }
    }
    __device__ __forceinline__ void st(int f, double v) const {
        switch (f) {
        case F_PTAU: c->ptau = v; break;
        case F_RVV: c->rvv = v; break;
        default: stg(f, v); break;
        }
    }
    __device__ __forceinline__ int64_t ldi(int f) const {
        switch (f) {
        case F_AT_TURN: return c->at_turn0;
        case F_AT_ELEMENT: return (int64_t) c->at_element0;
        default: return ldgi(f);
        }
    }
    // fill the cache from the SoA (kernel entry)
    __device__ __forceinline__ void load_cold(const bool with_rng) const {
        c->beta0 = ldg(F_BETA0);  c->gamma0 = ldg(F_GAMMA0);  c->p0c = ldg(F_P0C);
        c->charge_ratio = ldg(F_CHARGE_RATIO);  c->ptau = ldg(F_PTAU);  c->rvv = ldg(F_RVV);
        c->at_turn0 = ldgi(F_AT_TURN);  c->at_element0 = (int32_t) ldgi(F_AT_ELEMENT);
        c->rng_cached = with_rng ? 1 : 0;
        if (with_rng) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                c->rng[j] = reinterpret_cast<const uint32_t*>(p->field[F_RNG_S1 + j])[i];
        }
    }
    __device__ __forceinline__ void sti(int f, int64_t v) const {
        reinterpret_cast<int64_t*>(p->field[f])[i] = v;
    }
    __device__ __forceinline__ uint32_t ldu(int f) const {
        if (f >= F_RNG_S1 && f <= F_RNG_S4 && c->rng_cached) return c->rng[f - F_RNG_S1];
        return reinterpret_cast<const uint32_t*>(p->field[f])[i];
    }
    __device__ __forceinline__ void stu(int f, uint32_t v) const {
        if (f >= F_RNG_S1 && f <= F_RNG_S4 && c->rng_cached) { c->rng[f - F_RNG_S1] = v;  return; }
        reinterpret_cast<uint32_t*>(p->field[f])[i] = v;
    }
    // cached generator state -> SoA
    __device__ __forceinline__ void flush_rng() const {
        if (c->rng_cached) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                reinterpret_cast<uint32_t*>(p->field[F_RNG_S1 + j])[i] = c->rng[j];
        }
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
__device__ __forceinline__ void pstate_load(PHot& P, const PSlot& G) {
    P.x = G.ld(F_X);  P.px = G.ld(F_PX);  P.y = G.ld(F_Y);  P.py = G.ld(F_PY);
    P.zeta = G.ld(F_ZETA);  P.delta = G.ld(F_DELTA);  P.rpp = G.ld(F_RPP);
    P.chi = G.ld(F_CHI);  P.s = G.ld(F_S);
    P.rv0v = 1. / G.ld(F_RVV);
}

// hot state -> full state.  The SoA still holds the launch-entry at_turn / at_element
// of a particle that has not been stored yet, and the current rvv.
__device__ __forceinline__ PState pstate_full(const PState& P, const PSlot& G, const XtbPass& ps,
                                              const uint32_t eidx) {
    PState T = P;
    T.at_turn = G.ldi(F_AT_TURN) + ps.turn_inc;
    T.at_element = (ps.el_reset ? 0 : (int32_t) G.ldi(F_AT_ELEMENT)) + (int32_t) (ps.el_off + eidx);
    return T;
}
__device__ __forceinline__ PState pstate_full(const PHot& P, const PSlot& G, const XtbPass& ps,
                                              const uint32_t eidx) {
    PState T;
    T.x = P.x;  T.px = P.px;  T.y = P.y;  T.py = P.py;  T.zeta = P.zeta;  T.delta = P.delta;
    T.rpp = P.rpp;  T.rv0v = P.rv0v;  T.chi = P.chi;  T.s = P.s;  T.state = P.state;
    T.rvv = G.ld(F_RVV);
    T.at_turn = G.ldi(F_AT_TURN) + ps.turn_inc;
    T.at_element = (ps.el_reset ? 0 : (int32_t) G.ldi(F_AT_ELEMENT)) + (int32_t) (ps.el_off + eidx);
    return T;
}
// full state -> hot state after a generic op (rvv is written through to the SoA)
__device__ __forceinline__ void pstate_back(PState& P, const PState& T, const PSlot& G) { P = T; }
__device__ __forceinline__ void pstate_back(PHot& P, const PState& T, const PSlot& G) {
    P.x = T.x;  P.px = T.px;  P.y = T.y;  P.py = T.py;  P.zeta = T.zeta;  P.delta = T.delta;
    P.rpp = T.rpp;  P.rv0v = T.rv0v;  P.chi = T.chi;  P.s = T.s;  P.state = T.state;
    G.st(F_RVV, T.rvv);
}

// Write the register-resident fields back (exit, loss, monitor snapshots read
// the rest from the SoA, where write-through keeps them current).
__device__ __forceinline__ void pstate_store(const PState& P, const PSlot& G) {
    G.st(F_X, P.x);  G.st(F_PX, P.px);  G.st(F_Y, P.y);  G.st(F_PY, P.py);
    G.st(F_ZETA, P.zeta);  G.st(F_DELTA, P.delta);  G.st(F_RPP, P.rpp);
    G.stg(F_RVV, P.rvv);  G.st(F_S, P.s);
    G.stg(F_PTAU, G.c->ptau);
    G.sti(F_AT_TURN, P.at_turn);
    G.sti(F_AT_ELEMENT, (int64_t) P.at_element);
    G.sti(F_STATE, (int64_t) P.state);
    G.flush_rng();
}
