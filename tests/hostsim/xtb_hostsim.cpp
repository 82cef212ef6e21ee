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
static inline int __double2hiint(double v) { long long r; std::memcpy(&r, &v, 8); return (int) (r >> 32); }
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
using std::max;
using std::min;

#define XTB_WITH_HEAVY
#include "../../xtrack_b200/csrc/xtb_interp.cuh"

// Serial restatement of the turn loop of xtb_kernel.cuh around the SHARED op
// interpreter (xtb_interp.cuh); NPT slots are carried together as in the kernel.
template <int NPT, bool SYNRAD, bool FRZ>
static void run(const XtbTrackArgs& a) {
    for (int64_t base = 0; base < a.part.capacity; base += NPT) {
        PSlot G[NPT];
        PState P[NPT];
        bool live[NPT];
        bool any_live = false;
        for (int k = 0; k < NPT; ++k) {
            G[k].p = &a.part;
            G[k].i = base + k;
            live[k] = false;
            if (base + k < a.part.capacity) {
                P[k].state = (int32_t) G[k].ldi(F_STATE);
                live[k] = P[k].state > 0;
            }
            if (live[k]) {
                pstate_load(P[k], G[k]);
            } else {
                pstate_benign(P[k]);
                P[k].at_turn = 0;
                P[k].at_element = 0;
            }
            any_live = any_live || live[k];
        }
        if (!any_live) continue;
        for (int turn = 0; turn < a.num_turns; ++turn) {
            any_live = false;
            for (int k = 0; k < NPT; ++k) any_live = any_live || live[k];
            if (!any_live) break;
            if (a.flag_monitor == 1)
                for (int k = 0; k < NPT; ++k)
                    if (live[k]) monitor_record(a.mon, P[k], G[k]);
            uint32_t eidx = 0;
            xtb_interp<NPT, true, SYNRAD, FRZ>(a.prog + a.pc_start, a.prog + a.pc_stop, P, G, live,
                                               eidx, a);
            if (a.flag_monitor == 2)
                for (int k = 0; k < NPT; ++k)
                    if (live[k]) {
                        PState T = P[k];
                        T.at_element += (int32_t) eidx;
                        monitor_record(a.mon, T, G[k]);
                    }
            for (int k = 0; k < NPT; ++k) {
                if (a.flag_end_turn_actions > 0) {
                    P[k].at_turn += 1;
                    P[k].at_element = 0;
                    if (a.flag_reset_s > 0 && !FRZ) P[k].s = 0.;
                } else {
                    P[k].at_element += (int32_t) eidx;
                }
            }
        }
        for (int k = 0; k < NPT; ++k)
            if (live[k]) pstate_store(P[k], G[k]);
    }
}

extern "C" int xtb_hostsim_track(const uint64_t* words, const uint32_t* elem_offset,
                                  const xtb_particles_t* p, int64_t num_turns, int32_t ele_start,
                                  int32_t num_ele_track, int32_t flag_end_turn_actions,
                                  int32_t flag_reset_s, int32_t flag_monitor, const xtb_monitor_t* mon,
                                  uint64_t track_flags, double global_xy_limit, uint32_t variant,
                                  double line_length, const xtb_monitor_t* inline_mon,
                                  const xtb_last_turns_monitor_t* inline_ltm, int32_t npt) {
    XtbTrackArgs a;
    std::memset(&a, 0, sizeof(a));
    a.prog = words;
    a.part = *p;
    if (mon) a.mon = *mon;
    a.inline_mon = inline_mon;
    a.inline_ltm = inline_ltm;
    a.pc_start = elem_offset[ele_start];
    a.pc_stop = elem_offset[ele_start + num_ele_track];
    if (a.pc_start == XTB_NOT_ADDRESSABLE || a.pc_stop == XTB_NOT_ADDRESSABLE) return -1;
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
    if (npt == 2) {
        if (synrad && frz) run<2, true, true>(a);
        else if (synrad) run<2, true, false>(a);
        else if (frz) run<2, false, true>(a);
        else run<2, false, false>(a);
    } else {
        if (synrad && frz) run<1, true, true>(a);
        else if (synrad) run<1, true, false>(a);
        else if (frz) run<1, false, true>(a);
        else run<1, false, false>(a);
    }
    return 0;
}
