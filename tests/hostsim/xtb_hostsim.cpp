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
#include <type_traits>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __global__
static inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
static inline double __hiloint2double(int hi, int lo) { long long r = ((long long) hi << 32) | (unsigned) lo; double d; std::memcpy(&d, &r, 8); return d; }
static inline int __double2hiint(double v) { long long r; std::memcpy(&r, &v, 8); return (int) (r >> 32); }
#define __reduce_max_sync(m, v) (v)
#define __any_sync(m, v) (v)
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
using std::max;
using std::min;

#define XTB_WITH_HEAVY
#define XTB_COUNT_TRIG_MISS
#define XTB_COUNT_STOPS
#include "../../xtrack_b200/csrc/xtb_interp.cuh"

// (test hook) lookups / misses of the host-tabulated element trigonometry (xtb_thick.cuh)
extern "C" void xtb_hostsim_trig_stats(long long* lookups, long long* misses, int reset) {
    *lookups = xtb_trig_lookups;
    *misses = xtb_trig_misses;
    if (reset) { xtb_trig_lookups = 0;  xtb_trig_misses = 0; }
}

// Serial restatement of the turn loop of xtb_kernel.cuh around the SHARED op
// interpreter (xtb_interp.cuh); NPT slots are carried together as in the kernel.
// `a.prog` is the image of the element range: its ops + the XTB_OP_END sentinel.
template <int NPT, bool SYNRAD, bool FRZ, class S>
static void run(const XtbTrackArgs& a) {
    for (int64_t base = 0; base < a.part.capacity; base += NPT) {
        XtbLanes<NPT, S> lanes;
        PSlot G[NPT];
        S (&P)[NPT] = lanes.P;
        bool (&live)[NPT] = lanes.live;
        bool any_live = false, chi_one = true;
        for (int k = 0; k < NPT; ++k) {
            G[k].p = &a.part;
            G[k].i = (uint32_t) (base + k);
            G[k].c = &lanes.C[k];
            lanes.slot[k] = (uint32_t) (base + k);
            live[k] = (base + k < a.part.capacity) && G[k].ldi(F_STATE) > 0;
            if (live[k]) {
                G[k].load_cold(SYNRAD);
                pstate_load(P[k], G[k]);
                P[k].state = 1;
                chi_one = chi_one && (P[k].chi == 1.0);
            } else {
                pstate_benign(P[k]);
                pcold_benign(lanes.C[k]);
            }
            any_live = any_live || live[k];
        }
        if (!any_live) continue;
        // the kernel's fast state: chi == 1 and one common s on the whole group
        bool fast_state = chi_one && std::is_same<S, PHot>::value;
        {
            int first_live = -1;
            for (int k = NPT - 1; k >= 0; --k) if (live[k]) first_live = k;
            const double s_ref = P[first_live].s;
            for (int k = 0; k < NPT; ++k)
                if (live[k] && __double_as_longlong(P[k].s) != __double_as_longlong(s_ref)) fast_state = false;
            if (fast_state) for (int k = 0; k < NPT; ++k) P[k].s = s_ref;
        }
        XtbPass ps;
        ps.turn_inc = 0;  ps.el_off = 0;  ps.el_reset = 0;
        for (int turn = 0; turn < a.num_turns; ++turn) {
            any_live = false;
            for (int k = 0; k < NPT; ++k) any_live = any_live || live[k];
            if (!any_live) break;
            if (a.flag_monitor == 1)
                for (int k = 0; k < NPT; ++k)
                    if (live[k]) {
                        const PState T = pstate_full(P[k], G[k], ps, 0u);
                        monitor_record(a.mon, T, G[k]);
                    }
            lanes.eidx = 0;
            lanes.off = 0;
            if (fast_state) xtb_run_tile<NPT, true, SYNRAD, FRZ, true, true, true>(a.prog, lanes, ps, a);
            else xtb_run_tile<NPT, true, SYNRAD, FRZ, false, false, true>(a.prog, lanes, ps, a);
            const uint32_t eidx = a.num_ele_track;
            if (a.flag_monitor == 2)
                for (int k = 0; k < NPT; ++k)
                    if (live[k]) {
                        const PState T = pstate_full(P[k], G[k], ps, eidx);
                        monitor_record(a.mon, T, G[k]);
                    }
            xtb_end_pass<NPT, FRZ>(P, ps, eidx, a);
        }
        for (int k = 0; k < NPT; ++k)
            if (live[k]) {
                const PState T = pstate_full(P[k], G[k], ps, 0u);
                pstate_store(T, G[k]);
            }
    }
}

// (test hook) one block of the counter-based generator, as the device code computes it
extern "C" void xtb_hostsim_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* out) {
    uint32_t b[4];
    philox4x32_10(k0, k1, c0, c1, b);
    for (int j = 0; j < 4; ++j) out[j] = b[j];
}

// (test hook) inverse-CDF tables of the quantum-kick model for the following track calls
static const double* g_synrad_tables = nullptr;
extern "C" void xtb_hostsim_set_synrad_tables(const double* blob) { g_synrad_tables = blob; }

extern "C" int xtb_hostsim_track(const uint64_t* words, const uint32_t* elem_offset,
                                  const xtb_particles_t* p, int64_t num_turns, int32_t ele_start,
                                  int32_t num_ele_track, int32_t flag_end_turn_actions,
                                  int32_t flag_reset_s, int32_t flag_monitor, const xtb_monitor_t* mon,
                                  uint64_t track_flags, double global_xy_limit, uint32_t variant,
                                  double line_length, const xtb_monitor_t* inline_mon,
                                  const xtb_last_turns_monitor_t* inline_ltm, int32_t npt, int32_t force_full) {
    int npt_heavy = (npt >> 8) ? (npt >> 8) : 2;
    // XTB_NPT_SYNRAD of xtb_kernel_inst.cu: one lane per thread when photon-emission bodies exist
    if (variant & XTB_VARIANT_SYNRAD) {
        const uint64_t* w0 = words + elem_offset[ele_start];
        const uint64_t* w1 = words + elem_offset[ele_start + num_ele_track];
        for (const uint64_t* pw = w0; pw < w1; pw += (*pw >> 16) & 0xffffu)
            if ((*pw & 0xffu) == XTB_OP_MAGNET_BODY && (((uint32_t) (*pw >> 32) >> 10) & 3u) >= 2u) npt_heavy = 1;
    }
    npt &= 0xff;
    XtbTrackArgs a;
    std::memset(&a, 0, sizeof(a));
    if (elem_offset[ele_start] == XTB_NOT_ADDRESSABLE
        || elem_offset[ele_start + num_ele_track] == XTB_NOT_ADDRESSABLE) return -1;
    // image of the element range: its ops, the END sentinel, slack for the prefetches
    std::vector<uint64_t> image(words + elem_offset[ele_start],
                                words + elem_offset[ele_start + num_ele_track]);
    image.push_back(XTB_HDR(XTB_OP_END, 0, 2, 0));
    image.resize(image.size() + 9, 0);
    a.prog = image.data();
    a.num_ele_track = (uint32_t) num_ele_track;
    a.part = *p;
    if (mon) a.mon = *mon;
    a.inline_mon = inline_mon;
    a.inline_ltm = inline_ltm;
    a.num_turns = (int32_t) num_turns;
    a.flag_end_turn_actions = flag_end_turn_actions;
    a.flag_reset_s = flag_reset_s;
    a.flag_monitor = flag_monitor;
    a.ignore_global = (int32_t) ((track_flags >> XTB_FLAG_IGNORE_GLOBAL_APERTURE) & 1);
    a.ignore_local = (int32_t) ((track_flags >> XTB_FLAG_IGNORE_LOCAL_APERTURE) & 1);
    a.kill_cavity_kick = (int32_t) ((track_flags >> XTB_FLAG_KILL_CAVITY_KICK) & 1);
    a.rng_philox = (variant & XTB_VARIANT_PHILOX) ? 1 : 0;
    a.aperture_prefilter = 0;
    for (size_t pc = 0; pc + 2 < image.size();) {            // as xtb_api.cu::program_prepare
        const uint32_t hx = (uint32_t) image[pc], opx = hx & 0xffu;
        if (opx == XTB_OP_END) break;
        if (opx < XTB_GENERIC_FIRST && ((opx & ~(uint32_t) XTB_OPBIT_DRIFT) == XTB_OP_RECT || (opx & ~(uint32_t) XTB_OPBIT_DRIFT) == XTB_OP_ELLIPSE))
            a.aperture_prefilter = 1;
        pc += hx >> 16;
    }
    a.synrad_tables = g_synrad_tables;
    a.line_length = line_length;
    a.global_xy_limit = global_xy_limit;
    const bool synrad = variant & XTB_VARIANT_SYNRAD, frz = variant & XTB_VARIANT_FREEZE_LONG;
    // has_heavy mirrors the kernel choice: thick programs run on the full state
    bool heavy = synrad;
    for (size_t pc = 0; pc + 2 < image.size() && !heavy;) {
        const uint32_t hx = (uint32_t) image[pc];
        if ((hx & 0xffu) == XTB_OP_END) break;
        if ((hx & 0xffu) >= XTB_HEAVY_FIRST) heavy = true;
        pc += hx >> 16;
    }
    if (heavy || force_full) {
        // (npt_heavy mirrors XTB_NPT_HEAVY of xtb_kernel_inst.cu)
        if (npt_heavy == 2) {
            if (synrad && frz) run<2, true, true, PState>(a);
            else if (synrad) run<2, true, false, PState>(a);
            else if (frz) run<2, false, true, PState>(a);
            else run<2, false, false, PState>(a);
        } else {
            if (synrad && frz) run<1, true, true, PState>(a);
            else if (synrad) run<1, true, false, PState>(a);
            else if (frz) run<1, false, true, PState>(a);
            else run<1, false, false, PState>(a);
        }
    } else if (npt == 3) {
        if (frz) run<3, false, true, PHot>(a);
        else run<3, false, false, PHot>(a);
    } else if (npt == 2) {
        if (frz) run<2, false, true, PHot>(a);
        else run<2, false, false, PHot>(a);
    } else {
        if (frz) run<1, false, true, PHot>(a);
        else run<1, false, false, PHot>(a);
    }
    return 0;
}

// (test hook) returns of xtb_run_fast by reason (XtbStop) since the last reset
extern "C" void xtb_hostsim_stop_counts(unsigned long long* out8, int reset) {
    for (int i = 0; i < 8; ++i) { out8[i] = xtb_dbg_stops[i];  if (reset) xtb_dbg_stops[i] = 0; }
}
