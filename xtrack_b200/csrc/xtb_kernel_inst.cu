// xtb_kernel_inst.cu -- instantiates the tracking kernel variants.
//
// Compiled twice by xtrack_b200/build.py:
//   -DXTB_EXACT=0 -fmad=true   -> xtb_launch_track_fast   (default, FMA contraction)
//   -DXTB_EXACT=1 -fmad=false  -> xtb_launch_track_exact  (reference rounding order;
//                                 the reference CPU build has no FMA contraction)
#include <stdlib.h>
#include <string.h>
#include "xtb_kernel.cuh"

#if XTB_EXACT
#define XTB_LAUNCH_NAME xtb_launch_track_exact
#else
#define XTB_LAUNCH_NAME xtb_launch_track_fast
#endif

// particle slots per thread: 3 in the thin kernels (measured best of 1..4 on B200, see
// profiles/), 2 in the thick ones (their bodies run on both at once, xtb_thick.cuh)
#ifndef XTB_NPT_THIN
#define XTB_NPT_THIN 3
#endif
#ifndef XTB_NPT_HEAVY
#define XTB_NPT_HEAVY 2
#endif
// radiation kernels: photon emission is per particle, divergent and register-hungry; one
// particle per thread at 6 blocks / SM measured 1.3x (LEP) ... 1.5x (CLIC-DR) faster than two
#ifndef XTB_NPT_SYNRAD
#define XTB_NPT_SYNRAD 1
#endif
static_assert(XTB_NPT_SYNRAD == 1, "the photon-emission calls are compiled into the one-lane bodies only");
// small beams: particles per SM below which the thin kernel runs with 1 / 2 particles per thread
// (measured on hllhc_14 and SPS, 62 500 ... 500 000 particles on 148 SMs, profiles/r02_history.md:
// NPT 1 wins up to ~310 000 particles per GPU, NPT 2 around 375 000, NPT 3 from 500 000 on)
#ifndef XTB_NPT1_BELOW
#define XTB_NPT1_BELOW 2100
#endif
#ifndef XTB_NPT2_BELOW
#define XTB_NPT2_BELOW 2900
#endif


// ---- launch shape ------------------------------------------------------------------------
// One grid of XTB_THREADS-thread blocks, block b carrying slots [b * T * NPT, (b + 1) * T * NPT).
// Measured alternatives that LOST (profiles/r02_history.md): full waves + a balanced second
// grid of smaller blocks for the rest (the hardware scheduler already overlaps the ragged last
// wave: 10^6 particles = 4.4 waves run within 1.4 % of the 4.0-wave rate), and 32 / 64-thread
// blocks for small beams (same per-SM work, fewer warps to hide latency with).
// What a SMALL beam needs is more warps, not smaller blocks: with fewer than ~1500 particles per
// SM the NPT = 3 kernel leaves the schedulers with 2 warps each; NPT = 1 / 2 spread the same
// particles over 3x / 1.5x the warps (at the price of one op decode per 1 / 2 particles instead
// of 3).  `xtb_pick_npt` chooses by particles per SM; XTB_NPT_FORCE overrides (experiments).
static int xtb_pick_npt(const int64_t n, const int n_sm) {
    static const int forced = [] {
        const char* e = getenv("XTB_NPT_FORCE");
        return e ? atoi(e) : 0;
    }();
    if (forced >= 1 && forced <= 3) return forced;
    if (n_sm <= 0) return XTB_NPT_THIN;
    const double per_sm = (double) n / n_sm;
    if (per_sm < XTB_NPT1_BELOW) return 1;
    if (per_sm < XTB_NPT2_BELOW) return 2;
    return XTB_NPT_THIN;
}

// Block size: XTB_THREADS, halved only for beams too small to give every SM two blocks.
// XTB_THREADS_FORCE overrides (experiments).
static unsigned xtb_pick_threads(const int64_t n, const int npt, const int n_sm) {
    static const int forced = [] {
        const char* e = getenv("XTB_THREADS_FORCE");
        return e ? atoi(e) : 0;
    }();
    if (forced == 32 || forced == 64 || forced == 128) return (unsigned) forced;
    unsigned t = XTB_THREADS;
    if (n_sm <= 0) return t;
    // a beam of fewer blocks than SMs (10^4 particles: 40 - 80 blocks of 128 threads on 148 SMs)
    // leaves SMs idle: smaller blocks until every SM holds two.  (Above that, smaller blocks
    // only lose: measured 62 500 ... 500 000 particles, profiles/r02_history.md.)
    while (t > 32 && (double) n / ((double) t * npt) < 2.0 * n_sm) t >>= 1;
    return t;
}

// QUANTUM: the program contains photon-emission (radiation_flag 2) bodies.  Radiation with the
// deterministic mean model only runs like the other thick kernels (XTB_NPT_HEAVY lanes).
template <int NPT, bool HEAVY, bool SYNRAD, bool FRZ, bool BMON>
static cudaError_t launch_npt(const XtbTrackArgs& a0, int n_sm, int* n_launched, cudaStream_t stream) {
    XtbTrackArgs a = a0;
    a.slot_begin = 0;
    a.slot_end = a.part.capacity;
    const unsigned threads = xtb_pick_threads(a.part.capacity, NPT, n_sm);
    const int64_t per_block = (int64_t) threads * NPT;
    const unsigned grid = (unsigned) ((a.part.capacity + per_block - 1) / per_block);
    xtb_track_kernel<NPT, HEAVY, SYNRAD, FRZ, (XTB_EXACT != 0), BMON><<<grid, threads, 0, stream>>>(a);
    *n_launched += 1;
    return cudaGetLastError();
}

template <bool HEAVY, bool SYNRAD, bool FRZ, bool BMON = HEAVY, bool QUANTUM = true>
static cudaError_t launch(const XtbTrackArgs& a, int n_sm, int* n_launched, cudaStream_t stream) {
    if constexpr (!HEAVY && !FRZ && !BMON) {
        // the production thin kernel exists for 1, 2 and 3 particles per thread
        switch (xtb_pick_npt(a.part.capacity, n_sm)) {
        case 1: return launch_npt<1, false, false, false, false>(a, n_sm, n_launched, stream);
        case 2: return launch_npt<2, false, false, false, false>(a, n_sm, n_launched, stream);
        default: return launch_npt<XTB_NPT_THIN, false, false, false, false>(a, n_sm, n_launched, stream);
        }
    } else {
        constexpr int NPT = HEAVY ? ((SYNRAD && QUANTUM) ? XTB_NPT_SYNRAD : XTB_NPT_HEAVY) : XTB_NPT_THIN;
        return launch_npt<NPT, HEAVY, SYNRAD, FRZ, BMON>(a, n_sm, n_launched, stream);
    }
}

// variant bits: 1 = heavy ops present, 2 = synrad, 4 = freeze longitudinal,
// 8 = beam-monitor ops present (thin kernels only: the thick ones always contain them),
// 16 = photon-emission bodies present (radiation kernels: one lane per thread)
extern "C" cudaError_t XTB_LAUNCH_NAME(unsigned variant, const XtbTrackArgs* a, int n_sm,
                                       int* n_launched, cudaStream_t stream) {
    switch (variant & 7u) {
    case 0: return (variant & 8u) ? launch<false, false, false, true>(*a, n_sm, n_launched, stream)
                                  : launch<false, false, false, false>(*a, n_sm, n_launched, stream);
    case 4: return (variant & 8u) ? launch<false, false, true, true>(*a, n_sm, n_launched, stream)
                                  : launch<false, false, true, false>(*a, n_sm, n_launched, stream);
#ifdef XTB_WITH_HEAVY
    case 1: return launch<true, false, false>(*a, n_sm, n_launched, stream);
    case 5: return launch<true, false, true>(*a, n_sm, n_launched, stream);
    case 2: case 3: return (variant & 16u) ? launch<true, true, false, true, true>(*a, n_sm, n_launched, stream)
                                           : launch<true, true, false, true, false>(*a, n_sm, n_launched, stream);
    case 6: case 7: return (variant & 16u) ? launch<true, true, true, true, true>(*a, n_sm, n_launched, stream)
                                           : launch<true, true, true, true, false>(*a, n_sm, n_launched, stream);
#endif
    default: return cudaErrorNotSupported;
    }
}

#ifdef XTB_COUNT_STOPS
#if XTB_EXACT
extern "C" int xtb_debug_stops(unsigned long long* out8, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, xtb_dbg_stops, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0};
        e = cudaMemcpyToSymbol(xtb_dbg_stops, z, sizeof(z));
    }
    return (int) e;
}
#endif
#endif
