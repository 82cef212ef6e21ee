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

// ---- launch shape ------------------------------------------------------------------------
// A block carries T * NPT particles through all turns, so the unit of scheduling is coarse:
// B blocks of 128 threads are resident per SM (B * 128 * NPT particles in flight), and a beam
// of w.f "waves" of them takes ceil(w.f) wave times -- 10^6 particles on 148 SMs are 4.4
// waves (the last one 40 % full), 125 000 particles (one eighth of the 10^6-particle
// dynamic-aperture beam on each of 8 GPUs) are 326 blocks for 592 slots: 2 or 3 per SM.
// The FP64 pipe does not care how many threads feed it, only that all SMs hold EQUAL work:
//   * the full waves go out as one grid of 128-thread blocks;
//   * the rest (everything, for a small beam) goes out as a second grid on the same stream
//     whose block size T in {128, 64, 32} is chosen so that the blocks spread evenly over
//     the SMs -- T minimises ceil(blocks / n_sm) * T, the largest number of threads any SM
//     ends up with; all of them are resident at once (registers allow B * 128 / T blocks
//     per SM, the two tile buffers of a block 13).  It runs as long as its fill, not as
//     long as a full wave.
// XTB_LAUNCH_SHAPE=legacy in the environment restores the single 128-thread grid (A/B runs).
struct XtbGridPlan {
    int64_t n_main;          // slots in the grid of full waves (0: none)
    unsigned grid_main;
    int64_t n_rest;
    unsigned grid_rest, threads_rest;
};

static XtbGridPlan plan_grids(const int64_t n, const int npt, const int blocks_per_sm, const int n_sm) {
    XtbGridPlan g = {0, 0, 0, 0, XTB_THREADS};
    const int64_t per_block = (int64_t) XTB_THREADS * npt;
    static const bool legacy = [] {
        const char* e = getenv("XTB_LAUNCH_SHAPE");
        return e && !strcmp(e, "legacy");
    }();
    if (legacy || n_sm <= 0) {
        g.n_main = n;
        g.grid_main = (unsigned) ((n + per_block - 1) / per_block);
        return g;
    }
    const int64_t wave = (int64_t) n_sm * blocks_per_sm * per_block;
    const int64_t full = n / wave;
    g.n_main = full * wave;
    g.grid_main = (unsigned) (full * n_sm * blocks_per_sm);
    g.n_rest = n - g.n_main;
    if (g.n_rest > 0) {
        int64_t best = -1;
        for (unsigned t = XTB_THREADS; t >= 32; t >>= 1) {
            const int64_t blocks = (g.n_rest + (int64_t) t * npt - 1) / ((int64_t) t * npt);
            const int64_t resident = min((int64_t) blocks_per_sm * XTB_THREADS / t, (int64_t) 13);
            if (blocks > resident * n_sm) continue;             // would not fit one wave
            const int64_t load = ((blocks + n_sm - 1) / n_sm) * t;
            if (best < 0 || load < best) { best = load;  g.threads_rest = t; }
        }
        g.grid_rest = (unsigned) ((g.n_rest + (int64_t) g.threads_rest * npt - 1)
                                  / ((int64_t) g.threads_rest * npt));
    }
    return g;
}

// QUANTUM: the program contains photon-emission (radiation_flag 2) bodies.  Radiation with the
// deterministic mean model only runs like the other thick kernels (XTB_NPT_HEAVY lanes).
template <bool HEAVY, bool SYNRAD, bool FRZ, bool BMON = HEAVY, bool QUANTUM = true>
static cudaError_t launch(const XtbTrackArgs& a0, int n_sm, int* n_launched, cudaStream_t stream) {
    constexpr int NPT = HEAVY ? ((SYNRAD && QUANTUM) ? XTB_NPT_SYNRAD : XTB_NPT_HEAVY) : XTB_NPT_THIN;
    constexpr int BPS = HEAVY ? ((SYNRAD && NPT == 1) ? XTB_SYNRAD_BLOCKS_PER_SM : XTB_HEAVY_BLOCKS_PER_SM)
                              : XTB_THIN_BLOCKS_PER_SM;
    const XtbGridPlan g = plan_grids(a0.part.capacity, NPT, BPS, n_sm);
    XtbTrackArgs a = a0;
    if (g.grid_main) {
        a.slot_begin = 0;
        a.slot_end = g.n_main;
        xtb_track_kernel<NPT, HEAVY, SYNRAD, FRZ, (XTB_EXACT != 0), BMON><<<g.grid_main, XTB_THREADS, 0, stream>>>(a);
        *n_launched += 1;
    }
    if (g.grid_rest) {
        a.slot_begin = g.n_main;
        a.slot_end = g.n_main + g.n_rest;
        xtb_track_kernel<NPT, HEAVY, SYNRAD, FRZ, (XTB_EXACT != 0), BMON><<<g.grid_rest, g.threads_rest, 0, stream>>>(a);
        *n_launched += 1;
    }
    return cudaGetLastError();
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
