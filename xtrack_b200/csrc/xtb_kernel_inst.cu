// xtb_kernel_inst.cu -- instantiates the tracking kernel variants.
//
// Compiled twice by xtrack_b200/build.py:
//   -DXTB_EXACT=0 -fmad=true   -> xtb_launch_track_fast   (default, FMA contraction)
//   -DXTB_EXACT=1 -fmad=false  -> xtb_launch_track_exact  (reference rounding order;
//                                 the reference CPU build has no FMA contraction)
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

// QUANTUM: the program contains photon-emission (radiation_flag 2) bodies.  Radiation with the
// deterministic mean model only runs like the other thick kernels (XTB_NPT_HEAVY lanes).
template <bool HEAVY, bool SYNRAD, bool FRZ, bool BMON = HEAVY, bool QUANTUM = true>
static cudaError_t launch(const XtbTrackArgs& a, cudaStream_t stream) {
    constexpr int NPT = HEAVY ? ((SYNRAD && QUANTUM) ? XTB_NPT_SYNRAD : XTB_NPT_HEAVY) : XTB_NPT_THIN;
    const int64_t per_block = (int64_t) XTB_THREADS * NPT;
    const unsigned grid = (unsigned) ((a.part.capacity + per_block - 1) / per_block);
    xtb_track_kernel<NPT, HEAVY, SYNRAD, FRZ, (XTB_EXACT != 0), BMON><<<grid, XTB_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

// variant bits: 1 = heavy ops present, 2 = synrad, 4 = freeze longitudinal,
// 8 = beam-monitor ops present (thin kernels only: the thick ones always contain them),
// 16 = photon-emission bodies present (radiation kernels: one lane per thread)
extern "C" cudaError_t XTB_LAUNCH_NAME(unsigned variant, const XtbTrackArgs* a,
                                       cudaStream_t stream) {
    switch (variant & 7u) {
    case 0: return (variant & 8u) ? launch<false, false, false, true>(*a, stream)
                                  : launch<false, false, false, false>(*a, stream);
    case 4: return (variant & 8u) ? launch<false, false, true, true>(*a, stream)
                                  : launch<false, false, true, false>(*a, stream);
#ifdef XTB_WITH_HEAVY
    case 1: return launch<true, false, false>(*a, stream);
    case 5: return launch<true, false, true>(*a, stream);
    case 2: case 3: return (variant & 16u) ? launch<true, true, false, true, true>(*a, stream)
                                           : launch<true, true, false, true, false>(*a, stream);
    case 6: case 7: return (variant & 16u) ? launch<true, true, true, true, true>(*a, stream)
                                           : launch<true, true, true, true, false>(*a, stream);
#endif
    default: return cudaErrorNotSupported;
    }
}
