// xtb_kernel.cuh -- the fused turn x element tracking kernel for sm_100a.
//
// Replaces the generated `track_line` kernel of xtrack/tracker.py:545-749.
// Design:
//   * one thread carries NPT particle slots; their coordinates live in FP64
//     registers over the whole launch (PState), SoA traffic only at entry / exit /
//     loss / monitor records.  NPT = 3 in the thin kernels, 2 in the thick ones, 1 in the
//     radiation kernels (xtb_kernel_inst.cu): one op decode and one set of shared-memory
//     parameter loads serve all of them, and their independent dependency chains keep the
//     FP64 pipe fed;
//   * the lowered lattice is streamed through shared memory in tiles with
//     1-D bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP), double
//     buffered; all lanes read the same op -> shared-memory broadcast;
//   * a lost particle is written back the moment it is lost and its lane
//     continues on a benign on-axis state (no per-op predication); a warp whose
//     lanes are all done skips whole tiles (warp vote), a block stops at the
//     next turn boundary;
//   * template switches select the kernel variant: HEAVY (thick-magnet ops
//     compiled in), SYNRAD, FRZ (freeze_longitudinal), BMON (beam-monitor ops compiled
//     in: always in the thick kernels, on demand in the thin ones).  The translation unit is
//     compiled twice, with and without FMA contraction (xtb_kernel_inst.cu).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include "xtb_interp.cuh"

// Launch shape (measured on B200, profiles/r01_history.md): 128 threads, 4 blocks per SM for
// the thin kernels (128 registers per thread, 16 warps per SM; smaller blocks desynchronise
// the warps of an SM sub-partition a little better than 256 x 2), and 4 blocks per SM for
// the thick kernels too: they are bound by dependent division / square-root chains, and 16
// warps per SM at 128 registers (a few spills) beat 8 warps at 246 registers by 25 %.
#ifndef XTB_THREADS
#define XTB_THREADS 128
#endif
#ifndef XTB_THIN_BLOCKS_PER_SM
#define XTB_THIN_BLOCKS_PER_SM 4
#endif
#ifndef XTB_HEAVY_BLOCKS_PER_SM
#define XTB_HEAVY_BLOCKS_PER_SM 4
#endif
#ifndef XTB_SYNRAD_BLOCKS_PER_SM
#define XTB_SYNRAD_BLOCKS_PER_SM 6
#endif

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

#define XTB_TILE_BUF_WORDS (XTB_TILE_WORDS + 8)   // tile image + END sentinel + prefetch slack

// EXACT only distinguishes the symbols of the two builds of this translation unit
// (template instantiations are COMDAT: identical names would be merged at link time).
// resident blocks per SM the register allocation aims at.  Thin kernels with fewer particles
// per thread (the small-beam variants, xtb_kernel_inst.cu::xtb_pick_npt) need fewer registers
// and exist to put MORE warps on an SM: 6 blocks (85 registers) at NPT 2, 8 (64) at NPT 1.
template <int NPT, bool HEAVY, bool SYNRAD>
constexpr int xtb_blocks_per_sm() {
    return HEAVY ? ((SYNRAD && NPT == 1) ? XTB_SYNRAD_BLOCKS_PER_SM : XTB_HEAVY_BLOCKS_PER_SM)
                 : (NPT >= 3 ? XTB_THIN_BLOCKS_PER_SM : (NPT == 2 ? 6 : 8));
}

template <int NPT, bool HEAVY, bool SYNRAD, bool FRZ, bool EXACT, bool BMON>
__global__ void __launch_bounds__(XTB_THREADS, xtb_blocks_per_sm<NPT, HEAVY, SYNRAD>())
xtb_track_kernel(const __grid_constant__ XtbTrackArgs a) {
    using S = typename std::conditional<HEAVY, PState, PHot>::type;
    __shared__ __align__(128) uint64_t tile[XTB_NUM_BUF][XTB_TILE_BUF_WORDS];
    __shared__ __align__(8) uint64_t full_bar[XTB_NUM_BUF];

    XtbLanes<NPT, S> lanes;      // home of the particles between tiles (thread-local memory)
    PSlot G[NPT];
    S (&P)[NPT] = lanes.P;
    bool (&live)[NPT] = lanes.live;
    bool any_live = false;
    bool chi_one = true;
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        // (the block size is a launch parameter: XTB_THREADS or a smaller power of two)
        const int64_t slot = a.slot_begin + ((int64_t) blockIdx.x * NPT + k) * blockDim.x + threadIdx.x;
        G[k].p = &a.part;
        G[k].i = (uint32_t) slot;
        G[k].c = &lanes.C[k];
        lanes.slot[k] = (uint32_t) slot;
        live[k] = false;
        if (slot < a.slot_end) {
            // check_is_active (GPU), local_particle_custom_api.h:188
            live[k] = G[k].ldi(F_STATE) > 0;
        }
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
    if (!__syncthreads_or(any_live)) return;     // nothing to track in this block
    // chi == 1 for every particle of the block (any beam of one species): the products
    // chi * coefficient are then exact copies and are left out (CHI1 code path)
    const bool chi1 = __syncthreads_and(chi_one) && !HEAVY;
    // s bitwise identical on every live particle of the block (any beam tracked from one
    // place in the ring): it then stays identical, and the hot loop carries it once per
    // thread (SUNI code path, taken together with CHI1)
    bool fast_state = false;
    if (!HEAVY) {
        __shared__ int s_first_tid;
        __shared__ double s_ref_sh;
        if (threadIdx.x == 0) s_first_tid = (int) blockDim.x;
        __syncthreads();
        int first_live = -1;
#pragma unroll
        for (int k = NPT - 1; k >= 0; --k) if (live[k]) first_live = k;
        if (first_live >= 0) atomicMin(&s_first_tid, (int) threadIdx.x);
        __syncthreads();
        if ((int) threadIdx.x == s_first_tid) s_ref_sh = P[first_live].s;
        __syncthreads();
        const double s_ref = s_ref_sh;
        bool s_same = true;
#pragma unroll
        for (int k = 0; k < NPT; ++k)
            if (live[k]) s_same = s_same && (__double_as_longlong(P[k].s) == __double_as_longlong(s_ref));
        fast_state = __syncthreads_and(s_same) && chi1;
        if (fast_state) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) P[k].s = s_ref;      // (lanes without a particle too)
        }
    }

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

    XtbPass ps;
    ps.turn_inc = 0;
    ps.el_off = 0;
    ps.el_reset = 0;
    int64_t step = 0;
    for (int turn = 0; turn < a.num_turns; ++turn) {
        // every particle of this block lost: drain the copies in flight and stop
        if (turn > 0) {
            any_live = false;
#pragma unroll
            for (int k = 0; k < NPT; ++k) any_live = any_live || live[k];
            if (!__syncthreads_or(any_live)) {
                if (!resident) {
                    for (int64_t st = step; st < step + XTB_NUM_BUF && st < total_steps; ++st) {
                        mbar_wait(&full_bar[st % XTB_NUM_BUF], (uint32_t) ((st / XTB_NUM_BUF) & 1));
                    }
                }
                break;
            }
        }
        // xtrack/tracker.py:639-641: the turn-by-turn record precedes the turn
        if (a.flag_monitor == 1) {
#pragma unroll
            for (int k = 0; k < NPT; ++k)
                if (live[k]) {
                    const PState T = pstate_full(P[k], G[k], ps, 0u);
                    monitor_record(a.mon, T, G[k]);
                }
        }

        uint32_t eidx = 0;      // elements completed in this pass (uniform over the block)
        for (int j = 0; j < n_tiles; ++j, ++step) {
            const int kt = a.tile_first + j;
            const int b = resident ? 0 : (int) (step % XTB_NUM_BUF);
            const uint32_t w0 = a.tile_off[kt], w1 = a.tile_off[kt + 1];
            const uint32_t lo = max(w0, a.pc_start), hi = min(w1 - 2u, a.pc_stop);
            if (!resident || turn == 0) {
                mbar_wait(&full_bar[b], (uint32_t) ((step / XTB_NUM_BUF) & 1));
                if (hi < w1 - 2u) {
                    // the range stops inside this tile: plant the sentinel there
                    if (threadIdx.x == 0) {
                        tile[b][hi - w0] = XTB_HDR(XTB_OP_END, 0, 2, 0);
                        tile[b][hi - w0 + 1] = 0;
                    }
                    __syncthreads();
                }
            }
            any_live = false;
#pragma unroll
            for (int k = 0; k < NPT; ++k) any_live = any_live || live[k];
            if (__any_sync(0xffffffffu, any_live)) {
                lanes.eidx = eidx;
                lanes.off = lo - w0;
                if (fast_state)
                    xtb_run_tile<NPT, HEAVY, SYNRAD, FRZ, !HEAVY, !HEAVY, BMON>(xtb_tile_of(&tile[b][0]), lanes, ps, a);
                else
                    xtb_run_tile<NPT, HEAVY, SYNRAD, FRZ, false, false, BMON>(xtb_tile_of(&tile[b][0]), lanes, ps, a);
                eidx = lanes.eidx;
            }
            if (!resident) {
                __syncthreads();       // every warp is done with buffer b
                if (threadIdx.x == 0 && step + XTB_NUM_BUF < total_steps) {
                    // buffer b was read (and, for a range that stops inside the tile, written)
                    // through the generic proxy; the bulk copy writes it through the async proxy
#ifndef XTB_NO_PROXY_FENCE
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
                    issue(step + XTB_NUM_BUF);
                }
            }
        }
        eidx = a.num_ele_track;        // (warps that skipped dead tiles did not count)
        if (a.flag_monitor == 2) {
#pragma unroll
            for (int k = 0; k < NPT; ++k)
                if (live[k]) {
                    const PState T = pstate_full(P[k], G[k], ps, eidx);
                    monitor_record(a.mon, T, G[k]);
                }
        }
        xtb_end_pass<NPT, FRZ>(P, ps, eidx, a);
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k)
        if (live[k]) {
            const PState T = pstate_full(P[k], G[k], ps, 0u);
            pstate_store(T, G[k]);
        }
}
