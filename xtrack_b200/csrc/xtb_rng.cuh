// xtb_rng.cuh -- the per-particle random generators of the radiation kernels.
#pragma once
#include <stdint.h>

// rng_get_int32 / rng_get, rng_src/base_rng.h:23-42 (state in the SoA, as in the reference)
struct Rng {
    uint32_t s1, s2, s3, s4;
};
#define XTB_TAUSW(s, a, b, c, d) ((((s) & (c)) << (d)) ^ ((((s) << (a)) ^ (s)) >> (b)))
__device__ __forceinline__ uint32_t rng_u32(Rng& r) {
    r.s1 = XTB_TAUSW(r.s1, 13, 19, 4294967294u, 12);
    r.s2 = XTB_TAUSW(r.s2, 2, 25, 4294967288u, 4);
    r.s3 = XTB_TAUSW(r.s3, 3, 11, 4294967280u, 17);
    r.s4 = 1664525u * r.s4 + 1013904223u;
    return r.s1 ^ r.s2 ^ r.s3 ^ r.s4;
}

// Counter-based generator for production runs (north star: "a counter-based RNG for
// radiation"): Philox4x32-10 (Salmon et al., SC'11).  The particle's four state words hold the
// KEY (s1, s2: seed and particle id) and a 64-bit COUNTER of the draws made so far (s3 low,
// s4 high); draw number n is word (n & 3) of Philox(counter = n >> 2, key).  The stream of a
// particle depends on its key and on nothing else -- not on the slot, the block, the GPU it is
// tracked on or how a run is cut into launches -- and can be continued from a checkpoint of
// the SoA.  The Tausworthe generator above stays the parity mode (the reference's streams).
struct Philox {
    uint32_t blk[4];     // the block the next draws come from (valid iff have)
    bool have;
};
__device__ __forceinline__ void philox4x32_10(const uint32_t k0_, const uint32_t k1_, const uint32_t c0_,
                                              const uint32_t c1_, uint32_t (&out)[4]) {
    uint32_t c0 = c0_, c1 = c1_, c2 = 0u, c3 = 0u, k0 = k0_, k1 = k1_;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t) p1;
        c3 = (uint32_t) p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;  out[1] = c1;  out[2] = c2;  out[3] = c3;
}

