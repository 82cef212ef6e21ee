// xtb_libm.cuh -- sin / cos that return the bits of glibc's (x86-64, FMA-capable CPU).
//
// The reference's CPU contexts evaluate the cavity / RF-multipole phases with the C library's
// sin and cos.  CUDA's sin / cos are as accurate, but differ from glibc's in the last bit on a
// few percent of the arguments, and a ring amplifies that last bit to ~1e-11 of the beam size
// in ten turns: without the same libm no implementation can meet the 1e-12 parity bar on a
// lattice with RF (DESIGN.md "Parity").  These are therefore restatements of glibc's own
// algorithm -- the IBM Accurate Mathematical Library routines of
// sysdeps/ieee754/dbl-64/s_sin.c (glibc 2.28 ... 2.39: `__sin`, `__cos`, `do_sin`, `do_cos`,
// `TAYLOR_SIN`, `reduce_sincos`, table `__sincostab`) -- with every multiply-add fused exactly
// where the library's FMA build fuses it (x86-64 dispatches sin / cos to that build on any CPU
// with FMA + AVX2, which includes every host of a B200).  scripts/glibc/check_libm.c compares
// them with the installed libm on 10^9 arguments: no difference.  Arguments of 105414350 and
// above (glibc: Payne-Hanek reduction `__branred`) and non-finite ones go to the CUDA library
// function: no RF phase of a tracked particle is anywhere near.
//
// Shared by the CUDA kernels and the host build of the device code (tests/hostsim).
#pragma once
#include <math.h>
#include <stdint.h>
#include "xtb_sincostab.h"

#ifdef __CUDA_ARCH__
#define XTB_LIBM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define XTB_LIBM_TABLE xtb_sincostab_dev
#else
#define XTB_LIBM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define XTB_LIBM_TABLE xtb_sincostab_host
#endif
#ifdef __CUDACC__
static __device__ const double xtb_sincostab_dev[XTB_SINCOSTAB_N] = {XTB_SINCOSTAB_VALUES};
#define XTB_LIBM_FN static __host__ __device__ __forceinline__
#define XTB_LIBM_ENTRY static __host__ __device__ __noinline__
#else
#define XTB_LIBM_FN static inline
#define XTB_LIBM_ENTRY static inline
#endif
static const double xtb_sincostab_host[XTB_SINCOSTAB_N] = {XTB_SINCOSTAB_VALUES};

namespace xtb_libm {

// (hexadecimal literals: the library's constants to the bit -- usncs.h, s_sin.c)
constexpr double big = 0x1.8p45;                   // 1.5 * 2^45: x + big rounds x to 1/128
constexpr double toint = 0x1.8p52;
constexpr double hpinv = 0x1.45f306dc9c883p-1;     // 2 / pi
constexpr double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde973c000000p-27;
constexpr double pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;
constexpr double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;   // pi / 2
constexpr double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
constexpr double cs2 = 0.5, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7,
                 s3 = -0x1.a01a019db08b8p-13, s4 = 0x1.71de27b9a7ed9p-19,
                 s5 = -0x1.addffc2fcdf59p-26;

XTB_LIBM_FN uint32_t lo_word(const double v) {
#ifdef __CUDA_ARCH__
    return (uint32_t) __double2loint(v);
#else
    uint64_t u;
    __builtin_memcpy(&u, &v, 8);
    return (uint32_t) u;
#endif
}
XTB_LIBM_FN uint32_t hi_word(const double v) {
#ifdef __CUDA_ARCH__
    return (uint32_t) __double2hiint(v);
#else
    uint64_t u;
    __builtin_memcpy(&u, &v, 8);
    return (uint32_t) (u >> 32);
#endif
}

// TAYLOR_SIN (|x| < 0.126): x + ((POLY(xx) * x - 0.5 * dx) * xx + dx)
XTB_LIBM_FN double taylor_sin(const double xx, const double x, const double dx) {
    double p = s5;
    p = XTB_LIBM_FMA(xx, p, s4);
    p = XTB_LIBM_FMA(xx, p, s3);
    p = XTB_LIBM_FMA(xx, p, s2);
    p = XTB_LIBM_FMA(xx, p, s1);
    const double t0 = XTB_LIBM_FMA(p, x, -(0.5 * dx));
    const double t = XTB_LIBM_FMA(xx, t0, dx);
    return x + t;
}

// do_sin: sin(x + dx) for 0.126 <= |x| < 0.8555, table + degree-5 / degree-6 corrections
XTB_LIBM_FN double do_sin(const double x0, double dx) {
    if (fabs(x0) < 0.126) return taylor_sin(x0 * x0, x0, dx);
    if (x0 <= 0) dx = -dx;
    const double ax = fabs(x0);
    const double u = ax + big;
    const double x = ax - (u - big);
    const uint32_t k = lo_word(u) << 2;
    const double xx = x * x;
    const double s = x + XTB_LIBM_FMA(x * xx, XTB_LIBM_FMA(xx, sn5, sn3), dx);
    const double c = XTB_LIBM_FMA(x, dx, xx * XTB_LIBM_FMA(xx, XTB_LIBM_FMA(xx, cs6, cs4), cs2));
    const double sn = XTB_LIBM_TABLE[k], ssn = XTB_LIBM_TABLE[k + 1], cs = XTB_LIBM_TABLE[k + 2],
                 ccs = XTB_LIBM_TABLE[k + 3];
    const double cor = XTB_LIBM_FMA(s, cs, XTB_LIBM_FMA(-c, sn, XTB_LIBM_FMA(s, ccs, ssn)));
    return copysign(sn + cor, x0);
}

// do_cos: cos(x + dx) for |x| < 0.8555
XTB_LIBM_FN double do_cos(const double x0, double dx) {
    if (x0 < 0) dx = -dx;
    const double ax = fabs(x0);
    const double u = ax + big;
    const double x = (ax - (u - big)) + dx;
    const uint32_t k = lo_word(u) << 2;
    const double xx = x * x;
    const double s = XTB_LIBM_FMA(x * xx, XTB_LIBM_FMA(xx, sn5, sn3), x);
    const double c = xx * XTB_LIBM_FMA(xx, XTB_LIBM_FMA(xx, cs6, cs4), cs2);
    const double sn = XTB_LIBM_TABLE[k], ssn = XTB_LIBM_TABLE[k + 1], cs = XTB_LIBM_TABLE[k + 2],
                 ccs = XTB_LIBM_TABLE[k + 3];
    const double cor = XTB_LIBM_FMA(-s, sn, XTB_LIBM_FMA(-c, cs, XTB_LIBM_FMA(-s, ssn, ccs)));
    return cs + cor;
}

// reduce_sincos: x = n * pi/2 + (a + da), |a| <= pi/4, for |x| < 105414350; returns n & 3
XTB_LIBM_FN int reduce_sincos(const double x, double& a, double& da) {
    const double t = XTB_LIBM_FMA(x, hpinv, toint);
    const double xn = t - toint;
    const double y = XTB_LIBM_FMA(-xn, mp2, XTB_LIBM_FMA(-xn, mp1, x));
    const int n = (int) (lo_word(t) & 3u);
    const double t2 = XTB_LIBM_FMA(-xn, pp3, y);
    double db = XTB_LIBM_FMA(-pp3, xn, y - t2);
    const double b = XTB_LIBM_FMA(-xn, pp4, t2);
    db = db + XTB_LIBM_FMA(-xn, pp4, t2 - b);
    a = b;
    da = db;
    return n;
}

XTB_LIBM_FN double do_sincos(const double a, const double da, const int n) {
    const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
    return (n & 2) ? -r : r;
}

}  // namespace xtb_libm

// sin(x) with glibc's bits
XTB_LIBM_ENTRY double xtb_sin_glibc(const double x) {
    using namespace xtb_libm;
    const uint32_t k = hi_word(x) & 0x7fffffffu;
    if (k < 0x3e500000u) return x;                               // |x| < 2^-26
    if (k < 0x3feb6000u) return do_sin(x, 0.0);                  // |x| < 0.855469
    if (k < 0x400368fdu) {                                       // |x| < 2.426265
        const double t = hp0 - fabs(x);
        return copysign(do_cos(t, hp1), x);
    }
    if (k < 0x419921fbu) {                                       // |x| < 105414350
        double a, da;
        const int n = reduce_sincos(x, a, da);
        return do_sincos(a, da, n);
    }
    return sin(x);
}

// cos(x) with glibc's bits
XTB_LIBM_ENTRY double xtb_cos_glibc(const double x) {
    using namespace xtb_libm;
    const uint32_t k = hi_word(x) & 0x7fffffffu;
    if (k < 0x3e400000u) return 1.0;                             // |x| < 2^-27
    if (k < 0x3feb6000u) return do_cos(x, 0.0);
    if (k < 0x400368fdu) {
        const double y = hp0 - fabs(x);
        const double a = y + hp1;
        const double da = (y - a) + hp1;
        return do_sin(a, da);
    }
    if (k < 0x419921fbu) {
        double a, da;
        const int n = reduce_sincos(x, a, da);
        return do_sincos(a, da, n + 1);
    }
    return cos(x);
}
