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
static __device__ const unsigned long long xtb_exptab_dev[XTB_EXPTAB_N] = {XTB_EXPTAB_VALUES};
#define XTB_LIBM_FN static __host__ __device__ __forceinline__
#define XTB_LIBM_ENTRY static __host__ __device__ __noinline__
#else
#define XTB_LIBM_FN static inline
#define XTB_LIBM_ENTRY static inline
#endif
static const double xtb_sincostab_host[XTB_SINCOSTAB_N] = {XTB_SINCOSTAB_VALUES};
static const unsigned long long xtb_exptab_host[XTB_EXPTAB_N] = {XTB_EXPTAB_VALUES};
#ifdef __CUDA_ARCH__
#define XTB_LIBM_EXPTAB xtb_exptab_dev
#else
#define XTB_LIBM_EXPTAB xtb_exptab_host
#endif

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

// ---- exp, expm1 (-> sinh, cosh of the quadrupole map) -------------------------------------------
// glibc's exp (sysdeps/ieee754/dbl-64/e_exp.c, Szabolcs Nagy's table-driven algorithm, N = 128)
// and expm1 (s_expm1.c, the FreeBSD msun routine), FMA placement of the x86-64 FMA builds that
// libm dispatches to; sinh and cosh (e_sinh.c, e_cosh.c) have no FMA build: plain operations.
constexpr double exp_invln2N = 0x1.71547652b82fep+7, exp_shift = 0x1.8p52;
constexpr double exp_negln2hiN = -0x1.62e42fefa0000p-8, exp_negln2loN = -0x1.cf79abc9e3b3ap-47;
constexpr double exp_C2 = 0x1.ffffffffffdbdp-2, exp_C3 = 0x1.555555555543cp-3,
                 exp_C4 = 0x1.55555cf172b91p-5, exp_C5 = 0x1.1111167a4d017p-7;
constexpr double em_ln2_hi = 0x1.62e42fee00000p-1, em_ln2_lo = 0x1.a39ef35793c76p-33,
                 em_invln2 = 0x1.71547652b82fep+0;
constexpr double em_Q1 = -0x1.11111111110f4p-5, em_Q2 = 0x1.a01a019fe5585p-10,
                 em_Q3 = -0x1.4ce199eaadbb7p-14, em_Q4 = 0x1.0cfca86e65239p-18,
                 em_Q5 = -0x1.afdb76e09c32dp-23;

XTB_LIBM_FN unsigned long long bits_of(const double v) {
#ifdef __CUDA_ARCH__
    return (unsigned long long) __double_as_longlong(v);
#else
    unsigned long long u;
    __builtin_memcpy(&u, &v, 8);
    return u;
#endif
}
XTB_LIBM_FN double double_of(const unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long) u);
#else
    double v;
    __builtin_memcpy(&v, &u, 8);
    return v;
#endif
}
XTB_LIBM_FN double add_to_high_word(const double v, const int32_t k20) {      // hi word += k << 20
    return double_of(bits_of(v) + ((unsigned long long) (uint32_t) k20 << 32));
}

// exp(x) for 2^-54 <= |x| < 512 (the caller sends the rest elsewhere)
XTB_LIBM_FN double exp_core(const double x) {
    const double kd0 = XTB_LIBM_FMA(x, exp_invln2N, exp_shift);
    const unsigned long long ki = bits_of(kd0);
    const double kd = kd0 - exp_shift;
    const double r = XTB_LIBM_FMA(kd, exp_negln2loN, XTB_LIBM_FMA(kd, exp_negln2hiN, x));
    const unsigned idx = 2u * (unsigned) (ki & 127u);
    const unsigned long long top = ki << 45;
    const double tail = double_of(XTB_LIBM_EXPTAB[idx]);
    const unsigned long long sbits = XTB_LIBM_EXPTAB[idx + 1] + top;
    const double r2 = r * r;
    const double p23 = XTB_LIBM_FMA(r, exp_C3, exp_C2);
    const double tr = r + tail;
    const double p45 = XTB_LIBM_FMA(r, exp_C5, exp_C4);
    double tmp = XTB_LIBM_FMA(p23, r2, tr);
    tmp = XTB_LIBM_FMA(r2 * r2, p45, tmp);
    const double scale = double_of(sbits);
    return XTB_LIBM_FMA(scale, tmp, scale);
}

// expm1(x) for |x| < 56 ln 2
XTB_LIBM_FN double expm1_core(double x) {
    const uint32_t hx = hi_word(x);
    const bool neg = (hx & 0x80000000u) != 0;
    const uint32_t ax = hx & 0x7fffffffu;
    int k = 0;
    double c = 0.;
    if (ax > 0x3fd62e42u) {                              // |x| > 0.5 ln 2
        double hi, lo;
        if (ax < 0x3ff0a2b2u) {                          // and |x| < 1.5 ln 2
            if (!neg) { hi = x - em_ln2_hi;  lo = em_ln2_lo;  k = 1; }
            else { hi = x + em_ln2_hi;  lo = -em_ln2_lo;  k = -1; }
        } else {
            const double kf = (neg ? -0.5 : 0.5) + x * em_invln2;
            k = (int) kf;
            const double t = (double) k;
            hi = XTB_LIBM_FMA(-t, em_ln2_hi, x);
            lo = t * em_ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (ax < 0x3c900000u) {
        return x;                                        // |x| < 2^-54
    }
    const double hfx = 0.5 * x;
    const double hxs = x * hfx;
    const double R2 = XTB_LIBM_FMA(hxs, em_Q3, em_Q2);
    const double R3 = XTB_LIBM_FMA(hxs, em_Q5, em_Q4);
    const double h2 = hxs * hxs;
    const double R1 = XTB_LIBM_FMA(hxs, em_Q1, 1.0);
    const double h4 = h2 * h2;
    const double r1 = XTB_LIBM_FMA(h4, R3, XTB_LIBM_FMA(h2, R2, R1));
    const double t = XTB_LIBM_FMA(-r1, hfx, 3.0);
    double e = hxs * ((r1 - t) / XTB_LIBM_FMA(-x, t, 6.0));
    if (k == 0) return x - XTB_LIBM_FMA(e, x, -hxs);
    e = XTB_LIBM_FMA(e - c, x, -c);
    e -= hxs;
    if (k == -1) return XTB_LIBM_FMA(x - e, 0.5, -0.5);
    if (k == 1) {
        if (x < -0.25) return -2.0 * (e - (x + 0.5));
        return XTB_LIBM_FMA(x - e, 2.0, 1.0);
    }
    if (k <= -2 || k > 56) {
        double y = 1.0 - (e - x);
        y = add_to_high_word(y, k << 20);
        return y - 1.0;
    }
    if (k < 20) {
        const double tt = double_of((unsigned long long) (uint32_t) (0x3ff00000 - (0x200000 >> k)) << 32);
        const double y = tt - (e - x);
        return add_to_high_word(y, k << 20);
    }
    const double tt = double_of((unsigned long long) (uint32_t) ((0x3ff - k) << 20) << 32);
    double y = x - (e + tt);
    y += 1.0;
    return add_to_high_word(y, k << 20);
}

}  // namespace xtb_libm

// exp(x), expm1(x), sinh(x), cosh(x) with glibc's bits; arguments beyond the ranges restated
// here (|x| >= 512 for exp, >= 22 for the hyperbolic pair: no focusing strength x length of a
// lattice comes near) go to the CUDA library function
XTB_LIBM_ENTRY double xtb_exp_glibc(const double x) {
    using namespace xtb_libm;
    const uint32_t abstop = (hi_word(x) >> 20) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop < 0x3c9u) return 1.0 + x;
        return exp(x);
    }
    return exp_core(x);
}
XTB_LIBM_ENTRY double xtb_expm1_glibc(const double x) {
    using namespace xtb_libm;
    if ((hi_word(x) & 0x7fffffffu) >= 0x4043687au) return expm1(x);
    return expm1_core(x);
}
XTB_LIBM_ENTRY double xtb_sinh_glibc(const double x) {
    using namespace xtb_libm;
    const uint32_t jx = hi_word(x), ix = jx & 0x7fffffffu;
    if (ix >= 0x40360000u) return sinh(x);                       // |x| >= 22
    const double h = (jx & 0x80000000u) ? -0.5 : 0.5;
    if (ix < 0x3e300000u) return x;                              // |x| < 2^-28
    const double t = expm1_core(fabs(x));
    if (ix < 0x3ff00000u) return h * ((t + t) - t * t / (t + 1.0));
    return h * (t + t / (t + 1.0));
}
XTB_LIBM_ENTRY double xtb_cosh_glibc(const double x) {
    using namespace xtb_libm;
    const uint32_t ix = hi_word(x) & 0x7fffffffu;
    if (ix >= 0x40360000u) return cosh(x);
    if (ix < 0x3fd62e43u) {                                      // |x| < 0.5 ln 2
        if (ix < 0x3c800000u) return 1.0;
        const double t = expm1_core(fabs(x));
        const double w = 1.0 + t;
        return 1.0 + (t * t) / (w + w);
    }
    const double t = xtb_libm::exp_core(fabs(x));
    return 0.5 * t + 0.5 / t;
}

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
