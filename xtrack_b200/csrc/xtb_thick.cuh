// xtb_thick.cuh -- thick-magnet body, edges and synchrotron radiation, on registers.
//
// Restates, in the reference's operation order:
//   track_magnet_drift.h     body "drift" maps (drift models 0,1,2,3,4,5,7,8)
//   track_magnet_kick.h      multipolar kick with curvature corrections
//   track_magnet.h:26-285    integrators (teapot / uniform / yoshida4) + WITH_RADIATION
//   track_magnet_edge.h      full / dipole-only edge (models 1, 2)
//   track_dipole_fringe.h:16-85, track_mult_fringe.h:13-129, track_wedge.h:14-98
//   track_dipole_edge_nonlinear.h:12-44
//   track_magnet_radiation.h:64-94,233-267; headers/synrad_spectrum.h:22-77,80-245,463-535
//   random/random_src/{uniform,uniform_accurate,exponential}.h; rng_src/base_rng.h:23-42
// Element-constant configuration (model/integrator selection, kick counts,
// coefficient scaling) is resolved by xtrack_b200/lowering.py::_lower_magnet;
// the parameter layout of each op is documented there and below.
#pragma once
#include "xtb_state.cuh"
#include "xtb_thin.cuh"

// Inlining of the thick-body pieces (tuning knob, measured in profiles/r01_history.md):
//   0  polar drift inlined into ONE drift function, drift and kick out of line
//   1  drift and kick inlined into the integrator loops, polar drift out of line
#ifndef XTB_THICK_INLINE_MODE
#define XTB_THICK_INLINE_MODE 1
#endif
#if XTB_THICK_INLINE_MODE == 0
#define XTB_POLAR_INLINE __forceinline__
#define XTB_DRIFTKICK_INLINE __noinline__
#else
#define XTB_POLAR_INLINE __noinline__
#define XTB_DRIFTKICK_INLINE __forceinline__
#endif

#define XTB_QELEM 1.60217662e-19
#define XTB_EPSILON_0 8.854187817620e-12
#define XTB_POW2(X) ((X) * (X))
#define XTB_POW3(X) ((X) * (X) * (X))
#define XTB_POW4(X) ((X) * (X) * (X) * (X))

// ---------------------------------------------------------------- drifts ----
// Trigonometry of element constants.  cos(h*s), sin(h*s), sin(h*s/2) in the polar drift and
// the curved exact bend depend on the element alone, yet the reference evaluates them per
// particle and per integrator sub-step (a default bend: 32 polar drifts = 96 sin/cos per
// particle).  The host lowering tabulates them for every sub-step length the integrator of
// this op will ask for (lowering.py::_trig_table, computed with the host libm -- the one the
// reference's CPU build uses), keyed by the bit pattern of the length; a length that is not
// in the table (there should be none) is evaluated here.
struct TrigTab {
    const double* t;     // entries of 4 doubles: length, cos(h*s), sin(h*s), sin(0.5*h*s)
    int n;
    double rho;          // 1 / h (valid iff n > 0)
};
#ifdef XTB_COUNT_TRIG_MISS
static long long xtb_trig_lookups = 0, xtb_trig_misses = 0;
#endif
__device__ __forceinline__ void trig_of(const TrigTab& tt, const double h, const double s,
                                        double& ca, double& sa, double& sa2) {
#ifdef XTB_COUNT_TRIG_MISS
    xtb_trig_lookups++;
#endif
    const long long key = __double_as_longlong(s);
    for (int i = 0; i < tt.n; ++i) {
        if (__double_as_longlong(tt.t[4 * i]) == key) {
            ca = tt.t[4 * i + 1];
            sa = tt.t[4 * i + 2];
            sa2 = tt.t[4 * i + 3];
            return;
        }
    }
#ifdef XTB_COUNT_TRIG_MISS
    xtb_trig_misses++;
#endif
    ca = cos(h * s);
    sa = sin(h * s);
    sa2 = sin(0.5 * h * s);
}

// track_polar_drift_single_particle, track_magnet_drift.h:45-87
template <bool FRZ>
__device__ XTB_POLAR_INLINE void polar_drift(PState& P, const double length, const double h,
                                             const TrigTab tt) {
    const double rvv = P.rvv;
    const double x = P.x, y = P.y, px = P.px, py = P.py;
    const double s = length;
    const double one_plus_delta = P.delta + 1.0;
    const double pz = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(px) - XTB_POW2(py));
    const double rho = (tt.n > 0) ? tt.rho : 1 / h;
    double ca, sa, sa2;
    trig_of(tt, h, s, ca, sa, sa2);
    const double _pz = 1 / pz;
    const double pxt = px * _pz;
    const double _ptt = 1 / (ca - sa * pxt);
    const double pst = (x + rho) * sa * _pz * _ptt;
    const double new_x = (x + rho * (2 * sa2 * sa2 + sa * pxt)) * _ptt;
    const double new_px = ca * px + sa * pz;
    const double new_y = y + pst * py;
    const double delta_ell = one_plus_delta * (x + rho) * sa / ca / pz / (1 - px * sa / ca / pz);
    P.x = new_x;
    P.px = new_px;
    P.y = new_y;
    if (!FRZ) {
        P.zeta += length - delta_ell / rvv;
        P.s += s;
    }
}

// track_expanded_combined_dipole_quad_single_particle, track_magnet_drift.h:91-213
template <bool FRZ>
__device__ __noinline__ void combined_dipole_quad(PState& P, const double length, const double k0_,
                                                  const double k1_, const double h) {
    const double x = P.x, y = P.y, px = P.px, py = P.py, rvv = P.rvv;
    const double delta_plus_1 = P.delta + 1;
    const double chi = P.chi;
    const double k0 = chi * k0_ / delta_plus_1;
    const double k1 = chi * k1_ / delta_plus_1;
    const double Kx = k0 * h + k1;
    const double Ky = -k1;
    double Sx, Sy, Cx, Cy;
    if (Kx > 0.0) {
        const double sqrt_Kx = sqrt(Kx);
        Sx = sin(sqrt_Kx * length) / sqrt_Kx;
        Cx = cos(sqrt_Kx * length);
    } else if (Kx < 0.0) {
        const double sqrt_Kx = sqrt(-Kx);
        Sx = sinh(sqrt_Kx * length) / sqrt_Kx;
        Cx = cosh(sqrt_Kx * length);
    } else {
        Sx = length;
        Cx = 1.0;
    }
    if (Ky > 0.0) {
        const double sqrt_Ky = sqrt(Ky);
        Sy = sin(sqrt_Ky * length) / sqrt_Ky;
        Cy = cos(sqrt_Ky * length);
    } else if (Ky < 0.0) {
        const double sqrt_Ky = sqrt(-Ky);
        Sy = sinh(sqrt_Ky * length) / sqrt_Ky;
        Cy = cosh(sqrt_Ky * length);
    } else {
        Sy = length;
        Cy = 1.0;
    }
    const double xp = px / delta_plus_1;
    const double yp = py / delta_plus_1;
    const double A = -Kx * x - k0 + h;
    const double B = xp;
    const double C = -Ky * y;
    const double D = yp;
    double x_ = x * Cx + xp * Sx;
    const double y_ = y * Cy + yp * Sy;
    const double px_ = (A * Sx + B * Cx) * delta_plus_1;
    const double py_ = (C * Sy + D * Cy) * delta_plus_1;
    if (Kx != 0.0)
        x_ = x_ + (k0 - h) * (Cx - 1.0) / Kx;
    else
        x_ = x_ - (k0 - h) * 0.5 * XTB_POW2(length);
    double length_ = length;
    if (Kx != 0.0) {
        length_ -= (h * ((Cx - 1.0) * xp + Sx * A + length * (k0 - h))) / Kx;
        length_ += 0.5 * (-(XTB_POW2(A) * Cx * Sx) / (2.0 * Kx) + (XTB_POW2(B) * Cx * Sx) / 2.0
                          + (XTB_POW2(A) * length) / (2.0 * Kx) + (XTB_POW2(B) * length) / 2.0
                          - (A * B * XTB_POW2(Cx)) / Kx + (A * B) / Kx);
    } else {
        length_ += h * length * (3.0 * length * xp + 6.0 * x - (k0 - h) * XTB_POW2(length)) / 6.0;
        length_ += 0.5 * (XTB_POW2(B)) * length;
    }
    if (Ky != 0.0) {
        length_ += 0.5 * (-(XTB_POW2(C) * Cy * Sy) / (2.0 * Ky) + (XTB_POW2(D) * Cy * Sy) / 2.0
                          + (XTB_POW2(C) * length) / (2.0 * Ky) + (XTB_POW2(D) * length) / 2.0
                          - (C * D * XTB_POW2(Cy)) / Ky + (C * D) / Ky);
    } else {
        length_ += 0.5 * XTB_POW2(D) * length;
    }
    const double dzeta = length - length_ / rvv;
    P.x = x_;
    P.px = px_;
    P.y = y_;
    P.py = py_;
    if (!FRZ) {
        P.zeta += dzeta;
        P.s += length;
    }
}

// track_curved_exact_bend_single_particle, track_magnet_drift.h:272-345
template <bool FRZ>
__device__ __noinline__ void curved_exact_bend(PState& P, const double length, const double k0,
                                               const double h, const TrigTab tt) {
    const double k0_chi = k0 * P.chi;
    if (fabs(k0_chi) < 1e-8) {
        polar_drift<FRZ>(P, length, h, tt);
        return;
    }
    const double rvv = P.rvv;
    const double x0 = P.x, y0 = P.y, px0 = P.px, py = P.py;
    const double s = length;
    const double one_plus_delta = P.delta + 1.0;
    const double hs = h * s;
    // (sin(hs / 2) == sin(0.5 * h * s): halving is exact)
    double cos_hs, sin_hs, sin_hs_2;
    trig_of(tt, h, s, cos_hs, sin_hs, sin_hs_2);
    const double pz0 = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(px0) - XTB_POW2(py));
    const double C = pz0 - k0_chi * ((1.0 / h) + x0);
    const double pxs = px0 * cos_hs + C * sin_hs;
    const double pzs = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(pxs) - XTB_POW2(py));
    const double delta_pz = (px0 - pxs) * (px0 + pxs) / (pz0 + pzs);
    const double delta_D = -2 * C * XTB_POW2(sin_hs_2) - px0 * sin_hs;
    const double delta_x = (delta_pz - delta_D) / k0_chi;
    const double delta_px = -2 * px0 * XTB_POW2(sin_hs_2) + C * sin_hs;
    const double N_a = px0 * delta_pz - pz0 * delta_px;
    const double D_a = pz0 * pzs + px0 * pxs;
    const double delta_a = atan2(N_a, D_a);
    const double integ = (hs + delta_a) / k0_chi;
    const double new_y = y0 + py * integ;
    const double delta_ell = one_plus_delta * integ;
    P.x += delta_x;
    P.px = pxs;
    P.y = new_y;
    if (!FRZ) {
        P.zeta += length - delta_ell / rvv;
        P.s += s;
    }
}

// track_straight_exact_bend_single_particle, track_magnet_drift.h:349-394
template <bool FRZ>
__device__ __noinline__ void straight_exact_bend(PState& P, const double length, const double k0) {
    const double k0_chi = k0 * P.chi;
    if (fabs(k0_chi) < 1e-8) {
        drift_exact<FRZ>(P, length);
        return;
    }
    const double rvv = P.rvv;
    const double x = P.x, y = P.y, px = P.px, py = P.py;
    const double s = length;
    const double one_plus_delta = P.delta + 1.0;
    const double A = 1.0 / sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(py));
    const double pz = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(px) - XTB_POW2(py));
    const double new_px = px - k0_chi * s;
    const double new_x =
        x + (sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(new_px) - XTB_POW2(py)) - pz) / k0_chi;
    const double D = asin(A * px) - asin(A * new_px);
    const double new_y = y + (py / k0_chi) * D;
    const double delta_ell = (one_plus_delta / k0_chi) * D;
    P.x = new_x;
    P.px = new_px;
    P.y = new_y;
    if (!FRZ) {
        P.zeta += length - delta_ell / rvv;
        P.s += s;
    }
}

// track_magnet_drift_single_particle, track_magnet_drift.h:468-555
template <bool FRZ>
__device__ XTB_DRIFTKICK_INLINE void magnet_drift(PState& P, const double length, const double k0,
                                             const double k1, const double h, const int drift_model,
                                             const TrigTab tt) {
    if (drift_model == -1) return;
    if (length == 0.0) return;
    switch (drift_model) {
    case 0: drift_expanded<FRZ>(P, length); break;
    case 1: drift_exact<FRZ>(P, length); break;
    case 2: polar_drift<FRZ>(P, length, h, tt); break;
    case 3: combined_dipole_quad<FRZ>(P, length, k0, k1, h); break;
    case 4: curved_exact_bend<FRZ>(P, length, k0, h, tt); break;
    case 5: straight_exact_bend<FRZ>(P, length, k0); break;
    case 7: {
        // nested Yoshida-4 bend, track_magnet_drift.h:521-531.  One inlined polar drift in a
        // loop over the step table (not 4 copies): the thick kernels are bound by instruction
        // fetch and local-memory traffic, not by arithmetic (profiles/r01_ncu_lep.md)
        const double pd[4] = {0.6756035959798289, -0.17560359597982889, -0.17560359597982889,
                              0.6756035959798289};
        const double pk[4] = {1.3512071919596578, -1.7024143839193155, 1.3512071919596578, 0.};
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            polar_drift<FRZ>(P, pd[j] * length, h, tt);
            if (j < 3) P.px = P.px - pk[j] * k0 * P.chi * length;
        }
        break;
    }
    case 8: {
        // nested Yoshida-6 bend, track_magnet_drift.h:532-550
        const double d[8] = {3.922568052387799819591407413100e-01, 5.100434119184584780271052295575e-01,
                             -4.710533854097565531482416645304e-01, 6.875316825251809316199569366290e-02,
                             6.875316825251809316199569366290e-02, -4.710533854097565531482416645304e-01,
                             5.100434119184584780271052295575e-01, 3.922568052387799819591407413100e-01};
        const double k[8] = {7.845136104775599639182814826199e-01, 2.355732133593569921359289764951e-01,
                             -1.177679984178870098432412305556e+00, 1.315186320683906284756403692882e+00,
                             -1.177679984178870098432412305556e+00, 2.355732133593569921359289764951e-01,
                             7.845136104775599639182814826199e-01, 0.};
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            polar_drift<FRZ>(P, d[j] * length, h, tt);
            if (j < 7) P.px = P.px - k[j] * k0 * P.chi * length;
        }
        break;
    }
    default: break;
    }
}

// ------------------------------------------------------------ body params ----
// OP_MAGNET_BODY parameter block (lowering.py::_lower_magnet):
//  q[0] length  q[1] k0_drift  q[2] k1_drift  q[3] h_drift  q[4] h_kick  q[5] hxl
//  q[6] A0 = k0_h_correction*length + k0l   q[7] A1 = k1_h_correction*length + k1l
//  q[8] htot    q[9] (int) order_user | order_rel << 32
//  q[10..17] k0_tot k1_tot k2 k3 k0s k1s k2s k3s   (field evaluation for radiation)
//  q[18..25] main coefficients (order 3, Horner order, pairs)
//  then user coefficients (order_user+1 pairs), then rel coefficients (order_rel+1 pairs),
//  then the trig table: (int) n, 1/h, n x [length, cos(h*s), sin(h*s), sin(h*s/2)]
//  aux: integrator[0:2] drift_model+1[2:6] rot_frame[6] has_user[7] has_rel[8]
//       has_main[9] radiation_flag[10:12] drift_only[12] num_kicks[13:32]
struct BodyPar {
    const double* q;
    const double* cm;    // main
    const double* cu;    // user
    const double* cr;    // rel
    int order_user, order_rel;
    int integrator, drift_model, rot_frame, has_user, has_rel, has_main, radiation_flag, drift_only;
    int num_kicks;
    TrigTab trig;
};

__device__ __forceinline__ BodyPar body_par(const double* q, const int32_t aux) {
    BodyPar b;
    b.q = q;
    const unsigned long long w = (unsigned long long) __double_as_longlong(q[9]);
    b.order_user = (int) (w & 0xffffffffu);
    b.order_rel = (int) (w >> 32);
    b.cm = q + 18;
    b.cu = b.cm + 8;
    b.cr = b.cu + 2 * (b.order_user + 1);
    const double* tq = b.cr + 2 * (b.order_rel + 1);
    b.trig.n = (int) __double_as_longlong(tq[0]);
    b.trig.rho = tq[1];
    b.trig.t = tq + 2;
    const uint32_t a = (uint32_t) aux;
    b.integrator = a & 3;
    b.drift_model = (int) ((a >> 2) & 15) - 1;
    b.rot_frame = (a >> 6) & 1;
    b.has_user = (a >> 7) & 1;
    b.has_rel = (a >> 8) & 1;
    b.has_main = (a >> 9) & 1;
    b.radiation_flag = (a >> 10) & 3;
    b.drift_only = (a >> 12) & 1;
    b.num_kicks = (int) (a >> 13);
    return b;
}

// track_magnet_kick_single_particle, track_magnet_kick.h:24-144
template <bool FRZ>
__device__ XTB_DRIFTKICK_INLINE void magnet_kick(PState& P, const BodyPar& b, const double kick_weight) {
    const double chi = P.chi, x = P.x, y = P.y;
    const double length = b.q[0];
    double m, n;
    if (b.has_user) {
        horner_kick(x, y, chi, b.cu, b.order_user, m, n);
        P.px += kick_weight * (-m);
        P.py += kick_weight * n;
    }
    if (b.has_rel) {
        horner_kick(x, y, chi, b.cr, b.order_rel, m, n);
        P.px += kick_weight * (-m);
        P.py += kick_weight * n;
    }
    if (b.has_main) {
        horner_kick(x, y, chi, b.cm, 3, m, n);
        P.px += kick_weight * (-m);
        P.py += kick_weight * n;
    }
    double dpx = 0, dpy = 0, dzeta = 0;
    const double h = b.q[4], hxl = b.q[5];
    if (b.rot_frame) {
        const double hl = h * length * kick_weight + hxl * kick_weight;
        dpx += hl * (1. + P.delta);
        dzeta += -P.rv0v * hl * x;
    }
    const double htot = b.q[8];
    dpx += -chi * b.q[6] * kick_weight * htot * x;
    dpx += htot * chi * b.q[7] * kick_weight * (-x * x + 0.5 * y * y);
    dpy += htot * chi * b.q[7] * kick_weight * x * y;
    P.px += dpx;
    P.py += dpy;
    if (!FRZ) P.zeta += dzeta;
}

// ------------------------------------------------------------- radiation ----
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

struct RadCtx {      // per-call context: rng state + failure flag
    Rng r;
    bool seeded;
    bool rng_error;
};

// RandomUniform_generate, random_src/uniform.h:34-53
__device__ __forceinline__ double rand_uniform(RadCtx& c) {
    if (!c.seeded) { c.rng_error = true;  return 0; }
    return rng_u32(c.r) / 4294967296.0;
}
__device__ __forceinline__ uint32_t rand_u32(RadCtx& c) {
    if (!c.seeded) { c.rng_error = true;  return 0; }
    return rng_u32(c.r);
}
// RandomUniformAccurate_generate, uniform_accurate.h:21-37
__device__ __forceinline__ double rand_uniform_accurate(RadCtx& c) {
    const double T = 4294967296.0;
    double out = 0;
    out += rand_u32(c) / T;
    out += rand_u32(c) / (T * T);
    out += rand_u32(c) / (T * T * T);
    out += rand_u32(c) / (T * T * T * T);
    out += rand_u32(c) / (T * T * T * T * T);
    out += rand_u32(c) / (T * T * T * T * T * T);
    return out;
}
// RandomExponential_generate, exponential.h:19-25
__device__ __forceinline__ double rand_exponential(RadCtx& c) {
    double x1 = rand_uniform(c);
    while (x1 == 0.0 && !c.rng_error) x1 = rand_uniform(c);
    return -log(x1);
}

// SynRad, headers/synrad_spectrum.h:80-175 (Chebyshev series from H.Burkhardt)
static __device__ __noinline__ double synrad_fn(const double x) {
    double synrad = 0.;
    if (x > 0. && x < 800.) {
        if (x < 6.) {
            double a, b, z;
            z = x * x / 16. - 2.;
            b = .00000000000000000012;
            a = z * b + .00000000000000000460;
            b = z * a - b + .00000000000000031738;
            a = z * b - a + .00000000000002004426;
            b = z * a - b + .00000000000111455474;
            a = z * b - a + .00000000005407460944;
            b = z * a - b + .00000000226722011790;
            a = z * b - a + .00000008125130371644;
            b = z * a - b + .00000245751373955212;
            a = z * b - a + .00006181256113829740;
            b = z * a - b + .00127066381953661690;
            a = z * b - a + .02091216799114667278;
            b = z * a - b + .26880346058164526514;
            a = z * b - a + 2.61902183794862213818;
            b = z * a - b + 18.65250896865416256398;
            a = z * b - a + 92.95232665922707542088;
            b = z * a - b + 308.15919413131586030542;
            a = z * b - a + 644.86979658236221700714;
            double p;
            p = .5 * z * a - b + 414.56543648832546975110;
            a = .00000000000000000004;
            b = z * a + .00000000000000000289;
            a = z * b - a + .00000000000000019786;
            b = z * a - b + .00000000000001196168;
            a = z * b - a + .00000000000063427729;
            b = z * a - b + .00000000002923635681;
            a = z * b - a + .00000000115951672806;
            b = z * a - b + .00000003910314748244;
            a = z * b - a + .00000110599584794379;
            b = z * a - b + .00002581451439721298;
            a = z * b - a + .00048768692916240683;
            b = z * a - b + .00728456195503504923;
            a = z * b - a + .08357935463720537773;
            b = z * a - b + .71031361199218887514;
            a = z * b - a + 4.26780261265492264837;
            b = z * a - b + 17.05540785795221885751;
            a = z * b - a + 41.83903486779678800040;
            double q;
            q = .5 * z * a - b + 28.41787374362784178164;
            double y;
            y = pow(x, 2. / 3.);
            synrad = (p / y - q * y - 1.) * 1.81379936423421784215530788143;
        } else {
            double a, b, z;
            z = 20. / x - 2.;
            a = .00000000000000000001;
            b = z * a - .00000000000000000002;
            a = z * b - a + .00000000000000000006;
            b = z * a - b - .00000000000000000020;
            a = z * b - a + .00000000000000000066;
            b = z * a - b - .00000000000000000216;
            a = z * b - a + .00000000000000000721;
            b = z * a - b - .00000000000000002443;
            a = z * b - a + .00000000000000008441;
            b = z * a - b - .00000000000000029752;
            a = z * b - a + .00000000000000107116;
            b = z * a - b - .00000000000000394564;
            a = z * b - a + .00000000000001489474;
            b = z * a - b - .00000000000005773537;
            a = z * b - a + .00000000000023030657;
            b = z * a - b - .00000000000094784973;
            a = z * b - a + .00000000000403683207;
            b = z * a - b - .00000000001785432348;
            a = z * b - a + .00000000008235329314;
            b = z * a - b - .00000000039817923621;
            a = z * b - a + .00000000203088939238;
            b = z * a - b - .00000001101482369622;
            a = z * b - a + .00000006418902302372;
            b = z * a - b - .00000040756144386809;
            a = z * b - a + .00000287536465397527;
            b = z * a - b - .00002321251614543524;
            a = z * b - a + .00022505317277986004;
            b = z * a - b - .00287636803664026799;
            a = z * b - a + .06239591359332750793;
            double p;
            p = .5 * z * a - b + 1.06552390798340693166;
            synrad = p * sqrt(0.5 * XTB_PI / x) / exp(x);
        }
    }
    return synrad;
}

// synrad_gen_photon_energy_normalized, synrad_spectrum.h:178-207
static __device__ __noinline__ double synrad_gen_photon_energy_normalized(RadCtx& c) {
    const double xlow = 1.;
    const double a1 = 2.149528241534391;
    const double a2 = 1.770750801624037;
    const double c1 = 0.;
    const double ratio = 0.908250405131381;
    double appr = 0, exact = 1000, result = 0;
    do {
        if (c.rng_error) return 0.;
        if (rand_uniform(c) < ratio) {
            result = c1 + (1. - c1) * rand_uniform(c);
            const double tmp = result * result;
            result *= tmp;
            exact = synrad_fn(result);
            appr = a1 / tmp;
        } else {
            const double u = rand_uniform_accurate(c);
            if (u < 1.e-50) continue;
            result = xlow - log(u);
            exact = synrad_fn(result);
            appr = a2 * exp(-result);
        }
    } while (exact < appr * rand_uniform(c));
    return result;
}

// synrad_average_number_of_photons, synrad_spectrum.h:209-219
__device__ __forceinline__ double synrad_average_number_of_photons(const double mass0, const double q0,
                                                                   const double beta0_gamma0,
                                                                   const double B_T, const double lpath) {
    const double mass0_kg = mass0 * XTB_QELEM / XTB_C_LIGHT / XTB_C_LIGHT;
    const double P0_J = mass0_kg * beta0_gamma0 * XTB_C_LIGHT;
    const double Q0_coulomb = fabs(q0) * XTB_QELEM;
    const double curv = B_T / P0_J * Q0_coulomb;
    const double kick = curv * lpath;
    return 2.5 / 1.732050807568877 * 0.0072973525693 * beta0_gamma0 * fabs(kick);
}

// synrad_average_kick, synrad_spectrum.h:22-77 (mean model)
template <bool FRZ>
__device__ __noinline__ void synrad_average_kick(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                                 const double B_T, const double lpath) {
    const double gamma0 = G.ld(F_GAMMA0);
    const double mass0 = a.part.mass0;
    const double q0 = a.part.q0;
    const double Q0_coulomb = fabs(q0) * XTB_QELEM;
    const double mass0_kg = mass0 / XTB_C_LIGHT / XTB_C_LIGHT * XTB_QELEM;
    const double delta = P.delta;
    const double gamma = gamma0 * (1 + delta);
    const double r0_m = Q0_coulomb * Q0_coulomb
                        / (4 * XTB_PI * XTB_EPSILON_0 * mass0_kg * XTB_C_LIGHT * XTB_C_LIGHT);
    const double Ps_W = 2 * r0_m * XTB_C_LIGHT * Q0_coulomb * Q0_coulomb * gamma * gamma * B_T * B_T
                        / (3 * mass0_kg);
    const double Delta_E_eV = Ps_W * lpath / XTB_C_LIGHT / XTB_QELEM;
    const double f_t = 1 - Delta_E_eV / (gamma0 * mass0 * (1 + delta));
    update_delta<FRZ>(P, G, G.ld(F_BETA0), (delta + 1) * f_t - 1);
    P.px *= f_t;
    P.py *= f_t;
}

// synrad_emit_photons, synrad_spectrum.h:463-535 (quantum model, no photon log)
template <bool FRZ>
__device__ __noinline__ void synrad_emit_photons(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                                 const double B_T, const double lpath) {
    if (fabs(B_T) < 1e-4) return;
    const double mass0 = a.part.mass0;
    const double q0 = a.part.q0;
    const double gamma0 = G.ld(F_GAMMA0);
    const double beta0 = G.ld(F_BETA0);
    const double Q0_coulomb = fabs(q0) * XTB_QELEM;
    const double mass0_kg = mass0 * XTB_QELEM / XTB_C_LIGHT / XTB_C_LIGHT;
    const double P0_J = mass0_kg * beta0 * gamma0 * XTB_C_LIGHT;
    const double curv = B_T / P0_J * Q0_coulomb;
    const double delta = P.delta;
    const double gamma = gamma0 * (1 + delta);
    const double p0c = G.ld(F_P0C);
    const double initial_energy = sqrt(p0c * p0c + mass0 * mass0) + G.ld(F_PTAU) * p0c;
    double energy = initial_energy;

    RadCtx c;
    c.r.s1 = G.ldu(F_RNG_S1);  c.r.s2 = G.ldu(F_RNG_S2);
    c.r.s3 = G.ldu(F_RNG_S3);  c.r.s4 = G.ldu(F_RNG_S4);
    c.seeded = !(c.r.s1 == 0 && c.r.s2 == 0 && c.r.s3 == 0 && c.r.s4 == 0);
    c.rng_error = false;

    const double n_avg = synrad_average_number_of_photons(mass0, q0, beta0 * gamma0, B_T, lpath);
    double n = rand_exponential(c);
    while (n < n_avg && !c.rng_error) {
        const double c1 = 1.5 * 1.973269804593025e-07;
        const double energy_critical = c1 * (gamma * gamma * gamma0) * curv;
        const double energy_loss = synrad_gen_photon_energy_normalized(c) * energy_critical;
        if (energy_loss >= energy) {
            energy = 0.0;
            break;
        }
        energy -= energy_loss;
        n += rand_exponential(c);
    }
    G.stu(F_RNG_S1, c.r.s1);  G.stu(F_RNG_S2, c.r.s2);
    G.stu(F_RNG_S3, c.r.s3);  G.stu(F_RNG_S4, c.r.s4);
    if (c.rng_error) {       // RNG_ERR_SEEDS_NOT_SET, uniform.h:40-43
        kill_particle<FRZ>(P, G, -20);
        return;
    }
    if (energy <= 0.0) {
        P.state = -10;       // XT_LOST_ALL_E_IN_SYNRAD
    } else {
        const double f_t = energy / initial_energy;
        update_delta<FRZ>(P, G, beta0, (P.delta + 1) * f_t - 1);
        P.px *= f_t;
        P.py *= f_t;
    }
}

// evaluate_field_from_strengths, track_magnet_kick.h:265-370 (no solenoid terms)
__device__ __forceinline__ void field_from_strengths(const BodyPar& b, const double p0c, const double q0,
                                                     const double x, const double y, double& Bx_T,
                                                     double& By_T) {
    const double length = b.q[0];
    if (length == 0.0) { Bx_T = 0.0;  By_T = 0.0;  return; }
    double dpx_mul = 0., dpy_mul = 0., dpx_rel = 0., dpy_rel = 0., dpx_main = 0., dpy_main = 0.;
    double m, n;
    if (b.has_user) { horner_kick(x, y, 1., b.cu, b.order_user, m, n);  dpx_mul = -m;  dpy_mul = n; }
    if (b.has_rel) { horner_kick(x, y, 1., b.cr, b.order_rel, m, n);  dpx_rel = -m;  dpy_rel = n; }
    {   // main strengths include the part integrated by the drift map (k0_drift + k0_kick, ...)
        double knl_main[4], ksl_main[4];
        for (int i = 0; i < 4; ++i) { knl_main[i] = b.q[10 + i] * length;  ksl_main[i] = b.q[14 + i] * length; }
        double inv_factorial = 1. / (3 * 2);
        int index = 3;
        double pm = 1. * knl_main[index] * 1 * inv_factorial;
        double qm = 1. * ksl_main[index] * 1 * inv_factorial;
        while (index > 0) {
            const double zre = pm * x - qm * y;
            const double zim = pm * y + qm * x;
            inv_factorial *= index;
            index -= 1;
            pm = 1. * knl_main[index] * 1 * inv_factorial + zre;
            qm = 1. * ksl_main[index] * 1 * inv_factorial + zim;
        }
        dpx_main = -pm;
        dpy_main = qm;
    }
    const double dpx = dpx_mul + dpx_main + dpx_rel;
    const double dpy = dpy_mul + dpy_main + dpy_rel;
    const double brho_0 = p0c / XTB_C_LIGHT / q0;
    Bx_T = dpy * brho_0 / length - 0.5 * 0. * brho_0 * (x - 0.);
    By_T = -dpx * brho_0 / length - 0.5 * 0. * brho_0 * (y - 0.);
}

// compute_b_perp_mod, track_magnet_radiation.h:64-94 (incl. the reference's
// `1 - iix*iix + iiy*iiy` in direction_of_motion :22)
__device__ __forceinline__ double b_perp_mod(const double kin_px, const double kin_py, const double delta,
                                             const double Bx, const double By, const double Bz) {
    const double iix = kin_px / (1. + delta);
    const double iiy = kin_py / (1. + delta);
    const double iis = sqrt(1 - iix * iix + iiy * iiy);
    const double B_par = Bx * iix + By * iiy + Bz * iis;
    const double px_ = Bx - B_par * iix;
    const double py_ = By - B_par * iiy;
    const double pz_ = Bz - B_par * iis;
    return sqrt(px_ * px_ + py_ * py_ + pz_ * pz_);
}

// --------------------------------------------------------------- the body ----
// One integrator step wrapped by WITH_RADIATION, track_magnet.h:92-178.
struct RadSnapshot {
    double old_x, old_y, old_zeta, old_px, old_py;
};

template <bool SYNRAD>
__device__ __forceinline__ void rad_begin(RadSnapshot& s, const PState& P) {
    if (SYNRAD) { s.old_x = P.x;  s.old_y = P.y;  s.old_zeta = P.zeta;  s.old_px = P.px;  s.old_py = P.py; }
}

template <bool SYNRAD, bool FRZ>
__device__ __forceinline__ void rad_end(const RadSnapshot& s, PState& P, const PSlot& G,
                                        const XtbTrackArgs& a, const BodyPar& b, const double ll) {
    if (!SYNRAD) return;
    const double length = b.q[0];
    if (!(b.radiation_flag && length > 0)) return;   // spin is (0,0,0): magnet_spin is a no-op
    const double p0c = G.ld(F_P0C);
    const double q0 = a.part.q0;
    const double mean_x = 0.5 * (s.old_x + P.x);
    const double mean_y = 0.5 * (s.old_y + P.y);
    const double mean_kin_px = 0.5 * (s.old_px + P.px);
    const double mean_kin_py = 0.5 * (s.old_py + P.py);
    double Bx_T, By_T;
    field_from_strengths(b, p0c, q0, mean_x, mean_y, Bx_T, By_T);
    const double dzeta = P.zeta - s.old_zeta;
    const double l_path = P.rvv * (ll - dzeta);
    const double B_perp_T = b_perp_mod(mean_kin_px, mean_kin_py, P.delta, Bx_T, By_T, 0. * (p0c / XTB_C_LIGHT / q0));
    if (b.radiation_flag == 1) {
        synrad_average_kick<FRZ>(P, G, a, B_perp_T, l_path);
    } else if (b.radiation_flag == 2) {
        synrad_emit_photons<FRZ>(P, G, a, B_perp_T, l_path);
    }
}

// track_magnet_body_single_particle, track_magnet.h:26-285
template <bool SYNRAD, bool FRZ>
__device__ __noinline__ void magnet_body(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                         const double* __restrict__ q, const int32_t aux) {
    const BodyPar b = body_par(q, aux);
    const double length = q[0], k0d = q[1], k1d = q[2], hd = q[3];
    const int dm = b.drift_model;
    RadSnapshot snap;
#define XTB_DRIFT(dl) magnet_drift<FRZ>(P, (dl), k0d, k1d, hd, dm, b.trig)
#define XTB_KICK(w) magnet_kick<FRZ>(P, b, (w))
    if (b.drift_only) {
        rad_begin<SYNRAD>(snap, P);
        XTB_DRIFT(length);
        rad_end<SYNRAD, FRZ>(snap, P, G, a, b, length);
        return;
    }
    const int nk = b.num_kicks;
    if (b.integrator == 1) {            // teapot
        rad_begin<SYNRAD>(snap, P);
        const double kick_weight = 1. / nk;
        double edge_drift_weight = 0.5;
        double inside_drift_weight = 0;
        if (nk > 1) {
            edge_drift_weight = 1. / (2 * (1 + nk));
            inside_drift_weight = ((double) nk) / ((double) ((int64_t) nk * nk) - 1);
        }
        XTB_DRIFT(edge_drift_weight * length);
        for (int i = 0; i < nk - 1; ++i) {
            XTB_KICK(kick_weight);
            XTB_DRIFT(inside_drift_weight * length);
        }
        XTB_KICK(kick_weight);
        XTB_DRIFT(edge_drift_weight * length);
        rad_end<SYNRAD, FRZ>(snap, P, G, a, b, length);
    } else if (b.integrator == 3) {     // uniform
        const double kick_weight = 1. / nk;
        const double drift_weight = kick_weight;
        for (int i = 0; i < nk; ++i) {
            rad_begin<SYNRAD>(snap, P);
            XTB_DRIFT(0.5 * drift_weight * length);
            XTB_KICK(kick_weight);
            XTB_DRIFT(0.5 * drift_weight * length);
            rad_end<SYNRAD, FRZ>(snap, P, G, a, b, drift_weight * length);
        }
    } else if (b.integrator == 2) {     // yoshida 4
        const int num_slices = nk / 7 + (nk % 7 != 0);
        const double slice_length = length / (num_slices);
        const double kick_weight = 1. / num_slices;
        const double d0 = 3.922568052387799819591407413100e-01, d1 = 5.100434119184584780271052295575e-01,
                     d2 = -4.710533854097565531482416645304e-01, d3 = 6.875316825251809316199569366290e-02;
        const double y0 = 7.845136104775599639182814826199e-01, y1 = 2.355732133593569921359289764951e-01,
                     y2 = -1.177679984178870098432412305556e+00, y3 = 1.315186320683906284756403692882e+00;
        const double yd[8] = {d0, d1, d2, d3, d3, d2, d1, d0};
        const double yk[8] = {y0, y1, y2, y3, y2, y1, y0, 0.};
        for (int ii = 0; ii < num_slices; ++ii) {
            rad_begin<SYNRAD>(snap, P);
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {       // one copy of the drift and kick bodies
                XTB_DRIFT(slice_length * yd[j]);
                if (j < 7) XTB_KICK(kick_weight * yk[j]);
            }
            rad_end<SYNRAD, FRZ>(snap, P, G, a, b, slice_length);
        }
    }
#undef XTB_DRIFT
#undef XTB_KICK
}

// ------------------------------------------------------------------ edges ----
// DipoleFringe_single_particle (MAD-NG form), track_dipole_fringe.h:16-85
template <bool FRZ>
__device__ __noinline__ void dipole_fringe(PState& P, const PSlot& G, const double fint, const double hgap,
                                           const double k0) {
    if (fabs(k0) < 10e-10) return;
    const double beta0 = G.ld(F_BETA0);
    const double x = P.x, px = P.px, y = P.y, py = P.py;
    const double t = P.zeta / beta0;
    const double pt = G.ld(F_PTAU);
    const double delta = P.delta;
    const double fh = hgap * fint;
    const double fsad = (fh > 10e-10) ? 1. / (72 * fh) : 0;
    const double k0w = k0 * P.chi;
    const double _beta = 1. / beta0;
    const double b0 = k0w;
    const double dpp = XTB_POW2(1. + delta);
    const double pz = sqrt(dpp - XTB_POW2(px) - XTB_POW2(py));
    const double _pz = 1. / pz;
    const double relp = 1. / sqrt(dpp);
    const double tfac = -(_beta + pt);
    const double c2 = b0 * fh * 2;
    const double c3 = XTB_POW2(b0) * fsad * relp;
    const double xp = px / pz;
    const double yp = py / pz;
    const double xyp = xp * yp;
    const double yp2 = 1. + XTB_POW2(yp);
    const double xp2 = XTB_POW2(xp);
    const double _yp2 = 1. / yp2;
    const double fi0 = atan((xp * _yp2)) - c2 * (1 + xp2 * (1 + yp2)) * _pz;
    const double co2 = b0 / XTB_POW2(cos(fi0));
    const double co1 = co2 / (1 + XTB_POW2(xp * _yp2)) * _yp2;
    const double co3 = co2 * c2;
    const double fi1 = co1 - co3 * 2 * xp * (1 + yp2) * _pz;
    const double fi2 = -2 * co1 * xyp * _yp2 - co3 * 2 * xp * xyp * _pz;
    const double fi3 = +co3 * (1 + xp2 * (1 + yp2)) * XTB_POW2(_pz);
    const double kx = fi1 * (1 + xp2) * _pz + fi2 * xyp * _pz - fi3 * xp;
    const double ky = fi1 * xyp * _pz + fi2 * yp2 * _pz - fi3 * yp;
    const double kz = fi1 * tfac * xp * XTB_POW2(_pz) + fi2 * tfac * yp * XTB_POW2(_pz) - fi3 * tfac * _pz;
    const double new_y = 2 * y / (1 + sqrt(1 - 2 * ky * y));
    const double new_x = x + 0.5 * kx * XTB_POW2(new_y);
    const double new_py = py - 4 * c3 * XTB_POW3(new_y) - b0 * tan(fi0) * new_y;
    const double new_t = t + 0.5 * kz * XTB_POW2(new_y) + c3 * XTB_POW4(new_y) * XTB_POW2(relp) * tfac;
    const double new_zeta = new_t * beta0;
    P.x = new_x;
    P.y = new_y;
    P.py = new_py;
    if (!FRZ) P.zeta = new_zeta;
}

// MultFringe_track_single_particle, track_mult_fringe.h:13-129
template <bool FRZ>
__device__ __noinline__ void mult_fringe(PState& P, const PSlot& G, const double* __restrict__ kn,
                                         const double* __restrict__ ks, const int k_order,
                                         const double* __restrict__ knl, const double* __restrict__ ksl,
                                         const int kl_order, const double length, const int is_exit,
                                         const unsigned min_order) {
    if (k_order == -1 && kl_order == -1) return;
    const double beta0 = G.ld(F_BETA0);
    const double direction = is_exit ? -1 : 1;
    const double x = P.x, px = P.px, y = P.y, py = P.py;
    const double t = P.zeta / beta0;
    const double pt = G.ld(F_PTAU);
    const double rpp = P.rpp;
    const double chi = P.chi;
    double rx = 1, ix = 0, fx = 0, fxx = 0, fxy = 0, fy = 0, fyx = 0, fyy = 0;
    const unsigned order = (unsigned) ((k_order > kl_order) ? k_order : kl_order);
    double inv_factorial = 1;
    for (unsigned ii = 0; ii <= order; ii++) {
        if (ii > 1) inv_factorial /= ii;
        const double component = ii + 1;
        const double drx = rx;
        const double dix = ix;
        rx = drx * x - dix * y;
        ix = drx * y + dix * x;
        double kn_total = 0, ks_total = 0;
        if (ii >= min_order) {
            if ((int) ii <= k_order) {
                kn_total += kn[ii] * inv_factorial;
                ks_total += ks[ii] * inv_factorial;
            }
            if ((int) ii <= kl_order && length != 0.) {
                kn_total += knl[ii] / length * inv_factorial;
                ks_total += ksl[ii] / length * inv_factorial;
            }
        }
        const double nj = -direction / (4 * (component + 1));
        const double nf = (component + 2) / component;
        const double kj = kn_total * chi;
        const double ksj = ks_total * chi;
        double u, v, du, dv;
        if (ii == 0) {
            u = nj * (-ksj * ix);
            v = nj * (ksj * rx);
            du = nj * (-ksj * dix);
            dv = nj * (ksj * drx);
        } else {
            u = nj * (kj * rx - ksj * ix);
            v = nj * (kj * ix + ksj * rx);
            du = nj * (kj * drx - ksj * dix);
            dv = nj * (kj * dix + ksj * drx);
        }
        const double dux = component * du;
        const double dvx = component * dv;
        const double duy = -component * dv;
        const double dvy = component * du;
        fx = fx + u * x + nf * v * y;
        fy = fy + u * y - nf * v * x;
        fxx = fxx + dux * x + nf * dvx * y + u;
        fyy = fyy + duy * y - nf * dvy * x + u;
        fxy = fxy + duy * x + nf * (dvy * y + v);
        fyx = fyx + dux * y - nf * (dvx * x + v);
    }
    const double a = 1 - fxx * rpp;
    const double b = -fyx * rpp;
    const double c = -fxy * rpp;
    const double d = 1 - fyy * rpp;
    const double det = (a * d - b * c);
    const double new_px = (d * px - b * py) / det;
    const double new_py = (a * py - c * px) / det;
    const double delta_t = (1 / beta0 + pt) * (new_px * fx + new_py * fy) * XTB_POW3(rpp);
    P.x += -fx * rpp;
    P.y += -fy * rpp;
    P.px = new_px;
    P.py = new_py;
    if (!FRZ) P.zeta = (t + delta_t) * beta0;
}

// Wedge_single_particle, track_wedge.h:14-74
template <bool FRZ>
__device__ __noinline__ void wedge(PState& P, const PSlot& G, const double theta, const double k0) {
    const double b1 = k0 * P.chi;
    if (fabs(b1) < 10e-10) {
        const double sin_ = sin(theta), cos_ = cos(theta), tan_ = tan(theta);
        yrotation<FRZ>(P, G, -sin_, cos_, -tan_);
        return;
    }
    const double rvv = P.rvv;
    const double x = P.x, px = P.px, py = P.py;
    const double one_plus_delta = P.delta + 1.0;
    const double A = 1.0 / sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(py));
    const double pz = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(px) - XTB_POW2(py));
    const double new_px = px * cos(theta) + (pz - b1 * x) * sin(theta);
    const double new_pz = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(new_px) - XTB_POW2(py));
    const double new_x = x * cos(theta)
        + (x * px * sin(2 * theta) + XTB_POW2(sin(theta)) * (2 * x * pz - b1 * XTB_POW2(x)))
              / (new_pz + pz * cos(theta) - px * sin(theta));
    const double D = asin(A * px) - asin(A * new_px);
    const double delta_y = py * (theta + D) / b1;
    const double delta_ell = one_plus_delta * (theta + D) / b1;
    P.x = new_x;
    P.y += delta_y;
    P.px = new_px;
    if (!FRZ) P.zeta += -delta_ell / rvv;
}

// Quad_wedge_single_particle, track_wedge.h:77-98
__device__ __forceinline__ void quad_wedge(PState& P, const double theta, const double k1) {
    const double b2 = k1 * P.chi;
    const double x = P.x, y = P.y, px = P.px, py = P.py;
    P.px = px - b2 * x * x * theta + b2 * y * y / 2 * theta;
    P.py = py + b2 * x * y * theta;
}

// track_magnet_edge_particles models 1 (full) and 2 (dipole-only), track_magnet_edge.h:79-160.
// q: [0] k0 (sign already flipped at the exit) [1] sin [2] cos [3] tan of the face angle
//    [4] fint [5] hgap [6] face_angle [7] length/factor_knl_ksl [8..11] knorm [12..15] kskew
//    [16..] knl[nkl], ksl[nkl];  aux: is_exit[0] model[1:3] should_rotate[3] nkl[4:]
template <bool FRZ>
__device__ __noinline__ void magnet_edge(PState& P, const PSlot& G, const double* __restrict__ q,
                                         const int32_t aux) {
    const int is_exit = aux & 1;
    const int model = (aux >> 1) & 3;
    const int should_rotate = (aux >> 3) & 1;
    const int nkl = aux >> 4;
    const double k0 = q[0], sin_ = q[1], cos_ = q[2], tan_ = q[3], fint = q[4], hgap = q[5];
    const double face_angle = q[6], length_eff = q[7];
    const double* knorm = q + 8;
    const double* kskew = q + 12;
    const double* knl = q + 16;
    const double* ksl = knl + nkl;
    if (is_exit == 0) {
        if (should_rotate) yrotation<FRZ>(P, G, -sin_, cos_, -tan_);
        dipole_fringe<FRZ>(P, G, fint, hgap, k0);
        if (model == 1) {
            mult_fringe<FRZ>(P, G, knorm, kskew, 3, knl, ksl, nkl - 1, length_eff, is_exit, 1);
            if (should_rotate) quad_wedge(P, -face_angle, knorm[1]);
        }
        if (should_rotate) wedge<FRZ>(P, G, -face_angle, knorm[0]);
    } else {
        if (should_rotate) wedge<FRZ>(P, G, -face_angle, knorm[0]);
        if (model == 1) {
            if (should_rotate) quad_wedge(P, -face_angle, knorm[1]);
            mult_fringe<FRZ>(P, G, knorm, kskew, 3, knl, ksl, nkl - 1, length_eff, is_exit, 1);
        }
        dipole_fringe<FRZ>(P, G, fint, hgap, k0);
        if (should_rotate) yrotation<FRZ>(P, G, -sin_, cos_, -tan_);
    }
}

// DipoleEdgeNonLinear_single_particle, track_dipole_edge_nonlinear.h:12-44
// q: [0] k [1] e1 [2] fint [3] hgap [4] sin [5] cos [6] tan (or -999); aux = side
template <bool FRZ>
__device__ __noinline__ void dipole_edge_nonlinear(PState& P, const PSlot& G, const double* __restrict__ q,
                                                   const int32_t side) {
    const double k = q[0], e1 = q[1], fint = q[2], hgap = q[3], sin_ = q[4], cos_ = q[5], tan_ = q[6];
    if (side == 0) {
        if (sin_ > -99.) yrotation<FRZ>(P, G, -sin_, cos_, -tan_);
        dipole_fringe<FRZ>(P, G, fint, hgap, k);
        if (sin_ > -99.) wedge<FRZ>(P, G, -e1, k);
    } else if (side == 1) {
        if (sin_ > -99.) wedge<FRZ>(P, G, -e1, k);
        dipole_fringe<FRZ>(P, G, fint, hgap, -k);
        if (sin_ > -99.) yrotation<FRZ>(P, G, -sin_, cos_, -tan_);
    }
}

template <bool SYNRAD, bool FRZ>
__device__ __noinline__ void heavy_op(const uint32_t op, const int32_t aux, const double* __restrict__ q,
                                      PState& P, const PSlot& G, const XtbTrackArgs& a) {
    switch (op) {
    case XTB_OP_MAGNET_BODY: magnet_body<SYNRAD, FRZ>(P, G, a, q, aux); break;
    case XTB_OP_MAGNET_EDGE: magnet_edge<FRZ>(P, G, q, aux); break;
    case XTB_OP_DIPEDGE_NL: dipole_edge_nonlinear<FRZ>(P, G, q, aux); break;
    default: break;
    }
}
