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
#include "xtb_math.cuh"
#include "xtb_rng.cuh"

#define XTB_QELEM 1.60217662e-19
#define XTB_EPSILON_0 8.854187817620e-12
#define XTB_POW2(X) ((X) * (X))
#define XTB_POW3(X) ((X) * (X) * (X))
#define XTB_POW4(X) ((X) * (X) * (X) * (X))

// ---------------------------------------------------------------- drifts ----
// Explicit fused multiply-add (one rounding) in both builds.
#ifdef __CUDA_ARCH__
#define XTB_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define XTB_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

// a / b when rb = RN(1 / b) is already at hand (an element constant tabulated by the host,
// the 1/pz the reference computes anyway, the 1/rvv every state carries).  This is the
// closing correction step of every FMA-based IEEE division (Markstein 1990; the tail of the
// CUDA double-precision division itself): q = RN(a*rb), r = a - b*q exactly (FMA),
// result = RN(q + r*rb) = RN(a / b).  3 FP64 instructions instead of ~10 and no slow-path
// test.  Checked against `a / b` on 2*10^9 random operand pairs of the ranges met here
// (b ~ cos(small angle), pz ~ 1 + delta, generic exponents +-20): no mismatch.
__device__ __forceinline__ double div_by(const double a, const double b, const double rb) {
    const double q = a * rb;
    const double r = XTB_FMA(-b, q, a);
    return XTB_FMA(r, rb, q);
}

// Trigonometry of element constants.  cos(h*s), sin(h*s), sin(h*s/2) in the polar drift and
// the curved exact bend depend on the element alone, yet the reference evaluates them per
// particle and per integrator sub-step (a default bend: 32 polar drifts = 96 sin/cos per
// particle).  The host lowering tabulates them (lowering.py::_trig_table, computed with the
// host libm -- the one the reference's CPU build uses) for every sub-step length the
// integrator of this op asks for, in the order the integrator asks: entry
// `outer class * n_inner + inner class` (see magnet_drift_n / magnet_body_n), so the device
// indexes, it does not search.  The entry carries the length it was made for; if the bit
// pattern does not match (there should be no such case; the host test build counts them)
// the values are evaluated here.
#define XTB_TRIG_STRIDE 6   // [length, cos(h*s), sin(h*s), sin(h*s/2), RN(1/cos(h*s)), 0]
struct TrigTab {
    // block in the op's parameters: [(int) n | n_inner << 32, 1/h, n entries]; read on demand
    // (shared-memory broadcast loads) instead of being carried in registers through the body
    const double* base;
    __device__ __forceinline__ int n() const {
        return (int) ((unsigned long long) __double_as_longlong(base[0]) & 0xffffffffu);
    }
    // inner classes per outer class (1: models 2, 4; 2: model 7; 4: model 8)
    __device__ __forceinline__ int n_inner() const {
        return (int) ((unsigned long long) __double_as_longlong(base[0]) >> 32);
    }
    __device__ __forceinline__ double rho() const { return base[1]; }      // 1 / h (iff n > 0)
    __device__ __forceinline__ const double* entry(const int idx) const {
        return base + 2 + XTB_TRIG_STRIDE * idx;
    }
};
#ifdef XTB_COUNT_TRIG_MISS
static long long xtb_trig_lookups = 0, xtb_trig_misses = 0;
#endif
__device__ __forceinline__ void trig_of(const TrigTab& tt, const int idx, const double h,
                                        const double s, double& ca, double& sa, double& sa2,
                                        double& rca) {
#ifdef XTB_COUNT_TRIG_MISS
    xtb_trig_lookups++;
#endif
    if (idx < tt.n()) {
        const double* e = tt.entry(idx);
        if (__double_as_longlong(e[0]) == __double_as_longlong(s)) {
            ca = e[1];  sa = e[2];  sa2 = e[3];  rca = e[4];
            return;
        }
    }
#ifdef XTB_COUNT_TRIG_MISS
    xtb_trig_misses++;
#endif
    ca = xtb_cos_glibc(h * s);
    sa = xtb_sin_glibc(h * s);
    sa2 = xtb_sin_glibc(0.5 * h * s);
    rca = 1. / ca;
}

// Step tables of the nested Yoshida bends (track_magnet_drift.h:521-550) and of the
// Yoshida-4 integrator (track_magnet.h:236-250): [drift fractions..., kick fractions...]
#ifdef __CUDACC__
#define XTB_CONST_TABLE static __device__ __constant__ double
#else
#define XTB_CONST_TABLE static const double
#endif
XTB_CONST_TABLE XTB_Y4_NESTED[8] = {0.6756035959798289, -0.17560359597982889, -0.17560359597982889,
                                    0.6756035959798289,
                                    1.3512071919596578, -1.7024143839193155, 1.3512071919596578, 0.};
XTB_CONST_TABLE XTB_Y6[16] = {
    3.922568052387799819591407413100e-01, 5.100434119184584780271052295575e-01,
    -4.710533854097565531482416645304e-01, 6.875316825251809316199569366290e-02,
    6.875316825251809316199569366290e-02, -4.710533854097565531482416645304e-01,
    5.100434119184584780271052295575e-01, 3.922568052387799819591407413100e-01,
    7.845136104775599639182814826199e-01, 2.355732133593569921359289764951e-01,
    -1.177679984178870098432412305556e+00, 1.315186320683906284756403692882e+00,
    -1.177679984178870098432412305556e+00, 2.355732133593569921359289764951e-01,
    7.845136104775599639182814826199e-01, 0.};

// track_polar_drift_single_particle, track_magnet_drift.h:45-87, on N particles at once
// (N independent dependency chains for the FP64 pipe: straight-line code, the square root,
// reciprocals and division being the guard-free fast paths of xtb_math.cuh).  The divisions by
// cos(h*s), pz and rvv use the reciprocals at hand (div_by): same correctly rounded quotients.
template <int N, bool FRZ>
__device__ __forceinline__ void polar_drift_n(PState (&P)[N], const double length, const double h,
                                              const TrigTab& tt, const int idx, const double rho) {
    const double s = length;
    double ca, sa, sa2, rca;
    trig_of(tt, idx, h, s, ca, sa, sa2, rca);
    // (each statement of the reference's map, for all N particles in turn: see XTB_LANES)
    double opd[N], pz2[N], pz[N], ipz[N], pxt[N], dtt[N], iptt[N], xr[N];
    XTB_LANES opd[k] = P[k].delta + 1.0;
    XTB_LANES pz2[k] = XTB_POW2(opd[k]) - XTB_POW2(P[k].px) - XTB_POW2(P[k].py);
    xtb_vsqrt<N>(pz, pz2);
    xtb_vrcp<N>(ipz, pz);                              // _pz = 1 / pz
    XTB_LANES pxt[k] = P[k].px * ipz[k];
    XTB_LANES dtt[k] = ca - sa * pxt[k];
    xtb_vrcp<N>(iptt, dtt);                            // _ptt = 1 / (ca - sa * pxt)
    XTB_LANES xr[k] = P[k].x + rho;
    double pst[N], new_x[N], new_px[N], new_y[N], num[N], den[N], dell[N];
    XTB_LANES pst[k] = xr[k] * sa * ipz[k] * iptt[k];
    XTB_LANES new_x[k] = (P[k].x + rho * (2 * sa2 * sa2 + sa * pxt[k])) * iptt[k];
    XTB_LANES new_px[k] = ca * P[k].px + sa * pz[k];
    XTB_LANES new_y[k] = P[k].y + pst[k] * P[k].py;
    // delta_ell = one_plus_delta * (x + rho) * sa / ca / pz / (1 - px * sa / ca / pz)
    XTB_LANES num[k] = opd[k] * xr[k] * sa;
    XTB_LANES den[k] = P[k].px * sa;
    XTB_LANES num[k] = div_by(num[k], ca, rca);
    XTB_LANES den[k] = div_by(den[k], ca, rca);
    XTB_LANES num[k] = div_by(num[k], pz[k], ipz[k]);
    XTB_LANES den[k] = 1 - div_by(den[k], pz[k], ipz[k]);
    xtb_vdiv<N>(dell, num, den);
    XTB_LANES {
        P[k].x = new_x[k];
        P[k].px = new_px[k];
        P[k].y = new_y[k];
    }
    if (!FRZ) {
        XTB_LANES dell[k] = div_by(dell[k], P[k].rvv, P[k].rv0v);
        XTB_LANES P[k].zeta += length - dell[k];
        XTB_LANES P[k].s += s;
    }
}
template <bool FRZ>
__device__ __noinline__ void polar_drift(PState& P, const double length, const double h,
                                         const TrigTab& tt, const int idx) {
    polar_drift_n<1, FRZ>(reinterpret_cast<PState(&)[1]>(P), length, h, tt, idx,
                          (tt.n() > 0) ? tt.rho() : 1 / h);
}

// S = sin(sqrt(K) L) / sqrt(K), C = cos(sqrt(K) L) for K > 0, the hyperbolic pair for K < 0,
// (L, 1) for K == 0: track_magnet_drift.h:121-147, once per plane.  Out of line, one copy: the
// four libm bodies inlined twice made the quadrupole map miss the instruction cache (ncu:
// 31 % of its samples were no_instruction).
static __device__ __noinline__ void focusing_terms(const double K, const double length, double& S,
                                                   double& C) {
    if (K > 0.0) {
        const double sqrt_K = sqrt(K);
        S = xtb_sin_glibc(sqrt_K * length) / sqrt_K;
        C = xtb_cos_glibc(sqrt_K * length);
    } else if (K < 0.0) {
        const double sqrt_K = sqrt(-K);
        S = xtb_sinh_glibc(sqrt_K * length) / sqrt_K;
        C = xtb_cosh_glibc(sqrt_K * length);
    } else {
        S = length;
        C = 1.0;
    }
}

// track_expanded_combined_dipole_quad_single_particle, track_magnet_drift.h:91-213.
// Its 13 divisions by 1 + delta, Kx and Ky share three reciprocals (div_by: same correctly
// rounded quotients; `t / (2 K)` is `(t / K) / 2` exactly).
template <bool FRZ>
__device__ __noinline__ void combined_dipole_quad(PState& P, const double length, const double k0_,
                                                  const double k1_, const double h) {
    const double x = P.x, y = P.y, px = P.px, py = P.py, rvv = P.rvv;
    const double delta_plus_1 = P.delta + 1;
    const double r_dp1 = xtb_rcp(delta_plus_1);
    const double chi = P.chi;
    const double k0 = div_by(chi * k0_, delta_plus_1, r_dp1);
    const double k1 = div_by(chi * k1_, delta_plus_1, r_dp1);
    const double Kx = k0 * h + k1;
    const double Ky = -k1;
    double Sx, Sy, Cx, Cy;
    focusing_terms(Kx, length, Sx, Cx);
    focusing_terms(Ky, length, Sy, Cy);
    const double xp = div_by(px, delta_plus_1, r_dp1);
    const double yp = div_by(py, delta_plus_1, r_dp1);
    const double A = -Kx * x - k0 + h;
    const double B = xp;
    const double C = -Ky * y;
    const double D = yp;
    double x_ = x * Cx + xp * Sx;
    const double y_ = y * Cy + yp * Sy;
    const double px_ = (A * Sx + B * Cx) * delta_plus_1;
    const double py_ = (C * Sy + D * Cy) * delta_plus_1;
    double length_ = length;
    if (Kx != 0.0) {
        const double r_Kx = xtb_rcp(Kx);
        x_ = x_ + div_by((k0 - h) * (Cx - 1.0), Kx, r_Kx);
        length_ -= div_by(h * ((Cx - 1.0) * xp + Sx * A + length * (k0 - h)), Kx, r_Kx);
        length_ += 0.5 * (div_by(-(XTB_POW2(A) * Cx * Sx), Kx, r_Kx) * 0.5 + (XTB_POW2(B) * Cx * Sx) / 2.0
                          + div_by(XTB_POW2(A) * length, Kx, r_Kx) * 0.5 + (XTB_POW2(B) * length) / 2.0
                          - div_by(A * B * XTB_POW2(Cx), Kx, r_Kx) + div_by(A * B, Kx, r_Kx));
    } else {
        x_ = x_ - (k0 - h) * 0.5 * XTB_POW2(length);
        length_ += h * length * (3.0 * length * xp + 6.0 * x - (k0 - h) * XTB_POW2(length)) / 6.0;
        length_ += 0.5 * (XTB_POW2(B)) * length;
    }
    if (Ky != 0.0) {
        const double r_Ky = xtb_rcp(Ky);
        length_ += 0.5 * (div_by(-(XTB_POW2(C) * Cy * Sy), Ky, r_Ky) * 0.5 + (XTB_POW2(D) * Cy * Sy) / 2.0
                          + div_by(XTB_POW2(C) * length, Ky, r_Ky) * 0.5 + (XTB_POW2(D) * length) / 2.0
                          - div_by(C * D * XTB_POW2(Cy), Ky, r_Ky) + div_by(C * D, Ky, r_Ky));
    } else {
        length_ += 0.5 * XTB_POW2(D) * length;
    }
    const double dzeta = length - div_by(length_, rvv, P.rv0v);
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
                                               const double h, const TrigTab& tt, const int idx) {
    const double k0_chi = k0 * P.chi;
    if (fabs(k0_chi) < 1e-8) {
        polar_drift<FRZ>(P, length, h, tt, idx);
        return;
    }
    const double rvv = P.rvv;
    const double x0 = P.x, y0 = P.y, px0 = P.px, py = P.py;
    const double s = length;
    const double one_plus_delta = P.delta + 1.0;
    const double hs = h * s;
    // (sin(hs / 2) == sin(0.5 * h * s): halving is exact)
    double cos_hs, sin_hs, sin_hs_2, rcos_hs;
    trig_of(tt, idx, h, s, cos_hs, sin_hs, sin_hs_2, rcos_hs);
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

// track_magnet_drift_single_particle, track_magnet_drift.h:468-555, on N particles.
// `oc` = class of the outer (integrator) sub-step, for the trig table index.
template <int N, bool FRZ>
__device__ __forceinline__ void magnet_drift_n(PState (&P)[N], const double length, const double k0,
                                               const double k1, const double h, const int drift_model,
                                               const TrigTab& tt, const int oc) {
    if (drift_model == -1) return;
    if (length == 0.0) return;
    switch (drift_model) {
    case 0:
#pragma unroll
        for (int k = 0; k < N; ++k) drift_expanded<FRZ>(P[k], length);
        break;
    case 1:
#pragma unroll
        for (int k = 0; k < N; ++k) drift_exact<FRZ>(P[k], length);
        break;
    // (these out-of-line maps take the lane by address, which keeps the lane array of the
    // caller in thread-local memory.  Measured, r02 session 15: handing them a copy instead --
    // lanes in registers, 600-900 bytes of spills under the 128-register cap -- is SLOWER:
    // LEP thick 1.61e10 -> 1.48e10 PET/s, CLIC-DR mean 1.08e11 -> 0.81e11)
    case 3:
        for (int k = 0; k < N; ++k) combined_dipole_quad<FRZ>(P[k], length, k0, k1, h);
        break;
    case 4:
        for (int k = 0; k < N; ++k) curved_exact_bend<FRZ>(P[k], length, k0, h, tt, oc);
        break;
    case 5:
        for (int k = 0; k < N; ++k) straight_exact_bend<FRZ>(P[k], length, k0);
        break;
    case 2:
    case 7:
    case 8: {
        // polar drift (2) and the nested Yoshida-4 / Yoshida-6 bends (7, 8:
        // track_magnet_drift.h:521-550) as ONE loop around ONE inlined polar drift: the
        // state stays in registers for the whole body, the code is there once
        const int n_in = (drift_model == 2) ? 1 : ((drift_model == 7) ? 4 : 8);
        const double* __restrict__ tab = (drift_model == 7) ? XTB_Y4_NESTED : XTB_Y6;
        const int n_inner = tt.n_inner();                          // (read once per call)
        const double rho = (tt.n() > 0) ? tt.rho() : 1 / h;
#pragma unroll 1
        for (int j = 0; j < n_in; ++j) {
            const double lj = (n_in == 1) ? length : tab[j] * length;
            const int ic = (j < n_in - 1 - j) ? j : n_in - 1 - j;
            polar_drift_n<N, FRZ>(P, lj, h, tt, oc * n_inner + ic, rho);
            if (j < n_in - 1) {
                const double kj = tab[n_in + j];
#pragma unroll
                for (int k = 0; k < N; ++k) P[k].px = P[k].px - kj * k0 * P[k].chi * length;
            }
        }
        break;
    }
    default: break;
    }
}

// ------------------------------------------------------------ body params ----
// OP_MAGNET_BODY parameter block (lowering.py::_lower_magnet):
//  q[0] length  q[1] k0_drift (kick-only bodies: 1 / length)  q[2] k1_drift  q[3] h_drift  q[4] h_kick  q[5] hxl
//  q[6] A0 = k0_h_correction*length + k0l   q[7] A1 = k1_h_correction*length + k1l
//  q[8] htot    q[9] (int) order_user | order_rel << 32
//  q[10..17] k0_tot k1_tot k2 k3 k0s k1s k2s k3s   (field evaluation for radiation)
//  q[18..25] main coefficients (order 3, Horner order, pairs)
//  then user coefficients (order_user+1 pairs), then rel coefficients (order_rel+1 pairs),
//  then the trig table: (int) n | n_inner << 32, 1/h, n x XTB_TRIG_STRIDE doubles
//  then, if the element's linear edges were merged in (lowering.py::_merge_linear_edges):
//  r21, r43 of the entry edge, r21, r43 of the exit edge (track_dipole_edge_linear.h:30-39)
//  aux: integrator[0:2] drift_model+1[2:6] rot_frame[6] has_user[7] has_rel[8]
//       has_main[9] radiation_flag[10:12] drift_only[12] edge_in[13] edge_out[14]
//       num_kicks[15:32]
struct BodyPar {
    // only the parameter pointer and the flag word are carried; everything else is decoded
    // where it is used (a few integer instructions against ~30 registers held through the
    // whole body, which kept ptxas from interleaving the lanes)
    const double* q;
    uint32_t a;
    __device__ __forceinline__ int order_user() const {
        return (int) ((unsigned long long) __double_as_longlong(q[9]) & 0xffffffffu);
    }
    __device__ __forceinline__ int order_rel() const {
        return (int) ((unsigned long long) __double_as_longlong(q[9]) >> 32);
    }
    __device__ __forceinline__ const double* cm() const { return q + 18; }
    __device__ __forceinline__ const double* cu() const { return q + 26; }
    __device__ __forceinline__ const double* cr() const { return q + 26 + 2 * (order_user() + 1); }
    __device__ __forceinline__ TrigTab trig() const {
        TrigTab tt;
        tt.base = cr() + 2 * (order_rel() + 1);
        return tt;
    }
    __device__ __forceinline__ const double* edges() const {     // [r21_in, r43_in, r21_out, r43_out]
        const TrigTab tt = trig();
        return tt.entry(tt.n());
    }
    __device__ __forceinline__ int integrator() const { return a & 3; }
    __device__ __forceinline__ int drift_model() const { return (int) ((a >> 2) & 15) - 1; }
    __device__ __forceinline__ int rot_frame() const { return (a >> 6) & 1; }
    __device__ __forceinline__ int has_user() const { return (a >> 7) & 1; }
    __device__ __forceinline__ int has_rel() const { return (a >> 8) & 1; }
    __device__ __forceinline__ int has_main() const { return (a >> 9) & 1; }
    __device__ __forceinline__ int radiation_flag() const { return (a >> 10) & 3; }
    __device__ __forceinline__ int drift_only() const { return (a >> 12) & 1; }
    __device__ __forceinline__ int edge_in() const { return (a >> 13) & 1; }
    __device__ __forceinline__ int edge_out() const { return (a >> 14) & 1; }
    __device__ __forceinline__ int num_kicks() const { return (int) (a >> 15); }
};

__device__ __forceinline__ BodyPar body_par(const double* q, const int32_t aux) {
    BodyPar b;
    b.q = q;
    b.a = (uint32_t) aux;
    return b;
}

template <int N, bool FRZ>
__device__ __forceinline__ void magnet_kick_curv_n(PState (&P)[N], const BodyPar& b, const double kick_weight);

// track_magnet_kick_single_particle, track_magnet_kick.h:24-144, on N particles
template <int N, bool FRZ>
__device__ __forceinline__ void magnet_kick_n(PState (&P)[N], const BodyPar& b, const double kick_weight) {
    const double length = b.q[0];
    if (b.has_user()) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double m, n;
            horner_kick(P[k].x, P[k].y, P[k].chi, b.cu(), b.order_user(), m, n);
            P[k].px += kick_weight * (-m);
            P[k].py += kick_weight * n;
        }
    }
    if (b.has_rel()) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double m, n;
            horner_kick(P[k].x, P[k].y, P[k].chi, b.cr(), b.order_rel(), m, n);
            P[k].px += kick_weight * (-m);
            P[k].py += kick_weight * n;
        }
    }
    if (b.has_main()) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double m, n;
            // (order 3 whatever the zero pattern: trimming the leading zero orders with a
            // run-time order measured 1.5 % SLOWER than the unrolled constant-order loop)
            horner_kick(P[k].x, P[k].y, P[k].chi, b.cm(), 3, m, n);
            P[k].px += kick_weight * (-m);
            P[k].py += kick_weight * n;
        }
    }
    magnet_kick_curv_n<N, FRZ>(P, b, kick_weight);
}

// ... its curvature terms (track_magnet_kick.h:98-142)
template <int N, bool FRZ>
__device__ __forceinline__ void magnet_kick_curv_n(PState (&P)[N], const BodyPar& b, const double kick_weight) {
    const double length = b.q[0];
    const double h = b.q[4], hxl = b.q[5];
    const double htot = b.q[8];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double chi = P[k].chi, x = P[k].x, y = P[k].y;
        double dpx = 0, dpy = 0, dzeta = 0;
        if (b.rot_frame()) {
            const double hl = h * length * kick_weight + hxl * kick_weight;
            dpx += hl * (1. + P[k].delta);
            dzeta += -P[k].rv0v * hl * x;
        }
        dpx += -chi * b.q[6] * kick_weight * htot * x;
        dpx += htot * chi * b.q[7] * kick_weight * (-x * x + 0.5 * y * y);
        dpy += htot * chi * b.q[7] * kick_weight * x * y;
        P[k].px += dpx;
        P[k].py += dpy;
        if (!FRZ) P[k].zeta += dzeta;
    }
}

// ------------------------------------------------------------- radiation ----
struct RadCtx {      // per-call context: rng state + failure flag
    Rng r;
    Philox ph;
    bool philox;
    bool seeded;
    bool rng_error;
};

__device__ __forceinline__ uint32_t rng_next_u32(RadCtx& c) {
    if (!c.philox) return rng_u32(c.r);
    const uint32_t w = c.r.s3 & 3u;
    if (!c.ph.have || w == 0u) {
        // block number = draw counter >> 2 (62 bits)
        philox4x32_10(c.r.s1, c.r.s2, (c.r.s3 >> 2) | (c.r.s4 << 30), c.r.s4 >> 2, c.ph.blk);
        c.ph.have = true;
    }
    const uint32_t v = c.ph.blk[w];
    c.r.s3 += 1u;
    if (c.r.s3 == 0u) c.r.s4 += 1u;
    return v;
}

// RandomUniform_generate, random_src/uniform.h:34-53
__device__ __forceinline__ double rand_uniform(RadCtx& c) {
    if (!c.seeded) { c.rng_error = true;  return 0; }
    return rng_next_u32(c) / 4294967296.0;
}
__device__ __forceinline__ uint32_t rand_u32(RadCtx& c) {
    if (!c.seeded) { c.rng_error = true;  return 0; }
    return rng_next_u32(c);
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

// SynRad, headers/synrad_spectrum.h:80-175 (Chebyshev series from H.Burkhardt).  The
// reference writes each series out as a chain `a = z*b - a + c_k; b = z*a - b + c_k+1; ...`:
// t_k = z*t_(k-1) - t_(k-2) + c_k with t_0 = c_0 (t_(-1) = 0: `z*t_0 - 0` is `z*t_0` exactly)
// and a last step with z/2.  Here the coefficients sit in constant memory and the chain is a
// loop -- same operations in the same order -- because written out, with two 32-bit moves per
// 64-bit literal, the three series were a third of the radiation kernel's instructions and
// the kernel was bound by instruction fetch (ncu: 65 % of its samples `no_instruction`).
XTB_CONST_TABLE XTB_SYNRAD_P[19] = {
    .00000000000000000012, .00000000000000000460, .00000000000000031738, .00000000000002004426,
    .00000000000111455474, .00000000005407460944, .00000000226722011790, .00000008125130371644,
    .00000245751373955212, .00006181256113829740, .00127066381953661690, .02091216799114667278,
    .26880346058164526514, 2.61902183794862213818, 18.65250896865416256398, 92.95232665922707542088,
    308.15919413131586030542, 644.86979658236221700714, 414.56543648832546975110};
XTB_CONST_TABLE XTB_SYNRAD_Q[18] = {
    .00000000000000000004, .00000000000000000289, .00000000000000019786, .00000000000001196168,
    .00000000000063427729, .00000000002923635681, .00000000115951672806, .00000003910314748244,
    .00000110599584794379, .00002581451439721298, .00048768692916240683, .00728456195503504923,
    .08357935463720537773, .71031361199218887514, 4.26780261265492264837, 17.05540785795221885751,
    41.83903486779678800040, 28.41787374362784178164};
XTB_CONST_TABLE XTB_SYNRAD_R[30] = {
    .00000000000000000001, -.00000000000000000002, .00000000000000000006, -.00000000000000000020,
    .00000000000000000066, -.00000000000000000216, .00000000000000000721, -.00000000000000002443,
    .00000000000000008441, -.00000000000000029752, .00000000000000107116, -.00000000000000394564,
    .00000000000001489474, -.00000000000005773537, .00000000000023030657, -.00000000000094784973,
    .00000000000403683207, -.00000000001785432348, .00000000008235329314, -.00000000039817923621,
    .00000000203088939238, -.00000001101482369622, .00000006418902302372, -.00000040756144386809,
    .00000287536465397527, -.00002321251614543524, .00022505317277986004, -.00287636803664026799,
    .06239591359332750793, 1.06552390798340693166};
__device__ __forceinline__ double synrad_series(const double* __restrict__ c, const int n, const double z) {
    double t2 = 0., t1 = c[0];
#pragma unroll 1
    for (int k = 1; k < n - 1; ++k) {
        const double t = z * t1 - t2 + c[k];
        t2 = t1;
        t1 = t;
    }
    return .5 * z * t1 - t2 + c[n - 1];
}

static __device__ __noinline__ double synrad_fn(const double x) {
    double synrad = 0.;
    if (x > 0. && x < 800.) {
        if (x < 6.) {
            const double z = x * x / 16. - 2.;
            const double p = synrad_series(XTB_SYNRAD_P, 19, z);
            const double q = synrad_series(XTB_SYNRAD_Q, 18, z);
            const double y = pow(x, 2. / 3.);
            synrad = (p / y - q * y - 1.) * 1.81379936423421784215530788143;
        } else {
            const double z = 20. / x - 2.;
            const double p = synrad_series(XTB_SYNRAD_R, 30, z);
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
    c.philox = a.rng_philox != 0;
    c.ph.have = false;

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

// ---- quantum-kick model (radiation_flag 3): the TOTAL energy radiated in a slice ----------
// synrad_spectrum.h:220-459.  The number of photons is drawn (Poisson), the sum of their
// energies comes from tabulated inverse CDFs of the sum of N photon energies (N = 1..32 and
// 64, 128, 256; headers/_generate_synrad_total_energy_tables.py) -- a few table look-ups
// instead of hundreds of rejection-sampled photons.  The tables are data, uploaded with the
// lattice (xtb_lattice_set_synrad_tables); layout of the blob, in doubles:
//   [0] n_left  [1] n_center  [2] n_right  [3] tail probability max  [4] direct table max (32)
//   [5..7] reserved, then the left u grid, the centre u grid, the right v grid, then the
//   tables log(X_N) for N = 1..32, 64, 128, 256, each n_left + n_center + n_right long.
struct QkTables {
    const double* grid_l;
    const double* grid_c;
    const double* grid_r;
    const double* tables;
    int n_l, n_c, n_r, direct_max;
    double tail_max;
};
__device__ __forceinline__ QkTables qk_tables(const double* __restrict__ blob) {
    QkTables t;
    t.n_l = (int) blob[0];  t.n_c = (int) blob[1];  t.n_r = (int) blob[2];
    t.tail_max = blob[3];
    t.direct_max = (int) blob[4];
    t.grid_l = blob + 8;
    t.grid_c = t.grid_l + t.n_l;
    t.grid_r = t.grid_c + t.n_c;
    t.tables = t.grid_r + t.n_r;
    return t;
}
// table of the sum of n photons (n <= direct_max, or 64 / 128 / 256)
__device__ __forceinline__ const double* qk_table_of(const QkTables& t, const int64_t n) {
    const int64_t size = (int64_t) t.n_l + t.n_c + t.n_r;
    int64_t idx;
    if (n <= t.direct_max) idx = n - 1;
    else if (n == 64) idx = t.direct_max;
    else if (n == 128) idx = t.direct_max + 1;
    else idx = t.direct_max + 2;
    return t.tables + idx * size;
}

// synrad_gen_photon_count, synrad_spectrum.h:220-245
static __device__ __noinline__ int64_t synrad_gen_photon_count(RadCtx& c, const double average_nphotons) {
    if (average_nphotons <= 0.0) return 0;
    if (average_nphotons < 700.0) {
        const double threshold = exp(-average_nphotons);
        double product = 1.0;
        int64_t nphot = -1;
        do {
            nphot++;
            product *= rand_uniform(c);
        } while (product > threshold && !c.rng_error);
        return nphot;
    }
    double n = rand_exponential(c);
    int64_t nphot = 0;
    while (n < average_nphotons && !c.rng_error) {
        nphot++;
        n += rand_exponential(c);
    }
    return nphot;
}

// synrad_total_energy_find_grid_index_direct_segment, synrad_spectrum.h:257-288
__device__ __forceinline__ int64_t qk_find_index(const double value, const double* __restrict__ grid,
                                                 const int64_t size, const bool is_log_spaced) {
    if (value <= grid[0]) return 0;
    if (value >= grid[size - 1]) return size - 2;
    if (is_log_spaced) {
        if (value < grid[1]) return 0;
        const double position = ((log(value) - log(grid[1])) / (log(grid[size - 1]) - log(grid[1]))
                                 * (size - 2));
        int64_t i_low = 1 + (int64_t) floor(position);
        if (i_low < 0) i_low = 0;
        if (i_low > size - 2) i_low = size - 2;
        return i_low;
    }
    const double du = (grid[size - 1] - grid[0]) / (size - 1);
    int64_t i_low = (int64_t) floor((value - grid[0]) / du);
    if (i_low < 0) i_low = 0;
    if (i_low > size - 2) i_low = size - 2;
    return i_low;
}
// synrad_total_energy_interpolate_segment, synrad_spectrum.h:291-325
__device__ __forceinline__ double qk_interp_segment(const double value, const double* __restrict__ grid,
                                                    const double* __restrict__ table, const int64_t offset,
                                                    const int64_t size, const bool is_log_spaced) {
    if (value <= grid[0]) return table[offset];
    if (value >= grid[size - 1]) return table[offset + size - 1];
    int64_t i_low = qk_find_index(value, grid, size, is_log_spaced);
    if (i_low < 0) i_low = 0;
    if (i_low >= size - 1) i_low = size - 2;
    const int64_t i_high = i_low + 1;
    const double u_low = grid[i_low], u_high = grid[i_high];
    double w;
    if (is_log_spaced && u_low > 0.0 && value > 0.0) w = (log(value) - log(u_low)) / (log(u_high) - log(u_low));
    else w = (value - u_low) / (u_high - u_low);
    return (1.0 - w) * table[offset + i_low] + w * table[offset + i_high];
}
// synrad_gen_total_energy_normalized_from_log_table, synrad_spectrum.h:327-380
__device__ __forceinline__ double qk_sample(RadCtx& c, const QkTables& t, const double* __restrict__ table) {
    const double u = rand_uniform(c);
    double log_value;
    if (u < t.tail_max) {
        log_value = qk_interp_segment(u, t.grid_l, table, 0, t.n_l, true);
    } else if (u <= 1.0 - t.tail_max) {
        log_value = qk_interp_segment(u, t.grid_c, table, t.n_l, t.n_c, false);
    } else {
        const double v = 1.0 - u;
        log_value = qk_interp_segment(v, t.grid_r, table, (int64_t) t.n_l + t.n_c, t.n_r, true);
    }
    return exp(log_value);
}

// synrad_emit_total_energy_loss, synrad_spectrum.h:408-459
template <bool FRZ>
__device__ __noinline__ void synrad_emit_total_energy_loss(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                                           const double B_T, const double lpath) {
    if (fabs(B_T) < 1e-4) return;
    const double mass0 = a.part.mass0;
    const double q0 = a.part.q0;
    const double gamma0 = G.ld(F_GAMMA0);
    const double beta0 = G.ld(F_BETA0);

    RadCtx c;
    c.r.s1 = G.ldu(F_RNG_S1);  c.r.s2 = G.ldu(F_RNG_S2);
    c.r.s3 = G.ldu(F_RNG_S3);  c.r.s4 = G.ldu(F_RNG_S4);
    c.seeded = !(c.r.s1 == 0 && c.r.s2 == 0 && c.r.s3 == 0 && c.r.s4 == 0);
    c.rng_error = false;
    c.philox = a.rng_philox != 0;
    c.ph.have = false;

    const double n_avg = synrad_average_number_of_photons(mass0, q0, beta0 * gamma0, B_T, lpath);
    const int64_t nphot = synrad_gen_photon_count(c, n_avg);
    double total = 0.0;
    if (nphot > 0 && !c.rng_error) {
        const QkTables t = qk_tables(a.synrad_tables);
        int64_t left = nphot;
        while (left > t.direct_max && !c.rng_error) {
            int64_t chunk = 1;                      // synrad_largest_power_of_two_leq
            while (chunk <= left / 2) chunk *= 2;
            if (chunk > 256) chunk = 256;
            total += qk_sample(c, t, qk_table_of(t, chunk));
            left -= chunk;
        }
        if (left > 0) total += qk_sample(c, t, qk_table_of(t, left));
    }
    G.stu(F_RNG_S1, c.r.s1);  G.stu(F_RNG_S2, c.r.s2);
    G.stu(F_RNG_S3, c.r.s3);  G.stu(F_RNG_S4, c.r.s4);
    if (c.rng_error) {       // RNG_ERR_SEEDS_NOT_SET, uniform.h:40-43
        kill_particle<FRZ>(P, G, -20);
        return;
    }
    if (nphot == 0) return;

    const double Q0_coulomb = fabs(q0) * XTB_QELEM;
    const double mass0_kg = mass0 * XTB_QELEM / XTB_C_LIGHT / XTB_C_LIGHT;
    const double P0_J = mass0_kg * beta0 * gamma0 * XTB_C_LIGHT;
    const double curv = B_T / P0_J * Q0_coulomb;
    const double gamma = gamma0 * (1 + P.delta);
    const double p0c = G.ld(F_P0C);
    const double initial_energy = sqrt(p0c * p0c + mass0 * mass0) + G.ld(F_PTAU) * p0c;
    const double c1 = 1.5 * 1.973269804593025e-07;
    const double energy_critical = c1 * (gamma * gamma * gamma0) * curv;
    double energy_loss_total = total * energy_critical;
    if (energy_loss_total >= initial_energy) energy_loss_total = initial_energy;
    if (energy_loss_total >= initial_energy) {
        P.state = -10;       // XT_LOST_ALL_E_IN_SYNRAD
    } else {
        const double energy = initial_energy - energy_loss_total;
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
    if (b.has_user()) { horner_kick(x, y, 1., b.cu(), b.order_user(), m, n);  dpx_mul = -m;  dpy_mul = n; }
    if (b.has_rel()) { horner_kick(x, y, 1., b.cr(), b.order_rel(), m, n);  dpx_rel = -m;  dpy_rel = n; }
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
    if (!(b.radiation_flag() && length > 0)) return;   // spin is (0,0,0): magnet_spin is a no-op
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
    if (b.radiation_flag() == 1) {
        synrad_average_kick<FRZ>(P, G, a, B_perp_T, l_path);
    } else if (b.radiation_flag() == 2) {
        synrad_emit_photons<FRZ>(P, G, a, B_perp_T, l_path);
    } else if (b.radiation_flag() == 3) {
        synrad_emit_total_energy_loss<FRZ>(P, G, a, B_perp_T, l_path);
    }
}

// ---- thin radiating multipole: the whole element on N particles at once ----------------------
// A thin ring with radiation (CLIC-DR: 8 596 wiggler / bend / quadrupole slices per turn) is made
// of kick-only bodies (model -1, one uniform step without drifts) wrapped by WITH_RADIATION.
// Through the generic body they cost ~750 instructions per particle (ncu, profiles/
// r02_history.md): three Horner evaluations for the field where one has non-zero coefficients,
// out-of-line calls per particle for rad_end / synrad_average_kick / update_delta, a dozen
// guarded divisions and three guarded square roots, one dependency chain at a time.  This is
// the same arithmetic, operation for operation (track_magnet.h:92-178, track_magnet_kick.h:
// 265-370, track_magnet_radiation.h:64-94,233-267, synrad_spectrum.h:22-77,
// local_particle_custom_api.h:36-50) -- the host build stays bit-identical to the oracle --
// written across the particles of the thread with the guard-free IEEE sequences of
// xtb_math.cuh; what depends on the beam alone (classical radius, 2 r0 c Q^2 ...) is formed once
// per call, the products with literal-zero coefficient sets (main, relative strengths) are
// left out as the exact identities they are.  Mean model: all here; quantum model: the field
// and path length here, the photon loop per particle (synrad_emit_photons).
// What the thin radiating kick needs of the beam and of each particle's reference, formed once
// per RUN of ops (xtb_interp.cuh::xtb_run_heavy) instead of once per element: the operations
// are the reference's, on the same operands -- only not repeated.  The reciprocals serve the
// divisions by these constants (div_by: same correctly rounded quotients, 3 FP64 instructions
// instead of ~10).
template <int N>
struct ThinRadRun {
    double brho[N];      // p0c / C_LIGHT / q0
    double g0[N], g0m[N];   // gamma0, gamma0 * mass0
    double b0[N], rb0[N];   // beta0, RN(1 / beta0)
    double K1, K2, rK2, rC, rQ;
};
template <int N>
__device__ __forceinline__ void thin_rad_run_begin(ThinRadRun<N>& r, const PSlot (&G)[N], const XtbTrackArgs& a) {
    const double q0 = a.part.q0, mass0 = a.part.mass0;
    double pc[N], cl[N], qq[N], t0[N];
    XTB_LANES { pc[k] = G[k].ld(F_P0C);  cl[k] = XTB_C_LIGHT;  qq[k] = q0; }
    xtb_vdiv<N>(t0, pc, cl);
    xtb_vdiv<N>(r.brho, t0, qq);                          // brho_0 = p0c / C_LIGHT / q0
    XTB_LANES {
        r.g0[k] = G[k].ld(F_GAMMA0);
        r.g0m[k] = r.g0[k] * mass0;
        r.b0[k] = G[k].ld(F_BETA0);
    }
    xtb_vrcp<N>(r.rb0, r.b0);
    const double Q0_coulomb = fabs(q0) * XTB_QELEM;
    const double mass0_kg = mass0 / XTB_C_LIGHT / XTB_C_LIGHT * XTB_QELEM;
    const double r0_m = Q0_coulomb * Q0_coulomb
                        / (4 * XTB_PI * XTB_EPSILON_0 * mass0_kg * XTB_C_LIGHT * XTB_C_LIGHT);
    r.K1 = 2 * r0_m * XTB_C_LIGHT * Q0_coulomb * Q0_coulomb;
    r.K2 = 3 * mass0_kg;
    r.rK2 = 1. / r.K2;
    r.rC = 1. / XTB_C_LIGHT;
    r.rQ = 1. / XTB_QELEM;
}

template <int N, bool FRZ>
__device__ __forceinline__ void thin_rad_kick_run(PState (&P)[N], const bool (&live)[N], const PSlot (&G)[N],
                                                  const XtbTrackArgs& a, const BodyPar& b,
                                                  const ThinRadRun<N>& r) {
    const double length = b.q[0];
    double old_px[N], old_py[N], old_zeta[N];
    XTB_LANES { old_px[k] = P[k].px;  old_py[k] = P[k].py;  old_zeta[k] = P[k].zeta; }
    // the kick (magnet_kick_n with kick_weight 1; only the user coefficients are set here).
    // The Horner sums are kept: for chi == 1 they ARE the sums of the field evaluation below
    // (same coefficients, x and y do not move in a thin kick, 0.5 * (x + x) == x).
    double um[N], un[N];
    XTB_LANES { um[k] = 0.;  un[k] = 0.; }
    if (b.has_user()) {
        const double* __restrict__ cu = b.cu();
        const int ou = b.order_user();
        XTB_LANES {
            const double x = P[k].x, y = P[k].y, chi = P[k].chi;
            switch (ou) {           // (uniform: the op's)
            case 0: horner_kick_c<0>(x, y, chi, cu, um[k], un[k]);  break;
            case 1: horner_kick_c<1>(x, y, chi, cu, um[k], un[k]);  break;
            case 2: horner_kick_c<2>(x, y, chi, cu, um[k], un[k]);  break;
            default: horner_kick(x, y, chi, cu, ou, um[k], un[k]);  break;
            }
            P[k].px += 1.0 * (-um[k]);
            P[k].py += 1.0 * un[k];
        }
    }
    magnet_kick_curv_n<N, FRZ>(P, b, 1.0);
    if (!(b.radiation_flag() && length > 0)) return;

    // field at the mean position (x, y do not move in a thin kick): evaluate_field_from_strengths
    // with the user coefficients; the main and relative sets are all zero here
    double Bx[N], By[N], t0[N], t1[N];
    {
        const double rlen = b.q[1];          // RN(1 / length), folded by the host lowering
        XTB_LANES {
            double m = um[k], n = un[k];
            if (b.has_user() && P[k].chi != 1.0)
                horner_kick(0.5 * (P[k].x + P[k].x), 0.5 * (P[k].y + P[k].y), 1., b.cu(),
                            b.order_user(), m, n);
            const double dpx = (-m + -0.) + 0.;           // dpx_mul + dpx_main + dpx_rel
            const double dpy = (n + 0.) + 0.;
            Bx[k] = div_by(dpy * r.brho[k], length, rlen);        // Bx_T = dpy * brho_0 / length
            By[k] = div_by(-dpx * r.brho[k], length, rlen);       // By_T = -dpx * brho_0 / length
        }
    }
    // compute_b_perp_mod (Bz = 0 * brho_0: its terms are exact zeros)
    double opd[N], ropd[N], iix[N], iiy[N], iis[N], bperp[N], lpath[N];
    XTB_LANES opd[k] = 1. + P[k].delta;
    xtb_vrcp<N>(ropd, opd);
    XTB_LANES {
        iix[k] = div_by(0.5 * (old_px[k] + P[k].px), opd[k], ropd[k]);
        iiy[k] = div_by(0.5 * (old_py[k] + P[k].py), opd[k], ropd[k]);
        t0[k] = 1 - iix[k] * iix[k] + iiy[k] * iiy[k];
    }
    xtb_vsqrt<N>(iis, t0);
    XTB_LANES {
        const double B_par = Bx[k] * iix[k] + By[k] * iiy[k];
        const double px_ = Bx[k] - B_par * iix[k];
        const double py_ = By[k] - B_par * iiy[k];
        const double pz_ = -(B_par * iis[k]);
        t0[k] = px_ * px_ + py_ * py_ + pz_ * pz_;
    }
    xtb_vsqrt<N>(bperp, t0);
    // (a particle on the axis of a quadrupole sees no field: zero, and anything below the range
    // of the guard-free sequence, goes through the IEEE square root)
    XTB_LANES { if (!(t0[k] > 1e-280)) bperp[k] = sqrt(t0[k]); }
    XTB_LANES lpath[k] = P[k].rvv * (length - (P[k].zeta - old_zeta[k]));

    // (programs with random emission run on the one-lane kernels, xtb_kernel_inst.cu: the
    // calls are not even compiled into the two-lane run loop of the mean model, whose register
    // allocation they cost 100 bytes of spills)
    if constexpr (N == 1) {
        if (b.radiation_flag() >= 2) {
            if (live[0]) {
                if (b.radiation_flag() == 2) synrad_emit_photons<FRZ>(P[0], G[0], a, bperp[0], lpath[0]);
                else synrad_emit_total_energy_loss<FRZ>(P[0], G[0], a, bperp[0], lpath[0]);
            }
            return;
        }
    }
    // synrad_average_kick (mean model)
    double ft[N], nd[N];
    XTB_LANES {
        const double gamma = r.g0[k] * opd[k];
        const double Ps_W = div_by(r.K1 * gamma * gamma * bperp[k] * bperp[k], r.K2, r.rK2);
        // Delta_E_eV = Ps_W * lpath / C / QELEM
        t0[k] = div_by(div_by(Ps_W * lpath[k], XTB_C_LIGHT, r.rC), XTB_QELEM, r.rQ);
        t1[k] = r.g0m[k] * opd[k];
    }
    xtb_vdiv<N>(t1, t0, t1);
    XTB_LANES {
        ft[k] = 1 - t1[k];
        nd[k] = opd[k] * ft[k] - 1;                       // (delta + 1) * f_t - 1
    }
    if (!FRZ) {
        // LocalParticle_update_delta
        double db0[N], pb0[N], nopd[N], rvv[N], rpp[N], rv0v[N];
        XTB_LANES {
            db0[k] = nd[k] * r.b0[k];
            t0[k] = db0[k] * db0[k] + 2 * db0[k] * r.b0[k] + 1;
        }
        xtb_vsqrt<N>(pb0, t0);
        XTB_LANES { pb0[k] = pb0[k] - 1;  nopd[k] = 1 + nd[k];  t0[k] = 1 + pb0[k]; }
        xtb_vdiv<N>(rvv, nopd, t0);
        xtb_vrcp<N>(rpp, nopd);
        xtb_vrcp<N>(rv0v, rvv);
        XTB_LANES {
            if (live[k]) {
                P[k].delta = nd[k];  P[k].rvv = rvv[k];  P[k].rv0v = rv0v[k];  P[k].rpp = rpp[k];
                G[k].st(F_PTAU, div_by(pb0[k], r.b0[k], r.rb0[k]));
            }
        }
    }
    XTB_LANES {
        if (live[k]) { P[k].px *= ft[k];  P[k].py *= ft[k]; }
    }
}

// (one element on its own: the generic body dispatch)
template <int N, bool FRZ>
__device__ __forceinline__ void thin_rad_kick_n(PState (&P)[N], const bool (&live)[N], const PSlot (&G)[N],
                                                const XtbTrackArgs& a, const BodyPar& b) {
    ThinRadRun<N> r;
    thin_rad_run_begin<N>(r, G, a);
    thin_rad_kick_run<N, FRZ>(P, live, G, a, b, r);
}

// is this body op a kick-only element in one uniform step (what thin_rad_kick_* implement)?
__device__ __forceinline__ bool is_thin_kick_body(const uint32_t aux) {
    // integrator 3 (uniform), drift model -1, no relative / main strengths, not drift-only,
    // no merged edges, one kick
    const uint32_t mask = 3u | (15u << 2) | (1u << 8) | (1u << 9) | (1u << 12) | (1u << 13) | (1u << 14);
    return ((aux & mask) == 3u) && ((aux >> 15) == 1u);
}

// track_magnet_body_single_particle, track_magnet.h:26-285, on N particles at once.
// The three integrators are ONE loop nest -- radiation segments x sub-steps, each sub-step a
// drift and (except the last of a segment) a kick -- so that the drift and kick bodies exist
// once and the particles stay in registers from the first sub-step to the last:
//   teapot   (track_magnet.h:191-213)  1 segment,  n_kicks + 1 sub-steps (edge, inside.., edge)
//   uniform  (:214-227)                n_kicks segments of 2 sub-steps
//   yoshida4 (:228-274)                n_slices segments of 8 sub-steps
//   drift-only shortcut (:181-185)     1 segment, 1 sub-step, no kick
// Lengths and weights are formed by the reference's own expressions.  `live`: lanes whose
// particle is real (radiation touches the caller's SoA and the RNG stream: real lanes only).
template <int N, bool SYNRAD, bool FRZ>
__device__ __forceinline__ void magnet_body_n(PState (&P)[N], const bool (&live)[N], const PSlot (&G)[N],
                                              const XtbTrackArgs& a, const double* __restrict__ q,
                                              const int32_t aux) {
    const BodyPar b = body_par(q, aux);
    const double length = q[0], k0d = q[1], k1d = q[2], hd = q[3];
    const int dm = b.drift_model();
    const int nk = b.num_kicks();
    const int integ = b.drift_only() ? 0 : b.integrator();
    if (SYNRAD && is_thin_kick_body(b.a)) {
        thin_rad_kick_n<N, FRZ>(P, live, G, a, b);       // kick-only element, one step
        return;
    }

    int n_seg = 1, n_sub = 1;
    double seg_length = length;
    double kick_weight = 0., w_edge = 0., w_inside = 0., slice_length = 0.;
    if (integ == 1) {               // teapot
        kick_weight = 1. / nk;
        w_edge = 0.5;
        if (nk > 1) {
            w_edge = 1. / (2 * (1 + nk));
            w_inside = ((double) nk) / ((double) ((int64_t) nk * nk) - 1);
        }
        n_sub = nk + 1;
    } else if (integ == 3) {        // uniform
        kick_weight = 1. / nk;
        n_seg = nk;
        n_sub = 2;
        seg_length = kick_weight * length;      // drift_weight * length
    } else if (integ == 2) {        // yoshida 4
        const int num_slices = nk / 7 + (nk % 7 != 0);
        slice_length = length / (num_slices);
        kick_weight = 1. / num_slices;
        n_seg = num_slices;
        n_sub = 8;
        seg_length = slice_length;
    }
    if (b.edge_in()) {
#pragma unroll
        for (int k = 0; k < N; ++k) edge_linear(P[k], b.edges()[0], b.edges()[1]);
    }
    RadSnapshot snap[N];
    for (int seg = 0; seg < n_seg; ++seg) {
#pragma unroll
        for (int k = 0; k < N; ++k) rad_begin<SYNRAD>(snap[k], P[k]);
#pragma unroll 1
        for (int j = 0; j < n_sub; ++j) {
            double dl, kw = kick_weight;
            int oc = 0;
            if (integ == 1) {
                const bool edge = (j == 0) || (j == n_sub - 1);
                dl = (edge ? w_edge : w_inside) * length;
                oc = edge ? 0 : 1;
            } else if (integ == 3) {
                dl = 0.5 * kick_weight * length;
            } else if (integ == 2) {
                dl = slice_length * XTB_Y6[j];
                kw = kick_weight * XTB_Y6[8 + j];
                oc = (j < 7 - j) ? j : 7 - j;
            } else {
                dl = length;
            }
            magnet_drift_n<N, FRZ>(P, dl, k0d, k1d, hd, dm, b.trig(), oc);
            const bool has_kick = (integ == 3) ? (j == 0) : (j < n_sub - 1);
            if (has_kick) magnet_kick_n<N, FRZ>(P, b, kw);
        }
        if (SYNRAD) {
            for (int k = 0; k < N; ++k)
                if (live[k]) rad_end<SYNRAD, FRZ>(snap[k], P[k], G[k], a, b, seg_length);
        }
    }
    if (b.edge_out()) {
#pragma unroll
        for (int k = 0; k < N; ++k) edge_linear(P[k], b.edges()[2], b.edges()[3]);
    }
}

// single-particle entry (generic slow path)
template <bool SYNRAD, bool FRZ>
__device__ __noinline__ void magnet_body(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                         const double* __restrict__ q, const int32_t aux) {
    const bool live[1] = {true};
    magnet_body_n<1, SYNRAD, FRZ>(reinterpret_cast<PState(&)[1]>(P), live,
                                  reinterpret_cast<const PSlot(&)[1]>(G), a, q, aux);
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
    const double co2 = b0 / XTB_POW2(xtb_cos_glibc(fi0));
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

// Wedge_single_particle, track_wedge.h:14-74.  sin_t, cos_t, tan_t: sin / cos / tan of theta, an
// element constant, from the host (sin and tan are odd to the bit, cos even: the lowering's
// values of the face angle serve theta = -face_angle).
template <bool FRZ>
__device__ __noinline__ void wedge(PState& P, const PSlot& G, const double theta, const double k0,
                                   const double sin_t, const double cos_t, const double tan_t) {
    const double b1 = k0 * P.chi;
    if (fabs(b1) < 10e-10) {
        yrotation<FRZ>(P, G, -sin_t, cos_t, -tan_t);
        return;
    }
    const double rvv = P.rvv;
    const double x = P.x, px = P.px, py = P.py;
    const double one_plus_delta = P.delta + 1.0;
    const double A = 1.0 / sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(py));
    const double pz = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(px) - XTB_POW2(py));
    const double new_px = px * cos_t + (pz - b1 * x) * sin_t;
    const double new_pz = sqrt(XTB_POW2(one_plus_delta) - XTB_POW2(new_px) - XTB_POW2(py));
    const double new_x = x * cos_t
        + (x * px * xtb_sin_glibc(2 * theta) + XTB_POW2(sin_t) * (2 * x * pz - b1 * XTB_POW2(x)))
              / (new_pz + pz * cos_t - px * sin_t);
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
        if (should_rotate) wedge<FRZ>(P, G, -face_angle, knorm[0], -sin_, cos_, -tan_);
    } else {
        if (should_rotate) wedge<FRZ>(P, G, -face_angle, knorm[0], -sin_, cos_, -tan_);
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
        if (sin_ > -99.) wedge<FRZ>(P, G, -e1, k, -sin_, cos_, -tan_);
    } else if (side == 1) {
        if (sin_ > -99.) wedge<FRZ>(P, G, -e1, k, -sin_, cos_, -tan_);
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
