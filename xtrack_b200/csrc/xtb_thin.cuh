// xtb_thin.cuh -- single-particle maps of the thin element set, on registers.
//
// Every function restates the arithmetic of the cited reference routine in the
// reference's operation order, so that the EXACT build (-fmad=false) rounds
// like the reference's CPU build; the default build lets nvcc contract to FMA.
// Element-constant work (coefficient scaling, factorials, trigonometry of
// element parameters) is done once by the host lowering (xtrack_b200/lowering.py)
// with the same IEEE operations, so only per-particle arithmetic remains here.
#pragma once
#include <type_traits>
#include "xtb_state.cuh"
#include "xtb_libm.cuh"
#include "xtb_math.cuh"

// 1 + t/2 as the reference writes it (track_drift.h:19).  t/2 is exact, so the fused
// multiply-add rounds the same real number once: bit-identical, one FP64 instruction less.
// (Host build: plain arithmetic, same value.)
#ifdef __CUDA_ARCH__
#define XTB_ONE_PLUS_HALF(t) __fma_rn((t), 0.5, 1.)
#else
#define XTB_ONE_PLUS_HALF(t) (1. + (t) / 2.)
#endif

#define XTB_C_LIGHT 299792458.0
#define XTB_PI 3.1415926535897932384626433832795028841971693993751
#define XTB_DEG2RAD 0.0174532925199432957692369076848861271344287188854

// FRZ = freeze_longitudinal: writes to zeta, s, delta, ptau, rpp, rvv are
// suppressed (FREEZE_VAR_*, xtrack/particles/particles.py:193-227, line.py:4446-4508).

// xtrack/beam_elements/elements_src/track_drift.h:11-22
template <bool FRZ, class S>
__device__ __forceinline__ void drift_expanded(S& P, const double length) {
    const double xp = P.px * P.rpp;
    const double yp = P.py * P.rpp;
    const double dzeta = 1 - P.rv0v * XTB_ONE_PLUS_HALF(xp * xp + yp * yp);
    P.x += xp * length;
    P.y += yp * length;
    if (!FRZ) {
        P.s += length;
        P.zeta += length * dzeta;
    }
}

// drift_expanded without the `s += length`: for the hot loop when s is carried once per
// thread (block-uniform s, see xtb_run_fast SUNI)
template <bool FRZ, class S>
__device__ __forceinline__ void drift_expanded_nos(S& P, const double length) {
    const double xp = P.px * P.rpp;
    const double yp = P.py * P.rpp;
    const double dzeta = 1 - P.rv0v * XTB_ONE_PLUS_HALF(xp * xp + yp * yp);
    P.x += xp * length;
    P.y += yp * length;
    if (!FRZ) P.zeta += length * dzeta;
}

// track_drift.h:26-40
template <bool FRZ>
__device__ __forceinline__ void drift_exact(PState& P, const double length) {
    const double one_plus_delta = 1. + P.delta;
    const double one_over_pz =
        1. / sqrt(one_plus_delta * one_plus_delta - P.px * P.px - P.py * P.py);
    const double dzeta = 1 - P.rv0v * one_plus_delta * one_over_pz;
    P.x += P.px * one_over_pz * length;
    P.y += P.py * one_over_pz * length;
    if (!FRZ) {
        P.zeta += dzeta * length;
        P.s += length;
    }
}

// Horner evaluation of kick_simple_single_coordinates, track_magnet_kick.h:183-228.
// c[] holds the host-scaled coefficients (knl[i]*factor*inv_factorial_i, same
// product order as the reference) as pairs (normal, skew) from the highest
// order down; chi multiplies each coefficient as in the reference.
__device__ __forceinline__ void horner_kick(const double x, const double y, const double chi,
                                            const double* __restrict__ c, const int order,
                                            double& dpx_mul, double& dpy_mul) {
    dpx_mul = chi * c[0];
    dpy_mul = chi * c[1];
    for (int i = 1; i <= order; ++i) {
        const double zre = dpx_mul * x - dpy_mul * y;
        const double zim = dpx_mul * y + dpy_mul * x;
        dpx_mul = chi * c[2 * i] + zre;
        dpy_mul = chi * c[2 * i + 1] + zim;
    }
}

// (the same with the order known at compile time: no loop counter, no address arithmetic)
template <int ORDER>
__device__ __forceinline__ void horner_kick_c(const double x, const double y, const double chi,
                                              const double* __restrict__ c, double& dpx_mul,
                                              double& dpy_mul) {
    dpx_mul = chi * c[0];
    dpy_mul = chi * c[1];
#pragma unroll
    for (int i = 1; i <= ORDER; ++i) {
        const double zre = dpx_mul * x - dpy_mul * y;
        const double zim = dpx_mul * y + dpy_mul * x;
        dpx_mul = chi * c[2 * i] + zre;
        dpy_mul = chi * c[2 * i + 1] + zim;
    }
}

// Thin multipole without curvature: Multipole (model -1) of multipole.h:16-75
// -> track_magnet_kick_single_particle, track_magnet_kick.h:24-144, with
// hxl == 0 and all-zero knl_rel / main strengths (their kicks are exact zeros).
__device__ __forceinline__ void mult_kick(PState& P, const double* __restrict__ c, const int order) {
    double dpx_mul, dpy_mul;
    horner_kick(P.x, P.y, P.chi, c, order, dpx_mul, dpy_mul);
    P.px += -dpx_mul;
    P.py += dpy_mul;
}

// Thin multipole with hxl != 0: adds the curvature terms of
// track_magnet_kick.h:98-142.  q = [hl, B0, B1, 0], c = coefficients.
//   hl = h*length*kw + hxl*kw ; B0 = -(k0_h_corr*length + k0l)*kw*htot ;
//   B1 = htot*(k1_h_corr*length + k1l)*kw          (host, reference order, chi = 1)
template <bool FRZ>
__device__ __forceinline__ void mult_kick_h(PState& P, const double* __restrict__ q,
                                            const double* __restrict__ c, const int order,
                                            const bool has_b1) {
    const double x = P.x, y = P.y, chi = P.chi;
    double dpx_mul, dpy_mul;
    horner_kick(x, y, chi, c, order, dpx_mul, dpy_mul);
    P.px += -dpx_mul;
    P.py += dpy_mul;

    const double hl = q[0];
    double dpx = hl * (1. + P.delta);
    const double dzeta = -P.rv0v * hl * x;
    dpx += chi * q[1] * x;
    if (has_b1) {
        const double b1 = chi * q[2];
        dpx += b1 * (-x * x + 0.5 * y * y);
        P.px += dpx;
        P.py += b1 * x * y;
    } else {
        P.px += dpx;
    }
    if (!FRZ) P.zeta += dzeta;
}

// LocalParticle_update_ptau, local_particle_custom_api.h:21-32
template <bool FRZ>
__device__ __forceinline__ void update_ptau(PState& P, const PSlot& G, const double beta0,
                                            const double ptau) {
    if (FRZ) return;
    const double irpp = sqrt(ptau * ptau + 2 * ptau / beta0 + 1);
    const double new_rpp = 1. / irpp;
    const double new_rvv = irpp / (1 + beta0 * ptau);
    P.delta = irpp - 1;
    P.rvv = new_rvv;
    P.rv0v = 1. / new_rvv;
    P.rpp = new_rpp;
    G.st(F_PTAU, ptau);
}

// LocalParticle_update_delta, local_particle_custom_api.h:36-50
template <bool FRZ>
__device__ __forceinline__ void update_delta(PState& P, const PSlot& G, const double beta0,
                                             const double new_delta) {
    if (FRZ) return;
    const double delta_beta0 = new_delta * beta0;
    const double ptau_beta0 = sqrt(delta_beta0 * delta_beta0 + 2 * delta_beta0 * beta0 + 1) - 1;
    const double one_plus_delta = 1 + new_delta;
    const double rvv = one_plus_delta / (1 + ptau_beta0);
    const double rpp = 1 / one_plus_delta;
    const double ptau = ptau_beta0 / beta0;
    P.delta = new_delta;
    P.rvv = rvv;
    P.rv0v = 1. / rvv;
    P.rpp = rpp;
    G.st(F_PTAU, ptau);
}

// LocalParticle_add_to_energy, local_particle_custom_api.h:196-216
template <bool FRZ>
__device__ __forceinline__ void add_to_energy(PState& P, const PSlot& G, const double beta0,
                                              const double delta_energy, const int pz_only) {
    double ptau = G.ld(F_PTAU);
    const double p0c = G.ld(F_P0C);
    const double charge_ratio = G.ld(F_CHARGE_RATIO);
    const double mass_ratio = charge_ratio / P.chi;
    ptau += delta_energy / p0c / mass_ratio;
    const double old_rpp = P.rpp;
    update_ptau<FRZ>(P, G, beta0, ptau);
    if (!pz_only) {
        const double f = old_rpp / P.rpp;
        P.px *= f;
        P.py *= f;
    }
}

// LocalParticle_kill_particle, local_particle_custom_api.h:248-256
template <bool FRZ>
__device__ __forceinline__ void kill_particle(PState& P, const PSlot& G, const int kill_state) {
    P.x = 1e30;  P.px = 1e30;  P.y = 1e30;  P.py = 1e30;
    if (!FRZ) P.zeta = 1e30;
    update_delta<FRZ>(P, G, G.ld(F_BETA0), -1.);
    P.state = kill_state;
}

// Thin cavity kick: track_rf_kick_single_particle, track_rf.h:18-167 with
// order = -1 and no transverse voltage.  q = [V, f, harmonic, lag, phase].
template <bool FRZ>
__device__ __forceinline__ void cavity_kick(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                            const double voltage, double frequency,
                                            const double harmonic, const double lag,
                                            const double phase, const int absolute_time) {
    double phase0 = 0;
    const double beta0 = G.ld(F_BETA0);
    if (harmonic != 0) {
        const double t_rev0 = a.line_length / (beta0 * XTB_C_LIGHT);
        frequency += (harmonic / t_rev0);
    }
    if (absolute_time == 1) {
        phase0 += 2 * XTB_PI * P.at_turn * frequency * a.part.t_sim;
    }
    const double q = fabs(a.part.q0) * G.ld(F_CHARGE_RATIO);
    const double tau = P.zeta / beta0;
    // voltage == 0 (block-uniform): q*0*sin(.) is an exact zero, the sine is not evaluated
    // (xtb_sin_glibc: the C library's sine to the bit, xtb_libm.cuh)
    const double energy_kick = (voltage == 0.) ? 0. : q * voltage
        * xtb_sin_glibc(phase0 + XTB_DEG2RAD * lag + phase - (2.0 * XTB_PI) / XTB_C_LIGHT * frequency * tau);
    if (!a.kill_cavity_kick) {
        add_to_energy<FRZ>(P, G, beta0, energy_kick + 0., 1);
    }
}

// RF multipole kick: track_rf.h:18-167 with order >= 0 (RFMultipole, rfmultipole.h).
// q = [V, f, lag, phase, 0], then (bal_n, bal_s, pn, ps, phase_n, phase_s) per order, with
// bal = factor_knl_ksl * k[kk] / kk! folded by the host; `order` is the highest order with a
// non-zero strength (-1: none).  Terms whose strength is a literal zero are exact zeros in
// the reference (cos * (0 * z)); they, and the sin/cos that only they use, are left out --
// the branches are on element constants, hence uniform over the block.
// (the transverse kick and the energy change, before they are applied)
__device__ __forceinline__ void rfmult_terms(const double x, const double y, const double zeta,
                                             const double beta0, const double p0c, const double qq,
                                             const double* __restrict__ q, const int order,
                                             double& dpx, double& dpy, double& delta_energy) {
    const double voltage = q[0], frequency = q[1], lag = q[2], phase = q[3];
    const double* __restrict__ t = q + 5;
    const double phase0 = 0;
    const double tau = zeta / beta0;
    const double energy_kick = (voltage == 0.) ? 0. : qq * voltage
        * xtb_sin_glibc(phase0 + XTB_DEG2RAD * lag + phase - (2.0 * XTB_PI) / XTB_C_LIGHT * frequency * tau);

    double dptr = 0.0, zre = 1.0, zim = 0.0;
    dpx = 0.0;
    dpy = 0.0;
    for (int kk = 0; kk <= order; kk++) {
        const double* __restrict__ e = t + 6 * kk;
        const double bal_n_kk = e[0];
        const double bal_s_kk = e[1];
        const bool has_n = (bal_n_kk != 0.), has_s = (bal_s_kk != 0.);
        const double pn_kk = phase0 + XTB_DEG2RAD * e[2] + e[4] - (2.0 * XTB_PI) / XTB_C_LIGHT * frequency * tau;
        const double ps_kk = phase0 + XTB_DEG2RAD * e[3] + e[5] - (2.0 * XTB_PI) / XTB_C_LIGHT * frequency * tau;
        double cn = 0., cs = 0., sn = 0., ss = 0.;
        if (has_n) { cn = xtb_cos_glibc(pn_kk);  sn = xtb_sin_glibc(pn_kk); }
        if (has_s) {
            if (has_n && ps_kk == pn_kk) { cs = cn;  ss = sn; }
            else { cs = xtb_cos_glibc(ps_kk);  ss = xtb_sin_glibc(ps_kk); }
        }
        if (has_n && has_s) {
            dpx += cn * (bal_n_kk * zre) - cs * (bal_s_kk * zim);
            dpy += cs * (bal_s_kk * zre) + cn * (bal_n_kk * zim);
        } else if (has_n) {
            dpx += cn * (bal_n_kk * zre);
            dpy += cn * (bal_n_kk * zim);
        } else if (has_s) {
            dpx += -(cs * (bal_s_kk * zim));
            dpy += cs * (bal_s_kk * zre);
        }
        const double zret = zre * x - zim * y;
        zim = zim * x + zre * y;
        zre = zret;
        if (has_n && has_s) dptr += sn * (bal_n_kk * zre) - ss * (bal_s_kk * zim);
        else if (has_n) dptr += sn * (bal_n_kk * zre);
        else if (has_s) dptr += -(ss * (bal_s_kk * zim));
    }
    const double rf_energy_kick = -qq * ((frequency * (2.0 * XTB_PI / XTB_C_LIGHT) * p0c) * dptr);
    delta_energy = energy_kick + rf_energy_kick;
}

template <bool FRZ>
__device__ __forceinline__ void rfmult_kick(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                            const double* __restrict__ q, const int order) {
    const double beta0 = G.ld(F_BETA0);
    const double qq = fabs(a.part.q0) * G.ld(F_CHARGE_RATIO);
    double dpx, dpy, delta_energy;
    rfmult_terms(P.x, P.y, P.zeta, beta0, G.ld(F_P0C), qq, q, order, dpx, dpy, delta_energy);
    P.px += -P.chi * dpx;
    P.py += P.chi * dpy;
    if (!a.kill_cavity_kick) {
        add_to_energy<FRZ>(P, G, beta0, delta_energy, 1);
    }
}

// Crab cavity: track_rf_kick_single_particle (track_rf.h:18-167) with voltage = 0, order = -1 and
// a transverse voltage: an RF dipole kick whose strength is transverse_voltage / p0c of the
// particle.  q = [V_t, f, lag, phase].  (energy_kick = q * 0 * sin(.) is an exact zero.)
template <bool FRZ>
__device__ __forceinline__ void crab_kick(PState& P, const PSlot& G, const XtbTrackArgs& a,
                                          const double* __restrict__ q, const int absolute_time) {
    const double tv = q[0], frequency = q[1], tlag = q[2], tphase = q[3];
    double phase0 = 0;
    const double beta0 = G.ld(F_BETA0);
    if (absolute_time == 1) phase0 += 2 * XTB_PI * P.at_turn * frequency * a.part.t_sim;
    const double qq = fabs(a.part.q0) * G.ld(F_CHARGE_RATIO);
    const double tau = P.zeta / beta0;
    double delta_energy = 0.;
    if (tv != 0) {
        const double zre = 1.0, zim = 0.0;
        const double x = P.x, y = P.y;
        const double p0c = G.ld(F_P0C);
        const double pn_kk = phase0 + XTB_DEG2RAD * (tlag + 90.) + tphase
                             - (2.0 * XTB_PI) / XTB_C_LIGHT * frequency * tau;
        const double k0l = tv / p0c;
        const double cn = xtb_cos_glibc(pn_kk);
        const double sn = xtb_sin_glibc(pn_kk);
        double dpx = 0.0, dpy = 0.0, dptr = 0.0;
        dpx += cn * (k0l * zre);
        dpy += cn * (k0l * zim);
        const double zret = zre * x - zim * y;
        dptr += sn * (k0l * zret);
        delta_energy += -qq * ((frequency * (2.0 * XTB_PI / XTB_C_LIGHT) * p0c) * dptr);
        P.px += -P.chi * dpx;
        P.py += P.chi * dpy;
    }
    if (!a.kill_cavity_kick) add_to_energy<FRZ>(P, G, beta0, 0. + delta_energy, 1);
}

// ---- RF elements on all the particles of a thread at once -------------------------------------
// A ring has a handful of RF elements per turn, yet they took ~10 % of the thin kernel: each
// went through the generic out-of-line path once per particle -- full state assembled and
// taken apart again, a square root, two reciprocals and five divisions with their slow-path
// guards, one dependency chain at a time.  These work on the thread's N particles together,
// straight on their hot state and cached cold fields, the energy update (LocalParticle_
// add_to_energy with pz_only = 1 -> LocalParticle_update_ptau, local_particle_custom_api.h:
// 196-216, 21-32) written step by step ACROSS the particles with the guard-free IEEE sequences
// of xtb_math.cuh: N interleaved chains, identical bits.  Only lanes with a real particle are
// written back (a lane without one keeps its benign state).
template <int N, bool FRZ, class S>
__device__ __forceinline__ void add_to_energy_lanes(S (&P)[N], PCold (&C)[N], const bool (&live)[N],
                                                    const double (&delta_energy)[N]) {
    if (FRZ) return;
    double cr[N], chi[N], p0c[N], b0[N], mr[N], t1[N], t2[N], ptau[N], tp[N], u[N], arg[N], irpp[N],
        rpp[N], den[N], rvv[N], rv0v[N];
    XTB_LANES { cr[k] = C[k].charge_ratio;  chi[k] = P[k].chi;  p0c[k] = C[k].p0c;  b0[k] = C[k].beta0; }
    xtb_vdiv<N>(mr, cr, chi);                       // mass_ratio = charge_ratio / chi
    xtb_vdiv<N>(t1, delta_energy, p0c);             // ptau += delta_energy / p0c / mass_ratio
    xtb_vdiv<N>(t2, t1, mr);
    XTB_LANES ptau[k] = C[k].ptau + t2[k];
    XTB_LANES tp[k] = 2 * ptau[k];
    xtb_vdiv<N>(u, tp, b0);
    XTB_LANES arg[k] = ptau[k] * ptau[k] + u[k] + 1;
    xtb_vsqrt<N>(irpp, arg);                        // irpp = sqrt(ptau^2 + 2 ptau / beta0 + 1)
    xtb_vrcp<N>(rpp, irpp);
    XTB_LANES den[k] = 1 + b0[k] * ptau[k];
    xtb_vdiv<N>(rvv, irpp, den);
    xtb_vrcp<N>(rv0v, rvv);
    XTB_LANES {
        if (live[k]) {
            P[k].delta = irpp[k] - 1;
            P[k].rpp = rpp[k];
            P[k].rv0v = rv0v[k];
            if constexpr (std::is_same<S, PState>::value) P[k].rvv = rvv[k];
            C[k].rvv = rvv[k];
            C[k].ptau = ptau[k];
        }
    }
}

// Cavity (not absolute_time) on the thread's particles: cavity_kick above, lane-parallel
template <int N, bool FRZ, class S>
static __device__ __noinline__ void cavity_lanes(S (&P)[N], PCold (&C)[N], const bool (&live)[N],
                                                 const double* __restrict__ q, const XtbTrackArgs& a) {
    const double voltage = q[0], harmonic = q[2], lag = q[3], phase = q[4];
    const double phase0 = 0;
    double dE[N];
    XTB_LANES {
        double frequency = q[1];
        const double beta0 = C[k].beta0;
        if (harmonic != 0) {
            const double t_rev0 = a.line_length / (beta0 * XTB_C_LIGHT);
            frequency += (harmonic / t_rev0);
        }
        const double qk = fabs(a.part.q0) * C[k].charge_ratio;
        const double tau = P[k].zeta / beta0;
        const double energy_kick = (voltage == 0.) ? 0. : qk * voltage
            * xtb_sin_glibc(phase0 + XTB_DEG2RAD * lag + phase - (2.0 * XTB_PI) / XTB_C_LIGHT * frequency * tau);
        dE[k] = energy_kick + 0.;
    }
    if (!a.kill_cavity_kick) add_to_energy_lanes<N, FRZ>(P, C, live, dE);
}

// RF multipole on the thread's particles: rfmult_kick above, the energy update lane-parallel
template <int N, bool FRZ, class S>
static __device__ __noinline__ void rfmult_lanes(S (&P)[N], PCold (&C)[N], const bool (&live)[N],
                                                 const double* __restrict__ q, const int order,
                                                 const XtbTrackArgs& a) {
    double dE[N];
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const double qq = fabs(a.part.q0) * C[k].charge_ratio;
        double dpx, dpy;
        rfmult_terms(P[k].x, P[k].y, P[k].zeta, C[k].beta0, C[k].p0c, qq, q, order, dpx, dpy, dE[k]);
        if (live[k]) {
            P[k].px += -P[k].chi * dpx;
            P[k].py += P[k].chi * dpy;
        }
    }
    if (!a.kill_cavity_kick) add_to_energy_lanes<N, FRZ>(P, C, live, dE);
}

// DipoleEdgeLinear_single_particle, track_dipole_edge_linear.h:30-39
__device__ __forceinline__ void edge_linear(PState& P, const double r21, const double r43) {
    P.px += P.chi * r21 * P.x;
    P.py += P.chi * r43 * P.y;
}

// SRotation_single_particle, track_srotation.h:12-40 (spin left untouched: spin
// tracking is outside the contract and the reference skips it for zero spin)
__device__ __forceinline__ void srotation(PState& P, const double sin_z, const double cos_z) {
    const double x = P.x, y = P.y, px = P.px, py = P.py;
    P.x = cos_z * x + sin_z * y;
    P.y = -sin_z * x + cos_z * y;
    P.px = cos_z * px + sin_z * py;
    P.py = -sin_z * px + cos_z * py;
}

// YRotation_single_particle, track_yrotation.h:12-45
template <bool FRZ>
__device__ __forceinline__ void yrotation(PState& P, const PSlot& G, const double sin_angle,
                                          const double cos_angle, const double tan_angle) {
    const double beta0 = G.ld(F_BETA0);
    const double x = P.x, y = P.y, px = P.px, py = P.py;
    const double t = P.zeta / beta0;
    const double pt = (G.ld(F_PTAU) / beta0) * beta0;
    const double pz = sqrt(1.0 + 2.0 * pt / beta0 + pt * pt - px * px - py * py);
    const double ptt = 1.0 + tan_angle * px / pz;
    const double x_hat = x / (cos_angle * ptt);
    const double px_hat = cos_angle * px - sin_angle * pz;
    const double y_hat = y - tan_angle * x * py / (pz * ptt);
    const double t_hat = t + tan_angle * x * (1.0 / beta0 + pt) / (pz * ptt);
    P.x = x_hat;
    P.px = px_hat;
    P.y = y_hat;
    if (!FRZ) P.zeta = t_hat * beta0;
}

// XRotation_single_particle, track_xrotation.h:12-45
template <bool FRZ>
__device__ __forceinline__ void xrotation(PState& P, const PSlot& G, const double sin_angle,
                                          const double cos_angle, const double tan_angle) {
    const double beta0 = G.ld(F_BETA0);
    const double x = P.x, y = P.y, px = P.px, py = P.py;
    const double t = P.zeta / beta0;
    const double pt = (G.ld(F_PTAU) / beta0) * beta0;
    const double pz = sqrt(1.0 + 2.0 * pt / beta0 + pt * pt - px * px - py * py);
    const double ptt = 1.0 - tan_angle * py / pz;
    const double y_hat = y / (cos_angle * ptt);
    const double py_hat = cos_angle * py + sin_angle * pz;
    const double x_hat = x + tan_angle * y * px / (pz * ptt);
    const double t_hat = t - tan_angle * y * (1.0 / beta0 + pt) / (pz * ptt);
    P.x = x_hat;
    P.py = py_hat;
    P.y = y_hat;
    if (!FRZ) P.zeta = t_hat * beta0;
}

// LimitPolygon_track_local_particle, limitpolygon.h:63-88 (the CPU-only
// inscribed-circle shortcut :28-61 does not change the result)
__device__ __forceinline__ bool polygon_contains(const double x, const double y,
                                                 const double* __restrict__ vx,
                                                 const double* __restrict__ vy, const int n) {
    int is_alive = 0;
    int jj = n - 1;
    for (int ii = 0; ii < n; ++ii) {
        const double Vx_ii = vx[ii], Vx_jj = vx[jj], Vy_ii = vy[ii], Vy_jj = vy[jj];
        if (((Vy_ii > y) != (Vy_jj > y))
            && (x < (Vx_jj - Vx_ii) * (y - Vy_ii) / (Vy_jj - Vy_ii) + Vx_ii)) {
            is_alive = !is_alive;
        }
        jj = ii;
    }
    return is_alive != 0;
}

// global_aperture_check, local_particle_custom_api.h:262-289
__device__ __forceinline__ void global_aperture_check(PState& P, const double lim) {
    const bool inside = (P.x >= -lim) && (P.x <= lim) && (P.y >= -lim) && (P.y <= lim);
    if (P.state > 0 && !inside) P.state = -1;
}

// ---- order-specialised forms used by the fast opcodes (xtb_ops.h) ----------
// Same arithmetic and operation order as horner_kick / mult_kick above, with the
// coefficients already in registers (loaded once per op, shared by the particles
// a thread carries) and the loop unrolled.
// CHI1: every particle of the block has chi == 1.0, so `chi * c` is `c` bit for bit and
// the multiplication is left out.
template <int ORDER, bool CHI1, class S>
__device__ __forceinline__ void mult_kick_c(S& P, const double (&c)[2 * (ORDER + 1)]) {
    const double x = P.x, y = P.y, chi = P.chi;
    double dpx_mul = CHI1 ? c[0] : chi * c[0];
    double dpy_mul = CHI1 ? c[1] : chi * c[1];
#pragma unroll
    for (int i = 1; i <= ORDER; ++i) {
        const double zre = dpx_mul * x - dpy_mul * y;
        const double zim = dpx_mul * y + dpy_mul * x;
        dpx_mul = (CHI1 ? c[2 * i] : chi * c[2 * i]) + zre;
        dpy_mul = (CHI1 ? c[2 * i + 1] : chi * c[2 * i + 1]) + zim;
    }
    P.px += -dpx_mul;
    P.py += dpy_mul;
}

// mult_kick_h with order 0 and no k1 term: q = [hl, B0], c = [cn_0, cs_0]
template <bool FRZ, bool CHI1, class S>
__device__ __forceinline__ void mult_kick_h0(S& P, const double hl, const double b0,
                                             const double cn0, const double cs0) {
    const double x = P.x, chi = P.chi;
    const double dpx_mul = CHI1 ? cn0 : chi * cn0;
    const double dpy_mul = CHI1 ? cs0 : chi * cs0;
    P.px += -dpx_mul;
    P.py += dpy_mul;
    double dpx = hl * (1. + P.delta);
    const double dzeta = -P.rv0v * hl * x;
    dpx += (CHI1 ? b0 : chi * b0) * x;
    P.px += dpx;
    if (!FRZ) P.zeta += dzeta;
}

// Plain normal multipole of order ORDER >= 1: mult_kick_c with every coefficient but cn_ORDER
// a literal zero, the zero operations left out (xtb_ops.h "Zero-coefficient specialisation"):
//   Horner step 1:  zre = cn*x - 0*y = cn*x ;  zim = cn*y + 0*x = cn*y ;  then 0 + z = z
template <bool CHI1, class S>
__device__ __forceinline__ void mult_kick_p1(S& P, const double cn) {
    const double a = CHI1 ? cn : P.chi * cn;
    P.px += -(a * P.x);
    P.py += a * P.y;
}
// (order known at compile time: a sextupole / octupole kick without the loop -- the run-time
// loop below was 26 % branches and 28 % integer instructions, ncu r02e2)
template <int ORDER, bool CHI1, class S>
__device__ __forceinline__ void mult_kick_pn_c(S& P, const double cn) {
    const double x = P.x, y = P.y;
    const double a = CHI1 ? cn : P.chi * cn;
    double dpx = a * x, dpy = a * y;
#pragma unroll
    for (int i = 2; i <= ORDER; ++i) {
        const double zre = dpx * x - dpy * y;
        const double zim = dpx * y + dpy * x;
        dpx = zre;
        dpy = zim;
    }
    P.px += -dpx;
    P.py += dpy;
}
template <bool CHI1, class S>
__device__ __forceinline__ void mult_kick_pn(S& P, const double cn, const uint32_t order) {
    const double x = P.x, y = P.y;
    const double a = CHI1 ? cn : P.chi * cn;
    double dpx = a * x, dpy = a * y;
    for (uint32_t i = 2; i <= order; ++i) {
        const double zre = dpx * x - dpy * y;
        const double zim = dpx * y + dpy * x;
        dpx = zre;
        dpy = zim;
    }
    P.px += -dpx;
    P.py += dpy;
}

// mult_kick_h0 with cs_0 == 0: `py += 0` left out
template <bool FRZ, bool CHI1, class S>
__device__ __forceinline__ void mult_kick_h0n(S& P, const double hl, const double b0,
                                              const double cn0) {
    const double x = P.x, chi = P.chi;
    const double dpx_mul = CHI1 ? cn0 : chi * cn0;
    P.px += -dpx_mul;
    double dpx = hl * (1. + P.delta);
    const double dzeta = -P.rv0v * hl * x;
    dpx += (CHI1 ? b0 : chi * b0) * x;
    P.px += dpx;
    if (!FRZ) P.zeta += dzeta;
}

// mult_kick_h with order 1, the k1*h term, and cs_1 = cs_0 = 0 (combined-function magnet):
// Horner step with the zero operations left out, then track_magnet_kick.h:98-142
template <bool FRZ, bool CHI1, class S>
__device__ __forceinline__ void mult_kick_h1n(S& P, const double hl, const double b0,
                                              const double b1c, const double cn1,
                                              const double cn0) {
    const double x = P.x, y = P.y, chi = P.chi;
    const double a = CHI1 ? cn1 : chi * cn1;
    const double dpx_mul = (CHI1 ? cn0 : chi * cn0) + a * x;
    const double dpy_mul = a * y;
    P.px += -dpx_mul;
    P.py += dpy_mul;
    double dpx = hl * (1. + P.delta);
    const double dzeta = -P.rv0v * hl * x;
    dpx += (CHI1 ? b0 : chi * b0) * x;
    const double b1 = CHI1 ? b1c : chi * b1c;
    dpx += b1 * (-x * x + 0.5 * y * y);
    P.px += dpx;
    P.py += b1 * x * y;
    if (!FRZ) P.zeta += dzeta;
}

// DipoleEdgeLinear_single_particle (track_dipole_edge_linear.h:30-39), fast-op form
template <bool CHI1, class S>
__device__ __forceinline__ void edge_linear_c(S& P, const double r21, const double r43) {
    P.px += (CHI1 ? r21 : P.chi * r21) * P.x;
    P.py += (CHI1 ? r43 : P.chi * r43) * P.y;
}

// outside the global aperture?  (negation of the test in global_aperture_check, NaN -> outside)
template <class S>
__device__ __forceinline__ bool outside_global(const S& P, const double lim) {
    return !((P.x >= -lim) && (P.x <= lim) && (P.y >= -lim) && (P.y <= lim));
}
