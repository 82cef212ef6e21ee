// xtb_interp.cuh -- the op interpreter: monitors, rarely used ops, and the loop that
// executes a range of program words on the particles one thread carries.
//
// Shared verbatim by the CUDA kernel (xtb_kernel.cuh) and by the host build of
// the device code that the CPU test tier uses (tests/hostsim), so that the op
// semantics checked against the oracle without a GPU are the ones the kernel runs.
//
// Element loop semantics restated from the generated `track_line` kernel,
// xtrack/tracker.py:646-711: per element  track -> [global aperture check for
// statically thick classes :681-689] -> is-active check :702-705 (a lost particle
// stops; at_element stays on the element where it was lost) -> at_element++ :707-711.
#pragma once
#include "xtb_state.cuh"
#include "xtb_thin.cuh"
#ifdef XTB_WITH_HEAVY
#include "xtb_thick.cuh"
#endif

// ParticlesMonitor record: LocalParticle_to_Particles(part, data, store_at, 0)
// of monitors/particles_monitor.h:13-77 -- all 32 per-particle fields.
static __device__ __noinline__ void monitor_store(const xtb_monitor_t& m, const int64_t at,
                                                  const PState& P, const PSlot& G) {
    auto D = [&](int f) { return reinterpret_cast<double*>(m.field[f]) + at; };
    auto I = [&](int f) { return reinterpret_cast<int64_t*>(m.field[f]) + at; };
    auto U = [&](int f) { return reinterpret_cast<uint32_t*>(m.field[f]) + at; };
    *D(F_P0C) = G.ld(F_P0C);  *D(F_GAMMA0) = G.ld(F_GAMMA0);  *D(F_BETA0) = G.ld(F_BETA0);
    *D(F_S) = P.s;  *D(F_ZETA) = P.zeta;  *D(F_X) = P.x;  *D(F_Y) = P.y;
    *D(F_PX) = P.px;  *D(F_PY) = P.py;  *D(F_PTAU) = G.ld(F_PTAU);  *D(F_DELTA) = P.delta;
    *D(F_RPP) = P.rpp;  *D(F_RVV) = P.rvv;  *D(F_CHI) = P.chi;
    *D(F_CHARGE_RATIO) = G.ld(F_CHARGE_RATIO);  *D(F_WEIGHT) = G.ld(F_WEIGHT);
    *D(F_AX) = G.ld(F_AX);  *D(F_AY) = G.ld(F_AY);  *D(F_SPIN_X) = G.ld(F_SPIN_X);
    *D(F_SPIN_Y) = G.ld(F_SPIN_Y);  *D(F_SPIN_Z) = G.ld(F_SPIN_Z);  *D(F_ANOM) = G.ld(F_ANOM);
    *I(F_PDG_ID) = G.ldi(F_PDG_ID);  *I(F_PARTICLE_ID) = G.ldi(F_PARTICLE_ID);
    *I(F_AT_ELEMENT) = (int64_t) P.at_element;  *I(F_AT_TURN) = P.at_turn;
    *I(F_STATE) = (int64_t) P.state;  *I(F_PARENT_ID) = G.ldi(F_PARENT_ID);
    *U(F_RNG_S1) = G.ldu(F_RNG_S1);  *U(F_RNG_S2) = G.ldu(F_RNG_S2);
    *U(F_RNG_S3) = G.ldu(F_RNG_S3);  *U(F_RNG_S4) = G.ldu(F_RNG_S4);
}

// monitors/particles_monitor.h:13-77 (P.at_element must hold the CURRENT element index)
static __device__ __noinline__ void monitor_record(const xtb_monitor_t& m, const PState& P,
                                                   const PSlot& G) {
    const int64_t n_turns_record = m.stop_at_turn - m.start_at_turn;
    const int64_t at_turn = m.ebe_mode ? (int64_t) P.at_element : P.at_turn;
    const int64_t particle_id = G.ldi(F_PARTICLE_ID);
    if (m.n_repetitions == 1) {
        if (at_turn >= m.start_at_turn && at_turn < m.stop_at_turn
            && particle_id < m.part_id_end && particle_id >= m.part_id_start) {
            monitor_store(m, n_turns_record * (particle_id - m.part_id_start) + at_turn - m.start_at_turn,
                          P, G);
        }
    } else if (m.n_repetitions > 1) {
        if (at_turn < m.start_at_turn) return;
        const int64_t i_frame = (at_turn - m.start_at_turn) / m.repetition_period;
        if (i_frame < m.n_repetitions && at_turn >= m.start_at_turn + i_frame * m.repetition_period
            && at_turn < m.stop_at_turn + i_frame * m.repetition_period
            && particle_id < m.part_id_end && particle_id >= m.part_id_start) {
            monitor_store(m,
                          n_turns_record * (m.part_id_end - m.part_id_start) * i_frame
                              + n_turns_record * (particle_id - m.part_id_start)
                              + (at_turn - i_frame * m.repetition_period) - m.start_at_turn,
                          P, G);
        }
    }
}

// LastTurnsMonitor_track_local_particle, monitors/last_turns_monitor.h:16-55
static __device__ __noinline__ void last_turns_record(const xtb_last_turns_monitor_t& m,
                                                      const PState& P, const PSlot& G) {
    const int64_t particle_id = G.ldi(F_PARTICLE_ID);
    const int64_t at_turn = P.at_turn;
    const int64_t stop = m.particle_id_start + m.num_particles;
    if (at_turn >= 0 && at_turn % m.every_n_turns == 0 && m.particle_id_start <= particle_id
        && particle_id < stop) {
        const int64_t offset = (at_turn / m.every_n_turns) % m.n_last_turns;
        const int64_t ip = particle_id - m.particle_id_start;
        const int64_t slot = m.n_last_turns * ip + offset;
        reinterpret_cast<uint32_t*>(m.field[0])[ip] = (uint32_t) offset;
        reinterpret_cast<uint32_t*>(m.field[1])[slot] = (uint32_t) particle_id;
        reinterpret_cast<uint32_t*>(m.field[2])[slot] = (uint32_t) at_turn;
        reinterpret_cast<float*>(m.field[3])[slot] = (float) P.x;
        reinterpret_cast<float*>(m.field[4])[slot] = (float) P.px;
        reinterpret_cast<float*>(m.field[5])[slot] = (float) P.y;
        reinterpret_cast<float*>(m.field[6])[slot] = (float) P.py;
        reinterpret_cast<float*>(m.field[7])[slot] = (float) P.delta;
        reinterpret_cast<float*>(m.field[8])[slot] = (float) P.zeta;
    }
}

// BeamPositionMonitor / BeamSizeMonitor, monitors/beam_position_monitor.h:16-58 and
// beam_size_monitor.h:16-64: per time slot, the count and the sums of x, y (and x^2, y^2)
// of the particles that cross the monitor in that slot.
//   q: (int) start_at_turn, (int) particle_id_start, (int) particle_id_stop, frev,
//      sampling_frequency, (int) n_slots, (ptr) record = [count | x_sum | y_sum | x2_sum |
//      y2_sum], n_slots doubles each;  aux = number of sums (3: position, 5: size)
// The reference adds with one atomicAdd per particle and quantity.  Here the lanes of a warp
// that fall into the same slot (a bunch: all of them) are summed in the warp first and ONE
// lane adds to memory: 5 atomics per warp instead of 160 to five addresses everyone wants.
// (The order of a floating-point sum is free in the reference too: OpenMP / GPU atomics.)
static __device__ __noinline__ void beam_monitor_record(const double* __restrict__ q, const int32_t n_sums,
                                                        const PState& P, const PSlot& G) {
    const int64_t start_at_turn = __double_as_longlong(q[0]);
    const int64_t id_start = __double_as_longlong(q[1]), id_stop = __double_as_longlong(q[2]);
    const double frev = q[3], sampling_frequency = q[4];
    const int64_t max_slot = __double_as_longlong(q[5]);
    double* rec = reinterpret_cast<double*>(__double_as_longlong(q[6]));
    const int64_t particle_id = G.ldgi(F_PARTICLE_ID);
    int64_t slot = -1;
    if (id_stop < 0 || (id_start <= particle_id && particle_id < id_stop)) {
        const double at_turn = (double) P.at_turn;
        const double beta0 = G.ld(F_BETA0);
        slot = (int64_t) round(sampling_frequency
                               * ((at_turn - start_at_turn) / frev - P.zeta / beta0 / XTB_C_LIGHT));
        if (!(slot >= 0 && slot < max_slot)) slot = -1;
    }
    double v[5] = {1.0, P.x, P.y, P.x * P.x, P.y * P.y};
#ifdef __CUDA_ARCH__
    const unsigned active = __activemask();
    int same = 0;
    __match_all_sync(active, (long long) slot, &same);
    if (same) {
        if (slot < 0) return;
        const int leader = __ffs(active) - 1;
        double sum[5] = {0., 0., 0., 0., 0.};
        for (unsigned m = active; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            for (int j = 0; j < 5; ++j)
                if (j < n_sums) sum[j] += __shfl_sync(active, v[j], src);
        }
        if ((int) (threadIdx.x & 31) == leader)
            for (int j = 0; j < n_sums; ++j) atomicAdd(rec + j * max_slot + slot, sum[j]);
        return;
    }
    if (slot < 0) return;
    for (int j = 0; j < n_sums; ++j) atomicAdd(rec + j * max_slot + slot, v[j]);
#else
    if (slot < 0) return;
    for (int j = 0; j < n_sums; ++j) rec[j * max_slot + slot] += v[j];
#endif
}

// BeamStatsMonitor, monitors/beam_stats_monitor.h:11-419: weighted primitive moments (sum of
// weights, weighted sums of the 9 first and 27 second moments, optional profiles) per logged
// turn -- of the whole beam, per bunch slot, per slice of a bunch, or per slice of a full turn
// (coasting beam).  q[0] is the DEVICE address of the monitor's descriptor, 8-byte words
// (xtrack_b200/monitors.py:BeamStatsMonitor.allocate):
//   0 start_at_turn  1 stop_at_turn  2 every_n_turns  3 mode  4 n_records  5 n_selected
//   6 n_slices  7 z_min_edge (f64)  8 dzeta (f64)  9 bunch_spacing_zeta (f64)  10, 11 particle
//   id range (start < 0: all)  12 len(slot_to_selected)  13 selected_slots[0]
//   14 &touched_records  15 n_profiles  16 &profile_counts
//   17 .. 54 addresses of the 38 record arrays in the reference's field order (0: not kept)
//   then slot_to_selected[], then per profile {offset, num_bins, coord_id, min, bin_width}.
// The reference adds with one atomicAdd per particle and quantity; here the lanes of a warp
// that fall into the same bin (a bunch: all of them) are summed in the warp and ONE lane adds.
#define XTB_BSM_HEADER 17
#define XTB_BSM_NFIELDS 38
static __device__ __noinline__ void beam_stats_record(const double* __restrict__ q, const PState& P,
                                                      const PSlot& G, const XtbTrackArgs& a) {
    const int64_t* __restrict__ d = reinterpret_cast<const int64_t*>(__double_as_longlong(q[0]));
    const double* __restrict__ df = reinterpret_cast<const double*>(d);
    const int64_t start_at_turn = d[0], stop_at_turn = d[1], every_n_turns = d[2], mode = d[3];
    const int64_t n_records = d[4], n_selected = d[5], n_slices = d[6];
    const double z_min_edge = df[7], dzeta = df[8], bunch_spacing_zeta = df[9];
    const int64_t pid_start = d[10], pid_stop = d[11], n_s2s = d[12];
    const bool coasting = (mode == 3);
    const int64_t* __restrict__ s2s = d + XTB_BSM_HEADER + XTB_BSM_NFIELDS;

    int64_t effective_turn = P.at_turn;
    double zeta = P.zeta;
    int64_t coasting_slice = 0;
    bool accepted = true;
    if (pid_start >= 0) {
        const int64_t particle_id = G.ldgi(F_PARTICLE_ID);
        if (particle_id < pid_start || particle_id >= pid_stop) accepted = false;
    }
    if (accepted && coasting) {
        const double line_length = a.line_length;
        if (line_length <= 0.0) {
            accepted = false;
        } else {
            const double u = ((double) effective_turn - zeta / line_length);
            effective_turn = (int64_t) floor(u + 0.5);
            const double relative_turn_fraction = u - (double) effective_turn;
            const double slice_position = (relative_turn_fraction + 0.5) * (double) n_slices;
            coasting_slice = (int64_t) floor(slice_position);
            if (coasting_slice < 0 || coasting_slice >= n_slices) accepted = false;
            zeta = -relative_turn_fraction * line_length;
        }
    }
    const int64_t turn_offset = effective_turn - start_at_turn;
    int64_t index = -1, i_record = -1;
    if (accepted && effective_turn >= start_at_turn && effective_turn < stop_at_turn
        && turn_offset % every_n_turns == 0) {
        i_record = turn_offset / every_n_turns;
        if (i_record >= 0 && i_record < n_records) {
            index = i_record;
            if (coasting) {
                index = (i_record * n_selected) * n_slices + coasting_slice;
            } else if (mode > 0) {
                int64_t slot = 0, i_selected = 0;
                if (bunch_spacing_zeta != 0.0) {
                    slot = (int64_t) -floor((zeta - z_min_edge) / bunch_spacing_zeta);
                    i_selected = (slot >= 0 && slot < n_s2s) ? s2s[slot] : -1;
                } else if (n_selected == 1) {
                    slot = d[13];
                    i_selected = 0;
                } else {
                    i_selected = -1;
                }
                if (i_selected < 0) {
                    index = -1;
                } else if (mode == 1) {
                    index = i_record * n_selected + i_selected;
                } else {
                    const double z_min_edge_bunch = z_min_edge - slot * bunch_spacing_zeta;
                    const int64_t i_slice = (int64_t) floor((zeta - z_min_edge_bunch) / dzeta);
                    if (i_slice < 0 || i_slice >= n_slices) index = -1;
                    else index = (i_record * n_selected + i_selected) * n_slices + i_slice;
                }
            }
        }
    }
    // the weighted quantities, in the order of the record arrays
    double v[XTB_BSM_NFIELDS];
    double coords[7];
    double weight = 0.;
    if (index >= 0) {
        weight = G.ldg(F_WEIGHT);
        const double beta0 = G.ld(F_BETA0);
        const double charge_ratio = G.ld(F_CHARGE_RATIO);
        coords[0] = P.x;  coords[1] = P.px;  coords[2] = P.y;  coords[3] = P.py;  coords[4] = zeta;
        coords[5] = P.delta;  coords[6] = G.ld(F_PTAU) / beta0;
        v[0] = weight;
        v[1] = weight * (beta0 * G.ld(F_GAMMA0));
        for (int j = 0; j < 7; ++j) v[2 + j] = weight * coords[j];
        v[9] = weight * charge_ratio;
        v[10] = weight * (charge_ratio / P.chi);
        int k = 11;
        for (int i = 0; i < 7; ++i)
            for (int j = i; j < 7; ++j) {
                if (i == 5 && j == 6) continue;              // (no delta_pzeta moment)
                v[k++] = weight * coords[i] * coords[j];
            }
    } else {
        for (int j = 0; j < XTB_BSM_NFIELDS; ++j) v[j] = 0.;
    }
    int64_t* touched = reinterpret_cast<int64_t*>(d[14]);
#ifdef __CUDA_ARCH__
    const unsigned active = __activemask();
    int same = 0;
    __match_all_sync(active, (long long) index, &same);
    if (same) {
        if (index >= 0) {
            const int leader = __ffs(active) - 1;
            const bool lead = ((int) (threadIdx.x & 31) == leader);
            if (lead) touched[i_record] = 1;
            for (int j = 0; j < XTB_BSM_NFIELDS; ++j) {
                double* dst = reinterpret_cast<double*>(d[XTB_BSM_HEADER + j]);
                if (!dst) continue;                          // (uniform: the descriptor's)
                double sum = 0.;
                for (unsigned m = active; m; m &= m - 1) sum += __shfl_sync(active, v[j], __ffs(m) - 1);
                if (lead) atomicAdd(dst + index, sum);
            }
        }
    } else if (index >= 0) {
        touched[i_record] = 1;
        for (int j = 0; j < XTB_BSM_NFIELDS; ++j) {
            double* dst = reinterpret_cast<double*>(d[XTB_BSM_HEADER + j]);
            if (dst) atomicAdd(dst + index, v[j]);
        }
    }
#else
    if (index >= 0) {
        touched[i_record] = 1;
        for (int j = 0; j < XTB_BSM_NFIELDS; ++j) {
            double* dst = reinterpret_cast<double*>(d[XTB_BSM_HEADER + j]);
            if (dst) dst[index] += v[j];
        }
    }
#endif
    // weighted profiles (bins are spread over the beam: plain per-particle adds)
    const int64_t n_profiles = d[15];
    if (index >= 0 && n_profiles > 0) {
        double* counts = reinterpret_cast<double*>(d[16]);
        const int64_t* pw = s2s + n_s2s;
        for (int64_t ip = 0; ip < n_profiles; ++ip, pw += 5) {
            const int64_t off = pw[0], nb = pw[1], cid = pw[2];
            const double vmin = reinterpret_cast<const double*>(pw)[3];
            const double width = reinterpret_cast<const double*>(pw)[4];
            const double value = (cid >= 0 && cid < 7) ? coords[cid] : 0.0;
            const int64_t ib = (int64_t) floor((value - vmin) / width);
            if (ib >= 0 && ib < nb) {
#ifdef __CUDA_ARCH__
                atomicAdd(counts + off + index * nb + ib, weight);
#else
                counts[off + index * nb + ib] += weight;
#endif
            }
        }
    }
}

// BeamProfileMonitor, monitors/beam_profile_monitor.h:15-80: per time sample a histogram of
// x and one of y (particle counts per bin; the bins are spread over the beam, so the plain
// per-particle atomic add of the reference is kept).
//   q: (int) start_at_turn, (int) particle_id_start, (int) particle_id_stop, frev,
//      sampling_frequency, (int) sample_size, (int) nx, x_min, dx, (int) ny, y_min, dy,
//      (ptr) counts_x [sample_size * nx], (ptr) counts_y [sample_size * ny]
static __device__ __noinline__ void beam_profile_record(const double* __restrict__ q, const PState& P,
                                                        const PSlot& G) {
    const int64_t start_at_turn = __double_as_longlong(q[0]);
    const int64_t id_start = __double_as_longlong(q[1]), id_stop = __double_as_longlong(q[2]);
    const double frev = q[3], sampling_frequency = q[4];
    const int64_t max_sample = __double_as_longlong(q[5]);
    const int64_t nx = __double_as_longlong(q[6]), ny = __double_as_longlong(q[9]);
    const double x_min = q[7], dx = q[8], y_min = q[10], dy = q[11];
    double* counts_x = reinterpret_cast<double*>(__double_as_longlong(q[12]));
    double* counts_y = reinterpret_cast<double*>(__double_as_longlong(q[13]));
    const int64_t particle_id = G.ldgi(F_PARTICLE_ID);
    if (!(id_stop < 0 || (id_start <= particle_id && particle_id < id_stop))) return;
    const double at_turn = (double) P.at_turn;
    const double beta0 = G.ld(F_BETA0);
    const int64_t sample = (int64_t) round(sampling_frequency
                                           * ((at_turn - start_at_turn) / frev - P.zeta / beta0 / XTB_C_LIGHT));
    if (!(sample >= 0 && sample < max_sample)) return;
    const int64_t bin_x = (int64_t) floor((P.x - x_min) / dx);
    const int64_t bin_y = (int64_t) floor((P.y - y_min) / dy);
    // (slot < len(counts) of the reference holds by construction: len = sample_size * n)
#ifdef __CUDA_ARCH__
    if (bin_x >= 0 && bin_x < nx) atomicAdd(counts_x + sample * nx + bin_x, 1.0);
    if (bin_y >= 0 && bin_y < ny) atomicAdd(counts_y + sample * ny + bin_y, 1.0);
#else
    if (bin_x >= 0 && bin_x < nx) counts_x[sample * nx + bin_x] += 1.0;
    if (bin_y >= 0 && bin_y < ny) counts_y[sample * ny + bin_y] += 1.0;
#endif
}

// Generic thin ops, kept out of line so that they do not weigh on the hot loop's
// register allocation.  `P.at_element` holds the current element index here.
// BMON: the beam-monitor ops are compiled in.  They are cold, yet their mere presence in the
// thin kernels cost the hot loop 3.6 % (EXACT variant, same handler SASS, gpurun_out/r01t14;
// not a matter of the hot function's alignment, gpurun_out/r01pad), so lattices without such
// monitors -- the production case -- run the instantiation that does not contain them.
template <bool FRZ, bool BMON>
static __device__ __noinline__ void generic_op(const uint32_t op, const int32_t aux,
                                               const double* __restrict__ q, PState& P,
                                               const PSlot& G, const XtbTrackArgs& a,
                                               const bool live) {
    switch (op) {
    case XTB_OP_DRIFT:
        drift_expanded<FRZ>(P, q[0]);
        break;
    case XTB_OP_DRIFT_EXACT:
        drift_exact<FRZ>(P, q[0]);
        break;
    case XTB_OP_MULT:
        mult_kick(P, q, aux);
        break;
    case XTB_OP_MULT_H:
        mult_kick_h<FRZ>(P, q, q + 4, aux & 0xff, (aux >> 8) & 1);
        break;
    case XTB_OP_CAVITY:
        cavity_kick<FRZ>(P, G, a, q[0], q[1], q[2], q[3], q[4], aux);
        break;
    case XTB_OP_RFMULT:
        rfmult_kick<FRZ>(P, G, a, q, aux);
        break;
    case XTB_OP_CRAB:
        crab_kick<FRZ>(P, G, a, q, aux);
        break;
    case XTB_OP_EDGE_LIN:
        edge_linear(P, q[0], q[1]);
        break;
    case XTB_OP_SROT:
        srotation(P, q[0], q[1]);
        break;
    case XTB_OP_XYSHIFT:
        P.x += -q[0];
        P.y += -q[1];
        break;
    case XTB_OP_SSHIFT:
        drift_exact<FRZ>(P, q[0]);
        if (!FRZ) { P.zeta += -q[0];  P.s += -q[0]; }
        break;
    case XTB_OP_YROT:
        yrotation<FRZ>(P, G, q[0], q[1], q[2]);
        break;
    case XTB_OP_XROT:
        xrotation<FRZ>(P, G, q[0], q[1], q[2]);
        break;
    case XTB_OP_LIMIT_RECT:
        if (!a.ignore_local) {
            const bool in = (P.x >= q[0]) && (P.x <= q[1]) && (P.y >= q[2]) && (P.y <= q[3]);
            if (!in) P.state = 0;
        }
        break;
    case XTB_OP_LIMIT_ELLIPSE:
        if (!a.ignore_local) {
            const double temp = P.x * P.x * q[1] + P.y * P.y * q[0];
            if (!(temp <= q[2])) P.state = 0;
        }
        break;
    case XTB_OP_LIMIT_POLYGON:
        if (!a.ignore_local && !polygon_contains(P.x, P.y, q, q + aux, aux)) P.state = 0;
        break;
    case XTB_OP_MONITOR:
        if (live) monitor_record(a.inline_mon[aux], P, G);
        break;
    case XTB_OP_LAST_TURNS:
        if (live) last_turns_record(a.inline_ltm[aux], P, G);
        break;
    case XTB_OP_BEAM_MON:
        if constexpr (BMON) { if (live) beam_monitor_record(q, aux, P, G); }
        break;
    case XTB_OP_BEAM_PROFILE:
        if constexpr (BMON) { if (live) beam_profile_record(q, P, G); }
        break;
    case XTB_OP_BEAM_STATS:
        if constexpr (BMON) { if (live) beam_stats_record(q, P, G, a); }
        break;
    case XTB_OP_KILL:
        kill_particle<FRZ>(P, G, aux);
        break;
    case XTB_OP_SET_STATE:
        P.state = aux;
        break;
    case XTB_OP_ADD_S_ZETA:
        if (!FRZ) { P.s += q[0];  P.zeta += q[0]; }
        break;
    case XTB_OP_ADD_X:
        P.x += q[0];
        break;
    default:
        break;
    }
}

// A lane without a particle to track executes the ops on this benign on-axis state
// (no per-op predication in the hot loop); nothing of it is ever written back.
template <class S>
__device__ __forceinline__ void pstate_benign(S& P) {
    P.x = P.px = P.y = P.py = P.zeta = P.delta = P.s = 0.;
    P.rpp = P.rv0v = P.chi = 1.;
    P.state = 1;
}
__device__ __forceinline__ void pstate_benign(PState& P) {
    P.x = P.px = P.y = P.py = P.zeta = P.delta = P.s = 0.;
    P.rpp = P.rvv = P.rv0v = P.chi = 1.;
    P.state = 1;
    P.at_turn = 0;
    P.at_element = 0;
}

// ---- cold paths, out of line: their register needs must not weigh on the hot loop ----
// one generic / heavy op on one lane; returns false when the particle was lost and stored
template <bool HEAVY, bool SYNRAD, bool FRZ, bool BMON, class S>
static __device__ __noinline__ bool xtb_slow_op(S& Pk, const PSlot Gk, const XtbPass ps,
                                                const uint32_t eidx, const uint32_t h,
                                                const int32_t aux, const double* __restrict__ q,
                                                const XtbTrackArgs& a) {
    const uint32_t op = h & 0xffu;
    bool alive = true;
    PState T = pstate_full(Pk, Gk, ps, eidx);
    if ((a.flag_monitor == 2) && (h & (XTB_F_START << 8))) monitor_record(a.mon, T, Gk);
#ifdef XTB_WITH_HEAVY
    if (HEAVY && op >= XTB_HEAVY_FIRST) heavy_op<SYNRAD, FRZ>(op, aux, q, T, Gk, a);
    else
#endif
        generic_op<FRZ, BMON>(op, aux, q, T, Gk, a, true);
    if ((h & (XTB_F_GLOBAL << 8)) && !a.ignore_global) global_aperture_check(T, a.global_xy_limit);
    if ((h & (XTB_F_END << 8)) && T.state <= 0) {
        // tracker.py:702-711: a lost particle stops here, at_element stays on the
        // element where it was lost
        pstate_store(T, Gk);
        alive = false;
        T.state = 1;
    }
    pstate_back(Pk, T, Gk);
    return alive;
}

#ifndef XTB_UNROLL_RUNS
#define XTB_UNROLL_RUNS 0      /* measured slower on B200 (0.538 vs 0.559): instruction footprint */
#endif
#ifndef XTB_VOLATILE_PARAMS
#define XTB_VOLATILE_PARAMS 0
#endif
#ifndef XTB_HOT_PAD
#define XTB_HOT_PAD 0
#endif

struct __align__(16) xtb_w128 { uint64_t x, y; };
struct __align__(16) xtb_d2 { double x, y; };

// Tile access.  On the device a tile is addressed by its 32-bit shared-window address
// and read with explicit ld.shared (no 64-bit generic pointer arithmetic in the hot
// loop); the host build of this file reads through an ordinary pointer.
#ifdef __CUDA_ARCH__
typedef uint32_t xtb_tile_t;
__device__ __forceinline__ xtb_tile_t xtb_tile_of(const uint64_t* p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ xtb_w128 xtb_ld_w(const xtb_tile_t tb, const uint32_t off) {
    xtb_w128 r;
    // (volatile: keeps the load where it is written -- ahead of the arithmetic that covers
    // its latency; otherwise it is sunk to its first use)
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "r"(tb + off * 8u));
    return r;
}
__device__ __forceinline__ xtb_d2 xtb_ld_d(const xtb_tile_t tb, const uint32_t off) {
    xtb_d2 r;
#if XTB_VOLATILE_PARAMS
    // ld.volatile: ptxas may not sink the load below the (never taken) loss branch that ends
    // the drift prefix, so the parameters of the main op are fetched while the prefix computes
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(tb + off * 8u));
#else
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(tb + off * 8u));
#endif
    return r;
}
__device__ __forceinline__ const uint64_t* xtb_tile_ptr(const xtb_tile_t tb, const uint32_t off) {
    return reinterpret_cast<const uint64_t*>(__cvta_shared_to_generic(tb + off * 8u));
}
#else
typedef const uint64_t* xtb_tile_t;
static inline xtb_tile_t xtb_tile_of(const uint64_t* p) { return p; }
static inline xtb_w128 xtb_ld_w(const xtb_tile_t tb, const uint32_t off) {
    return *reinterpret_cast<const xtb_w128*>(tb + off);
}
static inline xtb_d2 xtb_ld_d(const xtb_tile_t tb, const uint32_t off) {
    return *reinterpret_cast<const xtb_d2*>(tb + off);
}
static inline const uint64_t* xtb_tile_ptr(const xtb_tile_t tb, const uint32_t off) { return tb + off; }
#endif

#ifndef XTB_GLOBAL_FILTER
#define XTB_GLOBAL_FILTER 2
#endif
#ifndef XTB_UNLIKELY
#define XTB_UNLIKELY(c) __builtin_expect(!!(c), 0)
#endif

// The particles one thread carries, as they sit in (thread-local) memory between tiles.
template <int NPT, class S>
struct XtbLanes {
    S P[NPT];
    PCold C[NPT];        // cached cold fields of lane k's particle (xtb_state.cuh)
    uint32_t slot[NPT];  // index of lane k's particle in the caller's SoA
    bool live[NPT];      // lane k still tracks a real, active particle
    uint32_t eidx;       // elements completed so far in this pass (identical for all threads)
    uint32_t off;        // word offset in the tile of the op to execute next
};

// why xtb_run_fast came back
enum XtbStop {
    XTB_STOP_END = 0,        // sentinel reached: tile done
    XTB_STOP_SLOW,           // the op at `off` is a generic / heavy op (not executed)
    XTB_STOP_GLOBAL_PREFIX,  // drift prefix of the op at `off` done; some lane is outside the global limit
    XTB_STOP_GLOBAL_MAIN,    // XTB_OP_FDRIFT at `off` done; some lane is outside the global limit
    XTB_STOP_RECT,           // XTB_OP_RECT at `off`: some lane is outside the aperture
    XTB_STOP_ELLIPSE         // XTB_OP_ELLIPSE at `off`: some lane is outside the aperture
};

#ifdef XTB_COUNT_STOPS
static __device__ unsigned long long xtb_dbg_stops[8];     // returns of xtb_run_fast by reason
#ifndef __CUDA_ARCH__
#define atomicAdd(p, v) (*(p) += (v))
#endif
#endif

// The hot loop.  Executes fast ops from `lb->off` on until something happens that is
// not fast-path work (XtbStop) and returns what; the caller (xtb_run_tile) deals with
// it and calls again.  The function
//   * is deliberately NOT inlined and contains nothing but the fast handlers -- no global
//     memory access, no call, no cold code -- so that ptxas allocates registers for the
//     hot handlers alone (with the cold paths in the same function it kept spilling the
//     tile address and offset, the two values every handler needs first);
//   * loads the lanes from `lb` into registers at entry and writes them back at exit
//     (~50 instructions per call, against ~10^4 per tile);
//   * is threaded code: each handler fetches its parameters and the header of the NEXT op
//     first (XTB_FETCH: their shared-memory latency is covered by the arithmetic), then
//     branches straight to the handler of the next op (XTB_NEXT).  One taken branch per op,
//     no loop back-edge, no bounds test (the tile ends at a sentinel op), no second
//     dispatch for the drift prefix (prefixed forms have their own handlers).
// The loss tests are single not-taken branches; the exact test and the bookkeeping are
// redone by the caller.  `skip_prefix`: the drift prefix of the first op was already done.
// SUNI: every lane of the block entered the launch with the same s (bitwise), so s is the
// same number on every lane for ever (every op adds the same element constants in the same
// order): it is carried ONCE per thread -- one DADD per drift instead of NPT, and no
// register pair per particle for it.
// APF: the fast aperture ops test an integer box first (XTB_INSIDE_FOR_SURE).  A separate
// instantiation, chosen per launch (XtbTrackArgs::aperture_prefilter: does the program hold
// such ops?): compiled into the one hot function it cost a lattice WITHOUT apertures 0.6 %
// (register allocation of the other handlers; sessions 25, 26).
template <int NPT, bool FRZ, bool CHI1, bool SUNI, bool APF, class S>
static __device__ __noinline__ int xtb_run_fast(const xtb_tile_t tb, XtbLanes<NPT, S>* __restrict__ lb,
                                                const uint32_t lim_hi, const int skip_prefix) {
    S P[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        P[k] = lb->P[k];
        if (CHI1) P[k].chi = 1.0;      // known constant on this path: no register for it
    }
    uint32_t eidx = lb->eidx;
    uint32_t off = lb->off;
    int stop;
    double s_u = lb->P[0].s;
    // Code placement.  The throughput of the handlers below moves by +-3.6 % with their
    // position relative to the instruction-cache lines (measured: adding 40 KB of COLD code
    // elsewhere in the kernel took the EXACT variant from 1.284e12 to 1.239e12 PET/s with the
    // handlers' SASS unchanged, gpurun_out/r01t14).  XTB_HOT_PAD single-instruction no-ops
    // (16 bytes each, executed once per call) shift the whole function body; the value is
    // swept on the GPU whenever the kernel's code changes (scripts/gpu_pad_sweep.sh,
    // profiles/r01_history.md).
#ifdef __CUDA_ARCH__
#if XTB_HOT_PAD >= 1
    asm volatile("pmevent 0;");
#endif
#if XTB_HOT_PAD >= 2
    asm volatile("pmevent 0;");
#endif
#if XTB_HOT_PAD >= 3
    asm volatile("pmevent 0;");
#endif
#if XTB_HOT_PAD >= 4
    asm volatile("pmevent 0;");
#endif
#if XTB_HOT_PAD >= 5
    asm volatile("pmevent 0;");
#endif
#if XTB_HOT_PAD >= 6
    asm volatile("pmevent 0;");
#endif
#if XTB_HOT_PAD >= 7
    asm volatile("pmevent 0;");
#endif
#endif

    uint32_t h, op, cur;
    xtb_d2 c0, c1;
    xtb_w128 hw, hwn;
    double L;

    // "is any lane outside the global aperture?" (XTB_GLOBAL_FILTER, see xtb_run_tile)
#if XTB_GLOBAL_FILTER == 2
#define XTB_ANY_OUTSIDE(any)                                                 \
    {                                                                        \
        uint32_t m_ = 0;                                                     \
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) {                    \
            m_ = max(m_, (uint32_t) __double2hiint(P[k].x) & 0x7fffffffu);   \
            m_ = max(m_, (uint32_t) __double2hiint(P[k].y) & 0x7fffffffu);   \
        }                                                                    \
        any = m_ >= lim_hi;                                                  \
    }
#else
#define XTB_ANY_OUTSIDE(any)                                                 \
    {                                                                        \
        const double lim_ = __hiloint2double((int) lim_hi, 0);              \
        any = false;                                                         \
        _Pragma("unroll") for (int k = 0; k < NPT; ++k)                      \
            any = any | !((fabs(P[k].x) < lim_) && (fabs(P[k].y) < lim_));   \
    }
#endif
    // fast aperture ops: is every lane inside the box the op's aux word stands for (lowering.py::
    // _inside_for_sure)?  Integer compares of the high words of |x|, |y|; aux 0: never.
#define XTB_INSIDE_FOR_SURE(sure)                                            \
    if constexpr (!APF) { sure = false; } else                               \
    {                                                                        \
        const uint32_t aux_ = (uint32_t) (hw.x >> 32);                       \
        uint32_t mx_ = 0, my_ = 0;                                           \
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) {                    \
            mx_ = max(mx_, (uint32_t) __double2hiint(P[k].x) & 0x7fffffffu); \
            my_ = max(my_, (uint32_t) __double2hiint(P[k].y) & 0x7fffffffu); \
        }                                                                    \
        sure = (mx_ < (aux_ & 0xffff0000u)) & (my_ < (aux_ << 16));          \
    }
#define XTB_DRIFT(LEN)                                                       \
    if (SUNI) {                                                              \
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) drift_expanded_nos<FRZ>(P[k], LEN); \
        if (!FRZ) s_u += LEN;                                                \
    } else {                                                                 \
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(P[k], LEN); \
    }
#define XTB_FETCH()                                                          \
    h = (uint32_t) hw.x;                                                     \
    L = __longlong_as_double((long long) hw.y);                              \
    cur = off;                                                               \
    off += (h >> 16);                                                        \
    c0 = xtb_ld_d(tb, cur + 2);                                              \
    c1 = xtb_ld_d(tb, cur + 4);                                              \
    hwn = xtb_ld_w(tb, off);
#define XTB_D(OPC) ((OPC) | XTB_OPBIT_DRIFT)
#define XTB_ADV()                                                            \
    hw = hwn;                                                                \
    op = (uint32_t) hw.x & 0xffu;
#define XTB_NEXT()                                                           \
    XTB_ADV()                                                                \
    XTB_DISPATCH()
#define XTB_DISPATCH()                                                       \
    if (op == XTB_D(XTB_OP_MULTH0N)) goto H_D_MULTH0N;                       \
    if (op == XTB_D(XTB_OP_MULTP1)) goto H_D_MULTP1;                         \
    if (op == XTB_D(XTB_OP_EDGE)) goto H_D_EDGE;                             \
    if (op == XTB_D(XTB_OP_RECT)) goto H_D_RECT;                             \
    if (op == XTB_D(XTB_OP_MULTPN)) goto H_D_MULTPN;                         \
    if (op == XTB_OP_MULTH0N) goto H_MULTH0N;                                \
    goto L_SWITCH;
    // the Drift element in front of an op: track, global check (-> caller), at_element++
#define XTB_PREFIX()                                                         \
    {                                                                        \
        XTB_DRIFT(L)                                                         \
        bool any_;                                                           \
        XTB_ANY_OUTSIDE(any_)                                                \
        if (XTB_UNLIKELY(any_)) { stop = XTB_STOP_GLOBAL_PREFIX;  off = cur;  goto L_STOP; } \
        eidx += 1;                                                           \
    }
#define XTB_HANDLER(NAME, ...)                                               \
    H_##NAME: { XTB_FETCH(); { __VA_ARGS__ } eidx += 1; XTB_NEXT(); }        \
    H_D_##NAME: { XTB_FETCH(); XTB_PREFIX(); { __VA_ARGS__ } eidx += 1; XTB_NEXT(); }
    // Two thirds of the dispatches of a thin-sliced ring go to the handler that is running
    // (runs of bend / quadrupole slices): the prefixed handler of the hottest ops carries a
    // second copy of itself that is entered by FALLING THROUGH when the next op is of the
    // same kind -- one taken branch per two ops instead of one per op.
#if XTB_UNROLL_RUNS
#define XTB_HANDLER_X2(NAME, ...)                                            \
    H_##NAME: { XTB_FETCH(); { __VA_ARGS__ } eidx += 1; XTB_NEXT(); }        \
    H_D_##NAME: { XTB_FETCH(); XTB_PREFIX(); { __VA_ARGS__ } eidx += 1;      \
                  XTB_ADV()                                                  \
                  if (op == XTB_D(XTB_OP_##NAME)) {                          \
                      XTB_FETCH(); XTB_PREFIX(); { __VA_ARGS__ } eidx += 1;  \
                      XTB_NEXT();                                            \
                  }                                                          \
                  XTB_DISPATCH(); }
#else
#define XTB_HANDLER_X2(NAME, ...) XTB_HANDLER(NAME, __VA_ARGS__)
#endif

    hwn = xtb_ld_w(tb, off);
    hw = hwn;
    op = (uint32_t) hw.x & 0xffu;
    if (skip_prefix) op &= ~(uint32_t) XTB_OPBIT_DRIFT;     // (fast ops only: see xtb_run_tile)
    goto L_SWITCH;

    XTB_HANDLER(NOP, )
    XTB_HANDLER(MULT0, {
        const double c[2] = {c0.x, c0.y};
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) mult_kick_c<0, CHI1>(P[k], c);
    })
    XTB_HANDLER(MULT1, {
        const double c[4] = {c0.x, c0.y, c1.x, c1.y};
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) mult_kick_c<1, CHI1>(P[k], c);
    })
    XTB_HANDLER(MULTN, {
        // order >= 2: Horner loop, one coefficient pair per step from shared memory
        // (same arithmetic as mult_kick_c, not unrolled: small register footprint)
        const uint32_t order = (uint32_t) (hw.x >> 32);
        double dpx[NPT], dpy[NPT];
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) {
            dpx[k] = CHI1 ? c0.x : P[k].chi * c0.x;
            dpy[k] = CHI1 ? c0.y : P[k].chi * c0.y;
        }
        xtb_d2 cc = c1;
        for (uint32_t i = 1; i <= order; ++i) {
            const xtb_d2 cn = xtb_ld_d(tb, cur + 4 + 2 * i);
            _Pragma("unroll") for (int k = 0; k < NPT; ++k) {
                const double zre = dpx[k] * P[k].x - dpy[k] * P[k].y;
                const double zim = dpx[k] * P[k].y + dpy[k] * P[k].x;
                dpx[k] = (CHI1 ? cc.x : P[k].chi * cc.x) + zre;
                dpy[k] = (CHI1 ? cc.y : P[k].chi * cc.y) + zim;
            }
            cc = cn;
        }
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) {
            P[k].px += -dpx[k];
            P[k].py += dpy[k];
        }
    })
    XTB_HANDLER(MULTH0, {
        _Pragma("unroll") for (int k = 0; k < NPT; ++k)
            mult_kick_h0<FRZ, CHI1>(P[k], c0.x, c0.y, c1.x, c1.y);
    })
    XTB_HANDLER_X2(MULTH0N, {
        _Pragma("unroll") for (int k = 0; k < NPT; ++k)
            mult_kick_h0n<FRZ, CHI1>(P[k], c0.x, c0.y, c1.x);
    })
    XTB_HANDLER(MULTH1N, {
        const xtb_d2 c2 = xtb_ld_d(tb, cur + 6);
        _Pragma("unroll") for (int k = 0; k < NPT; ++k)
            mult_kick_h1n<FRZ, CHI1>(P[k], c0.x, c0.y, c1.x, c1.y, c2.x);
    })
    XTB_HANDLER_X2(MULTP1, {
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) mult_kick_p1<CHI1>(P[k], c0.x);
    })
    XTB_HANDLER(MULTPN, {
        const uint32_t order = (uint32_t) (hw.x >> 32);
#ifdef XTB_NO_PN_SPECIAL      /* (A/B switch of the tuning sessions) */
        if (false) {
#else
        if (order == 2) {
#endif
            _Pragma("unroll") for (int k = 0; k < NPT; ++k) mult_kick_pn_c<2, CHI1>(P[k], c0.x);
#ifndef XTB_NO_PN_SPECIAL
        } else if (order == 3) {
            _Pragma("unroll") for (int k = 0; k < NPT; ++k) mult_kick_pn_c<3, CHI1>(P[k], c0.x);
#endif
        } else {
            _Pragma("unroll") for (int k = 0; k < NPT; ++k) mult_kick_pn<CHI1>(P[k], c0.x, order);
        }
    })
    XTB_HANDLER_X2(EDGE, {
        _Pragma("unroll") for (int k = 0; k < NPT; ++k) edge_linear_c<CHI1>(P[k], c0.x, c0.y);
    })
    XTB_HANDLER(RECT, {
        // (a beam well inside the chamber never reaches the 4 DSETP per particle of the exact
        // test: the FP64 pipe is the bound of this kernel, the integer compare is free)
        bool sure;
        XTB_INSIDE_FOR_SURE(sure)
        if (!sure) {
            bool any = false;
            _Pragma("unroll") for (int k = 0; k < NPT; ++k)
                any = any | !((P[k].x >= c0.x) && (P[k].x <= c0.y) && (P[k].y >= c1.x)
                              && (P[k].y <= c1.y));
            if (XTB_UNLIKELY(any)) { stop = XTB_STOP_RECT;  off = cur;  goto L_STOP; }
        }
    })
    XTB_HANDLER(ELLIPSE, {
        bool sure;
        XTB_INSIDE_FOR_SURE(sure)
        if (!sure) {
            bool any = false;
            _Pragma("unroll") for (int k = 0; k < NPT; ++k)
                any = any | !(P[k].x * P[k].x * c0.y + P[k].y * P[k].y * c0.x <= c1.x);
            if (XTB_UNLIKELY(any)) { stop = XTB_STOP_ELLIPSE;  off = cur;  goto L_STOP; }
        }
    })
    XTB_HANDLER(FDRIFT, {
        XTB_DRIFT(c0.x)
        bool any;
        XTB_ANY_OUTSIDE(any)
        if (XTB_UNLIKELY(any)) { stop = XTB_STOP_GLOBAL_MAIN;  off = cur;  goto L_STOP; }
    })

L_SWITCH:
    switch (op) {
#define XTB_ROUTE(NAME)                                  \
    case XTB_OP_##NAME: goto H_##NAME;                   \
    case XTB_D(XTB_OP_##NAME): goto H_D_##NAME;
    XTB_ROUTE(NOP) XTB_ROUTE(MULT0) XTB_ROUTE(MULT1) XTB_ROUTE(MULTN) XTB_ROUTE(MULTH0)
    XTB_ROUTE(EDGE) XTB_ROUTE(RECT) XTB_ROUTE(ELLIPSE) XTB_ROUTE(FDRIFT)
    XTB_ROUTE(MULTPN) XTB_ROUTE(MULTP1) XTB_ROUTE(MULTH0N) XTB_ROUTE(MULTH1N)
#undef XTB_ROUTE
    case XTB_OP_END:
        stop = XTB_STOP_END;
        break;
    default:
        stop = XTB_STOP_SLOW;
        break;
    }
#undef XTB_FETCH
#undef XTB_DRIFT
#undef XTB_NEXT
#undef XTB_ADV
#undef XTB_DISPATCH
#undef XTB_HANDLER_X2
#undef XTB_PREFIX
#undef XTB_HANDLER
#undef XTB_ANY_OUTSIDE
#undef XTB_D
L_STOP:
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        if (SUNI) P[k].s = s_u;
        lb->P[k] = P[k];
    }
    lb->eidx = eidx;
    lb->off = off;
    return stop;
}

#ifdef XTB_WITH_HEAVY
// The run loop of the thick lattices.  A thick ring is a sequence of magnet bodies and
// drifts (LEP: 92 % of the ops); going back to xtb_run_fast / xtb_run_tile between them cost
// a fifth of the kernel in state round trips through thread-local memory (ncu,
// profiles/r01_ncu_lep.md).  This loop keeps the lanes in registers over a whole RUN of
//   XTB_OP_MAGNET_BODY (any flags, drift prefix), XTB_OP_FDRIFT (with / without prefix) and
//   generic XTB_OP_DRIFT ops,
// the bodies on all lanes at once (xtb_thick.cuh::magnet_body_n), and returns at the first op
// of another kind (lanes.off points at it).  Element bookkeeping exactly as xtb_slow_op /
// xtb_run_tile: global aperture check after statically thick elements, loss check and
// at_element + 1 at the end of an element (tracker.py:681-711); a lost particle is stored at
// once and its lane goes on, benign.
// THIN (radiation kernels): the run loop of a thin radiating ring (CLIC-DR: drift, edge, wiggler
// pole, edge, drift ...) -- the same loop with the kick-only bodies alone (thin_rad_kick_run,
// the beam constants formed once per run); it returns at the first thick body.  Without the
// thick maps in the function the lanes stay in registers (the general loop spills ~300 bytes
// around them, and its thin bodies waited on thread-local memory: profiles/r02_history.md).
template <int NPT, bool SYNRAD, bool FRZ, bool THIN>
static __device__ __noinline__ void xtb_run_heavy(const xtb_tile_t tb, XtbLanes<NPT, PState>& lanes,
                                                  const XtbPass ps, const XtbTrackArgs& a) {
    PState T[NPT];
    PSlot G[NPT];
    bool live[NPT];
    int32_t ae_base[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        G[k].p = &a.part;  G[k].i = lanes.slot[k];  G[k].c = &lanes.C[k];
        live[k] = lanes.live[k];
        T[k] = lanes.P[k];
        ae_base[k] = 0;
        if (live[k])
            ae_base[k] = (ps.el_reset ? 0 : (int32_t) G[k].ldi(F_AT_ELEMENT)) + (int32_t) ps.el_off;
    }
#ifdef __CUDA_ARCH__
    // The lane array stays ADDRESSABLE, i.e. in thread-local memory that the L1 serves, and the
    // registers go to the temporaries of the maps.  Measured (profiles/r02_history.md, section 7):
    // with the lanes promoted to registers ptxas spills 0.6 - 9 KB per thread under the
    // 128-register cap and every workload of these loops gets slower (LEP thick 1.61 -> 1.48e10
    // PET/s, CLIC-DR mean 1.08 -> 0.81e11).  Whether the array is promoted used to hinge on
    // which out-of-line map happened to take a lane by address; this pins it.
    asm volatile("" ::"l"(&T[0]) : "memory");
#endif
    uint32_t off = lanes.off, eidx = lanes.eidx;
    // launch constants read once (through `a` they are loads the compiler will not hoist over
    // the stores of the loop: ncu showed long-scoreboard stalls on them per op)
    const double lim = a.global_xy_limit;
    const bool ignore_global = a.ignore_global != 0;
    const bool ebe_monitor = (a.flag_monitor == 2);
    ThinRadRun<NPT> rad_run;
    if (THIN) thin_rad_run_begin<NPT>(rad_run, G, a);

    // lane k is lost in the current element (index eidx): write it back, go on benign
    auto retire = [&](const int k) {
        T[k].at_turn = G[k].ldi(F_AT_TURN) + ps.turn_inc;
        T[k].at_element = ae_base[k] + (int32_t) eidx;
        pstate_store(T[k], G[k]);
        live[k] = false;
        lanes.live[k] = false;
        pstate_benign(T[k]);
    };
    // end of a Drift element: global aperture check, loss check, at_element + 1.  As in the
    // thin kernels an integer test on the high words of x, y comes first (|x| >= lim / 2 implies
    // hi32(|x|) >= hi32(lim / 2); NaN passes it too): the exact test of the reference, the
    // write-back of a lost particle and the re-pinning of a lane without particle that
    // wandered off (see xtb_run_tile) only run when some lane is beyond half the limit.
    const uint32_t half_lim_hi = (lim > 0.) ? (uint32_t) __double2hiint(0.5 * lim) : 0u;
    auto end_thick = [&]() {
        uint32_t m_ = 0;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            m_ = max(m_, (uint32_t) __double2hiint(T[k].x) & 0x7fffffffu);
            m_ = max(m_, (uint32_t) __double2hiint(T[k].y) & 0x7fffffffu);
        }
#ifdef XTB_NO_HEAVY_PREFILTER      /* (A/B switch of the tuning sessions) */
        m_ = 0xffffffffu;
#endif
        if (m_ >= half_lim_hi) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                if (!live[k]) {
                    if (!(fabs(T[k].x) < 0.5 * lim && fabs(T[k].y) < 0.5 * lim)) pstate_benign(T[k]);
                    continue;
                }
                if (!ignore_global) global_aperture_check(T[k], lim);
                if (T[k].state <= 0) retire(k);
            }
        }
        eidx += 1;
    };

    for (;;) {
        const xtb_w128 hw = xtb_ld_w(tb, off);
        const uint32_t h = (uint32_t) hw.x;
        const uint32_t op = h & 0xffu;
        const double L = __longlong_as_double((long long) hw.y);
        if (op == XTB_OP_MAGNET_BODY || op == XTB_OP_DRIFT) {
            const int32_t aux = (int32_t) (hw.x >> 32);
            if (THIN && op == XTB_OP_MAGNET_BODY && !is_thin_kick_body((uint32_t) aux)) break;
            const double* __restrict__ q = reinterpret_cast<const double*>(xtb_tile_ptr(tb, off + 2));
            if (h & (XTB_F_DRIFT << 8)) {           // the Drift element in front
#pragma unroll
                for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(T[k], L);
                end_thick();
            }
            if (ebe_monitor && (h & (XTB_F_START << 8))) {
                for (int k = 0; k < NPT; ++k)
                    if (live[k]) {
                        T[k].at_turn = G[k].ldi(F_AT_TURN) + ps.turn_inc;
                        T[k].at_element = ae_base[k] + (int32_t) eidx;
                        monitor_record(a.mon, T[k], G[k]);
                    }
            }
            if (op == XTB_OP_MAGNET_BODY) {
                if constexpr (THIN) thin_rad_kick_run<NPT, FRZ>(T, live, G, a, body_par(q, aux), rad_run);
                else magnet_body_n<NPT, SYNRAD, FRZ>(T, live, G, a, q, aux);
            } else {
                const double len = q[0];
#pragma unroll
                for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(T[k], len);
            }
            if (h & (XTB_F_END << 8)) {
#pragma unroll
                for (int k = 0; k < NPT; ++k) {
                    if (!live[k]) {
                        // (a lane without particle: back on the axis once it has wandered off)
                        if (!(fabs(T[k].x) < 0.5 * lim && fabs(T[k].y) < 0.5 * lim)) pstate_benign(T[k]);
                        T[k].state = 1;
                        continue;
                    }
                    if ((h & (XTB_F_GLOBAL << 8)) && !ignore_global) global_aperture_check(T[k], lim);
                    if (T[k].state <= 0) retire(k);
                }
                eidx += 1;
            } else if ((h & (XTB_F_GLOBAL << 8)) && !ignore_global) {
#pragma unroll
                for (int k = 0; k < NPT; ++k)
                    if (live[k]) global_aperture_check(T[k], lim);
            }
        } else if (op == XTB_OP_FDRIFT || op == (XTB_OP_FDRIFT | XTB_OPBIT_DRIFT)) {
            if (op & XTB_OPBIT_DRIFT) {
#pragma unroll
                for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(T[k], L);
                end_thick();
            }
            const double len = __longlong_as_double((long long) xtb_ld_w(tb, off + 2).x);
#pragma unroll
            for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(T[k], len);
            end_thick();
        } else if ((op & ~(uint32_t) XTB_OPBIT_DRIFT) < XTB_NUM_FAST && !ebe_monitor
                   && (op & ~(uint32_t) XTB_OPBIT_DRIFT) != XTB_OP_RECT
                   && (op & ~(uint32_t) XTB_OPBIT_DRIFT) != XTB_OP_ELLIPSE) {
            // the thin fast ops between the bodies (dipole edges, markers, thin correctors ...):
            // a radiating thin ring (CLIC-DR: drift, edge, wiggler pole, edge, drift ...) went
            // back to xtb_run_fast for every one of them -- two round trips of the lanes through
            // thread-local memory per pole.  Same arithmetic as the handlers of xtb_run_fast.
            if (op & XTB_OPBIT_DRIFT) {
#pragma unroll
                for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(T[k], L);
                end_thick();
            }
            const uint32_t fop = op & ~(uint32_t) XTB_OPBIT_DRIFT;
            const double* __restrict__ c = reinterpret_cast<const double*>(xtb_tile_ptr(tb, off + 2));
            const uint32_t order = (uint32_t) (hw.x >> 32);
#define XTB_EACH_LANE _Pragma("unroll") for (int k = 0; k < NPT; ++k)
            switch (fop) {           // (one decode for all lanes)
            case XTB_OP_MULT0: { const double cc[2] = {c[0], c[1]};  XTB_EACH_LANE mult_kick_c<0, false>(T[k], cc);  break; }
            case XTB_OP_MULT1: { const double cc[4] = {c[0], c[1], c[2], c[3]};  XTB_EACH_LANE mult_kick_c<1, false>(T[k], cc);  break; }
            case XTB_OP_MULTN: XTB_EACH_LANE mult_kick(T[k], c, (int) order);  break;
            case XTB_OP_MULTPN: XTB_EACH_LANE mult_kick_pn<false>(T[k], c[0], order);  break;
            case XTB_OP_MULTP1: XTB_EACH_LANE mult_kick_p1<false>(T[k], c[0]);  break;
            case XTB_OP_MULTH0: XTB_EACH_LANE mult_kick_h0<FRZ, false>(T[k], c[0], c[1], c[2], c[3]);  break;
            case XTB_OP_MULTH0N: XTB_EACH_LANE mult_kick_h0n<FRZ, false>(T[k], c[0], c[1], c[2]);  break;
            case XTB_OP_MULTH1N: XTB_EACH_LANE mult_kick_h1n<FRZ, false>(T[k], c[0], c[1], c[2], c[3], c[4]);  break;
            case XTB_OP_EDGE: XTB_EACH_LANE edge_linear_c<false>(T[k], c[0], c[1]);  break;
            default: break;      // XTB_OP_NOP
            }
#undef XTB_EACH_LANE
            eidx += 1;               // (these ops cannot lose a particle)
        } else {
            break;
        }
        off += h >> 16;
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) {      // (at_turn / at_element of lanes.P are never read)
        PState& Q = lanes.P[k];
        Q.x = T[k].x;  Q.px = T[k].px;  Q.y = T[k].y;  Q.py = T[k].py;  Q.zeta = T[k].zeta;
        Q.delta = T[k].delta;  Q.rpp = T[k].rpp;  Q.rvv = T[k].rvv;  Q.rv0v = T[k].rv0v;
        Q.chi = T[k].chi;  Q.s = T[k].s;  Q.state = T[k].state;
    }
    lanes.off = off;
    lanes.eidx = eidx;
}
#endif


// Executes the ops of one tile, from word `lb->off` of the tile up to the XTB_OP_END
// sentinel, on the NPT particles of this thread: the fast ops in xtb_run_fast, everything
// else here --
//   * generic and heavy ops through their out-of-line bodies (xtb_slow_op);
//   * the loss bookkeeping when a fast-path test fired: the exact test of the reference
//     (global_aperture_check, local_particle_custom_api.h:262-289 -> state -1; LimitRect /
//     LimitEllipse -> state 0) on every lane, then for a lost particle the write-back to
//     the caller's SoA at once (pstate_store; at_element stays on the element where it was
//     lost, tracker.py:702-711).  The lane goes on as a dead lane on the benign state.
// XTB_GLOBAL_FILTER selects the fast path's "is any lane outside the global limit?":
//   2  integer pre-filter on the high words: |x| > lim implies hi32(|x|) >= hi32(lim), so
//      lanes whose high words are all below hi32(lim) are inside for sure -- no FP64
//      instruction, no false negative; the exact test here sorts out the false positives;
//   1  |x| < 2^e(lim) (2 DSETP per particle) as the pre-filter.
#ifndef XTB_RUN_TILE_INLINE
#define XTB_RUN_TILE_INLINE __forceinline__
#endif
template <int NPT, bool HEAVY, bool SYNRAD, bool FRZ, bool CHI1, bool SUNI, bool BMON, class S>
__device__ XTB_RUN_TILE_INLINE void xtb_run_tile(const xtb_tile_t tb, XtbLanes<NPT, S>& lanes,
                                             const XtbPass& ps, const XtbTrackArgs& a) {
    const double lim = a.global_xy_limit;
    // high word of the limit for the pre-filter (0: every check goes to the exact test)
    const uint32_t lim_hi = (lim > 0.) ? (uint32_t) __double2hiint(lim) : 0u;
    int skip_prefix = 0;

    // lane k lost in the current element with state code `code`
    auto retire = [&](const int k, const int32_t code) {
        const PSlot Gk{&a.part, lanes.slot[k], &lanes.C[k]};
        PState T = pstate_full(lanes.P[k], Gk, ps, lanes.eidx);
        T.state = code;
        pstate_store(T, Gk);
        lanes.live[k] = false;
        const double s_keep = lanes.P[k].s;      // (SUNI: s stays the same on every lane)
        pstate_benign(lanes.P[k]);
        lanes.P[k].s = s_keep;
    };
    // A lane without a particle keeps executing the ops from the on-axis state it was given;
    // where the reference orbit is not the axis (crossing bumps) that is a large-amplitude
    // trajectory, which leaves the limit within a turn and would then trip the pre-filter of
    // the hot loop at EVERY drift for the rest of the launch (measured: 1.3e6 returns for 37
    // lost particles in 200 turns of hllhc_14, the blocks with a loss 3.5x slower).  Such a
    // lane is put back on the axis here: one return per excursion.
    auto repin = [&](const int k) {
        const double s_keep = lanes.P[k].s;      // (SUNI: s stays the same on every lane)
        pstate_benign(lanes.P[k]);
        lanes.P[k].s = s_keep;
    };
    auto global_check = [&]() {
        for (int k = 0; k < NPT; ++k) {
            const S& Q = lanes.P[k];
            if (!lanes.live[k]) {
                // (well inside what any form of the pre-filter lets pass; catches NaN)
                if (!(fabs(Q.x) < 0.5 * lim && fabs(Q.y) < 0.5 * lim)) repin(k);
            } else if (!a.ignore_global && outside_global(Q, lim)) retire(k, -1);
        }
    };

    for (;;) {
        const int stop = a.aperture_prefilter
                             ? xtb_run_fast<NPT, FRZ, CHI1, SUNI, true>(tb, &lanes, lim_hi, skip_prefix)
                             : xtb_run_fast<NPT, FRZ, CHI1, SUNI, false>(tb, &lanes, lim_hi, skip_prefix);
        skip_prefix = 0;
#ifdef XTB_COUNT_STOPS
        atomicAdd(&xtb_dbg_stops[stop], 1ull);          // (per thread)
#endif
        if (stop == XTB_STOP_END) break;
        const xtb_w128 hw = xtb_ld_w(tb, lanes.off);
        const uint32_t h = (uint32_t) hw.x;
        const uint32_t op = h & 0xffu;
        const double L = __longlong_as_double((long long) hw.y);
        const uint32_t cur = lanes.off;
        switch (stop) {
        case XTB_STOP_GLOBAL_PREFIX:
            // the drift prefix is done: finish its element, then redo the op without it
            global_check();
            lanes.eidx += 1;
            skip_prefix = 1;
            if (op >= XTB_GENERIC_FIRST) goto slow_main;
            break;
        case XTB_STOP_GLOBAL_MAIN:
            global_check();
            lanes.eidx += 1;
            lanes.off = cur + (h >> 16);
            break;
        case XTB_STOP_RECT: {
            const xtb_d2 c0 = xtb_ld_d(tb, cur + 2), c1 = xtb_ld_d(tb, cur + 4);
            if (!a.ignore_local)
                for (int k = 0; k < NPT; ++k) {
                    const S& Q = lanes.P[k];
                    if (lanes.live[k]
                        && !((Q.x >= c0.x) && (Q.x <= c0.y) && (Q.y >= c1.x) && (Q.y <= c1.y)))
                        retire(k, 0);
                }
            lanes.eidx += 1;
            lanes.off = cur + (h >> 16);
            break;
        }
        case XTB_STOP_ELLIPSE: {
            const xtb_d2 c0 = xtb_ld_d(tb, cur + 2), c1 = xtb_ld_d(tb, cur + 4);
            if (!a.ignore_local)
                for (int k = 0; k < NPT; ++k) {
                    const S& Q = lanes.P[k];
                    if (lanes.live[k] && !(Q.x * Q.x * c0.y + Q.y * Q.y * c0.x <= c1.x)) retire(k, 0);
                }
            lanes.eidx += 1;
            lanes.off = cur + (h >> 16);
            break;
        }
        default: {      // XTB_STOP_SLOW: generic / heavy op, flags honoured
#ifdef XTB_WITH_HEAVY
            if constexpr (HEAVY && std::is_same<S, PState>::value) {
                if (op == XTB_OP_MAGNET_BODY || op == XTB_OP_DRIFT) {
                    // a whole run of them
                    if (SYNRAD && !(op == XTB_OP_MAGNET_BODY && !is_thin_kick_body((uint32_t) (hw.x >> 32))))
                        xtb_run_heavy<NPT, SYNRAD, FRZ, true>(tb, lanes, ps, a);
                    else
                        xtb_run_heavy<NPT, SYNRAD, FRZ, false>(tb, lanes, ps, a);
                    break;
                }
            }
#endif
            if (h & (XTB_F_DRIFT << 8)) {
                for (int k = 0; k < NPT; ++k) drift_expanded<FRZ>(lanes.P[k], L);
                global_check();
                lanes.eidx += 1;
            }
        slow_main:
            const int32_t aux = (int32_t) (hw.x >> 32);
            const double* __restrict__ q = reinterpret_cast<const double*>(xtb_tile_ptr(tb, cur + 2));
            if (((op == XTB_OP_CAVITY && aux != 1) || op == XTB_OP_RFMULT) && a.flag_monitor != 2) {
                // RF elements: all the lanes of the thread at once (xtb_thin.cuh::cavity_lanes)
                if (op == XTB_OP_CAVITY) cavity_lanes<NPT, FRZ>(lanes.P, lanes.C, lanes.live, q, a);
                else rfmult_lanes<NPT, FRZ>(lanes.P, lanes.C, lanes.live, q, aux, a);
                if (h & (XTB_F_GLOBAL << 8)) global_check();
                if (h & (XTB_F_END << 8)) lanes.eidx += 1;
                lanes.off = cur + (h >> 16);
                skip_prefix = 0;
                break;
            }
            for (int k = 0; k < NPT; ++k) {
                if (!lanes.live[k]) continue;      // these bodies touch the caller's SoA
                const PSlot Gk{&a.part, lanes.slot[k], &lanes.C[k]};
                lanes.live[k] = xtb_slow_op<HEAVY, SYNRAD, FRZ, BMON>(lanes.P[k], Gk, ps, lanes.eidx, h, aux, q, a);
                if (!lanes.live[k]) {
                    const double s_keep = lanes.P[k].s;
                    pstate_benign(lanes.P[k]);
                    lanes.P[k].s = s_keep;
                }
            }
            if (SUNI) {
                // dead lanes skipped the op: give them the s of the lanes that did it
                for (int k = 0; k < NPT; ++k)
                    if (lanes.live[k]) {
                        for (int j = 0; j < NPT; ++j) lanes.P[j].s = lanes.P[k].s;
                        break;
                    }
            }
            if (h & (XTB_F_END << 8)) lanes.eidx += 1;
            lanes.off = cur + (h >> 16);
            skip_prefix = 0;
            break;
        }
        }
    }
}

// End of a pass over the element range: increment_at_turn (local_particle_custom_api.h:
// 76-84) expressed on the block-uniform counters, s reset on the lanes.
template <int NPT, bool FRZ, class S>
__device__ __forceinline__ void xtb_end_pass(S (&P)[NPT], XtbPass& ps, const uint32_t eidx,
                                             const XtbTrackArgs& a) {
    if (a.flag_end_turn_actions > 0) {
        ps.turn_inc += 1;
        ps.el_off = 0;
        ps.el_reset = 1;
        if (a.flag_reset_s > 0 && !FRZ) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) P[k].s = 0.;
        }
    } else {
        ps.el_off += eidx;
    }
}
