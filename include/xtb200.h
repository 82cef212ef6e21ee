/* xtb200.h -- C-ABI of libxtb200.so: the B200-native replacement for xtrack's
 * generated `track_line` kernel and the small kernels around it.
 *
 * Plain C, plain pointers and sizes, no torch / C++ types.  Every function
 * returns 0 on success or a negative error code (XTB_E_*); the message is
 * available through xtb_last_error_string().  The library never owns caller
 * memory: particle and monitor arrays are DEVICE pointers owned by the caller
 * (in the Python host: torch CUDA tensors), mutated in place, exactly as the
 * reference kernel mutates the xobjects buffers it is handed.
 *
 * Reference interfaces replaced (paths relative to the xtrack tree):
 *   xtb_lattice_create   <- TrackerData / ElementRefData construction,
 *                           xtrack/tracker_data.py:67-255 (elements frozen into one buffer)
 *   xtb_track            <- kernel `track_line`, xtrack/tracker.py:546-564 (arg list),
 *                           :793-813 (typed description); launched at :1372-1436
 *   xtb_rng_init         <- kernel `Particles_initialize_rand_gen`,
 *                           xtrack/particles/rng_src/particles_rng.h:12-28
 *   xtb_reduce_stats     <- (no reference equivalent) per-GPU partial sums that the host
 *                           all-reduces over NCCL at the end of a sharded run
 *   xtb_compact_*        <- Particles.reorganize(), xtrack/particles/particles.py:1198-1259
 *                           and LocalParticle_exchange based check_is_active,
 *                           xtrack/particles/local_particle_custom_api.h:108-164
 */
#ifndef XTB200_H
#define XTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XTB_NUM_FIELDS 32      /* per-particle arrays, order of particles.py:49-83 */

/* error codes */
#define XTB_OK               0
#define XTB_E_INVALID       -1   /* bad argument / malformed lattice blob        */
#define XTB_E_CUDA          -2   /* CUDA runtime error (see error string)       */
#define XTB_E_UNSUPPORTED   -3   /* feature outside the contract (backtrack...) */
#define XTB_E_NOMEM         -4

/* variant flags of xtb_track (compile-time switches of the reference kernel
 * that became run-time kernel variants, xtrack/tracker.py:1543-1552) */
#define XTB_VARIANT_EXACT        1u   /* no FMA contraction: arithmetic in the reference's order */
#define XTB_VARIANT_SYNRAD       2u   /* reference built WITHOUT XTRACK_MULTIPOLE_NO_SYNRAD      */
#define XTB_VARIANT_FREEZE_LONG  4u   /* FREEZE_VAR_{zeta,delta,ptau,rpp,rvv,s} (line.py:4446)   */
#define XTB_VARIANT_PLAIN_PROGRAM 8u  /* force the unfused program (diagnostics / tests)         */
#define XTB_VARIANT_PHILOX      16u   /* _rng_s1..4 hold key + draw counter of Philox4x32-10     */
                                      /* (production generator) instead of the reference's       */
                                      /* Tausworthe / LCG state (rng_src/base_rng.h:23-31)       */

/* track flags: bit positions of xtrack/track_flags.py:5-12 */
#define XTB_FLAG_BACKTRACK              0
#define XTB_FLAG_KILL_CAVITY_KICK       2
#define XTB_FLAG_IGNORE_GLOBAL_APERTURE 3
#define XTB_FLAG_IGNORE_LOCAL_APERTURE  4
#define XTB_FLAG_SR_TAPER               5
#define XTB_FLAG_SR_KICK_SAME_AS_FIRST  6

/* ParticlesData (xtrack/particles/particles.py:26-83): SoA of device pointers.
 * field[] order: p0c gamma0 beta0 s zeta x y px py ptau delta rpp rvv chi
 * charge_ratio weight ax ay spin_x spin_y spin_z anomalous_magnetic_moment (f64)
 * pdg_id particle_id at_element at_turn state parent_particle_id (i64)
 * _rng_s1.._rng_s4 (u32). */
typedef struct xtb_particles {
    int64_t capacity;
    double  q0, mass0, t_sim;
    void*   field[XTB_NUM_FIELDS];
} xtb_particles_t;

/* ParticlesMonitorData (xtrack/monitors/particles_monitor.py:180-192); `data`
 * is a ParticlesData of n_records rows, zero-initialised by the caller. */
typedef struct xtb_monitor {
    int64_t start_at_turn, stop_at_turn, part_id_start, part_id_end;
    int64_t ebe_mode, n_repetitions, repetition_period;
    void*   field[XTB_NUM_FIELDS];
} xtb_monitor_t;

/* LastTurnsMonitorData (xtrack/monitors/last_turns_monitor.py:18-44).
 * field[]: lost_at_offset, particle_id, at_turn (u32), x px y py delta zeta (f32). */
typedef struct xtb_last_turns_monitor {
    int64_t particle_id_start, num_particles, n_last_turns, every_n_turns;
    void*   field[9];
} xtb_last_turns_monitor_t;

/* per-GPU partial statistics (all sums over particles with state > 0) */
typedef struct xtb_stats {
    int64_t n_alive, n_lost;
    double  sum[6];        /* x px y py zeta delta */
    double  sum2[21];      /* upper triangle of the 6x6 second-moment matrix, row-major */
} xtb_stats_t;

typedef struct xtb_lattice* xtb_lattice_handle;

/* Upload a lowered lattice to `device`: two op streams ("programs", 8-byte words, format
 * in xtrack_b200/csrc/xtb_ops.h) of the same line -- FUSED (drift-prefixed fast ops; may
 * be NULL) and PLAIN (one element = its own ops).  `*_elem_offset[n_elements+1]` is the
 * word offset of each element's first op; in the fused program an element absorbed in
 * the op of its predecessor carries 0xffffffff.  Immutable afterwards. */
int xtb_lattice_create(const uint64_t* fused_words, size_t n_fused_words,
                       const uint32_t* fused_elem_offset,
                       const uint64_t* plain_words, size_t n_plain_words,
                       const uint32_t* plain_elem_offset, size_t n_elements,
                       double line_length, int device, xtb_lattice_handle* out);
int xtb_lattice_destroy(xtb_lattice_handle h);

/* In-line monitors referenced by OP_MONITOR / OP_LAST_TURNS ops (index = op aux). */
int xtb_lattice_set_inline_monitors(xtb_lattice_handle h,
                                    const xtb_monitor_t* mons, size_t n_mons,
                                    const xtb_last_turns_monitor_t* ltms, size_t n_ltms);

/* Inverse-CDF tables of the `quantum-kick` radiation model (magnet bodies with radiation_flag
 * 3; xtrack/headers/synrad_spectrum.h:257-406, data of the reference's generated header
 * synrad_total_energy_tables.h).  `blob` is HOST memory, copied to the lattice's device;
 * layout in doubles: [0] n_left [1] n_center [2] n_right [3] tail probability max [4] direct
 * table max (32) [5..7] 0, the left u / centre u / right v probability grids, then the tables
 * log(X_N) for N = 1..32, 64, 128, 256, each n_left + n_center + n_right long. */
int xtb_lattice_set_synrad_tables(xtb_lattice_handle h, const double* blob, size_t n_doubles);

/* One `track_line` launch; argument meaning of xtrack/tracker.py:546-564.
 * Asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream). */
int xtb_track(xtb_lattice_handle h, const xtb_particles_t* particles,
              int64_t num_turns, int32_t ele_start, int32_t num_ele_track,
              int32_t flag_end_turn_actions, int32_t flag_reset_s_at_end_turn,
              int32_t flag_monitor, const xtb_monitor_t* tbt_monitor /* nullable */,
              uint64_t track_flags, double global_xy_limit,
              uint32_t variant_flags, void* cuda_stream);

/* rng_set() of xtrack/particles/rng_src/base_rng.h:45-62 for slots [0, n). */
int xtb_rng_init(const xtb_particles_t* particles, const uint32_t* seeds_dev,
                 int64_t n, int device, void* cuda_stream);

/* Partial beam statistics of this GPU's shard into `out_dev` (device memory). */
int xtb_reduce_stats(const xtb_particles_t* particles, xtb_stats_t* out_dev,
                     int device, void* cuda_stream);

/* Loss histogram by element: hist_dev[e] += #particles with state<=0 (and
 * allocated) whose at_element == e; hist_dev has n_elements+1 int64 entries. */
int xtb_loss_histogram(const xtb_particles_t* particles, int64_t* hist_dev,
                       int64_t n_elements, int device, void* cuda_stream);

/* Stream compaction in place of the CPU reorganize(): stable partition of the
 * particle slots into [active | lost | unallocated].  `perm_dev` (capacity
 * int64) receives the source slot of each destination slot; `scratch_dev`
 * must hold xtb_compact_scratch_bytes(capacity) bytes.  counts_dev[0..1]
 * receive n_active, n_lost. */
size_t xtb_compact_scratch_bytes(int64_t capacity);
int xtb_compact(const xtb_particles_t* particles, int64_t* perm_dev,
                int64_t* counts_dev, void* scratch_dev, int device, void* cuda_stream);

/* Register-resident DFMA chain: measures this GPU's FP64 FMA peak (the
 * roofline denominator).  Returns achieved FLOP/s in *flops_out. */
int xtb_measure_dfma_peak(int device, double seconds, double* flops_out);

/* Self-test: the guard-free FP64 reciprocal / square root / division sequences of the thick
 * maps (csrc/xtb_math.cuh) against the built-in IEEE operators on `n_samples` random operands
 * with binary exponents in [-exponent_range, exponent_range].  mismatches_out[3] receives the
 * number of results that differ in any bit (rcp, sqrt, div): expected 0, 0, 0. */
int xtb_selftest_math(int device, int64_t n_samples, uint64_t seed, int exponent_range,
                      uint64_t* mismatches_out);

/* Self-test: the device sin / cos / exp / expm1 / sinh / cosh that reproduce the C library's
 * results to the bit (csrc/xtb_libm.cuh: the RF phases of cavities and RF multipoles, the
 * focusing terms of the thick quadrupole map -- their reference is glibc) evaluated on `n` HOST
 * arguments; out_host[6 n] receives sin, cos, exp, expm1, sinh, cosh (n values each) for the
 * caller to compare with its libm. */
int xtb_eval_libm(int device, const double* x_host, int64_t n, double* out_host);

/* Self-test: `n` blocks of Philox4x32-10 (counter = c0 + i, c1, 0, 0; key k0, k1) evaluated on
 * the device into the host array out[4 n] (known-answer and cross-implementation tests). */
int xtb_eval_philox(int device, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, int64_t n,
                    uint32_t* out_host);

/* Kernel launches issued by this library since load (bench bookkeeping). */
int64_t xtb_launch_count(void);

const char* xtb_last_error_string(void);
const char* xtb_version(void);

/* Version of the op-stream format this library interprets (csrc/xtb_ops.h
 * XTB_OPS_ABI_VERSION); the host lowering refuses a library that reports another one. */
int xtb_ops_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XTB200_H */
