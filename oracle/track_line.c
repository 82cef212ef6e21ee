/* ORACLE -- test infrastructure and CPU baseline; never linked into the product.
 *
 * C restatement of the `track_line` kernel that xtrack generates as a Python
 * string (xtrack/tracker.py:545-749): the turn loop, the element loop with the
 * per-class dispatch + static-thick global aperture check (:660-697), the
 * active check / `at_element` bookkeeping (:702-711), the end-of-turn actions
 * (:727-731) and, for the OpenMP context, the chunking and the two serial
 * re-partition passes (:566-593, :741-748).  Backtracking and the
 * multi-element monitor are outside the contract and omitted.
 *
 * Element physics is NOT restated here: `xt_generated.h` includes the
 * reference's own headers from /root/reference, so this translation unit
 * compiles the reference's arithmetic unmodified.
 *
 * Build variants (oracle/build_ref.py):
 *   -DXO_CONTEXT_CPU_SERIAL               bit reference (one thread)
 *   -DXO_CONTEXT_CPU_OPENMP -fopenmp      timing baseline ("ContextCpu(omp)")
 */
#ifdef XO_CONTEXT_CPU_OPENMP
#include <omp.h>
#endif

double xtb_oracle_global_xy_limit = 1.0;       /* line.config XTRACK_GLOBAL_XY_LIMIT */
#define XTRACK_GLOBAL_XY_LIMIT xtb_oracle_global_xy_limit

#include "xt_generated.h"

void xt_ref_set_global_xy_limit(double v){ xtb_oracle_global_xy_limit = v; }

/* inverse-CDF tables of the quantum-kick radiation model: the blob the product uploads to the
 * GPU (shim/xtrack/headers/synrad_total_energy_tables.h); the caller keeps it alive */
void xt_ref_set_synrad_tables(const double* blob){
#ifdef XTB_ORACLE_SYNRAD_TABLES_SHIM_H
    xtb_qk_blob = blob;
#else
    (void) blob;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline wants every core */
void xt_ref_set_num_threads(int n){
#ifdef XO_CONTEXT_CPU_OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void) n;
#endif
}

int xt_ref_num_threads(void){
#ifdef XO_CONTEXT_CPU_OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void xt_ref_track_line(
        ParticlesData particles,
        void** elements,             /* ElementRefData: pointer per element ... */
        const int64_t* type_ids,     /* ... and its class id                    */
        int num_turns,
        int ele_start,
        int num_ele_track,
        int flag_end_turn_actions,
        int flag_reset_s_at_end_turn,
        int flag_monitor,
        int num_ele_line,
        double line_length,
        ParticlesMonitorData tbt_monitor,
        uint64_t track_flags)
{
#ifdef XO_CONTEXT_CPU_OPENMP
    const int64_t capacity = ParticlesData_get__capacity(particles);
    const int num_threads = omp_get_max_threads();
    const int64_t num_particles_to_track = ParticlesData_get__num_active_particles(particles);
    {
        LocalParticle lpart;
        lpart.io_buffer = NULL;
        Particles_to_LocalParticle(particles, &lpart, 0, capacity);
        check_is_active(&lpart);
        count_reorganized_particles(&lpart);
        LocalParticle_to_Particles(&lpart, particles, 0, capacity);
    }
    const int64_t chunk_size = (num_particles_to_track + num_threads - 1)/num_threads;
    #pragma omp parallel for
    for (int chunk = 0; chunk < num_threads; chunk++) {
    int64_t part_id = chunk * chunk_size;
    int64_t end_id = (chunk + 1) * chunk_size;
    if (end_id > num_particles_to_track) end_id = num_particles_to_track;
#else
    int64_t part_id = 0;
    int64_t end_id = 0;
    {
#endif
    LocalParticle lpart;
    lpart.io_buffer = NULL;
    lpart.track_flags = track_flags;
    lpart.line_length = line_length;

    int64_t part_capacity = ParticlesData_get__capacity(particles);
    if (part_id < part_capacity){
        Particles_to_LocalParticle(particles, &lpart, part_id, end_id);
        int64_t isactive = check_is_active(&lpart);

        for (int64_t iturn = 0; iturn < num_turns; iturn++){
            if (!isactive) break;
            int64_t const ele_stop = ele_start + num_ele_track;

            /* XS_FLAG_BACKTRACK (tracker.py:628-646,707-731): the elements in reverse order,
             * at_element counted down, the turn bookkeeping at the START of the pass */
            const int backtrack = LocalParticle_check_track_flag(&lpart, XS_FLAG_BACKTRACK);
            int64_t elem_idx, increm;
            if (backtrack){
                elem_idx = ele_stop - 1;
                increm = -1;
                if (flag_end_turn_actions > 0){
                    increment_at_turn_backtrack(&lpart, flag_reset_s_at_end_turn,
                                                line_length, num_ele_line);
                }
            } else {
                if (flag_monitor == 1){
                    ParticlesMonitor_track_local_particle(tbt_monitor, &lpart);
                }
                elem_idx = ele_start;
                increm = 1;
            }
            for (; (elem_idx >= ele_start) && (elem_idx < ele_stop); elem_idx += increm){
                if (flag_monitor == 2){
                    ParticlesMonitor_track_local_particle(tbt_monitor, &lpart);
                }
                xtb_oracle_dispatch(type_ids[elem_idx], elements[elem_idx], &lpart);

                isactive = check_is_active(&lpart);
                if (!isactive) break;
                increment_at_element(&lpart, backtrack ? -1 : 1);
            }
            if (flag_monitor == 2){
                ParticlesMonitor_track_local_particle(tbt_monitor, &lpart);
            }
            if (backtrack){
                if (flag_monitor == 1){
                    ParticlesMonitor_track_local_particle(tbt_monitor, &lpart);
                }
            } else if (flag_end_turn_actions > 0){
                if (isactive){
                    increment_at_turn(&lpart, flag_reset_s_at_end_turn);
                }
            }
        }
        LocalParticle_to_Particles(&lpart, particles, part_id, 1);
    }
    }
#ifdef XO_CONTEXT_CPU_OPENMP
    {
        LocalParticle lpart;
        lpart.io_buffer = NULL;
        Particles_to_LocalParticle(particles, &lpart, 0, capacity);
        check_is_active(&lpart);
        count_reorganized_particles(&lpart);
        LocalParticle_to_Particles(&lpart, particles, 0, capacity);
    }
#endif
}

/* Per-element kernel `X_track_particles` (base_element.py:141-214): one
 * element applied to all particles, track_flags = 0, optional at_element++. */
void xt_ref_track_element(ParticlesData particles, void* el, int64_t type_id,
                          int64_t flag_increment_at_element, double line_length,
                          uint64_t track_flags)
{
    int64_t type_ids[1] = {type_id};
    void* elements[1] = {el};
    (void) flag_increment_at_element;
    xt_ref_track_line(particles, elements, type_ids, 1, 0, 1, 0, 0, 0, 1,
                      line_length, NULL, track_flags);
}

/* `Particles_initialize_rand_gen` (particles/rng_src/particles_rng.h:12-28) is
 * compiled from the reference header included by xt_generated.h; exported as is. */
void xt_ref_init_rand_gen(ParticlesData particles, uint32_t* seeds, int n_init){
    Particles_initialize_rand_gen(particles, seeds, n_init);
}
