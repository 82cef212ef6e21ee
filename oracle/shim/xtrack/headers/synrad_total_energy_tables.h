/* ORACLE SHIM -- test infrastructure, not product code.
 *
 * Stand-in for `xtrack/headers/synrad_total_energy_tables.h`, which synrad_spectrum.h:11
 * includes but which is a generated blob absent from /root/reference (.MISSING_LARGE_BLOBS;
 * generator: xtrack/headers/_generate_synrad_total_energy_tables.py).  The reference's header
 * holds the probability grids and the tables log(X_N) as static arrays plus two getters; here
 * the same names are POINTERS and SIZES set at run time from the blob that the product
 * uploads to the GPU (xt_ref_set_synrad_tables; layout: include/xtb200.h), so that the
 * reference's quantum-kick code (synrad_spectrum.h:257-459) runs on exactly the product's
 * data.  Without tables the getters return NULL, for which the reference's code draws
 * nothing (synrad_spectrum.h:372).
 */
#ifndef XTB_ORACLE_SYNRAD_TABLES_SHIM_H
#define XTB_ORACLE_SYNRAD_TABLES_SHIM_H

#include <stdint.h>

static const double* xtb_qk_blob = 0;

#define XTB_QK_N_LEFT   ((int64_t) (xtb_qk_blob ? xtb_qk_blob[0] : 1))
#define XTB_QK_N_CENTER ((int64_t) (xtb_qk_blob ? xtb_qk_blob[1] : 1))
#define XTB_QK_N_RIGHT  ((int64_t) (xtb_qk_blob ? xtb_qk_blob[2] : 1))

#define XTRACK_SYNRAD_TOTAL_ENERGY_DIRECT_TABLE_MAX 32
#define XTRACK_SYNRAD_TOTAL_ENERGY_TAIL_PROBABILITY_MAX (xtb_qk_blob ? xtb_qk_blob[3] : 9.8e-2)
#define XTRACK_SYNRAD_TOTAL_ENERGY_LEFT_OFFSET 0
#define XTRACK_SYNRAD_TOTAL_ENERGY_LEFT_SIZE XTB_QK_N_LEFT
#define XTRACK_SYNRAD_TOTAL_ENERGY_CENTER_OFFSET XTB_QK_N_LEFT
#define XTRACK_SYNRAD_TOTAL_ENERGY_CENTER_SIZE XTB_QK_N_CENTER
#define XTRACK_SYNRAD_TOTAL_ENERGY_RIGHT_OFFSET (XTB_QK_N_LEFT + XTB_QK_N_CENTER)
#define XTRACK_SYNRAD_TOTAL_ENERGY_RIGHT_SIZE XTB_QK_N_RIGHT

#define synrad_total_energy_left_u_grid   (xtb_qk_blob + 8)
#define synrad_total_energy_center_u_grid (xtb_qk_blob + 8 + XTB_QK_N_LEFT)
#define synrad_total_energy_right_v_grid  (xtb_qk_blob + 8 + XTB_QK_N_LEFT + XTB_QK_N_CENTER)

static inline const double* xtb_qk_table(int64_t index){
    const int64_t size = XTB_QK_N_LEFT + XTB_QK_N_CENTER + XTB_QK_N_RIGHT;
    return xtb_qk_blob + 8 + size * (1 + index);
}
static inline const double* synrad_get_total_energy_log_table_direct32(int64_t n){
    if (!xtb_qk_blob || n < 1 || n > 32) return 0;
    return xtb_qk_table(n - 1);
}
static inline const double* synrad_get_total_energy_log_table_power2(int64_t n){
    if (!xtb_qk_blob) return 0;
    switch (n){
        case 1: case 2: case 4: case 8: case 16: case 32: return xtb_qk_table(n - 1);
        case 64: return xtb_qk_table(32);
        case 128: return xtb_qk_table(33);
        case 256: return xtb_qk_table(34);
        default: return 0;
    }
}

#endif
