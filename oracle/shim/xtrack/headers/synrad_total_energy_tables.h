/* ORACLE SHIM -- test infrastructure, not product code.
 *
 * Stand-in for `xtrack/headers/synrad_total_energy_tables.h`, which
 * synrad_spectrum.h:11 includes unconditionally but which is a generated blob that is
 * absent from /root/reference (listed in .MISSING_LARGE_BLOBS; its generator
 * `_generate_synrad_total_energy_tables.py` is a long offline computation).  Only the
 * `quantum-kick` radiation model (radiation_flag = 3, synrad_spectrum.h:383-460) reads
 * these tables; that model is outside the contract (DESIGN.md "Out of scope").  The
 * stub declares just enough for the header to compile: empty grids and table getters
 * that return NULL, for which the reference's own code draws nothing
 * (synrad_spectrum.h:372).  The `mean` and `quantum` models never touch it.
 */
#ifndef XTB_ORACLE_SYNRAD_TABLES_STUB_H
#define XTB_ORACLE_SYNRAD_TABLES_STUB_H

#define XTRACK_SYNRAD_TOTAL_ENERGY_DIRECT_TABLE_MAX 32
#define XTRACK_SYNRAD_TOTAL_ENERGY_TAIL_PROBABILITY_MAX 9.8e-2
#define XTRACK_SYNRAD_TOTAL_ENERGY_LEFT_OFFSET 0
#define XTRACK_SYNRAD_TOTAL_ENERGY_LEFT_SIZE 1
#define XTRACK_SYNRAD_TOTAL_ENERGY_CENTER_OFFSET 1
#define XTRACK_SYNRAD_TOTAL_ENERGY_CENTER_SIZE 1
#define XTRACK_SYNRAD_TOTAL_ENERGY_RIGHT_OFFSET 2
#define XTRACK_SYNRAD_TOTAL_ENERGY_RIGHT_SIZE 1

static const double synrad_total_energy_left_u_grid[1] = {0.0};
static const double synrad_total_energy_center_u_grid[1] = {0.5};
static const double synrad_total_energy_right_v_grid[1] = {0.0};

static inline const double* synrad_get_total_energy_log_table_power2(int64_t n){ (void) n; return 0; }
static inline const double* synrad_get_total_energy_log_table_direct32(int64_t n){ (void) n; return 0; }

#endif
