/* ORACLE SHIM -- test infrastructure.  Stand-in for xobjects/headers/atomicadd.h
 * (included by xtrack/headers/track.h:10); nothing on the single-particle hot
 * path calls atomicAdd when in-kernel record logging is off. */
#ifndef XTB_ORACLE_XO_ATOMICADD_H
#define XTB_ORACLE_XO_ATOMICADD_H
/* The beam monitors (monitors/beam_position_monitor.h, beam_size_monitor.h) do call it: on
 * the CPU contexts xobjects defines it as a plain (OpenMP: atomic) add. */
static inline void atomicAdd(double* addr, double val) {
#ifdef XO_CONTEXT_CPU_OPENMP
#pragma omp atomic
#endif
    *addr += val;
}
#endif
