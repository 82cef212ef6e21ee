/* ORACLE SHIM -- test infrastructure, not product code.
 *
 * Stand-in for `xobjects/headers/common.h`, which the reference's physics
 * headers include (xtrack/headers/track.h:9) but which lives in the external
 * `xobjects` package that is absent from /root/reference.  It provides the
 * CPU-context meaning of the macros the hot-path headers use (SURVEY.md
 * Appendix B).  The definitions are the natural ones; xobjects' exact text is
 * not verifiable offline and is documented as "assumed" in oracle/README.md.
 */
#ifndef XTB_ORACLE_XO_COMMON_H
#define XTB_ORACLE_XO_COMMON_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <math.h>
#ifdef XTB_ORACLE_ULP_NOISE
#include "ulp_noise.h"      /* +-1 ulp on every transcendental result: sensitivity yardstick */
#endif

#define XO_CONTEXT_CPU
#if !defined(XO_CONTEXT_CPU_SERIAL) && !defined(XO_CONTEXT_CPU_OPENMP)
#error "define XO_CONTEXT_CPU_SERIAL or XO_CONTEXT_CPU_OPENMP"
#endif

#define GPUFUN    static inline
#define GPUKERN
#define GPUGLMEM
#define RESTRICT  restrict

#define POW2(X) ((X)*(X))
#define POW3(X) ((X)*(X)*(X))
#define POW4(X) ((X)*(X)*(X)*(X))
#define NONZERO(X) ((X) != 0.0)

#define VECTORIZE_OVER(INDEX, COUNT) \
    for (int64_t INDEX = 0; INDEX < (COUNT); INDEX++) {
#define END_VECTORIZE }

#endif
