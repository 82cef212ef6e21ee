/* ORACLE SHIM -- test infrastructure, not product code.
 *
 * "libm noise" build of the oracle (variant `noise` of oracle/build_ref.py): every
 * result of a transcendental libm call in the reference's headers is moved by
 * -1, 0 or +1 ulp, chosen by a hash of the argument.  No libm promises more than
 * ~1 ulp and two libms (glibc on the host, CUDA's on the device) differ at that
 * level, so the distance between this build and the clean oracle after N turns is
 * the reference's OWN sensitivity to the libm it happens to link: the yardstick
 * against which the GPU-vs-oracle deviation of lattices with per-particle
 * sin/cos/... calls is judged (tests/test_gpu_parity.py, DESIGN.md "Parity").
 */
#ifndef XTB_ORACLE_ULP_NOISE_H
#define XTB_ORACLE_ULP_NOISE_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double xtn_nudge(double r, double x){
    if (!(r == r) || r == 0.0 || isinf(r)) return r;
    uint64_t b; memcpy(&b, &x, 8);
    b ^= b >> 29; b *= 0x9E3779B97F4A7C15ull; b ^= b >> 32;
    const unsigned sel = (unsigned)(b % 3u);
    if (sel == 0u) return r;
    return nextafter(r, sel == 1u ? INFINITY : -INFINITY);
}
static inline double xtn_sin(double x){ return xtn_nudge(sin(x), x); }
static inline double xtn_cos(double x){ return xtn_nudge(cos(x), x); }
static inline double xtn_tan(double x){ return xtn_nudge(tan(x), x); }
static inline double xtn_sinh(double x){ return xtn_nudge(sinh(x), x); }
static inline double xtn_cosh(double x){ return xtn_nudge(cosh(x), x); }
static inline double xtn_asin(double x){ return xtn_nudge(asin(x), x); }
static inline double xtn_atan(double x){ return xtn_nudge(atan(x), x); }
static inline double xtn_atan2(double y, double x){ return xtn_nudge(atan2(y, x), y + 3.0 * x); }
static inline double xtn_exp(double x){ return xtn_nudge(exp(x), x); }
static inline double xtn_log(double x){ return xtn_nudge(log(x), x); }
#define sin(x)  xtn_sin(x)
#define cos(x)  xtn_cos(x)
#define tan(x)  xtn_tan(x)
#define sinh(x) xtn_sinh(x)
#define cosh(x) xtn_cosh(x)
#define asin(x) xtn_asin(x)
#define atan(x) xtn_atan(x)
#define atan2(y, x) xtn_atan2(y, x)
#define exp(x)  xtn_exp(x)
#define log(x)  xtn_log(x)
#endif
