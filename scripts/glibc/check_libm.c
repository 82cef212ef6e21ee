/* Compares xtrack_b200/csrc/xtb_libm.cuh (host build) with the installed libm, bit for bit.
 *   g++ -O2 -mfma -ffp-contract=off -fopenmp -x c++ scripts/glibc/check_libm.c -o /tmp/check_libm -lm
 *   /tmp/check_libm [n_samples_per_range]                                                       */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../xtrack_b200/csrc/xtb_libm.cuh"

static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t bits(double v) { uint64_t u; memcpy(&u, &v, 8); return u; }

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 100000000L;
    /* ranges: every branch of __sin / __cos, dense where the RF phases of the lattices live */
    const double lo[] = {0.0, 0.126, 0.855469, 2.426265, 0.0, 0.0, 0.0, 1e4};
    const double hi[] = {0.126, 0.855469, 2.426265, 7.0, 1e-7, 100.0, 1e4, 105414350.0};
    long bad_total = 0;
    for (int r = 0; r < 8; ++r) {
        long bad_s = 0, bad_c = 0;
#pragma omp parallel for reduction(+ : bad_s, bad_c)
        for (long i = 0; i < n; ++i) {
            const uint64_t h = mix64((uint64_t) i * 0x100000001B3ull + (uint64_t) r);
            double x = lo[r] + (hi[r] - lo[r]) * ((double) (h >> 11) * 0x1p-53);
            if (h & 1) x = -x;
            if (bits(xtb_sin_glibc(x)) != bits(sin(x))) bad_s++;
            if (bits(xtb_cos_glibc(x)) != bits(cos(x))) bad_c++;
        }
        printf("range [%g, %g): %ld samples, sin mismatches %ld, cos mismatches %ld\n", lo[r], hi[r], n,
               bad_s, bad_c);
        bad_total += bad_s + bad_c;
    }
    /* exp, expm1, sinh, cosh */
    const double lo2[] = {0.0, 0.3465, 1.0397, 1e-9, 2.0, 0.0, 0.0};
    const double hi2[] = {0.3466, 1.04, 2.0, 1e-3, 21.99, 21.99, 500.0};
    for (int r = 0; r < 7; ++r) {
        long be = 0, bm = 0, bs = 0, bc = 0;
#pragma omp parallel for reduction(+ : be, bm, bs, bc)
        for (long i = 0; i < n; ++i) {
            const uint64_t h = mix64((uint64_t) i * 0x100000001B3ull + (uint64_t) (r + 100));
            double x = lo2[r] + (hi2[r] - lo2[r]) * ((double) (h >> 11) * 0x1p-53);
            if (h & 1) x = -x;
            if (bits(xtb_exp_glibc(x)) != bits(exp(x))) be++;
            if (r < 6) {
                if (bits(xtb_expm1_glibc(x)) != bits(expm1(x))) bm++;
                if (bits(xtb_sinh_glibc(x)) != bits(sinh(x))) bs++;
                if (bits(xtb_cosh_glibc(x)) != bits(cosh(x))) bc++;
            }
        }
        printf("range [%g, %g): %ld samples, mismatches exp %ld expm1 %ld sinh %ld cosh %ld\n", lo2[r],
               hi2[r], n, be, bm, bs, bc);
        bad_total += be + bm + bs + bc;
    }
    return bad_total != 0;
}
