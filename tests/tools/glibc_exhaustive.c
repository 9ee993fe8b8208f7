/* tests/tools/glibc_exhaustive.c — exhaustive check of the float libm replicas against the live libm.
 * build: gcc -O2 -ffp-contract=off -fopenmp glibc_exhaustive.c ../../oracle/glibc_replica.c -I../../oracle -lm
 * Sweeps every non-negative float bit pattern (0 .. 0x7f800000) plus a strided sample of negatives. */
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include "oracle_common.h"
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static int same(float a, float b) { return f2u(a) == f2u(b) || (a != a && b != b); }
int main(int argc, char** argv)
{
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    long bad_log = 0, bad_log10 = 0, bad_log2 = 0;
#pragma omp parallel for reduction(+:bad_log,bad_log10,bad_log2) schedule(static)
    for (int64_t u = 0; u <= 0x7f800000LL; u += stride) {
        float x = u2f((uint32_t)u);
        if (!same(og_logf(x), logf(x))) bad_log++;
        if (!same(og_log10f(x), log10f(x))) bad_log10++;
        if (!same(og_log2f(x), log2f(x))) bad_log2++;
    }
    printf("stride %u: mismatches logf=%ld log10f=%ld log2f=%ld\n", stride, bad_log, bad_log10, bad_log2);
    return (bad_log || bad_log10 || bad_log2) ? 1 : 0;
}
