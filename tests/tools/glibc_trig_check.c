/* tests/tools/glibc_trig_check.c — sweeps the oracle's restatements of glibc 2.39 sincosf / atan / sin /
 * cos / sincos (oracle/glibc_trig_replica.c) against the LIVE libm of this box.
 *   gcc -O2 -ffp-contract=off -fopenmp -I../../oracle glibc_trig_check.c ../../oracle/glibc_trig_replica.c -lm
 *   ./a.out [which] ; which: 1 sincosf (all 2^32 floats), 2 atan, 4 sin/cos/sincos, default all */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void og_sincosf(float y, float* sinp, float* cosp);
double og_atan(double x);
double og_sin(double x);
double og_cos(double x);

static uint64_t rng(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static int same_d(double a, double b) { return (a != a && b != b) || memcmp(&a, &b, 8) == 0; }
static int same_f(float a, float b) { return (a != a && b != b) || memcmp(&a, &b, 4) == 0; }

int main(int argc, char** argv)
{
    const int which = argc > 1 ? atoi(argv[1]) : 7;
    long bad = 0;
    if (which & 1) {
        long nb = 0;
#pragma omp parallel for reduction(+:nb) schedule(static)
        for (long long u = 0; u < (1LL << 32); u++) {
            uint32_t ui = (uint32_t)u;
            float y, s0, c0, s1, c1;
            memcpy(&y, &ui, 4);
            sincosf(y, &s0, &c0);
            og_sincosf(y, &s1, &c1);
            if (!same_f(s0, s1) || !same_f(c0, c1)) { if (nb < 5) fprintf(stderr, "sincosf %a: %a %a vs %a %a\n", y, s0, c0, s1, c1); nb++; }
            if (!same_f(s0, sinf(y)) || !same_f(c0, cosf(y))) nb++;
        }
        printf("sincosf: all 2^32 floats, mismatches %ld\n", nb);
        bad += nb;
    }
    if (which & 2) {
        long nb = 0;
#pragma omp parallel for reduction(+:nb) schedule(static)
        for (long long k = 0; k < 400000000LL; k++) {
            uint64_t s = 0x9E3779B97F4A7C15ULL * (uint64_t)(k + 1);
            uint64_t r = rng(&s);
            double x;
            if (k & 1) { memcpy(&x, &r, 8); }                              /* any bit pattern */
            else { x = ldexp((double)(rng(&s) >> 11) / 9007199254740992.0 + 0.5, (int)(r % 70) - 35); if (r & (1ULL << 40)) x = -x; }
            if (!same_d(atan(x), og_atan(x))) { if (nb < 5) fprintf(stderr, "atan %a: %a vs %a\n", x, atan(x), og_atan(x)); nb++; }
        }
        printf("atan: 4e8 samples, mismatches %ld\n", nb);
        bad += nb;
    }
#ifdef WITH_SIN
    if (which & 4) {
        long nb = 0;
#pragma omp parallel for reduction(+:nb) schedule(static)
        for (long long k = 0; k < 400000000LL; k++) {
            uint64_t s = 0xD1B54A32D192ED03ULL * (uint64_t)(k + 1);
            uint64_t r = rng(&s);
            double x;
            const int m = (int)(k % 4);
            if (m == 0) x = ldexp((double)(rng(&s) >> 11) / 9007199254740992.0 + 0.5, (int)(r % 60) - 32);   /* 2^-33 .. 2^27 */
            else if (m == 1) x = (double)(rng(&s) >> 11) / 9007199254740992.0 * 3.2;                          /* [0, 3.2) */
            else if (m == 2) x = (double)(float)((double)(rng(&s) >> 11) / 9007199254740992.0 * 420.0);       /* float-valued, [0, 420) */
            else x = (double)(rng(&s) >> 11) / 9007199254740992.0 * 1.05e8;                         /* up to the __branred switch at 105414350 */
            if (r & (1ULL << 40)) x = -x;
            if (fabs(x) >= 105414350.0) continue;                        /* __branred territory: not restated */
            double ss, cc;
            sincos(x, &ss, &cc);
            if (!same_d(sin(x), og_sin(x))) { if (nb < 5) fprintf(stderr, "sin %a: %a vs %a\n", x, sin(x), og_sin(x)); nb++; }
            if (!same_d(cos(x), og_cos(x))) { if (nb < 5) fprintf(stderr, "cos %a: %a vs %a\n", x, cos(x), og_cos(x)); nb++; }
            if (!same_d(ss, sin(x)) || !same_d(cc, cos(x))) { if (nb < 5) fprintf(stderr, "sincos != sin,cos at %a\n", x); nb++; }
        }
        printf("sin/cos/sincos: 4e8 samples, mismatches %ld\n", nb);
        bad += nb;
    }
#endif
    return bad != 0;
}
