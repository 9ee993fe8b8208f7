/*
 * oracle/glibc_replica.c — TEST INFRASTRUCTURE ONLY.
 *
 * The reference calls the box's libm per frame: log10f (src/transient_detector.cpp:81),
 * log2f (src/atrac/at3/atrac3_bitstream.cpp:269, src/atrac3denc.cpp:277-286,526-527),
 * double log/exp (src/atrac/atrac_psy_common.cpp:184,194).  Third-party dependency outside
 * /root/reference: GNU libc 2.39 (Ubuntu 2.39-0ubuntu8.5), libm.so.6.  On x86-64 CPUs with
 * FMA+AVX2 its ifunc resolvers pick the "-fma" builds of e_logf.c / e_log2f.c / e_log.c /
 * e_exp.c (sysdeps/x86_64/fpu/multiarch), i.e. the published ARM optimized-routines
 * algorithms compiled with contraction.  The functions below restate those algorithms with
 * the SAME fused operations the shipped binary performs (read off `objdump -d` of
 * libm-2.39.a: e_logf-fma.o, e_log2f-fma.o, e_log-fma.o, e_exp-fma.o); log10f is the older
 * fdlibm-style wrapper around logf (sysdeps/ieee754/flt-32/e_log10f.c, not multiarch, no FMA).
 * tests/test_glibc_replica.py checks them against the live libm (exhaustively for the float
 * functions).
 */
#include "oracle_common.h"
#include "glibc239_tables.h"
#include <math.h>
#include <string.h>

static inline double u2d(unsigned long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* e_logf.c (__logf), fma build */
float og_logf(float x)
{
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u)
        return 0;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0)
            return -INFINITY;               /* __math_divzerof(1) */
        if (ix == 0x7f800000u)
            return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u)
            return (x - x) / (x - x);       /* __math_invalidf */
        ix = f2u(x * 0x1p23f);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (tmp >> 19) & 15;
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = u2d(og239_logf_data[2 * i]), logc = u2d(og239_logf_data[2 * i + 1]);
    double ln2 = u2d(og239_logf_data[32]);
    double A0 = u2d(og239_logf_data[33]), A1 = u2d(og239_logf_data[34]), A2 = u2d(og239_logf_data[35]);
    double z = (double)u2f(iz);
    double r = fma(z, invc, -1.0);
    double y0 = fma((double)k, ln2, logc);
    double r2 = r * r;
    double y = fma(r, A1, A2);
    y = fma(r2, A0, y);
    y = fma(r2, y, y0 + r);
    return (float)y;
}

/* e_log10f.c (__ieee754_log10f): float arithmetic, no contraction */
float og_log10f(float x)
{
    static const float two25 = 3.3554432000e+07f;
    const float ivln10 = u2f(0x3ede5bd9u), log10_2hi = u2f(0x3e9a2080u), log10_2lo = u2f(0x355427dbu);
    int32_t hx = (int32_t)f2u(x);
    int32_t k = 0;
    if (hx < 0x00800000) {
        if ((hx & 0x7fffffff) == 0)
            return -two25 / fabsf(x);
        if (hx < 0)
            return (x - x) / (x - x);
        k -= 25;
        x *= two25;
        hx = (int32_t)f2u(x);
    }
    if (hx >= 0x7f800000)
        return x + x;
    k += (hx >> 23) - 127;
    int32_t i = ((uint32_t)k & 0x80000000u) >> 31;
    hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
    float y = (float)(k + i);
    x = u2f((uint32_t)hx);
    float z = y * log10_2lo + ivln10 * og_logf(x);
    return z + y * log10_2hi;
}

/* e_log2f.c (__log2f), fma build */
float og_log2f(float x)
{
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u)
        return 0;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0)
            return -INFINITY;
        if (ix == 0x7f800000u)
            return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u)
            return (x - x) / (x - x);
        ix = f2u(x * 0x1p23f);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (tmp >> 19) & 15;
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)tmp >> 23;
    double invc = u2d(og239_log2f_data[2 * i]), logc = u2d(og239_log2f_data[2 * i + 1]);
    double A0 = u2d(og239_log2f_data[32]), A1 = u2d(og239_log2f_data[33]);
    double A2 = u2d(og239_log2f_data[34]), A3 = u2d(og239_log2f_data[35]);
    double z = (double)u2f(iz);
    double r = fma(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double y = fma(r, A1, A2);
    double p = fma(r, A3, y0);
    y = fma(r2, A0, y);
    y = fma(r2, y, p);
    return (float)y;
}
