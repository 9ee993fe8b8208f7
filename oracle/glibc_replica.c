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

/* e_log.c (__log), fma build (__FP_FAST_FMA path: r = fma(z, invc, -1), no T2 table).
 * Operation order read off libm-2.39.a:e_log-fma.o. */
double og_log(double x)
{
    const unsigned long long* D = og239_log_data;   /* ln2hi, ln2lo, A[5] @2, B[11] @7, T[128]{invc,logc} @18 */
    unsigned long long ix;
    memcpy(&ix, &x, 8);
    const unsigned top = (unsigned)(ix >> 48);
    const unsigned long long LO = 0x3fee000000000000ULL, HI = 0x3ff1090000000000ULL;
    if (ix - LO < HI - LO) {
        if (ix == 0x3ff0000000000000ULL)
            return 0;
        const double* B = (const double*)(const void*)(D + 7);
        double r = x - 1.0;
        double r2 = r * r;
        double r3 = r * r2;
        double p2 = fma(r, B[2], B[1]);
        double p3 = fma(r, B[5], B[4]);
        double p5 = fma(r, B[8], B[7]);
        p2 = fma(r2, B[3], p2);
        p3 = fma(r2, B[6], p3);
        double p1 = fma(r2, B[9], p5);
        p1 = fma(r3, B[10], p1);
        p1 = fma(p1, r3, p3);
        p1 = fma(p1, r3, p2);
        double t = fma(r, 0x1p27, r);
        double rhi = fma(-0x1p27, r, t);
        double rhi2 = rhi * rhi;
        double rlo = r - rhi;
        double hi = fma(rhi2, B[0], r);
        double lo = fma(rhi2, B[0], r - hi);
        lo = fma(B[0] * rlo, r + rhi, lo);
        double y = fma(p1, r3, lo);
        return hi + y;
    }
    if (top - 0x0010 >= 0x7ff0 - 0x0010) {
        if (ix * 2 == 0)
            return -INFINITY;
        if (ix == 0x7ff0000000000000ULL)
            return x;
        if ((top & 0x8000) || (top & 0x7ff0) == 0x7ff0)
            return (x - x) / (x - x);
        double xs = x * 0x1p52;
        memcpy(&ix, &xs, 8);
        ix -= 52ULL << 52;
    }
    unsigned long long tmp = ix - 0x3fe6000000000000ULL;
    int i = (int)((tmp >> 45) & 127);
    int k = (int)((long long)tmp >> 52);
    unsigned long long iz = ix - (tmp & 0xfffULL << 52);
    double invc = u2d(D[18 + 2 * i]), logc = u2d(D[18 + 2 * i + 1]);
    double ln2hi = u2d(D[0]), ln2lo = u2d(D[1]);
    double A0 = u2d(D[2]), A1 = u2d(D[3]), A2 = u2d(D[4]), A3 = u2d(D[5]), A4 = u2d(D[6]);
    double z = u2d(iz);
    double kd = (double)k;
    double w = fma(kd, ln2hi, logc);
    double r = fma(z, invc, -1.0);
    double q = fma(r, A2, A1);
    double hi = r + w;
    double r2 = r * r;
    double lo = (w - hi) + r;
    lo = fma(kd, ln2lo, lo);
    double r3 = r * r2;
    double q2 = fma(r, A4, A3);
    lo = fma(r2, A0, lo);
    q2 = fma(q2, r2, q);
    double y = fma(r3, q2, lo);
    return y + hi;
}

/* e_exp.c (__exp), fma build.  Only the ranges the encoder can reach are restated exactly
 * (|x| < 512); overflow/underflow special cases return what libm returns for +-inf. */
double og_exp(double x)
{
    const unsigned long long* D = og239_exp_data;   /* invln2N, shift, negln2hiN, negln2loN, C2..C5 @4, ..., tab @22 */
    unsigned long long ix;
    memcpy(&ix, &x, 8);
    unsigned abstop = (unsigned)(ix >> 52) & 0x7ff;
    if (abstop - 0x3c9 >= 0x3f) {
        if (abstop - 0x3c9 >= 0x80000000u)
            return 1.0 + x;
        if (abstop >= 0x409) {
            if (ix == 0xfff0000000000000ULL) return 0.0;
            if (abstop >= 0x7ff) return 1.0 + x;
            return (ix >> 63) ? 0.0 : INFINITY;
        }
        abstop = 0;
    }
    double invln2N = u2d(D[0]), shift = u2d(D[1]), nhi = u2d(D[2]), nlo = u2d(D[3]);
    double C2 = u2d(D[4]), C3 = u2d(D[5]), C4 = u2d(D[6]), C5 = u2d(D[7]);
    double kd = fma(x, invln2N, shift);
    unsigned long long ki;
    memcpy(&ki, &kd, 8);
    kd -= shift;
    double r = fma(kd, nhi, x);
    r = fma(kd, nlo, r);
    unsigned long long idx = 2 * (ki % 128);
    unsigned long long topb = ki << 45;
    double tail = u2d(D[22 + idx]);
    unsigned long long sbits = D[22 + idx + 1] + topb;
    double p = fma(r, C3, C2);
    double t3 = r + tail;
    double r2 = r * r;
    double q = fma(r, C5, C4);
    p = fma(p, r2, t3);
    double r4 = r2 * r2;
    double tmp = fma(r4, q, p);
    if (abstop == 0) {
        /* specialcase(): huge |x|, not reachable from the encoder; approximate via scaling */
        double scale;
        if ((ki & 0x80000000) == 0) {
            sbits -= 1009ULL << 52;
            scale = u2d(sbits);
            return 0x1p1009 * (scale + scale * tmp);
        }
        sbits += 1022ULL << 52;
        scale = u2d(sbits);
        double y = scale + scale * tmp;
        return 0x1p-1022 * y;
    }
    double scale = u2d(sbits);
    return fma(scale, tmp, scale);
}
