/*
 * oracle/glibc_trig_replica.c — TEST INFRASTRUCTURE ONLY.
 *
 * The ATRAC3plus tone search of the reference (src/lib/libgha/src/gha.c, built -O2 -ffp-contract=off)
 * calls libm at run time: sincosf (gha.c:271-272, sinf+cosf of the same argument merged by gcc),
 * sincos (gha.c:176-177), atan (gha.c:228), sin (gha.c:240).  Third-party dependency outside
 * /root/reference: GNU libc 2.39 libm.so.6; on x86-64 CPUs with FMA+AVX2 the ifunc resolvers pick
 * __sincosf_fma, __sincos_fma, __atan_fma, __sin_fma (sysdeps/x86_64/fpu/multiarch).  The functions
 * below restate those algorithms (sysdeps/ieee754/flt-32/s_sincosf.c + sysdeps/x86/fpu/sincosf_poly.h,
 * sysdeps/ieee754/dbl-64/s_atan.c, s_sin.c) with the SAME fused operations the shipped binary performs,
 * read off `objdump -d` of libm-2.39.a (s_sincosf-fma.o, s_atan-fma.o, s_sin-fma.o).
 * tests/tools/glibc_trig_check.c sweeps them against the live libm.
 */
#include "oracle_common.h"
#include "glibc239_trig_tables.h"
#include <math.h>
#include <string.h>
#include <stdint.h>

static inline double u2d(unsigned long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline unsigned long long d2u(double d) { unsigned long long u; memcpy(&u, &d, 8); return u; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* ---- sincosf ------------------------------------------------------------------------------- */
/* sincosf_poly of sysdeps/x86/fpu/sincosf_poly.h, every a + b*c fused:
 *   {s1', c2'} = fma({x2,x2}, {s3,c4}, {s2,c3});  c1 = fma(x2, C1, C0)
 *   {s, c} = fma({x3,x4}, {s1,c2}, {x, c1});  {sin, cos} = fma({x5,x6}, {s1',c2'}, {s, c}) */
static void sincosf_poly(double x, double x2, const unsigned long long* p, int n, float* sinp, float* cosp)
{
    const double c0 = u2d(p[6]), c1c = u2d(p[7]);
    const double s1 = u2d(p[8]), c2 = u2d(p[9]), s2 = u2d(p[10]), c3 = u2d(p[11]), s3 = u2d(p[12]), c4 = u2d(p[13]);
    const double x3 = x2 * x, x4 = x2 * x2;
    const double s1p = fma(x2, s3, s2), c2p = fma(x2, c4, c3);
    const double c1 = fma(x2, c1c, c0);
    const double x5 = x3 * x2, x6 = x4 * x2;
    const double s = fma(x3, s1, x), c = fma(x4, c2, c1);
    const float rs = (float)fma(x5, s1p, s), rc = (float)fma(x6, c2p, c);
    if (n & 1) { *cosp = rs; *sinp = rc; } else { *sinp = rs; *cosp = rc; }
}

void og_sincosf(float y, float* sinp, float* cosp)
{
    const unsigned long long* p = og239_sincosf_table;
    const uint32_t top = (f2u(y) >> 20) & 0x7ff;
    double x = y;
    if (top < 0x3f4) {                                   /* |y| < pi/4 */
        const double x2 = x * x;
        if (top < 0x398) { *sinp = y; *cosp = 1.0f; return; }   /* |y| < 2^-12 */
        sincosf_poly(x, x2, p, 0, sinp, cosp);
    } else if (top < 0x42f) {                            /* |y| < 120: reduce_fast */
        const double r = x * u2d(p[4]);
        const int n = ((int32_t)r + 0x800000) >> 24;
        x = fma(-(double)n, u2d(p[5]), x);               /* x - n*hpi, fused (vfnmadd) */
        const double s = u2d(p[n & 3]);
        if (n & 2) p += 14;
        sincosf_poly(x * s, x * x, p, n, sinp, cosp);
    } else if (top < 0x7f8) {                            /* reduce_large */
        uint32_t xi = f2u(y);
        const int sign = xi >> 31;
        const uint32_t* arr = &og239_inv_pio4[(xi >> 26) & 15];
        const int shift = (xi >> 23) & 7;
        uint64_t n, res0, res1, res2;
        xi = (xi & 0xffffff) | 0x800000;
        xi <<= shift;
        res0 = (uint32_t)(xi * arr[0]);
        res1 = (uint64_t)xi * arr[4];
        res2 = (uint64_t)xi * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        n = (res0 + (1ULL << 61)) >> 62;
        res0 -= n << 62;
        x = (double)(int64_t)res0 * 0x1.921FB54442D18p-62;
        const int ns = (int)n + sign;
        const double s = u2d(p[ns & 3]);
        if (ns & 2) p += 14;
        sincosf_poly(x * s, x * x, p, (int)n, sinp, cosp);
    } else {
        *sinp = *cosp = y - y;
    }
}

/* ---- atan ---------------------------------------------------------------------------------- */
double og_atan(double x)
{
    static const double HPI = 0x1.921fb54442d18p0, HPI1 = 0x1.1a62633145c07p-54;
    const unsigned long long ux = d2u(x);
    if (((ux >> 32) & 0x7ff00000u) == 0x7ff00000u && (ux & 0x000fffffffffffffULL)) return x + x;
    const double d13 = u2d(0x3fb375f08b31cbceULL), d11 = u2d(0xbfb7458022b13c25ULL), d9 = u2d(0x3fbc71c6e5129a3bULL),
                 d7 = u2d(0xbfc24924923f7603ULL), d5 = u2d(0x3fc99999999997fdULL), d3 = u2d(0xbfd5555555555555ULL);
    const double A = u2d(0x3e4bb67a00000000ULL), B = 0.0625, E = u2d(0x43349ff200000000ULL);
    const double u = x < 0 ? -x : x;
    if (u < 1.0) {
        if (u < B) {
            if (u < A) return x;
            const double v = x * x;
            double yy = d13;
            yy = fma(yy, v, d11); yy = fma(yy, v, d9); yy = fma(yy, v, d7); yy = fma(yy, v, d5); yy = fma(yy, v, d3);
            return fma(x * v, yy, x);
        }
        const int i = (int)(fma(u, 256.0, 0x1p52) - 0x1p52) - 16;
        const unsigned long long* c = &og239_atan_cij[7 * i];
        const double z = u - u2d(c[0]);
        double yy = u2d(c[6]);
        yy = fma(yy, z, u2d(c[5])); yy = fma(yy, z, u2d(c[4])); yy = fma(yy, z, u2d(c[3])); yy = fma(yy, z, u2d(c[2]));
        return copysign(fma(yy, z, u2d(c[1])), x);
    }
    if (u < 16.0) {
        const double w = 1.0 / u;
        const double t1 = w * u, t2 = fma(u, w, -t1);
        const double r = (1.0 - t1) - t2;
        const int i = (int)(fma(w, 256.0, 0x1p52) - 0x1p52) - 16;
        const unsigned long long* c = &og239_atan_cij[7 * i];
        const double z = fma(r, w, w - u2d(c[0]));
        double yy = u2d(c[6]);
        yy = fma(yy, z, u2d(c[5])); yy = fma(yy, z, u2d(c[4])); yy = fma(yy, z, u2d(c[3])); yy = fma(yy, z, u2d(c[2]));
        yy = fma(-z, yy, HPI1);
        return copysign((HPI - u2d(c[1])) + yy, x);
    }
    if (u < E) {
        const double w = 1.0 / u;
        const double t1 = w * u, t3 = HPI - w, v = w * w;
        double yy = d13;
        yy = fma(yy, v, d11); yy = fma(yy, v, d9); yy = fma(yy, v, d7); yy = fma(yy, v, d5); yy = fma(yy, v, d3);
        const double cor = ((HPI - t3) - w) + HPI1;
        const double t2 = fma(u, w, -t1);
        const double r = (1.0 - t1) - t2;
        double acc = fma(-r, w, cor);
        acc = fma(-(w * v), yy, acc);
        return copysign(t3 + acc, x);
    }
    return x > 0 ? HPI : -HPI;
}

/* ---- sin / cos (sysdeps/ieee754/dbl-64/s_sin.c, fma build) ----------------------------------- */
/* |x| < 105414350 only; beyond that glibc switches to __branred, which the encoder never reaches
 * (arguments are below 2^9) — those inputs return NaN here. */
static const double k_big = 0x1.8p45, k_toint = 0x1.8p52;
#define TRIG_C(name, bits) static const unsigned long long name##_u = bits
TRIG_C(s1, 0xbfc5555555555555ULL); TRIG_C(s2, 0x3f81111111110eceULL); TRIG_C(s3, 0xbf2a01a019db08b8ULL);
TRIG_C(s4, 0x3ec71de27b9a7ed9ULL); TRIG_C(s5, 0xbe5addffc2fcdf59ULL);
TRIG_C(sn3, 0xbfc5555555555515ULL); TRIG_C(sn5, 0x3f811110e829872fULL);
TRIG_C(cs2, 0x3fe0000000000000ULL); TRIG_C(cs4, 0xbfa5555555555535ULL); TRIG_C(cs6, 0x3f56c16bedd9e239ULL);
TRIG_C(hp0, 0x3ff921fb54442d18ULL); TRIG_C(hp1, 0x3c91a62633145c07ULL);
TRIG_C(hpinv, 0x3fe45f306dc9c883ULL);
TRIG_C(mp1, 0x3ff921fb58000000ULL); TRIG_C(mp2, 0xbe4dde973c000000ULL);
TRIG_C(pp3, 0xbc8cb3b398000000ULL); TRIG_C(pp4, 0xbacd747f23e32ed7ULL);
#define K(name) u2d(name##_u)

static double taylor_sin(double xx, double x, double dx)
{
    double p = K(s5);
    p = fma(p, xx, K(s4)); p = fma(p, xx, K(s3)); p = fma(p, xx, K(s2)); p = fma(p, xx, K(s1));
    const double t = fma(xx, fma(p, x, -(dx * 0.5)), dx);
    return t + x;
}

static double do_sin(double x, double dx)
{
    const double a = fabs(x);
    if (a < u2d(0x3fc020c49ba5e354ULL))                     /* 0.126 */
        return taylor_sin(x * x, x, dx);
    if (!(0.0 < x)) dx = -dx;
    const double u = a + k_big;
    const int k = (int)((uint32_t)d2u(u) << 2);
    const double xr = a - (u - k_big);
    const double sn = u2d(og239_sincostab[k]), ssn = u2d(og239_sincostab[k + 1]);
    const double cs = u2d(og239_sincostab[k + 2]), ccs = u2d(og239_sincostab[k + 3]);
    const double xx = xr * xr;
    const double s = xr + fma(xr * xx, fma(xx, K(sn5), K(sn3)), dx);
    const double c = fma(xr, dx, xx * fma(xx, fma(xx, K(cs6), K(cs4)), K(cs2)));
    const double cor = fma(s, cs, fma(-c, sn, fma(s, ccs, ssn)));
    return copysign(sn + cor, x);
}

static double do_cos(double x, double dx)
{
    if (x < 0.0) dx = -dx;
    const double a = fabs(x);
    const double u = a + k_big;
    const int k = (int)((uint32_t)d2u(u) << 2);
    const double xr = (a - (u - k_big)) + dx;
    const double sn = u2d(og239_sincostab[k]), ssn = u2d(og239_sincostab[k + 1]);
    const double cs = u2d(og239_sincostab[k + 2]), ccs = u2d(og239_sincostab[k + 3]);
    const double xx = xr * xr;
    const double s = fma(xr * xx, fma(xx, K(sn5), K(sn3)), xr);
    const double c = xx * fma(xx, fma(xx, K(cs6), K(cs4)), K(cs2));
    const double cor = fma(-s, sn, fma(-c, cs, fma(-s, ssn, ccs)));
    return cs + cor;
}

static int reduce_sincos(double x, double* a, double* da)
{
    const double t = fma(x, K(hpinv), k_toint);
    const double xn = t - k_toint;
    const int n = (int)(d2u(t) & 3);
    const double y = fma(-xn, K(mp2), fma(-xn, K(mp1), x));
    const double t2 = fma(-xn, K(pp3), y);
    double db = fma(-xn, K(pp3), y - t2);
    const double b = fma(-xn, K(pp4), t2);
    db = db + fma(-xn, K(pp4), t2 - b);
    *a = b;
    *da = db;
    return n;
}

static double do_sincos(double a, double da, int n)
{
    const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
    return (n & 2) ? -r : r;
}

double og_sin(double x)
{
    const int k = (int)((d2u(x) >> 32) & 0x7fffffff);
    if (k < 0x3e500000) return x;
    if (k < 0x3feb6000) return do_sin(x, 0.0);
    if (k < 0x400368fd) return copysign(do_cos(K(hp0) - fabs(x), K(hp1)), x);
    if (k < 0x419921fb) {
        double a, da;
        const int n = reduce_sincos(x, &a, &da);
        return do_sincos(a, da, n);
    }
    return NAN;
}

double og_cos(double x)
{
    const int k = (int)((d2u(x) >> 32) & 0x7fffffff);
    if (k < 0x3e400000) return 1.0;
    if (k < 0x3feb6000) return do_cos(x, 0.0);
    if (k < 0x400368fd) {
        const double y = K(hp0) - fabs(x);
        const double a = y + K(hp1);
        const double da = (y - a) + K(hp1);
        return do_sin(a, da);
    }
    if (k < 0x419921fb) {
        double a, da;
        const int n = reduce_sincos(x, &a, &da);
        return do_sincos(a, da, n + 1);
    }
    return NAN;
}
