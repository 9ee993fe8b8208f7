// kissfft_dev.cuh — butterflies with exactly kissfft's arithmetic (src/lib/fft/kissfft_impl/kiss_fft.c).
//
// The reference's kf_work recursion is decimation in time; unrolled it is
//   (1) a mixed-radix digit-reversal gather (host_tables.cpp: kiss_perm), then
//   (2) the stages innermost-first; a stage of radix p and sub-length m combines elements
//       F[base + q*m], q < p, with twiddles tw[q*k*fstride], fstride = n/(p*m), base = g*p*m + k.
// Every butterfly of a stage is independent, so a stage is one parallel phase; the arithmetic
// inside a butterfly is kept in the reference's order with every product and sum rounded on its
// own (no FMA), which makes the result bit-identical to the CPU library.
#pragma once
#include "atde_cuda.h"

namespace atde {

// kf_bfly2 body, kiss_fft.c:32-41
ATDE_D void kf_bfly2(cpx& a, cpx& b, cpx tw)
{
    const cpx t = cmul(b, tw);
    b.r = fsub(a.r, t.r);
    b.i = fsub(a.i, t.i);
    a.r = fadd(a.r, t.r);
    a.i = fadd(a.i, t.i);
}

// kf_bfly4 body, kiss_fft.c:60-89 (same temporaries scratch[0..5], same order)
template <bool INVERSE>
ATDE_D void kf_bfly4(cpx& f0, cpx& f1, cpx& f2, cpx& f3, cpx t1, cpx t2, cpx t3)
{
    const cpx s0 = cmul(f1, t1);
    const cpx s1 = cmul(f2, t2);
    const cpx s2 = cmul(f3, t3);
    cpx s5, s3, s4;
    s5.r = fsub(f0.r, s1.r);  s5.i = fsub(f0.i, s1.i);
    f0.r = fadd(f0.r, s1.r);  f0.i = fadd(f0.i, s1.i);
    s3.r = fadd(s0.r, s2.r);  s3.i = fadd(s0.i, s2.i);
    s4.r = fsub(s0.r, s2.r);  s4.i = fsub(s0.i, s2.i);
    f2.r = fsub(f0.r, s3.r);  f2.i = fsub(f0.i, s3.i);
    f0.r = fadd(f0.r, s3.r);  f0.i = fadd(f0.i, s3.i);
    if (INVERSE) {
        f1.r = fsub(s5.r, s4.i);  f1.i = fadd(s5.i, s4.r);
        f3.r = fadd(s5.r, s4.i);  f3.i = fsub(s5.i, s4.r);
    } else {
        f1.r = fadd(s5.r, s4.i);  f1.i = fsub(s5.i, s4.r);
        f3.r = fsub(s5.r, s4.i);  f3.i = fadd(s5.i, s4.r);
    }
}

// ---- the same butterfly on packed fp32 pairs (FMUL2 / FFMA2): 22 issue slots instead of 34 ----
// A twiddle t is spread to (t.r, t.r | t.i, -t.i).  With a = (a.r, a.i) as one pair,
//   P = a * (t.r, t.r)  = (a.r t.r,  a.i t.r)          Q = a * (t.i, -t.i) = (a.r t.i, -(a.i t.i))
// hold C_MUL's four rounded products (a product by a negated factor is the negated product), and
//   m.r = P.x + Q.y = a.r t.r - a.i t.i,   m.i = Q.x + P.y = a.r t.i + a.i t.r
// are its two rounded sums.  The six complex additions of the butterfly are packed adds (spelled as FMAs by an opaque
// +-1, see add2), the closing four mix real and imaginary parts and stay scalar.
struct __align__(16) tw4 { float rr0, rr1, ii0, ii1; };
// (spread on the fly, two extra instructions: 16-byte table entries were measured slower — they double the shared-memory
// wavefronts of a warp load and the L1 footprint of the global tables)
ATDE_D tw4 spread_twiddle_dev(cpx t, float mone) { tw4 w; w.rr0 = t.r; w.rr1 = t.r; w.ii0 = t.i; w.ii1 = fmul(t.i, mone); return w; }
ATDE_D cpx cmul_tw(cpx a, tw4 t)
{
    f32x2 av, rr, ii;
    av.x = a.r; av.y = a.i;
    rr.x = t.rr0; rr.y = t.rr1;
    ii.x = t.ii0; ii.y = t.ii1;
    const f32x2 P = mul2(av, rr), Q = mul2(av, ii);
    cpx m;
    m.r = fadd(P.x, Q.y);
    m.i = fadd(Q.x, P.y);
    return m;
}
ATDE_D f32x2 as2(cpx a) { f32x2 v; v.x = a.r; v.y = a.i; return v; }
ATDE_D cpx as_cpx(f32x2 v) { cpx a; a.r = v.x; a.i = v.y; return a; }
template <bool INVERSE>
ATDE_D void kf_bfly4_packed(cpx& f0, cpx& f1, cpx& f2, cpx& f3, tw4 t1, tw4 t2, tw4 t3, f32x2 one, f32x2 mone)
{
    const f32x2 s0 = as2(cmul_tw(f1, t1));
    const f32x2 s1 = as2(cmul_tw(f2, t2));
    const f32x2 s2 = as2(cmul_tw(f3, t3));
    const f32x2 s5 = sub2(as2(f0), s1, mone);
    f32x2 g0 = add2(as2(f0), s1, one);
    const f32x2 s3 = add2(s0, s2, one);
    const f32x2 s4 = sub2(s0, s2, mone);
    f2 = as_cpx(sub2(g0, s3, mone));
    f0 = as_cpx(add2(g0, s3, one));
    if (INVERSE) {
        f1.r = fsub(s5.x, s4.y);  f1.i = fadd(s5.y, s4.x);
        f3.r = fadd(s5.x, s4.y);  f3.i = fsub(s5.y, s4.x);
    } else {
        f1.r = fadd(s5.x, s4.y);  f1.i = fsub(s5.y, s4.x);
        f3.r = fsub(s5.x, s4.y);  f3.i = fadd(s5.y, s4.x);
    }
}

// One radix-4 butterfly of a stage, operating in place on `buf` (shared memory).
//   v      : butterfly number within this FFT instance, 0 <= v < n/4
//   m      : sub-length of the stage;  fstride = n/(4*m)
template <bool INVERSE>
ATDE_D void kf_stage4(cpx* buf, const cpx* tw, int v, int m, int fstride)
{
    const int g = v / m, k = v - g * m;
    cpx* F = buf + g * 4 * m + k;
    cpx f0 = F[0], f1 = F[m], f2 = F[2 * m], f3 = F[3 * m];
    kf_bfly4<INVERSE>(f0, f1, f2, f3, tw[k * fstride], tw[2 * k * fstride], tw[3 * k * fstride]);
    F[0] = f0; F[m] = f1; F[2 * m] = f2; F[3 * m] = f3;
}

// One radix-2 butterfly of a stage (only ever the innermost stage, m = 1, for n = 2*4^k).
ATDE_D void kf_stage2(cpx* buf, const cpx* tw, int v, int m, int fstride)
{
    const int g = v / m, k = v - g * m;
    cpx* F = buf + g * 2 * m + k;
    cpx a = F[0], b = F[m];
    kf_bfly2(a, b, tw[k * fstride]);
    F[0] = a; F[m] = b;
}

} // namespace atde
