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
