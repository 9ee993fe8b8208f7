/*
 * oracle/oracle_common.c — TEST INFRASTRUCTURE ONLY.  See oracle_common.h.
 * CPU restatement of the shared primitives of the reference's encode path.
 */
#include "oracle_common.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ kissfft ---------------- */
/* kf_factor, kiss_fft.c:308-330: powers of 4 first, then 2, then odd primes. */
static void okf_factor(int n, int* facbuf)
{
    int p = 4;
    double floor_sqrt = floor(sqrt((double)n));
    do {
        while (n % p) {
            switch (p) {
                case 4: p = 2; break;
                case 2: p = 3; break;
                default: p += 2; break;
            }
            if (p > floor_sqrt)
                p = n;
        }
        n /= p;
        *facbuf++ = p;
        *facbuf++ = n;
    } while (n > 1);
}

okiss* okiss_alloc(int nfft, int inverse)
{
    okiss* st = (okiss*)calloc(1, sizeof(okiss));
    st->nfft = nfft;
    st->inverse = inverse;
    st->tw = (ocpx*)malloc(sizeof(ocpx) * nfft);
    for (int i = 0; i < nfft; ++i) {
        /* kiss_fft.c:357-363 : double phase, double cos/sin, cast to float */
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / nfft;
        if (inverse)
            phase *= -1;
        st->tw[i].r = (float)cos(phase);
        st->tw[i].i = (float)sin(phase);
    }
    okf_factor(nfft, st->factors);
    return st;
}

void okiss_free(okiss* st)
{
    if (!st) return;
    free(st->tw);
    free(st);
}

/* C_MUL, _kiss_fft_guts.h:87-89 : four products, one sub, one add, each rounded */
static inline ocpx cmul(ocpx a, ocpx b)
{
    ocpx m;
    m.r = a.r * b.r - a.i * b.i;
    m.i = a.r * b.i + a.i * b.r;
    return m;
}

/* one radix-2 butterfly of kf_bfly2 (kiss_fft.c:21-42) */
static inline void bfly2(ocpx* a, ocpx* b, ocpx tw)
{
    ocpx t = cmul(*b, tw);
    b->r = a->r - t.r;  b->i = a->i - t.i;
    a->r += t.r;        a->i += t.i;
}

/* one radix-4 butterfly of kf_bfly4 (kiss_fft.c:44-90), same temporaries, same order */
static inline void bfly4(ocpx* f0, ocpx* f1, ocpx* f2, ocpx* f3, ocpx t1, ocpx t2, ocpx t3, int inverse)
{
    ocpx s0 = cmul(*f1, t1);
    ocpx s1 = cmul(*f2, t2);
    ocpx s2 = cmul(*f3, t3);
    ocpx s5, s3, s4;
    s5.r = f0->r - s1.r;  s5.i = f0->i - s1.i;
    f0->r += s1.r;        f0->i += s1.i;
    s3.r = s0.r + s2.r;   s3.i = s0.i + s2.i;
    s4.r = s0.r - s2.r;   s4.i = s0.i - s2.i;
    f2->r = f0->r - s3.r; f2->i = f0->i - s3.i;
    f0->r += s3.r;        f0->i += s3.i;
    if (inverse) {
        f1->r = s5.r - s4.i;  f1->i = s5.i + s4.r;
        f3->r = s5.r + s4.i;  f3->i = s5.i - s4.r;
    } else {
        f1->r = s5.r + s4.i;  f1->i = s5.i - s4.r;
        f3->r = s5.r - s4.i;  f3->i = s5.i + s4.r;
    }
}

/*
 * kf_work (kiss_fft.c:237-302) is a decimation-in-time recursion.  Unrolled: output slot
 *   o = q1*m1 + q2*m2 + ... + qk   receives input   i = q1 + q2*p1 + q3*p1*p2 + ...
 * (mixed-radix digit reversal), then the stages run innermost (last factor) first; the stage
 * with radix p and sub-length m uses twiddle stride fstride = nfft/(p*m).  Only radix 2 and 4
 * occur for the sizes on this path (8,16,64,128,256,2048).
 */
void okiss_fft(const okiss* st, const ocpx* in, ocpx* out)
{
    int n = st->nfft;
    int p[32], m[32], ns = 0;
    for (const int* f = st->factors;; f += 2) {
        p[ns] = f[0]; m[ns] = f[1]; ns++;
        if (f[1] == 1) break;
    }
    for (int o = 0; o < n; o++) {
        int rem = o, idx = 0, stride = 1;
        for (int s = 0; s < ns; s++) {
            int q = rem / m[s];
            rem -= q * m[s];
            idx += q * stride;
            stride *= p[s];
        }
        out[o] = in[idx];
    }
    for (int s = ns - 1; s >= 0; s--) {
        int P = p[s], M = m[s];
        int fstride = n / (P * M);
        for (int base = 0; base < n; base += P * M) {
            for (int k = 0; k < M; k++) {
                ocpx* F = out + base + k;
                if (P == 2) {
                    bfly2(F, F + M, st->tw[k * fstride]);
                } else if (P == 4) {
                    bfly4(F, F + M, F + 2 * M, F + 3 * M, st->tw[k * fstride], st->tw[2 * k * fstride],
                          st->tw[3 * k * fstride], st->inverse);
                } else {
                    abort();
                }
            }
        }
    }
}

okissr* okissr_alloc(int nfft, int inverse)
{
    okissr* st = (okissr*)calloc(1, sizeof(okissr));
    nfft >>= 1;
    st->ncfft = nfft;
    st->sub = okiss_alloc(nfft, inverse);
    st->tmp = (ocpx*)malloc(sizeof(ocpx) * nfft);
    st->super_tw = (ocpx*)malloc(sizeof(ocpx) * (nfft / 2));
    for (int i = 0; i < nfft / 2; ++i) {
        /* kiss_fftr.c:50-56 */
        double phase = -3.14159265358979323846264338327 * ((double)(i + 1) / nfft + .5);
        if (inverse)
            phase *= -1;
        st->super_tw[i].r = (float)cos(phase);
        st->super_tw[i].i = (float)sin(phase);
    }
    return st;
}

void okissr_free(okissr* st)
{
    if (!st) return;
    okiss_free(st->sub);
    free(st->tmp);
    free(st->super_tw);
    free(st);
}

void okiss_fftr(const okissr* st, const float* timedata, ocpx* freq)
{
    int ncfft = st->ncfft;
    okiss_fft(st->sub, (const ocpx*)timedata, st->tmp);
    ocpx tdc = st->tmp[0];
    freq[0].r = tdc.r + tdc.i;
    freq[ncfft].r = tdc.r - tdc.i;
    freq[ncfft].i = freq[0].i = 0;
    for (int k = 1; k <= ncfft / 2; ++k) {
        ocpx fpk = st->tmp[k];
        ocpx fpnk = { st->tmp[ncfft - k].r, -st->tmp[ncfft - k].i };
        ocpx f1k = { fpk.r + fpnk.r, fpk.i + fpnk.i };
        ocpx f2k = { fpk.r - fpnk.r, fpk.i - fpnk.i };
        ocpx tw = cmul(f2k, st->super_tw[k - 1]);
        /* HALF_OF(x) = x*.5 : float sum promoted, halved, rounded back == exact halving */
        freq[k].r = (float)((f1k.r + tw.r) * .5);
        freq[k].i = (float)((f1k.i + tw.i) * .5);
        freq[ncfft - k].r = (float)((f1k.r - tw.r) * .5);
        freq[ncfft - k].i = (float)((tw.i - f1k.i) * .5);
    }
}

void okiss_fftri(const okissr* st, const ocpx* freq, float* timedata)
{
    int ncfft = st->ncfft;
    st->tmp[0].r = freq[0].r + freq[ncfft].r;
    st->tmp[0].i = freq[0].r - freq[ncfft].r;
    for (int k = 1; k <= ncfft / 2; ++k) {
        ocpx fk = freq[k];
        ocpx fnkc = { freq[ncfft - k].r, -freq[ncfft - k].i };
        ocpx fek = { fk.r + fnkc.r, fk.i + fnkc.i };
        ocpx tmp = { fk.r - fnkc.r, fk.i - fnkc.i };
        ocpx fok = cmul(tmp, st->super_tw[k - 1]);
        st->tmp[k].r = fek.r + fok.r;
        st->tmp[k].i = fek.i + fok.i;
        st->tmp[ncfft - k].r = fek.r - fok.r;
        st->tmp[ncfft - k].i = fek.i - fok.i;
        st->tmp[ncfft - k].i *= -1;
    }
    okiss_fft(st->sub, st->tmp, (ocpx*)timedata);
}

/* ------------------------------------------------------------------ MDCT ------------------- */
omdct* omdct_alloc(int n, float scale)
{
    omdct* m = (omdct*)calloc(1, sizeof(omdct));
    m->n = n;
    m->sincos = (float*)malloc(sizeof(float) * (n >> 1));
    /* CalcSinCos, mdct.cpp:25-36 : alpha, omiga, scale are FLOAT variables */
    const float alpha = 2.0 * M_PI / (8.0 * n);
    const float omiga = 2.0 * M_PI / n;
    scale = sqrtf(scale / n);
    for (int i = 0; i < (n >> 2); ++i) {
        m->sincos[2 * i + 0] = scale * cosf(omiga * i + alpha);
        m->sincos[2 * i + 1] = scale * sinf(omiga * i + alpha);
    }
    m->fft = okiss_alloc(n >> 2, 0);
    m->fin = (ocpx*)malloc(sizeof(ocpx) * (n >> 2));
    m->fout = (ocpx*)malloc(sizeof(ocpx) * (n >> 2));
    return m;
}

void omdct_free(omdct* m)
{
    if (!m) return;
    okiss_free(m->fft);
    free(m->sincos); free(m->fin); free(m->fout); free(m);
}

/* TMDCT::operator(), mdct.h:51-104 */
void omdct_run(const omdct* m, const float* in, float* out)
{
    const int N = m->n, n2 = N >> 1, n4 = N >> 2, n34 = 3 * n4, n54 = 5 * n4;
    const float* cs = m->sincos;
    int n;
    for (n = 0; n < n4; n += 2) {
        float r0 = in[n34 - 1 - n] + in[n34 + n];
        float i0 = in[n4 + n] - in[n4 - 1 - n];
        float c = cs[n], s = cs[n + 1];
        m->fin[n / 2].r = r0 * c + i0 * s;
        m->fin[n / 2].i = i0 * c - r0 * s;
    }
    for (; n < n2; n += 2) {
        float r0 = in[n34 - 1 - n] - in[n - n4];
        float i0 = in[n4 + n] + in[n54 - 1 - n];
        float c = cs[n], s = cs[n + 1];
        m->fin[n / 2].r = r0 * c + i0 * s;
        m->fin[n / 2].i = i0 * c - r0 * s;
    }
    okiss_fft(m->fft, m->fin, m->fout);
    for (n = 0; n < n2; n += 2) {
        float r0 = m->fout[n / 2].r, i0 = m->fout[n / 2].i;
        float c = cs[n], s = cs[n + 1];
        out[n] = -r0 * c - i0 * s;
        out[n2 - 1 - n] = -r0 * s + i0 * c;
    }
}

/* ------------------------------------------------------------------ QMF -------------------- */
void oqmf_window(float w[48])
{
    /* qmf.cpp:25-45 */
    static const float TapHalf[24] = {
        -0.00001461907,  -0.00009205479, -0.000056157569,  0.00030117269,
        0.0002422519,    -0.00085293897, -0.0005205574,    0.0020340169,
        0.00078333891,   -0.0042153862,  -0.00075614988,   0.0078402944,
        -0.000061169922, -0.01344162,    0.0024626821,     0.021736089,
        -0.007801671,    -0.034090221,   0.01880949,       0.054326009,
        -0.043596379,    -0.099384367,   0.13207909,       0.46424159
    };
    for (int i = 0; i < 24; i++)
        w[i] = w[47 - i] = TapHalf[i] * 2.0;
}

void oqmf_analysis(const float w[48], float* hist46, const float* in, int n_in, float* lower, float* upper)
{
    /* qmf.h:47-64 with PcmBuffer = [hist46 | in] */
    float* buf = (float*)malloc(sizeof(float) * (n_in + 46));
    memcpy(buf, hist46, 46 * sizeof(float));
    memcpy(buf + 46, in, n_in * sizeof(float));
    for (int j = 0; j < n_in; j += 2) {
        float lo = 0.0, up = 0.0;
        for (int i = 0; i < 24; i++) {
            lo += w[2 * i] * buf[48 - 1 + j - (2 * i)];
            up += w[(2 * i) + 1] * buf[48 - 1 + j - (2 * i) - 1];
        }
        upper[j / 2] = lo - up;
        lower[j / 2] = lo + up;
    }
    memcpy(hist46, buf + n_in, 46 * sizeof(float));
    free(buf);
}

/* ------------------------------------------------------------------ bit writer ------------- */
void obits_init(obits* b) { memset(b, 0, sizeof(*b)); }

void obits_write(obits* b, uint32_t val, int n)
{
    /* bitstream.cpp:40-63, including its buffer growth rule (GetBytes() length depends on it) */
    if (n > 23 || n < 0) abort();
    const int bitsLeft = b->size * 8 - b->bits_used;
    const int bitsReq = n - bitsLeft;
    const int bytesPos = b->bits_used / 8;
    const int overlap = b->bits_used % 8;
    if (overlap || bitsReq >= 0)
        b->size += bitsReq / 8 + (overlap ? 2 : 1);
    uint32_t t = (val << (32 - n) >> overlap);
    for (int i = 0; i < n / 8 + (overlap ? 2 : 1); ++i)
        b->buf[bytesPos + i] |= (uint8_t)(t >> (8 * (3 - i)));
    b->bits_used += n;
}

int omake_sign(int val, unsigned bits)
{
    unsigned shift = 8 * sizeof(int) - bits;
    union { unsigned u; int s; } v = { (unsigned)val << shift };
    return v.s >> shift;
}

/* ------------------------------------------------------------------ bisection -------------- */
void obisect_start(obisect* b, size_t target, float mn, float mx)
{
    b->target = target; b->min_l = mn; b->max_l = mx; b->last = mx; b->need_repeat = 0;
}

float obisect_continue(obisect* b)
{
    if (b->max_l <= b->min_l)
        return b->last;
    b->cur = (b->max_l + b->min_l) / 2.0;
    return b->cur;
}

int obisect_submit(obisect* b, size_t got)
{
    if (b->max_l <= b->min_l) {
        b->need_repeat = 0;
    } else {
        if (got < b->target) {
            b->last = b->cur;
            b->max_l = b->cur - 0.01f;
            b->need_repeat = 1;
        } else if (got > b->target) {
            b->min_l = b->cur + 0.01f;
            b->need_repeat = 1;
        } else {
            b->need_repeat = 0;
        }
    }
    return !b->need_repeat;
}

/* ------------------------------------------------------------------ psy tables ------------- */
/* ATHformula_Frank, atrac_psy_common.cpp:33-95 (table borrowed from Musepack by the reference) */
static float oath_formula(float freq)
{
    static const short tab[] = {
        9669, 9669, 9626, 9512, 9353, 9113, 8882, 8676, 8469, 8243, 7997, 7748, 7492, 7239, 7000, 6762,
        6529, 6302, 6084, 5900, 5717, 5534, 5351, 5167, 5004, 4812, 4638, 4466, 4310, 4173, 4050, 3922,
        3723, 3577, 3451, 3281, 3132, 3036, 2902, 2760, 2658, 2591, 2441, 2301, 2212, 2125, 2018, 1900,
        1770, 1682, 1594, 1512, 1430, 1341, 1260, 1198, 1136, 1057,  998,  943,  887,  846,  744,  712,
         693,  668,  637,  606,  580,  555,  529,  502,  475,  448,  422,  398,  375,  351,  327,  322,
         312,  301,  291,  268,  246,  215,  182,  146,  107,   61,   13,  -35,  -96, -156, -179, -235,
        -295, -350, -401, -421, -446, -499, -532, -535, -513, -476, -431, -313, -179,    8,  203,  403,
         580,  736,  881, 1022, 1154, 1251, 1348, 1421, 1479, 1399, 1285, 1193, 1287, 1519, 1914, 2369,
        3352, 4352, 5352, 6352, 7352, 8352, 9352, 9999, 9999, 9999, 9999, 9999,
    };
    double freq_log;
    unsigned index;
    if (freq < 10.) freq = 10.;
    if (freq > 29853.) freq = 29853.;
    freq_log = 40. * log10(0.1 * freq);
    index = (unsigned)freq_log;
    return 0.01 * (tab[index] * (1 + index - freq_log) + tab[index + 1] * (freq_log - index));
}

void ocalc_ath(int len, int sample_rate, float* res)
{
    float mf = (float)sample_rate / 2000.0;
    for (int i = 0; i < len; i++) {
        const float f = (float)(i + 1) * mf / len;
        float trh = oath_formula(1.e3 * f) - 100;
        trh -= f * f * 0.015;
        res[i] = trh;
    }
}

void ocreate_loudness_curve(int sz, float* res)
{
    for (int i = 0; i < sz; i++) {
        float f = (float)(i + 3) * 0.5 * 44100 / (float)sz;
        float t = log10f(f) - 3.5;            /* std::log10(float) -> log10f */
        t = -10 * t * t + 3 - f / 3000;
        t = pow(10, (0.1 * t));
        res[i] = t;
    }
}

float otrack_loudness2(float prev, float l0, float l1) { return 0.98 * prev + 0.01 * (l0 + l1); }
float otrack_loudness1(float prev, float l) { return 0.98 * prev + 0.02 * l; }

/* ------------------------------------------------------------------ self-test hooks -------- */
/* Restates the fixture of src/lib/bs_encode/encode_ut.cpp:27-33,138-176: one part calls
 * Start(1000,-15,-1), the next runs Continue/Submit on a synthetic cost function until done.
 * kind 1 = SomeBitFn1, kind 2 = SomeBitFn2.  Reports Encode() calls and the final bit count. */
void obisect_selftest(int kind, int* calls, long* bits)
{
    obisect b;
    obisect_start(&b, 1000, -15, -1);
    int n = 0;
    size_t got = 0;
    for (;;) {
        float lambda = obisect_continue(&b);
        size_t f1 = sqrtf(lambda * (-1.0f)) * 300;
        got = (kind == 1) ? f1 : 1 + (f1 & (~(size_t)7));
        n++;
        if (obisect_submit(&b, got))
            break;
    }
    *calls = n;
    *bits = (long)got;
}
