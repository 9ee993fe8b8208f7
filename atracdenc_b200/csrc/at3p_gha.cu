// at3p_gha.cu — ATRAC3plus tone search on sm_100a (first, correctness-first version).
//
// Replaces TSubbandGhaProcessor::AnalyzeChannels / DoRound (src/atrac/at3p/at3p_gha.cpp:732-953),
// CheckResuidalAndApply (:492-579), CheckNextFrame (:780-813), PsyPreCheck (:955-973), FillResultBuf /
// FillFolowerRes / AdjustEnvelope (:1499-1664) and libgha's gha_analyze_one / gha_adjust_info_newton_md /
// sle_solve (src/lib/libgha/src/gha.c:141-470, sle.c:9-60).
//
// Parallel structure.  A frame's search is a sequence of rounds; in a round every (channel, subband)
// re-fits its tones (Newton) and extracts one more.  The reference walks them in order, but a
// (channel, subband) step only reads that subband's own tones and samples; the order matters only for
// the shared tone budget (48) and for the "too close to a neighbour" test, which look at the other
// subbands' state.  So: one THREAD per (channel, subband) computes its step from the state at the start
// of the round into a staging area (kGhaTask = 16 threads per frame, two frames per warp); then one
// lane per frame commits the 16 staged steps in the reference's order, applying the budget and the
// neighbour tests against the partially committed state, exactly as the sequential code would see it.
// All arithmetic is the reference's: un-fused doubles / floats in its operation order, glibc's
// sin / cos / atan / sincosf restated (glibc_trig.cuh), kissfft restated (kissfft_dev.cuh).
// Across frames the only carried state is the previous valid result's stop envelopes
// (ResultBufHistory), scanned per stream by at3p_gha_result_kernel.
#include "at3p_kernels.cuh"
#include "kissfft_dev.cuh"
#include "glibc_trig.cuh"
#include "host_tables.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace atde {
namespace at3p {

constexpr int kGhaSb = 8;                 // SUBBANDS, at3p_gha.cpp:216
constexpr int kGhaTask = 16;              // (channel, subband) tasks per frame
constexpr int kMaxDim = 16;               // tones per subband: SubbandDone counts to 15, +1 being fitted
constexpr unsigned kEmpty = 0xffffffffu;  // TAt3PGhaData::EMPTY_POINT
constexpr unsigned kInit = 0xfffffffeu;   // TAt3PGhaData::INIT
constexpr double kPi = 3.14159265358979323846;

struct GhaTables {
    float window[128];                    // gha_init_window, gha.c:41-56
    cpx tw64[64];                         // forward, kiss_fftr(128) -> FFT-64
    cpx super128[32];
    cpx tw128i[128];                      // inverse, kiss_fftri(256) -> FFT-128
    cpx super256i[64];
    float subband_ath[kGhaSb];            // FillSubbandAth, at3p_gha.cpp:453-465
    float amp_sf_tab[64];                 // CreateAmpSfTab, :467-474
    float sine_tab[2048];                 // :281-283
};

struct GhaInfo { float frequency, phase, magnitude; };

// TChannelData of one (channel, subband), at3p_gha.cpp:225-247
struct SbState {
    int n;                                // tones of this subband in GhaInfos, ascending key
    unsigned key[kMaxDim];
    GhaInfo info[kMaxDim];
    unsigned env_first, env_second;
    int gapless, done;                    // done == 16: MarkSubbandDone
    float max_mag, last_res_energy;
    unsigned last_added;
};

// what one step wants to do, applied (or dropped) by the commit pass
struct Staged {
    int part1;                            // 0 no tones, 1 fit ok (replace tones), 2 failed: erase last added, mark done
    int n_new;
    GhaInfo fit[kMaxDim];                 // re-fitted tones, sorted by frequency
    unsigned env_first, env_second;       // CheckResuidalAndApply's writes (also on its error paths)
    float last_res_energy;
    int gapless;
    int resid_valid;                      // the staged residual replaces Buf
    int analyzed;
    GhaInfo found;                        // gha_analyze_one on the (new) residual
    int psy_ok;
    float max_mag;
};

// per-frame raw search result, consumed by the per-stream result pass
struct GhaFrameOut {
    int total_tones;
    int n[2][kGhaSb];
    unsigned key[2][kGhaSb][kMaxDim];
    GhaInfo info[2][kGhaSb][kMaxDim];
    unsigned env[2][kGhaSb][2];
};

struct TaskScratch {                      // LOCAL memory (the kernel's stack frame): transient within a step; the hardware
                                          // interleaves local memory across the lanes of a warp, so lanes walking their own
                                          // arrays in step touch one line per access instead of 32
    float tmp[128];                       // libgha's ctx->tmp_buf (the Repeat call of gha_adjust_info reads its stale tail)
    float s[kMaxDim][128], c[kMaxDim][128];
    cpx fa[64];
    cpx fout[65];
    cpx fb[128];
};

// per CUDA device, like at3p_kernels.cu:device_tables()
constexpr int kMaxGhaDevices = 64;
static GhaTables* g_gha_tables[kMaxGhaDevices] = {};
static std::mutex g_gha_mu;

static GhaTables* build_gha_tables()
{
    GhaTables* h = new GhaTables();
    memset(h, 0, sizeof(*h));
    {
        const size_t size = 128, n = size + 1, half = size / 2;
        for (size_t i = 0; i < half; i++) {
            h->window[i] = sinf(M_PI * (i + 1) / n);
            h->window[i] *= h->window[i];
        }
        for (size_t i = half; i < size; i++) h->window[i] = h->window[size - 1 - i];
    }
    const auto t64 = kiss_twiddles(64, false);
    memcpy(h->tw64, t64.data(), sizeof(h->tw64));
    const auto s128 = kiss_super_twiddles(128, false);
    memcpy(h->super128, s128.data(), sizeof(h->super128));
    const auto t128 = kiss_twiddles(128, true);
    memcpy(h->tw128i, t128.data(), sizeof(h->tw128i));
    const auto s256 = kiss_super_twiddles(256, true);
    memcpy(h->super256i, s256.data(), sizeof(h->super256i));
    {
        const auto ath = calc_ath(16 * 1024, 44100);
        for (size_t sb = 0; sb < (size_t)kGhaSb; sb++) {
            float m = 999.;
            for (size_t f = sb * 1024, i = 0; i < 1024; f++, i++) m = fmin(m, ath[f]);
            h->subband_ath[sb] = pow(10, 0.1 * (m + 90));
        }
    }
    for (int i = 0; i < 64; i++) h->amp_sf_tab[i] = exp2f((i - 3) / 4.0f);
    for (int i = 0; i < 2048; i++) h->sine_tab[i] = sin(2 * M_PI * i / 2048);
    GhaTables* d = nullptr;
    if (cudaMalloc(&d, sizeof(GhaTables)) == cudaSuccess &&
        cudaMemcpy(d, h, sizeof(GhaTables), cudaMemcpyHostToDevice) == cudaSuccess) {
        delete h;
        return d;
    }
    if (d) cudaFree(d);
    delete h;
    return nullptr;
}

const GhaTables* gha_tables()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxGhaDevices) return nullptr;
    std::lock_guard<std::mutex> lock(g_gha_mu);
    if (!g_gha_tables[dev]) g_gha_tables[dev] = build_gha_tables();
    return g_gha_tables[dev];
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
ATDE_D double dmul(double a, double b) { return __dmul_rn(a, b); }
ATDE_D double dadd(double a, double b) { return __dadd_rn(a, b); }
ATDE_D double dsub(double a, double b) { return __dsub_rn(a, b); }
ATDE_D double ddiv(double a, double b) { return __ddiv_rn(a, b); }
ATDE_D float d2f(double a) { return __double2float_rn(a); }

// GhaFreqToIndex (at3p_gha.cpp:50-53): lrintf(1024.0f * (f / M_PI)) & 1023 | sb << 10
ATDE_D unsigned freq_to_index(float f, unsigned sb)
{
    const float v = d2f(dmul(1024.0, ddiv((double)f, kPi)));
    return ((unsigned)__float2int_rn(v) & 1023u) | (sb << 10);
}
// GhaPhaseToIndex (:55-58): lrintf(32.0 * (p / (2 * M_PI))) & 31
ATDE_D unsigned phase_to_index(float p)
{
    const float v = d2f(dmul(32.0, ddiv((double)p, 2 * kPi)));
    return (unsigned)__float2int_rn(v) & 31u;
}
// AmplitudeToSf (:1666-1673): upper_bound - 1, clamped at 0
ATDE_D unsigned amplitude_to_sf(const GhaTables* G, float amp)
{
    int lo = 0, hi = 64;                                 // first element > amp
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (amp < G->amp_sf_tab[mid]) hi = mid; else lo = mid + 1;
    }
    return (unsigned)(lo > 0 ? lo - 1 : 0);
}

// in-place kissfft on a buffer already in gather order; radix-4 stages with sub-lengths m0, 4 m0, ...
template <bool INVERSE>
ATDE_D void fft_stages4(cpx* buf, const cpx* tw, int n, int m0)
{
    for (int m = m0; m < n; m *= 4) {
        const int fstride = n / (4 * m);
        for (int v = 0; v < n / 4; v++) kf_stage4<INVERSE>(buf, tw, v, m, fstride);
    }
}

// ---------------------------------------------------------------------------------------------
// gha_analyze_one (gha.c:404-431) with upsample = 1
// ---------------------------------------------------------------------------------------------
ATDE_D GhaInfo analyze_one(const GhaTables* G, const float* pcm, TaskScratch* ws)
{
    float* tmp = ws->tmp;
    for (int i = 0; i < 128; i++) tmp[i] = fmul(pcm[i], G->window[i]);
    // kiss_fftr(128): FFT-64 of the packed pairs (64 = 4 x 4 x 4: slot 16 d0 + 4 d1 + d2 holds input d0 + 4 d1 + 16 d2)
    cpx* fa = ws->fa;
    for (int slot = 0; slot < 64; slot++) {
        const int j = (slot >> 4) + 4 * ((slot >> 2) & 3) + 16 * (slot & 3);
        fa[slot].r = tmp[2 * j];
        fa[slot].i = tmp[2 * j + 1];
    }
    fft_stages4<false>(fa, G->tw64, 64, 1);
    cpx* fo = ws->fout;                                   // freqdata[0..64]; [65..128] stay zero (calloc)
    {
        const float tr = fa[0].r, ti = fa[0].i;
        fo[0].r = fadd(tr, ti);  fo[0].i = 0.0f;
        fo[64].r = fsub(tr, ti); fo[64].i = 0.0f;
        for (int k = 1; k <= 32; k++) {
            const cpx fpk = fa[k];
            cpx fpnk; fpnk.r = fa[64 - k].r; fpnk.i = -fa[64 - k].i;
            cpx f1k, f2k;
            f1k.r = fadd(fpk.r, fpnk.r); f1k.i = fadd(fpk.i, fpnk.i);
            f2k.r = fsub(fpk.r, fpnk.r); f2k.i = fsub(fpk.i, fpnk.i);
            const cpx t2 = cmul(f2k, G->super128[k - 1]);
            fo[k].r = fmul(fadd(f1k.r, t2.r), 0.5f);
            fo[k].i = fmul(fadd(f1k.i, t2.i), 0.5f);
            fo[64 - k].r = fmul(fsub(f1k.r, t2.r), 0.5f);
            fo[64 - k].i = fmul(fsub(t2.i, f1k.i), 0.5f);
        }
    }
    // gha_estimate_bin (:141-157)
    int bin = 0;
    {
        float mx = 0.0f;
        for (int i = 0; i < 65; i++) {
            const float t = fadd(fmul(fo[i].r, fo[i].r), fmul(fo[i].i, fo[i].i));
            if (t > mx) { mx = t; bin = i; }
        }
    }
    // resample_fft (:159-166): kiss_fftri(256) of the zero-extended spectrum, then / 128
    cpx* fb = ws->fb;                                     // gather order of 128 = 4 x 4 x 4 x 2
    {
        auto freq = [&](int k) { cpx z; if (k <= 64) z = fo[k]; else { z.r = 0.0f; z.i = 0.0f; } return z; };
        auto put = [&](int idx, cpx v) {
            // input index d0 + 4 d1 + 16 d2 + 64 d3 sits at slot 32 d0 + 8 d1 + 2 d2 + d3
            const int slot = ((idx & 3) << 5) | (((idx >> 2) & 3) << 3) | (((idx >> 4) & 3) << 1) | (idx >> 6);
            fb[slot] = v;
        };
        cpx t0;
        t0.r = fadd(freq(0).r, freq(128).r);
        t0.i = fsub(freq(0).r, freq(128).r);
        put(0, t0);
        for (int k = 1; k <= 64; k++) {
            const cpx fk = freq(k);
            cpx fnkc; fnkc.r = freq(128 - k).r; fnkc.i = -freq(128 - k).i;
            cpx fek, tp;
            fek.r = fadd(fk.r, fnkc.r); fek.i = fadd(fk.i, fnkc.i);
            tp.r = fsub(fk.r, fnkc.r);  tp.i = fsub(fk.i, fnkc.i);
            const cpx fok = cmul(tp, G->super256i[k - 1]);
            cpx a, bq;
            a.r = fadd(fek.r, fok.r); a.i = fadd(fek.i, fok.i);
            bq.r = fsub(fek.r, fok.r); bq.i = fmul(fsub(fek.i, fok.i), -1.0f);
            put(k, a);
            put(128 - k, bq);                             // k == 64 overwrites, as in the reference
        }
        for (int v = 0; v < 64; v++) kf_stage2(fb, G->tw128i, v, 1, 64);
        fft_stages4<true>(fb, G->tw128i, 128, 2);
    }
    float* res = reinterpret_cast<float*>(fb);            // 256 reals
    for (int i = 0; i < 256; i++) res[i] = __fdiv_rn(res[i], 128.0f);
    // gha_search_omega_newton (:173-236) on the 256 resampled points
    GhaInfo out;
    {
        double omega = ddiv(dmul((double)(bin * 2), kPi), 256.0);
        for (int loop = 0; loop <= 7; loop++) {
            double Xr = 0, Xi = 0, dXr = 0, dXi = 0, ddXr = 0, ddXs = 0;
            const double a = g_cos(omega), b = g_sin(omega);
            double c = 1.0, s = 0.0, dn = 0.0;
            for (int n = 0; n < 256; n++, dn += 1.0) {      // dn == (double)n exactly
                const double p = (double)res[n];
                const double cm = dmul(p, c), sm = dmul(p, s);
                Xr = dadd(Xr, cm);
                Xi = dadd(Xi, sm);
                const double tc = dmul(dn, cm), ts = dmul(dn, sm);
                dXr = dsub(dXr, ts);
                dXi = dadd(dXi, tc);
                ddXr = dsub(ddXr, dmul(dn, tc));
                ddXs = dsub(ddXs, dmul(dn, ts));
                const double nc = dsub(dmul(a, c), dmul(b, s));
                const double ns = dadd(dmul(b, c), dmul(a, s));
                c = nc; s = ns;
            }
            const double F = dadd(dmul(Xr, dXr), dmul(Xi, dXi));
            const double G2 = dadd(dmul(Xr, Xr), dmul(Xi, Xi));
            const double dF = dadd(dadd(dadd(dmul(Xr, ddXr), dmul(dXr, dXr)), dmul(Xi, ddXs)), dmul(dXi, dXi));
            const double dw = ddiv(F, dsub(dF, ddiv(dmul(F, F), G2)));
            omega = dsub(omega, dw);
            if (omega < 0) omega = dmul(omega, -1.0);
            while (omega > kPi * 2.0) omega = dsub(omega, kPi * 2.0);
            if (omega > kPi) omega = dsub(kPi * 2.0, omega);
            if (loop == 7) {
                out.frequency = d2f(omega);
                out.phase = d2f(dsub(kPi / 2, g_atan(ddiv(Xi, Xr))));
                if (Xr < 0) out.phase = d2f(dadd((double)out.phase, kPi));
            }
        }
        out.frequency = d2f(dmul((double)out.frequency, 2.0));
    }
    // gha_generate_sine (:238-244) + gha_estimate_magnitude (:246-257)
    double t1 = 0, t2 = 0;
    float fi = 0.0f;
    for (int i = 0; i < 128; i++, fi += 1.0f) {
        const float arg = fadd(fmul(out.frequency, fi), out.phase);
        const float r = d2f(g_sin((double)arg));
        tmp[i] = r;
        t1 = dadd(t1, (double)fmul(pcm[i], r));
        t2 = dadd(t2, (double)fmul(r, r));
    }
    out.magnitude = d2f(ddiv(t1, t2));
    return out;
}

// ---------------------------------------------------------------------------------------------
// sle_solve (sle.c:9-60) on one dim x (dim+1) block.  The reference solves the 3dim system at once, but
// gha_adjust_info_newton_md overwrites the coupling blocks with never-written (zero) entries
// (gha.c:341-343), which leaves three independent blocks; rows of the other blocks carry exact zeros
// in this block's columns, never win a pivot search and are skipped by the |t| < eps test.
// ---------------------------------------------------------------------------------------------
ATDE_D int sle_block(double* a, int n, double* x)
{
    const int col = n + 1;
    const double eps = (double)0.00001f;
    for (int k = 0; k < n; k++) {
        double mx = fabs(a[col * k + k]);
        int index = k;
        for (int i = k + 1; i < n; i++) {
            const double t = fabs(a[col * i + k]);
            if (t > mx) { mx = t; index = i; }
        }
        if (mx < eps) return -1;
        if (index != k)
            for (int i = 0; i < col; i++) { const double t = a[col * k + i]; a[col * k + i] = a[col * index + i]; a[col * index + i] = t; }
        for (int i = k; i < n; i++) {
            const double t = a[col * i + k];
            if (fabs(t) < eps) continue;
            for (int j = 0; j < col; j++) a[i * col + j] = ddiv(a[i * col + j], t);
            if (i != k)
                for (int j = 0; j < col; j++) a[i * col + j] = dsub(a[i * col + j], a[k * col + j]);
        }
    }
    for (int k = n - 1; k >= 0; k--) {
        x[k] = a[col * k + n];
        for (int i = 0; i < k; i++) a[col * i + n] = dsub(a[col * i + n], dmul(a[col * i + k], x[k]));
    }
    return 0;
}

// The end of one Newton loop for one tone (gha.c:348-399): damped step, then the sign / range repairs.  The
// reference runs the four repair loops one after the other over all tones; they only touch their own tone, so
// doing all four for tone k before tone k+1 gives the same values.
ATDE_D void newton_update(GhaInfo& t, double da, double dw, double dp)
{
    t.magnitude = d2f(dsub((double)t.magnitude, dmul(da, 0.8)));
    t.frequency = d2f(dsub((double)t.frequency, dmul(dw, 0.8)));
    t.phase = d2f(dsub((double)t.phase, dmul(dp, 0.8)));
    if (t.magnitude < 0) {
        t.magnitude = fmul(t.magnitude, -1.0f);
        t.phase = d2f(dadd((double)t.phase, kPi));
    }
    if (t.magnitude > 32768.0f) t.magnitude = d2f(dmul(32768.0, 0.5));
    if (t.frequency < 0) {
        t.frequency = fmul(t.frequency, -1.0f);
        t.phase = d2f(dsub(2 * kPi, (double)t.phase));
    }
    while ((double)t.frequency > kPi * 2.0) t.frequency = d2f(dsub((double)t.frequency, kPi * 2.0));
    if ((double)t.frequency > kPi) t.frequency = d2f(dsub(2 * kPi, (double)t.frequency));
    while ((double)t.phase > kPi * 2.0) t.phase = d2f(dsub((double)t.phase, kPi * 2));
    while (t.phase < 0) t.phase = d2f(dadd((double)t.phase, kPi * 2));
}

// ---------------------------------------------------------------------------------------------
// gha_adjust_info_newton_md (gha.c:259-402); leaves the last loop's residual in ws->tmp[0..sz)
// ---------------------------------------------------------------------------------------------
ATDE_D int adjust_newton(const float* pcm, GhaInfo* info, int dim, int sz, TaskScratch* ws)
{
    float* tmp = ws->tmp;
    double M[3][kMaxDim * (kMaxDim + 1)];
    double fx[3][kMaxDim];
    for (int loop = 0; loop < 7; loop++) {
        for (int n = 0; n < sz; n++) tmp[n] = pcm[n];
        for (int k = 0; k < dim; k++) {
            const float fr = info[k].frequency, ph = info[k].phase, mg = info[k].magnitude;
            float fn = 0.0f;                                 // == (float)n exactly
            for (int n = 0; n < sz; n++, fn += 1.0f) {
                const float t = fadd(fmul(fr, fn), ph);
                float s, c;
                g_sincosf(t, s, c);
                tmp[n] = fsub(tmp[n], fmul(mg, s));
                ws->s[k][n] = s;
                ws->c[k][n] = c;
            }
        }
        // three blocks: A (ba = -s), w (bw = -A n c), p (bp = -A c); right-hand sides from the residual.
        // One pass over n per tone builds its three diagonal entries and right-hand sides, one pass per tone
        // pair the three off-diagonal entries (each accumulator still adds its terms in the order n = 0, 1, ...).
        {
            const int col = dim + 1;
            for (int i = 0; i < dim; i++) {
                const double Ai = (double)info[i].magnitude;
                double aa = 0.0, ww = 0.0, pp = 0.0, ra = 0.0, rw = 0.0, rp = 0.0, dn = 0.0;
                for (int n = 0; n < sz; n++, dn += 1.0) {
                    const float sf = ws->s[i][n], t = tmp[n];
                    const double c = (double)ws->c[i][n], sd = (double)sf, td = (double)t;
                    aa = dadd(aa, dmul(-sd, -sd));
                    const double bw = dmul(dmul(-Ai, dn), c);
                    const double bww = dmul(dmul(dmul(Ai, dn), dn), sd);
                    ww = dadd(ww, dadd(dmul(td, bww), dmul(bw, bw)));
                    const double bp = dmul(-Ai, c);
                    const float bpp = d2f(dmul(Ai, sd));
                    pp = dadd(pp, dadd((double)fmul(t, bpp), dmul(bp, bp)));
                    ra = dadd(ra, (double)fmul(t, -sf));
                    rw = dadd(rw, (double)fmul(t, d2f(bw)));
                    rp = dadd(rp, (double)fmul(t, d2f(bp)));
                }
                M[0][i * col + i] = dmul(aa, 2.0); M[1][i * col + i] = dmul(ww, 2.0); M[2][i * col + i] = dmul(pp, 2.0);
                M[0][i * col + dim] = dmul(ra, 2.0); M[1][i * col + dim] = dmul(rw, 2.0); M[2][i * col + dim] = dmul(rp, 2.0);
                // M[i][j] and M[j][i] are the same sum of commutative products in the same order: compute j > i, mirror
                for (int j = i + 1; j < dim; j++) {
                    const double Aj = (double)info[j].magnitude;
                    double xa = 0.0, xw = 0.0, xp = 0.0, dn = 0.0;
                    for (int n = 0; n < sz; n++, dn += 1.0) {
                        const double ci = (double)ws->c[i][n], cj = (double)ws->c[j][n];
                        xa = dadd(xa, dmul(-(double)ws->s[i][n], -(double)ws->s[j][n]));
                        xw = dadd(xw, dmul(dmul(dmul(-Ai, dn), ci), dmul(dmul(-Aj, dn), cj)));
                        xp = dadd(xp, dmul(dmul(-Ai, ci), dmul(-Aj, cj)));
                    }
                    M[0][i * col + j] = M[0][j * col + i] = dmul(xa, 2.0);
                    M[1][i * col + j] = M[1][j * col + i] = dmul(xw, 2.0);
                    M[2][i * col + j] = M[2][j * col + i] = dmul(xp, 2.0);
                }
            }
            for (int blk = 0; blk < 3; blk++) {
                for (int i = 0; i < dim; i++) fx[blk][i] = 0.0;
                if (sle_block(M[blk], dim, fx[blk])) return -1;
            }
        }
        for (int k = 0; k < dim; k++) newton_update(info[k], fx[0][k], fx[1][k], fx[2][k]);
    }
    return 0;
}

#ifndef ATDE_GHA_SMALL_MAX
#define ATDE_GHA_SMALL_MAX 3              // fits of up to this many tones keep every accumulator in registers
                                          // (measured per 250,880 frames: 3 -> 253 ms; 2 -> 279 ms; 3 blocks per SM at
                                          //  80 registers -> 287 ms with 3, 300 ms with 2)
#endif
#ifndef ATDE_GHA_NEWTON_UNROLL
#define ATDE_GHA_NEWTON_UNROLL 2
#endif
constexpr int kNewtonUnroll = ATDE_GHA_NEWTON_UNROLL;   // sample-loop unroll of the register-resident Newton fits
// The same fit for 1..3 tones — by far the most frequent sizes — with one pass over the samples per Newton
// loop: every sample's sin/cos pairs, the residual and all matrix terms are produced together and the
// accumulators live in registers, so the [tone][sample] sin/cos arrays (the bulk of the local-memory traffic of
// the general version) are never stored.  Every accumulator still adds its terms in the order n = 0, 1, ...
template <int DIM>
ATDE_D int adjust_newton_small(const float* pcm, GhaInfo* info, int sz, TaskScratch* ws)
{
    float* tmp = ws->tmp;
    constexpr int col = DIM + 1;
    for (int loop = 0; loop < 7; loop++) {
        double aa[DIM], ww[DIM], pp[DIM], ra[DIM], rw[DIM], rp[DIM];
        double xa[DIM][DIM], xw[DIM][DIM], xp[DIM][DIM];             // only j > i used
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            aa[i] = ww[i] = pp[i] = ra[i] = rw[i] = rp[i] = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; j++) xa[i][j] = xw[i][j] = xp[i][j] = 0.0;
        }
        float fn = 0.0f;
        double dn = 0.0;
#pragma unroll kNewtonUnroll
        for (int n = 0; n < sz; n++, fn += 1.0f, dn += 1.0) {
            float sf[DIM], cf[DIM];
            float t = pcm[n];
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                g_sincosf(fadd(fmul(info[k].frequency, fn), info[k].phase), sf[k], cf[k]);
                t = fsub(t, fmul(info[k].magnitude, sf[k]));
            }
            tmp[n] = t;
            const double td = (double)t;
            double bw[DIM], bp[DIM];
#pragma unroll
            for (int i = 0; i < DIM; i++) {
                const double Ai = (double)info[i].magnitude, c = (double)cf[i], sd = (double)sf[i];
                aa[i] = dadd(aa[i], dmul(-sd, -sd));
                bw[i] = dmul(dmul(-Ai, dn), c);
                const double bww = dmul(dmul(dmul(Ai, dn), dn), sd);
                ww[i] = dadd(ww[i], dadd(dmul(td, bww), dmul(bw[i], bw[i])));
                bp[i] = dmul(-Ai, c);
                const float bpp = d2f(dmul(Ai, sd));
                pp[i] = dadd(pp[i], dadd((double)fmul(t, bpp), dmul(bp[i], bp[i])));
                ra[i] = dadd(ra[i], (double)fmul(t, -sf[i]));
                rw[i] = dadd(rw[i], (double)fmul(t, d2f(bw[i])));
                rp[i] = dadd(rp[i], (double)fmul(t, d2f(bp[i])));
            }
#pragma unroll
            for (int i = 0; i < DIM; i++)
#pragma unroll
                for (int j = i + 1; j < DIM; j++) {
                    xa[i][j] = dadd(xa[i][j], dmul(-(double)sf[i], -(double)sf[j]));
                    xw[i][j] = dadd(xw[i][j], dmul(bw[i], bw[j]));
                    xp[i][j] = dadd(xp[i][j], dmul(bp[i], bp[j]));
                }
        }
        double M[3][DIM * col], fx[3][DIM];
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            M[0][i * col + i] = dmul(aa[i], 2.0); M[1][i * col + i] = dmul(ww[i], 2.0); M[2][i * col + i] = dmul(pp[i], 2.0);
            M[0][i * col + DIM] = dmul(ra[i], 2.0); M[1][i * col + DIM] = dmul(rw[i], 2.0); M[2][i * col + DIM] = dmul(rp[i], 2.0);
#pragma unroll
            for (int j = i + 1; j < DIM; j++) {
                M[0][i * col + j] = M[0][j * col + i] = dmul(xa[i][j], 2.0);
                M[1][i * col + j] = M[1][j * col + i] = dmul(xw[i][j], 2.0);
                M[2][i * col + j] = M[2][j * col + i] = dmul(xp[i][j], 2.0);
            }
        }
        for (int blk = 0; blk < 3; blk++) {
            for (int i = 0; i < DIM; i++) fx[blk][i] = 0.0;
            if (sle_block(M[blk], DIM, fx[blk])) return -1;
        }
        for (int k = 0; k < DIM; k++) newton_update(info[k], fx[0][k], fx[1][k], fx[2][k]);
    }
    return 0;
}

// GenWaves (at3p_gha.cpp:476-490) over the 64 look-ahead samples + the energy test of CheckNextFrame (:780-813)
ATDE_D bool check_next_frame(const GhaTables* G, const float* next_src, const GhaInfo* tones, int n)
{
    float buf[64];
    for (int i = 0; i < 64; i++) buf[i] = 0.0f;
    for (int w = 0; w < n; w++) {
        const float amp = G->amp_sf_tab[amplitude_to_sf(G, tones[w].magnitude)];
        const int inc = (int)freq_to_index(tones[w].frequency, 0);
        int pos = ((int)((phase_to_index(tones[w].phase) & 0x1f) << 6) + (0 ^ 128) * inc) & 2047;
        for (int i = 0; i < 64; i++) {
            buf[i] = fadd(buf[i], fmul(G->sine_tab[pos], amp));
            pos = (pos + inc) & 2047;
        }
    }
    float before = 0.0f, after = 0.0f;
    for (int i = 0; i < 64; i++) {
        before = fadd(before, fmul(next_src[i], next_src[i]));
        const float t = fsub(next_src[i], buf[i]);
        after = fadd(after, fmul(t, t));
    }
    return after < before;
}

// One (channel, subband) step of DoRound computed from the start-of-round state into `st`, in two halves so
// that the kernel can run each half over a list of its own (every lane of a warp then executes the same code).
// First half: re-fit the subband's tones (nothing to do without tones).  Returns false when the step ends
// here (the fit failed: the commit pass drops the last added tone and closes the subband).
ATDE_D bool task_fit(const GhaTables* G, const SbState& sbs, int sb, const float* src, const float* next_src,
                     float* buf_new, TaskScratch* ws, Staged& st)
{
    st.part1 = 0; st.n_new = 0; st.resid_valid = 0; st.analyzed = 0; st.psy_ok = 0;
    st.env_first = sbs.env_first; st.env_second = sbs.env_second;
    st.last_res_energy = sbs.last_res_energy; st.gapless = sbs.gapless; st.max_mag = sbs.max_mag;
    {
        const int dim = sbs.n;
        GhaInfo tmp_info[kMaxDim];
        for (int i = 0; i < dim; i++) tmp_info[i] = sbs.info[i];
        // do { gha_adjust_info(...) } while (Repeat)   (at3p_gha.cpp:840-846, CheckResuidalAndApply :492-579)
        int status = 1;                                  // 0 Error, 1 Ok, 2 Repeat
        int frame_sz = 0;
        for (int call = 0; call < 2; call++) {
            const int sz = (frame_sz && frame_sz < 128) ? frame_sz : 128;
            int ar;
            if (dim == 1) ar = adjust_newton_small<1>(src, tmp_info, sz, ws);
            else if (dim == 2) ar = adjust_newton_small<2>(src, tmp_info, sz, ws);
#if ATDE_GHA_SMALL_MAX >= 3
            else if (dim == 3) ar = adjust_newton_small<3>(src, tmp_info, sz, ws);
#endif
            else ar = adjust_newton(src, tmp_info, dim, sz, ws);
            if (ar < 0) { status = 0; break; }
            // the callback reads 128 samples: beyond sz they are what the previous (full-size) call left in
            // tmp_buf, which is exactly what ws->tmp still holds there
            float res_energy = 0.0f;
            unsigned start = 0, cur_start = 0, count = 0, len = 0;
            bool found = false;
            for (int i = 0; i < 128; i += 4) {
                float ein = 0.0f, eout = 0.0f;
                for (int j = 0; j < 4; j++) {
                    ein = fadd(ein, fmul(src[i + j], src[i + j]));
                    eout = fadd(eout, fmul(ws->tmp[i + j], ws->tmp[i + j]));
                }
                ein = __fsqrt_rn(__fdiv_rn(ein, 4.0f));
                eout = __fsqrt_rn(__fdiv_rn(eout, 4.0f));
                res_energy = fadd(res_energy, eout);
                if (__fdiv_rn(ein, eout) < 1.0f) {
                    count = 0; found = false; cur_start = (unsigned)i + 4;
                } else {
                    count++;
                    if (count > len) { len = count; if (!found) { start = cur_start; found = true; } }
                }
            }
            if (len < 4) { status = 0; break; }
            const unsigned end = start + len * 4;
            if (status != 2 && end != 128) {
                frame_sz = (int)end; status = 2;
                continue;
            }
            if (st.last_res_energy == 0.0f) {               // static_cast<bool>(x) == false
                st.last_res_energy = res_energy;
            } else if (st.last_res_energy < fmul(res_energy, 1.05f)) {
                status = 0; break;
            } else {
                st.last_res_energy = res_energy;
            }
            st.env_first = start;
            if (st.env_second == kEmpty && end != 128) { status = 0; break; }
            st.env_second = end;
            status = 1;
            for (int i = 0; i < 128; i++) buf_new[i] = ws->tmp[i];
            st.resid_valid = 1;
            break;
        }
        bool ok = status == 1;                          // (the second call never asks for another repeat)
        if (ok) {
            // std::sort by frequency (unique result unless two frequencies are equal, which the duplicate test rejects)
            for (int i = 1; i < dim; i++) {
                const GhaInfo v = tmp_info[i];
                int j = i - 1;
                while (j >= 0 && v.frequency < tmp_info[j].frequency) { tmp_info[j + 1] = tmp_info[j]; j--; }
                tmp_info[j + 1] = v;
            }
            bool dup = false;
            unsigned idx1 = freq_to_index(tmp_info[0].frequency, (unsigned)sb);
            for (int i = 1; i < dim; i++) {
                const unsigned idx2 = freq_to_index(tmp_info[i].frequency, (unsigned)sb);
                if (idx2 == idx1) { dup = true; break; }
                idx1 = idx2;
            }
            if (dup) ok = false;
            if (ok && (st.env_second == 128u || st.env_second == kEmpty)) {
                const bool cont = check_next_frame(G, next_src, tmp_info, dim);
                if (st.gapless && !cont) ok = false;
                else if (st.env_second == 128u && cont) { st.env_second = kEmpty; st.gapless = 1; }
            }
        }
        if (!ok) { st.part1 = 2; return false; }
        st.part1 = 1;
        st.n_new = dim;
        for (int i = 0; i < dim; i++) {
            st.fit[i] = tmp_info[i];
            st.max_mag = fmaxf(st.max_mag, tmp_info[i].magnitude);
        }
    }
    return true;
}

// Second half: extract the next tone from the residual (the staged one if the fit just replaced it).
ATDE_D void task_analyze(const GhaTables* G, int sb, const float* buf, const float* buf_new, TaskScratch* ws, Staged& st)
{
    const float* analysis_src = st.resid_valid ? buf_new : buf;
    st.found = analyze_one(G, analysis_src, ws);
    st.analyzed = 1;
    // PsyPreCheck (:955-973)
    const float mg = st.found.magnitude;
    st.psy_ok = !(mg != mg) && fmul(mg, mg) > G->subband_ath[sb] && mg > __fdiv_rn(st.max_mag, 10.0f);
}

// map<uint32_t, gha_info> of one channel = the eight per-subband lists; helpers for the commit pass
ATDE_D void sb_erase_key(SbState& s, unsigned key)
{
    for (int i = 0; i < s.n; i++)
        if (s.key[i] == key) {
            for (int j = i; j + 1 < s.n; j++) { s.key[j] = s.key[j + 1]; s.info[j] = s.info[j + 1]; }
            s.n--;
            return;
        }
}
ATDE_D bool sb_insert(SbState& s, unsigned key, const GhaInfo& v)     // map::insert: keeps the old element on a key clash
{
    int pos = 0;
    while (pos < s.n && s.key[pos] < key) pos++;
    if (pos < s.n && s.key[pos] == key) return false;
    if (s.n >= kMaxDim) return false;
    for (int j = s.n; j > pos; j--) { s.key[j] = s.key[j - 1]; s.info[j] = s.info[j - 1]; }
    s.key[pos] = key; s.info[pos] = v; s.n++;
    return true;
}

// The search kernel.  A block works on kGhaFB frames at a time.  Every round it compacts the (frame, channel,
// subband) steps that still have work into a list and spreads that list over its threads — subbands and
// frames need very different numbers of rounds, and a fixed step-per-lane mapping left three quarters of
// the lanes idle — then one thread per frame commits that frame's staged steps in the reference's order.
#ifndef ATDE_GHA_THREADS
#define ATDE_GHA_THREADS 256
#endif
#ifndef ATDE_GHA_MINBLOCKS
#define ATDE_GHA_MINBLOCKS 2              // resident blocks per SM the register budget is capped for (128 registers)
#endif
constexpr int kGhaThreads = ATDE_GHA_THREADS;
constexpr int kGhaFB = kGhaThreads;       // most frames per block batch (one committing thread per frame); the launch picks fb <= kGhaFB
constexpr int kGhaItems = kGhaFB * kGhaTask;

struct alignas(16) ItemState {            // global memory, per (frame slot, channel, subband) of a block; float4 copies
    float buf[128];                       // Buf[sb]: residual so far
    float buf_new[128];                   // staged residual of the running step
    SbState sb;
    Staged st;
};

__device__ float g_zero64[64];            // look-ahead of the last frame of a stage-test run

__global__ void __launch_bounds__(kGhaThreads, ATDE_GHA_MINBLOCKS) at3p_gha_search_kernel(const GhaTables* __restrict__ G,
                                                                       const float* __restrict__ bands,
                                                                       int S, int C, int F, int L, int j0, int fb,
                                                                       ItemState* items_g, GhaFrameOut* out)
{
    // bands [S][C][L][2048]; analysis (s, f), f < F, reads frame j0 + f with look-ahead frame j0 + f + 1 (zeros past L)
    __shared__ int s_total[kGhaFB], s_go[kGhaFB];
    __shared__ unsigned short s_list[kGhaItems];      // steps with work this round, then: steps to analyse
    __shared__ unsigned short s_fit[kGhaItems];       // steps with tones to re-fit, grouped by tone count
    __shared__ unsigned char s_adopt[kGhaItems];
    __shared__ int s_n, s_nfit, s_cnt[kMaxDim + 1], s_off[kMaxDim + 1];
    __shared__ int s_next;                            // dynamic work fetch: warps take 32 list entries at a time
    const int tid = threadIdx.x;
    const int n_items = fb * kGhaTask;                // steps of this block's batch
    const long long n_frames = (long long)S * F;
    TaskScratch scratch_local;
    TaskScratch* ws = &scratch_local;
    ItemState* items = items_g + (size_t)blockIdx.x * kGhaItems;
    for (long long base = (long long)blockIdx.x * fb; base < n_frames; base += (long long)gridDim.x * fb) {
        for (int idx = tid; idx < n_items; idx += kGhaThreads) {
            const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
            const bool live = base + fs < n_frames && ch < C;
            SbState& me = items[idx].sb;
            me.n = 0;
            // pair<> Envelopes[SUBBANDS] = {{INIT, INIT}}: only element 0 gets INIT, the rest are value-initialised
            me.env_first = sb == 0 ? kInit : 0u;
            me.env_second = sb == 0 ? kInit : 0u;
            me.gapless = 0; me.done = live ? 0 : 16; me.max_mag = 0.0f; me.last_res_energy = 0.0f; me.last_added = 0;
            s_adopt[idx] = 0;
        }
        for (int w = tid; w < n_items * 32; w += kGhaThreads) {        // Buf[sb] = the subband's samples (float4 granules)
            const int idx = w >> 5, q = w & 31;
            const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
            const long long frame = base + fs;
            if (frame < n_frames && ch < C) {
                const int s = (int)(frame / F), f = (int)(frame % F);
                const float* src = bands + (((size_t)s * C + ch) * L + j0 + f) * kFrame + sb * kSbSamples;
                reinterpret_cast<float4*>(items[idx].buf)[q] = reinterpret_cast<const float4*>(src)[q];
            }
        }
        if (tid < fb) { s_total[tid] = 0; s_go[tid] = base + tid < n_frames; }
        __syncthreads();
        for (;;) {
            if (tid == 0) { s_n = 0; s_nfit = 0; s_next = 0; }
            if (tid <= kMaxDim) s_cnt[tid] = 0;
            __syncthreads();
            // steps with work; those with tones are counted per tone count (a counting sort keeps warps homogeneous)
            int my_active = 0;
            for (int idx = tid; idx < n_items; idx += kGhaThreads)
                if (s_go[idx >> 4] && items[idx].sb.done != 16) {
                    my_active++;
                    const int n = items[idx].sb.n;
                    if (n > 0) atomicAdd(&s_cnt[n], 1);
                }
            if (my_active) atomicAdd(&s_n, my_active);
            __syncthreads();
            if (s_n == 0) break;
            if (tid == 0) {
                int acc = 0;
                for (int d = kMaxDim; d >= 1; d--) { s_off[d] = acc; acc += s_cnt[d]; }     // largest fits first
                s_nfit = acc;
            }
            __syncthreads();
            for (int idx = tid; idx < n_items; idx += kGhaThreads)
                if (s_go[idx >> 4] && items[idx].sb.done != 16 && items[idx].sb.n > 0)
                    s_fit[atomicAdd(&s_off[items[idx].sb.n], 1)] = (unsigned short)idx;
            if (tid == 0) s_n = 0;
            __syncthreads();
            // first half: fits
            const int n_fit = s_nfit;
            for (;;) {
                // the list is sorted by tone count, largest first: a warp's 32 entries cost about the same, and the
                // expensive ones start first
                int it0 = 0;
                if ((tid & 31) == 0) it0 = atomicAdd(&s_next, 32);
                it0 = __shfl_sync(0xffffffffu, it0, 0);
                if (it0 >= n_fit) break;
                const int it = it0 + (tid & 31);
                if (it >= n_fit) continue;
                const int idx = s_fit[it];
                const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
                const long long frame = base + fs;
                const int s = (int)(frame / F), f = (int)(frame % F);
                const float* src = bands + (((size_t)s * C + ch) * L + j0 + f) * kFrame + sb * kSbSamples;
                const float* next_src = j0 + f + 1 < L ? src + kFrame : g_zero64;
                if (task_fit(G, items[idx].sb, sb, src, next_src, items[idx].buf_new, ws, items[idx].st))
                    s_list[atomicAdd(&s_n, 1)] = (unsigned short)idx;
            }
            // steps without tones go straight to the analysis
            for (int idx = tid; idx < n_items; idx += kGhaThreads)
                if (s_go[idx >> 4] && items[idx].sb.done != 16 && items[idx].sb.n == 0) {
                    Staged& st = items[idx].st;
                    const SbState& sbs = items[idx].sb;
                    st.part1 = 0; st.n_new = 0; st.resid_valid = 0; st.analyzed = 0; st.psy_ok = 0;
                    st.env_first = sbs.env_first; st.env_second = sbs.env_second;
                    st.last_res_energy = sbs.last_res_energy; st.gapless = sbs.gapless; st.max_mag = sbs.max_mag;
                    s_list[atomicAdd(&s_n, 1)] = (unsigned short)idx;
                }
            __syncthreads();
            if (tid == 0) s_next = 0;
            __syncthreads();
            // second half: analyses
            const int n_list = s_n;
            for (;;) {
                int it0 = 0;
                if ((tid & 31) == 0) it0 = atomicAdd(&s_next, 32);
                it0 = __shfl_sync(0xffffffffu, it0, 0);
                if (it0 >= n_list) break;
                const int it = it0 + (tid & 31);
                if (it >= n_list) continue;
                const int idx = s_list[it];
                task_analyze(G, idx & 7, items[idx].buf, items[idx].buf_new, ws, items[idx].st);
            }
            __syncthreads();
            if (tid < fb && s_go[tid]) {
                // commit in the reference's order: channel 0 subbands 0..7, then channel 1
                const int fs = tid;
                int total = s_total[fs];
                bool progress[2] = {false, false};
                for (int c2 = 0; c2 < C; c2++) {
                    bool prog = false;
                    for (int b2 = 0; b2 < kGhaSb; b2++) {
                        SbState& z = items[fs * 16 + c2 * 8 + b2].sb;
                        if (z.done == 16) continue;
                        if (total >= 48) { prog = false; break; }          // return false
                        const Staged& g = items[fs * 16 + c2 * 8 + b2].st;
                        if (g.part1 != 0) {
                            // what CheckResuidalAndApply / the look-ahead test wrote, also on their failure paths
                            z.env_first = g.env_first; z.env_second = g.env_second;
                            z.last_res_energy = g.last_res_energy; z.gapless = g.gapless;
                            if (g.part1 == 2) {                            // :866-871, :899-910: drop the last added tone, done
                                sb_erase_key(z, z.last_added);
                                total--;
                                z.done = 16;
                                continue;
                            }
                            z.n = 0;                                       // :888-897: replace the subband's tones
                            for (int i = 0; i < g.n_new; i++) {
                                z.max_mag = fmaxf(z.max_mag, g.fit[i].magnitude);
                                sb_insert(z, freq_to_index(g.fit[i].frequency, (unsigned)b2), g.fit[i]);
                            }
                            if (g.resid_valid) s_adopt[fs * 16 + c2 * 8 + b2] = 1;
                        }
                        const unsigned fi = freq_to_index(g.found.frequency, (unsigned)b2);
                        if (!g.psy_ok) { z.done = 16; continue; }
                        if (z.done == 0) {
                            sb_insert(z, fi, g.found);
                            z.last_added = fi;
                        } else {
                            // lower_bound over the channel's whole map
                            unsigned next_key = 0, prev_key = 0;
                            bool has_nxt = false, has_prev = false;
                            for (int b3 = 0; b3 < kGhaSb; b3++) {
                                const SbState& y = items[fs * 16 + c2 * 8 + b3].sb;
                                for (int i = 0; i < y.n; i++) {
                                    if (y.key[i] >= fi) { if (!has_nxt) { has_nxt = true; next_key = y.key[i]; } }
                                    else { has_prev = true; prev_key = y.key[i]; }
                                }
                            }
                            if (has_nxt && (next_key == fi || next_key - fi < 20u)) { z.done = 16; continue; }
                            if (has_prev && fi - prev_key < 20u) { z.done = 16; continue; }
                            if (z.done == 15) { z.done = 16; continue; }
                            sb_insert(z, fi, g.found);
                            z.last_added = fi;
                        }
                        z.done++;
                        total++;
                        prog = true;
                    }
                    progress[c2] = prog;
                }
                s_total[fs] = total;
                s_go[fs] = (progress[0] || progress[1]) && total < 48;
            }
            __syncthreads();
            // adopt the staged residual where the callback accepted it and the step was committed
            for (int w = tid; w < n_list * 32; w += kGhaThreads) {
                const int idx = s_list[w >> 5], q = w & 31;
                if (s_adopt[idx])
                    reinterpret_cast<float4*>(items[idx].buf)[q] = reinterpret_cast<const float4*>(items[idx].buf_new)[q];
            }
            __syncthreads();
            for (int it = tid; it < n_list; it += kGhaThreads) s_adopt[s_list[it]] = 0;
        }
        for (int idx = tid; idx < n_items; idx += kGhaThreads) {
            const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
            const long long frame = base + fs;
            if (frame < n_frames && ch < C) {
                const SbState& me = items[idx].sb;
                GhaFrameOut& o = out[frame];
                if (t == 0) o.total_tones = s_total[fs];
                o.n[ch][sb] = me.n;
                for (int i = 0; i < me.n; i++) { o.key[ch][sb][i] = me.key[i]; o.info[ch][sb][i] = me.info[i]; }
                o.env[ch][sb][0] = me.env_first; o.env[ch][sb][1] = me.env_second;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// EXPERIMENT (not launched unless built with -DATDE_GHA_DYNAMIC; see launch_gha_search).
// The same search without block-wide rounds.  Frames are independent, and inside a frame the only ordering is
// "all steps of a round, then the commit, then the next round" — so a frame advances on its own: every step
// decrements its frame's counter when it finishes, the thread that brings it to zero commits the frame and
// queues the next round's steps.  Warps pull up to 32 steps at a time from five queues (analyses; fits of 1, 2, 3
// and >= 4 tones, so that a warp's lanes run the same code).  No barrier inside a batch: the tail rounds of
// slow frames overlap with the busy rounds of others.
// ---------------------------------------------------------------------------------------------
#ifdef ATDE_GHA_DYNAMIC
constexpr int kGhaQ = 5;
struct DynShared {
    unsigned short q[kGhaQ][kGhaItems];   // rings of step ids, 0xffff = empty slot
    int head[kGhaQ], tail[kGhaQ];
    int pending[kGhaFB], total[kGhaFB];
    unsigned char cur[kGhaItems];         // which of the step's two residual buffers is Buf[sb]
    int active_frames;
};

ATDE_D int vload(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

ATDE_D void dyn_push(DynShared& sh, int qi, int id)
{
    const int pos = atomicAdd(&sh.tail[qi], 1);
    *reinterpret_cast<volatile unsigned short*>(&sh.q[qi][pos & (kGhaItems - 1)]) = (unsigned short)id;
}

// queue the steps of frame slot fs's next round; returns false if there is none (every subband closed)
ATDE_D bool dyn_start_round(DynShared& sh, ItemState* items, int fs)
{
    int pend = 0;
    for (int t = 0; t < kGhaTask; t++) pend += items[fs * 16 + t].sb.done != 16;
    if (pend == 0) return false;
    sh.pending[fs] = pend;
    __threadfence_block();
    for (int t = 0; t < kGhaTask; t++) {
        const int idx = fs * 16 + t;
        const SbState& sbs = items[idx].sb;
        if (sbs.done == 16) continue;
        if (sbs.n == 0) {
            Staged& st = items[idx].st;
            st.part1 = 0; st.n_new = 0; st.resid_valid = 0; st.analyzed = 0; st.psy_ok = 0;
            st.env_first = sbs.env_first; st.env_second = sbs.env_second;
            st.last_res_energy = sbs.last_res_energy; st.gapless = sbs.gapless; st.max_mag = sbs.max_mag;
            __threadfence_block();
            dyn_push(sh, 0, idx);
        } else {
            dyn_push(sh, sbs.n < 4 ? sbs.n : 4, idx);
        }
    }
    return true;
}

// the commit pass of one frame (the reference's order: channel 0 subbands 0..7, then channel 1); returns `go`
ATDE_D bool dyn_commit(DynShared& sh, ItemState* items, int fs, int C)
{
    int total = sh.total[fs];
    bool progress[2] = {false, false};
    for (int c2 = 0; c2 < C; c2++) {
        bool prog = false;
        for (int b2 = 0; b2 < kGhaSb; b2++) {
            const int idx = fs * 16 + c2 * 8 + b2;
            SbState& z = items[idx].sb;
            if (z.done == 16) continue;
            if (total >= 48) { prog = false; break; }                      // return false
            const Staged& g = items[idx].st;
            if (g.part1 != 0) {
                z.env_first = g.env_first; z.env_second = g.env_second;
                z.last_res_energy = g.last_res_energy; z.gapless = g.gapless;
                if (g.part1 == 2) {
                    sb_erase_key(z, z.last_added);
                    total--;
                    z.done = 16;
                    continue;
                }
                z.n = 0;
                for (int i = 0; i < g.n_new; i++) {
                    z.max_mag = fmaxf(z.max_mag, g.fit[i].magnitude);
                    sb_insert(z, freq_to_index(g.fit[i].frequency, (unsigned)b2), g.fit[i]);
                }
                if (g.resid_valid) sh.cur[idx] ^= 1;                      // the staged residual becomes Buf[sb]
            }
            const unsigned fi = freq_to_index(g.found.frequency, (unsigned)b2);
            if (!g.psy_ok) { z.done = 16; continue; }
            if (z.done == 0) {
                sb_insert(z, fi, g.found);
                z.last_added = fi;
            } else {
                unsigned next_key = 0, prev_key = 0;
                bool has_nxt = false, has_prev = false;
                for (int b3 = 0; b3 < kGhaSb; b3++) {
                    const SbState& y = items[fs * 16 + c2 * 8 + b3].sb;
                    for (int i = 0; i < y.n; i++) {
                        if (y.key[i] >= fi) { if (!has_nxt) { has_nxt = true; next_key = y.key[i]; } }
                        else { has_prev = true; prev_key = y.key[i]; }
                    }
                }
                if (has_nxt && (next_key == fi || next_key - fi < 20u)) { z.done = 16; continue; }
                if (has_prev && fi - prev_key < 20u) { z.done = 16; continue; }
                if (z.done == 15) { z.done = 16; continue; }
                sb_insert(z, fi, g.found);
                z.last_added = fi;
            }
            z.done++;
            total++;
            prog = true;
        }
        progress[c2] = prog;
    }
    sh.total[fs] = total;
    return (progress[0] || progress[1]) && total < 48;
}

// a step of frame slot fs has finished; the last one of the round commits and moves the frame on
ATDE_D void dyn_step_done(DynShared& sh, ItemState* items, int fs, int C)
{
    __threadfence_block();
    if (atomicSub(&sh.pending[fs], 1) != 1) return;
    __threadfence_block();
    const bool go = dyn_commit(sh, items, fs, C);
    __threadfence_block();
    if (!go || !dyn_start_round(sh, items, fs)) atomicSub(&sh.active_frames, 1);
}

__global__ void __launch_bounds__(kGhaThreads, ATDE_GHA_MINBLOCKS) at3p_gha_search_dyn_kernel(const GhaTables* __restrict__ G,
                                                                       const float* __restrict__ bands,
                                                                       int S, int C, int F, int L, int j0, int fb,
                                                                       ItemState* items_g, GhaFrameOut* out)
{
    __shared__ DynShared sh;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_items = fb * kGhaTask;
    const long long n_frames = (long long)S * F;
    TaskScratch scratch_local;
    TaskScratch* ws = &scratch_local;
    ItemState* items = items_g + (size_t)blockIdx.x * kGhaItems;
    for (int i = tid; i < kGhaQ * kGhaItems; i += kGhaThreads) (&sh.q[0][0])[i] = 0xffffu;
    if (tid < kGhaQ) { sh.head[tid] = 0; sh.tail[tid] = 0; }
    __syncthreads();
    for (long long base = (long long)blockIdx.x * fb; base < n_frames; base += (long long)gridDim.x * fb) {
        for (int idx = tid; idx < n_items; idx += kGhaThreads) {
            const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
            const bool live = base + fs < n_frames && ch < C;
            SbState& me = items[idx].sb;
            me.n = 0;
            me.env_first = sb == 0 ? kInit : 0u;               // Envelopes[SUBBANDS] = {{INIT, INIT}}: element 0 only
            me.env_second = sb == 0 ? kInit : 0u;
            me.gapless = 0; me.done = live ? 0 : 16; me.max_mag = 0.0f; me.last_res_energy = 0.0f; me.last_added = 0;
            sh.cur[idx] = 0;
        }
        for (int w = tid; w < n_items * 32; w += kGhaThreads) {
            const int idx = w >> 5, q = w & 31;
            const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
            const long long frame = base + fs;
            if (frame < n_frames && ch < C) {
                const int s = (int)(frame / F), f = (int)(frame % F);
                const float* src = bands + (((size_t)s * C + ch) * L + j0 + f) * kFrame + sb * kSbSamples;
                reinterpret_cast<float4*>(items[idx].buf)[q] = reinterpret_cast<const float4*>(src)[q];
            }
        }
        if (tid == 0) {
            long long nf = n_frames - base;
            sh.active_frames = (int)(nf < fb ? nf : fb);
        }
        if (tid < fb) sh.total[tid] = 0;
        __syncthreads();
        if (tid < fb && base + tid < n_frames)
            if (!dyn_start_round(sh, items, tid)) atomicSub(&sh.active_frames, 1);
        __syncthreads();
        for (;;) {
            int qi = -1, h0 = 0, take = 0;
            if (lane == 0) {
                for (;;) {
                    int best = -1, bestn = 0;
                    for (int k = 0; k < kGhaQ; k++) {
                        const int n = vload(&sh.tail[k]) - vload(&sh.head[k]);
                        if (n > bestn) { bestn = n; best = k; }
                    }
                    if (best >= 0) {
                        const int h = vload(&sh.head[best]);
                        const int n = vload(&sh.tail[best]) - h;
                        if (n <= 0) continue;
                        const int tk = n < 32 ? n : 32;
                        if (atomicCAS(&sh.head[best], h, h + tk) == h) { qi = best; h0 = h; take = tk; break; }
                        continue;
                    }
                    if (vload(&sh.active_frames) <= 0) { qi = -2; break; }
                    __nanosleep(100);                       // steps of other warps are still running
                }
            }
            qi = __shfl_sync(0xffffffffu, qi, 0);
            h0 = __shfl_sync(0xffffffffu, h0, 0);
            take = __shfl_sync(0xffffffffu, take, 0);
            if (qi == -2) break;
            if (lane < take) {
                volatile unsigned short* slot = &sh.q[qi][(h0 + lane) & (kGhaItems - 1)];
                unsigned id;
                while ((id = *slot) == 0xffffu) __nanosleep(20);
                *slot = 0xffffu;
                __threadfence_block();
                const int idx = (int)id;
                const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
                float* bcur = sh.cur[idx] ? items[idx].buf_new : items[idx].buf;
                float* bnew = sh.cur[idx] ? items[idx].buf : items[idx].buf_new;
                if (qi == 0) {
                    task_analyze(G, sb, bcur, bnew, ws, items[idx].st);
                    dyn_step_done(sh, items, fs, C);
                } else {
                    const long long frame = base + fs;
                    const int s = (int)(frame / F), f = (int)(frame % F);
                    const float* src = bands + (((size_t)s * C + ch) * L + j0 + f) * kFrame + sb * kSbSamples;
                    const float* next_src = j0 + f + 1 < L ? src + kFrame : g_zero64;
                    if (task_fit(G, items[idx].sb, sb, src, next_src, bnew, ws, items[idx].st)) {
                        __threadfence_block();
                        dyn_push(sh, 0, idx);
                    } else {
                        dyn_step_done(sh, items, fs, C);
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();
        for (int idx = tid; idx < n_items; idx += kGhaThreads) {
            const int fs = idx >> 4, t = idx & 15, ch = t >> 3, sb = t & 7;
            const long long frame = base + fs;
            if (frame < n_frames && ch < C) {
                const SbState& me = items[idx].sb;
                GhaFrameOut& o = out[frame];
                if (t == 0) o.total_tones = sh.total[fs];
                o.n[ch][sb] = me.n;
                for (int i = 0; i < me.n; i++) { o.key[ch][sb][i] = me.key[i]; o.info[ch][sb][i] = me.info[i]; }
                o.env[ch][sb][0] = me.env_first; o.env[ch][sb][1] = me.env_second;
            }
        }
        __syncthreads();
    }
}
#endif // ATDE_GHA_DYNAMIC

// ---------------------------------------------------------------------------------------------
// FillResultBuf / FillFolowerRes / AdjustEnvelope (at3p_gha.cpp:1499-1664) + the ResultBufHistory carry:
// one thread per stream walks its frames in order.
// ---------------------------------------------------------------------------------------------
struct GhaHistory {                       // what later frames read of ResultBufHistory
    int n_sb[2];
    unsigned env_second[2][16];
};

constexpr unsigned kNeedHist = 0xfffffffdu;   // envelope start that depends on the previous valid result (patched by the history pass)

ATDE_D void adjust_envelope(int* env /*[2]*/, unsigned src_first, unsigned src_second)
{
    // AdjustEnvelope(.., history): first = (src.first == 0 && history == EMPTY) ? EMPTY : src.first / 4
    if (src_first == 0) env[0] = (int)kNeedHist;
    else env[0] = (int)(src_first / 4);
    if (src_second == kEmpty) env[1] = (int)kEmpty;
    else env[1] = (int)((src_second - 1) / 4);
}

// One thread per frame: everything of FillResultBuf that does not look at ResultBufHistory.
__global__ void at3p_gha_result_kernel(const GhaTables* __restrict__ G, const GhaFrameOut* __restrict__ in,
                                       int S, int C, int F, ToneBlock* out, int out_stride, int out_off)
{
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (long long)S * F) return;
    const int s = (int)(gi / F), f = (int)(gi % F);
    {
        const GhaFrameOut& o = in[(size_t)s * F + f];
        ToneBlock& tb = out[(size_t)s * out_stride + out_off + f];
        tb.present = 0; tb.num_tone_bands = 0; tb.second_is_leader = 0;
        tb.n_sb[0] = tb.n_sb[1] = 0; tb.n_params[0] = tb.n_params[1] = 0;
        for (int i = 0; i < 16; i++) tb.tone_sharing[i] = 0;
        if (o.total_tones == 0) return;                        // DoAnalize returns nullptr, history untouched
        int used[2] = {0, 0};
        for (int ch = 0; ch < C; ch++)
            for (int sb = 0; sb < kGhaSb; sb++) if (o.n[ch][sb] > 0) used[ch] = sb + 1;
        const int leader = used[1] > used[0] ? 1 : 0;
        const int ntb = used[leader];
        tb.present = 1; tb.second_is_leader = leader; tb.num_tone_bands = ntb;
        tb.n_sb[0] = ntb;
        if (C == 2) tb.n_sb[1] = ntb;
        for (int ch = 0; ch < 2; ch++)
            for (int sb = 0; sb < 16; sb++) { tb.sb[ch][sb][0] = 0; tb.sb[ch][sb][1] = 0; tb.sb[ch][sb][2] = (int)kEmpty; tb.sb[ch][sb][3] = (int)kEmpty; }
        int np0 = 0, np1 = 0;
        const int fol = 1 - leader;
        for (int sb = 0; sb < ntb; sb++) {
            const int index = np0;
            for (int i = 0; i < o.n[leader][sb]; i++) {
                int* prm = tb.params[0][np0++];
                prm[0] = (int)(o.key[leader][sb][i] & 1023u);
                prm[1] = (int)amplitude_to_sf(G, o.info[leader][sb][i].magnitude);
                prm[2] = 1;
                prm[3] = (int)phase_to_index(o.info[leader][sb][i].phase);
                tb.sb[0][sb][1]++;
            }
            if (tb.sb[0][sb][1] > 0) {
                tb.sb[0][sb][0] = index;
                adjust_envelope(&tb.sb[0][sb][2], o.env[leader][sb][0], o.env[leader][sb][1]);
            }
            if (C == 2) {
                unsigned mode = 0;
                int added = 0;
                for (int i = 0; i < o.n[fol][sb]; i++) {
                    const unsigned key = o.key[fol][sb][i];
                    bool in_leader = false;
                    for (int j = 0; j < o.n[leader][sb]; j++) in_leader |= o.key[leader][sb][j] == key;
                    mode |= in_leader ? 1u : 2u;
                    int* prm = tb.params[1][np1++];
                    prm[0] = (int)(key & 1023u);
                    prm[1] = (int)amplitude_to_sf(G, o.info[fol][sb][i].magnitude);
                    prm[2] = 1;
                    prm[3] = (int)phase_to_index(o.info[fol][sb][i].phase);
                    added++;
                }
                if (mode == 0) { tb.tone_sharing[sb] = 0; tb.sb[1][sb][1] = 0; }
                else if (mode == 1) { tb.tone_sharing[sb] = 1; np1 -= added; }
                else {
                    tb.tone_sharing[sb] = 0;
                    tb.sb[1][sb][0] = np1 - added;
                    tb.sb[1][sb][1] = added;
                    adjust_envelope(&tb.sb[1][sb][2], o.env[fol][sb][0], o.env[fol][sb][1]);
                }
            }
        }
        tb.n_params[0] = np0; tb.n_params[1] = np1;
    }
}

// ResultBufHistory is the previous NON-NULL result; what later frames read of it (WaveSbInfos.size() and the stop
// envelopes) does not depend on history itself, so the carry is a look-back, not a recurrence: one thread per
// frame finds its predecessor (inside the batch, else the carried state) and settles the envelope starts that
// were waiting for it.
__global__ void at3p_gha_history_kernel(int S, int C, int F, const GhaHistory* hist_state, ToneBlock* out, int out_stride, int out_off)
{
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (long long)S * F) return;
    const int s = (int)(gi / F), f = (int)(gi % F);
    ToneBlock* row = out + (size_t)s * out_stride + out_off;
    ToneBlock& tb = row[f];
    if (!tb.present) return;
    int g = f - 1;
    while (g >= 0 && !row[g].present) g--;
    for (int ch = 0; ch < C; ch++)
        for (int sb = 0; sb < tb.num_tone_bands; sb++) {
            if ((unsigned)tb.sb[ch][sb][2] != kNeedHist) continue;
            unsigned hs = kInit;
            if (g >= 0) { if (row[g].n_sb[ch] > sb) hs = (unsigned)row[g].sb[ch][sb][3]; }
            else if (hist_state[s].n_sb[ch] > sb) hs = hist_state[s].env_second[ch][sb];
            tb.sb[ch][sb][2] = hs == kEmpty ? (int)kEmpty : 0;
        }
}

// after the look-backs: the last non-null result of the batch becomes the carried history
__global__ void at3p_gha_history_commit_kernel(int S, int C, int F, GhaHistory* hist_state, const ToneBlock* out, int out_stride, int out_off)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const ToneBlock* row = out + (size_t)s * out_stride + out_off;
    int g = F - 1;
    while (g >= 0 && !row[g].present) g--;
    if (g < 0) return;
    GhaHistory h = hist_state[s];
    const int ntb = row[g].num_tone_bands;
    h.n_sb[0] = ntb;
    if (C == 2) h.n_sb[1] = ntb;
    for (int ch = 0; ch < C; ch++)
        for (int sb = 0; sb < ntb; sb++) h.env_second[ch][sb] = (unsigned)row[g].sb[ch][sb][3];
    hist_state[s] = h;
}

size_t gha_scratch_bytes(int blocks) { return (size_t)blocks * kGhaItems * sizeof(ItemState); }
size_t gha_frame_out_bytes() { return sizeof(GhaFrameOut); }
size_t gha_history_bytes() { return sizeof(GhaHistory); }
// frames per block batch: as many as keep one resident wave of blocks (4 per SM) busy, at most kGhaFB
static int gha_fb_for(long long n_analyses)
{
    long long fb = (n_analyses + 148 * ATDE_GHA_MINBLOCKS - 1) / (148 * ATDE_GHA_MINBLOCKS);
    if (fb < 8) fb = 8;
    if (fb > kGhaFB) fb = kGhaFB;
    return (int)fb;
}
int gha_blocks_for(long long n_analyses)
{
    const int fb = gha_fb_for(n_analyses);
    long long b = (n_analyses + fb - 1) / fb;
    if (b > 148 * ATDE_GHA_MINBLOCKS) b = 148 * ATDE_GHA_MINBLOCKS;       // one resident wave
    return (int)(b < 1 ? 1 : b);
}
void launch_gha_search(const float* bands, int S, int C, int nA, int L, int j0, void* scratch, void* frame_out, int blocks, cudaStream_t st)
{
#ifdef ATDE_GHA_DYNAMIC
    // the barrier-free per-frame scheduler: exact, but measured slower on B200 (137 ms vs 104 ms per 62,720 frames):
    // warps often find fewer than 32 ready steps and a commit holds its whole warp; kept for the next round of work
    ATDE_LAUNCH(at3p_gha_search_dyn_kernel, (unsigned)blocks, kGhaThreads, 0, st, gha_tables(), bands, S, C, nA, L, j0,
                gha_fb_for((long long)S * nA), (ItemState*)scratch, (GhaFrameOut*)frame_out);
#else
    ATDE_LAUNCH(at3p_gha_search_kernel, (unsigned)blocks, kGhaThreads, 0, st, gha_tables(), bands, S, C, nA, L, j0,
                gha_fb_for((long long)S * nA), (ItemState*)scratch, (GhaFrameOut*)frame_out);
#endif
}
void launch_gha_result(const void* frame_out, int S, int C, int nA, void* hist_state, ToneBlock* tones, int stride, int off, cudaStream_t st)
{
    const long long n = (long long)S * nA;
    ATDE_LAUNCH(at3p_gha_result_kernel, (unsigned)((n + 63) / 64), 64, 0, st, gha_tables(), (const GhaFrameOut*)frame_out, S, C, nA,
                tones, stride, off);
    ATDE_LAUNCH(at3p_gha_history_kernel, (unsigned)((n + 63) / 64), 64, 0, st, S, C, nA, (const GhaHistory*)hist_state, tones, stride, off);
    ATDE_LAUNCH(at3p_gha_history_commit_kernel, (unsigned)((S + 63) / 64), 64, 0, st, S, C, nA, (GhaHistory*)hist_state,
                (const ToneBlock*)tones, stride, off);
}
bool gha_tables_ready() { return gha_tables() != nullptr; }

} // namespace at3p
} // namespace atde

// ---- stage entry: the whole tone search on fresh streams (tests/) ----
extern "C" int atde_at3p_stage_gha(const float* bands, int S, int C, int F, void* tones)
{
    using namespace atde::at3p;
    if (!gha_tables_ready()) return -2;
    const size_t n = (size_t)S * C * F * kFrame;
    float* d_bands = nullptr;
    unsigned char *d_scr = nullptr, *d_out = nullptr, *d_hist = nullptr;
    ToneBlock* d_tb = nullptr;
    const long long n_frames = (long long)S * F;
    const int blocks = gha_blocks_for(n_frames);
    int rc = 0;
    if (cudaMalloc(&d_bands, n * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&d_scr, gha_scratch_bytes(blocks)) != cudaSuccess ||
        cudaMalloc(&d_out, (size_t)n_frames * gha_frame_out_bytes()) != cudaSuccess ||
        cudaMalloc(&d_hist, (size_t)S * gha_history_bytes()) != cudaSuccess ||
        cudaMalloc(&d_tb, (size_t)n_frames * sizeof(ToneBlock)) != cudaSuccess) rc = -3;
    if (!rc && (cudaMemcpy(d_bands, bands, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemset(d_hist, 0, (size_t)S * gha_history_bytes()) != cudaSuccess ||
                cudaMemset(d_out, 0, (size_t)n_frames * gha_frame_out_bytes()) != cudaSuccess)) rc = -2;
    if (!rc) {
        launch_gha_search(d_bands, S, C, F, F, 0, d_scr, d_out, blocks, nullptr);
        launch_gha_result(d_out, S, C, F, d_hist, d_tb, F, 0, nullptr);
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = -2;
    }
    if (!rc && cudaMemcpy(tones, d_tb, (size_t)n_frames * sizeof(ToneBlock), cudaMemcpyDeviceToHost) != cudaSuccess) rc = -2;
    cudaFree(d_bands); cudaFree(d_scr); cudaFree(d_out); cudaFree(d_hist); cudaFree(d_tb);
    return rc;
}
