// at3_analysis.cu — ATRAC3 encode hot path on sm_100a, analysis half.
//
// Replaces, for thousands of frames at once (reference: dcherednik/atracdenc):
//   K1 at3_qmf_kernel        lambda prologue /4.0 + Atrac3AnalysisFilterBank::Analysis (src/atrac3denc.cpp:700-713,
//                            src/atrac/at3/atrac3_qmf.h:37-41, src/qmf/qmf.h:47-64), Matrixing (atrac3denc.cpp:665-677)
//   K2 at3_gain_kernel       TSpectralUpsampler::Process (src/transient_spectral_upsampler.cpp:77-180),
//                            kiss_fftr / kiss_fftri (src/lib/fft/kissfft_impl/tools/kiss_fftr.c:61-153),
//                            AnalyzeGain (src/transient_detector.cpp:95-136), FindPlateau target (:190-262,298-312)
//   K3 at3_gain_scan_kernel  the TCurveBuilderCtx recurrence (atrac3denc.cpp:319-346, transient_detector.cpp:322-325)
//   K4 at3_curve_kernel      CalcCurve (transient_detector.cpp:276-482) + CreateSubbandInfo (atrac3denc.cpp:299-579)
//   K5 at3_mdct_kernel       CalcGainEnergyScale (atrac3denc.cpp:154-224), TGainProcessor::Modulate (src/gain_processor.h:87-121),
//                            TAtrac3MDCT::Mdct (atrac3denc.cpp:33-58), TMDCT<512> (src/lib/mdct/mdct.h:51-104),
//                            per-channel loudness term (atrac3denc.cpp:811-820)
//   K6 at3_loudness_kernel   TrackLoudness (atrac3denc.cpp:833-841, src/atrac/atrac_psy_common.h:46-54)
//
// Frames of a stream are processed in parallel.  The only true recurrences are the three scalars of
// TCurveBuilderCtx per (channel, band) and the loudness scalar per stream; K3 / K6 scan them.  Every
// other piece of carried encoder state (QMF histories, look-ahead buffer, the windowed+modulated MDCT
// half, PrevOverlapGainScale) is a finite function of the neighbouring frames' band samples and gain
// curves and is recomputed from them with the reference's operations in the reference's order.
// All arithmetic is un-fused IEEE fp32 (bit-exact contract).
#include "at3_kernels.cuh"
#include "kissfft_dev.cuh"
#include "glibc_math.cuh"

namespace atde {
namespace at3 {



// 48-tap half-band split (qmf.h:54-63), register-tiled: one thread task = kQR consecutive output
// pairs j0 .. j0+kQR-1 of one source array.  Output j is
//     lower = sum_{i=0..23} W[2i]   * src[2j + 49 - 2i]        (sequential, i ascending)
//     upper = sum_{i=0..23} W[2i+1] * src[2j + 48 - 2i]
// i.e. tap pair i multiplies the ALIGNED sample pair p = j + 24 - i = (src[2p], src[2p+1]) by
// (W[2i+1], W[2i]); the two running sums advance with one packed multiply and one packed add per tap
// pair.  The kQR outputs of a task share their sample pairs (31 distinct pairs for 8 outputs), which
// are loaded once into registers: 16 x LDS.128 per 384 packed-math instructions instead of one
// LDS.64 per tap pair — the kernel is bound by the FP32 pipe, not by shared-memory bandwidth.
// Arrays are stored with one float4 of padding after every four (qphys) so that the 64-byte thread
// stride of the task loads spreads over all banks.
constexpr int kQR = 8;
ATDE_HD int qphys(int e) { return e + ((e >> 4) << 2); }

__constant__ __align__(8) float c_qmf3p[48];       // tap pairs (W[2i+1], W[2i])

void upload_qmf_window(const float w[48])
{
    float p[48];
    for (int i = 0; i < 24; i++) { p[2 * i] = w[2 * i + 1]; p[2 * i + 1] = w[2 * i]; }
    cudaMemcpyToSymbol(c_qmf3p, p, 48 * sizeof(float));
}

// src: padded array; task: outputs 8*task .. 8*task+7.  lower[r] / upper[r] as TQmf::Analysis returns them.
ATDE_D void qmf_task(const float* src, const float* cw, int task, f32x2 one, float* lower, float* upper)
{
    f32x2 xp[4 * kQR];                               // sample pairs 8*task .. 8*task+31
    const float4* q = reinterpret_cast<const float4*>(src + qphys(2 * kQR * task));
#pragma unroll
    for (int u = 0; u < 2 * kQR; u++) {
        const float4 v = q[u + (u >> 2)];
        xp[2 * u].x = v.x; xp[2 * u].y = v.y;
        xp[2 * u + 1].x = v.z; xp[2 * u + 1].y = v.w;
    }
    f32x2 acc[kQR];
#pragma unroll
    for (int r = 0; r < kQR; r++) { acc[r].x = 0.0f; acc[r].y = 0.0f; }
#pragma unroll
    for (int i = 0; i < 24; i++) {
        const float2 c = *reinterpret_cast<const float2*>(cw + 2 * i);
        f32x2 cc;
        cc.x = c.x; cc.y = c.y;
#pragma unroll
        for (int r = 0; r < kQR; r++) acc[r] = add2(acc[r], mul2(cc, xp[r + 24 - i]), one);
    }
#pragma unroll
    for (int r = 0; r < kQR; r++) {                  // acc.x = upper sum, acc.y = lower sum (qmf.h:60-62)
        upper[r] = fsub(acc[r].y, acc[r].x);
        lower[r] = fadd(acc[r].y, acc[r].x);
    }
}

// =====================================================================================
// K1: PCM -> four 11 kHz bands per channel (two QMF stages), optional M/S matrixing
// =====================================================================================
// One block = kQT consecutive band samples u of one stream; threads 0..127 work on channel 0,
// threads 128..255 on channel 1.
//   band sample u <-> m = u - 128 (sample m of the extended sequence, 256 per frame)
//   stage-1 output index k (512 per frame) feeds stage 2: band m reads s1[2m-47 .. 2m+1]
//   input sample n (1024 per frame): s1[k] reads x[2k-47 .. 2k+1]
// Tile-local arrays (m0 = first band sample of the tile):
//   s1[j] = stage-1 sample k = 2*m0 - 48 + j,  j in [0, kQS1 = 1024)  -> 128 tasks per channel
//   x[t]  = input sample n = 4*m0 - 144 + t,   t in [0, kQX = 2096)
//   stage 2: 488 outputs per half-band = 61 tasks each
constexpr int kQT = 488;
constexpr int kQS1 = 2 * kQT + 48;
constexpr int kQX = 2 * kQS1 + 48;
static_assert(kQS1 == 128 * kQR && kQT % kQR == 0, "tile geometry");
constexpr int kQXP = kQX + (kQX / 16 + 1) * 4;       // padded sizes
constexpr int kQS1P = kQS1 + (kQS1 / 16 + 1) * 4;

ATDE_D float virt_pcm(const Geometry& g, const Buffers& b, int s, int c, long long n, bool started)
{
    // sample n of the extended sequence of stream s (n < 0: the frame before ext frame 0)
    if (n < 0) {
        if (!started || n < -1024) return 0.0f;
        return b.pcm_hist[(((size_t)s * 2 + 0) * 1024 + (size_t)(n + 1024)) * g.C + c];
    }
    if (started) {
        if (n < 1024) return b.pcm_hist[(((size_t)s * 2 + 1) * 1024 + (size_t)n) * g.C + c];
        n -= 1024;
    }
    if (n >= (long long)g.N * 1024) return 0.0f;
    return b.pcm[((size_t)s * g.N * 1024 + (size_t)n) * g.C + c];
}

__global__ void __launch_bounds__(256, 4) at3_qmf_kernel(Geometry g, Buffers b)
{
    __shared__ __align__(16) float xs[2][kQXP];          // input tile per channel; later the output tile
    __shared__ __align__(16) float s1[2][2][kQS1P];      // [channel][lo, hi]
    __shared__ __align__(8) float cw[48];

    const int s = blockIdx.y;
    const int u0 = blockIdx.x * kQT;
    const int m0 = u0 - 128;
    const bool started = b.started[s] != 0;
    const int tid = threadIdx.x;
    const int ch = tid >> 7, t7 = tid & 127;
    f32x2 one;
    one.x = g.one; one.y = g.one;

    if (tid < 48) cw[tid] = c_qmf3p[tid];
    {
        const long long n0 = 4LL * m0 - 144 - (started ? 1024 : 0);    // tile sample 0 as an index into b.pcm
        const long long nlim = (long long)g.N * 1024;
        if (g.C == 2 && n0 >= 0 && n0 + kQX <= nlim) {
            // interior tile: every load independent and in flight at once
            const float2* src = reinterpret_cast<const float2*>(b.pcm + ((size_t)s * g.N * 1024 + (size_t)n0) * 2);
            constexpr int kIters = (kQX + 255) / 256;
            float2 v[kIters];
#pragma unroll
            for (int it = 0; it < kIters; it++) {
                const int t = tid + 256 * it;
                v[it] = t < kQX ? src[t] : make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int it = 0; it < kIters; it++) {
                const int t = tid + 256 * it;
                if (t < kQX) {
                    xs[0][qphys(t)] = fmul(v[it].x, 0.25f);      // data / 4.0 (atrac3denc.cpp:704)
                    xs[1][qphys(t)] = fmul(v[it].y, 0.25f);
                }
            }
        } else
        ATDE_PAR_FOR(t, kQX) {
            const long long nn = n0 + t;
            float v0, v1 = 0.0f;
            if (nn >= 0 && nn < nlim) {
                if (g.C == 2) {
                    const float2 v = *reinterpret_cast<const float2*>(b.pcm + ((size_t)s * g.N * 1024 + (size_t)nn) * 2);
                    v0 = v.x; v1 = v.y;
                } else {
                    v0 = b.pcm[(size_t)s * g.N * 1024 + (size_t)nn];
                }
            } else {
                const long long n = 4LL * m0 - 144 + t;
                v0 = virt_pcm(g, b, s, 0, n, started);
                if (g.C == 2) v1 = virt_pcm(g, b, s, 1, n, started);
            }
            xs[0][qphys(t)] = fmul(v0, 0.25f);                   // data / 4.0 (atrac3denc.cpp:704)
            xs[1][qphys(t)] = fmul(v1, 0.25f);
        }
    }
    __syncthreads();
    if (ch < g.C) {
        float lo[kQR], hi[kQR];
        qmf_task(xs[ch], cw, t7, one, lo, hi);
        float4* dl = reinterpret_cast<float4*>(&s1[ch][0][qphys(kQR * t7)]);
        float4* dh = reinterpret_cast<float4*>(&s1[ch][1][qphys(kQR * t7)]);
        dl[0] = make_float4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_float4(lo[4], lo[5], lo[6], lo[7]);
        dh[0] = make_float4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_float4(hi[4], hi[5], hi[6], hi[7]);
    }
    __syncthreads();
    // Qmf2(Buf1) -> subs[0], subs[1];  Qmf3(Buf2) -> subs[3], subs[2]   (atrac3_qmf.h:37-41)
    float* outb = &xs[0][0];                                     // [channel][band][kQT], the input tile is dead
    static_assert(2 * 4 * kQT <= 2 * kQXP, "output tile fits the input tile");
    if (ch < g.C && t7 < 2 * (kQT / kQR)) {
        const int which = t7 >= kQT / kQR, task = t7 - which * (kQT / kQR);
        float lo[kQR], hi[kQR];
        qmf_task(s1[ch][which], cw, task, one, lo, hi);
        float* ol = outb + (ch * 4 + (which ? 3 : 0)) * kQT + kQR * task;
        float* oh = outb + (ch * 4 + (which ? 2 : 1)) * kQT + kQR * task;
#pragma unroll
        for (int r = 0; r < kQR; r++) { ol[r] = lo[r]; oh[r] = hi[r]; }
    }
    __syncthreads();
    const int nvalid = min(kQT, g.BL - u0);
    ATDE_PAR_FOR(w, 4 * kQT) {
        const int band = w / kQT, q = w - band * kQT;
        if (q < nvalid) {
            if (g.js) {
                const float l = outb[band * kQT + q], r = outb[(4 + band) * kQT + q];
                b.bands[(((size_t)s * 2 + 0) * 4 + band) * g.BL + u0 + q] = fmul(fadd(l, r), 0.5f);
                b.bands[(((size_t)s * 2 + 1) * 4 + band) * g.BL + u0 + q] = fmul(fsub(l, r), 0.5f);
            } else {
                for (int c = 0; c < g.C; c++)
                    b.bands[(((size_t)s * g.C + c) * 4 + band) * g.BL + u0 + q] = outb[(c * 4 + band) * kQT + q];
            }
        }
    }
}

void launch_qmf(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    dim3 grid((g.BL + kQT - 1) / kQT, g.S);
    ATDE_LAUNCH(at3_qmf_kernel, grid, 256, 0, st, g, b);
}

// =====================================================================================
// K2: spectral upsampler + sub-frame envelope of one (stream, channel, band, frame)
// =====================================================================================
// RelationToIdx (transient_detector.cpp:139-147 and atrac3denc.h:44-52: same function of x)
ATDE_D int relation_to_idx(float x)
{
    if (x <= 0.5f) {
        x = __fdiv_rn(1.0f, fmaxf(x, 0.00048828125f));
        const unsigned v = (unsigned)__float2int_rz(x);
        return 4 + (v ? 31 - __clz((int)v) : 0);
    }
    x = fminf(x, 16.0f);
    const unsigned v = (unsigned)__float2int_rz(x);
    return 4 - (v ? 31 - __clz((int)v) : 0);
}

ATDE_D float median3(float a, float b, float c)
{
    return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}

// MedianFilter<1> (transient_detector.cpp:149-163): edges use a 2-element window and take w[1] = max.
ATDE_D void median_filter1(const float* in, float* out, int n)
{
    for (int i = 0; i < n; i++) {
        if (i == 0) out[i] = fmaxf(in[0], in[1]);
        else if (i == n - 1) out[i] = fmaxf(in[n - 2], in[n - 1]);
        else out[i] = median3(in[i - 1], in[i], in[i + 1]);
    }
}

// FindPlateau (transient_detector.cpp:175-262) + target selection (:286-299); in[] has 32 entries.
ATDE_D float plateau_target(const float* in)
{
    const int n = 32, minc = 3;
    float max_raw = 0.0f;
    for (int i = 0; i < n; i++) max_raw = fmaxf(max_raw, in[i]);
    float filt[32];
    median_filter1(in, filt, n);
    float best = 0.0f;
    int best_end = -1;
    for (int j = 0; j + minc <= n; j++) {
        float mv = filt[j];
        for (int k = 1; k < minc; k++) mv = fminf(mv, filt[j + k]);
        if (mv > best) { best = mv; best_end = j + minc - 1; }
    }
    float level = best;
    bool release = false;
    if (best < 1e-6f) {
        level = 0.0f;
    } else {
        while (best_end + 1 < n && filt[best_end + 1] >= best) ++best_end;
        if (best_end < n - 1) {
            if (in[n - 1] < fmul(best, 0.1f)) {
                release = true;
            } else {
                bool any_high = false;
                for (int i = best_end + 1; i < n; i++)
                    if (in[i] >= fmul(best, 0.7f)) { any_high = true; break; }
                release = !any_high && (in[n - 1] < fmul(best, 0.5f));
            }
        }
    }
    const bool use_plateau = level > 1e-6f && !release && level >= fmul(max_raw, 0.4f);
    return use_plateau ? level : in[n - 1];
}

constexpr int kGainThreads = 128;         // FFT threads
constexpr int kGainBlock = kGainThreads + 32;   // + one helper warp for the double-precision hfr sums

// The 2048-point buffer is padded by one element per 8 (pass 1 stores 8 consecutive slots per thread:
// a 9-element thread stride is conflict-free) and by 8 more per 128 (so the eight-lane groups of
// pass 2, 128 slots apart, land on the other half of the banks).
ATDE_D int gphys(int i) { return i + (i >> 3) + ((i >> 7) << 3); }
// The real output signal is padded by 4 floats per 64 so that the 32 sequential 64-sample RMS sums
// (one lane each, 64 floats apart) read different banks.
ATDE_D int sphys(int j) { return j + ((j >> 6) << 2); }
// Spectrum-sized arrays (forward FFT buffer, frequency bins, staged super-twiddles) are padded by one
// element per 16: both the natural-order accesses and the base-4 digit-reversed ones (lane stride 64,
// 16, 4 elements) then fall on 16 different bank pairs.
ATDE_D int fq(int k) { return k + (k >> 4); }

// One radix-4 butterfly of the forward FFT-256 on the padded buffer; F = first element, m = sub-length.
ATDE_D void fwd_bfly(cpx* buf, int F, int m, cpx t1, cpx t2, cpx t3)
{
    cpx f0 = buf[fq(F)], f1 = buf[fq(F + m)], f2 = buf[fq(F + 2 * m)], f3 = buf[fq(F + 3 * m)];
    kf_bfly4<false>(f0, f1, f2, f3, t1, t2, t3);
    buf[fq(F)] = f0; buf[fq(F + m)] = f1; buf[fq(F + 2 * m)] = f2; buf[fq(F + 3 * m)] = f3;
}

// One (stream, channel, band, frame) per block.
//
// kiss_fftri(4096) = pre-processing + inverse complex FFT-2048 (4x4x4x4x4x2).  kissfft's decimation
// in time is restated as a digit-reversed gather followed by the stages innermost first; the stages
// are executed two at a time on registers (every butterfly keeps the library's exact operation
// order, so regrouping them changes nothing numerically):
//   pass 1  radix-2 (m=1) + radix-4 (m=2)   on the 8 consecutive slots of a group
//   pass 2  radix-4 (m=8) + radix-4 (m=32)  on 16 slots  base + k + 8a + 32b
//   pass 3  radix-4 (m=128) + radix-4 (m=512) on 16 slots k + 128a + 512b
// Only input bins LowCutBin..256 and their mirrors are non-zero, which leaves two (for one group,
// three) non-zero inputs per pass-1 group: slot 0 = tmpbuf[low], slot 7 = tmpbuf[1792 + low],
// slot 2 = tmpbuf[256] (low == 0 only).  Pass 1 is written out for that case; adding or multiplying
// by the structural zeros is exact, so the values equal the full transform's (up to the sign of zero,
// which no consumer can see: the output is squared).
// Only output samples [1024, 3072) are consumed (AnalyzeGain), i.e. complex slots [512, 1536).
__global__ void __launch_bounds__(kGainBlock, 6) at3_gain_kernel(Geometry g, Buffers b)
{
    __shared__ __align__(16) cpx big[2048 + 256 + 128];    // inverse FFT buffer, padded; later the real output
    __shared__ __align__(8) cpx tw2c[15][8];         // pass-2 twiddles of lane group k, compact
    __shared__ __align__(16) cpx fwd[256 + 16];
    __shared__ __align__(16) cpx freq[257 + 17];
    __shared__ __align__(16) cpx sup[257 + 17];      // super4096[k - 1] at fq(k), k = 1..256
    __shared__ float micro[256];
    __shared__ float sgain[96];
    __shared__ __align__(16) double2 ee[257];        // (|X_k|^2, |X_k H_k|^2)
    __shared__ double esum2[2];
    __shared__ float sstat[2];

    const DevTables* __restrict__ T = b.tab;
    const int f = blockIdx.x;
    const int band = blockIdx.y % kGainBands, c = blockIdx.y / kGainBands;
    const int s = blockIdx.z;
    const float* __restrict__ in = b.bands + (((size_t)s * g.C + c) * 4 + band) * g.BL + 256 * (size_t)f;
    const int tid = threadIdx.x;
    const cpx* __restrict__ tw = T->tw2048;

    if (tid < 120) (&tw2c[0][0])[tid] = (&T->gtw2[0][0])[tid];
    // the only super-twiddles kiss_fftri(4096) meets with a non-zero operand
    for (int k = 1 + tid; k <= 256; k += kGainBlock) sup[fq(k)] = T->super4096[k - 1];
    // 1. Planck window, packed as the complex input of the half-size FFT; loaded in natural order and
    //    stored at its digit-reversed slot (base-4 reversal of 4 digits is an involution)
    ATDE_PAR_FOR(j, 256) {
        const float2 x = *reinterpret_cast<const float2*>(in + 2 * j);
        const float2 w = *reinterpret_cast<const float2*>(&T->planck[2 * j]);
        cpx z;
        z.r = fmul(x.x, w.x);
        z.i = fmul(x.y, w.y);
        const int o = ((j & 3) << 6) | (((j >> 2) & 3) << 4) | (((j >> 4) & 3) << 2) | (j >> 6);
        fwd[fq(o)] = z;
    }
    __syncthreads();
    // 2. forward complex FFT-256 = 4x4x4x4, innermost stage first; the lane -> butterfly map of every
    //    stage is chosen so that 16 consecutive lanes touch 16 different bank pairs
    if (tid < 64) fwd_bfly(fwd, 4 * tid, 1, T->ftw[0][0][0], T->ftw[0][1][0], T->ftw[0][2][0]);
    __syncthreads();
    if (tid < 64) {
        const int gq = tid & 15, k = tid >> 4;
        fwd_bfly(fwd, 16 * gq + k, 4, T->ftw[1][0][k], T->ftw[1][1][k], T->ftw[1][2][k]);
    }
    __syncthreads();
    if (tid < 64) {
        const int gq = tid >> 4, k = tid & 15;
        fwd_bfly(fwd, 64 * gq + k, 16, T->ftw[2][0][k], T->ftw[2][1][k], T->ftw[2][2][k]);
    }
    __syncthreads();
    if (tid < 64) fwd_bfly(fwd, tid, 64, T->ftw[3][0][tid], T->ftw[3][1][tid], T->ftw[3][2][tid]);
    __syncthreads();
    // kiss_fftr post-processing (kiss_fftr.c:84-115)
    ATDE_PAR_FOR(k, 129) {
        if (k == 0) {
            const float tr = fwd[0].r, ti = fwd[0].i;
            freq[0].r = fadd(tr, ti);   freq[0].i = 0.0f;
            freq[fq(256)].r = fsub(tr, ti); freq[fq(256)].i = 0.0f;
        } else {
            const cpx fpk = fwd[fq(k)];
            cpx fpnk; fpnk.r = fwd[fq(256 - k)].r; fpnk.i = -fwd[fq(256 - k)].i;
            cpx f1k, f2k;
            f1k.r = fadd(fpk.r, fpnk.r); f1k.i = fadd(fpk.i, fpnk.i);
            f2k.r = fsub(fpk.r, fpnk.r); f2k.i = fsub(fpk.i, fpnk.i);
            const cpx t2 = cmul(f2k, T->super512[k - 1]);
            cpx a, bb;
            a.r = fmul(fadd(f1k.r, t2.r), 0.5f);  a.i = fmul(fadd(f1k.i, t2.i), 0.5f);
            bb.r = fmul(fsub(f1k.r, t2.r), 0.5f); bb.i = fmul(fsub(t2.i, f1k.i), 0.5f);
            freq[fq(k)] = a;                   // k == 128 writes the same element twice: the second
            freq[fq(256 - k)] = bb;            // store (freqdata[ncfft-k]) wins, as in the reference
        }
    }
    __syncthreads();
    // 2a. high-frequency energy ratio (upsampler.cpp:99-118): two SEQUENTIAL double sums over the 257
    //     bins.  The helper warp takes them (per-bin terms lane-parallel, then both chains interleaved
    //     on one lane) while the four FFT warps run the inverse transform; they only meet again at the
    //     final barrier.
    const int lcb = T->low_cut_bin;
    if (tid >= kGainThreads) {
        const int hl = tid - kGainThreads;
        for (int k = hl; k < 257; k += 32) {
            const double r = (double)freq[fq(k)].r, i = (double)freq[fq(k)].i;
            const double e = __dadd_rn(__dmul_rn(r, r), __dmul_rn(i, i));
            float H = 0.0f;
            if (k >= lcb + 2) H = 1.0f;
            else if (k >= lcb) H = T->hpf_h[k - lcb];
            ee[k] = make_double2(e, __dmul_rn(__dmul_rn(e, (double)H), (double)H));
        }
        __syncwarp();
        if (hl == 0) {
            double tot = 0.0, hi = 0.0;
#pragma unroll 4
            for (int k = 0; k <= 256; k++) { const double2 v = ee[k]; tot = __dadd_rn(tot, v.x); hi = __dadd_rn(hi, v.y); }
            esum2[0] = tot; esum2[1] = hi;
        }
        __syncthreads();                                      // the final barrier of the FFT warps
        return;
    }
    // 3/4. inverse FFT input Y[k] = 8*X[k]*H[k] (Nyquist bin halved), kiss_fftri pre-processing
    //      (kiss_fftr.c:131-151) with Y[2048-k] == 0, and pass 1 of the inverse FFT.
    auto tmp_pair = [&](int k, cpx& lo, cpx& hi) {
        // tmpbuf[k] and tmpbuf[2048-k] for 1 <= k <= 256; zero when the bin is cut
        lo.r = lo.i = hi.r = hi.i = 0.0f;
        if (k < lcb || k < 1) return;
        cpx fk;
        if (k == 256) {
            if (lcb + 2 > 256) return;
            fk.r = fmul(fmul(freq[fq(256)].r, 8.0f), 0.5f);
            fk.i = 0.0f;
        } else if (k >= lcb + 2) {
            fk.r = fmul(freq[fq(k)].r, 8.0f);
            fk.i = fmul(freq[fq(k)].i, 8.0f);
        } else {
            const float w = T->hpf_h[k - lcb];
            fk.r = fmul(fmul(freq[fq(k)].r, 8.0f), w);
            fk.i = fmul(fmul(freq[fq(k)].i, 8.0f), w);
        }
        // fnkc = conj(Y[2048-k]) = (0, -0):  fek = fk + fnkc, tmp = fk - fnkc
        cpx fek, tp;
        fek.r = fadd(fk.r, 0.0f);  fek.i = fadd(fk.i, -0.0f);
        tp.r = fsub(fk.r, 0.0f);   tp.i = fsub(fk.i, -0.0f);
        const cpx fok = cmul(tp, sup[fq(k)]);
        lo.r = fadd(fek.r, fok.r);  lo.i = fadd(fek.i, fok.i);
        hi.r = fsub(fek.r, fok.r);  hi.i = fmul(fsub(fek.i, fok.i), -1.0f);
    };
    for (int grp = tid; grp < 256; grp += kGainThreads) {
        const int low = ((grp >> 6) & 3) | (((grp >> 4) & 3) << 2) | (((grp >> 2) & 3) << 4) | ((grp & 3) << 6);
        cpx x0, x2, x7, dummy;
        tmp_pair(low, x0, dummy);                        // slot 0: input index low (tmpbuf[0] is zero: bin 0 is cut)
        tmp_pair(256 - low, dummy, x7);                  // slot 7: input index 1792 + low = 2048 - (256 - low)
        x2.r = x2.i = 0.0f;
        if (low == 0) { x0.r = x0.i = 0.0f; tmp_pair(256, x2, dummy); }   // slot 2: input index 256
        if (lcb == 0 && low == 0) {
            // bin 0 kept (not the encoder's configuration, kept for completeness): tmpbuf[0]
            const float y0 = fmul(freq[0].r, 8.0f);
            x0.r = fadd(y0, 0.0f); x0.i = fsub(y0, 0.0f);
        }
        // radix-2, m = 1, twiddle tw[0]: (F0,F1) = (x0, x0); (F2,F3) = (x2, x2); (F6,F7) = (x7, -x7)
        const cpx t7 = cmul(x7, tw[0]);
        cpx F6, F7;
        F7.r = fsub(0.0f, t7.r); F7.i = fsub(0.0f, t7.i);
        F6.r = fadd(0.0f, t7.r); F6.i = fadd(0.0f, t7.i);
        cpx o0, o1, o2, o3, o4, o5, o6, o7;
        {   // radix-4, m = 2, k = 0: elements (x0, x2, 0, F6), twiddles tw[0]
            cpx f0 = x0, f1 = x2, f2, f3 = F6;
            f2.r = f2.i = 0.0f;
            kf_bfly4<true>(f0, f1, f2, f3, tw[0], tw[0], tw[0]);
            o0 = f0; o2 = f1; o4 = f2; o6 = f3;
        }
        {   // k = 1: elements (x0, x2, 0, F7), twiddles tw[256], tw[512], tw[768]
            cpx f0 = x0, f1 = x2, f2, f3 = F7;
            f2.r = f2.i = 0.0f;
            kf_bfly4<true>(f0, f1, f2, f3, tw[256], tw[512], tw[768]);
            o1 = f0; o3 = f1; o5 = f2; o7 = f3;
        }
        cpx* dst = big + gphys(8 * grp);
        dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3; dst[4] = o4; dst[5] = o5; dst[6] = o6; dst[7] = o7;
    }
    atde_named_barrier(1, kGainThreads);
    // pass 2: radix-4 m = 8 (fstride 64), then m = 32 (fstride 16)
    {
        const int k = tid & 7, base = (tid >> 3) << 7;
        cpx x[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) x[a][q] = big[gphys(base + k + 8 * a + 32 * q)];
        {
            const cpx t1 = tw2c[0][k], t2 = tw2c[1][k], t3 = tw2c[2][k];      // tw[64k], tw[128k], tw[192k]
#pragma unroll
            for (int q = 0; q < 4; q++) kf_bfly4<true>(x[0][q], x[1][q], x[2][q], x[3][q], t1, t2, t3);
        }
#pragma unroll
        for (int a = 0; a < 4; a++)                                            // kk = k + 8a: tw[16kk], tw[32kk], tw[48kk]
            kf_bfly4<true>(x[a][0], x[a][1], x[a][2], x[a][3], tw2c[3 + 3 * a][k], tw2c[4 + 3 * a][k], tw2c[5 + 3 * a][k]);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) big[gphys(base + k + 8 * a + 32 * q)] = x[a][q];
    }
    atde_named_barrier(1, kGainThreads);
    // pass 3: radix-4 m = 128 (fstride 4), then m = 512 (fstride 1); keep slots [512, 1536), normalised
    {
        const int k = tid;
        cpx x[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) x[a][q] = big[gphys(k + 128 * a + 512 * q)];
        atde_named_barrier(1, kGainThreads);                                  // every slot is in registers: big can be overwritten
        {
            const cpx t1 = T->gtw3a[0][k], t2 = T->gtw3a[1][k], t3 = T->gtw3a[2][k];     // tw[4k], tw[8k], tw[12k]
#pragma unroll
            for (int q = 0; q < 4; q++) kf_bfly4<true>(x[0][q], x[1][q], x[2][q], x[3][q], t1, t2, t3);
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int kk = k + 128 * a;
            kf_bfly4<true>(x[a][0], x[a][1], x[a][2], x[a][3], T->gtw3b[a][0][k], T->gtw3b[a][1][k], T->gtw3b[a][2][k]);  // tw[kk], tw[2kk], tw[3kk]
            // 5. normalise (norm = 1/4096); complex slot kk + 512q -> output samples 2*slot, 2*slot+1
            cpx u, v;
            u.r = fmul(x[a][1].r, 1.0f / 4096.0f); u.i = fmul(x[a][1].i, 1.0f / 4096.0f);
            v.r = fmul(x[a][2].r, 1.0f / 4096.0f); v.i = fmul(x[a][2].i, 1.0f / 4096.0f);
            float* sigw = reinterpret_cast<float*>(big);
            *reinterpret_cast<cpx*>(sigw + sphys(2 * kk)) = u;            // slot 512 + kk  -> samples 1024 + 2kk, +1
            *reinterpret_cast<cpx*>(sigw + sphys(1024 + 2 * kk)) = v;     // slot 1024 + kk -> samples 2048 + 2kk, +1
        }
    }
    atde_named_barrier(1, kGainThreads);
    // AnalyzeGain(signal + 1024, 2048, 32, rms): 64-sample RMS, plus 8 micro-chunk RMS values each
    const float* sig = reinterpret_cast<const float*>(big);       // sig[sphys(i)] = output sample 1024 + i
    for (int l = tid; l < 256; l += kGainThreads) {
        const int q = (l & ~12) | ((l & 4) << 1) | ((l & 8) >> 1);    // bits 2 and 3 swapped: conflict-free float4 reads
        const float* p = sig + sphys(8 * q);
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) a = fadd(a, fmul(p[i], p[i]));
        micro[q] = __fsqrt_rn(__fdiv_rn(a, 8.0f));
    }
    if (tid >= 32 && tid < 64) {
        const int sf = tid - 32;
        const float* p = sig + sphys(64 * sf);
        float a = 0.0f;
        for (int i = 0; i < 64; i += 4) {
            const float4 q = *reinterpret_cast<const float4*>(p + i);
            a = fadd(a, fmul(q.x, q.x)); a = fadd(a, fmul(q.y, q.y));
            a = fadd(a, fmul(q.z, q.z)); a = fadd(a, fmul(q.w, q.w));
        }
        sgain[sf] = __fsqrt_rn(__fdiv_rn(a, 64.0f));
    }
    atde_named_barrier(1, kGainThreads);
    if (tid < 32) {
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = micro[8 * tid + i];
        // ascending sort of 8 plain floats (any correct sort gives std::sort's result)
#pragma unroll
        for (int i = 1; i < 8; i++) {
#pragma unroll
            for (int j = 7; j >= 1; j--) {
                if (j >= i) {
                    const float lo2 = fminf(m[j - 1], m[j]), hi3 = fmaxf(m[j - 1], m[j]);
                    m[j - 1] = lo2; m[j] = hi3;
                }
            }
        }
        sgain[32 + tid] = m[2];
        sgain[64 + tid] = m[6];
    }
    // the closing scalar work: curHpfEnergy and the plateau target on two different warps
    if (tid == 64) {
        float cur = 0.0f;
        for (int i = 0; i < 32; i++) cur = fadd(cur, sgain[i]);
        sstat[0] = __fdiv_rn(cur, 32.0f);
    } else if (tid == 96) {
        sstat[1] = plateau_target(sgain);
    }
    __syncthreads();
    const size_t item = ((((size_t)s * g.C + c) * kGainBands + band) * g.n_out + f);
    if (tid < 96) b.gain[item * 96 + tid] = sgain[tid];
    if (tid == 0) {
        float4 st4;
        st4.x = (esum2[0] > 0.0) ? __double2float_rn(__ddiv_rn(esum2[1], esum2[0])) : 0.0f;
        st4.y = sstat[0];
        st4.z = sstat[1];
        st4.w = sgain[31];
        reinterpret_cast<float4*>(b.gstat)[item] = st4;
    }
}

void launch_gain_analysis(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    dim3 grid(g.n_out, g.C * kGainBands, g.S);
    ATDE_LAUNCH(at3_gain_kernel, grid, kGainBlock, 0, st, g, b);
}

// =====================================================================================
// K3: TCurveBuilderCtx recurrence, one thread per (stream, channel, band)
// =====================================================================================
__global__ void at3_gain_scan_kernel(Geometry g, Buffers b)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (s*C + c)*3 + band
    if (idx >= g.S * g.C * kGainBands) return;
    float4* ctxp = reinterpret_cast<float4*>(b.ctx) + idx;
    float4 ctx = *ctxp;                                              // x LastLevel, y LastHpfEnergy, z LastTarget
    const float4* st = reinterpret_cast<const float4*>(b.gstat) + (size_t)idx * g.n_out;
    float4* pv = reinterpret_cast<float4*>(b.gprev) + (size_t)idx * g.n_out;
    for (int f = 0; f < g.n_out; f++) {
        const float4 v = st[f];
        float4 o;
        o.x = ctx.y; o.y = ctx.x; o.z = ctx.z; o.w = 0.0f;
        if (v.x < 0.05f) {                                           // kHighFreqThreshold: LastLevel = 0, continue
            ctx.x = 0.0f;
        } else {
            ctx.y = v.y;
            ctx.x = v.w;
            ctx.z = v.z;
        }
        pv[f] = o;
    }
    *ctxp = ctx;
}

void launch_gain_scan(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const int n = g.S * g.C * kGainBands;
    ATDE_LAUNCH(at3_gain_scan_kernel, (n + 63) / 64, 64, 0, st, g, b);
}

// =====================================================================================
// K4: gain curve of one (stream, channel, band, frame), one thread each
// =====================================================================================
struct Pts {
    int n;
    int level[8];
    int loc[8];
};

// BuildSubframeDivisors (atrac3denc.cpp:228-255): mean divisor of each 8-sample sub-frame.  Point i
// holds its level up to sample 8*loc_i and ramps over the 8 samples of sub-frame loc_i, so a
// sub-frame is either constant (sum of 8 equal powers of two, exact) or one point's ramp.
ATDE_D void subframe_divisors(const DevTables* T, const Pts& p, float* out_div)
{
    int sf = 0;
    for (int i = 0; i < p.n; i++) {
        const float level0 = T->gain_level[p.level[i]];
        for (; sf < p.loc[i] && sf < 32; ++sf) {
            float sum = 0.0f;
            for (int k = 0; k < 8; k++) sum = fadd(sum, level0);
            out_div[sf] = __fdiv_rn(sum, 8.0f);
        }
        if (sf < 32 && sf == p.loc[i]) {
            const int inc = ((i + 1) < p.n ? p.level[i + 1] : 4) - p.level[i] + 15;
            const float ginc = T->gain_interp[inc];
            float level = level0, sum = 0.0f;
            for (int k = 0; k < 8; k++) { sum = fadd(sum, level); level = fmul(level, ginc); }
            out_div[sf] = __fdiv_rn(sum, 8.0f);
            ++sf;
        }
    }
    for (; sf < 32; ++sf) out_div[sf] = 1.0f;
}

// CalcCurveEarlyMismatchScore (atrac3denc.cpp:259-297)
ATDE_D float mismatch_score(const DevTables* T, const float* gain, float target, const Pts& p)
{
    if (target <= 1e-9f) return 0.0f;
    float div[32];
    subframe_divisors(T, p, div);
    int max_loc = 0;
    for (int i = 0; i < p.n; i++) max_loc = max(max_loc, p.loc[i]);
    const int eval = min(32, max(3, max_loc + 3));
    const float eps = 1e-9f;
    float fit = 0.0f;
    for (int sf = 0; sf < eval; sf++) {
        const float mod = __fdiv_rn(gain[sf], fmaxf(div[sf], eps));
        const float e = g_log2f(__fdiv_rn(fmaxf(mod, eps), fmaxf(target, eps)));
        fit = fadd(fit, fmul(e, e));
    }
    fit = __fdiv_rn(fit, (float)eval);
    float leak = 0.0f, wsum = 0.0f;
    for (int sf = 0; sf + 1 < eval; sf++) {
        const float a = g_log2f(fmaxf(div[sf], eps));
        const float bb = g_log2f(fmaxf(div[sf + 1], eps));
        const float d = fsub(bb, a);
        const float w = fmul(0.5f, fadd(gain[sf], gain[sf + 1]));
        leak = fadd(leak, fmul(fmul(d, d), w));
        wsum = fadd(wsum, w);
    }
    if (wsum > eps) leak = __fdiv_rn(leak, wsum);
    return fadd(fit, fmul(0.25f, leak));
}

// CalcCurve (transient_detector.cpp:276-482) after the ctx update; returns the points in p
ATDE_D void calc_curve(const float* in, const float* low, const float* high, float target,
                       float saved_last_level, float saved_last_target, float min_score, Pts& p)
{
    p.n = 0;
    const int n = 32;
    if (target < 1e-6f) return;
    if (saved_last_level < 1e-6f) return;
    float filt[32];
    median_filter1(in, filt, n);
    float max_gain = 0.0f;
    for (int i = 0; i < n; i++) max_gain = fmaxf(max_gain, in[i]);
    const float intra = __fdiv_rn(max_gain, fmaxf(target, 1e-9f));
    float inter = 1.0f;
    if (saved_last_target > 1e-6f) {
        const float hi = fmaxf(saved_last_target, target), lo = fminf(saved_last_target, target);
        inter = __fdiv_rn(hi, fmaxf(lo, 1e-9f));
    }
    const bool sticky = intra <= 7.0f && inter <= 10.0f;
    int lev[32];
    for (int i = 0; i < n; i++) {
        int level = relation_to_idx(__fdiv_rn(filt[i], target));
        if (i > 0 && sticky) {
            float rlo = __fdiv_rn(low[i], target), rhi = __fdiv_rn(high[i], target);
            if (rlo > rhi) { const float t = rlo; rlo = rhi; rhi = t; }
            const int ilo = relation_to_idx(rlo), ihi = relation_to_idx(rhi);
            const int mn = min(ilo, ihi), mx = max(ilo, ihi);
            const int prev = lev[i - 1];
            if (mx - mn <= 1 && abs(level - prev) == 1 && prev >= mn && prev <= mx) level = prev;
        }
        lev[i] = level;
    }
    int target_sf = 0;
    for (int sf = n - 2; sf >= 0; --sf)
        if (lev[sf] != 4) { target_sf = sf + 1; break; }
    if (target_sf == 0) return;
    // transitions, scanned leftward from target_sf (stored right-to-left, reversed below)
    int t_loc[32], t_lev[32], t_delta[32], nt = 0;
    {
        int prev = 4;
        for (int sf = target_sf - 1; sf >= 0; --sf) {
            const int l = lev[sf];
            if (l != prev) {
                const int loc = sf + 1;
                const int delta = abs(l - prev);
                bool keep = (loc == target_sf) || (delta >= 2);
                if (!keep) {
                    // BoundaryTransientScore(filtered, loc, 3) (transient_detector.cpp:251-274)
                    float lm = 0.0f, rm = 0.0f;
                    for (int i = max(0, loc - 3); i < loc; i++) lm = fmaxf(lm, filt[i]);
                    for (int i = loc; i < min(n, loc + 3); i++) rm = fmaxf(rm, filt[i]);
                    const float eps = 1e-9f;
                    const float attack = __fdiv_rn(fadd(rm, eps), fadd(lm, eps));
                    const float release = __fdiv_rn(fadd(lm, eps), fadd(rm, eps));
                    keep = fmaxf(attack, release) >= min_score;
                }
                if (keep) { t_loc[nt] = loc; t_lev[nt] = l; t_delta[nt] = delta; nt++; prev = l; }
            }
        }
    }
    if (nt == 0) return;
    // after std::reverse the transitions are in ascending Loc order: index nt-1-k
    if (nt > 6) {
        // keep the 6 largest by (Delta desc, Loc desc) — a strict total order since Locs are unique —
        // then emit in ascending Loc.
        bool keepf[32];
        for (int i = 0; i < nt; i++) keepf[i] = false;
        for (int r = 0; r < 6; r++) {
            int best = -1;
            for (int i = 0; i < nt; i++) {
                if (keepf[i]) continue;
                if (best < 0 || t_delta[i] > t_delta[best] || (t_delta[i] == t_delta[best] && t_loc[i] > t_loc[best]))
                    best = i;
            }
            keepf[best] = true;
        }
        for (int i = nt - 1; i >= 0; --i)
            if (keepf[i]) { p.level[p.n] = t_lev[i]; p.loc[p.n] = t_loc[i]; p.n++; }
    } else {
        for (int i = nt - 1; i >= 0; --i) { p.level[p.n] = t_lev[i]; p.loc[p.n] = t_loc[i]; p.n++; }
    }
}

__global__ void __launch_bounds__(128) at3_curve_kernel(Geometry g, Buffers b)
{
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // ((s*C+c)*3+band)*n_out + f
    const long long total = (long long)g.S * g.C * kGainBands * g.n_out;
    if (item >= total) return;
    const DevTables* __restrict__ T = b.tab;
    const int f = (int)(item % g.n_out);
    const long long scb = item / g.n_out;
    const int band = (int)(scb % kGainBands);
    const long long sc = scb / kGainBands;
    Curve out;
    out.n = 0; out.pad = 0;
    for (int i = 0; i < 7; i++) { out.level[i] = 0; out.loc[i] = 0; }
    Curve* dst = b.curves + ((size_t)sc * 4 + band) * g.n_out + f;

    const float4 st = reinterpret_cast<const float4*>(b.gstat)[item];
    const float4 pv = reinterpret_cast<const float4*>(b.gprev)[item];
    const float hfr = st.x, cur_hpf = st.y, target = st.z;
    const float prev_hpf = pv.x, saved_ll = pv.y, saved_lt = pv.z;
    if (hfr < 0.05f) { *dst = out; return; }

    float gain[32], low[32], high[32];
    const float* gp = b.gain + (size_t)item * 96;
    for (int i = 0; i < 32; i++) { gain[i] = gp[i]; low[i] = gp[32 + i]; high[i] = gp[64 + i]; }

    const float ratio = (cur_hpf > 1e-9f && prev_hpf > 1e-9f) ? __fdiv_rn(prev_hpf, cur_hpf) : 1.0f;
    const float min_score = fmul(1.9f, fminf(1.5f, fmaxf(1.0f, ratio)));
    const float prev_target = saved_lt;
    Pts p;
    calc_curve(gain, low, high, target, saved_ll, saved_lt, min_score, p);
    if (p.n == 0) { *dst = out; return; }                     // "skip: no_curve" (atrac3denc.cpp:395-400)

    float max_gain = 0.0f;
    for (int i = 0; i < 32; i++) max_gain = fmaxf(max_gain, gain[i]);
    if (max_gain < 1e-4f) p.n = 0;
    if (hfr < 0.3f) p.n = 0;

    // explicit point 0 (atrac3denc.cpp:457-554)
    const Pts before = p;
    bool changed = false;
    float next_mod = 0.0f;
    bool valid = false;
    if (p.n > 0 && p.loc[0] > 0) {
        const int nb = p.loc[0];
        float sum = 0.0f;
        for (int sf = 0; sf < nb; sf++) sum = fadd(sum, gain[sf]);
        next_mod = __fdiv_rn(__fdiv_rn(sum, (float)nb), T->gain_level[p.level[0]]);
        valid = true;
    } else if (p.n == 0) {
        float sum = 0.0f;
        for (int i = 0; i < 32; i++) sum = fadd(sum, gain[i]);
        next_mod = __fdiv_rn(sum, 32.0f);
        valid = true;
    }
    const bool have0 = valid && prev_target > 1e-6f && next_mod > 1e-6f;
    if (have0) {
        const int l0 = relation_to_idx(__fdiv_rn(prev_target, next_mod));
        int at = -1;
        for (int i = 0; i < p.n; i++) if (p.loc[i] == 0) { at = i; break; }
        if (at >= 0) {
            if (p.level[at] != l0) { p.level[at] = l0; changed = true; }
        } else if (l0 != 4 || p.n > 0) {
            for (int i = p.n; i > 0; --i) { p.level[i] = p.level[i - 1]; p.loc[i] = p.loc[i - 1]; }
            p.level[0] = l0; p.loc[0] = 0; p.n++;
            changed = true;
        }
    }
    if (changed) {
        const float score_before = mismatch_score(T, gain, target, before);
        const float score_after = mismatch_score(T, gain, target, p);
        bool keep_by_boundary = false;
        if (have0) {
            const float desired = fminf(fmaxf(__fdiv_rn(prev_target, next_mod), T->gain_level[15]), T->gain_level[0]);
            const float sb = T->gain_level[before.n ? before.level[0] : 4];
            const float sa = T->gain_level[p.n ? p.level[0] : 4];
            const float eps = 1e-9f;
            const float eb = fabsf(g_log2f(__fdiv_rn(fmaxf(sb, eps), fmaxf(desired, eps))));
            const float ea = fabsf(g_log2f(__fdiv_rn(fmaxf(sa, eps), fmaxf(desired, eps))));
            keep_by_boundary = fadd(ea, 0.20f) < eb;
        }
        if (!keep_by_boundary && score_after > fmul(score_before, fadd(1.0f, 0.02f)))
            p = before;
    }
    if (p.n >= 2 && p.loc[0] == 0 && p.level[0] == p.level[1]) {
        for (int i = 0; i + 1 < p.n; i++) { p.level[i] = p.level[i + 1]; p.loc[i] = p.loc[i + 1]; }
        p.n--;
    }
    out.n = (unsigned char)p.n;
    for (int i = 0; i < p.n && i < 7; i++) { out.level[i] = (unsigned char)p.level[i]; out.loc[i] = (unsigned char)p.loc[i]; }
    *dst = out;
}

void launch_gain_curve(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const long long total = (long long)g.S * g.C * kGainBands * g.n_out;
    ATDE_LAUNCH(at3_curve_kernel, (unsigned)((total + 127) / 128), 128, 0, st, g, b);
}

// =====================================================================================
// K5: gain modulation + window + MDCT-512 x4 + energy scales + loudness term, one block per
//     (stream, frame, channel)
// =====================================================================================
// Divisor applied by TGainProcessor::Modulate / BuildSampleDivisors to sample `pos` of the frame the
// curve belongs to (gain_processor.h:87-121, atrac3denc.cpp:154-173); 1.0 beyond the last ramp.
ATDE_D float curve_level(const DevTables* T, const Curve& cv, int pos)
{
    for (int i = 0; i < cv.n; i++) {
        const int last = (int)cv.loc[i] << 3;
        if (pos < last + 8) {
            float level = T->gain_level[cv.level[i]];
            if (pos >= last) {
                const int inc = ((i + 1) < cv.n ? (int)cv.level[i + 1] : 4) - (int)cv.level[i] + 15;
                const float ginc = T->gain_interp[inc];
                for (int r = last; r < pos; r++) level = fmul(level, ginc);
            }
            return level;
        }
    }
    return 1.0f;
}


// The same divisors for all 256 positions of a band at once, by a whole warp: lane g owns positions 8g .. 8g+7,
// which lie in one 8-sample step of the curve.  curve_level() picks the first point i with pos < 8 loc[i] + 8,
// i.e. g <= loc[i]; left of the ramp (g < loc[i]) the level is constant, inside it (g == loc[i]) it is
// level * ginc^r built by r successive multiplications, exactly the chain of the per-sample loop.
ATDE_D void curve_levels_warp(const DevTables* T, const Curve& cv, int lane, float* dst)
{
    float v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = 1.0f;
    for (int i = 0; i < cv.n; i++) {
        const int loc = cv.loc[i];
        if (lane <= loc) {
            float level = T->gain_level[cv.level[i]];
            float ginc = 1.0f;
            if (lane == loc) {
                const int inc = ((i + 1) < cv.n ? (int)cv.level[i + 1] : 4) - (int)cv.level[i] + 15;
                ginc = T->gain_interp[inc];
            }
#pragma unroll
            for (int r = 0; r < 8; r++) {
                v[r] = level;
                if (lane == loc) level = fmul(level, ginc);
            }
            break;
        }
    }
    *reinterpret_cast<float4*>(dst + 8 * lane) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 8 * lane + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// SafeEnergyScale (atrac3denc.cpp:143-152)
ATDE_D float safe_energy_scale(float orig, float mod)
{
    const float eps = 1.0e-20f;
    const float inf = __int_as_float(0x7f800000);
    if (orig <= eps || mod <= eps || !(fabsf(orig) < inf) || !(fabsf(mod) < inf)) return 1.0f;
    const float sc = __fdiv_rn(orig, mod);
    return (fabsf(sc) < inf && sc > 0.0f) ? sc : 1.0f;
}

// One WARP per (stream, frame, channel), no block-wide barrier.  The unit's four MDCT-512 run as two
// iterations of two bands, 16 lanes per band:
//   phase A  load the band samples of this and the previous frame (coalesced float4), divide by the
//            gain curves, window, store the two 512-sample MDCT inputs to the warp's tile
//   phase B  fold + pre-twiddle (mdct.h:56-76) straight into kissfft's gather order; lane k16 of a band
//            owns gather block k16 (8 consecutive slots) and runs the two innermost stages
//            (radix-2 m=1, radix-4 m=2) on registers
//   exchange through the tile: lane (k, bp) takes elements k + 8a + 32b, a = 0..3, b = 2bp, 2bp+1
//   phase C  radix-4 m=8 over a; lane pairs (k16, k16^8) swap halves by shuffle so that lane (k, ap)
//            holds k + 8a + 32b, a = 2ap, 2ap+1, b = 0..3; radix-4 m=32 over b; post-twiddle
//            (mdct.h:92-101); spectrum staged in the tile and written out as coalesced float4
// Every butterfly keeps kissfft's operation order (kissfft_dev.cuh), so the regrouping is exact.
// The loop body is kept small on purpose (two bands per iteration): the fully unrolled four-band version missed
// the instruction cache on every warp.  Gain-curve divisors are built once per band by the whole warp
// (curve_levels_warp) instead of per sample: the per-sample curve walk was 28 % of this kernel's instructions
// and 45 % of its stall samples on the bench signal (ncu v6).
constexpr int kMdctWarps = 4;
constexpr int kMdctBandStride = 520;                 // floats; keeps float4 alignment, shifts banks by 8
constexpr int kMdctXchStride = 152;                  // cpx per band in the exchange layout (16 blocks x 9, +8)
constexpr int kMdctOutStride = 264;
constexpr int kMdctTermStride = 132;                 // floats between the seven energy-term rows: the seven summing lanes hit 28 different banks
constexpr int kMdctLevels = 2 * kMdctBandStride;     // floats: divisor tables of the band in work (its own curve, the previous frame's)
constexpr int kMdctTile = kMdctLevels + 512;         // floats per warp: two MDCT inputs (or 7 x 128 energy terms) + the two tables

__global__ void __launch_bounds__(kMdctWarps * 32, 7) at3_mdct_kernel(Geometry g, Buffers b)
{
    __shared__ __align__(16) float tile[kMdctWarps][kMdctTile];
    __shared__ __align__(16) float s_sincos[256];
    __shared__ __align__(16) float s_win[256];
    __shared__ __align__(16) cpx s_tw[128];
    __shared__ Curve s_cv[kMdctWarps][4][2];         // [band][0] previous frame, [1] this frame
    __shared__ float s_esum[kMdctWarps][4][8];

    const DevTables* __restrict__ T = b.tab;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ATDE_PAR_FOR(i, 256) { s_sincos[i] = T->sincos512[i]; s_win[i] = T->encode_window[i]; }
    ATDE_PAR_FOR(i, 128) s_tw[i] = T->tw128[i];
    __syncthreads();

    float* tl = tile[warp];
    const long long n_units = (long long)g.S * g.n_out * g.C;
    for (long long unit = (long long)blockIdx.x * kMdctWarps + warp; unit < n_units;
         unit += (long long)gridDim.x * kMdctWarps) {
        const int c = (int)(unit % g.C);
        const long long sf = unit / g.C;
        const int f = (int)(sf % g.n_out), s = (int)(sf / g.n_out);
        const size_t sc = (size_t)s * g.C + c;

        __syncwarp();                                    // previous unit's tile reads are done
        if (lane < 8) {
            const int band = lane >> 1, which = lane & 1;
            Curve cv;
            cv.n = 0;
            const int ff = f - 1 + which;
            if (!g.no_gain && band < kGainBands && ff >= 0)
                cv = b.curves[(sc * 4 + band) * g.n_out + ff];
            s_cv[warp][band][which] = cv;
        }
        __syncwarp();
        // A band whose own curve is empty and whose overlap scale is 1 has every energy scale exactly 1.0
        // (SafeEnergyScale(x, x)); the sequential energy sums only run for bands that touch a curve.
        unsigned trivial = 0;                              // bit per band (a mask, so that the band loop below stays rolled)
#pragma unroll
        for (int band = 0; band < 4; band++)
            if (s_cv[warp][band][1].n == 0 &&
                (f == 0 ? b.next_scale[sc * 4 + band] == 1.0f : s_cv[warp][band][0].n == 0)) trivial |= 1u << band;
        // ---- CalcGainEnergyScale's sequential sums (atrac3denc.cpp:189-216), bands that touch a curve only:
        // squared terms sample-parallel into the tile, seven lanes then add them up in order.
        //  0 prevStored  1 curOriginal  2 curModulated  3 nextOriginal  4 nextModulated
        //  5/6 nextOriginal/nextModulated of the PREVIOUS frame (-> its NextOverlapScale)
#pragma unroll 1
        for (int band = 0; band < 4; band++) {
            if ((trivial >> band) & 1u) continue;
            const Curve& cc = s_cv[warp][band][1];
            const Curve& pc = s_cv[warp][band][0];
            const float* bp = b.bands + (sc * 4 + band) * g.BL + 128 + 256 * (size_t)f;
            float* lv_c = tl + kMdctLevels;
            float* lv_p = lv_c + 256;
            __syncwarp();
            if (cc.n) curve_levels_warp(T, cc, lane, lv_c);
            if (f != 0 && pc.n) curve_levels_warp(T, pc, lane, lv_p);
            __syncwarp();
            float a = 0.0f;                                    // lane < 7: running sum of term `lane`
#pragma unroll 1
            for (int half = 0; half < 256; half += 128) {
#pragma unroll 1
                for (int j = lane; j < 128; j += 32) {
                    const int i = half + j;
                    const float x = bp[i];
                    const float xm = cc.n ? __fdiv_rn(x, lv_c[i]) : x;
                    const float wi = s_win[i], wr = s_win[255 - i];
                    float prev, y = 0.0f, ym = 0.0f;
                    if (f == 0) {
                        prev = b.prevhalf[(sc * 4 + band) * 256 + i];
                    } else {
                        y = bp[i - 256];
                        ym = pc.n ? __fdiv_rn(y, lv_p[i]) : y;
                        prev = fmul(wi, ym);
                    }
                    float v;
                    tl[0 * kMdctTermStride + j] = fmul(prev, prev);
                    v = fmul(x, wr);  tl[1 * kMdctTermStride + j] = fmul(v, v);
                    v = fmul(xm, wr); tl[2 * kMdctTermStride + j] = fmul(v, v);
                    v = fmul(x, wi);  tl[3 * kMdctTermStride + j] = fmul(v, v);
                    v = fmul(xm, wi); tl[4 * kMdctTermStride + j] = fmul(v, v);
                    v = fmul(y, wi);  tl[5 * kMdctTermStride + j] = fmul(v, v);
                    v = fmul(ym, wi); tl[6 * kMdctTermStride + j] = fmul(v, v);
                }
                __syncwarp();
                if (lane < 7) {
#pragma unroll 4
                    for (int j = 0; j < 128; j += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(&tl[lane * kMdctTermStride + j]);
                        a = fadd(a, q.x); a = fadd(a, q.y); a = fadd(a, q.z); a = fadd(a, q.w);
                    }
                }
                __syncwarp();
            }
            if (lane < 7) s_esum[warp][band][lane] = a;
            __syncwarp();
        }
        if (lane < 4) {
            const int band = lane;
            const Curve& cc = s_cv[warp][band][1];
            const bool cur_empty = cc.n == 0, prev_empty = s_cv[warp][band][0].n == 0;
            float sc0 = 1.0f, sc1 = 1.0f, sc2 = 1.0f, sc3 = 1.0f;
            if (!((trivial >> band) & 1u)) {
                const float* es = s_esum[warp][band];
                float pos_scale;                                       // PrevOverlapGainScale[channel][band]
                if (f == 0) pos_scale = b.next_scale[sc * 4 + band];
                else if (prev_empty) pos_scale = 1.0f;                 // SafeEnergyScale(e, e)
                else pos_scale = safe_energy_scale(es[5], es[6]);
                const float inf = __int_as_float(0x7f800000);
                if (!(fabsf(pos_scale) < inf) || pos_scale <= 0.0f) pos_scale = 1.0f;
                const float prev_div = cc.n ? T->gain_level[cc.level[0]] : 1.0f;
                const float prev_stored = es[0];
                const float prev_orig = fmul(prev_stored, pos_scale);
                const float prev_mod = __fdiv_rn(prev_stored, fmul(prev_div, prev_div));
                const float cur_orig = es[1], cur_mod = cur_empty ? es[1] : es[2];
                const float nxt_orig = es[3], nxt_mod = cur_empty ? es[3] : es[4];
                sc0 = safe_energy_scale(prev_orig, prev_mod);
                sc1 = safe_energy_scale(cur_orig, cur_mod);
                sc2 = safe_energy_scale(fadd(prev_orig, cur_orig), fadd(prev_mod, cur_mod));
                sc3 = safe_energy_scale(nxt_orig, nxt_mod);
            }
            *reinterpret_cast<float4*>(&b.gscale[(size_t)unit * 16 + band * 4]) = make_float4(sc0, sc1, sc2, sc3);
            if (f == g.n_out - 1) b.next_scale_out[sc * 4 + band] = sc3;
        }
        float* const outp = b.specs + (size_t)unit * 1024;
#pragma unroll 1
        for (int bp2 = 0; bp2 < 2; bp2++) {                 // bands 2*bp2, 2*bp2 + 1
        __syncwarp();                                       // the tile is free (rare path / previous iteration)
        // ---- phase A: MDCT input (atrac3denc.cpp:39-49): in[j] = stored half / scale, in[256+j] = win[255-j] * modulated cur
#pragma unroll 1
        for (int bb = 0; bb < 2; bb++) {
            const int band = 2 * bp2 + bb;
            // BL = 128 + 256 L: every band row and every frame inside it starts 16-byte aligned
            const float* bp = b.bands + (sc * 4 + band) * g.BL + 128 + 256 * (size_t)f;
            const Curve& cc = s_cv[warp][band][1];
            const Curve& pc = s_cv[warp][band][0];
            float* in = tl + bb * kMdctBandStride;
            float* lv_c = tl + kMdctLevels;
            float* lv_p = lv_c + 256;
            if (cc.n | pc.n) {                                        // divisor tables of this band (rare)
                __syncwarp();
                if (cc.n) curve_levels_warp(T, cc, lane, lv_c);
                if (f != 0 && pc.n) curve_levels_warp(T, pc, lane, lv_p);
                __syncwarp();
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i0 = 4 * (lane + 32 * h);
                float x[4], y[4], pv[4], cu[4];
                {
                    const float4 q = *reinterpret_cast<const float4*>(bp + i0);
                    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
                }
                if (f == 0) {
                    const float4 q = *reinterpret_cast<const float4*>(b.prevhalf + (sc * 4 + band) * 256 + i0);
                    pv[0] = q.x; pv[1] = q.y; pv[2] = q.z; pv[3] = q.w;
                } else {
                    const float4 q = *reinterpret_cast<const float4*>(bp + i0 - 256);
                    y[0] = q.x; y[1] = q.y; y[2] = q.z; y[3] = q.w;
                }
                if (cc.n) {                                           // gain modulation (gain_processor.h:87-121)
                    const float4 d = *reinterpret_cast<const float4*>(lv_c + i0);
                    x[0] = __fdiv_rn(x[0], d.x); x[1] = __fdiv_rn(x[1], d.y); x[2] = __fdiv_rn(x[2], d.z); x[3] = __fdiv_rn(x[3], d.w);
                }
                if (f != 0 && pc.n) {
                    const float4 d = *reinterpret_cast<const float4*>(lv_p + i0);
                    y[0] = __fdiv_rn(y[0], d.x); y[1] = __fdiv_rn(y[1], d.y); y[2] = __fdiv_rn(y[2], d.z); y[3] = __fdiv_rn(y[3], d.w);
                }
                const float4 wf = *reinterpret_cast<const float4*>(&s_win[i0]);          // win[i0 .. i0+3]
                const float4 wb = *reinterpret_cast<const float4*>(&s_win[252 - i0]);    // win[255-i0-3 .. 255-i0]
                const float wi[4] = {wf.x, wf.y, wf.z, wf.w}, wr[4] = {wb.w, wb.z, wb.y, wb.x};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (f != 0) pv[e] = fmul(wi[e], y[e]);
                    cu[e] = fmul(wr[e], x[e]);
                }
                if (cc.n) {
                    const float d0 = T->gain_level[cc.level[0]];
#pragma unroll
                    for (int e = 0; e < 4; e++) pv[e] = __fdiv_rn(pv[e], d0);
                }
                if (f == g.n_out - 1) {                               // the half this frame leaves behind (next batch)
                    *reinterpret_cast<float4*>(b.prevhalf_out + (sc * 4 + band) * 256 + i0) =
                        make_float4(fmul(wi[0], x[0]), fmul(wi[1], x[1]), fmul(wi[2], x[2]), fmul(wi[3], x[3]));
                }
                *reinterpret_cast<float4*>(in + i0) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                *reinterpret_cast<float4*>(in + 256 + i0) = make_float4(cu[0], cu[1], cu[2], cu[3]);
            }
        }
        __syncwarp();
        // ---- phase B: fold + pre-twiddle + the two innermost FFT stages on gather block k16
        const int bsel = lane >> 4, k16 = lane & 15, k = lane & 7, hp = (lane >> 3) & 1;
        cpx e[4][2];                                       // after the exchange: element k + 8a + 32(2 hp + bb)
        {
            const float* in = tl + bsel * kMdctBandStride;
            cpx v[8];
            // slot = 8 k16 + j of the 4x4x4x2 digit reversal: i = d0 + 4 d1 + 16 d2 + 64 d3
            const int ibase = (k16 >> 2) + 4 * (k16 & 3);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int n = 2 * (ibase + 16 * (j >> 1) + 64 * (j & 1));   // N = 512, n4 = 128, n34 = 384, n54 = 640
                float r0, i0;
                if ((j & 1) == 0) { r0 = fadd(in[383 - n], in[384 + n]); i0 = fsub(in[128 + n], in[127 - n]); }   // n < 128
                else              { r0 = fsub(in[383 - n], in[n - 128]); i0 = fadd(in[128 + n], in[639 - n]); }
                const float2 cs = *reinterpret_cast<const float2*>(&s_sincos[n]);
                v[j].r = fadd(fmul(r0, cs.x), fmul(i0, cs.y));
                v[j].i = fsub(fmul(i0, cs.x), fmul(r0, cs.y));
            }
            // radix-2, m = 1 (fstride 64): pairs (2q, 2q+1), twiddle tw[0]
#pragma unroll
            for (int q = 0; q < 4; q++) kf_bfly2(v[2 * q], v[2 * q + 1], s_tw[0]);
            // radix-4, m = 2 (fstride 16): elements kk + 2q, twiddles tw[16 kk q]
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                kf_bfly4<false>(v[kk], v[kk + 2], v[kk + 4], v[kk + 6], s_tw[16 * kk], s_tw[32 * kk], s_tw[48 * kk]);
            __syncwarp();                                  // every lane has read its MDCT input
            cpx* xch = reinterpret_cast<cpx*>(tl) + bsel * kMdctXchStride;
#pragma unroll
            for (int j = 0; j < 8; j++) xch[k16 * 9 + j] = v[j];
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int bb = 0; bb < 2; bb++) e[a][bb] = xch[(a + 4 * (2 * hp + bb)) * 9 + k];
        }
        // ---- phase C: radix-4 m = 8 (fstride 4) over a
        {
            const cpx t1 = s_tw[4 * k], t2 = s_tw[8 * k], t3 = s_tw[12 * k];
#pragma unroll
            for (int bb = 0; bb < 2; bb++) kf_bfly4<false>(e[0][bb], e[1][bb], e[2][bb], e[3][bb], t1, t2, t3);
        }
        // lane pair swap: lane hp keeps a = 2hp + aa and receives the partner's b range for those a
        cpx gq[2][4];                                      // element k + 8 (2 hp + aa) + 32 b
#pragma unroll
        for (int aa = 0; aa < 2; aa++)
#pragma unroll
            for (int bb = 0; bb < 2; bb++) {
                const cpx keep = hp ? e[2 + aa][bb] : e[aa][bb];
                const cpx send = hp ? e[aa][bb] : e[2 + aa][bb];
                cpx recv;
                recv.r = __shfl_xor_sync(0xffffffffu, send.r, 8);
                recv.i = __shfl_xor_sync(0xffffffffu, send.i, 8);
                gq[aa][bb] = hp ? recv : keep;
                gq[aa][2 + bb] = hp ? keep : recv;
            }
        // radix-4 m = 32 (fstride 1) over b
#pragma unroll
        for (int aa = 0; aa < 2; aa++) {
            const int v = k + 8 * (2 * hp + aa);
            kf_bfly4<false>(gq[aa][0], gq[aa][1], gq[aa][2], gq[aa][3], s_tw[v], s_tw[2 * v], s_tw[3 * v]);
        }
        __syncwarp();                                      // exchange reads done: the tile becomes the output stage
        // post-twiddle (mdct.h:92-101); odd bands reversed (atrac3denc.cpp:53-55)
        {
            float* sp = tl + bsel * kMdctOutStride;
#pragma unroll
            for (int aa = 0; aa < 2; aa++)
#pragma unroll
                for (int bq = 0; bq < 4; bq++) {
                    const int n = 2 * (k + 8 * (2 * hp + aa) + 32 * bq);
                    const cpx z = gq[aa][bq];
                    const float2 cs = *reinterpret_cast<const float2*>(&s_sincos[n]);
                    const float va = fsub(fmul(-z.r, cs.x), fmul(z.i, cs.y));
                    const float vb = fadd(fmul(-z.r, cs.y), fmul(z.i, cs.x));
                    int pa = n, pb = 255 - n;
                    if (bsel) { pa = 255 - pa; pb = 255 - pb; }       // band 2*bp2 + bsel is odd iff bsel
                    sp[pa] = va;
                    sp[pb] = vb;
                }
        }
        __syncwarp();
        float4* out = reinterpret_cast<float4*>(outp + 512 * bp2);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int w = lane + 32 * q;                   // float4 index 0..127; band within the pair = w >> 6
            out[w] = *reinterpret_cast<const float4*>(tl + (w >> 6) * kMdctOutStride + 4 * (w & 63));
        }
        }
    }
}

void launch_mdct(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const long long n_units = (long long)g.S * g.n_out * g.C;
    long long blocks = (n_units + kMdctWarps - 1) / kMdctWarps;
    if (blocks > 148 * 8 * 4) blocks = 148 * 8 * 4;       // a few waves of resident blocks, warps stride over units
    ATDE_LAUNCH(at3_mdct_kernel, (unsigned)blocks, kMdctWarps * 32, 0, st, g, b);
}

// =====================================================================================
// K6: loudness recurrence, one thread per stream
// =====================================================================================
__global__ void at3_loudness_kernel(Geometry g, Buffers b)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.S) return;
    float L = b.loud_state[s];
    for (int f = 0; f < g.n_out; f++) {
        const size_t o = ((size_t)s * g.n_out + f) * g.C;
        if (g.C == 2 && !g.js) {
            const float sum = fadd(b.chloud[o], b.chloud[o + 1]);
            L = __double2float_rn(__dadd_rn(__dmul_rn(0.98, (double)L), __dmul_rn(0.01, (double)sum)));
        } else {
            L = __double2float_rn(__dadd_rn(__dmul_rn(0.98, (double)L), __dmul_rn(0.02, (double)b.chloud[o])));
        }
        b.loud[(size_t)s * g.n_out + f] = L;
    }
    b.loud_state[s] = L;
}

void launch_loudness(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    ATDE_LAUNCH(at3_loudness_kernel, (g.S + 63) / 64, 64, 0, st, g, b);
}

// =====================================================================================
// carry: keep the last two extended PCM frames of every stream for the next batch
// =====================================================================================
__global__ void at3_carry_kernel(Geometry g, Buffers b)
{
    float* hist_out = b.hist_tmp;
    const int s = blockIdx.x;
    const bool started = b.started[s] != 0;
    const int n = 2 * 1024 * g.C;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i % g.C;
        const int t = i / g.C;                                  // 0..2047: ext frames L-2, L-1
        const long long ext = (long long)(g.L - 2) * 1024 + t;
        hist_out[(size_t)s * n + i] = virt_pcm(g, b, s, c, ext, started);
    }
}

__global__ void at3_carry_commit_kernel(Geometry g, Buffers b)
{
    const float* hist_in = b.hist_tmp;
    unsigned char* started = b.started;
    const int s = blockIdx.x;
    const int n = 2 * 1024 * g.C;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        b.pcm_hist[(size_t)s * n + i] = hist_in[(size_t)s * n + i];
    if (g.n_out > 0) {
        const int m = g.C * 4 * 256;
        for (int i = threadIdx.x; i < m; i += blockDim.x)
            b.prevhalf[(size_t)s * m + i] = b.prevhalf_out[(size_t)s * m + i];
        for (int i = threadIdx.x; i < g.C * 4; i += blockDim.x)
            b.next_scale[(size_t)s * g.C * 4 + i] = b.next_scale_out[(size_t)s * g.C * 4 + i];
    }
    if (threadIdx.x == 0) started[s] = 1;
}

void launch_carry(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    ATDE_LAUNCH(at3_carry_kernel, g.S, 256, 0, st, g, b);
    ATDE_LAUNCH(at3_carry_commit_kernel, g.S, 256, 0, st, g, b);
}

} // namespace at3
} // namespace atde
