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
#ifndef ATDE_K1_NO_TMA_LOAD
    __shared__ mbar_t ldbar;
    if (tid == 0) mbar_init(&ldbar, 1);
    async_proxy_fence();
    __syncthreads();
#endif
    {
        const long long n0 = 4LL * m0 - 144 - (started ? 1024 : 0);    // tile sample 0 as an index into b.pcm
        const long long nlim = (long long)g.N * 1024;
        if (g.C == 2 && n0 >= 0 && n0 + kQX <= nlim) {
            const float2* src = reinterpret_cast<const float2*>(b.pcm + ((size_t)s * g.N * 1024 + (size_t)n0) * 2);
#ifndef ATDE_K1_NO_TMA_LOAD
            // interior tile: the interleaved PCM (16.4 KB, 32-byte aligned) comes in by ONE bulk asynchronous copy into the
            // stage-1 output area, which is idle until the tile has been de-interleaved, scaled and padded
            float* raw = &s1[0][0][0];
            static_assert(4 * kQS1P >= 2 * kQX && (kQX * 8) % 16 == 0 && kQX % 2 == 0, "staging area and copy size");
            if (tid == 0) {
                mbar_expect_tx(&ldbar, (unsigned)(kQX * 8));
                bulk_g2s(raw, src, (unsigned)(kQX * 8), &ldbar);
            }
            mbar_wait(&ldbar, 0);
            for (int t2 = tid; t2 < kQX / 2; t2 += 256) {                // two stereo samples per 16-byte load
                const float4 v = reinterpret_cast<const float4*>(raw)[t2];
                const int p = qphys(2 * t2);                              // samples 2 t2, 2 t2 + 1 are neighbours in the padded layout
                *reinterpret_cast<float2*>(&xs[0][p]) = make_float2(fmul(v.x, 0.25f), fmul(v.z, 0.25f));   // data / 4.0 (atrac3denc.cpp:704)
                *reinterpret_cast<float2*>(&xs[1][p]) = make_float2(fmul(v.y, 0.25f), fmul(v.w, 0.25f));
            }
#else
            // interior tile: every load independent and in flight at once
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
#endif
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
        static_assert(kQR == 8 && (kQT * 4) % 16 == 0, "two 16-byte stores per task and row");
        reinterpret_cast<float4*>(ol)[0] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        reinterpret_cast<float4*>(ol)[1] = make_float4(lo[4], lo[5], lo[6], lo[7]);
        reinterpret_cast<float4*>(oh)[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        reinterpret_cast<float4*>(oh)[1] = make_float4(hi[4], hi[5], hi[6], hi[7]);
    }
    __syncthreads();
    if (g.js) {
        // Matrixing (atrac3denc.cpp:665-677): (L + R) / 2, (L - R) / 2, in place in the output tile
        ATDE_PAR_FOR(w, kQT) {                                    // 4 bands x kQT samples = kQT float4 columns
            const int band = w / (kQT / 4), q4 = w - band * (kQT / 4);
            float4* pl = reinterpret_cast<float4*>(outb + band * kQT) + q4;
            float4* pr = reinterpret_cast<float4*>(outb + (4 + band) * kQT) + q4;
            const float4 l = *pl, r = *pr;
            *pl = make_float4(fmul(fadd(l.x, r.x), 0.5f), fmul(fadd(l.y, r.y), 0.5f), fmul(fadd(l.z, r.z), 0.5f), fmul(fadd(l.w, r.w), 0.5f));
            *pr = make_float4(fmul(fsub(l.x, r.x), 0.5f), fmul(fsub(l.y, r.y), 0.5f), fmul(fsub(l.z, r.z), 0.5f), fmul(fsub(l.w, r.w), 0.5f));
        }
    }
    // The eight (channel, band) rows of the tile leave by bulk asynchronous copies (1-D TMA): rows are 16-byte aligned
    // on both sides (kQT * 4 and BL * 4 are multiples of 16) and so is the length of a last, shorter tile.
    async_proxy_fence();
    __syncthreads();
    if (tid == 0) {
        const int nvalid = min(kQT, g.BL - u0);
        for (int c = 0; c < g.C; c++)
            for (int band = 0; band < 4; band++)
                bulk_s2g(b.bands + (((size_t)s * g.C + c) * 4 + band) * g.BL + u0, outb + (c * 4 + band) * kQT, (unsigned)nvalid * 4u);
        bulk_store_commit();
        bulk_store_wait_read();                                  // shared memory is released when the block retires
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

// FindPlateau (transient_detector.cpp:175-262) + target selection (:286-299) over the 32 sub-frame levels:
//   filt = MedianFilter<1>(in); best = max over j of min(filt[j..j+2]) (first j wins), the run is extended while
//   filt >= best; "release" if the frame ends far below the plateau; the plateau is the target unless it is released,
//   negligible or below 0.4 of the frame's maximum — then the last level is.
// By a whole warp: lane i holds in[i] (all values >= 0: they are RMS levels, so their bit patterns order
// like the floats and redux.sync finds maxima).  Every lane returns the target.
ATDE_D float plateau_target_warp(float a, int lane)
{
    const unsigned full = 0xffffffffu;
    const float left = __shfl_up_sync(full, a, 1), right = __shfl_down_sync(full, a, 1);
    const float filt = lane == 0 ? fmaxf(a, right) : (lane == 31 ? fmaxf(left, a) : median3(left, a, right));
    const float max_raw = __uint_as_float(__reduce_max_sync(full, __float_as_uint(a)));
    const float f1 = __shfl_down_sync(full, filt, 1), f2 = __shfl_down_sync(full, filt, 2);
    const float mv = lane <= 29 ? fminf(fminf(filt, f1), f2) : 0.0f;
    const float best = __uint_as_float(__reduce_max_sync(full, __float_as_uint(mv)));      // 0 when no run is above 0
    const float last = __shfl_sync(full, a, 31);
    bool release = false;
    float level = best;
    if (best < 1e-6f) {
        level = 0.0f;
    } else {
        // first run that reaches the maximum (the serial scan only replaces on '>'), extended while filt >= best
        int best_end = __ffs((int)__ballot_sync(full, lane <= 29 && mv == best)) - 1 + 2;
        const unsigned below = ~__ballot_sync(full, filt >= best) & (best_end >= 31 ? 0u : (full << (best_end + 1)));
        best_end = below ? __ffs((int)below) - 2 : 31;
        const unsigned high = __ballot_sync(full, a >= fmul(best, 0.7f));
        if (best_end < 31) {
            if (last < fmul(best, 0.1f)) release = true;
            else release = !(high & (full << (best_end + 1))) && (last < fmul(best, 0.5f));
        }
    }
    const bool use_plateau = level > 1e-6f && !release && level >= fmul(max_raw, 0.4f);
    return use_plateau ? level : last;
}

constexpr int kGainThreads = 128;         // threads per block
constexpr int kGainBlock = kGainThreads;

// The 2048-point buffer is padded by one element per 16.  With 8-byte elements a half-warp access is conflict-free when
// its 16 addresses differ mod 16, and this one padding gives that to all three passes:
//   pass 1  thread grp stores slots 8 grp + j:             8 grp + grp / 2 + j       -> grp / 2 + 8 (grp & 1) + j  (mod 16)
//   pass 2  lanes (k < 8, two groups G):                   136 G + k + const         -> k + 8 (G & 1) + const
//   pass 3  16 consecutive k:                              k + k / 16 + const        -> a rotation of 0..15
ATDE_D int gphys(int i) { return i + (i >> 4); }
// The real output signal is padded by 4 floats per 64 so that the 32 sequential 64-sample RMS sums
// (one lane each, 64 floats apart) read different banks.
ATDE_D int sphys(int j) { return j + ((j >> 6) << 2); }
// Spectrum-sized arrays (forward FFT buffer, frequency bins, staged super-twiddles) are padded by one
// element per 16: both the natural-order accesses and the base-4 digit-reversed ones (lane stride 64,
// 16, 4 elements) then fall on 16 different bank pairs.
ATDE_D int fq(int k) { return k + (k >> 4); }

// One radix-4 butterfly of the forward FFT-256 on the padded buffer: elements p, p + d, p + 2d, p + 3d in PADDED
// coordinates (the callers fold fq() into p and d: every stage touches elements whose padding is an affine function of
// the butterfly number, so the addresses are base + constant and need no per-access index arithmetic).
ATDE_D void fwd_bfly(cpx* buf, int p, int d, cpx t1, cpx t2, cpx t3)
{
    cpx f0 = buf[p], f1 = buf[p + d], f2 = buf[p + 2 * d], f3 = buf[p + 3 * d];
    kf_bfly4<false>(f0, f1, f2, f3, t1, t2, t3);
    buf[p] = f0; buf[p + d] = f1; buf[p + 2 * d] = f2; buf[p + 3 * d] = f3;
}

// Two consecutive frames of one (stream, channel, band) per block.
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
#ifndef ATDE_GAIN_BLOCKS
#define ATDE_GAIN_BLOCKS 8
#endif
// TRACE = true is the instance behind the reference's `--yaml-log` gain-control trace (atrac3denc.cpp:305-400): it also
// covers band 3 (analysed and logged by the reference although it never carries a curve), takes the high-frequency
// ratio by the reference's sequential sums and forms `next_level` — the RMS of the first 64 up-sampled look-ahead
// samples, output samples [3072, 3136) — and writes to the trace buffers only.  The encode path never launches it.
template <bool TRACE>
__global__ void __launch_bounds__(kGainBlock, ATDE_GAIN_BLOCKS) at3_gain_kernel(Geometry g, Buffers b)
{
    constexpr int kNb = TRACE ? kBands : kGainBands;
    __shared__ __align__(16) cpx big[2048 + 128];    // inverse FFT buffer, padded; later the real output
    __shared__ __align__(8) cpx tw2c[15][8];         // pass-2 twiddles of lane group k (8-byte elements: a 16-byte pre-spread
                                                     // form costs four wavefronts per warp load instead of two, and the
                                                     // kernel is bound by the shared-memory pipe)
    __shared__ __align__(16) cpx freq2[2][257 + 17]; // spectra of the block's two frames
    __shared__ __align__(16) cpx sup[257 + 17];      // super4096[k - 1] at fq(k), k = 1..256
    // Shared memory is the kernel's occupancy limit (8 blocks per SM) and every KB saved is L1 for the twiddle tables, so
    // the big buffer is used three times over: the two forward FFT buffers live in it while the forward transforms run,
    // and once a frame's real output lies in its first 2176 floats, the 256 micro-chunk RMS values and the 96 output
    // values are staged behind them
    cpx (*const fwd2)[257 + 17] = reinterpret_cast<cpx (*)[257 + 17]>(big);
    float* const micro = reinterpret_cast<float*>(big) + 2304;
    float* const sgain = reinterpret_cast<float*>(big) + 2560;
    __shared__ double esum_part[kGainThreads / 32][2];   // per-warp partial sums of (|X_k|^2, |X_k H_k|^2): warps 2h, 2h+1 = frame h
    __shared__ float sstat[2];

    const DevTables* __restrict__ T = b.tab;
    // A block takes TWO consecutive frames of one (stream, channel, band).  The forward FFT-256 has 64 butterflies per
    // stage: the two halves of the block run the two frames' forward transforms side by side (every thread busy, half
    // the barriers per frame); the inverse FFT-2048 wants all 128 threads and runs for one frame after the other.
    const int f0 = 2 * blockIdx.x;
    const int n_fr = min(2, g.n_out - f0);
    const int band = blockIdx.y % kNb, c = blockIdx.y / kNb;
    const int s = blockIdx.z;
    const int tid = threadIdx.x;
    const int half = tid >> 6, t = tid & 63;
    float* const nxt = reinterpret_cast<float*>(big) + 2688;     // TRACE: output samples [3072, 3136)
    const bool act = half < n_fr;
    const float* __restrict__ in = b.bands + (((size_t)s * g.C + c) * 4 + band) * g.BL + 256 * (size_t)(f0 + half);
    cpx* const fwd = fwd2[half];
    const cpx* __restrict__ tw = T->tw2048;

    // the frame's 512 band samples come from DRAM: their loads are issued before anything else (four per thread, all in
    // flight while the twiddles are staged)
    float2 xin[4];
#pragma unroll
    for (int q = 0; q < 4; q++) xin[q] = act ? *reinterpret_cast<const float2*>(in + 2 * (t + 64 * q)) : make_float2(0.0f, 0.0f);
    // ... and so are the thread's twiddles of the four forward stages (each stage used to fetch its own right behind a barrier)
    cpx ft[4][3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        ft[0][q] = T->ftw[0][q][0];
        ft[1][q] = T->ftw[1][q][t >> 4];
        ft[2][q] = T->ftw[2][q][t & 15];
        ft[3][q] = T->ftw[3][q][t];
    }
    if (tid < 120) (&tw2c[0][0])[tid] = (&T->gtw2[0][0])[tid];
    f32x2 one2, mone2;
    one2.x = one2.y = g.one;
    mone2.x = mone2.y = -g.one;
    // the only super-twiddles kiss_fftri(4096) meets with a non-zero operand
    sup[fq(1 + tid)] = T->super4096[tid];
    sup[fq(1 + tid + kGainBlock)] = T->super4096[tid + kGainBlock];
    // 1. Planck window, packed as the complex input of the half-size FFT; loaded in natural order and
    //    stored at its digit-reversed slot (base-4 reversal of 4 digits is an involution)
    if (act) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = t + 64 * q;
            const float2 x = xin[q];
            const float2 w = *reinterpret_cast<const float2*>(&T->planck[2 * j]);
            cpx z;
            z.r = fmul(x.x, w.x);
            z.i = fmul(x.y, w.y);
            const int o = ((j & 3) << 6) | (((j >> 2) & 3) << 4) | (((j >> 4) & 3) << 2) | (j >> 6);
            fwd[fq(o)] = z;
        }
    }
    __syncthreads();
    // 2. forward complex FFT-256 = 4x4x4x4, innermost stage first; the lane -> butterfly map of every
    //    stage is chosen so that 16 consecutive lanes touch 16 different bank pairs
    //    (padded addresses: element F + m q of a butterfly lies at fq(F) + q * (m + m / 16))
    if (act) fwd_bfly(fwd, 4 * t + (t >> 2), 1, ft[0][0], ft[0][1], ft[0][2]);
    __syncthreads();
    if (act) fwd_bfly(fwd, 17 * (t & 15) + (t >> 4), 4, ft[1][0], ft[1][1], ft[1][2]);
    __syncthreads();
    if (act) fwd_bfly(fwd, 68 * (t >> 4) + (t & 15), 17, ft[2][0], ft[2][1], ft[2][2]);
    __syncthreads();
    if (act) fwd_bfly(fwd, t + (t >> 4), 68, ft[3][0], ft[3][1], ft[3][2]);
    __syncthreads();
    // kiss_fftr post-processing (kiss_fftr.c:84-115)
    if (act) {
        cpx* const fo = freq2[half];
        for (int k = t; k < 129; k += 64) {
            if (k == 0) {
                const float tr = fwd[0].r, ti = fwd[0].i;
                fo[0].r = fadd(tr, ti);   fo[0].i = 0.0f;
                fo[fq(256)].r = fsub(tr, ti); fo[fq(256)].i = 0.0f;
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
                fo[fq(k)] = a;                     // k == 128 writes the same element twice: the second
                fo[fq(256 - k)] = bb;              // store (freqdata[ncfft-k]) wins, as in the reference
            }
        }
    }
    __syncthreads();
    // 2a. high-frequency energy ratio (upsampler.cpp:99-118).  The reference adds the 257 per-bin energies (and the
    //     HPF-weighted ones) SEQUENTIALLY in double and hands float(hi / tot) to the curve builder, where the value is
    //     only ever COMPARED with kHighFreqThreshold = 0.05f and with 0.3f (atrac3denc.cpp:352,431).  Here the two sums
    //     are taken as a tree over the frame's two warps — any summation order of 257 non-negative doubles agrees with
    //     the sequential one to 257 * 2^-53 relative, the ratio to ~1.2e-13 — and a result that lands within 1e-7
    //     (relative) of one of the two thresholds, where the float rounding of the ratio could decide a comparison
    //     differently, is recomputed with the reference's sequential loop at the end of the kernel.
    const int lcb = T->low_cut_bin;
    {
        double tot = 0.0, hi = 0.0;
        if (act) {
            const cpx* const fi = freq2[half];
            for (int k = t; k < 257; k += 64) {
                const cpx z = fi[fq(k)];
                const double r = (double)z.r, i = (double)z.i;
                const double e = __dadd_rn(__dmul_rn(r, r), __dmul_rn(i, i));
                float H = 0.0f;
                if (k >= lcb + 2) H = 1.0f;
                else if (k >= lcb) H = T->hpf_h[k - lcb];
                tot = __dadd_rn(tot, e);
                hi = __dadd_rn(hi, __dmul_rn(__dmul_rn(e, (double)H), (double)H));
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            tot = __dadd_rn(tot, __shfl_xor_sync(0xffffffffu, tot, d));
            hi = __dadd_rn(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        }
        if ((tid & 31) == 0) { esum_part[tid >> 5][0] = tot; esum_part[tid >> 5][1] = hi; }
    }
    __syncthreads();
    for (int it = 0; it < n_fr; it++) {
    const cpx* const freq = freq2[it];
    const int f = f0 + it;
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
    // tmpbuf[k] alone / tmpbuf[2048-k] alone for a kept bin 1 <= k < 256 with the structural zeros folded away:
    // fek = tp = fk (adding +-0), hi.i = -(fk.i - fok.i) = fok.i - fk.i
    auto scaled_bin = [&](int k) {
        cpx fk;
        fk.r = fmul(freq[fq(k)].r, 8.0f);
        fk.i = fmul(freq[fq(k)].i, 8.0f);
        if (k < lcb + 2) {
            const float w = T->hpf_h[k - lcb];
            fk.r = fmul(fk.r, w);
            fk.i = fmul(fk.i, w);
        }
        return fk;
    };
    for (int grp = tid; grp < 256; grp += kGainThreads) {
        const int low = ((grp >> 6) & 3) | (((grp >> 4) & 3) << 2) | (((grp >> 2) & 3) << 4) | ((grp & 3) << 6);
        cpx o0, o1, o2, o3, o4, o5, o6, o7;
        if (low == 0) {
            // the one group with a third input (tmpbuf[256] in slot 2), and with tmpbuf[0] when bin 0 is kept: written out in full
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
        } else {
            // slots 0 and 7 only: x0 = tmpbuf[low], x7 = tmpbuf[1792 + low].  tw[0] = (1, 0) and every other input of the two
            // butterflies is a structural zero, so (each line below is what kf_bfly2 / kf_bfly4 leave once the exact
            // operations on zeros and the multiplications by one are dropped; only the sign of a zero can differ, and
            // the consumers square the output):
            //   radix-2:  F6 = x7, F7 = -x7
            //   radix-4, k = 0: s2 = x7, s3 = x7, s4 = -x7          radix-4, k = 1: s2 = -c, s3 = -c, s4 = c, c = x7 * tw[768]
            cpx x0, x7;
            x0.r = x0.i = x7.r = x7.i = 0.0f;
            if (low >= lcb) {
                const cpx fk = scaled_bin(low);
                const cpx fok = cmul(fk, sup[fq(low)]);
                x0.r = fadd(fk.r, fok.r); x0.i = fadd(fk.i, fok.i);
            }
            if (256 - low >= lcb) {
                const cpx fk = scaled_bin(256 - low);
                const cpx fok = cmul(fk, sup[fq(256 - low)]);
                x7.r = fsub(fk.r, fok.r); x7.i = fsub(fok.i, fk.i);
            }
            o0.r = fadd(x0.r, x7.r); o0.i = fadd(x0.i, x7.i);
            o4.r = fsub(x0.r, x7.r); o4.i = fsub(x0.i, x7.i);
            o2.r = fadd(x0.r, x7.i); o2.i = fsub(x0.i, x7.r);
            o6.r = fsub(x0.r, x7.i); o6.i = fadd(x0.i, x7.r);
            const cpx c = cmul(x7, tw[768]);
            o1.r = fsub(x0.r, c.r); o1.i = fsub(x0.i, c.i);
            o5.r = fadd(x0.r, c.r); o5.i = fadd(x0.i, c.i);
            o3.r = fsub(x0.r, c.i); o3.i = fadd(x0.i, c.r);
            o7.r = fadd(x0.r, c.i); o7.i = fsub(x0.i, c.r);
        }
        cpx* dst = big + gphys(8 * grp);
        dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3; dst[4] = o4; dst[5] = o5; dst[6] = o6; dst[7] = o7;
    }
    atde_named_barrier(1, kGainThreads);
    // pass 2: radix-4 m = 8 (fstride 64), then m = 32 (fstride 16)
    {
        // gphys(base + k + 8a + 32q) = gphys(base + k) + 8a + a / 2 + 34q  (k < 8, base a multiple of 128)
        const int k = tid & 7;
        cpx* const bg = big + gphys(((tid >> 3) << 7) + k);
        cpx x[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) x[a][q] = bg[8 * a + (a >> 1) + 34 * q];
        {
            const tw4 t1 = spread_twiddle_dev(tw2c[0][k], mone2.x), t2 = spread_twiddle_dev(tw2c[1][k], mone2.x),
                      t3 = spread_twiddle_dev(tw2c[2][k], mone2.x);           // tw[64k], tw[128k], tw[192k]
#pragma unroll
            for (int q = 0; q < 4; q++) kf_bfly4_packed<true>(x[0][q], x[1][q], x[2][q], x[3][q], t1, t2, t3, one2, mone2);
        }
#pragma unroll
        for (int a = 0; a < 4; a++)                                            // kk = k + 8a: tw[16kk], tw[32kk], tw[48kk]
            kf_bfly4_packed<true>(x[a][0], x[a][1], x[a][2], x[a][3], spread_twiddle_dev(tw2c[3 + 3 * a][k], mone2.x),
                                  spread_twiddle_dev(tw2c[4 + 3 * a][k], mone2.x), spread_twiddle_dev(tw2c[5 + 3 * a][k], mone2.x), one2, mone2);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) bg[8 * a + (a >> 1) + 34 * q] = x[a][q];
    }
    atde_named_barrier(1, kGainThreads);
    // pass 3: radix-4 m = 128 (fstride 4), then m = 512 (fstride 1); keep slots [512, 1536), normalised
    {
        // gphys(k + 128a + 512q) = k + k / 16 + 136a + 544q  (k < 128)
        const int k = tid;
        const cpx* const bg = big + k + (k >> 4);
        // (the first stage's twiddles are requested before the shared-memory reads and the barrier, not behind them)
        const cpx g3a0 = T->gtw3a[0][k], g3a1 = T->gtw3a[1][k], g3a2 = T->gtw3a[2][k];
        cpx x[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) x[a][q] = bg[136 * a + 544 * q];
        atde_named_barrier(1, kGainThreads);                                  // every slot is in registers: big can be overwritten
        {
            const float mone = mone2.x;
            const tw4 t1 = spread_twiddle_dev(g3a0, mone), t2 = spread_twiddle_dev(g3a1, mone),
                      t3 = spread_twiddle_dev(g3a2, mone);                               // tw[4k], tw[8k], tw[12k]
#pragma unroll
            for (int q = 0; q < 4; q++) kf_bfly4_packed<true>(x[0][q], x[1][q], x[2][q], x[3][q], t1, t2, t3, one2, mone2);
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            kf_bfly4_packed<true>(x[a][0], x[a][1], x[a][2], x[a][3], spread_twiddle_dev(T->gtw3b[a][0][k], mone2.x),
                                  spread_twiddle_dev(T->gtw3b[a][1][k], mone2.x), spread_twiddle_dev(T->gtw3b[a][2][k], mone2.x), one2, mone2);  // tw[kk], tw[2kk], tw[3kk]
            // 5. normalise (norm = 1/4096); complex slot kk + 512q -> output samples 2*slot, 2*slot+1
            cpx u, v;
            u.r = fmul(x[a][1].r, 1.0f / 4096.0f); u.i = fmul(x[a][1].i, 1.0f / 4096.0f);
            v.r = fmul(x[a][2].r, 1.0f / 4096.0f); v.i = fmul(x[a][2].i, 1.0f / 4096.0f);
            // sphys(2kk) = 2k + 4 (k / 32) + 272a;  sphys(1024 + 2kk) = that + 1088
            float* sigw = reinterpret_cast<float*>(big) + 2 * k + 4 * (k >> 5);
            *reinterpret_cast<cpx*>(sigw + 272 * a) = u;                  // slot 512 + kk  -> samples 1024 + 2kk, +1
            *reinterpret_cast<cpx*>(sigw + 272 * a + 1088) = v;           // slot 1024 + kk -> samples 2048 + 2kk, +1
            if (TRACE && a == 0 && k < 32) {                              // slot 1536 + k -> samples 3072 + 2k, +1
                nxt[2 * k] = fmul(x[0][3].r, 1.0f / 4096.0f);
                nxt[2 * k + 1] = fmul(x[0][3].i, 1.0f / 4096.0f);
            }
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
        micro[q] = __fsqrt_rn(fmul(a, 0.125f));               // a / 8: a product by a power of two is the same rounded quotient
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
        sgain[sf] = __fsqrt_rn(fmul(a, 0.015625f));         // a / 64
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
    } else if (tid >= 96) {
        const float tgt = plateau_target_warp(sgain[tid - 96], tid - 96);
        if (tid == 96) sstat[1] = tgt;
    }
    __syncthreads();
    const size_t item = ((((size_t)s * g.C + c) * kNb + band) * g.n_out + f);
    if (tid < 96) (TRACE ? b.trace_gain : b.gain)[item * 96 + tid] = sgain[tid];
    if (tid == 0) {
        double tot = __dadd_rn(esum_part[2 * it][0], esum_part[2 * it + 1][0]);
        double hi = __dadd_rn(esum_part[2 * it][1], esum_part[2 * it + 1][1]);
        if (tot > 0.0) {
            const double ratio = __ddiv_rn(hi, tot);
            const double t1 = (double)0.05f, t2 = (double)0.3f;
            if (TRACE || fabs(ratio - t1) <= t1 * 1e-7 || fabs(ratio - t2) <= t2 * 1e-7) {
                // too close to a decision threshold for a reordered sum: the reference's loop, bin by bin
                tot = 0.0; hi = 0.0;
                for (int k = 0; k <= 256; k++) {
                    const double r = (double)freq[fq(k)].r, i = (double)freq[fq(k)].i;
                    const double e = __dadd_rn(__dmul_rn(r, r), __dmul_rn(i, i));
                    float H = 0.0f;
                    if (k >= lcb + 2) H = 1.0f;
                    else if (k >= lcb) H = T->hpf_h[k - lcb];
                    tot = __dadd_rn(tot, e);
                    hi = __dadd_rn(hi, __dmul_rn(__dmul_rn(e, (double)H), (double)H));
                }
            }
        }
        float4 st4;
        st4.x = (tot > 0.0) ? __double2float_rn(__ddiv_rn(hi, tot)) : 0.0f;
        st4.y = sstat[0];
        st4.z = sstat[1];
        st4.w = sgain[31];
        if (TRACE) {
            // AnalyzeGain(signal + 3072, 64, 1, rms)[0] (atrac3denc.cpp:335, transient_detector.cpp:33-40)
            float a = 0.0f;
            for (int i = 0; i < 64; i++) a = fadd(a, fmul(nxt[i], nxt[i]));
            st4.w = __fsqrt_rn(fmul(a, 0.015625f));
            reinterpret_cast<float4*>(b.trace_stat)[item] = st4;           // hfr, curHpfEnergy, target, next_level
        } else {
            reinterpret_cast<float4*>(b.gstat)[item] = st4;
        }
    }
        __syncthreads();                       // the staged values are read: pass 1 of the next frame may overwrite them
    }   // frames of the block
}

void launch_gain_analysis(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    dim3 grid((g.n_out + 1) / 2, g.C * kGainBands, g.S);      // two frames per block
    ATDE_LAUNCH(at3_gain_kernel<false>, grid, kGainBlock, 0, st, g, b);
}

void launch_gain_trace(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    dim3 grid((g.n_out + 1) / 2, g.C * kBands, g.S);
    ATDE_LAUNCH(at3_gain_kernel<true>, grid, kGainBlock, 0, st, g, b);
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
    for (int f0 = 0; f0 < g.n_out; f0 += 8) {                        // (eight frames' statistics in flight at once)
        float4 vv[8];
#pragma unroll
        for (int q = 0; q < 8; q++) vv[q] = f0 + q < g.n_out ? st[f0 + q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (f0 + q < g.n_out) {
                const float4 v = vv[q];
                float4 o;
                o.x = ctx.y; o.y = ctx.x; o.z = ctx.z; o.w = 0.0f;
                if (v.x < 0.05f) {                                   // kHighFreqThreshold: LastLevel = 0, continue
                    ctx.x = 0.0f;
                } else {
                    ctx.y = v.y;
                    ctx.x = v.w;
                    ctx.z = v.z;
                }
                pv[f0 + q] = o;
            }
        }
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

ATDE_D Curve* curve_slot(const Geometry& g, const Buffers& b, long long item)
{
    const int f = (int)(item % g.n_out);
    const long long scb = item / g.n_out;
    const int band = (int)(scb % kGainBands);
    const long long sc = scb / kGainBands;
    return b.curves + ((size_t)sc * 4 + band) * g.n_out + f;
}

// The full curve construction of one item that passed the cheap tests of the kernel below.
ATDE_D void curve_item(const Geometry& g, const Buffers& b, long long item)
{
    const DevTables* __restrict__ T = b.tab;
    Curve out;
    out.n = 0; out.pad = 0;
    for (int i = 0; i < 7; i++) { out.level[i] = 0; out.loc[i] = 0; }
    Curve* dst = curve_slot(g, b, item);

    const float4 st = reinterpret_cast<const float4*>(b.gstat)[item];
    const float4 pv = reinterpret_cast<const float4*>(b.gprev)[item];
    const float hfr = st.x, cur_hpf = st.y, target = st.z;
    const float prev_hpf = pv.x, saved_ll = pv.y, saved_lt = pv.z;

    float gain[32], low[32], high[32];
    {   // the item's 384-byte record by 16-byte loads (a lane's record is contiguous: a quarter of the load instructions
        // and of the L1 wavefronts of scalar loads)
        const float4* __restrict__ gp4 = reinterpret_cast<const float4*>(b.gain + (size_t)item * 96);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 a = gp4[i], l = gp4[8 + i], h = gp4[16 + i];
            gain[4 * i] = a.x; gain[4 * i + 1] = a.y; gain[4 * i + 2] = a.z; gain[4 * i + 3] = a.w;
            low[4 * i] = l.x;  low[4 * i + 1] = l.y;  low[4 * i + 2] = l.z;  low[4 * i + 3] = l.w;
            high[4 * i] = h.x; high[4 * i + 1] = h.y; high[4 * i + 2] = h.z; high[4 * i + 3] = h.w;
        }
    }

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

// Most items leave after a glance at their statistics (no high-frequency content: atrac3denc.cpp:319-326; no target or
// no previous level: transient_detector.cpp:287-295) and the rest is long, branchy, per-item work: with one thread per
// item a warp ran its few surviving lanes at 7 of 32.  So the block first sorts its items — the cheap exits are
// answered on the spot, the others are collected in a list — and then walks the list with dense warps.
constexpr int kCurveBlock = 256;
__global__ void __launch_bounds__(kCurveBlock) at3_curve_kernel(Geometry g, Buffers b)
{
    __shared__ int list[kCurveBlock];
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const long long item0 = (long long)blockIdx.x * kCurveBlock;                 // item = ((s*C+c)*3+band)*n_out + f
    const long long total = (long long)g.S * g.C * kGainBands * g.n_out;
    const long long item = item0 + threadIdx.x;
    if (item < total) {
        const float4 st = reinterpret_cast<const float4*>(b.gstat)[item];
        const float4 pv = reinterpret_cast<const float4*>(b.gprev)[item];
        if (st.x < 0.05f || st.z < 1e-6f || pv.y < 1e-6f) {                      // hfr | target | saved last level
            Curve out;
            out.n = 0; out.pad = 0;
            for (int i = 0; i < 7; i++) { out.level[i] = 0; out.loc[i] = 0; }
            *curve_slot(g, b, item) = out;
        } else {
            list[atomicAdd(&cnt, 1)] = (int)threadIdx.x;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kCurveBlock) curve_item(g, b, item0 + list[i]);
}

void launch_gain_curve(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const long long total = (long long)g.S * g.C * kGainBands * g.n_out;
    ATDE_LAUNCH(at3_curve_kernel, (unsigned)((total + kCurveBlock - 1) / kCurveBlock), kCurveBlock, 0, st, g, b);
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

// One WARP walks a RUN of consecutive frames of one (stream, channel); the band samples come in by bulk
// asynchronous copies (1-D TMA) and the spectra leave the same way:
//   ring    per band two 256-sample slots: the previous frame's samples (already gain-modulated) and this frame's.
//           While frame f is transformed, lane 0 has the copy engine fetch frame f+1 into the slots frame f-1
//           just vacated (cp.async.bulk + mbarrier); nothing is read twice and the warp never waits on a load it
//           issued itself.  (The first frame of a stream takes the carried, already windowed half instead.)
//   energy  bands that touch a gain curve: CalcGainEnergyScale's sequential sums (atrac3denc.cpp:189-216) from the ring,
//           then the frame's samples are divided by the curve IN PLACE — the next frame needs them modulated too —
//           and the overlap scale is handed on in a register.
//   fold    all four MDCT-512 at once, 8 lanes per band, 16 complex points per lane: window (atrac3denc.cpp:39-49),
//           fold + pre-twiddle (mdct.h:56-76) straight from the ring into kissfft's gather order; radix-2 m=1 and
//           radix-4 m=2 on registers
//   exchange through the warp's tile (padded 1 per 16: conflict-free both ways)
//   finish  radix-4 m=8 and m=32 on registers, post-twiddle (mdct.h:92-101), odd bands reversed
//           (atrac3denc.cpp:53-55), spectrum staged in the tile and stored with cp.async.bulk
// Every butterfly keeps kissfft's operation order (kissfft_dev.cuh), so the regrouping is exact.
constexpr int kMdWarps = 4;
#ifndef ATDE_MD_RUN
#define ATDE_MD_RUN 16                               // (the CPU-emulation build uses 3 so that tiny batches cross run boundaries)
#endif
constexpr int kMdRun = ATDE_MD_RUN;                  // frames a warp walks before it takes the next item
constexpr int kMdRingBand = 2 * 256 + 8;             // floats per band: two slots; +8 spreads the bands over the banks
constexpr int kMdRingFloats = 4 * kMdRingBand;       // 2080
constexpr int kMdXchBand = 136;                      // cpx per band in the exchange layout (128 + 1 per 16)
constexpr int kMdOutBand = 2 * kMdXchBand;           // floats per band in the output layout (272: 16-byte aligned rows)
constexpr int kMdTileFloats = 4 * kMdOutBand;        // 1088 (exchange / output / energy-term scratch)
constexpr int kMdTermStride = 132;                   // floats between energy-term rows

struct MdWarpShared {
    float ring[kMdRingFloats];
    float tile[kMdTileFloats];
    Curve cv[4][2];                                  // [band][0] previous frame, [1] this frame
    float esum[4][8];                                // energy sums 0..4; [5] = divisor of the stored half (d0), [6] = overlap scale
    mbar_t bar;
    unsigned long long pad_;
};
struct MdBlockShared {
    float sincos[256];
    float win[256];
    float ones[256];
    cpx tw[128];
    MdWarpShared w[kMdWarps];
};

__global__ void __launch_bounds__(kMdWarps * 32, 4) at3_mdct_kernel(Geometry g, Buffers b)
{
    ATDE_DYN_SMEM(smem_raw);
    MdBlockShared& sh = *reinterpret_cast<MdBlockShared*>(smem_raw);

    const DevTables* __restrict__ T = b.tab;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ATDE_PAR_FOR(i, 256) { sh.sincos[i] = T->sincos512[i]; sh.win[i] = T->encode_window[i]; sh.ones[i] = 1.0f; }
    ATDE_PAR_FOR(i, 128) sh.tw[i] = T->tw128[i];
    MdWarpShared& ws = sh.w[warp];
    if (lane == 0) mbar_init(&ws.bar, 1);
    async_proxy_fence();
    __syncthreads();

    float* const ring = ws.ring;
    float* const tile = ws.tile;
    unsigned phase = 0;                                  // parity of the barrier phase the next wait looks at
    const int runs = (g.n_out + kMdRun - 1) / kMdRun;
    const long long n_items = (long long)g.S * g.C * runs;
    const int bnd = lane >> 3, L = lane & 7;             // fold / FFT ownership: band, lane within the band

    for (long long item = (long long)blockIdx.x * kMdWarps + warp; item < n_items; item += (long long)gridDim.x * kMdWarps) {
        const int run = (int)(item % runs);
        const size_t sc = (size_t)(item / runs);
        const int c = (int)(sc % g.C), s = (int)(sc / g.C);
        const int f0 = run * kMdRun, f1 = min(f0 + kMdRun, g.n_out);
        const float* const brow = b.bands + sc * 4 * g.BL + 128;          // band q, frame f: brow + q BL + 256 f
        const bool gain = !g.no_gain;

        // ---- prologue: stored half + first frame into the ring ----
        if (lane == 0) bulk_store_wait_read();           // the previous item's spectrum has left the tile
        async_proxy_fence();
        __syncwarp();                                    // ... and nobody still reads the ring
        if (lane == 0) {
            mbar_expect_tx(&ws.bar, (f0 > 0 ? 8u : 4u) * 1024u);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (f0 > 0) bulk_g2s(ring + q * kMdRingBand, brow + (size_t)q * g.BL + 256 * (size_t)(f0 - 1), 1024, &ws.bar);
                bulk_g2s(ring + q * kMdRingBand + 256, brow + (size_t)q * g.BL + 256 * (size_t)f0, 1024, &ws.bar);
            }
        }
        if (f0 == 0) {                                   // the carried half is windowed + modulated already
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int h = 0; h < 2; h++)
                    *reinterpret_cast<float4*>(ring + q * kMdRingBand + 4 * (lane + 32 * h)) =
                        *reinterpret_cast<const float4*>(b.prevhalf + (sc * 4 + q) * 256 + 4 * (lane + 32 * h));
        }
        // gain curves: lane q < 3 carries band q's curve of the frame in work and of the one before
        Curve cc, pc;
        cc.n = 0; pc.n = 0;
        if (gain && lane < kGainBands) {
            const Curve* cp = b.curves + (sc * 4 + lane) * g.n_out;
            cc = cp[f0];
            if (f0 > 0) pc = cp[f0 - 1];
        }
        float pos_scale = 1.0f;                          // lane q < 4: PrevOverlapGainScale[channel][q]
        if (f0 == 0 && lane < 4) pos_scale = b.next_scale[sc * 4 + lane];
        mbar_wait(&ws.bar, phase);
        phase ^= 1u;
        if (f0 > 0) {
            // the stored half of a run that starts inside the stream: the previous frame's samples, modulated by ITS
            // curve; its NextOverlapScale = SafeEnergyScale(sum (y w)^2, sum (ym w)^2) (atrac3denc.cpp:205-216)
            if (lane < 4) ws.cv[lane][0] = pc;
            __syncwarp();
#pragma unroll 1
            for (int q = 0; q < kGainBands; q++) {
                const Curve& pcv = ws.cv[q][0];
                if (pcv.n == 0) continue;
                float* P = ring + q * kMdRingBand;
                float* lv = tile + 2 * kMdTermStride;
                __syncwarp();
                curve_levels_warp(T, pcv, lane, lv);
                __syncwarp();
                float a = 0.0f;
#pragma unroll 1
                for (int half = 0; half < 256; half += 128) {
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const int j = lane + 32 * t, i = half + j;
                        const float y = P[i];
                        const float ym = __fdiv_rn(y, lv[i]);
                        const float wi = sh.win[i];
                        float v;
                        v = fmul(y, wi);  tile[0 * kMdTermStride + j] = fmul(v, v);
                        v = fmul(ym, wi); tile[1 * kMdTermStride + j] = fmul(v, v);
                        P[i] = ym;
                    }
                    __syncwarp();
                    if (lane < 2) {
#pragma unroll 4
                        for (int j = 0; j < 128; j += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(&tile[lane * kMdTermStride + j]);
                            a = fadd(a, t4.x); a = fadd(a, t4.y); a = fadd(a, t4.z); a = fadd(a, t4.w);
                        }
                    }
                    __syncwarp();
                }
                const float e_org = __shfl_sync(0xffffffffu, a, 0);
                const float e_mod = __shfl_sync(0xffffffffu, a, 1);
                if (lane == q) {
                    float ps = safe_energy_scale(e_org, e_mod);
                    const float inf = __int_as_float(0x7f800000);
                    if (!(fabsf(ps) < inf) || ps <= 0.0f) ps = 1.0f;
                    pos_scale = ps;
                }
            }
        }

        int par = 0;                                     // ring slot of the stored half
#pragma unroll 1
        for (int f = f0; f < f1; f++) {
            const size_t unit = ((size_t)s * g.n_out + f) * g.C + c;
            if (f != f0) {
                mbar_wait(&ws.bar, phase);               // frame f's samples have landed
                phase ^= 1u;
            }
            Curve nc;                                    // next frame's curve, fetched early
            nc.n = 0;
            if (gain && lane < kGainBands && f + 1 < f1) nc = b.curves[(sc * 4 + lane) * g.n_out + f + 1];
            if (lane == 0) bulk_store_wait_read();       // the previous frame's spectrum has left the tile
            if (lane < 4) {
                ws.cv[lane][1] = cc;
                ws.esum[lane][6] = pos_scale;
                ws.esum[lane][5] = cc.n ? T->gain_level[cc.level[0]] : 1.0f;
            }
            __syncwarp();
            // A band whose own curve is empty and whose overlap scale is 1 has every energy scale exactly 1.0
            // (SafeEnergyScale(x, x)); the sequential energy sums only run for bands that touch a curve.
            unsigned trivial = 0, curved = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (ws.cv[q][1].n) curved |= 1u << q;
                else if (ws.esum[q][6] == 1.0f) trivial |= 1u << q;
            }
            const float* wprev = (f == 0) ? sh.ones : sh.win;      // window of the stored half (applied already at f == 0)
            // ---- CalcGainEnergyScale's sequential sums, bands that touch a curve only:
            //  0 prevStored  1 curOriginal  2 curModulated  3 nextOriginal  4 nextModulated
#pragma unroll 1
            for (int q = 0; q < 4; q++) {
                if ((trivial >> q) & 1u) continue;
                const bool has = (curved >> q) & 1u;
                const float* P = ring + q * kMdRingBand + 256 * par;
                float* X = ring + q * kMdRingBand + 256 * (par ^ 1);
                float* lv = tile + 5 * kMdTermStride;
                __syncwarp();
                if (has) curve_levels_warp(T, ws.cv[q][1], lane, lv);
                __syncwarp();
                float a = 0.0f;                                    // lane < 5: running sum of term `lane`
#pragma unroll 1
                for (int half = 0; half < 256; half += 128) {
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const int j = lane + 32 * t, i = half + j;
                        const float x = X[i];
                        const float xm = has ? __fdiv_rn(x, lv[i]) : x;
                        const float wi = sh.win[i], wr = sh.win[255 - i];
                        const float prev = fmul(wprev[i], P[i]);
                        float v;
                        tile[0 * kMdTermStride + j] = fmul(prev, prev);
                        v = fmul(x, wr);  tile[1 * kMdTermStride + j] = fmul(v, v);
                        v = fmul(xm, wr); tile[2 * kMdTermStride + j] = fmul(v, v);
                        v = fmul(x, wi);  tile[3 * kMdTermStride + j] = fmul(v, v);
                        v = fmul(xm, wi); tile[4 * kMdTermStride + j] = fmul(v, v);
                        if (has) X[i] = xm;                        // gain modulation (gain_processor.h:87-121), in place
                    }
                    __syncwarp();
                    if (lane < 5) {
#pragma unroll 4
                        for (int j = 0; j < 128; j += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(&tile[lane * kMdTermStride + j]);
                            a = fadd(a, t4.x); a = fadd(a, t4.y); a = fadd(a, t4.z); a = fadd(a, t4.w);
                        }
                    }
                    __syncwarp();
                }
                if (lane < 5) ws.esum[q][lane] = a;
            }
            __syncwarp();
            if (lane < 4) {
                const int q = lane;
                const bool cur_empty = cc.n == 0;
                float sc0 = 1.0f, sc1 = 1.0f, sc2 = 1.0f, sc3 = 1.0f;
                if (!((trivial >> q) & 1u)) {
                    const float* es = ws.esum[q];
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
                *reinterpret_cast<float4*>(&b.gscale[unit * 16 + q * 4]) = make_float4(sc0, sc1, sc2, sc3);
                if (f == g.n_out - 1) b.next_scale_out[sc * 4 + q] = sc3;
                // PrevOverlapGainScale for the next frame (atrac3denc.cpp:781-786)
                float ps = sc3;
                const float inf = __int_as_float(0x7f800000);
                if (!(fabsf(ps) < inf) || ps <= 0.0f) ps = 1.0f;
                pos_scale = ps;
            }
            if (f == g.n_out - 1) {                                // the half this frame leaves behind (next batch)
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i0 = 4 * (lane + 32 * h);
                        const float4 x4 = *reinterpret_cast<const float4*>(ring + q * kMdRingBand + 256 * (par ^ 1) + i0);
                        const float4 w4 = *reinterpret_cast<const float4*>(&sh.win[i0]);
                        *reinterpret_cast<float4*>(b.prevhalf_out + (sc * 4 + q) * 256 + i0) =
                            make_float4(fmul(w4.x, x4.x), fmul(w4.y, x4.y), fmul(w4.z, x4.z), fmul(w4.w, x4.w));
                    }
            }
            __syncwarp();                                          // ring modulated, tile free, esum[.][5] visible
            const float* wfold = wprev;
            if (curved) {
                // some band has a curve: its stored half is divided by the curve's first level (atrac3denc.cpp:46-47).
                // Done here, on the windowed values and in place (the slot is dead after this frame), so that the fold
                // below carries no division; bands without a curve divide by 1.0f, which changes nothing.
#pragma unroll 1
                for (int q = 0; q < 4; q++) {
                    float* P = ring + q * kMdRingBand + 256 * par;
                    const float d0 = ws.esum[q][5];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i0 = 4 * (lane + 32 * h);
                        const float4 p4 = *reinterpret_cast<const float4*>(P + i0);
                        const float4 w4 = *reinterpret_cast<const float4*>(wprev + i0);
                        *reinterpret_cast<float4*>(P + i0) =
                            make_float4(__fdiv_rn(fmul(w4.x, p4.x), d0), __fdiv_rn(fmul(w4.y, p4.y), d0),
                                        __fdiv_rn(fmul(w4.z, p4.z), d0), __fdiv_rn(fmul(w4.w, p4.w), d0));
                    }
                }
                wfold = sh.ones;
                __syncwarp();
            }

            // ---- fold + pre-twiddle + the two innermost FFT stages: lane (bnd, L) owns gather blocks 2L, 2L + 1 ----
            cpx e[4][4];                                           // after the exchange: element L + 8a + 32b
            {
                const float* P = ring + bnd * kMdRingBand + 256 * par;
                const float* X = ring + bnd * kMdRingBand + 256 * (par ^ 1);
                cpx* xch = reinterpret_cast<cpx*>(tile) + bnd * kMdXchBand;
#pragma unroll 1
                for (int h = 0; h < 2; h++) {                      // (rolled: the kernel is instruction-fetch sensitive)
                    const int blk = 2 * L + h;
                    // slot = 8 blk + j of the 4x4x4x2 digit reversal: i = d0 + 4 d1 + 16 d2 + 64 d3
                    const int ibase = (blk >> 2) + 4 * (blk & 3);
                    cpx v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int n = 2 * (ibase + 16 * (j >> 1) + 64 * (j & 1));   // N = 512, n4 = 128, n34 = 384, n54 = 640
                        float r0, i0;
                        if ((j & 1) == 0) {                        // n < 128: r0 = in[383-n] + in[384+n], i0 = in[128+n] - in[127-n]
                            const float wa = sh.win[128 + n], wb = sh.win[127 - n];
                            const float p1 = fmul(wfold[128 + n], P[128 + n]), p2 = fmul(wfold[127 - n], P[127 - n]);
                            r0 = fadd(fmul(wa, X[127 - n]), fmul(wb, X[128 + n]));
                            i0 = fsub(p1, p2);
                        } else {                                   // r0 = in[383-n] - in[n-128], i0 = in[128+n] + in[639-n]
                            const float wa = sh.win[383 - n], wb = sh.win[n - 128];
                            const float p1 = fmul(wfold[383 - n], P[383 - n]), p2 = fmul(wfold[n - 128], P[n - 128]);
                            r0 = fsub(p1, p2);
                            i0 = fadd(fmul(wa, X[n - 128]), fmul(wb, X[383 - n]));
                        }
                        const float2 cs = *reinterpret_cast<const float2*>(&sh.sincos[n]);
                        v[j].r = fadd(fmul(r0, cs.x), fmul(i0, cs.y));
                        v[j].i = fsub(fmul(i0, cs.x), fmul(r0, cs.y));
                    }
                    // radix-2, m = 1 (fstride 64): pairs (2q, 2q+1), twiddle tw[0]
#pragma unroll
                    for (int q = 0; q < 4; q++) kf_bfly2(v[2 * q], v[2 * q + 1], sh.tw[0]);
                    // radix-4, m = 2 (fstride 16): elements kk + 2q, twiddles tw[16 kk q]
#pragma unroll
                    for (int kk = 0; kk < 2; kk++)
                        kf_bfly4<false>(v[kk], v[kk + 2], v[kk + 4], v[kk + 6], sh.tw[16 * kk], sh.tw[32 * kk], sh.tw[48 * kk]);
                    // element 8 blk + j at padded index 17 L + 8 h + j (one pad per 16 elements)
#pragma unroll
                    for (int j = 0; j < 8; j++) xch[17 * L + 8 * h + j] = v[j];
                }
                async_proxy_fence();
                __syncwarp();                                      // every lane is done with the ring's stored-half slots
                if (lane == 0 && f + 1 < f1) {                     // fetch frame f+1 into them while this one is transformed
                    mbar_expect_tx(&ws.bar, 4096u);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        bulk_g2s(ring + q * kMdRingBand + 256 * par, brow + (size_t)q * g.BL + 256 * (size_t)(f + 1), 1024, &ws.bar);
                }
                // element L + 8a + 32b at padded index L + 8a + (a >> 1) + 34 b
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) e[a][q] = xch[L + 8 * a + (a >> 1) + 34 * q];
            }
            // ---- radix-4 m = 8 (fstride 4) over a, radix-4 m = 32 (fstride 1) over b ----
            {
                const cpx t1 = sh.tw[4 * L], t2 = sh.tw[8 * L], t3 = sh.tw[12 * L];
#pragma unroll
                for (int q = 0; q < 4; q++) kf_bfly4<false>(e[0][q], e[1][q], e[2][q], e[3][q], t1, t2, t3);
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int v = L + 8 * a;
                kf_bfly4<false>(e[a][0], e[a][1], e[a][2], e[a][3], sh.tw[v], sh.tw[2 * v], sh.tw[3 * v]);
            }
            __syncwarp();                                          // exchange reads done: the tile becomes the output stage
            // post-twiddle (mdct.h:92-101); odd bands reversed (atrac3denc.cpp:53-55)
            {
                float* sp = tile + bnd * kMdOutBand;
                const bool odd = bnd & 1;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int n = 2 * (L + 8 * a + 32 * q);
                        const cpx z = e[a][q];
                        const float2 cs = *reinterpret_cast<const float2*>(&sh.sincos[n]);
                        const float va = fsub(fmul(-z.r, cs.x), fmul(z.i, cs.y));
                        const float vb = fadd(fmul(-z.r, cs.y), fmul(z.i, cs.x));
                        sp[odd ? 255 - n : n] = va;
                        sp[odd ? n : 255 - n] = vb;
                    }
            }
            async_proxy_fence();                                   // generic-proxy stores -> visible to the copy engine
            __syncwarp();
            if (lane == 0) {
                float* outp = b.specs + unit * 1024;
#pragma unroll
                for (int q = 0; q < 4; q++) bulk_s2g(outp + 256 * q, tile + q * kMdOutBand, 1024);
                bulk_store_commit();
            }
            par ^= 1;
            pc = cc;
            cc = nc;
        }
    }
    if (lane == 0) bulk_store_wait_all();                          // the last spectrum is in global memory before the block retires
}

void launch_mdct(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const int runs = (g.n_out + kMdRun - 1) / kMdRun;
    const long long n_items = (long long)g.S * g.C * runs;
    long long blocks = (n_items + kMdWarps - 1) / kMdWarps;
    if (blocks > 148 * 4) blocks = 148 * 4;               // persistent: every block resident (4 per SM), warps stride over items
    cudaFuncSetAttribute(at3_mdct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MdBlockShared));
    ATDE_LAUNCH(at3_mdct_kernel, (unsigned)blocks, kMdWarps * 32, sizeof(MdBlockShared), st, g, b);
}

// =====================================================================================
// K6: loudness recurrence, one thread per stream
// =====================================================================================
__global__ void at3_loudness_kernel(Geometry g, Buffers b)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.S) return;
    float L = b.loud_state[s];
    const bool two = g.C == 2 && !g.js;
    // eight frames' terms are fetched before the recurrence runs over them (the stores of one frame would otherwise stand
    // between the loads of the next: one exposed memory latency per frame on a thread that has nothing else to do)
    for (int f0 = 0; f0 < g.n_out; f0 += 8) {
        float t0[8], t1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const size_t o = ((size_t)s * g.n_out + f0 + q) * g.C;
            const bool in = f0 + q < g.n_out;
            t0[q] = in ? b.chloud[o] : 0.0f;
            t1[q] = in && two ? b.chloud[o + 1] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (f0 + q < g.n_out) {
                if (two) {
                    const float sum = fadd(t0[q], t1[q]);
                    L = __double2float_rn(__dadd_rn(__dmul_rn(0.98, (double)L), __dmul_rn(0.01, (double)sum)));
                } else {
                    L = __double2float_rn(__dadd_rn(__dmul_rn(0.98, (double)L), __dmul_rn(0.02, (double)t0[q])));
                }
                b.loud[(size_t)s * g.n_out + f0 + q] = L;
            }
        }
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
