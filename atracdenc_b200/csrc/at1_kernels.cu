// at1_kernels.cu — ATRAC1 encode hot path on sm_100a.
//
// Replaces, for thousands of frames at once (reference: dcherednik/atracdenc):
//   K1 at1_analysis_kernel  Atrac1AnalysisFilterBank::Analysis (src/atrac/at1/atrac1_qmf.h:37-43)
//                           TQmf::Analysis                      (src/qmf/qmf.h:47-64)
//                           TTransientDetector::Detect          (src/transient_detector.cpp:52-93)
//                           TAtrac1MDCT::Mdct + TMDCT           (src/atrac1denc.cpp:70-102, src/lib/mdct/mdct.h:51-104)
//                           kiss_fft                            (src/lib/fft/kissfft_impl/kiss_fft.c)
//                           per-channel loudness term           (src/atrac1denc.cpp:235-240)
//   K4 at1_loudness_kernel  TrackLoudness recurrence            (src/atrac1denc.cpp:243-247)
//   K5 at1_pack_kernel      TScaler::ScaleFrame                 (src/atrac/atrac_scale.cpp:141-188)
//                           TAt1BitAlloc::Write                 (src/atrac/at1/atrac1_bitalloc.cpp:80-409)
//                           TBitStreamEncoder bisection         (src/lib/bs_encode/encode.cpp:57-129)
//                           TBitStream::Write                   (src/lib/bitstream/bitstream.cpp:40-63)
//
// Frames of a stream are NOT processed sequentially.  Every piece of encoder state the reference
// carries from frame to frame is either a finite function of the previous PCM frame (QMF history,
// hi-band delay line, MDCT overlap tail, transient-detector history) — recomputed here from a
// 304-sample input halo — or the scalar loudness recurrence, which K4 scans per stream.
// All arithmetic is un-fused IEEE fp32 in the reference's operation order (bit-exact contract).
#include "at1_kernels.cuh"
#include "kissfft_dev.cuh"
#include "glibc_math.cuh"
#include "qmf_dev.cuh"

namespace atde {
namespace at1 {

__constant__ __align__(8) float c_qmf1p[48];   // QmfWindow (qmf.cpp:36-45) as tap pairs (W[2i+1], W[2i]), see qmf_dev.cuh

void upload_qmf_window(const float w[48])
{
    float p[48];
    for (int i = 0; i < 24; i++) { p[2 * i] = w[2 * i + 1]; p[2 * i + 1] = w[2 * i]; }
    cudaMemcpyToSymbol(c_qmf1p, p, 48 * sizeof(float));
}

// ---- BFU geometry, atrac1.h:89-104 ----
__device__ const unsigned char kSpecsPerBlock[kMaxBfus] = {
    8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 7, 7, 7, 7, 9, 9, 9, 9, 10, 10, 10, 10,
    12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20, 20, 20, 20, 20, 20};
__device__ const unsigned short kSpecsStartLong[kMaxBfus] = {
    0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116, 122,
    128, 134, 140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226, 236, 246,
    256, 268, 280, 292, 304, 316, 328, 340, 352, 372, 392, 412, 432, 452, 472, 492};
__device__ const unsigned short kSpecsStartShort[kMaxBfus] = {
    0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90, 122,
    128, 160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182, 214, 246,
    256, 288, 320, 352, 384, 416, 448, 480, 268, 300, 332, 364, 396, 428, 460, 492};
// fixed allocation tables and boost mask, atrac1_bitalloc.cpp:37-67
__device__ const unsigned char kFixLong[kMaxBfus] = {
    7, 7, 7, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4,
    4, 4, 3, 3, 3, 3, 3, 3, 2, 1, 1, 1, 1, 0, 0, 0};
__device__ const unsigned char kFixShort[kMaxBfus] = {
    6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
    4, 4, 4, 4, 4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0};

ATDE_D int bfu_band(int b) { return b < 20 ? 0 : (b < 36 ? 1 : 2); }
ATDE_D int bfu_amount(int idx) { return idx == 0 ? 20 : (idx == 1 ? 28 : 28 + 4 * (idx - 1)); }   // {20,28,32,...,52}

// =====================================================================================
// K1: QMF tree + transient detection + windowed MDCT + loudness term
// =====================================================================================
//
// One block = kTile consecutive frames of one stream; channels are processed one after another
// through the same shared-memory tile.  Index conventions (t0 = first frame of the tile):
//   x[k]     = input sample  512*t0 - 304 + k          k in [0, kNX)
//   s1lo/hi[j] = stage-1 QMF output n = 256*t0 - 128 + j  j in [0, kNS1)   (taps read x[2j+48-2i], x[2j+49-2i])
//   lo/mi[j] = stage-2 output    m = 128*t0 - 40 + j    j in [0, kNS2)   (taps read s1lo[2j+48-2i], s1lo[2j+49-2i])
//   hi band sample q (delayed 39, atrac1_qmf.h:26,38-42) = s1hi[q - 256*t0 + 89]
constexpr int kNX = kTile * 512 + 304;
constexpr int kNS1 = kTile * 256 + 128;
constexpr int kNS2 = kTile * 128 + 40;
constexpr int kNFilt = kTile * 512 + 48;         // HPF output incl. one 16-sample halo block per band
constexpr int kNEner = kTile * 32 + 3;

struct BandView {
    const float* base;   // sample k of tile-frame 0 is base[k]; frame tl adds tl*size
    int size;            // 128 / 128 / 256
};

// Stage-1 tasks of 9 output pairs (kNS1 = 1152 = 128 x 9: one task per thread), stage-2 tasks of 5 (111 tasks cover
// kNS2 = 552 and three scratch outputs).  Odd task sizes keep the 64-bit loads conflict-free (qmf_dev.cuh).
constexpr int kQ1 = 9, kQ2 = 5;
constexpr int kAnaThreads = 128;
static_assert(kNS1 == kAnaThreads * kQ1, "stage-1 tasks fill the block exactly");
constexpr int kNS2T = ((kNS2 + kQ2 - 1) / kQ2) * kQ2;           // 555

// ---- the MDCT of one frame by one warp ---------------------------------------------------------------------------
// kissfft stage with compile-time geometry (kissfft_dev.cuh: kf_stage4 divides by a run-time m)
template <int M, int FSTRIDE>
ATDE_D void stage4c(cpx* buf, const cpx* __restrict__ tw, int v)
{
    const int g = v / M, k = v % M;
    cpx* F = buf + g * 4 * M + k;
    cpx f0 = F[0], f1 = F[M], f2 = F[2 * M], f3 = F[3 * M];
    kf_bfly4<false>(f0, f1, f2, f3, tw[k * FSTRIDE], tw[2 * k * FSTRIDE], tw[3 * k * FSTRIDE]);
    F[0] = f0; F[M] = f1; F[2 * M] = f2; F[3 * M] = f3;
}

// Window + fold + pre-twiddle of one band (mdct.h:56-76 on TAtrac1MDCT::Mdct's input, atrac1denc.cpp:80-90), written in
// kissfft's gather order: one long block, or (SHORT) the band's 4 / 8 short blocks of 32 samples, 16 slots each.  The
// four input positions of an output slot are a property of the slot: the host tabulates them (DevTables::fold128 /
// fold256 / fold64: four 10-bit positions t into the windowed stretch, 1023 = outside of it, where the reference's buffer
// holds 0; bits 40.. = the slot's index i = n / 2), so the kernel neither permutes nor range-tests.
template <bool SHORT>
ATDE_D void fold_band(const unsigned long long* __restrict__ tab, int nslots, const float* __restrict__ wt,
                      const float* fr, const float* __restrict__ cs, int q4, cpx* dst, int lane)
{
    for (int s = lane; s < nslots; s += 32) {
        const unsigned long long e = tab[SHORT ? (s & 15) : s];
        const float* const frb = SHORT ? fr + 32 * (s >> 4) : fr;  // short block kb windows the 64 samples around it
        const unsigned el = (unsigned)e, eh = (unsigned)(e >> 32);
        auto term = [&](unsigned t) {
            const bool ok = t != 1023u;
            const int tc = ok ? (int)t : 0;
            const float v = fmul(wt[tc], frb[tc - 32]);          // the stretch starts 32 samples before the block
            return ok ? v : 0.0f;
        };
        const float a0 = term(el & 1023u), a1 = term((el >> 10) & 1023u), b0 = term((el >> 20) & 1023u),
                    b1 = term(((el >> 30) | (eh << 2)) & 1023u);
        const int i = (int)((eh >> 8) & 127u);
        float r0, i0;
        if (i < q4) { r0 = fadd(a0, a1); i0 = fsub(b0, b1); }
        else        { r0 = fsub(a0, a1); i0 = fadd(b0, b1); }
        const float2 c = *reinterpret_cast<const float2*>(cs + 2 * i);
        cpx X;
        X.r = fadd(fmul(r0, c.x), fmul(i0, c.y));
        X.i = fsub(fmul(i0, c.x), fmul(r0, c.y));
        dst[s] = X;
    }
}

// post-twiddle (mdct.h:92-101) of NB blocks of N2 spectral lines each, the band reversal of the mid / hi bands and the
// level fix of the hi band's short blocks (atrac1denc.cpp:92-98)
template <int N2, int NB, bool REV, bool TWICE>
ATDE_D void post_band(const cpx* src, const float* __restrict__ cs, float* out, int lane)
{
#pragma unroll
    for (int s = lane; s < NB * N2 / 2; s += 32) {
        const int kb = s / (N2 / 2), i = s % (N2 / 2), n = 2 * i;
        const cpx z = src[s];
        const float2 c = *reinterpret_cast<const float2*>(cs + n);
        float va = fsub(fmul(-z.r, c.x), fmul(z.i, c.y));
        float vb = fadd(fmul(-z.r, c.y), fmul(z.i, c.x));
        if (TWICE) { va = fmul(va, 2.0f); vb = fmul(vb, 2.0f); }
        const int pa = REV ? N2 - 1 - n : n, pb = REV ? n : N2 - 1 - n;
        out[N2 * kb + pa] = va;
        out[N2 * kb + pb] = vb;
    }
}

// One frame: window mask `msk`, bit b = band b (low, mid, hi) takes short blocks.  Every branch is uniform over the warp.
//   low / mid   long: FFT64 = 4x4x4;  short: four FFT16 = 4x4.  The first two stages are THE SAME butterflies on the same
//               slots for both (stage m=1 has no twiddle; stage m=4 pairs slots 16g+k+{0,4,8,12} with twiddles
//               tw64[4k q] == tw16[k q], bit for bit: the phase -2 pi 4k / 64 is the phase -2 pi k / 16), so the two
//               bands share the warp, 16 butterflies each, whatever their window
//   hi          long: FFT128 = 4x4x4x2 (radix-2 innermost);  short: eight FFT16
ATDE_D void mdct_frame(const DevTables* __restrict__ T, const float* f0, const float* f1, const float* f2, unsigned msk,
                       cpx* fft, float* spf, int lane)
{
    const bool s0 = msk & 1u, s1 = msk & 2u, s2 = msk & 4u;
    if (s0) fold_band<true>(T->fold64, 64, T->win_short, f0, T->sincos64, 8, fft, lane);
    else fold_band<false>(T->fold128, 64, T->win_long128, f0, T->sincos256, 32, fft, lane);
    if (s1) fold_band<true>(T->fold64, 64, T->win_short, f1, T->sincos64, 8, fft + 64, lane);
    else fold_band<false>(T->fold128, 64, T->win_long128, f1, T->sincos256, 32, fft + 64, lane);
    if (s2) fold_band<true>(T->fold64, 128, T->win_short, f2, T->sincos64, 8, fft + 128, lane);
    else fold_band<false>(T->fold256, 128, T->win_long256, f2, T->sincos512, 64, fft + 128, lane);
    __syncwarp();
    cpx* const b01 = fft + 64 * (lane >> 4);
    cpx* const b2 = fft + 128;
    const int v = lane & 15;
    stage4c<1, 16>(b01, T->tw64, v);
    if (s2) {
        stage4c<1, 4>(b2 + 16 * (lane >> 2), T->tw16, lane & 3);
    } else {
        kf_stage2(b2, T->tw128, lane, 1, 64);
        kf_stage2(b2, T->tw128, lane + 32, 1, 64);
    }
    __syncwarp();
    stage4c<4, 4>(b01, T->tw64, v);
    if (s2) stage4c<4, 1>(b2 + 16 * (lane >> 2), T->tw16, lane & 3);
    else stage4c<2, 16>(b2, T->tw128, lane);
    __syncwarp();
    if (!((lane >> 4) ? s1 : s0)) stage4c<16, 1>(b01, T->tw64, v);
    if (!s2) {
        stage4c<8, 4>(b2, T->tw128, lane);
        __syncwarp();
        stage4c<32, 1>(b2, T->tw128, lane);
    }
    __syncwarp();
    if (s0) post_band<32, 4, false, false>(fft, T->sincos64, spf, lane);
    else post_band<128, 1, false, false>(fft, T->sincos256, spf, lane);
    if (s1) post_band<32, 4, true, false>(fft + 64, T->sincos64, spf + 128, lane);
    else post_band<128, 1, true, false>(fft + 64, T->sincos256, spf + 128, lane);
    if (s2) post_band<32, 8, true, true>(fft + 128, T->sincos64, spf + 256, lane);
    else post_band<256, 1, true, false>(fft + 128, T->sincos512, spf + 256, lane);
}

__global__ void __launch_bounds__(kAnaThreads, 6) at1_analysis_kernel(AnalysisParams p)
{
    __shared__ __align__(16) float x[kNX + 8];      // input tile; later HPF output; later FFT buffer
    __shared__ __align__(16) float s1lo[kNS1 + 8];  // (+8: the three scratch outputs of stage 2 read past kNS1)
    __shared__ __align__(16) float s1hi_raw[kNS1 + 4];
    float* const s1hi = s1hi_raw + 3;               // hi-band sample 0 of the tile sits at s1hi + 89: 16-byte aligned this way
    __shared__ __align__(16) float lo[kNS2T + 1];
    __shared__ __align__(16) float mi[kNS2T + 1];
    __shared__ __align__(16) float sp[kTile * 512];
    __shared__ float ener[kNEner];
    __shared__ __align__(8) float cw[48];
    __shared__ unsigned char smask[kTile];

    const int s = blockIdx.y;
    const int c = blockIdx.z;                              // one block per channel: the channels never meet
    const int t0 = blockIdx.x * kTile;
    const int C = p.C, F = p.F;
    const DevTables* __restrict__ T = p.tab;
    const bool fresh = !(p.started && p.started[s]);       // stream starts at frame 0 of this batch
    f32x2 one;
    one.x = p.one; one.y = p.one;

    ATDE_PAR_FOR(i, 48) cw[i] = c_qmf1p[i];
    ATDE_PAR_FOR(i, 8) { x[kNX + i] = 0.0f; s1lo[kNS1 + i] = 0.0f; }

    {
        // ---- load input tile with halo ----
        const float* __restrict__ pcm = p.pcm + (size_t)s * F * 512 * C + c;
        const int nbase = 512 * t0 - 304;
        if (nbase >= 0 && nbase + kNX <= F * 512) {
            // interior tile: every load independent and in flight at once
            constexpr int kIters = (kNX + kAnaThreads - 1) / kAnaThreads;
            float v[kIters];
#pragma unroll
            for (int it = 0; it < kIters; it++) {
                const int k = threadIdx.x + kAnaThreads * it;
                v[it] = k < kNX ? pcm[(size_t)(nbase + k) * C] : 0.0f;
            }
#pragma unroll
            for (int it = 0; it < kIters; it++) {
                const int k = threadIdx.x + kAnaThreads * it;
                if (k < kNX) x[k] = v[it];
            }
        } else
        ATDE_PAR_FOR(k, kNX) {
            const int n = 512 * t0 - 304 + k;
            float v = 0.0f;
            if (n >= 0) {
                if (n < F * 512) v = pcm[(size_t)n * C];
            } else if (p.hist && !fresh) {
                v = p.hist[((size_t)s * 512 + (512 + n)) * C + c];
            }
            x[k] = v;
        }
        __syncthreads();
        // ---- QMF stage 1: full band -> (low+mid, hi); thread t owns outputs 9t .. 9t+8 ----
        {
            float l[kQ1], u[kQ1];
            qmf_task64<kQ1>(x, cw, kQ1 * (int)threadIdx.x, one, l, u);
#pragma unroll
            for (int r = 0; r < kQ1; r++) { s1lo[kQ1 * threadIdx.x + r] = l[r]; s1hi[kQ1 * threadIdx.x + r] = u[r]; }
        }
        __syncthreads();
        // ---- QMF stage 2: (low+mid) -> (low, mid) ----
        if (threadIdx.x < kNS2T / kQ2) {
            float l[kQ2], u[kQ2];
            qmf_task64<kQ2>(s1lo, cw, kQ2 * (int)threadIdx.x, one, l, u);
#pragma unroll
            for (int r = 0; r < kQ2; r++) { lo[kQ2 * threadIdx.x + r] = l[r]; mi[kQ2 * threadIdx.x + r] = u[r]; }
        }
        __syncthreads();

        // sample 0 of tile-frame 0 of band b: lo + 40, mi + 40, s1hi + 89
        float* filt = x;                                          // x is dead from here on
        static_assert(kNFilt % 8 == 0 && (16 + kTile * 128) % 8 == 0, "HPF tasks of 8 outputs");

        if (p.window_auto) {
            // ---- 21-tap HPF (transient_detector.cpp:52-70) over every band sample of the tile
            //      plus the last 16 samples of the frame before it (for LastEnergy) ----
            //      A task = 8 consecutive outputs of one band-frame (frame sizes and the 16-sample halo are multiples of 8,
            //      so a task never straddles frames): its 29 input samples are loaded and sign-flipped once
            //      (InvertSpectr negates the even samples of the mid / hi bands; the window starts at an even sample,
            //      so the flipped positions are compile-time constants), then the 8 x 21 taps run on registers.
            constexpr int kHpfTasks = kNFilt / 8;                         // 66 + 66 + 130
            for (int task = threadIdx.x; task < kHpfTasks; task += blockDim.x) {
                const int nb01 = (16 + kTile * 128) / 8;
                const int b = task < nb01 ? 0 : (task < 2 * nb01 ? 1 : 2);
                const int q0 = 8 * (task - b * nb01);                     // first output of the task within the band's region
                const int size = (b == 2) ? 256 : 128;
                int tl, i0;
                if (q0 < 16) { tl = -1; i0 = size - 16 + q0; }
                else { tl = (q0 - 16) / size; i0 = (q0 - 16) - tl * size; }
                const float* base = b == 0 ? lo + 40 : (b == 1 ? mi + 40 : s1hi + 89);
                const float* fr = base + tl * size + i0 - 20;            // w[t] = frame sample i0 - 20 + t
                const unsigned flip = b != 0 ? 0x80000000u : 0u;
                // (fr is 16-byte aligned: i0 - 20 is a multiple of 4 and so are the band origins; eight 16-byte loads, the
                //  32-byte stride between the tasks of neighbouring lanes costs two wavefronts instead of eight)
                float w[32];
#pragma unroll
                for (int t4 = 0; t4 < 8; t4++) {
                    const float4 v = reinterpret_cast<const float4*>(fr)[t4];
                    w[4 * t4 + 0] = __uint_as_float(__float_as_uint(v.x) ^ flip);    // i0 - 20 is even: even t <-> even sample
                    w[4 * t4 + 1] = v.y;
                    w[4 * t4 + 2] = __uint_as_float(__float_as_uint(v.z) ^ flip);
                    w[4 * t4 + 3] = v.w;
                }
                // inBuf[w] = frame sample (w - 20); inBuf[size + 20] is never written by the reference => 0
                const float c0 = -8.65163e-18 * 2.0, c1 = -0.00851586 * 2.0, c2 = -6.74764e-18 * 2.0,
                            c3 = 0.0209036 * 2.0, c4 = -3.36639e-17 * 2.0, c5 = -0.0438162 * 2.0,
                            c6 = -1.54175e-17 * 2.0, c7 = 0.0931738 * 2.0, c8 = -5.52212e-17 * 2.0,
                            c9 = -0.313819 * 2.0;
                const float cf[10] = {c0, c1, c2, c3, c4, c5, c6, c7, c8, c9};
                const bool frame_end = i0 + 8 == size;                    // output 7 is the frame's last sample
                float o[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    float sacc = w[r + 10];                               // sample i - 10
                    float s2 = 0.0f;
#pragma unroll
                    for (int j = 0; j < 9; j += 2) {
                        // right = sample i + 1 - j; for j == 0 at the frame's last sample it lies past the frame: 0
                        const float right = (j == 0 && r == 7 && frame_end) ? 0.0f : w[r + 21 - j];
                        sacc = fadd(sacc, fmul(cf[j], fadd(w[r + j], right)));
                        s2 = fadd(s2, fmul(cf[j + 1], fadd(w[r + j + 1], w[r + 20 - j])));
                    }
                    o[r] = __fdiv_rn(fadd(sacc, s2), 2.0f);
                }
                float4* dst = reinterpret_cast<float4*>(filt + 8 * task);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
            __syncthreads();
            // ---- RMS (dB-ish) of each 16-sample short block (transient_detector.cpp:33-40,81) ----
            ATDE_PAR_FOR(e, kNEner) {
                const float* f = filt + 16 * e;
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < 16; i++)
                    acc = fadd(acc, fmul(f[i], f[i]));
                acc = __fdiv_rn(acc, 16.0f);
                const float rms = __fsqrt_rn(acc);
                ener[e] = __double2float_rn(__dmul_rn(19.0, (double)g_log10f(rms)));
            }
            __syncthreads();
            // ---- window decision per frame (transient_detector.cpp:78-91, atrac1denc.cpp:214-229) ----
            ATDE_PAR_FOR(tl, kTile) {
                unsigned m = 0;
                for (int b = 0; b < 3; b++) {
                    const int nshort = (b == 2) ? 16 : 8;
                    const int rb = (b == 0) ? 0 : (b == 1 ? 1 + kTile * 8 : 2 + kTile * 16);
                    float prev = ener[rb + tl * nshort];
                    if (t0 + tl == 0 && fresh)
                        prev = 0.0f;                                     // LastEnergy initial value
                    bool trans = false;
                    for (int k = 0; k < nshort; k++) {
                        const float cur = ener[rb + 1 + tl * nshort + k];
                        if (fsub(cur, prev) > 16.0f) trans = true;
                        if (fsub(prev, cur) > 20.0f) trans = true;
                        prev = cur;
                    }
                    if (trans) m |= 1u << b;
                }
                smask[tl] = (unsigned char)m;
            }
        } else {
            ATDE_PAR_FOR(tl, kTile) smask[tl] = (unsigned char)p.window_mask;
        }
        __syncthreads();

        // ---- MDCT (TAtrac1MDCT::Mdct, atrac1denc.cpp:70-102): ONE WARP PER FRAME of the tile, no block barrier until the
        //      spectra are stored; long and short blocks are chosen per band, uniformly over the warp ----
        {
            cpx* const fft_all = reinterpret_cast<cpx*>(x);        // kTile*256 complex (x is dead: the energies are taken)
            const int lane = threadIdx.x & 31;
            for (int tl = threadIdx.x >> 5; tl < kTile; tl += kAnaThreads / 32) {
                const unsigned msk = smask[tl];
                const float* f0 = lo + 40 + tl * 128;
                const float* f1 = mi + 40 + tl * 128;
                const float* f2 = s1hi + 89 + tl * 256;
                mdct_frame(T, f0, f1, f2, msk, fft_all + tl * 256, sp + tl * 512, lane);
            }
        }
        // ---- store spectra and masks (the loudness term has its own kernel: it is one sequential sum per frame).
        //      A frame's 512 spectral lines are one 2 KB row in shared and in global memory: bulk asynchronous stores ----
        async_proxy_fence();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int tl = 0; tl < kTile; tl++)
                if (t0 + tl < F)
                    bulk_s2g(p.specs + (((size_t)s * F + t0 + tl) * C + c) * 512, sp + tl * 512, 2048u);
            bulk_store_commit();
        }
        ATDE_PAR_FOR(tl, kTile) {
            if (t0 + tl < F) p.masks[((size_t)s * F + t0 + tl) * C + c] = smask[tl];
        }
        if (threadIdx.x == 0) bulk_store_wait_read();                 // shared memory is released when the block retires
    }
}

// Loudness term of every channel-frame (atrac1denc.cpp:235-240): l += spec^2 * curve over the 512 lines in order — one
// sequential chain per channel-frame, so ONE LANE per channel-frame; the spectra are staged through shared memory in
// 32x32 tiles to keep the global reads coalesced.
constexpr int kLoudWarps = 4;
__global__ void __launch_bounds__(kLoudWarps * 32) at1_loudterm_kernel(AnalysisParams p)
{
    __shared__ float tile_all[kLoudWarps][32][33];
    const DevTables* __restrict__ T = p.tab;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long total = (long long)p.S * p.F * p.C;
    const long long unit0 = ((long long)blockIdx.x * kLoudWarps + wib) * 32;
    if (unit0 >= total) return;
    float (*tile)[33] = tile_all[wib];
    const long long mine = unit0 + lane;
    const bool live = mine < total;
    const int nrows = (int)min(32LL, total - unit0);
    float l = 0.0f;
    for (int t = 0; t < 16; t++) {
        for (int r = 0; r < nrows; r++)
            tile[r][lane] = p.specs[(size_t)(unit0 + r) * 512 + 32 * t + lane];
        __syncwarp();
        if (live) {
#pragma unroll 8
            for (int k = 0; k < 32; k++) {
                const float v = tile[lane][k];
                l = fadd(l, fmul(fmul(v, v), T->loud_curve[32 * t + k]));
            }
        }
        __syncwarp();
    }
    if (live) p.chloud[mine] = l;
}

void launch_analysis(const AnalysisParams& p, cudaStream_t st)
{
    dim3 grid((p.F + kTile - 1) / kTile, p.S, p.C);
    ATDE_LAUNCH(at1_analysis_kernel, grid, kAnaThreads, 0, st, p);
    const long long total = (long long)p.S * p.F * p.C;
    const long long per_block = kLoudWarps * 32;
    ATDE_LAUNCH(at1_loudterm_kernel, (unsigned)((total + per_block - 1) / per_block), kLoudWarps * 32, 0, st, p);
}

// =====================================================================================
// K4: loudness recurrence, one thread per stream (atrac1denc.cpp:243-247, atrac_psy_common.h:46-54)
// =====================================================================================
__global__ void at1_loudness_kernel(LoudnessParams p)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.S) return;
    float L = p.loud_in ? p.loud_in[s] : kLoudFactor;
    // eight frames' masks and terms are fetched before the recurrence runs over them (the store of one frame would
    // otherwise stand between the loads of the next: one exposed memory latency per frame)
    const bool two = p.C == 2;
    for (int f0 = 0; f0 < p.F; f0 += 8) {
        unsigned m0[8], m1[8];
        float t0[8], t1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const size_t o = ((size_t)s * p.F + f0 + q) * p.C;
            const bool in = f0 + q < p.F;
            m0[q] = in ? p.masks[o] : 0u;
            m1[q] = in && two ? p.masks[o + 1] : 0u;
            t0[q] = in ? p.chloud[o] : 0.0f;
            t1[q] = in && two ? p.chloud[o + 1] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (f0 + q < p.F) {
                if (two && m0[q] == 0 && m1[q] == 0) {
                    const float sum = fadd(t0[q], t1[q]);
                    L = __double2float_rn(__dadd_rn(__dmul_rn(0.98, (double)L), __dmul_rn(0.01, (double)sum)));
                } else if (m0[q] == 0) {
                    L = __double2float_rn(__dadd_rn(__dmul_rn(0.98, (double)L), __dmul_rn(0.02, (double)t0[q])));
                }
                p.loud[(size_t)s * p.F + f0 + q] = L;
            }
        }
    }
}

void launch_loudness(const LoudnessParams& p, cudaStream_t st)
{
    ATDE_LAUNCH(at1_loudness_kernel, (p.S + 63) / 64, 64, 0, st, p);
}

// =====================================================================================
// K5: scale + bit allocation search + boost + bitstream, one warp per channel-frame
// =====================================================================================
constexpr int kPackWarps = 4;

ATDE_D unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

// MSB-first bit field into a zeroed big-endian word array (value already masked to n bits)
ATDE_D void put_bits(unsigned* words, int pos, int n, unsigned val)
{
    const int w = pos >> 5, off = pos & 31;
    const int room = 32 - off;
    if (n <= room) {
        atomicOr(&words[w], val << (room - n));
    } else {
        atomicOr(&words[w], val >> (n - room));
        atomicOr(&words[w + 1], val << (32 - (n - room)));
    }
}

// TBitStream::Write buffer growth (bitstream.cpp:43-52): only the resulting vector size matters here
ATDE_D void bs_grow(int& size, int& used, int n)
{
    const int bits_left = size * 8 - used;
    const int bits_req = n - bits_left;
    const int overlap = used & 7;
    if (overlap || bits_req >= 0)
        size += bits_req / 8 + (overlap ? 2 : 1);
    used += n;
}

__global__ void __launch_bounds__(kPackWarps * 32) at1_pack_kernel(PackParams p)
{
    __shared__ __align__(16) float sv_all[kPackWarps][512];
    __shared__ unsigned words_all[kPackWarps][56];
    __shared__ unsigned char wl_all[kPackWarps][kMaxBfus + 4];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * kPackWarps + wib;
    const long long total = (long long)p.S * p.F * p.C;
    if (unit >= total) return;                                    // whole warp leaves together
    float* sv = sv_all[wib];
    unsigned* words = words_all[wib];
    unsigned char* wls = wl_all[wib];
    const DevTables* __restrict__ T = p.tab;

    const long long sf = unit / p.C;                              // s*F + f
    const unsigned mask = p.masks[unit];
    const float loud = __fdiv_rn(p.loud[sf], kLoudFactor);        // Loudness / LoudFactor, atrac1denc.cpp:250

    const float* __restrict__ src = p.specs + (size_t)unit * 512;
    {   // the channel-frame's 2 KB: four 16-byte loads per lane, all in flight before the first store
        const float4* __restrict__ src4 = reinterpret_cast<const float4*>(src);
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) v[r] = src4[lane + 32 * r];
#pragma unroll
        for (int r = 0; r < 4; r++) reinterpret_cast<float4*>(sv)[lane + 32 * r] = v[r];
    }
    for (int i = lane; i < 56; i += 32) words[i] = 0;
    __syncwarp();

    // ---- TScaler::Scale for this lane's BFUs (b = lane, lane + 32) ----
    int sfi[2] = {0, 0};
    float energy[2] = {0.0f, 0.0f};
    for (int h = 0; h < 2; h++) {
        const int b = lane + 32 * h;
        if (b < kMaxBfus) {
            const bool shrt = (mask >> bfu_band(b)) & 1;
            const int start = shrt ? kSpecsStartShort[b] : kSpecsStartLong[b];
            const int len = kSpecsPerBlock[b];
            float mx = 0.0f;
            for (int j = 0; j < len; j++) {
                const float a = fabsf(sv[start + j]);
                if (a > mx) mx = a;
            }
            if (mx > 1.0f) mx = 1.0f;
            int lo_i = 0, hi_i = 63;                               // lower_bound over ScaleTable
            while (lo_i < hi_i) {
                const int mid = (lo_i + hi_i) >> 1;
                if (T->scale_table[mid] < mx) lo_i = mid + 1; else hi_i = mid;
            }
            const float scale = T->scale_table[lo_i];
            sfi[h] = lo_i;
            float en = 0.0f;
            for (int j = 0; j < len; j++) {
                const float xin = sv[start + j];
                float v = __fdiv_rn(xin, scale);
                en = fadd(en, fmul(xin, xin));
                if (fabsf(v) >= 1.0f)
                    v = (v > 0.0f) ? 0.99999f : -0.99999f;
                sv[start + j] = v;
            }
            energy[h] = en;
            if (p.tap_sfi) p.tap_sfi[(size_t)unit * kMaxBfus + b] = (unsigned char)lo_i;
        }
    }
    __syncwarp();

    // ---- CalcLowToMidTilt ingredients (integer sums are exact in fp32) ----
    const unsigned sum_low = warp_sum(lane < 20 ? (unsigned)sfi[0] : 0u);
    const unsigned sum_m28 = warp_sum((lane >= 20 && lane < 28) ? (unsigned)sfi[0] : 0u);
    const unsigned sum_m32 = warp_sum(lane >= 20 ? (unsigned)sfi[0] : 0u);
    const unsigned sum_m36 = sum_m32 + warp_sum(lane < 4 ? (unsigned)sfi[1] : 0u);

    // CalcBitsAllocation (atrac1_bitalloc.cpp:185-203) evaluates spread * (sfi / 3.2f) + (1 - spread) * fix - shift + bias
    // left to right: the first two terms and the audibility test do not depend on the search variable and are taken once
    // per BFU (the search runs ~25 steps per channel-frame)
    float wl_base[2];
    bool audible[2], shrt[2];
    int len[2];
    for (int h = 0; h < 2; h++) {
        const int b = lane + 32 * h;
        const int bb = b < kMaxBfus ? b : kMaxBfus - 1;
        shrt[h] = (mask >> bfu_band(bb)) & 1;
        const float fix = shrt[h] ? (float)kFixShort[bb] : (float)kFixLong[bb];
        const float spread = 0.4f;
        audible[h] = shrt[h] || !(energy[h] < fmul(T->ath_long[bb], loud));
        wl_base[h] = fadd(fmul(spread, __fdiv_rn((float)sfi[h], 3.2f)), fmul(fsub(1.0f, spread), fix));
        len[h] = kSpecsPerBlock[bb];
    }

    // ---- TBitStreamEncoder over {TConfigure, TBfuAlloc} (encode.cpp:100-129) ----
    int bfu_idx = p.bfu_idx_const ? p.bfu_idx_const - 1 : 7;
    const bool auto_bfu = !p.bfu_idx_const;
    unsigned wl[2] = {0, 0};
    int nbfu;
    unsigned bits_used;
    for (;;) {
        nbfu = bfu_amount(bfu_idx);
        const unsigned target = 212 * 8 - 3 - 32 - 2 - 3 - nbfu * 10;
        // band bias from the low/mid tilt (atrac1_bitalloc.cpp:147-161,176-179)
        float tilt = 0.0f;
        if (nbfu > 20) {
            const int n_mid = (nbfu < 36 ? nbfu : 36) - 20;
            const unsigned sm = nbfu == 28 ? sum_m28 : (nbfu == 32 ? sum_m32 : sum_m36);
            tilt = fsub(__fdiv_rn((float)sum_low, 20.0f), __fdiv_rn((float)sm, (float)n_mid));
        }
        const float mid_bias = fminf(1.5f, fmul(0.3f, fmaxf(0.0f, fsub(tilt, 7.0f))));
        const float bias_of_band[3] = {0.0f, mid_bias, fmul(mid_bias, 0.5f)};
        float bias[2];
        for (int h = 0; h < 2; h++) {
            const int b = lane + 32 * h;
            bias[h] = bias_of_band[bfu_band(b < kMaxBfus ? b : kMaxBfus - 1)];
        }
        float mn = -3.0f, mx = 15.0f, cur = 0.0f, last = 15.0f;       // Start(target, -3, 15)
        for (;;) {
            const bool exhausted = mx <= mn;
            const float shift = exhausted ? last
                                          : __double2float_rn(__dmul_rn((double)fadd(mx, mn), 0.5));   // (max + min) / 2.0: halving is exact
            cur = shift;
            unsigned my_bits = 0;
            for (int h = 0; h < 2; h++) {
                const int b = lane + 32 * h;
                wl[h] = 0;
                if (b < nbfu && audible[h]) {
                    const int tmp = __float2int_rz(fadd(fsub(wl_base[h], shift), bias[h]));
                    wl[h] = tmp > 16 ? 16u : (tmp < 2 ? 0u : (unsigned)tmp);
                    my_bits += (unsigned)len[h] * wl[h];
                }
            }
            bits_used = warp_sum(my_bits);
            if (exhausted)
                break;
            if (bits_used < target) { last = cur; mx = fsub(cur, 0.01f); }
            else if (bits_used > target) { mn = fadd(cur, 0.01f); }
            else break;
        }
        if (auto_bfu) {
            // GetMaxUsedBfuId (atrac1_bitalloc.cpp:207-230): amount-table segment of the highest non-zero BFU
            const unsigned nz0 = __ballot_sync(0xffffffffu, wl[0] != 0);
            const unsigned nz1 = __ballot_sync(0xffffffffu, wl[1] != 0);
            const int top = nz1 ? 32 + (31 - __clz((int)nz1)) : (nz0 ? 31 - __clz((int)nz0) : -1);
            int used = 0;
            while (used < 7 && bfu_amount(used) <= top) used++;
            if (top < 0) used = 0;
            if (used < bfu_idx) { bfu_idx--; continue; }
        }
        break;
    }

    // ---- TBitsBooster::ApplyBoost (atrac1_bitalloc.cpp:80-114): serial, tiny ----
    if (lane + 0 < kMaxBfus) wls[lane] = (unsigned char)wl[0];
    if (lane + 32 < kMaxBfus) wls[lane + 32] = (unsigned char)wl[1];
    __syncwarp();
    if (lane == 0) {
        const unsigned target = 212 * 8 - 3 - 32 - 2 - 3 - nbfu * 10;
        // multimap (bits -> position) in key order, insertion order within a key
        const unsigned char bpos[12] = {18, 19, 20, 21, 22, 32, 33, 34, 35, 36, 37, 38};
        const unsigned char bbits[12] = {6, 6, 6, 6, 6, 10, 10, 10, 10, 12, 12, 12};
        unsigned surplus = target - bits_used;
        const unsigned key = surplus > 12u ? 12u : surplus;
        int max_it = 0;
        while (max_it < 12 && bbits[max_it] <= key) max_it++;
        if (max_it != 0) {
            while (surplus >= 6u) {
                bool done = true;
                for (int it = 0; it < max_it; ++it) {
                    const unsigned cb = bbits[it];
                    const int cp = bpos[it];
                    if (cp >= nbfu) break;
                    const unsigned w = wls[cp];
                    if (w == 16u) continue;
                    const unsigned per = w ? 1u : 2u;
                    if (w == 0u && cb * 2 > surplus) continue;
                    if (cb * per > surplus) continue;
                    wls[cp] = (unsigned char)(w + per);
                    surplus -= cb * per;
                    done = false;
                }
                if (done) break;
            }
        }
    }
    __syncwarp();

    // ---- TBfuAlloc::Dump (atrac1_bitalloc.cpp:279-327) ----
    unsigned mbits[2];
    for (int h = 0; h < 2; h++) {
        const int b = lane + 32 * h;
        wl[h] = (b < nbfu) ? wls[b] : 0;
        mbits[h] = (wl[h] >= 2) ? (unsigned)len[h] * wl[h] : 0u;
    }
    // exclusive prefix of mantissa bits in BFU order (b = 0..31 on h=0, 32..51 on h=1)
    unsigned inc0 = mbits[0], inc1 = mbits[1];
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned a = __shfl_up_sync(0xffffffffu, inc0, d);
        const unsigned c2 = __shfl_up_sync(0xffffffffu, inc1, d);
        if (lane >= d) { inc0 += a; inc1 += c2; }
    }
    const unsigned tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const unsigned start_bits[2] = {inc0 - mbits[0], tot0 + inc1 - mbits[1]};

    if (lane == 0) {
        const unsigned lc0 = (mask & 1) ? 2 : 0, lc1 = (mask & 2) ? 2 : 0, lc2 = (mask & 4) ? 3 : 0;
        const unsigned hdr = ((2 - lc0) << 14) | ((2 - lc1) << 12) | ((3 - lc2) << 10) | ((unsigned)bfu_idx << 5);
        put_bits(words, 0, 16, hdr);
    }
    const int mant_base = 16 + 10 * nbfu;
    for (int h = 0; h < 2; h++) {
        const int b = lane + 32 * h;
        if (b < nbfu) {
            put_bits(words, 16 + 4 * b, 4, wl[h] ? wl[h] - 1 : 0);
            put_bits(words, 16 + 4 * nbfu + 6 * b, 6, (unsigned)sfi[h]);
            if (wl[h] >= 2) {
                const int start = shrt[h] ? kSpecsStartShort[b] : kSpecsStartLong[b];
                const float mult = (float)((1 << (wl[h] - 1)) - 1);
                int pos = mant_base + (int)start_bits[h];
                const unsigned vm = (1u << wl[h]) - 1u;
                for (int j = 0; j < len[h]; j++) {
                    const int q = __float2int_rn(fmul(sv[start + j], mult));
                    put_bits(words, pos, (int)wl[h], (unsigned)q & vm);
                    pos += (int)wl[h];
                }
            }
        }
        if (p.tap_wl && b < kMaxBfus)
            p.tap_wl[(size_t)unit * kMaxBfus + b] = (b < nbfu) ? (unsigned char)wl[h] : 0xff;
    }
    __syncwarp();
    // big-endian words -> byte stream, 53 words = 212 bytes
    unsigned* __restrict__ dst = reinterpret_cast<unsigned*>(p.out + (size_t)unit * kUnitBytes);
    for (int i = lane; i < 53; i += 32)
        dst[i] = __byte_perm(words[i], 0, 0x0123);

    if (p.sizes && lane == 0) {
        int size = 0, used = 0;
        bs_grow(size, used, 2); bs_grow(size, used, 2); bs_grow(size, used, 2); bs_grow(size, used, 2);
        bs_grow(size, used, 3); bs_grow(size, used, 2); bs_grow(size, used, 3);
        for (int b = 0; b < nbfu; b++) bs_grow(size, used, 4);
        for (int b = 0; b < nbfu; b++) bs_grow(size, used, 6);
        for (int b = 0; b < nbfu; b++) {
            const int w = wls[b];
            if (w < 2) continue;
            const int cnt = kSpecsPerBlock[b];
            for (int j = 0; j < cnt; j++) bs_grow(size, used, w);
        }
        bs_grow(size, used, 8); bs_grow(size, used, 8); bs_grow(size, used, 8);
        p.sizes[unit] = size;
    }
}

void launch_pack(const PackParams& p, cudaStream_t st)
{
    const long long total = (long long)p.S * p.F * p.C;
    const unsigned grid = (unsigned)((total + kPackWarps - 1) / kPackWarps);
    ATDE_LAUNCH(at1_pack_kernel, grid, kPackWarps * 32, 0, st, p);
}

// =====================================================================================
// carry: keep the last PCM frame and loudness of every stream for the next batch
// =====================================================================================
__global__ void at1_carry_kernel(CarryParams p)
{
    const int s = blockIdx.x;
    const int n = 512 * p.C;
    const float* src = p.pcm + ((size_t)s * p.F + (p.F - 1)) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        p.hist[(size_t)s * n + i] = src[i];
    if (threadIdx.x == 0) {
        p.loud_state[s] = p.loud[(size_t)s * p.F + p.F - 1];
        p.started[s] = 1;
    }
}

void launch_carry(const CarryParams& p, cudaStream_t st)
{
    ATDE_LAUNCH(at1_carry_kernel, p.S, 256, 0, st, p);
}

} // namespace at1
} // namespace atde
