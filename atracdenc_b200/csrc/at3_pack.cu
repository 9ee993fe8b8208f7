// at3_pack.cu — ATRAC3 encode hot path on sm_100a, coding half.
//
// Replaces (reference: dcherednik/atracdenc):
//   K7 at3_scale_tonal_kernel  CalcSpectralFlatnessPerBfu (src/atrac/atrac_psy_common.cpp:158-199),
//                              ExtractTonalComponents / MapTonalComponents (src/atrac3denc.cpp:581-662),
//                              TScaler::Scale / ScaleFrame (src/atrac/atrac_scale.cpp:141-188)
//   K8 at3_alloc_pack_kernel   TAtrac3BitStreamWriter::WriteSoundUnit (src/atrac/at3/atrac3_bitstream.cpp:759-847):
//                              CalcMSBytesShift (:741-757), TConfigure/TAlloc under TBitStreamEncoder's bisection
//                              (:587-682, src/lib/bs_encode/encode.cpp:57-129), CalcBitsAllocation (:272-336),
//                              QuantMantisas (src/atrac/atrac_scale.cpp:40-130) behind the (bfu, wordlen) cache
//                              (:151-227), ConsiderEnergyErr (:241-257), GroupTonalComponents /
//                              EncodeTonalComponents (:338-524), EncodeSpecs with CLCEnc / VLCEnc (:92-149,526-565),
//                              TBitStream::Write (src/lib/bitstream/bitstream.cpp:40-63)
//
// One warp owns one channel of one frame; lane i owns BFU i (ATRAC3 has exactly 32 BFUs).
#include "at3_kernels.cuh"
#include "glibc_math.cuh"
#include "stdsort_dev.cuh"

// The allocation / pack kernel is bound by instruction fetch, not by arithmetic (16 warps per SM, each in a
// different phase of a long program): loops the compiler would unroll are kept rolled so that the program
// executed per frame stays small (183 -> 141 ms per 10^6 frames for the first three, see profiles/README.md).
#ifdef ATDE_PACK_UNROLLED
#define ATDE_ROLLED
#else
#define ATDE_ROLLED _Pragma("unroll 1")
#endif

namespace atde {
namespace at3 {

// atrac3.h:83-105
__device__ const unsigned short kBlockStart[33] = {
    0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176,
    192, 224, 256, 288, 320, 352, 384, 416, 448, 480, 512, 576, 640, 704, 768, 896, 1024};
// atrac3_bitstream.cpp:44-49
__device__ const unsigned char kFixedAlloc[32] = {
    6, 6, 5, 4, 4, 4, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 0, 0, 0};
__device__ const float kMaxQuant[8] = {0.0f, 1.5f, 2.5f, 3.5f, 4.5f, 7.5f, 15.5f, 31.5f};   // atrac3.h:79-82
__device__ const unsigned char kClcLen[8] = {0, 4, 3, 3, 4, 4, 5, 6};                       // atrac3.h:106

// Huffman tables, atrac3.h:110-176.  Selector s uses table s-1 of {1, 2, 3, 1, 5, 6, 7}.
__device__ const unsigned char kHuffOff[8] = {0, 0, 9, 14, 0, 21, 36, 67};                  // by selector
__device__ const unsigned char kHuffCode[130] = {
    // table 1 (selectors 1 and 4)
    0x0, 0x4, 0x5, 0xC, 0xD, 0x1C, 0x1D, 0x1E, 0x1F,
    // table 2
    0x0, 0x4, 0x5, 0x6, 0x7,
    // table 3
    0x0, 0x4, 0x5, 0xC, 0xD, 0xE, 0xF,
    // table 5
    0x0, 0x2, 0x3, 0x8, 0x9, 0xA, 0xB, 0x1C, 0x1D, 0x3C, 0x3D, 0x3E, 0x3F, 0xC, 0xD,
    // table 6
    0x0, 0x2, 0x3, 0x4, 0x5, 0x6, 0x7, 0x14, 0x15, 0x16, 0x17, 0x18, 0x19,
    0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x3B, 0x78, 0x79, 0x7A, 0x7B, 0x7C, 0x7D, 0x7E, 0x7F, 0x8, 0x9,
    // table 7
    0x0, 0x8, 0x9, 0xA, 0xB, 0xC, 0xD, 0xE, 0xF, 0x10, 0x11,
    0x24, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x2B, 0x2C, 0x2D, 0x2E, 0x2F, 0x30, 0x31, 0x32, 0x33,
    0x68, 0x69, 0x6A, 0x6B, 0x6C, 0x6D, 0x6E, 0x6F, 0x70, 0x71, 0x72, 0x73, 0x74, 0x75,
    0xEC, 0xED, 0xEE, 0xEF, 0xF0, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA, 0xFB, 0xFC, 0xFD, 0xFE, 0xFF,
    0x2, 0x3};
__device__ const unsigned char kHuffBits[130] = {
    1, 3, 3, 4, 4, 5, 5, 5, 5,
    1, 3, 3, 3, 3,
    1, 3, 3, 4, 4, 4, 4,
    2, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6, 4, 4,
    3, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7, 4, 4,
    3, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
    6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
    7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7,
    8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8,
    4, 4};
__device__ const unsigned char kVlcPairIdx[9] = {8, 4, 7, 2, 0, 1, 6, 3, 5};                // atrac3.h:227-233
__device__ const unsigned char kClcIdx[4] = {2, 3, 0, 1};                                   // atrac3.h:221-226

ATDE_D int huff_index(int m)
{
    int h = (m < 0) ? (((-m) << 1) | 1) : (m << 1);
    if (h) h -= 1;
    return h;
}

ATDE_D int bfu_band(int i) { return i >= 30 ? 3 : (i >= 26 ? 2 : (i >= 18 ? 1 : 0)); }     // BlocksPerBand {0,18,26,30,32}

// BFU of spectral line i (atrac3.h:83-105)
ATDE_D int elem_bfu_s(int i)
{
    if (i < 64) return i >> 3;
    if (i < 192) return 8 + ((i - 64) >> 4);
    if (i < 512) return 16 + ((i - 192) >> 5);
    if (i < 768) return 26 + ((i - 512) >> 6);
    return 30 + ((i - 768) >> 7);
}

ATDE_D unsigned warp_sum_u(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

// lower_bound over ScaleTable + TScaler::Scale (atrac_scale.cpp:141-172); writes the scaled values
// back over v[0..len) and returns sfi / energy.
ATDE_D int scale_block(const DevTables* T, float* v, int len, float& energy)
{
    float mx = 0.0f;
    for (int j = 0; j < len; j++) {
        const float a = fabsf(v[j]);
        if (a > mx) mx = a;
    }
    if (mx > 1.0f) mx = 1.0f;
    int lo = 0, hi = 63;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (T->scale_table[mid] < mx) lo = mid + 1; else hi = mid;
    }
    const float scale = T->scale_table[lo];
    float en = 0.0f;
    for (int j = 0; j < len; j++) {
        const float x = v[j];
        float q = __fdiv_rn(x, scale);
        en = fadd(en, fmul(x, x));
        if (fabsf(q) >= 1.0f) q = (q > 0.0f) ? 0.99999f : -0.99999f;
        v[j] = q;
    }
    energy = en;
    return lo;
}

// TScaler::Scale split for the frame path: (1) per BFU, sequential — max |x|, scale factor index, energy
// (len is a multiple of 8; 8 lines per trip keep the shared-memory loads off the dependent chain);
// (2) per line, parallel — x / scale with the +-0.99999 clip.
ATDE_D int scale_analyse(const DevTables* T, const float* v, int len, float& energy, float& scale)
{
    float mx = 0.0f, en = 0.0f;
    for (int j = 0; j < len; j += 8) {
        const float4 a = *reinterpret_cast<const float4*>(v + j);
        const float4 c = *reinterpret_cast<const float4*>(v + j + 4);
        const float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            mx = fmaxf(mx, fabsf(x[k]));
            en = fadd(en, fmul(x[k], x[k]));
        }
    }
    if (mx > 1.0f) mx = 1.0f;
    int lo = 0, hi = 63;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (T->scale_table[mid] < mx) lo = mid + 1; else hi = mid;
    }
    scale = T->scale_table[lo];
    energy = en;
    return lo;
}

// =====================================================================================
// K7: tonal extraction + scaling, one warp per channel-frame
// =====================================================================================
constexpr int kScaleWarps = 4;


__global__ void __launch_bounds__(kScaleWarps * 32) at3_scale_tonal_kernel(Geometry g, Buffers b)
{
    __shared__ __align__(16) float sv_all[kScaleWarps][1024];
    __shared__ float run_val[kScaleWarps][32][5];
    __shared__ short run_start[kScaleWarps][32];
    __shared__ signed char run_len[kScaleWarps][32];
    __shared__ double lg_all[kScaleWarps][64];          // logs of the lines of the BFU under exact evaluation

    const DevTables* __restrict__ T = b.tab;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * kScaleWarps + wib;
    const long long total = (long long)g.S * g.n_out * g.C;
    if (unit >= total) return;
    float* sv = sv_all[wib];
    float* gsp = b.specs + (size_t)unit * 1024;
    {   // eight 16-byte loads per lane, all in flight before the first store
        const float4* g4 = reinterpret_cast<const float4*>(gsp);
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = g4[lane + 32 * r];
#pragma unroll
        for (int r = 0; r < 8; r++) reinterpret_cast<float4*>(sv)[lane + 32 * r] = v[r];
    }
    run_len[wib][lane] = 0;
    __syncwarp();

    const int start = kBlockStart[lane], len = kBlockStart[lane + 1] - kBlockStart[lane];
    TonalList* tl = b.tonal + unit;
    if (!g.no_tonal) {
        // CalcSpectralFlatnessPerBfu: geometric / arithmetic mean of the line energies, in double, compared with 0.01
        // (atrac3denc.cpp:586-590) — the value itself is used for nothing else.  Nearly every BFU is far from tonal
        // (flatness ~0.5), so each BFU is first screened with plain fp32 and the hardware log2 (error of the screen
        // ~1e-3 relative, generously): above 0.0125 the exact value cannot be below 0.01.  Only the BFUs the screen
        // cannot clear take the reference's arithmetic: per-line double logs (glibc's log, line-parallel over the warp),
        // the two sequential double sums, glibc's exp.
        const double floor_d = (double)1e-12f;
        bool maybe = false;
        if (lane >= 8 && lane < 29) {
            float aa = 0.0f, al = 0.0f;
            for (int i = 0; i < len; i += 4) {
                const float4 q = *reinterpret_cast<const float4*>(sv + start + i);
                const float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float ef = x[k] * x[k];
                    aa += ef;
                    al += __log2f(fmaxf(ef, 1e-12f));
                }
            }
            const float A = aa / (float)len;
            // near the energy floor (where the reference returns 1.0) or not clearly flat: the exact path decides
            maybe = !(A > 4e-12f) || !(exp2f(al / (float)len) > 0.0125f * A);
        }
        double* lg = lg_all[wib];
        float flat = 1.0f;
        for (unsigned mm = __ballot_sync(0xffffffffu, maybe); mm; mm &= mm - 1) {
            const int bf = __ffs((int)mm) - 1;
            const int b0 = kBlockStart[bf], bl = kBlockStart[bf + 1] - b0;
            __syncwarp();
            for (int i = lane; i < bl; i += 32) {
                const float ef = fmul(sv[b0 + i], sv[b0 + i]);
                const double e = (double)fmaxf(0.0f, ef);
                lg[i] = g_log(e > floor_d ? e : floor_d);
            }
            __syncwarp();
            if (lane == bf) {
                double arith = 0.0, mean_log = 0.0;
                for (int i = 0; i < len; i += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(sv + start + i);
                    const float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float ef = fmul(x[k], x[k]);
                        const double e = (double)fmaxf(0.0f, ef);
                        arith = __dadd_rn(arith, e);
                        mean_log = __dadd_rn(mean_log, lg[i + k]);
                    }
                }
                arith = __ddiv_rn(arith, (double)len);
                mean_log = __ddiv_rn(mean_log, (double)len);
                if (!(arith <= floor_d)) {
                    const double ratio = __ddiv_rn(g_exp(mean_log), arith);
                    const double cl = ratio < 0.0 ? 0.0 : (ratio > 1.0 ? 1.0 : ratio);   // min(1, max(0, ratio))
                    flat = __double2float_rn(cl);
                }
            }
        }
        __syncwarp();
        if (lane >= 8 && lane < 29) {
            if (flat < 0.01f) {
                // ExtractTonalComponents: best run of <= 5 lines by summed magnitude
                const int max_len = min(5, len);
                float best = -1.0f;
                int best_start = start, best_len = 1;
                for (int st = start; st < start + len; st++) {
                    const int ml = min(max_len, start + len - st);
                    float score = 0.0f;
                    for (int l = 1; l <= ml; l++) {
                        score = fadd(score, fabsf(sv[st + l - 1]));
                        if (score > best) { best = score; best_start = st; best_len = l; }
                    }
                }
                if (best > 0.0f) {
                    run_start[wib][lane] = (short)best_start;
                    run_len[wib][lane] = (signed char)best_len;
                    for (int n = 0; n < best_len; n++) run_val[wib][lane][n] = sv[best_start + n];
                }
            }
        }
        __syncwarp();
        for (int n = 0; n < run_len[wib][lane]; n++) sv[run_start[wib][lane] + n] = 0.0f;
        __syncwarp();
        if (lane == 0) {
            // MapTonalComponents: runs of consecutive positions (<= 7 values) become tonal blocks
            int nblk = 0;
            int bfu = 8, k = 0;                                  // cursor: component k of BFU bfu's run
            auto advance = [&]() {
                while (bfu < 29 && k >= run_len[wib][bfu]) { bfu++; k = 0; }
            };
            advance();
            while (bfu < 29 && nblk < kMaxTonal) {
                TonalBlock blk;
                float tmp[8];
                const int first_pos = run_start[wib][bfu] + k;
                blk.pos = (unsigned short)first_pos;
                blk.bfu = (unsigned char)bfu;
                int cnt = 0, cur_pos;
                do {
                    cur_pos = run_start[wib][bfu] + k;
                    tmp[cnt++] = run_val[wib][bfu][k];
                    k++;
                    advance();
                } while (bfu < 29 && (run_start[wib][bfu] + k) == cur_pos + 1 && cnt < 7);
                float en;
                blk.sfi = (unsigned char)scale_block(T, tmp, cnt, en);
                blk.len = (unsigned char)cnt;
                blk.pad[0] = blk.pad[1] = blk.pad[2] = 0;
                for (int j = 0; j < 7; j++) blk.val[j] = j < cnt ? tmp[j] : 0.0f;
                tl->b[nblk++] = blk;
            }
            tl->n = nblk;
        }
        __syncwarp();
    } else if (lane == 0) {
        tl->n = 0;
    }
    // ScaleFrame on what is left of the spectrum
    float en, scale;
    const int sfi = scale_analyse(T, sv + start, len, en, scale);
    b.sfi[(size_t)unit * 32 + lane] = (unsigned char)sfi;
    b.energy[(size_t)unit * 32 + lane] = en;
    for (int i = lane; i < 1024; i += 32) {
        const float sc = __shfl_sync(0xffffffffu, scale, elem_bfu_s(i));
        float q = __fdiv_rn(sv[i], sc);
        if (fabsf(q) >= 1.0f) q = (q > 0.0f) ? 0.99999f : -0.99999f;
        gsp[i] = q;
    }
}

// =====================================================================================
// loudness term of every channel-frame (atrac3denc.cpp:811-820): l += e * Frame * curve over the 1024
// lines in order -- one sequential chain per channel-frame, so ONE LANE per channel-frame; the
// spectra are staged through shared memory in 32x32 tiles to keep the global reads coalesced.
// =====================================================================================
constexpr int kLoudWarps = 4;

__global__ void __launch_bounds__(kLoudWarps * 32) at3_loudterm_kernel(Geometry g, Buffers b)
{
    __shared__ float tile_all[kLoudWarps][32][33];
    const DevTables* __restrict__ T = b.tab;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long total = (long long)g.S * g.n_out * g.C;
    const long long unit0 = ((long long)blockIdx.x * kLoudWarps + wib) * 32;
    if (unit0 >= total) return;
    float (*tile)[33] = tile_all[wib];
    const long long mine = unit0 + lane;
    const bool live = mine < total;
    float fr[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    if (live) {
        const float* gs = b.gscale + (size_t)mine * 16;
        fr[0] = gs[2]; fr[1] = gs[6]; fr[2] = gs[10]; fr[3] = gs[14];
    }
    const int nrows = (int)min(32LL, total - unit0);
    float l = 0.0f;
    for (int t = 0; t < 32; t++) {
        // (eight rows' loads in flight at a time)
        for (int r0 = 0; r0 < nrows; r0 += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
                v[q] = r0 + q < nrows ? b.specs[(size_t)(unit0 + r0 + q) * 1024 + 32 * t + lane] : 0.0f;
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (r0 + q < nrows) tile[r0 + q][lane] = v[q];
        }
        __syncwarp();
        const float f = fr[t >> 3];
        if (live) {
#pragma unroll 8
            for (int k = 0; k < 32; k++) {
                const float x = tile[lane][k];
                l = fadd(l, fmul(fmul(fmul(x, x), f), T->loud_curve[32 * t + k]));
            }
        }
        __syncwarp();
    }
    if (live) b.chloud[mine] = l;
}

void launch_loudterm(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const long long total = (long long)g.S * g.n_out * g.C;
    const long long per_block = kLoudWarps * 32;
    ATDE_LAUNCH(at3_loudterm_kernel, (unsigned)((total + per_block - 1) / per_block), kLoudWarps * 32, 0, st, g, b);
}

void launch_scale_tonal(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const long long total = (long long)g.S * g.n_out * g.C;
    ATDE_LAUNCH(at3_scale_tonal_kernel, (unsigned)((total + kScaleWarps - 1) / kScaleWarps), kScaleWarps * 32, 0, st, g, b);
}

// =====================================================================================
// K8: bit allocation search + quantisation + entropy coding + frame assembly
// =====================================================================================
// QuantMantisas of one BFU at one word length, plus its CLC / VLC cost (TAt3SpecUnit::Provide).
// `in` = scaled values of the BFU.  m_out (optional) receives the mantissas.
struct UnitCost {
    unsigned clc, vlc;
    float err;
};

// Exact restatement with the complete candidate list and libstdc++'s sort order; only used when two
// candidates that can actually be re-rounded share the same |delta| (see quant_unit).
ATDE_NOINLINE float quant_unit_exact(const float* in, int len, float mul, float inv2, signed char* m)
{
    SortCand cand[128];
    int nc = 0;
    float e1 = 0.0f, e2 = 0.0f;
    for (int j = 0; j < len; j++) {
        const float t = fmul(in[j], mul);
        e1 = fadd(e1, fmul(in[j], in[j]));
        const int q = __float2int_rn(t);
        m[j] = (signed char)q;
        e2 = fadd(e2, fmul((float)(q * q), inv2));
        const float delta = fsub(t, fadd(truncf(t), 0.5f));
        if (fabsf(delta) < 0.25f) { cand[nc].delta = delta; cand[nc].idx = j; nc++; }
    }
    if (nc > 0) {
        std_sort_cands(cand, nc);
        if (e2 < e1) {
            for (int k = 0; k < nc; k++) {
                const int j = cand[k].idx;
                const float t = fmul(in[j], mul);
                const int q = m[j];
                const float aq = (float)abs(q);
                if (aq < fabsf(t) && aq < fsub(mul, 1.0f)) {
                    int q2 = q;
                    if (q > 0) q2++;
                    if (q < 0) q2--;
                    if (q == 0) q2 = t > 0.0f ? 1 : -1;
                    float ex = e2;
                    ex = fsub(ex, fmul((float)(q * q), inv2));
                    ex = fadd(ex, fmul((float)(q2 * q2), inv2));
                    if (fabsf(fsub(ex, e1)) < fabsf(fsub(e2, e1))) { m[j] = (signed char)q2; e2 = ex; }
                }
            }
        } else if (e2 > e1) {
            for (int k = 0; k < nc; k++) {
                const int j = cand[k].idx;
                const float t = fmul(in[j], mul);
                const int q = m[j];
                if ((float)abs(q) > fabsf(t)) {
                    int q2 = q;
                    if (q > 0) q2--;
                    if (q < 0) q2++;
                    float ex = e2;
                    ex = fsub(ex, fmul((float)(q * q), inv2));
                    ex = fadd(ex, fmul((float)(q2 * q2), inv2));
                    if (fabsf(fsub(ex, e1)) < fabsf(fsub(e2, e1))) { m[j] = (signed char)q2; e2 = ex; }
                }
            }
        }
    }
    return __fdiv_rn(e1, e2);
}

// ---- BFU geometry by 32-line row (atrac3.h:83-105): rows 0-1 hold four 8-line BFUs, rows 2-5 two
//      16-line BFUs, rows 6-15 one 32-line BFU each, rows 16-23 half a 64-line BFU, rows 24-31 a
//      quarter of a 128-line BFU.
ATDE_D int elem_bfu(int i)
{
    if (i < 64) return i >> 3;
    if (i < 192) return 8 + ((i - 64) >> 4);
    if (i < 512) return 16 + ((i - 192) >> 5);
    if (i < 768) return 26 + ((i - 512) >> 6);
    return 30 + ((i - 768) >> 7);
}
// BFU of line 32 r + lane = kRowBase[r] + (lane >> kRowShift[r]): rows 0-1 hold four 8-line BFUs, rows 2-5 two 16-line
// BFUs, every later row lies inside one BFU.  (Constant-bank tables: the row index is warp-uniform.)
__constant__ unsigned char kRowBase[32] = {0, 4, 8, 10, 12, 14, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
                                           26, 26, 27, 27, 28, 28, 29, 29, 30, 30, 30, 30, 31, 31, 31, 31};
__constant__ unsigned char kRowShift[32] = {3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
                                            5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5};
ATDE_D int row_bfu(int r, int lane) { return kRowBase[r] + (lane >> kRowShift[r]); }
ATDE_D unsigned row_bfu_mask(int r)
{
    if (r < 2) return 0xFu << (4 * r);
    if (r < 6) return 0x3u << (8 + 2 * (r - 2));
    if (r < 16) return 1u << (16 + (r - 6));
    if (r < 24) return 1u << (26 + ((r - 16) >> 1));
    return 1u << (30 + ((r - 24) >> 2));
}

// rows (of 32 lines) touched by a set of BFUs
ATDE_D unsigned rows_of_bfus(unsigned bm)
{
    unsigned rows = ((bm & 0xFu) ? 1u : 0u) | ((bm & 0xF0u) ? 2u : 0u);
    const unsigned p = (bm >> 8) & 0xFFu;                                  // BFUs 8..15: two per row
    rows |= ((p & 0x03u) ? 1u << 2 : 0u) | ((p & 0x0Cu) ? 1u << 3 : 0u) | ((p & 0x30u) ? 1u << 4 : 0u) | ((p & 0xC0u) ? 1u << 5 : 0u);
    rows |= ((bm >> 16) & 0x3FFu) << 6;                                    // BFUs 16..25: one row each
    const unsigned q = (bm >> 26) & 0xFu;                                  // BFUs 26..29: two rows each
    rows |= ((q & 1u) ? 0x3u << 16 : 0u) | ((q & 2u) ? 0x3u << 18 : 0u) | ((q & 4u) ? 0x3u << 20 : 0u) | ((q & 8u) ? 0x3u << 22 : 0u);
    rows |= ((bm >> 30) & 1u) ? 0xFu << 24 : 0u;                           // BFUs 30, 31: four rows each
    rows |= ((bm >> 31) & 1u) ? 0xFu << 28 : 0u;
    return rows;
}
// (float)(1.0 / (mul * mul)) for MaxQuant[1..7] (atrac_scale.cpp:45), evaluated in double like the reference
__device__ const float kInv2[8] = {0.0f, (float)(1.0 / 2.25), (float)(1.0 / 6.25), (float)(1.0 / 12.25), (float)(1.0 / 20.25),
                                   (float)(1.0 / 56.25), (float)(1.0 / 240.25), (float)(1.0 / 992.25)};

ATDE_D unsigned vlc_bits_of(int wl, int q) { return kHuffBits[kHuffOff[wl] + huff_index(q)]; }
ATDE_D unsigned vlc_pair_bits(int qa, int qb) { return kHuffBits[kVlcPairIdx[3 * (qa + 1) + (qb + 1)]]; }

// MSB-first bit field into a zeroed big-endian word array (value already masked to n bits);
// fields that start beyond the buffer are dropped (the reference truncates with resize()).
ATDE_D void put_bits3(unsigned* words, int cap_bits, int pos, int n, unsigned val)
{
    if (n <= 0 || pos + n > cap_bits) return;
    const int w = pos >> 5, off = pos & 31;
    const int room = 32 - off;
    if (n <= room) {
        atomicOr(&words[w], val << (room - n));
    } else {
        atomicOr(&words[w], val >> (n - room));
        atomicOr(&words[w + 1], val << (32 - (n - room)));
    }
}

// sequential writer used by the header / tonal section (one lane, many call sites: kept out of line
// so the hot search loop stays small in the instruction cache)
ATDE_NOINLINE int put_seq(unsigned* words, int cap_bits, int pos, unsigned v, int n)
{
    put_bits3(words, cap_bits, pos, n, v & ((1u << n) - 1u));
    return pos + n;
}

constexpr int kWordsPerCh = kMaxUnitBytes / 4 + 8;             // bitstream of one channel

struct __align__(16) PackShared {
    float sv[1024];                    // scaled spectrum of the channel
    float pq[1024];                    // per line: q*q/mul^2 of the unit being quantised (energy chain input);
                                       //   afterwards, per BFU region: |delta| of the re-rounding candidates;
                                       //   once the allocation is final: the channel's bitstream words (kWordsPerCh)
    unsigned char hb[1024];            // per line: VLC bits (pairs: on the even line); afterwards: candidate line offsets
    signed char mant[1024];            // mantissas of the unit quantised last, per BFU region
    unsigned char wl_now[32];          // word length being quantised per BFU (0 = not requested)
    unsigned char walk_dir[32];        // 0 none, 1 e2 < e1 (round up), 2 e2 > e1 (round down)
    float walk_thr[32];                // acceptance pre-filter per BFU
    int walk_cnt[32];                  // candidates collected per BFU
    unsigned short cache_vlc[8][32];   // VLC bits of the unit, indexed [wordlen][bfu] (the CLC bits are len * kClcLen[wordlen])
    float cache_err[8][10];            // energy ratio e1 / e2; only BFUs < 10 (BOOST_NAQ_END) ever look at it
    unsigned short ton_pos[kMaxTonal];
    unsigned char ton_bfu[kMaxTonal], ton_len[kMaxTonal], ton_sfi[kMaxTonal];
    unsigned char ton_vlc[kMaxTonal][8];   // VLC bits of the block's mantissas at quantiser 2..7
    unsigned char prec[32];
    int hdr_bits;                      // header + gain info bits of this channel
};

// TAt3SpecUnit::Provide (atrac3_bitstream.cpp:157-173) = QuantMantisas (atrac_scale.cpp:40-130) + CLC /
// VLC cost, for every BFU whose lane has `need` set, at that lane's word length `wl` — computed by the
// whole warp together.  All 32 lanes must call it.  Phases:
//   E   line-parallel: t = x*mul, q = rint(t), q*q/mul^2 and the VLC bits of every line of a requested BFU
//   C   lane per BFU: the SEQUENTIAL energy sum e2 over the BFU's lines (the reference's order) + VLC sum
//   E2  line-parallel, BFUs > 18 only (energy-aware re-rounding): collect the candidates that can
//       change something into a compact per-BFU list
//   W   lane per BFU: walk the candidates in ascending |delta|
// The energy-aware branch of the reference sorts every candidate by |delta| and walks the sorted list,
// re-rounding a value when that brings the quantised energy e2 closer to e1.  Restated without the sort:
//   * only candidates that pass the walk's own test (|m| < |t| && |m| < mul-1 when e2 < e1, |m| > |t|
//     when e2 > e1 -- a static property of the line) can change anything;
//   * a re-rounding moves e2 by d = (2|q|+1)/mul^2 (up) or (2|q|-1)/mul^2 (down) and is accepted only if
//     it lands closer to e1, i.e. d < 2*gap up to rounding; the gap only shrinks during the walk, so a
//     candidate with d >= 2*gap0 + slack can never be accepted (slack = 1e-4 relative to the energies,
//     far above the few ulps the float updates can be off) and is dropped;
//   * the rest is visited in ascending |delta| by repeated selection, and the walk stops once e2 has
//     reached or crossed e1 (every later candidate moves it further away by at least 1/mul^2 ~ 1e-3);
//   * with distinct |delta| among the visited candidates this is the reference's order exactly; if two
//     of them tie, the library's sort order matters and quant_unit_exact redoes the block.
// Returns (for need lanes) clc | vlc << 16 and the energy ratio e1/e2.
#ifdef ATDE_PACK_STATS
long long g_stats[64];
struct StatsPrinter { ~StatsPrinter() { fprintf(stderr, "compute_units calls %lld, units %lld, walkers %lld, frames*ch %lld, bisect iters %lld, memo hits %lld, settled passes %lld\n", g_stats[0], g_stats[1], g_stats[2], g_stats[3], g_stats[4], g_stats[5], g_stats[6]); for (int q = 0; q < 32; q++) fprintf(stderr, "%lld ", g_stats[8 + q]); fprintf(stderr, "\n"); } } g_stats_printer;
#endif
ATDE_NOINLINE unsigned compute_units(PackShared& sh, int lane, bool need, int wl, int start, int len, float e1, float& err_out)
{
    const unsigned nm = __ballot_sync(0xffffffffu, need);
#ifdef ATDE_PACK_STATS
    if (lane == 0) { g_stats[0]++; g_stats[1] += __popc(nm); g_stats[2] += __popc(nm >> 19); for (int q = 0; q < 32; q++) if ((nm >> q) & 1u) g_stats[8 + q]++; }
#endif
    sh.wl_now[lane] = need ? (unsigned char)wl : 0;
    __syncwarp();
    // ---- E ----
    for (unsigned rows = rows_of_bfus(nm); rows; rows &= rows - 1) {
        const int r = __ffs((int)rows) - 1;
        const int i = 32 * r + lane;
        const int w = sh.wl_now[row_bfu(r, lane)];
        const float mul = kMaxQuant[w];
        const float x = sh.sv[i];
        const int q = __float2int_rn(fmul(x, mul));
        const int qn = __shfl_down_sync(0xffffffffu, q, 1);
        if (w) {
            const float inv2 = kInv2[w];
            sh.mant[i] = (signed char)q;
            sh.pq[i] = fmul((float)(q * q), inv2);
            sh.hb[i] = (unsigned char)(w > 1 ? vlc_bits_of(w, q) : ((lane & 1) ? 0u : vlc_pair_bits(q, qn)));
        }
    }
    __syncwarp();
    // ---- C ----
    float e2 = 0.0f;
    unsigned vlc = 0;
    if (need) {
        // 8 lines per trip (every BFU length is a multiple of 8): the loads stay off the dependent chain
#pragma unroll 1
        for (int j = 0; j < len; j += 8) {
            const float4 a = *reinterpret_cast<const float4*>(sh.pq + start + j);
            const float4 c = *reinterpret_cast<const float4*>(sh.pq + start + j + 4);
            const uint2 h = *reinterpret_cast<const uint2*>(sh.hb + start + j);
            e2 = fadd(e2, a.x); e2 = fadd(e2, a.y); e2 = fadd(e2, a.z); e2 = fadd(e2, a.w);
            e2 = fadd(e2, c.x); e2 = fadd(e2, c.y); e2 = fadd(e2, c.z); e2 = fadd(e2, c.w);
            // byte sums: add the two words, then fold the four byte lanes (each lane <= 2*8 bits ... < 256)
            const unsigned t2 = (h.x & 0x00ff00ffu) + ((h.x >> 8) & 0x00ff00ffu) + (h.y & 0x00ff00ffu) + ((h.y >> 8) & 0x00ff00ffu);
            vlc += (t2 & 0xffffu) + (t2 >> 16);
        }
    }
    const float mulw = kMaxQuant[need ? wl : 0];
    const float inv2w = kInv2[need ? wl : 0];
    bool used_exact = false;
    float exact_err = 0.0f;
    const bool walk = need && lane > 18 /* LOSY_NAQ_START */ && e2 != e1;
    const unsigned wm = __ballot_sync(0xffffffffu, walk);
    if (wm) {
        // ---- E2 ----
        const bool up = e2 < e1;
        sh.walk_dir[lane] = walk ? (up ? 1 : 2) : 0;
        sh.walk_thr[lane] = fadd(fmul(2.0f, fabsf(fsub(e1, e2))), fmul(1e-4f, fadd(fadd(e1, e2), 1.0f)));
        sh.walk_cnt[lane] = 0;
        __syncwarp();
        for (unsigned rows = rows_of_bfus(wm); rows; rows &= rows - 1) {
            const int r = __ffs((int)rows) - 1;
            const int i = 32 * r + lane;
            const int bf = row_bfu(r, lane);
            const int dir = sh.walk_dir[bf];
            if (dir) {
                const int wq = sh.wl_now[bf];
                const float mul = kMaxQuant[wq];
                const float t = fmul(sh.sv[i], mul);
                const float ad = fabsf(fsub(t, fadd(truncf(t), 0.5f)));
                if (ad < 0.25f) {
                    const int aqi = abs((int)sh.mant[i]);
                    const float aq = (float)aqi;
                    const bool qual = dir == 1 ? (aq < fabsf(t) && aq < fsub(mul, 1.0f)) : (aq > fabsf(t));
                    const float d = fmul((float)(dir == 1 ? 2 * aqi + 1 : 2 * aqi - 1), kInv2[wq]);
                    if (qual && d < sh.walk_thr[bf]) {
                        const int b0 = kBlockStart[bf];
                        const int slot = atomicAdd(&sh.walk_cnt[bf], 1);
                        sh.pq[b0 + slot] = ad;                    // the chain inputs of this BFU are dead by now
                        sh.hb[b0 + slot] = (unsigned char)(i - b0);
                    }
                }
            }
        }
        __syncwarp();
        // ---- S ---- sort every walking BFU's candidate list by |delta| (rank by counting, the whole warp on one list at a
        //      time: the walk below then visits candidates in order instead of searching the minimum again at every step).
        //      Equal keys keep their list order; the walk falls back to the exact restatement before it would visit them.
        for (unsigned bm = wm; bm; bm &= bm - 1) {
            const int bf = __ffs((int)bm) - 1;
            const int nc = sh.walk_cnt[bf];
            if (nc < 2) continue;
            const int b0 = kBlockStart[bf];
            float* ck = sh.pq + b0;
            unsigned char* ci = sh.hb + b0;
            if (nc <= 32) {
                const bool mine = lane < nc;
                const float key = mine ? ck[lane] : 0.0f;
                const unsigned char idx = mine ? ci[lane] : 0;
                int rank = 0;
                ATDE_ROLLED
                for (int k = 0; k < nc; k++) {
                    const float kk = ck[k];
                    rank += (kk < key || (kk == key && k < lane)) ? 1 : 0;
                }
                __syncwarp();
                if (mine) { ck[rank] = key; ci[rank] = idx; }
            } else {                                               // up to 128 candidates: four slots per lane
                float key[4];
                unsigned char idx[4];
                int rank[4];
#pragma unroll
                for (int rr = 0; rr < 4; rr++) {
                    const int sl = lane + 32 * rr;
                    key[rr] = sl < nc ? ck[sl] : 0.0f;
                    idx[rr] = sl < nc ? ci[sl] : 0;
                    rank[rr] = 0;
                }
                ATDE_ROLLED
                for (int k = 0; k < nc; k++) {
                    const float kk = ck[k];
#pragma unroll
                    for (int rr = 0; rr < 4; rr++)
                        rank[rr] += (kk < key[rr] || (kk == key[rr] && k < lane + 32 * rr)) ? 1 : 0;
                }
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < 4; rr++)
                    if (lane + 32 * rr < nc) { ck[rank[rr]] = key[rr]; ci[rank[rr]] = idx[rr]; }
            }
        }
        __syncwarp();
        // ---- W ----
        if (walk) {
            const int nc = sh.walk_cnt[lane];
            const float* ckey = sh.pq + start;
            const unsigned char* cidx = sh.hb + start;
            signed char* m = sh.mant + start;
            int k = 0;
            while ((up ? (e2 < e1) : (e2 > e1)) && k < nc) {
                if (k + 1 < nc && ckey[k + 1] == ckey[k]) {
                    // two candidates that would be visited share |delta|: the reference's order is libstdc++'s
                    const float er = quant_unit_exact(sh.sv + start, len, mulw, inv2w, m);
                    unsigned v = 0;
                    if (wl > 1) {
                        ATDE_ROLLED
                        for (int j = 0; j < len; j++) v += vlc_bits_of(wl, m[j]);
                    } else {
                        ATDE_ROLLED
                        for (int j = 0; j < len / 2; j++) v += vlc_pair_bits(m[2 * j], m[2 * j + 1]);
                    }
                    vlc = v;
                    exact_err = er;
                    used_exact = true;
                    break;
                }
                const int j = cidx[k];
                k++;
                const int q = m[j];
                int q2 = q;
                if (up) {
                    if (q > 0) q2++;
                    if (q < 0) q2--;
                    if (q == 0) q2 = fmul(sh.sv[start + j], mulw) > 0.0f ? 1 : -1;
                } else {
                    if (q > 0) q2--;
                    if (q < 0) q2++;
                }
                float ex = e2;
                ex = fsub(ex, fmul((float)(q * q), inv2w));
                ex = fadd(ex, fmul((float)(q2 * q2), inv2w));
                if (fabsf(fsub(ex, e1)) < fabsf(fsub(e2, e1))) {
                    if (wl > 1) {
                        vlc += vlc_bits_of(wl, q2) - vlc_bits_of(wl, q);
                    } else {
                        const int other = m[j ^ 1];
                        vlc += (j & 1) ? vlc_pair_bits(other, q2) - vlc_pair_bits(other, q)
                                       : vlc_pair_bits(q2, other) - vlc_pair_bits(q, other);
                    }
                    m[j] = (signed char)q2;
                    e2 = ex;
                }
            }
        }
        __syncwarp();
    }
    err_out = used_exact ? exact_err : __fdiv_rn(e1, e2);
    const unsigned clc = wl > 1 ? (unsigned)kClcLen[wl] * len : 4u * len / 2;
    return clc | (vlc << 16);
}

// bits EncodeTonalComponents would emit for the current allocation (atrac3_bitstream.cpp:382-524),
// warp-cooperative: lane t owns tonal block t.
ATDE_D unsigned tonal_bits(const PackShared& sh, int n_ton, int lane, unsigned prec, int num_bfu)
{
    if (n_ton == 0) return 5;
    const int t = lane < n_ton ? lane : 0;
    const int my_bfu = sh.ton_bfu[t];
    const unsigned my_prec = __shfl_sync(0xffffffffu, prec, my_bfu);
    const bool active = lane < n_ton && my_bfu < num_bfu;
    const int quant = (int)max(2u, min(my_prec + 4u, 7u));
    const int key = active ? quant * 8 + sh.ton_len[t] : -1;
    const int pos = sh.ton_pos[t];
    // subgroup automaton of GroupTonalComponents, run by the first member of every (quant, len) group
    bool leader = active;
    int start_val = 0, limiter = 0, nsub = 0;
    unsigned flags = 0, bits = 0;
    ATDE_ROLLED
    for (int j = 0; j < n_ton; j++) {
        const int kj = __shfl_sync(0xffffffffu, key, j);
        const int pj = __shfl_sync(0xffffffffu, pos, j);
        if (kj < 0 || kj != key) continue;
        if (j < lane) { leader = false; continue; }
        if (!leader) continue;
        if (j == lane) {
            nsub = 1; start_val = pj; limiter = 0; flags = 1u << (pj >> 8);
        } else {
            if (pj - (start_val & ~63) < 64) ++limiter; else { limiter = 0; start_val = pj; }
            if (limiter < 7) {
                flags |= 1u << (pj >> 8);
            } else {
                bits += 10u + 12u * __popc(flags);
                nsub++; start_val = pj; limiter = 0; flags = 1u << (pj >> 8);
            }
        }
    }
    if (leader) bits += 10u + 12u * __popc(flags);
    if (active) bits += 12u + sh.ton_vlc[t][quant];
    const unsigned tcsgn = warp_sum_u(leader ? (unsigned)nsub : 0u);
    const unsigned sum = warp_sum_u(bits);
    return 5u + (tcsgn ? 2u : 0u) + sum;
}

#ifndef ATDE_PACK_MINBLOCKS
#define ATDE_PACK_MINBLOCKS 9          // 23.5 KB of shared memory per block: nine blocks (18 warps) per SM, <= 113 registers
#endif
__global__ void __launch_bounds__(64, ATDE_PACK_MINBLOCKS) at3_alloc_pack_kernel(Geometry g, Buffers b)
{
    __shared__ PackShared shm[2];
    __shared__ int s_shift;

    const DevTables* __restrict__ T = b.tab;
    const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    const long long frame = blockIdx.x;                         // s * n_out + f
    const int f = (int)(frame % g.n_out);
    const int s = (int)(frame / g.n_out);
    const long long unit = frame * g.C + ch;
    PackShared& sh = shm[ch];
    const int half = g.frame_sz >> 1;

    // ---- load; header + gain info size of both channels (WriteSoundUnit, :771-804) ----
    const float* gsp = b.specs + (size_t)unit * 1024;
    {   // the channel's 4 KB of scaled spectrum: eight 16-byte loads per lane, all in flight before the first store
        const float4* __restrict__ g4 = reinterpret_cast<const float4*>(gsp);
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = g4[lane + 32 * r];
#pragma unroll
        for (int r = 0; r < 8; r++) reinterpret_cast<float4*>(sh.sv)[lane + 32 * r] = v[r];
    }
    const TonalList* tl = b.tonal + unit;
    const int n_ton = g.no_tonal ? 0 : tl->n;
    if (lane < n_ton) {
        const TonalBlock& tb = tl->b[lane];
        sh.ton_pos[lane] = tb.pos; sh.ton_bfu[lane] = tb.bfu; sh.ton_len[lane] = tb.len; sh.ton_sfi[lane] = tb.sfi;
        ATDE_ROLLED
        for (int q = 2; q < 8; q++) {
            unsigned v = 0;
            const int off = kHuffOff[q];
            ATDE_ROLLED
            for (int z = 0; z < tb.len; z++)
                v += kHuffBits[off + huff_index(__float2int_rn(fmul(tb.val[z], kMaxQuant[q])))];
            sh.ton_vlc[lane][q] = (unsigned char)v;
        }
    }
    Curve cv[4];
#pragma unroll
    for (int band = 0; band < 4; band++) {
        cv[band].n = 0;
        if (!g.no_gain && band < kGainBands)
            cv[band] = b.curves[(((size_t)s * g.C + ch) * 4 + band) * g.n_out + f];
    }
    if (lane == 0) {
        int hb = (g.js && ch == 1) ? 14 : 6;
        hb += 2;
        for (int band = 0; band < 4; band++) hb += 3 + 9 * cv[band].n;
        sh.hdr_bits = hb;
    }
    if (g.js) __syncthreads();                  // the byte shift below reads both channels' header sizes
    else __syncwarp();                          // otherwise the two channel warps never meet
    // ---- byte budget of this channel ----
    int shift_bytes = 0;
    if (g.js) {
        if (threadIdx.x == 0) {
            const int b0 = -6 - shm[0].hdr_bits, b1 = -6 - shm[1].hdr_bits;
            const int used = 0 - b0 - b1;
            const int max_shift = half - (int)(1u + ((unsigned)used - 1u) / 8u);
            const float m_e = b.chloud[frame * 2], s_e = b.chloud[frame * 2 + 1];
            const float tot = fadd(s_e, m_e);
            float ratio = 0.0f;
            if (tot > 0.0f)
                ratio = __double2float_rn(__dsub_rn((double)__fdiv_rn(m_e, tot), 0.5));
            const int want = __float2int_rn(fmul((float)g.frame_sz, ratio));
            s_shift = max(min(want, max_shift), -max_shift);
        }
        __syncthreads();
        shift_bytes = s_shift;
    }
    if (g.js_mono) {
        // Mono input in a joint-stereo container (atrac3denc.cpp:843-849): the second element is EMPTY — one QMF
        // band, no gain points, no blocks: WriteJsParams + 3 (14 bits), numQmfBand - 1 (2), gain count (3) = 19
        // header bits, then an empty tonal section (5), numBfu - 1 = 0 (5), coding mode 1 (1), precision 0 (3) —
        // and CalcMSBytesShift hands the first element everything the second cannot use (:750-752).
        const int used = (6 + sh.hdr_bits) + (6 + 19);
        shift_bytes = half - (int)(1u + ((unsigned)used - 1u) / 8u);
    }
    const int my_bytes = (ch == 0) ? half + shift_bytes : half - shift_bytes;
    int to_alloc = -6 - sh.hdr_bits + 8 * my_bytes;
    const unsigned target = (unsigned)(unsigned short)max(1, to_alloc);
    const int cap_bits = kWordsPerCh * 32 - 32;
    const float loud = __fdiv_rn(b.loud[frame], kLoudFactor);

    // ---- per-BFU constants ----
    const int start = kBlockStart[lane], len = kBlockStart[lane + 1] - kBlockStart[lane];
    const int sfi = b.sfi[(size_t)unit * 32 + lane];
    const float energy = b.energy[(size_t)unit * 32 + lane];
    float ges = b.gscale[(size_t)unit * 16 + bfu_band(lane) * 4 + 2];
    {
        const float inf = __int_as_float(0x7f800000);
        if (!(fabsf(ges) < inf) || !(ges > 0.0f)) ges = 1.0f;   // SanitizeGainEnergyScale
    }
    const bool audible = !(fmul(energy, ges) < fmul(T->ath[lane], loud));
    const float sfi_corr = fmaxf(0.0f, fminf(63.0f, fadd((float)sfi, fmul(1.5f, g_log2f(ges)))));
    const float xdiv = lane < 3 ? 2.8f : (lane < 10 ? 2.6f : (lane < 15 ? 3.3f : (lane <= 20 ? 3.6f : (lane <= 28 ? 4.2f : 6.0f))));
    const float sfi_term = __fdiv_rn(sfi_corr, xdiv);
    const float fix = (float)kFixedAlloc[lane];
    int n_ton_mine = 0;                                          // tonal blocks whose first line is in this BFU
    ATDE_ROLLED
    for (int t = 0; t < n_ton; t++) n_ton_mine += (sh.ton_bfu[t] == lane);
    // AnalizeScaleFactorSpread (atrac_psy_common.cpp:105-124): sequential float sums
    float spread;
    {
        float sacc = 0.0f;
        ATDE_ROLLED
        for (int i = 0; i < 32; i++) sacc = fadd(sacc, (float)__shfl_sync(0xffffffffu, sfi, i));
        sacc = __fdiv_rn(sacc, 32.0f);
        float sigma = 0.0f;
        ATDE_ROLLED
        for (int i = 0; i < 32; i++) {
            float t = fsub((float)__shfl_sync(0xffffffffu, sfi, i), sacc);
            t = fmul(t, t);
            sigma = fadd(sigma, t);
        }
        sigma = __fsqrt_rn(__fdiv_rn(sigma, 32.0f));
        if (sigma > 14.0f) sigma = 14.0f;
        spread = __fdiv_rn(sigma, 14.0f);
    }
    const float fix_term = fmul(fsub(1.0f, spread), fix);
    unsigned cached = 0;                                         // bit w: (lane, w) is in the cache
    unsigned mant_wl = 0;                                        // word length whose mantissas sh.mant holds for this BFU
    float e1 = 0.0f;                                             // QuantMantisas' e1: energy of the scaled values, sequential
    ATDE_ROLLED
    for (int j = 0; j < len; j += 8) {
        const float4 a = *reinterpret_cast<const float4*>(sh.sv + start + j);
        const float4 c = *reinterpret_cast<const float4*>(sh.sv + start + j + 4);
        e1 = fadd(e1, fmul(a.x, a.x)); e1 = fadd(e1, fmul(a.y, a.y)); e1 = fadd(e1, fmul(a.z, a.z)); e1 = fadd(e1, fmul(a.w, a.w));
        e1 = fadd(e1, fmul(c.x, c.x)); e1 = fadd(e1, fmul(c.y, c.y)); e1 = fadd(e1, fmul(c.z, c.z)); e1 = fadd(e1, fmul(c.w, c.w));
    }

    // CalcInitialNumBfu (:567-585)
    int num_bfu = g.bfu_idx_const ? g.bfu_idx_const : 32;
    if (target < 101u) {
        int lim = 1;
        if (target > 5u) lim = (int)(target - 5u) / 3;
        lim = max(1, lim);
        num_bfu = min(num_bfu, lim);
    }
    num_bfu = max(1, num_bfu);

#ifdef ATDE_PACK_STATS
    if (lane == 0) g_stats[3]++;
#endif
    unsigned prec = 0;
    unsigned mode = 1;
    // CalcBitsAllocation (:272-336): this lane's word length at a given shift
    auto alloc_prec = [&](float shift) {
        unsigned p = 0;
        if (lane < num_bfu && audible) {
            const int tmp = __float2int_rz(fsub(fadd(fmul(spread, sfi_term), fix_term), shift));
            p = tmp > 7 ? 7u : (tmp < 0 ? 0u : (tmp == 0 ? 1u : (unsigned)tmp));
        }
        if (lane < num_bfu && n_ton_mine && p > 2u)
            p = max(2u, p - (unsigned)min(n_ton_mine, 8));
        return p;
    };
    for (;;) {                                                   // TConfigure: ba.Start(target, -8, 20)
        float mn = -8.0f, mx = 20.0f, last = 20.0f;
        bool settled = false;
        // The bit count of a step is a function of the word lengths CalcBitsAllocation hands out (for a fixed BFU
        // count), and late in the bisection consecutive shifts hand out the same ones.  The last evaluation on either
        // side of the target is kept (lane-wise: word lengths before and after the energy-error bumps; warp-wide:
        // total and coding mode) and a step that meets one of them again takes its result instead of recounting.
        unsigned key_lo = 0xffu, key_hi = 0xffu, res_lo = 0, res_hi = 0;   // 0xff: empty
        unsigned tot_lo = 0, tot_hi = 0, mode_lo = 1, mode_hi = 1;
        for (;;) {                                               // TAlloc::Encode
#ifdef ATDE_PACK_STATS
            if (lane == 0) g_stats[4]++;
#endif
            // CheckBfus (:587-600) repeats the whole search with one BFU less when the last BFU ends up without bits.
            // Its word length never grows with the shift, every shift this search can still visit (midpoints, and
            // `last` at exhaustion, which is >= mn up to rounding) lies above mn - 0.02, and a BFU without bits is not
            // bumped: once it has none at mn - 0.02 the outcome is settled and the rest of the search is skipped.
            if (!g.bfu_idx_const && num_bfu > 1) {
                const unsigned p_low = alloc_prec(fsub(mn, 0.02f));
                if (__shfl_sync(0xffffffffu, p_low, num_bfu - 1) == 0) { settled = true; break; }
            }
            const bool exhausted = mx <= mn;
            const float shift = exhausted ? last : __double2float_rn(__dmul_rn((double)fadd(mx, mn), 0.5));   // (max + min) / 2.0: halving is exact
            prec = alloc_prec(shift);
            const unsigned key = prec;
            unsigned total;
            if (__all_sync(0xffffffffu, key == key_lo)) {
                prec = res_lo; total = tot_lo; mode = mode_lo;
#ifdef ATDE_PACK_STATS
                if (lane == 0) g_stats[5]++;
#endif
            } else if (__all_sync(0xffffffffu, key == key_hi)) {
                prec = res_hi; total = tot_hi; mode = mode_hi;
#ifdef ATDE_PACK_STATS
                if (lane == 0) g_stats[5]++;
#endif
            } else {
                // CalcSpecsBitsConsumption + ConsiderEnergyErr (:190-257): every BFU's bump chain is independent;
                // missing (bfu, wordlen) units are quantised by the whole warp together
                unsigned cvb = 0;
                for (;;) {
                    const bool active = lane < num_bfu && prec != 0;
                    const bool need = active && !((cached >> prec) & 1u);
                    if (__any_sync(0xffffffffu, need)) {
                        float er;
                        const unsigned cv2 = compute_units(sh, lane, need, (int)prec, start, len, e1, er);
                        if (need) {
                            sh.cache_vlc[prec][lane] = (unsigned short)(cv2 >> 16);
                            if (lane < 10) sh.cache_err[prec][lane] = er;
                            cached |= 1u << prec;
                            mant_wl = prec;
                        }
                    }
                    bool bump = false;
                    if (active) {
                        cvb = ((prec > 1u ? (unsigned)kClcLen[prec] : 2u) * (unsigned)len) | ((unsigned)sh.cache_vlc[prec][lane] << 16);
                        if (lane < 10) {                             // BOOST_NAQ_END
                            const float e = sh.cache_err[prec][lane];
                            bump = ((e > 0.0f && e < 0.7f) || e > 1.2f) && prec < 7u;
                        }
                    }
                    if (bump) prec++;
                    if (!__any_sync(0xffffffffu, bump)) break;
                }
                const unsigned clc = warp_sum_u(prec ? (cvb & 0xffffu) : 0u);
                const unsigned vlc = warp_sum_u(prec ? (cvb >> 16) : 0u);
                const unsigned nz = __popc(__ballot_sync(0xffffffffu, prec != 0));
                mode = clc <= vlc;
                total = (unsigned)num_bfu * 3u + 6u * nz + (mode ? clc : vlc);
                total += tonal_bits(sh, n_ton, lane, prec, num_bfu);
                if (total < target) { key_lo = key; res_lo = prec; tot_lo = total; mode_lo = mode; }
                else if (total > target) { key_hi = key; res_hi = prec; tot_hi = total; mode_hi = mode; }
            }
            if (exhausted) break;
            if (total < target) { last = shift; mx = fsub(shift, 0.01f); }
            else if (total > target) { mn = fadd(shift, 0.01f); }
            else break;
        }
        if (settled) {
#ifdef ATDE_PACK_STATS
            if (lane == 0) g_stats[6]++;
#endif
            num_bfu--;
            continue;
        }
        if (!g.bfu_idx_const && num_bfu > 1) {
            const unsigned last_prec = __shfl_sync(0xffffffffu, prec, num_bfu - 1);
            if (last_prec == 0) { num_bfu--; continue; }         // CheckBfus -> EStatus::Repeat
        }
        break;
    }
    if (b.tap_prec) b.tap_prec[(size_t)unit * 32 + lane] = lane < num_bfu ? (unsigned char)prec : 0xff;
    sh.prec[lane] = (unsigned char)prec;
    __syncwarp();

    // ---- mantissas of the final allocation (units whose last quantisation was at another word length are redone;
    //      the procedure is deterministic, so they come out as they were costed) ----
    const bool in_use = lane < num_bfu;
    {
        const bool redo = in_use && prec && mant_wl != prec;
        if (__any_sync(0xffffffffu, redo)) {
            float er;
            compute_units(sh, lane, redo, (int)prec, start, len, e1, er);
        }
    }
    // ---- write the channel's sound unit; the bitstream words take over the quantiser's scratch ----
    unsigned* const words = reinterpret_cast<unsigned*>(sh.pq);
    __syncwarp();
    for (int i = lane; i < kWordsPerCh; i += 32) words[i] = 0;
    __syncwarp();
    int pos = 0;
    if (lane == 0) {
        unsigned* W = words;
        auto put = [&](unsigned v, int n) { pos = put_seq(W, cap_bits, pos, v, n); };
        if (g.js && ch == 1) {
            put(0, 1); put(7, 3);
            for (int i = 0; i < 4; i++) put(3, 2);
            put(3, 2);
        } else {
            put(0x28, 6);
        }
        put(3, 2);                                               // numQmfBand - 1
        ATDE_ROLLED
        for (int band = 0; band < 4; band++) {
            put(cv[band].n, 3);
            ATDE_ROLLED
            for (int i = 0; i < cv[band].n; i++) { put(cv[band].level[i], 4); put(cv[band].loc[i], 5); }
        }
        // EncodeTonalComponents (:382-524)
        int order[kMaxTonal], keys[kMaxTonal], na = 0;
        ATDE_ROLLED
        for (int t = 0; t < n_ton; t++) {
            if (sh.ton_bfu[t] >= num_bfu) continue;
            const int quant = max(2, min((int)sh.prec[sh.ton_bfu[t]] + 4, 7));
            order[na] = t; keys[na] = quant * 8 + sh.ton_len[t]; na++;
        }
        // count subgroups first (tcsgn is written ahead of them)
        int tcsgn = 0;
        ATDE_ROLLED
        for (int key = 16; key < 64; key++) {
            int startv = 0, limiter = 0; bool open = false;
            ATDE_ROLLED
            for (int a = 0; a < na; a++) {
                if (keys[a] != key) continue;
                const int p = sh.ton_pos[order[a]];
                if (!open) { open = true; tcsgn++; startv = p; limiter = 0; }
                else {
                    if (p - (startv & ~63) < 64) ++limiter; else { limiter = 0; startv = p; }
                    if (limiter >= 7) { tcsgn++; startv = p; limiter = 0; }
                }
            }
        }
        put((unsigned)tcsgn, 5);
        if (tcsgn) {
            put(0, 2);
            ATDE_ROLLED
            for (int key = 16; key < 64; key++) {
                int mem[kMaxTonal], nm = 0;
                ATDE_ROLLED
                for (int a = 0; a < na; a++) if (keys[a] == key) mem[nm++] = order[a];
                if (!nm) continue;
                const int quant = key >> 3, coded = key & 7;
                int a0 = 0;
                while (a0 < nm) {
                    // extent of the subgroup starting at a0
                    int startv = sh.ton_pos[mem[a0]], limiter = 0, a1 = a0 + 1;
                    while (a1 < nm) {
                        const int p = sh.ton_pos[mem[a1]];
                        if (p - (startv & ~63) < 64) ++limiter; else { limiter = 0; startv = p; }
                        if (limiter >= 7) break;
                        a1++;
                    }
                    unsigned char cnt[16];
                    ATDE_ROLLED
                    for (int j = 0; j < 16; j++) cnt[j] = 0;
                    ATDE_ROLLED
                    for (int a = a0; a < a1; a++) cnt[sh.ton_pos[mem[a]] >> 6]++;
                    unsigned bandf = 0;
                    ATDE_ROLLED
                    for (int j = 0; j < 16; j++) if (cnt[j]) bandf |= 1u << (j >> 2);
                    ATDE_ROLLED
                    for (int j = 0; j < 4; j++) put((bandf >> j) & 1u, 1);
                    put((unsigned)(coded - 1), 3);
                    put((unsigned)quant, 3);
                    int lastp = a0;
                    ATDE_ROLLED
                    for (int j = 0; j < 16; j++) {
                        if (!((bandf >> (j >> 2)) & 1u)) continue;
                        put(cnt[j], 3);
                        ATDE_ROLLED
                        for (int k = lastp; k < lastp + cnt[j]; k++) {
                            const int t = mem[k];
                            put(sh.ton_sfi[t], 6);
                            put((unsigned)(sh.ton_pos[t] - j * 64), 6);
                            const TonalBlock& tb = tl->b[t];
                            const int off = kHuffOff[quant];
                            ATDE_ROLLED
                            for (int z = 0; z < tb.len; z++) {
                                const int hi = huff_index(__float2int_rn(fmul(tb.val[z], kMaxQuant[quant])));
                                put(kHuffCode[off + hi], kHuffBits[off + hi]);
                            }
                        }
                        lastp += cnt[j];
                    }
                    a0 = a1;
                }
            }
        }
        put((unsigned)(num_bfu - 1), 5);
        put(mode, 1);
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    // precisions, scale factor indices, mantissas (EncodeSpecs :541-564)
    if (in_use) put_bits3(words, cap_bits, pos + 3 * lane, 3, prec);
    pos += 3 * num_bfu;
    const unsigned nzmask = __ballot_sync(0xffffffffu, in_use && prec != 0);
    if (in_use && prec) put_bits3(words, cap_bits, pos + 6 * __popc(nzmask & ((1u << lane) - 1u)), 6, (unsigned)sfi);
    pos += 6 * __popc(nzmask);
    // Mantissas, line-parallel: BFUs follow each other in line order and unused ones contribute no bits, so the bit
    // position of a line is the running sum of the code lengths of all lines before it (a warp scan per 32-line row).
    sh.wl_now[lane] = (in_use ? (unsigned char)prec : 0);
    __syncwarp();
    {
        int base = pos;
        for (unsigned rows = rows_of_bfus(nzmask); rows; rows &= rows - 1) {
            const int r = __ffs((int)rows) - 1;
            const int i = 32 * r + lane;
            const int p = sh.wl_now[row_bfu(r, lane)];
            const int m = sh.mant[i];
            const int mn = __shfl_down_sync(0xffffffffu, m, 1);
            unsigned nb = 0, code = 0;
            if (p > 1) {
                if (mode) {
                    nb = kClcLen[p];
                    code = (unsigned)m & ((1u << nb) - 1u);
                } else {
                    const int hi = kHuffOff[p] + huff_index(m);
                    nb = kHuffBits[hi];
                    code = kHuffCode[hi];
                }
            } else if (p == 1 && !(lane & 1)) {                  // word length 1 codes pairs of lines
                if (mode) {
                    nb = 4;
                    code = ((unsigned)kClcIdx[m + 2] << 2) | kClcIdx[mn + 2];
                } else {
                    const int hi = kVlcPairIdx[3 * (m + 1) + (mn + 1)];
                    nb = kHuffBits[hi];
                    code = kHuffCode[hi];
                }
            }
            unsigned incl = nb;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            if (nb) put_bits3(words, cap_bits, base + (int)(incl - nb), (int)nb, code);
            base += (int)__shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncwarp();
    // ---- frame assembly (:826-843): channel 0 bytes, then channel 1 (byte-reversed when joint stereo);
    //      mono without JS duplicates the half frame.  Every warp places its own channel's bytes (the split
    //      point is known since the byte-shift barrier), so a fast channel does not wait for the other one ----
    unsigned char* dst = b.out + (size_t)frame * g.frame_sz;
    const int n0 = half + shift_bytes;
    if (g.C == 2) {
        const int lo = ch == 0 ? 0 : n0, hi = ch == 0 ? n0 : g.frame_sz;
        for (int i = lo + lane; i < hi; i += 32) {
            const int q = ch == 0 ? i : (g.js ? (g.frame_sz - n0 - 1) - (i - n0) : i - n0);
            const unsigned w = words[q >> 2];
            dst[i] = (unsigned char)(w >> (24 - 8 * (q & 3)));
        }
    } else if (g.js_mono) {
        // the empty element's 33 bits 0 111 11111111 11 00 000 00000 00000 1 000, zero-padded, byte-reversed at the frame end
        for (int i = lane; i < g.frame_sz; i += 32) {
            unsigned char v;
            if (i < n0) {
                const unsigned w = words[i >> 2];
                v = (unsigned char)(w >> (24 - 8 * (i & 3)));
            } else {
                const int q = (g.frame_sz - 1) - i;
                v = q == 0 ? 0x7F : q == 1 ? 0xFC : q == 3 ? 0x04 : 0x00;
            }
            dst[i] = v;
        }
    } else {
        for (int i = lane; i < g.frame_sz; i += 32) {
            const int q = i < n0 ? i : i - n0;
            const unsigned w = words[q >> 2];
            dst[i] = (unsigned char)(w >> (24 - 8 * (q & 3)));
        }
    }
}

void launch_alloc_pack(const Geometry& g, const Buffers& b, cudaStream_t st)
{
    const long long frames = (long long)g.S * g.n_out;
    ATDE_LAUNCH(at3_alloc_pack_kernel, (unsigned)frames, 32 * g.C, 0, st, g, b);
}

} // namespace at3
} // namespace atde
