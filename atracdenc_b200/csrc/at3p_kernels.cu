// at3p_kernels.cu — ATRAC3plus encode path on sm_100a: PQF analysis, MDCT-256 x16, frame packer.
// See at3p_kernels.cuh for the reference chain and what is still missing (GHA).
//
// Bit-exactness rules are the ones of the other codecs: un-fused IEEE fp32 in the
// reference's operation order, doubles where the reference uses doubles (the PQF accumulators),
// kissfft restated stage by stage, tables computed on the host with the reference's expressions.
#include "at3p_kernels.cuh"
#include "kissfft_dev.cuh"
#include "glibc_trig.cuh"
#include "host_tables.h"
#include "at3p_tables_gen.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace atde {
namespace at3p {

__constant__ float c_fir[kPqfProto];       // analysis prototype, atrac3plus_pqf.c:59-78
__constant__ float c_dct_sc[16];           // TMIDCT<32>(32 * 128 * 512) pre/post twiddles, mdct.cpp:63-66
__constant__ cpx c_tw8[8];                 // forward kissfft twiddles of the 8-point FFT

// =====================================================================================
// host tables
// =====================================================================================
// One table set (and one upload of the __constant__ symbols) per CUDA DEVICE: a process may create ATRAC3plus encoders
// on several GPUs (atde_settings::device), and both a cudaMalloc'ed table and a __constant__ symbol belong to the
// device that was current when they were written.  A failed upload is not cached: the next call tries again.
constexpr int kMaxDevices = 64;
static DevTables* g_dev_tables[kMaxDevices] = {};
static std::mutex g_tables_mu;

static float bits_to_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }

static DevTables* build_tables()
{
    DevTables* h = new DevTables();
    memset(h, 0, sizeof(*h));
    const auto sc = mdct_sincos(256, 1.0f);
    memcpy(h->sincos256, sc.data(), sizeof(h->sincos256));
    for (int i = 0; i < 128; i++)                                     // at3p_mdct.cpp:36-40
        h->sine_win[i] = 2.0 * sinf((i + 0.5) * (M_PI / (2.0 * 128)));
    const auto tw = kiss_twiddles(64, false);
    memcpy(h->tw64, tw.data(), sizeof(h->tw64));
    for (int i = 0; i < 64; i++) h->scale_table[i] = bits_to_float(kAt3pScaleBits[i]);
    h->inv_mant[0] = 0.0f;
    for (int i = 1; i < 8; i++) h->inv_mant[i] = 1.0f / bits_to_float(kAt3pMantTabBits[i]);   // TUnit::Multiplier, at3p_bitstream.cpp:365-368
    memcpy(h->spec_tab, kAt3pSpecTab, sizeof(h->spec_tab));
    memcpy(h->vlc_off, kAt3pVlcOff, sizeof(h->vlc_off));
    for (int t = 0; t < 56; t++)
        h->spec_pack[t] = (unsigned)kAt3pSpecTab[t][0] | (unsigned)kAt3pSpecTab[t][1] << 4 | (unsigned)kAt3pSpecTab[t][2] << 8 |
                          (unsigned)kAt3pSpecTab[t][3] << 12 | kAt3pVlcOff[t] << 16;
    static_assert(sizeof(kAt3pVlc) == sizeof(DevTables::vlc), "generated VLC table size");
    memcpy(h->vlc, kAt3pVlc, sizeof(h->vlc));
    memcpy(h->wl_vlc, kAt3pWlVlc, sizeof(h->wl_vlc));
    memcpy(h->tone_bands_vlc, kAt3pToneBandsVlc, sizeof(h->tone_bands_vlc));
    memcpy(h->qu_to_subband, kAt3pQuToSubband, sizeof(h->qu_to_subband));
    memcpy(h->sb_to_powgrps, kAt3pSbToPowGrps, sizeof(h->sb_to_powgrps));
    {   // ff_atrac3p_init_dsp_static (ff/atrac3plusdsp.c:49-66), same expressions and types
        const double twopi = 2 * M_PI;
        for (int i = 0; i < 2048; i++) h->sine_table[i] = sin(twopi * i / 2048);
        for (int i = 0; i < 256; i++) h->hann_window[i] = (1.0f - cos(twopi * i / 256.0f)) * 0.5f;
        for (int i = 0; i < 64; i++) h->amp_sf_tab[i] = exp2f((i - 3) / 4.0f);
    }

    float fir[kPqfProto];
    for (int i = 0; i < kPqfProto; i++) fir[i] = bits_to_float(kAt3pFirBits[i]);
    // atde_create_dct4_16(128 * 512.0): TMIDCT<32>(32 * 65536) -> CalcSinCos(32, 32 * 65536 / 2)
    const auto dsc = mdct_sincos(32, 1048576.0f);
    const auto tw8 = kiss_twiddles(8, false);
    DevTables* d = nullptr;
    if (cudaMalloc(&d, sizeof(DevTables)) == cudaSuccess &&
        cudaMemcpy(d, h, sizeof(DevTables), cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpyToSymbol(c_fir, fir, sizeof(fir)) == cudaSuccess &&
        cudaMemcpyToSymbol(c_dct_sc, dsc.data(), 16 * sizeof(float)) == cudaSuccess &&
        cudaMemcpyToSymbol(c_tw8, tw8.data(), 8 * sizeof(cpx)) == cudaSuccess) {
        delete h;
        return d;
    }
    if (d) cudaFree(d);
    delete h;
    return nullptr;
}

const DevTables* device_tables()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    std::lock_guard<std::mutex> lock(g_tables_mu);
    if (!g_dev_tables[dev]) g_dev_tables[dev] = build_tables();
    return g_dev_tables[dev];
}

// =====================================================================================
// P1: PQF analysis, at3plus_pqf_do_analyse (atrac3plus_pqf.c:80-147)
// =====================================================================================
// One block per (frame, stream): threads 0..127 = output position i of channel 0, 128..255 of
// channel 1.  Position i reads the 384 samples x[0..383] = buf[16 i ..] (buf = 368 history samples |
// the frame) and produces one sample of each of the 16 subbands:
//   vectoring  y[t] = sum_{j<12} (double)(fir[12 t + j] * x[32 j + t]),  t < 32  (float product, double sum)
//   matrixing  yy[k] = (float)(y[k + 8] + y[7 - k]),  yy[k + 8] = (float)(y[k + 16] + y[31 - k])
//              res = DCT-IV-16(yy) (TMIDCT<32> + 8-point kissfft);  out[sb][i] = res[15 - sb]
// The tile is padded by one float per 16 so that the 16-float stride between lanes is conflict-free.
constexpr int kPqfTile = kFrame + kPqfOverlap;                       // 2416
ATDE_HD int pphys(int e) { return e + (e >> 4); }
constexpr int kPqfTilePad = kPqfTile + kPqfTile / 16 + 1;

// out[i] = -Buf[i + 8] of TMIDCT<32> (mdct.cpp:73-81, mdct.h:117-180) on 16 inputs.
ATDE_D void dct4_16(const float* in, float* out)
{
    cpx v[8];                                       // FFT input in kissfft's gather order
    // 8 = 4 x 2: slot = 2 d0 + d1 holds input d0 + 4 d1
#pragma unroll
    for (int slot = 0; slot < 8; slot++) {
        const int idx = (slot >> 1) + 4 * (slot & 1);
        const int n = 2 * idx;
        const float r0 = in[n], i0 = in[15 - n];
        const float c = c_dct_sc[n], s = c_dct_sc[n + 1];
        // xr = -2.0 * (i0 * s + r0 * c); xi = -2.0 * (i0 * c - r0 * s): the float sums, doubled and negated exactly
        v[slot].r = fmul(-2.0f, fadd(fmul(i0, s), fmul(r0, c)));
        v[slot].i = fmul(-2.0f, fsub(fmul(i0, c), fmul(r0, s)));
    }
    // radix-2, m = 1 (fstride 4), twiddle tw[0]
#pragma unroll
    for (int q = 0; q < 4; q++) kf_bfly2(v[2 * q], v[2 * q + 1], c_tw8[0]);
    // radix-4, m = 2 (fstride 1): elements k + 2q, twiddles tw[k], tw[2k], tw[3k]
#pragma unroll
    for (int k = 0; k < 2; k++) kf_bfly4<false>(v[k], v[k + 2], v[k + 4], v[k + 6], c_tw8[k], c_tw8[2 * k], c_tw8[3 * k]);
    float buf[32];
#pragma unroll
    for (int h = 0; h < 8; h++) {
        const int n = 2 * h;
        const float r0 = v[h].r, i0 = v[h].i;
        const float c = c_dct_sc[n], s = c_dct_sc[n + 1];
        const float r1 = fadd(fmul(r0, c), fmul(i0, s));
        const float i1 = fsub(fmul(r0, s), fmul(i0, c));
        if (n < 8) { buf[23 - n] = r1; buf[24 + n] = r1; buf[8 + n] = i1; buf[7 - n] = -i1; }
        else       { buf[23 - n] = r1; buf[n - 8] = -r1; buf[8 + n] = i1; buf[39 - n] = i1; }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = -buf[i + 8];
}

__global__ void __launch_bounds__(256) at3p_pqf_kernel(const float* __restrict__ pcm, const float* __restrict__ pcm_tail,
                                                        float* __restrict__ bands, int S, int C, int F, int L, int joff)
{
    __shared__ float xs[2][kPqfTilePad];
    const int f = blockIdx.x, s = blockIdx.y;
    const int tid = threadIdx.x;
    const long long n0 = (long long)f * kFrame - kPqfOverlap;        // tile sample 0 within the stream
    const float* src = pcm + (size_t)s * F * kFrame * C;
    ATDE_PAR_FOR(t, kPqfTile) {
        const long long n = n0 + t;
        float v0 = 0.0f, v1 = 0.0f;
        if (n >= 0) {
            if (C == 2) {
                const float2 v = *reinterpret_cast<const float2*>(src + (size_t)n * 2);
                v0 = v.x; v1 = v.y;
            } else {
                v0 = src[n];
            }
        } else if (pcm_tail) {                                        // the last 368 samples of the previous batch
            const float* tl = pcm_tail + ((size_t)s * kPqfOverlap + (size_t)(n + kPqfOverlap)) * C;
            v0 = tl[0];
            if (C == 2) v1 = tl[1];
        }
        xs[0][pphys(t)] = v0;
        xs[1][pphys(t)] = v1;
    }
    __syncthreads();
    const int ch = tid >> 7, i = tid & 127;
    if (ch >= C) return;
    const float* x = xs[ch];
    const int base = 16 * i;
    float yy[16];
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        // four accumulator chains at once: y[k+8], y[7-k], y[k+16], y[31-k]
        const int t0 = k + 8, t1 = 7 - k, t2 = k + 16, t3 = 31 - k;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            a0 = __dadd_rn(a0, (double)fmul(c_fir[12 * t0 + j], x[pphys(base + 32 * j + t0)]));
            a1 = __dadd_rn(a1, (double)fmul(c_fir[12 * t1 + j], x[pphys(base + 32 * j + t1)]));
            a2 = __dadd_rn(a2, (double)fmul(c_fir[12 * t2 + j], x[pphys(base + 32 * j + t2)]));
            a3 = __dadd_rn(a3, (double)fmul(c_fir[12 * t3 + j], x[pphys(base + 32 * j + t3)]));
        }
        const float lo = __double2float_rn(__dadd_rn(a0, a1));
        const float hi = __double2float_rn(__dadd_rn(a2, a3));
        // yy[k], yy[k + 8] without dynamic register indexing
#pragma unroll
        for (int q = 0; q < 8; q++) if (q == k) { yy[q] = lo; yy[q + 8] = hi; }
    }
    float res[16];
    dct4_16(yy, res);
    float* out = bands + (((size_t)s * C + ch) * L + joff + f) * kFrame + i;
#pragma unroll
    for (int sb = 0; sb < 16; sb++) out[sb * kSbSamples] = res[15 - sb];
}

void launch_pqf(const float* pcm, const float* pcm_tail, float* bands, int S, int C, int F, int L, int joff, cudaStream_t st)
{
    dim3 grid(F, S);
    ATDE_LAUNCH(at3p_pqf_kernel, grid, 256, 0, st, pcm, pcm_tail, bands, S, C, F, L, joff);
}

// =====================================================================================
// P3: tone filter — TGhaProcessorBase::ApplyFilter (at3p_gha.cpp:581-687) + ff_atrac3p_generate_tones
//     (ff/atrac3plusdsp.c:130-204) + the MDCT input scaling of EncodeFrame (at3p.cpp:147-153)
// =====================================================================================
// The reference keeps two frames of tone parameters in Atrac3pChanUnitCtx; what it subtracts from the
// frame being encoded is a function of three consecutive GHA results only:
//   tb_next  the result of this call (tones_info),  tb_now  of the previous call (tones_info_prev),
//   tb_old   of the call before (its pending envelope enters tones_now->curr_env, computed one call earlier).
struct WaveGroup {                    // Atrac3pWavesData after ApplyFilter's sharing / leader swaps
    int num_wavs, src_ch, first;      // waves = tb->params[src_ch][first ..]
    int has_start, start_pos, has_stop, stop_pos;    // pend_env
};

ATDE_D WaveGroup resolve_group(const ToneBlock* tb, int C, int ch, int sb)
{
    WaveGroup g = {0, 0, 0, 0, 0, 0, 0};
    if (!tb->present || sb >= tb->num_tone_bands) return g;           // memset-zero state
    int c = ch;
    if (C == 2 && tb->second_is_leader) c = 1 - c;                    // std::swap(channels[0], channels[1])
    if (c == 1 && tb->tone_sharing[sb]) c = 0;                        // channels[1] = channels[0]
    g.num_wavs = tb->sb[c][sb][1];
    g.src_ch = c;
    g.first = tb->sb[c][sb][0];
    const unsigned e0 = (unsigned)tb->sb[c][sb][2], e1 = (unsigned)tb->sb[c][sb][3];
    g.has_start = e0 != 0xffffffffu; g.start_pos = g.has_start ? (int)e0 : -1;
    g.has_stop = e1 != 0xffffffffu;  g.stop_pos = g.has_stop ? (int)e1 : 32;
    return g;
}

struct Envelope { int has_start, start_pos, has_stop, stop_pos; };

// curr_env of `next` given the pending envelopes of `now` and `next` (atrac3plusdsp.c:141-166)
ATDE_D Envelope current_envelope(const WaveGroup& now, const WaveGroup& next)
{
    Envelope e;
    if (next.has_start && next.start_pos < next.stop_pos) { e.has_start = 1; e.start_pos = next.start_pos + 32; }
    else if (now.has_start) { e.has_start = 1; e.start_pos = now.start_pos; }
    else { e.has_start = 0; e.start_pos = 0; }
    if (now.has_stop && now.stop_pos >= e.start_pos) { e.has_stop = 1; e.stop_pos = now.stop_pos; }
    else if (next.has_stop) { e.has_stop = 1; e.stop_pos = next.stop_pos + 32; }
    else { e.has_stop = 0; e.stop_pos = 64; }
    return e;
}

// sample i of waves_synth (atrac3plusdsp.c:79-128), amplitude_mode = 1, no phase inversion
ATDE_D float synth_sample(const DevTables* T, const ToneBlock* tb, const WaveGroup& g, const Envelope& env, int reg_offset, int i)
{
    float v = 0.0f;
    for (int wn = 0; wn < g.num_wavs; wn++) {
        const int* prm = tb->params[g.src_ch][g.first + wn];
        const double amp = (double)fmul(T->amp_sf_tab[prm[1]], 1.0f);
        const int inc = prm[0];
        const int pos = ((((prm[3] & 0x1f) << 6) - (reg_offset ^ 128) * inc) + i * inc) & 2047;
        v = __double2float_rn(__dadd_rn((double)v, __dmul_rn((double)T->sine_table[pos], amp)));
    }
    if (env.has_start) {
        const int pos = (env.start_pos << 2) - reg_offset;
        if (pos > 0 && pos <= 128) {
            if (i < pos) v = 0.0f;
            if ((!env.has_stop || env.start_pos != env.stop_pos) && i >= pos && i < pos + 4)
                v = fmul(v, T->hann_window[32 * (i - pos)]);
        }
    }
    if (env.has_stop) {
        const int pos = ((env.stop_pos + 1) << 2) - reg_offset;
        if (pos > 0 && pos <= 128) {
            if (i >= pos - 4 && i < pos) v = fmul(v, T->hann_window[96 - 32 * (i - (pos - 4))]);
            if (i >= pos) v = 0.0f;
        }
    }
    return v;
}

__global__ void __launch_bounds__(128) at3p_tone_filter_kernel(const DevTables* __restrict__ T, const float* __restrict__ bands,
                                                                const ToneBlock* __restrict__ tb_old,
                                                                const ToneBlock* __restrict__ tb_now,
                                                                const ToneBlock* __restrict__ tb_next,
                                                                float* __restrict__ resid, int units, int C, FilterLayout lay)
{
    const int u = blockIdx.x / C, ch = blockIdx.x % C, i = threadIdx.x;
    // unit u = (stream s, output q); flat layout: fo = units, everything indexed by u
    const int s = u / lay.fo, q = u % lay.fo;
    const size_t ti = (size_t)s * lay.tone_stride + q;
    const ToneBlock* old = tb_old + ti;
    const ToneBlock* now = tb_now + ti;
    const ToneBlock* next = tb_next + ti;
    // bands == nullptr: GHA_PASS_INPUT is off — the work buffer starts from zeros (at3p.cpp:165-169)
    const float* in = bands ? bands + (((size_t)s * C + ch) * lay.in_frames + lay.in_off + q) * kFrame : nullptr;
    float* out = resid + (((size_t)s * C + ch) * lay.out_frames + lay.out_off + q) * kFrame;
    const bool any = now->present || next->present;                    // tones_present || prev tones_present
    // the thread's 16 samples (one per subband) are fetched together — every load in flight at once — and parked in
    // shared memory (each thread reads back only what it wrote): the subband loop below stays rolled
    __shared__ float xs[kSubbands][128];
    {
        float xin[kSubbands];
#pragma unroll
        for (int sb = 0; sb < kSubbands; sb++) xin[sb] = in ? in[sb * kSbSamples + i] : 0.0f;
#pragma unroll
        for (int sb = 0; sb < kSubbands; sb++) xs[sb][i] = xin[sb];
    }
    for (int sb = 0; sb < kSubbands; sb++) {
        float x = xs[sb][i];
        if (sb < 8 && any) {
            const WaveGroup gn = resolve_group(now, C, ch, sb), gx = resolve_group(next, C, ch, sb);
            if (gn.num_wavs || gx.num_wavs) {
                const Envelope env_next = current_envelope(gn, gx);
                const Envelope env_now = current_envelope(resolve_group(old, C, ch, sb), gn);   // what the previous call left
                const bool reg1 = env_now.stop_pos >= 32, reg2 = env_next.start_pos < 32;
                float w1 = 0.0f, w2 = 0.0f;
                if (gn.num_wavs && reg1) w1 = synth_sample(T, now, gn, env_now, 128, i);
                if (gx.num_wavs && reg2) w2 = synth_sample(T, next, gx, env_next, 0, i);
                if (gn.num_wavs && gx.num_wavs && reg1 && reg2) {
                    w1 = fmul(w1, T->hann_window[128 + i]);
                    w2 = fmul(w2, T->hann_window[i]);
                } else {
                    if (gn.num_wavs && !env_now.has_stop) w1 = fmul(w1, T->hann_window[128 + i]);
                    if (gx.num_wavs && !env_next.has_start) w2 = fmul(w2, T->hann_window[i]);
                }
                x = fsub(x, fadd(w1, w2));
            }
        }
        // tmp[i] = x[i] / (32768.0 / 1.122018)  (at3p.cpp:150-153): double division, rounded to float on the store
        out[sb * kSbSamples + i] = __double2float_rn(__ddiv_rn((double)x, 32768.0 / 1.122018));
    }
}

void launch_tone_filter(const DevTables* T, const float* bands, const ToneBlock* tb_old, const ToneBlock* tb_now,
                        const ToneBlock* tb_next, float* resid, int units, int C, const FilterLayout& lay, cudaStream_t st)
{
    ATDE_LAUNCH(at3p_tone_filter_kernel, (unsigned)(units * C), 128, 0, st, T, bands, tb_old, tb_now, tb_next, resid, units, C, lay);
}

// =====================================================================================
// P4: TAt3pMDCT::Do with sine windows (at3p_mdct.cpp:52-96), TMDCT<256> (mdct.h:51-104)
// =====================================================================================
// One WARP per (stream, frame, channel); two iterations of eight subbands, four lanes per band.
//   phase A  in[j] = win[j] * prev[j] (the overlap half the previous frame left behind, recomputed from
//            the previous frame's samples), in[128 + j] = win[127 - j] * cur[j]
//   phase B  fold + pre-twiddle into kissfft's gather order (64 = 4 x 4 x 4); lane k4 owns slots
//            16 k4 .. 16 k4 + 15 and runs the stages m = 1 and m = 4 on registers
//   exchange through the tile; phase C: stage m = 16 on elements k + 16 q, k = k4 + 4 c; post-twiddle;
//            odd bands reversed (SwapArray); staged and written as coalesced float4
constexpr int kPmWarps = 4;
constexpr int kPmBandStride = 264;                   // floats per band in the tile (256 + 8)
constexpr int kPmXchStride = 68;                     // cpx per band in the exchange layout: 4 lanes x 17, and 68 = 4 mod 16 spreads the bands
constexpr int kPmOutStride = 136;                    // floats per band in the output stage (128 + 8)

__global__ void __launch_bounds__(kPmWarps * 32, 6) at3p_mdct_kernel(const DevTables* __restrict__ T,
                                                                      const float* __restrict__ resid,
                                                                      float* __restrict__ specs, int S, int C, int F, int lead)
{
    __shared__ __align__(16) float tile[kPmWarps][8 * kPmBandStride];
    __shared__ __align__(16) float s_sincos[128];
    __shared__ __align__(16) float s_win[128];
    __shared__ __align__(16) cpx s_tw[64];
    ATDE_PAR_FOR(i, 128) { s_sincos[i] = T->sincos256[i]; s_win[i] = T->sine_win[i]; }
    ATDE_PAR_FOR(i, 64) s_tw[i] = T->tw64[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tl = tile[warp];
    const long long n_units = (long long)S * F * C;
    for (long long unit = (long long)blockIdx.x * kPmWarps + warp; unit < n_units;
         unit += (long long)gridDim.x * kPmWarps) {
        const int c = (int)(unit % C);
        const long long sf = unit / C;
        const int f = (int)(sf % F), s = (int)(sf / F);
        // resid [S][C][lead + F][2048]: with lead = 1 frame 0 is the residual the previous batch ended with
        const float* cur = resid + (((size_t)s * C + c) * (F + lead) + f + lead) * kFrame;
        float* const outp = specs + (size_t)unit * kFrame;
#pragma unroll 1
        for (int it = 0; it < 2; it++) {                     // subbands 8 it .. 8 it + 7
            __syncwarp();
            // ---- phase A ----
#pragma unroll
            for (int h = 0; h < 8; h++) {
                const int w = lane + 32 * h;                 // float4 index within the 8 x 128 samples
                const int band = w >> 5, i0 = 4 * (w & 31);
                const float4 x = *reinterpret_cast<const float4*>(cur + (8 * it + band) * kSbSamples + i0);
                float4 y = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (f + lead > 0) y = *reinterpret_cast<const float4*>(cur - kFrame + (8 * it + band) * kSbSamples + i0);
                const float4 wf = *reinterpret_cast<const float4*>(&s_win[i0]);
                const float4 wb = *reinterpret_cast<const float4*>(&s_win[124 - i0]);
                float* in = tl + band * kPmBandStride;
                *reinterpret_cast<float4*>(in + i0) =
                    make_float4(fmul(wf.x, y.x), fmul(wf.y, y.y), fmul(wf.z, y.z), fmul(wf.w, y.w));
                *reinterpret_cast<float4*>(in + 128 + i0) =
                    make_float4(fmul(wb.w, x.x), fmul(wb.z, x.y), fmul(wb.y, x.z), fmul(wb.x, x.w));
            }
            __syncwarp();
            // ---- phase B ----
            const int band = lane >> 2, k4 = lane & 3;
            cpx e[4][4];                                     // after the exchange: element (k4 + 4 c) + 16 q -> e[c][q]
            {
                const float* in = tl + band * kPmBandStride;
                cpx v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    // slot 16 k4 + j of the 4x4x4 digit reversal: i = d0 + 4 d1 + 16 d2
                    const int n = 2 * (k4 + 4 * (j >> 2) + 16 * (j & 3));    // N = 256, n4 = 64, n34 = 192, n54 = 320
                    float r0, i0;
                    if ((j & 3) < 2) { r0 = fadd(in[191 - n], in[192 + n]); i0 = fsub(in[64 + n], in[63 - n]); }   // n < 64
                    else             { r0 = fsub(in[191 - n], in[n - 64]);  i0 = fadd(in[64 + n], in[319 - n]); }
                    const float2 cs = *reinterpret_cast<const float2*>(&s_sincos[n]);
                    v[j].r = fadd(fmul(r0, cs.x), fmul(i0, cs.y));
                    v[j].i = fsub(fmul(i0, cs.x), fmul(r0, cs.y));
                }
                // radix-4, m = 1 (fstride 16): elements 4 g + q, twiddles tw[0]
#pragma unroll
                for (int gq = 0; gq < 4; gq++)
                    kf_bfly4<false>(v[4 * gq], v[4 * gq + 1], v[4 * gq + 2], v[4 * gq + 3], s_tw[0], s_tw[0], s_tw[0]);
                // radix-4, m = 4 (fstride 4): elements k + 4 q, twiddles tw[4 k q']
#pragma unroll
                for (int k = 0; k < 4; k++)
                    kf_bfly4<false>(v[k], v[k + 4], v[k + 8], v[k + 12], s_tw[4 * k], s_tw[8 * k], s_tw[12 * k]);
                __syncwarp();                                // every lane has read its MDCT input
                cpx* xch = reinterpret_cast<cpx*>(tl) + band * kPmXchStride;
#pragma unroll
                for (int j = 0; j < 16; j++) xch[17 * k4 + j] = v[j];         // 17-element lane stride: conflict-free
                __syncwarp();
#pragma unroll
                for (int cq = 0; cq < 4; cq++)
#pragma unroll
                    for (int q = 0; q < 4; q++) e[cq][q] = xch[17 * q + (k4 + 4 * cq)];   // element (k4 + 4cq) + 16 q
            }
            // ---- phase C: radix-4, m = 16 (fstride 1) ----
#pragma unroll
            for (int cq = 0; cq < 4; cq++) {
                const int k = k4 + 4 * cq;
                kf_bfly4<false>(e[cq][0], e[cq][1], e[cq][2], e[cq][3], s_tw[k], s_tw[2 * k], s_tw[3 * k]);
            }
            __syncwarp();
            {
                float* sp = tl + band * kPmOutStride;
#pragma unroll
                for (int cq = 0; cq < 4; cq++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int n = 2 * (k4 + 4 * cq + 16 * q);
                        const cpx z = e[cq][q];
                        const float2 cs = *reinterpret_cast<const float2*>(&s_sincos[n]);
                        const float va = fsub(fmul(-z.r, cs.x), fmul(z.i, cs.y));
                        const float vb = fadd(fmul(-z.r, cs.y), fmul(z.i, cs.x));
                        int pa = n, pb = 127 - n;
                        if (band & 1) { pa = 127 - pa; pb = 127 - pb; }       // SwapArray for odd subbands (8 it is even)
                        sp[pa] = va;
                        sp[pb] = vb;
                    }
            }
            __syncwarp();
            float4* out = reinterpret_cast<float4*>(outp + 1024 * it);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int w = lane + 32 * q;                 // float4 index; band = w >> 5
                out[w] = *reinterpret_cast<const float4*>(tl + (w >> 5) * kPmOutStride + 4 * (w & 31));
            }
        }
    }
}

void launch_mdct(const DevTables* T, const float* resid, float* specs, int S, int C, int F, int lead, cudaStream_t st)
{
    const long long n_units = (long long)S * F * C;
    long long blocks = (n_units + kPmWarps - 1) / kPmWarps;
    if (blocks > 148 * 6 * 4) blocks = 148 * 6 * 4;
    ATDE_LAUNCH(at3p_mdct_kernel, (unsigned)blocks, kPmWarps * 32, 0, st, T, resid, specs, S, C, F, lead);
}

// =====================================================================================
// P5: TScaler<NAt3p::TScaleTable>::ScaleFrame (atrac_scale.cpp:141-188) + TAt3PBitStream::WriteFrame
//     (at3p_bitstream.cpp:703-726) with its five part encoders (:98-701)
// =====================================================================================
// One WARP per frame (both channels).  The reference's allocation is static: word length per quant unit
// from TConfigure's table (:106-112), so a unit's mantissas, its best code table and its bit cost do
// not depend on how many units are finally coded; the only search is TTonalComponentEncoder's
// "drop units until the frame fits" loop (:596-609), evaluated here for every candidate count at once.
__constant__ unsigned char c_alloc[32] = {7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7,
                                          7, 6, 6, 6, 6, 6, 6, 6, 6, 6, 5, 5, 4, 3, 2, 1};      // at3p_bitstream.cpp:106-112
__constant__ unsigned short c_qu_start[33] = {0, 16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256, 288, 320, 352,
                                              384, 448, 512, 576, 640, 704, 768, 896, 1024, 1152, 1280, 1408, 1536,
                                              1664, 1792, 1920, 2048};                            // at3p_tables.h:66-72
constexpr int kPackWarps = 4;
constexpr int kFrameBits = kFrameBytes * 8;
constexpr int kAllocBits = kFrameBits - 3;           // FrameSzToAllocBits, at3p_bitstream.cpp:441

struct __align__(16) PackSm {
    signed char mant[2][kFrame];
    unsigned words[kFrameBytes / 4 + 4];
    unsigned twords[96];                             // the tonal part's own buffer (TDumper::Buf), <= 3072 bits
    unsigned short qbits[2][32];                     // bits of the cheapest code table per unit
    unsigned char qtab[2][32];
    unsigned char sfi[2][32];
};

ATDE_D void put_bits(unsigned* words, int cap_bits, int pos, int n, unsigned val)
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

// TBitStream::Write keeps the low n bits of the value (bitstream.cpp:49)
ATDE_NOINLINE int put_field(unsigned* words, int cap_bits, int pos, unsigned v, int n)
{
    put_bits(words, cap_bits, pos, n, n >= 32 ? v : (v & ((1u << n) - 1u)));
    return pos + n;
}

// One VLC symbol of TQuantUnitsEncoder::EncodeQuSpectra (:283-343): optional group flag, code, sign bits.
template <int NC>
ATDE_D unsigned spec_symbol(const unsigned* __restrict__ vlc, int g, int bits, bool sgn, const signed char* m, int s, int& nbits)
{
    unsigned val = 0, signs = 0;
    int nsign = 0;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        int t = m[s * NC + i];
        if (!sgn && t != 0) {
            signs = (signs << 1) | (t < 0 ? 1u : 0u);
            nsign++;
            if (t < 0) t = -t;
        } else {
            t &= (1 << bits) - 1;
        }
        val |= (unsigned)t << (bits * i);
    }
    const unsigned e = vlc[val & 255u];
    unsigned out = e & 0xffffu;
    int n = (int)(e >> 16);
    out = (out << nsign) | signs;
    n += nsign;
    if (g != 1 && (s & (g - 1)) == 0) { out |= 1u << n; n++; }       // group sizes are 1, 2, 4
    nbits = n;
    return out;
}

// The bit count of the symbols of ONE GROUP OF FOUR LINES (mantissas mv[0..3], group gi of its quant unit) under the code
// table described by `sp` (DevTables::spec_pack): four symbols of one coefficient, two of two or one of four; a symbol
// costs its VLC code, a sign bit per non-zero coefficient of an unsigned table and a group flag every g-th symbol
// (TQuantUnitsEncoder::EncodeQuSpectra, at3p_bitstream.cpp:283-343).
ATDE_D unsigned group_cost(unsigned sp, const unsigned* __restrict__ vlc_all, const int* mv, int gi)
{
    const int g = sp & 15u, nc = (sp >> 4) & 15u, nb = (sp >> 8) & 15u;
    const bool sgn = (sp >> 12) & 1u;
    const unsigned* __restrict__ vlc = vlc_all + (sp >> 16);
    const unsigned mask = (1u << nb) - 1u;
    unsigned t[4], n = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (sgn) {
            t[i] = (unsigned)mv[i] & mask;
        } else {
            t[i] = (unsigned)abs(mv[i]);
            n += mv[i] != 0;
        }
    }
    const unsigned gm = (unsigned)g - 1u;                     // group sizes are 1, 2, 4
    if (nc == 4) {
        n += vlc[(t[0] | t[1] << nb | t[2] << (2 * nb) | t[3] << (3 * nb)) & 255u] >> 16;
        n += g != 1 && ((unsigned)gi & gm) == 0;
    } else if (nc == 2) {
        n += vlc[(t[0] | t[1] << nb) & 255u] >> 16;
        n += vlc[(t[2] | t[3] << nb) & 255u] >> 16;
        n += g != 1 && ((unsigned)(2 * gi) & gm) == 0;        // the odd symbol never opens a group
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            n += vlc[t[i] & 255u] >> 16;
            n += g != 1 && ((unsigned)(4 * gi + i) & gm) == 0;
        }
    }
    return n;
}

ATDE_D int first_set_bit(unsigned x) { return x ? 31 - __clz((int)x) : 0; }   // util.h:65-76

// TTonalComponentEncoder::WriteTonalBlock (:465-594) into the part's own buffer; one lane.  Returns bits used.
ATDE_D int write_tonal_block(const DevTables* T, unsigned* w, int cap, int pos, int channels, const ToneBlock* tb)
{
    pos = put_field(w, cap, pos, 1, 1);                                          // GHA amplitude mode 1
    const unsigned tbv = T->tone_bands_vlc[tb->num_tone_bands - 1];
    pos = put_field(w, cap, pos, tbv & 0xffffu, (int)(tbv >> 16));
    const int ntb = tb->num_tone_bands;
    auto flags = [&](const int* fl, int nfl) {                                   // WriteSubbandFlags (:444-463)
        int sum = 0;
        for (int i = 0; i < nfl; i++) sum += fl[i] ? 1 : 0;
        if (sum == 0) {
            pos = put_field(w, cap, pos, 0, 1);
        } else if (sum == nfl) {
            pos = put_field(w, cap, pos, 1, 1);
            pos = put_field(w, cap, pos, 0, 1);
        } else {
            pos = put_field(w, cap, pos, 1, 1);
            pos = put_field(w, cap, pos, 1, 1);
            for (int i = 0; i < nfl; i++) pos = put_field(w, cap, pos, fl[i] ? 1 : 0, 1);
        }
    };
    if (channels == 2) {
        flags(tb->tone_sharing, ntb);
        flags(&tb->second_is_leader, 1);
        pos = put_field(w, cap, pos, 0, 1);
    }
    for (int ch = 0; ch < channels; ch++) {
        if (ch) pos = put_field(w, cap, pos, 0, 1);                              // each channel has its own envelope
        for (int i = 0; i < ntb; i++) {
            if (ch && tb->tone_sharing[i]) continue;
            const unsigned e0 = (unsigned)tb->sb[ch][i][2], e1 = (unsigned)tb->sb[ch][i][3];
            if (e0 != 0xffffffffu) { pos = put_field(w, cap, pos, 1, 1); pos = put_field(w, cap, pos, e0, 5); }
            else pos = put_field(w, cap, pos, 0, 1);
            if (e1 != 0xffffffffu) { pos = put_field(w, cap, pos, 1, 1); pos = put_field(w, cap, pos, e1, 5); }
            else pos = put_field(w, cap, pos, 0, 1);
        }
        pos = put_field(w, cap, pos, 0, ch + 1);                                 // num waves mode
        for (int i = 0; i < ntb; i++) {
            if (ch && tb->tone_sharing[i]) continue;
            pos = put_field(w, cap, pos, (unsigned)tb->sb[ch][i][1], 4);
        }
        if (ch) pos = put_field(w, cap, pos, 0, 1);                              // frequencies coded independently
        for (int i = 0; i < ntb; i++) {
            if (ch && tb->tone_sharing[i]) continue;
            const int nw = tb->sb[ch][i][1];
            if (nw == 0) continue;
            const int (*prm)[4] = &tb->params[ch][tb->sb[ch][i][0]];
            // CreateFreqBitPack (:39-96): ascending vs descending delta coding, fewer bits wins (ties: descending)
            int bits_asc = 10, bits_desc = 10;
            {
                unsigned prev = (unsigned)prm[0][0] & 1023u;
                for (int k = 1; k < nw; k++) {
                    bits_asc += prev < 512 ? 10 : first_set_bit(1023 - prev) + 1;
                    prev = (unsigned)prm[k][0] & 1023u;
                }
                prev = (unsigned)prm[nw - 1][0] & 1023u;
                for (int k = nw - 2; k >= 0; k--) {
                    bits_desc += first_set_bit(prev) + 1;
                    prev = (unsigned)prm[k][0] & 1023u;
                }
            }
            const bool asc = nw == 1 || bits_asc < bits_desc;
            if (nw > 1) pos = put_field(w, cap, pos, asc ? 0 : 1, 1);
            if (asc) {
                unsigned prev = (unsigned)prm[0][0] & 1023u;
                pos = put_field(w, cap, pos, prev, 10);
                for (int k = 1; k < nw; k++) {
                    const unsigned cur = (unsigned)prm[k][0] & 1023u;
                    if (prev < 512) {
                        pos = put_field(w, cap, pos, cur, 10);
                    } else {
                        const int bq = first_set_bit(1023 - prev) + 1;
                        pos = put_field(w, cap, pos, (cur - (1024u - (1u << bq))) & 0xffffu, bq);
                    }
                    prev = cur;
                }
            } else {
                unsigned prev = (unsigned)prm[nw - 1][0] & 1023u;
                pos = put_field(w, cap, pos, prev, 10);
                for (int k = nw - 2; k >= 0; k--) {
                    const unsigned cur = (unsigned)prm[k][0] & 1023u;
                    pos = put_field(w, cap, pos, cur, first_set_bit(prev) + 1);
                    prev = cur;
                }
            }
        }
        pos = put_field(w, cap, pos, 0, ch + 1);                                 // amplitude mode
        for (int i = 0; i < ntb; i++) {
            if (ch && tb->tone_sharing[i]) continue;
            const int nw = tb->sb[ch][i][1];
            for (int k = 0; k < nw; k++) pos = put_field(w, cap, pos, (unsigned)tb->params[ch][tb->sb[ch][i][0] + k][1], 6);
        }
        for (int i = 0; i < ntb; i++) {
            if (ch && tb->tone_sharing[i]) continue;
            const int nw = tb->sb[ch][i][1];
            for (int k = 0; k < nw; k++) pos = put_field(w, cap, pos, (unsigned)tb->params[ch][tb->sb[ch][i][0] + k][3], 5);
        }
    }
    return pos;
}

ATDE_D unsigned warp_incl_scan(unsigned v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned a = __shfl_up_sync(0xffffffffu, v, d);
        v += lane >= d ? a : 0u;
    }
    return v;
}

// quant unit of spectral line i (c_qu_start: 8 units of 16 lines, 8 of 32, 6 of 64, 10 of 128) and its first line
ATDE_D int qu_of_line(int i) { return i < 128 ? i >> 4 : (i < 384 ? 8 + ((i - 128) >> 5) : (i < 768 ? 16 + ((i - 384) >> 6) : 22 + ((i - 768) >> 7))); }
ATDE_D int qu_first_line(int q) { return q < 8 ? 16 * q : (q < 16 ? 128 + 32 * (q - 8) : (q < 22 ? 384 + 64 * (q - 16) : 768 + 128 * (q - 22))); }

__global__ void __launch_bounds__(kPackWarps * 32) at3p_pack_kernel(const DevTables* __restrict__ T,
                                                                     const float* __restrict__ specs,
                                                                     const ToneBlock* __restrict__ tones,
                                                                     unsigned char* __restrict__ frames, int units, int C,
                                                                     int fo, int tone_stride)
{
    __shared__ PackSm sm[kPackWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kPackWarps + warp;
    if (unit >= units) return;
    PackSm& sh = sm[warp];
    for (int i = lane; i < kFrameBytes / 4 + 4; i += 32) sh.words[i] = 0;
    for (int i = lane; i < 96; i += 32) sh.twords[i] = 0;

    // ---- TScaler::Scale per quant unit + the fixed-word-length quantiser (QuantMantisas, ea = false).
    //      Lane q owns quant unit q for the maximum and the scale-factor search (32 units at once: the loads of a unit are
    //      independent of every other unit's search); the division and rounding then run over the lines, coalesced
    {
        const int my_start = qu_first_line(lane), my_len = qu_first_line(lane + 1) - my_start;
        const float my_mul = T->inv_mant[c_alloc[lane]];
        for (int ch = 0; ch < C; ch++) {
            const float* x = specs + ((size_t)unit * C + ch) * kFrame;
            float mx = 0.0f;
            for (int i = 0; i < my_len; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(x + my_start + i);
                mx = fmaxf(fmaxf(fmaxf(mx, fabsf(v.x)), fmaxf(fabsf(v.y), fabsf(v.z))), fabsf(v.w));
            }
            if (mx > 1.0f) mx = 1.0f;                                  // MAX_SCALE
            int lo = 0, hi = 63;                                       // lower_bound over the ascending table
#pragma unroll
            for (int it = 0; it < 6; it++) {                           // 64 entries: six halvings, then lo == hi
                const int mid = (lo + hi) >> 1;
                if (T->scale_table[mid] < mx) lo = mid + 1; else hi = mid;
            }
            const float sf = T->scale_table[lo];
            sh.sfi[ch][lane] = (unsigned char)lo;
#pragma unroll 4
            for (int r = 0; r < kFrame / 32; r++) {
                const int i = 32 * r + lane, q = qu_of_line(i);
                const float sfq = __shfl_sync(0xffffffffu, sf, q), mul = __shfl_sync(0xffffffffu, my_mul, q);
                float v = __fdiv_rn(x[i], sfq);
                if (fabsf(v) >= 1.0f) v = v > 0.0f ? 0.99999f : -0.99999f;
                sh.mant[ch][i] = (signed char)__float2int_rn(fmul(v, mul));
            }
        }
    }
    __syncwarp();
    // ---- TUnit::GetOrCompute (:370-397): cheapest of the 8 code tables per unit.  A lane takes FOUR LINES (one, two or
    //      four symbols, whatever the table); a pass of the warp covers 128 lines, i.e. eight units of 16 lines, four of
    //      32, two of 64 or one of 128 — lane segments of 4 / 8 / 16 / 32, summed with segmented shuffles.  The tables of
    //      a pass are uniform over the warp except in the one pass where the word length changes between two units.
    for (int ch = 0; ch < C; ch++)
        for (int p = 0; p < kFrame / 128; p++) {
            const int line0 = 128 * p + 4 * lane;
            const int qu = qu_of_line(line0), gi = (line0 - qu_first_line(qu)) >> 2;
            const int seg = p == 0 ? 4 : (p <= 2 ? 8 : (p <= 5 ? 16 : 32));     // lanes per unit in this pass
            const int wl = c_alloc[qu];
            const int mm = *reinterpret_cast<const int*>(sh.mant[ch] + line0);
            const int mv[4] = {(int)(signed char)(mm & 0xff), (int)(signed char)((mm >> 8) & 0xff),
                               (int)(signed char)((mm >> 16) & 0xff), (int)(signed char)((mm >> 24) & 0xff)};
            unsigned best = 0xffffffffu;
            for (int ti = 0; ti < 8; ti++) {
                unsigned bits = group_cost(T->spec_pack[wl - 1 + 7 * ti], T->vlc, mv, gi);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {                               // segmented sum, branch-free
                    const unsigned o = __shfl_xor_sync(0xffffffffu, bits, d);
                    bits += d < seg ? o : 0u;
                }
                best = min(best, bits * 8u + (unsigned)ti);                      // first minimum wins (t < consumed)
            }
            if ((lane & (seg - 1)) == 0) { sh.qbits[ch][qu] = (unsigned short)(best >> 3); sh.qtab[ch][qu] = (unsigned char)(best & 7u); }
        }
    __syncwarp();
    // ---- TTonalComponentEncoder::Encode (:611-669), once: its buffer survives the Repeat rounds
    // tones == nullptr: GHA_WRITE_TONAL is off — `delay` never holds a tone block (at3p.cpp:173-177)
    const ToneBlock* tb = tones ? tones + (size_t)(unit / fo) * tone_stride + unit % fo : nullptr;   // unit = (stream, output)
    int tonal_bits = 0;
    if (lane == 0) {
        unsigned* w = sh.twords;
        const int cap = 96 * 32;
        int p = 0;
        if (C == 2) p = put_field(w, cap, p, 0, 2);                      // swap_channels, negate_coeffs
        for (int ch = 0; ch < C; ch++) p = put_field(w, cap, p, 0, 1);   // every window is a sine window (at3p.cpp:161)
        for (int ch = 0; ch < C; ch++) p = put_field(w, cap, p, 0, 1);   // no gain compensation
        if (tb && tb->present && tb->num_tone_bands) {
            p = put_field(w, cap, p, 1, 1);
            p = write_tonal_block(T, w, cap, p, C, tb);
        } else {
            p = put_field(w, cap, p, 0, 1);
        }
        p = put_field(w, cap, p, 0, 1);                                  // no noise info
        p = put_field(w, cap, p, 3, 2);                                  // terminator
        tonal_bits = p;
    }
    tonal_bits = __shfl_sync(0xffffffffu, tonal_bits, 0);
    // ---- how many quant units fit: 32, else 28, 27, ... (:596-609).  Lane l evaluates n = l + 1.
    int num_qu;
    {
        const int n = lane + 1;
        // word-length part (:164-246): deltas of the static table
        int maxd = 0;
        for (int i = 1; i < n; i++) maxd |= abs((int)c_alloc[i] - (int)c_alloc[i - 1]);
        const int t0 = maxd >= 3 ? 2 : (maxd == 2 ? 1 : 0), t1 = maxd >= 3 ? 3 : t0;
        int wl_idx = 0;
        unsigned wl_best = 0xffffffffu;
        for (int t = t0; t <= t1; t++) {
            unsigned sum = 0;
            for (int i = 1; i < n; i++) sum += T->wl_vlc[t][((int)c_alloc[i] - (int)c_alloc[i - 1]) & 7] >> 16;
            if (sum < wl_best) { wl_best = sum; wl_idx = t; }
        }
        unsigned total = 6;                                              // TConfigure: units - 1 (5), mute flag (1)
        total += 2 + 2 + 2 + 2 + 3 + wl_best;
        if (C == 2) total += 2 + 2 + 2 + (unsigned)n * (T->wl_vlc[0][0] >> 16);
        total += (unsigned)C * (2 + 6 * (unsigned)n);                    // TSfIdxEncoder (:248-270)
        total += 1 + (unsigned)C * (1 + 2 + 1 + 3 * (unsigned)n);        // EncodeCodeTab (:272-295)
        const unsigned q0 = warp_incl_scan(sh.qbits[0][lane], lane);
        const unsigned q1 = C == 2 ? warp_incl_scan(sh.qbits[1][lane], lane) : 0u;
        total += q0 + q1;
        total += (unsigned)C * 4u * T->sb_to_powgrps[T->qu_to_subband[n - 1]];
        total += (unsigned)tonal_bits;
        const bool fits = total <= (unsigned)kAllocBits;
        const unsigned fm = __ballot_sync(0xffffffffu, fits);
        if (fm >> 31) num_qu = 32;
        else {
            const unsigned low = fm & 0x0fffffffu;                       // candidates 28 .. 1
            num_qu = low ? 32 - __clz((int)low) : 1;
        }
        // the chosen count's word-length table index is needed by the writer
        wl_idx = __shfl_sync(0xffffffffu, wl_idx, num_qu - 1);
        // ---- write: header + parts in Dump order (encode.cpp:116-119)
        unsigned* w = sh.words;
        int p = 0;
        if (lane == 0) {
            p = put_field(w, kFrameBits, p, 0, 1);
            p = put_field(w, kFrameBits, p, (unsigned)(C - 1), 2);
            p = put_field(w, kFrameBits, p, (unsigned)(num_qu - 1), 5);
            p = put_field(w, kFrameBits, p, 0, 1);
            p = put_field(w, kFrameBits, p, 3, 2);
            p = put_field(w, kFrameBits, p, 0, 2);
            p = put_field(w, kFrameBits, p, 0, 2);
            p = put_field(w, kFrameBits, p, (unsigned)wl_idx, 2);
            p = put_field(w, kFrameBits, p, c_alloc[0], 3);
            for (int i = 1; i < num_qu; i++) {
                const unsigned e = T->wl_vlc[wl_idx][((int)c_alloc[i] - (int)c_alloc[i - 1]) & 7];
                p = put_field(w, kFrameBits, p, e & 0xffffu, (int)(e >> 16));
            }
            if (C == 2) {
                p = put_field(w, kFrameBits, p, 1, 2);
                p = put_field(w, kFrameBits, p, 0, 2);
                p = put_field(w, kFrameBits, p, 0, 2);
                const unsigned e = T->wl_vlc[0][0];
                for (int i = 0; i < num_qu; i++) p = put_field(w, kFrameBits, p, e & 0xffffu, (int)(e >> 16));
            }
        }
        p = __shfl_sync(0xffffffffu, p, 0);
        for (int ch = 0; ch < C; ch++) {                                 // scale factor indices
            if (lane < num_qu) put_bits(w, kFrameBits, p + 2 + 6 * lane, 6, sh.sfi[ch][lane]);
            p += 2 + 6 * num_qu;
        }
        p += 1;                                                          // use full table = 1
        if (lane == 0) put_bits(w, kFrameBits, p - 1, 1, 1u);
        for (int ch = 0; ch < C; ch++) {
            if (lane < num_qu) put_bits(w, kFrameBits, p + 4 + 3 * lane, 3, sh.qtab[ch][lane]);
            p += 4 + 3 * num_qu;
        }
        for (int ch = 0; ch < C; ch++) {
            for (int qu = 0; qu < num_qu; qu++) {
                const int start = c_qu_start[qu], len = c_qu_start[qu + 1] - start;
                // (the table's geometry is uniform over the warp: one of three unrolled symbol builders)
                const unsigned sp = T->spec_pack[c_alloc[qu] - 1 + 7 * sh.qtab[ch][qu]];
                const int g = sp & 15u, nc = (sp >> 4) & 15u, cb = (sp >> 8) & 15u;
                const bool sgn = (sp >> 12) & 1u;
                const unsigned* __restrict__ vlc = T->vlc + (sp >> 16);
                const signed char* mq = sh.mant[ch] + start;
                const int nsym = len / nc;
                for (int s0 = 0; s0 < nsym; s0 += 32) {
                    const int sidx = s0 + lane;
                    int nb = 0;
                    unsigned code = 0;
                    if (sidx < nsym)
                        code = nc == 1 ? spec_symbol<1>(vlc, g, cb, sgn, mq, sidx, nb)
                             : nc == 2 ? spec_symbol<2>(vlc, g, cb, sgn, mq, sidx, nb)
                                       : spec_symbol<4>(vlc, g, cb, sgn, mq, sidx, nb);
                    const unsigned inc = warp_incl_scan((unsigned)nb, lane);
                    put_bits(w, kFrameBits, p + (int)(inc - (unsigned)nb), nb, code);
                    p += (int)__shfl_sync(0xffffffffu, inc, 31);
                }
            }
            const int npow = T->sb_to_powgrps[T->qu_to_subband[num_qu - 1]];
            if (lane < npow) put_bits(w, kFrameBits, p + 4 * lane, 4, 15u);
            p += 4 * npow;
        }
        // the tonal part's buffer, appended bit-exactly
        for (int i = lane; i * 32 < tonal_bits; i += 32) {
            const int nb = min(32, tonal_bits - 32 * i);
            put_bits(w, kFrameBits, p + 32 * i, nb, sh.twords[i] >> (32 - nb));
        }
    }
    __syncwarp();
    unsigned char* out = frames + (size_t)unit * kFrameBytes;
    for (int i = lane; i < kFrameBytes / 4; i += 32) {
        const unsigned v = sh.words[i];
        reinterpret_cast<unsigned*>(out)[i] = __byte_perm(v, 0, 0x0123);   // big-endian words -> byte stream
    }
}

void launch_pack(const DevTables* T, const float* specs, const ToneBlock* tones, unsigned char* frames,
                 int units, int C, int fo, int tone_stride, cudaStream_t st)
{
    ATDE_LAUNCH(at3p_pack_kernel, (unsigned)((units + kPackWarps - 1) / kPackWarps), kPackWarps * 32, 0, st,
                T, specs, tones, frames, units, C, fo, tone_stride);
}

} // namespace at3p
} // namespace atde

// =====================================================================================
// Stage entry points (host buffers in, host buffers out): they let tests/ drive each kernel on its own against the
// reference's taps (the whole codec runs through atde_create(ATDE_CODEC_ATRAC3PLUS) / at3p_pipeline.cu).
// Declared in at3p_stage_api.h, not in include/.
// =====================================================================================
namespace {
template <class T> struct ScopedDev {
    T* p = nullptr;
    ~ScopedDev() { if (p) cudaFree(p); }
    bool alloc(size_t n) { return cudaMalloc(&p, n * sizeof(T)) == cudaSuccess; }
};
} // namespace

extern "C" int atde_at3p_stage_pqf(const float* pcm, int S, int C, int F, float* bands)
{
    using namespace atde::at3p;
    if (!device_tables()) return -2;
    const size_t n = (size_t)S * C * F * kFrame;
    ScopedDev<float> d_in, d_out;
    if (!d_in.alloc(n) || !d_out.alloc(n)) return -3;
    if (cudaMemcpy(d_in.p, pcm, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    launch_pqf(d_in.p, nullptr, d_out.p, S, C, F, F, 0, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    return cudaMemcpy(bands, d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

extern "C" int atde_at3p_stage_mdct(const float* resid, int S, int C, int F, float* specs)
{
    using namespace atde::at3p;
    const DevTables* T = device_tables();
    if (!T) return -2;
    const size_t n = (size_t)S * C * F * kFrame;
    ScopedDev<float> d_in, d_out;
    if (!d_in.alloc(n) || !d_out.alloc(n)) return -3;
    if (cudaMemcpy(d_in.p, resid, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    launch_mdct(T, d_in.p, d_out.p, S, C, F, 0, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    return cudaMemcpy(specs, d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

extern "C" int atde_at3p_tone_block_size(void) { return (int)sizeof(atde::at3p::ToneBlock); }

extern "C" int atde_at3p_stage_pack(const float* specs, const void* tones, int units, int C, unsigned char* frames)
{
    using namespace atde::at3p;
    const DevTables* T = device_tables();
    if (!T) return -2;
    const size_t n = (size_t)units * C * kFrame;
    ScopedDev<float> d_in;
    ScopedDev<ToneBlock> d_t;
    ScopedDev<unsigned char> d_out;
    if (!d_in.alloc(n) || !d_t.alloc((size_t)units) || !d_out.alloc((size_t)units * kFrameBytes)) return -3;
    if (cudaMemcpy(d_in.p, specs, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    if (cudaMemcpy(d_t.p, tones, (size_t)units * sizeof(ToneBlock), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    launch_pack(T, d_in.p, d_t.p, d_out.p, units, C, 1, 1, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    return cudaMemcpy(frames, d_out.p, (size_t)units * kFrameBytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

extern "C" int atde_at3p_stage_tone_filter(const float* bands, const void* tb_old, const void* tb_now, const void* tb_next,
                                           int units, int C, float* resid)
{
    using namespace atde::at3p;
    const DevTables* T = device_tables();
    if (!T) return -2;
    const size_t n = (size_t)units * C * kFrame;
    ScopedDev<float> d_in, d_out;
    ScopedDev<ToneBlock> d_t[3];
    const void* src[3] = {tb_old, tb_now, tb_next};
    if (!d_in.alloc(n) || !d_out.alloc(n)) return -3;
    for (int k = 0; k < 3; k++) {
        if (!d_t[k].alloc((size_t)units)) return -3;
        if (cudaMemcpy(d_t[k].p, src[k], (size_t)units * sizeof(ToneBlock), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    }
    if (cudaMemcpy(d_in.p, bands, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    // flat test layout [U][C][2048]: one "stream" per unit with a single frame
    FilterLayout lay;
    lay.fo = 1; lay.tone_stride = 1; lay.in_frames = 1; lay.in_off = 0; lay.out_frames = 1; lay.out_off = 0;
    launch_tone_filter(T, d_in.p, d_t[0].p, d_t[1].p, d_t[2].p, d_out.p, units, C, lay, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    return cudaMemcpy(resid, d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

// ---- device libm replicas under test (tests/test_gpu_math.py) ----
namespace {
__global__ void trig_kernel(int fn, const double* x, double* y, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    if (fn == 0) y[i] = atde::g_sin(v);
    else if (fn == 1) y[i] = atde::g_cos(v);
    else if (fn == 2) y[i] = atde::g_atan(v);
    else {
        float s, c;
        atde::g_sincosf((float)v, s, c);
        y[i] = fn == 3 ? (double)s : (double)c;
    }
}
} // namespace

/* fn: 0 sin, 1 cos, 2 atan (double in, double out); 3 sinf, 4 cosf of sincosf((float)x) widened to double */
extern "C" int atde_at3p_debug_trig(int fn, const double* x, double* y, long long n)
{
    ScopedDev<double> dx, dy;
    if (n <= 0 || !dx.alloc((size_t)n) || !dy.alloc((size_t)n)) return -3;
    if (cudaMemcpy(dx.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    ATDE_LAUNCH(trig_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t) nullptr, fn, (const double*)dx.p, dy.p, n);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    return cudaMemcpy(y, dy.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
