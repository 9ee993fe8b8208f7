// atde_api.cu — C ABI (include/atde_b200.h) over the sm_100a kernels.
//
// Host-side plumbing only: device buffers, table upload, stream/chunk pipeline, error mapping.
// There is no CPU implementation of the encode path in this library.
#include "../../include/atde_b200.h"
#include "atde_cuda.h"
#include "at1_kernels.cuh"
#include "at3_kernels.cuh"
#include "at3p_kernels.cuh"
#include "glibc_math.cuh"
#include "host_tables.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ATDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;     // elements
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) return fail(ATDE_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
        cap = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Per-chunk scratch (one per pipeline slot)
struct Workspace {
    DevBuf<float> pcm;              // host path only
    DevBuf<short> pcm16;            // host path, int16 ingest only
    DevBuf<float> specs;
    DevBuf<unsigned char> masks;
    DevBuf<float> chloud;
    DevBuf<float> loud;
    DevBuf<unsigned char> out;      // host path only
    DevBuf<int> sizes;              // host path only
    DevBuf<unsigned char> tap_sfi, tap_wl;
    // ATRAC3
    DevBuf<float> bands, gain, gstat, gprev, gscale, energy, hist_tmp;
    DevBuf<float> trace_gain, trace_stat;   // gain-control trace only (atde_set_gain_trace)
    DevBuf<atde::at3::Curve> curves;
    DevBuf<atde::at3::TonalList> tonal;
    DevBuf<unsigned char> sfi;
    cudaStream_t stream = nullptr;
    void release()
    {
        pcm.release(); pcm16.release(); specs.release(); masks.release(); chloud.release(); loud.release();
        out.release(); sizes.release(); tap_sfi.release(); tap_wl.release();
        bands.release(); gain.release(); gstat.release(); gprev.release(); gscale.release();
        energy.release(); hist_tmp.release(); curves.release(); tonal.release(); sfi.release();
        trace_gain.release(); trace_stat.release();
    }
};

} // namespace

struct atde_encoder {
    atde_settings cfg;
    int frame_samples = 0, units_per_frame = 0, unit_bytes = 0, lookahead = 0;
    atde::at1::DevTables* d_at1_tab = nullptr;
    atde::at3::DevTables* d_at3_tab = nullptr;
    int at3_js = 0;
    atde::at3p::StreamState* at3p = nullptr;    // ATRAC3plus: carried stream state + workspaces (at3p_pipeline.cu)
    // ATRAC3 stream state beyond hist / loud_state / started
    DevBuf<float> prevhalf, next_scale, ctx;
    static constexpr int kSlots = 3;   // pipeline slots of the host path: chunk k+2 is copied in while k+1 waits and k computes
    Workspace ws[kSlots];
    cudaEvent_t chunk_done[kSlots] = {};   // host path: "the kernels of the chunk in this slot are finished"
    // ATRAC3 host path: PCM arrives through its own ring of staging buffers on its own stream, so that the copy-in of
    // chunk k+4 waits only for the FIRST kernel of chunk k (the QMF, the one reader of the PCM), not for the chunk's
    // whole pipeline: the H2D copies then run back to back from the start of the call
    static constexpr int kPcmRing = 4;
    DevBuf<float> pcm_ring[kPcmRing];
    DevBuf<short> pcm16_ring[kPcmRing];
    cudaEvent_t pcm_ready[kPcmRing] = {}, pcm_free[kPcmRing] = {};
    cudaStream_t copy_stream = nullptr;
    // stream state, sized for n_state_streams
    DevBuf<float> hist;
    DevBuf<float> loud_state;
    DevBuf<unsigned char> started;
    int n_state_streams = 0;
    bool have_state = false;
    bool streams_started = false;   // at least one batch went through since create/reset
    long long last_out = 0;         // output frames per stream of the last batch
    bool taps_enabled = false;
    bool gain_trace = false;        // ATRAC3: also run the trace instance of the gain kernel (atde_set_gain_trace)
    // geometry of the last batch, for taps
    int last_S = 0; long long last_F = 0;
    long long launches = 0;
    // optional per-kernel CUDA-event timing (atde_set_profiling)
    bool profiling = false;
    struct EvPair { cudaEvent_t a, b; int kind; };
    std::vector<EvPair> ev_pool;
    size_t ev_used = 0;
};

namespace {

int build_at1_tables(atde_encoder* e)
{
    using namespace atde;
    at1::DevTables* h = new (std::nothrow) at1::DevTables();
    if (!h) return fail(ATDE_ERR_NOMEM, "host alloc");
    float qmf[48];
    qmf_window(qmf);
    for (uint32_t i = 0; i < 32; i++)                                   // atrac1.h:128-132
        h->sine_window[i] = sin((i + 0.5) * (M_PI / (2.0 * 32.0)));
    for (uint32_t i = 0; i < 64; i++)                                   // atrac1.h:122-127
        h->scale_table[i] = pow(2.0, (double)(i / 3.0 - 21.0));
    for (int t = 0; t < 288; t++) {                                     // weights of TAtrac1MDCT::Mdct's input (atrac1denc.cpp:80-90)
        if (t < 160) h->win_long128[t] = t < 32 ? h->sine_window[t] : (t < 128 ? 1.0f : h->sine_window[128 + 31 - t]);
        h->win_long256[t] = t < 32 ? h->sine_window[t] : (t < 256 ? 1.0f : h->sine_window[256 + 31 - t]);
        if (t < 64) h->win_short[t] = t < 32 ? h->sine_window[t] : h->sine_window[63 - t];
    }
    const std::vector<float> curve = loudness_curve(512);
    memcpy(h->loud_curve, curve.data(), sizeof(h->loud_curve));
    {   // CalcAt1ATH (atrac1_bitalloc.cpp:118-135): min over the BFU's lines, dB -> power
        static const unsigned short start_long[52] = {
            0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116, 122,
            128, 134, 140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226, 236, 246,
            256, 268, 280, 292, 304, 316, 328, 340, 352, 372, 392, 412, 432, 452, 472, 492};
        static const unsigned char per_block[52] = {
            8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6,
            6, 6, 6, 6, 7, 7, 7, 7, 9, 9, 9, 9, 10, 10, 10, 10,
            12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20, 20, 20, 20, 20, 20};
        const std::vector<float> ath = calc_ath(512, 44100);
        for (int b = 0; b < 52; b++) {
            float x = 999;
            for (size_t line = start_long[b]; line < (size_t)start_long[b] + per_block[b]; line++)
                x = fmin(x, ath[line]);
            x = pow(10, 0.1 * x);
            h->ath_long[b] = x;
        }
    }
    const std::vector<float> sc512 = mdct_sincos(512, 1), sc256 = mdct_sincos(256, 0.5), sc64 = mdct_sincos(64, 0.5);
    memcpy(h->sincos512, sc512.data(), sizeof(h->sincos512));
    memcpy(h->sincos256, sc256.data(), sizeof(h->sincos256));
    memcpy(h->sincos64, sc64.data(), sizeof(h->sincos64));
    const auto tw128 = kiss_twiddles(128, false), tw64 = kiss_twiddles(64, false), tw16 = kiss_twiddles(16, false);
    memcpy(h->tw128, tw128.data(), sizeof(h->tw128));
    memcpy(h->tw64, tw64.data(), sizeof(h->tw64));
    memcpy(h->tw16, tw16.data(), sizeof(h->tw16));
    const auto p128 = kiss_perm(128), p64 = kiss_perm(64), p16 = kiss_perm(16);
    for (int i = 0; i < 128; i++) h->perm128[i] = (unsigned char)p128[i];
    for (int i = 0; i < 64; i++) h->perm64[i] = (unsigned char)p64[i];
    for (int i = 0; i < 16; i++) h->perm16[i] = (unsigned char)p16[i];
    for (int kind = 0; kind < 3; kind++) {
        // kind 0 / 1: long block of a 128-sample (low, mid: MDCT-256) / 256-sample band (hi: MDCT-512); kind 2: a short
        // block (MDCT-64).  TMDCT's fold (mdct.h:56-76) reads in[n34-1-n], in[n34+n | n-n4], in[n4+n], in[n4-1-n | n54-1-n];
        // TAtrac1MDCT::Mdct's buffer is the windowed stretch (32 samples before the block + the block; a short block's
        // stretch is 64 samples) at offset joff, zeros around it (atrac1denc.cpp:80-90)
        const int N = kind == 0 ? 256 : (kind == 1 ? 512 : 64), n4 = N / 4, n34 = 3 * n4, n54 = 5 * n4;
        const int joff = kind == 0 ? 48 : (kind == 1 ? 112 : 0), range = kind == 2 ? 64 : N / 2 + 32;
        unsigned long long* tab = kind == 0 ? h->fold128 : (kind == 1 ? h->fold256 : h->fold64);
        for (int slot = 0; slot < N / 4; slot++) {
            const int i = kind == 0 ? p64[slot] : (kind == 1 ? p128[slot] : p16[slot]), n = 2 * i;
            const int j[4] = {n34 - 1 - n, n < n4 ? n34 + n : n - n4, n4 + n, n < n4 ? n4 - 1 - n : n54 - 1 - n};
            unsigned long long e = (unsigned long long)i << 40;
            for (int q = 0; q < 4; q++) {
                const int t = j[q] - joff;
                e |= (unsigned long long)((t >= 0 && t < range) ? t : 1023) << (10 * q);
            }
            tab[slot] = e;
        }
    }
    // the analysis kernel runs the first two FFT stages of long and short blocks with one table (at1_kernels.cu: mdct_frame)
    for (int k = 0; k < 16; k++)
        if (memcmp(&tw64[4 * k], &tw16[k], sizeof(tw16[k])) != 0) { delete h; return fail(ATDE_ERR_INVALID, "kissfft twiddles of 64 and 16 points disagree"); }

    cudaError_t ce = cudaMalloc(&e->d_at1_tab, sizeof(at1::DevTables));
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_at1_tab, h, sizeof(at1::DevTables), cudaMemcpyHostToDevice);
    delete h;
    if (ce != cudaSuccess) return fail(ATDE_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(ce));
    at1::upload_qmf_window(qmf);
    return 0;
}


// Container parameters, atrac3.h:211-220; GetContainerParamsForBitrate (atrac3.cpp:45-51) is a
// lower_bound on Bitrate with 0 meaning LP2.
struct At3Container { unsigned bitrate; int frame_sz; int js; };
const At3Container kAt3Containers[8] = {
    {66150, 192, 1}, {93713, 272, 1}, {104738, 304, 0}, {132300, 384, 0},
    {146081, 424, 0}, {176400, 512, 0}, {264600, 768, 0}, {352800, 1024, 0}};

const At3Container* at3_container_for(unsigned bitrate)
{
    if (bitrate == 0) bitrate = 132300;
    for (int i = 0; i < 8; i++)
        if (!(kAt3Containers[i].bitrate < bitrate)) return &kAt3Containers[i];
    return nullptr;                                   // past the end: the reference would read out of bounds
}

int build_at3_tables(atde_encoder* e)
{
    using namespace atde;
    at3::DevTables* h = new (std::nothrow) at3::DevTables();
    if (!h) return fail(ATDE_ERR_NOMEM, "host alloc");
    memset(h, 0, sizeof(*h));
    float qmf[48];
    qmf_window(qmf);
    // TAtrac3Data::TAtrac3Data (atrac3.h:178-198)
    for (uint32_t i = 0; i < 64; i++) h->scale_table[i] = pow(2.0, (double)(i / 3.0 - 21.0));
    for (int i = 0; i < 256; i++) h->encode_window[i] = (sin(((i + 0.5) / 256.0 - 0.5) * M_PI) + 1.0);
    for (int i = 0; i < 16; i++) h->gain_level[i] = pow(2.0, 4 - i);
    for (int i = 0; i < 31; i++) h->gain_interp[i] = pow(2.0, -1.0 / 8 * (i - 15));
    const std::vector<float> curve = loudness_curve(1024);
    memcpy(h->loud_curve, curve.data(), sizeof(h->loud_curve));
    {   // TAtrac3BitStreamWriter ctor (atrac3_bitstream.cpp:705-717)
        static const unsigned short start[33] = {
            0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176,
            192, 224, 256, 288, 320, 352, 384, 416, 448, 480, 512, 576, 640, 704, 768, 896, 1024};
        const std::vector<float> ath = calc_ath(1024, 44100);
        for (int b = 0; b < 32; b++) {
            float x = 999;
            for (size_t line = start[b]; line < start[b + 1]; line++) x = fmin(x, ath[line]);
            x = pow(10, 0.1f * x);
            h->ath[b] = x;
        }
    }
    const std::vector<float> sc512 = mdct_sincos(512, 1);
    memcpy(h->sincos512, sc512.data(), sizeof(h->sincos512));
    const auto tw128 = kiss_twiddles(128, false);
    memcpy(h->tw128, tw128.data(), sizeof(h->tw128));
    const auto p128 = kiss_perm(128);
    for (int i = 0; i < 128; i++) h->perm128[i] = (unsigned char)p128[i];
    {   // TSpectralUpsampler(11025.0f, 800.0f, 0.15f) (transient_spectral_upsampler.cpp:32-69)
        const float sampleRate = 11025.0f, lowCutHz = 800.0f, epsilon = 0.15f;
        const int kInN = 512;
        h->low_cut_bin = static_cast<int>(std::ceil(lowCutHz * kInN / sampleRate));
        const float eN = epsilon * static_cast<float>(kInN);
        const float fN = static_cast<float>(kInN);
        for (int n = 0; n < kInN; ++n) {
            const float fn = static_cast<float>(n);
            if (n == 0) {
                h->planck[n] = 0.0f;
            } else if (fn < eN) {
                const float Zp = eN * (1.0f / fn + 1.0f / (fn - eN));
                h->planck[n] = 1.0f / (1.0f + std::exp(Zp));
            } else if (fn <= fN - eN) {
                h->planck[n] = 1.0f;
            } else {
                const float m = fN - fn;
                const float Zp = eN * (1.0f / m + 1.0f / (m - eN));
                h->planck[n] = 1.0f / (1.0f + std::exp(Zp));
            }
        }
        for (int i = 1; i < 3; ++i)                                      // :110-111,152
            h->hpf_h[i - 1] = 0.5f * (1.0f - std::cos(static_cast<float>(M_PI) * i / 2.0f));
    }
    const auto tw256 = kiss_twiddles(256, false);
    memcpy(h->tw256, tw256.data(), sizeof(h->tw256));
    const auto s512 = kiss_super_twiddles(512, false);
    memcpy(h->super512, s512.data(), sizeof(h->super512));
    const auto p256 = kiss_perm(256);
    for (int i = 0; i < 256; i++) h->perm256[i] = (unsigned char)p256[i];
    const auto tw2048 = kiss_twiddles(2048, true);
    memcpy(h->tw2048, tw2048.data(), sizeof(h->tw2048));
    const auto s4096 = kiss_super_twiddles(4096, true);
    memcpy(h->super4096, s4096.data(), sizeof(h->super4096));
    const auto p2048 = kiss_perm(2048);
    for (int o = 0; o < 2048; o++) h->iperm2048[p2048[o]] = (unsigned short)o;
    memset(h->ftw, 0, sizeof(h->ftw));
    for (int st = 0; st < 4; st++) {
        const int m = 1 << (2 * st), fstride = 64 / m;
        for (int q = 0; q < 3; q++)
            for (int k = 0; k < m; k++) memcpy(&h->ftw[st][q][k], &tw256[(size_t)(q + 1) * k * fstride], sizeof(at3::DevTables::ftw[0][0][0]));
    }
    for (int k = 0; k < 8; k++) {
        for (int q = 0; q < 3; q++) {
            h->gtw2[q][k] = h->tw2048[(size_t)64 * k * (q + 1)];
            for (int a = 0; a < 4; a++)
                h->gtw2[3 + 3 * a + q][k] = h->tw2048[(size_t)16 * (k + 8 * a) * (q + 1)];
        }
    }
    for (int k = 0; k < 128; k++)
        for (int q = 0; q < 3; q++) {
            h->gtw3a[q][k] = h->tw2048[(size_t)4 * k * (q + 1)];
            for (int a = 0; a < 4; a++)
                h->gtw3b[a][q][k] = h->tw2048[(size_t)(k + 128 * a) * (q + 1)];
        }

    cudaError_t ce = cudaMalloc(&e->d_at3_tab, sizeof(at3::DevTables));
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_at3_tab, h, sizeof(at3::DevTables), cudaMemcpyHostToDevice);
    delete h;
    if (ce != cudaSuccess) return fail(ATDE_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(ce));
    at3::upload_qmf_window(qmf);
    return 0;
}

struct KernelTimer {
    atde_encoder* e; cudaStream_t st; int idx = -1;
    KernelTimer(atde_encoder* e_, cudaStream_t st_, int kind) : e(e_), st(st_)
    {
        if (!e->profiling) return;
        if (e->ev_used == e->ev_pool.size()) {
            atde_encoder::EvPair p;
            if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
            e->ev_pool.push_back(p);
        }
        idx = (int)e->ev_used++;
        e->ev_pool[idx].kind = kind;
        cudaEventRecord(e->ev_pool[idx].a, st);
    }
    ~KernelTimer() { if (idx >= 0) cudaEventRecord(e->ev_pool[idx].b, st); }
};

int at3p_prof_begin(void* ctx, cudaStream_t st, int kind)
{
    atde_encoder* e = (atde_encoder*)ctx;
    if (!e->profiling) return -1;
    if (e->ev_used == e->ev_pool.size()) {
        atde_encoder::EvPair p;
        if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return -1;
        e->ev_pool.push_back(p);
    }
    const int idx = (int)e->ev_used++;
    e->ev_pool[idx].kind = kind;
    cudaEventRecord(e->ev_pool[idx].a, st);
    return idx;
}
void at3p_prof_end(void* ctx, cudaStream_t st, int idx)
{
    atde_encoder* e = (atde_encoder*)ctx;
    if (idx >= 0) cudaEventRecord(e->ev_pool[idx].b, st);
}

// Runs the ATRAC1 pipeline for `S` streams x F frames whose PCM is at d_pcm (device), first
// stream index s0 inside the handle's state arrays.
int run_at1(atde_encoder* e, Workspace& w, const float* d_pcm, int s0, int S, long long F,
            unsigned char* d_out, int* d_sizes)
{
    using namespace atde::at1;
    const int C = e->cfg.channels;
    const size_t units = (size_t)S * F * C;
    int rc;
    if ((rc = w.specs.ensure(units * 512))) return rc;
    if ((rc = w.masks.ensure(units))) return rc;
    if ((rc = w.chloud.ensure(units))) return rc;
    if ((rc = w.loud.ensure((size_t)S * F))) return rc;
    if (e->taps_enabled) {
        if ((rc = w.tap_sfi.ensure(units * 52))) return rc;
        if ((rc = w.tap_wl.ensure(units * 52))) return rc;
    }
    AnalysisParams a;
    a.pcm = d_pcm;
    a.hist = e->hist.p + (size_t)s0 * 512 * C;
    a.started = e->started.p + s0;
    a.specs = w.specs.p; a.masks = w.masks.p; a.chloud = w.chloud.p;
    a.tab = e->d_at1_tab;
    a.S = S; a.C = C; a.F = (int)F;
    a.window_auto = e->cfg.window_mode == 1;
    a.window_mask = (int)e->cfg.window_mask;
    a.one = 1.0f;
    { KernelTimer kt(e, w.stream, 0); launch_analysis(a, w.stream); }

    LoudnessParams l;
    l.masks = w.masks.p; l.chloud = w.chloud.p;
    l.loud_in = e->loud_state.p + s0;
    l.loud = w.loud.p;
    l.S = S; l.C = C; l.F = (int)F;
    { KernelTimer kt(e, w.stream, 1); launch_loudness(l, w.stream); }

    PackParams k;
    k.specs = w.specs.p; k.masks = w.masks.p; k.loud = w.loud.p;
    k.out = d_out; k.sizes = d_sizes;
    k.tap_sfi = e->taps_enabled ? w.tap_sfi.p : nullptr;
    k.tap_wl = e->taps_enabled ? w.tap_wl.p : nullptr;
    k.tab = e->d_at1_tab;
    k.S = S; k.C = C; k.F = (int)F;
    k.bfu_idx_const = (int)e->cfg.bfu_idx_const;
    { KernelTimer kt(e, w.stream, 2); launch_pack(k, w.stream); }

    CarryParams c;
    c.pcm = d_pcm; c.loud = w.loud.p;
    c.hist = e->hist.p + (size_t)s0 * 512 * C;
    c.loud_state = e->loud_state.p + s0;
    c.started = e->started.p + s0;
    c.S = S; c.C = C; c.F = (int)F;
    launch_carry(c, w.stream);
    e->launches += 5;
    CK(cudaGetLastError());
    return 0;
}

// Runs the ATRAC3 pipeline for S streams x N new frames whose PCM is at d_pcm (device); s0 = first
// stream index inside the handle's state arrays; `started` says whether the streams carry a frame.
int run_at3(atde_encoder* e, Workspace& w, const float* d_pcm, int s0, int S, long long N, bool started,
            unsigned char* d_out, cudaEvent_t pcm_consumed = nullptr)
{
    using namespace atde::at3;
    const int C = e->cfg.channels;
    Geometry g;
    g.S = S; g.C = C; g.N = (int)N;
    g.L = (int)N + (started ? 1 : 0);
    g.n_out = g.L - 1;
    g.BL = 128 + 256 * g.L;
    g.js = e->at3_js && C == 2;
    g.js_mono = e->at3_js && C == 1;
    g.frame_sz = e->unit_bytes;
    g.no_gain = e->cfg.no_gain_control != 0;
    g.no_tonal = e->cfg.no_tonal != 0;
    g.bfu_idx_const = (int)e->cfg.bfu_idx_const;
    g.one = 1.0f;
    // (work areas are sized for the continuation case — N output frames, one carried frame in the band buffer — also on
    // a stream's first batch, which produces one frame less: otherwise the second call of a run re-allocates every
    // buffer in mid-pipeline, and a cudaMalloc synchronises the device)
    const size_t units = (size_t)S * (size_t)(N > 0 ? N : 1) * C;
    const size_t items = (size_t)S * C * kGainBands * (size_t)(N > 0 ? N : 1);
    int rc;
    if ((rc = w.bands.ensure((size_t)S * C * 4 * (128 + 256 * ((size_t)N + 1))))) return rc;
    if ((rc = w.hist_tmp.ensure((size_t)S * (2 * 1024 * C + C * 4 * 256 + C * 4)))) return rc;
    if ((rc = w.specs.ensure(units * 1024))) return rc;
    if ((rc = w.gscale.ensure(units * 16))) return rc;
    if ((rc = w.chloud.ensure(units))) return rc;
    if ((rc = w.loud.ensure((size_t)S * (N > 0 ? N : 1)))) return rc;
    if ((rc = w.sfi.ensure(units * 32))) return rc;
    if ((rc = w.energy.ensure(units * 32))) return rc;
    if ((rc = w.tonal.ensure(units))) return rc;
    if ((rc = w.curves.ensure((size_t)S * C * 4 * (N > 0 ? N : 1)))) return rc;
    if (!g.no_gain) {
        if ((rc = w.gain.ensure(items * 96))) return rc;
        if ((rc = w.gstat.ensure(items * 4))) return rc;
        if ((rc = w.gprev.ensure(items * 4))) return rc;
    }
    if (e->taps_enabled && (rc = w.tap_wl.ensure(units * 32))) return rc;
    const bool trace = e->gain_trace && !g.no_gain;
    if (trace) {
        const size_t titems = (size_t)S * C * kBands * (size_t)(N > 0 ? N : 1);
        if ((rc = w.trace_gain.ensure(titems * 96))) return rc;
        if ((rc = w.trace_stat.ensure(titems * 4))) return rc;
    }

    Buffers b;
    memset(&b, 0, sizeof(b));
    b.pcm = d_pcm;
    b.started = e->started.p + s0;
    b.hist_tmp = w.hist_tmp.p;
    b.prevhalf_out = w.hist_tmp.p + (size_t)S * 2 * 1024 * C;
    b.next_scale_out = b.prevhalf_out + (size_t)S * C * 4 * 256;
    b.pcm_hist = e->hist.p + (size_t)s0 * 2 * 1024 * C;
    b.prevhalf = e->prevhalf.p + (size_t)s0 * C * 4 * 256;
    b.next_scale = e->next_scale.p + (size_t)s0 * C * 4;
    b.ctx = e->ctx.p + (size_t)s0 * C * kGainBands * 4;
    b.loud_state = e->loud_state.p + s0;
    b.bands = w.bands.p; b.gain = w.gain.p; b.gstat = w.gstat.p; b.gprev = w.gprev.p;
    b.curves = w.curves.p; b.specs = w.specs.p; b.gscale = w.gscale.p; b.chloud = w.chloud.p;
    b.loud = w.loud.p; b.sfi = w.sfi.p; b.energy = w.energy.p; b.tonal = w.tonal.p;
    b.out = d_out;
    b.tap_prec = e->taps_enabled ? w.tap_wl.p : nullptr;
    b.trace_gain = trace ? w.trace_gain.p : nullptr;
    b.trace_stat = trace ? w.trace_stat.p : nullptr;
    b.tab = e->d_at3_tab;

    { KernelTimer kt(e, w.stream, 0); launch_qmf(g, b, w.stream); }
    e->launches += 1;
    if (pcm_consumed) CK(cudaEventRecord(pcm_consumed, w.stream));     // nothing after the QMF reads the PCM
    if (g.n_out > 0) {
        if (!g.no_gain) {
            { KernelTimer kt(e, w.stream, 3); launch_gain_analysis(g, b, w.stream); }
            { KernelTimer kt(e, w.stream, 4); launch_gain_scan(g, b, w.stream); launch_gain_curve(g, b, w.stream); }
            e->launches += 3;
            if (trace) { launch_gain_trace(g, b, w.stream); e->launches += 1; }
        }
        { KernelTimer kt(e, w.stream, 0); launch_mdct(g, b, w.stream); }
        { KernelTimer kt(e, w.stream, 1); launch_loudterm(g, b, w.stream); launch_loudness(g, b, w.stream); }
        { KernelTimer kt(e, w.stream, 5); launch_scale_tonal(g, b, w.stream); }
        { KernelTimer kt(e, w.stream, 2); launch_alloc_pack(g, b, w.stream); }
        e->launches += 5;
    }
    launch_carry(g, b, w.stream);
    e->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

__global__ void init_state_kernel(float* loud, unsigned char* started, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { loud[i] = atde::at1::kLoudFactor; started[i] = 0; }
}

__global__ void fill_kernel(float* p, float v, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int ensure_state(atde_encoder* e, int S)
{
    if (e->have_state && e->n_state_streams == S) return 0;
    if (e->have_state && e->n_state_streams != S)
        return fail(ATDE_ERR_INVALID, "batch has %d streams but the handle carries state for %d; call atde_reset() first", S, e->n_state_streams);
    int rc;
    const int C = e->cfg.channels;
    const bool at3 = e->cfg.codec == ATDE_CODEC_ATRAC3;
    if (e->cfg.codec == ATDE_CODEC_ATRAC3PLUS) {              // carried state lives in at3p_pipeline.cu
        atde::at3p::pipeline_reset(e->at3p);
        e->streams_started = false;
        e->n_state_streams = S;
        e->have_state = true;
        return 0;
    }
    if ((rc = e->hist.ensure((size_t)S * e->frame_samples * C * (at3 ? 2 : 1)))) return rc;
    if ((rc = e->loud_state.ensure(S))) return rc;
    if ((rc = e->started.ensure(S))) return rc;
    ATDE_LAUNCH(init_state_kernel, (S + 255) / 256, 256, 0, e->ws[0].stream, e->loud_state.p, e->started.p, S);
    e->launches += 1;
    if (at3) {
        // PcmBuffer zero-initialised (delay_buffer.h:28-31), PrevOverlapGainScale = 1 (atrac3denc.cpp:100-101),
        // CurveCtx = {} (atrac3denc.h:113)
        const size_t nh = (size_t)S * C * 4 * 256, ns = (size_t)S * C * 4, nc = (size_t)S * C * atde::at3::kGainBands * 4;
        if ((rc = e->prevhalf.ensure(nh))) return rc;
        if ((rc = e->next_scale.ensure(ns))) return rc;
        if ((rc = e->ctx.ensure(nc))) return rc;
        CK(cudaMemsetAsync(e->prevhalf.p, 0, nh * sizeof(float), e->ws[0].stream));
        CK(cudaMemsetAsync(e->ctx.p, 0, nc * sizeof(float), e->ws[0].stream));
        ATDE_LAUNCH(fill_kernel, (unsigned)((ns + 255) / 256), 256, 0, e->ws[0].stream, e->next_scale.p, 1.0f, (long long)ns);
        e->launches += 1;
    }
    e->streams_started = false;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->ws[0].stream));
    e->n_state_streams = S;
    e->have_state = true;
    return 0;
}

} // namespace

namespace {
// int16 PCM -> the normalised float the reference's reader hands to the PCM engine: libsndfile's sf_readf_float
// on a PCM_16 file multiplies by 1.0 / 0x8000 (src/pcm_io_sndfile.cpp:111-113 -> SndfileHandle::readf(float*)),
// exact in fp32.  8 samples per thread (one 16-byte load, two 16-byte stores).
__global__ void pcm_i16_to_f32_kernel(const short* __restrict__ in, float* __restrict__ out, long long n)
{
    const long long n8 = n >> 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 v = reinterpret_cast<const uint4*>(in)[i];
        const int w[4] = {(int)v.x, (int)v.y, (int)v.z, (int)v.w};
        float f[8];
        for (int k = 0; k < 4; k++) {
            f[2 * k] = __fmul_rn((float)(short)(w[k] & 0xffff), 1.0f / 32768.0f);
            f[2 * k + 1] = __fmul_rn((float)(short)(w[k] >> 16), 1.0f / 32768.0f);
        }
        reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fmul_rn((float)in[i], 1.0f / 32768.0f);
}

} // namespace

extern "C" {

const char* atde_last_error(void) { return g_err.c_str(); }
const char* atde_version(void) { return "atde_b200 0.1 (sm_100a)"; }

void atde_default_settings(atde_settings* s, int32_t codec, int32_t channels)
{
    memset(s, 0, sizeof(*s));
    s->codec = codec;
    s->channels = channels;
    s->window_mode = 1;
    s->gha_flags = 7;                       // TAt3PEnc::TSettings(): UseGha = GHA_ENABLED
}

int atde_create(const atde_settings* s, atde_encoder** out)
{
    if (!s || !out) return fail(ATDE_ERR_INVALID, "null argument");
    *out = nullptr;
    if (s->channels < 1 || s->channels > 2) return fail(ATDE_ERR_INVALID, "channels must be 1 or 2");
    if (s->codec != ATDE_CODEC_ATRAC1 && s->codec != ATDE_CODEC_ATRAC3 && s->codec != ATDE_CODEC_ATRAC3PLUS)
        return fail(ATDE_ERR_UNSUPPORTED, "codec %d is not built yet", s->codec);
    const At3Container* cont = nullptr;
    if (s->codec == ATDE_CODEC_ATRAC1) {
        if (s->bfu_idx_const > 8) return fail(ATDE_ERR_INVALID, "bfu_idx_const must be 0..8");
    } else if (s->codec == ATDE_CODEC_ATRAC3PLUS) {
        // TAt3PEnc::TSettings (src/atrac3p.h:29-57): the three processing flags; the wideband experiment is not built
        if (s->gha_flags & ~7u)
            return fail(ATDE_ERR_UNSUPPORTED, "ATRAC3plus GHA_WIDEBAND (gha_flags bit 3) is not built");
    } else {
        if (s->bfu_idx_const > 32) return fail(ATDE_ERR_INVALID, "bfu_idx_const must be 0..32");
        cont = at3_container_for(s->bitrate);
        if (!cont) return fail(ATDE_ERR_INVALID, "bitrate %u is above the largest ATRAC3 container (352800)", s->bitrate);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(ATDE_ERR_CUDA, "no CUDA device: libatde_b200 has no CPU fallback");
    CK(cudaSetDevice(s->device));
    atde_encoder* e = new (std::nothrow) atde_encoder();
    if (!e) return fail(ATDE_ERR_NOMEM, "host alloc");
    e->cfg = *s;
    if (s->codec == ATDE_CODEC_ATRAC1) {
        e->frame_samples = 512;
        e->units_per_frame = s->channels;
        e->unit_bytes = atde::at1::kUnitBytes;
        e->lookahead = 0;
    } else if (s->codec == ATDE_CODEC_ATRAC3PLUS) {
        e->frame_samples = atde::at3p::kFrame;
        e->units_per_frame = 1;
        e->unit_bytes = atde::at3p::kFrameBytes;
        e->lookahead = 1;
        e->at3p = atde::at3p::pipeline_create(s->channels, (int)s->gha_flags);
        if (!e->at3p) { delete e; return fail(ATDE_ERR_NOMEM, "host alloc"); }
    } else {
        e->frame_samples = 1024;
        e->units_per_frame = 1;
        e->unit_bytes = cont->frame_sz;
        e->lookahead = 1;
        e->at3_js = cont->js;
    }
    for (int i = 0; i < atde_encoder::kSlots; i++) {
        cudaError_t ce = cudaEventCreateWithFlags(&e->chunk_done[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->ws[i].stream, cudaStreamNonBlocking);
        if (ce != cudaSuccess) { delete e; return fail(ATDE_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)); }
    }
    if (s->codec == ATDE_CODEC_ATRAC3) {
        cudaError_t ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
        for (int i = 0; i < atde_encoder::kPcmRing && ce == cudaSuccess; i++) {
            ce = cudaEventCreateWithFlags(&e->pcm_ready[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->pcm_free[i], cudaEventDisableTiming);
        }
        if (ce != cudaSuccess) { atde_destroy(e); return fail(ATDE_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)); }
    }
    int rc = 0;
    if (s->codec == ATDE_CODEC_ATRAC1) rc = build_at1_tables(e);
    else if (s->codec == ATDE_CODEC_ATRAC3) rc = build_at3_tables(e);
    else if (!atde::at3p::device_tables() || !atde::at3p::gha_tables_ready()) rc = fail(ATDE_ERR_CUDA, "ATRAC3plus table upload failed");
    if (rc) { atde_destroy(e); return rc; }
    *out = e;
    return 0;
}

void atde_destroy(atde_encoder* e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < atde_encoder::kSlots; i++) {
        e->ws[i].release();
        if (e->ws[i].stream) cudaStreamDestroy(e->ws[i].stream);
        if (e->chunk_done[i]) cudaEventDestroy(e->chunk_done[i]);
    }
    for (int i = 0; i < atde_encoder::kPcmRing; i++) {
        e->pcm_ring[i].release(); e->pcm16_ring[i].release();
        if (e->pcm_ready[i]) cudaEventDestroy(e->pcm_ready[i]);
        if (e->pcm_free[i]) cudaEventDestroy(e->pcm_free[i]);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    e->hist.release(); e->loud_state.release(); e->started.release();
    e->prevhalf.release(); e->next_scale.release(); e->ctx.release();
    atde::at3p::pipeline_destroy(e->at3p);
    if (e->d_at1_tab) cudaFree(e->d_at1_tab);
    if (e->d_at3_tab) cudaFree(e->d_at3_tab);
    for (auto& p : e->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    delete e;
}

int atde_frame_samples(const atde_encoder* e) { return e ? e->frame_samples : ATDE_ERR_INVALID; }
int atde_units_per_frame(const atde_encoder* e) { return e ? e->units_per_frame : ATDE_ERR_INVALID; }
int atde_unit_bytes(const atde_encoder* e) { return e ? e->unit_bytes : ATDE_ERR_INVALID; }
int atde_lookahead_frames(const atde_encoder* e) { return e ? e->lookahead : ATDE_ERR_INVALID; }
void* atde_cuda_stream(atde_encoder* e) { return e ? (void*)e->ws[0].stream : nullptr; }
int64_t atde_launch_count(const atde_encoder* e) { return e ? e->launches : 0; }

int atde_reset(atde_encoder* e)
{
    if (!e) return fail(ATDE_ERR_INVALID, "null handle");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    e->have_state = false;
    e->n_state_streams = 0;
    e->streams_started = false;
    atde::at3p::pipeline_reset(e->at3p);
    return 0;
}

int64_t atde_output_frames(const atde_encoder* e, int64_t n_frames)
{
    if (!e || n_frames < 0) return ATDE_ERR_INVALID;
    if (e->lookahead && !(e->have_state && e->streams_started)) return n_frames > 0 ? n_frames - 1 : 0;
    return n_frames;
}

int atde_sync(atde_encoder* e)
{
    if (!e) return fail(ATDE_ERR_INVALID, "null handle");
    for (int i = 0; i < atde_encoder::kSlots; i++) CK(cudaStreamSynchronize(e->ws[i].stream));
    return 0;
}

int atde_encode_batch_device(atde_encoder* e, const float* d_pcm, int32_t S, int64_t F,
                             uint8_t* d_out, int32_t* d_sizes)
{
    if (!e || !d_pcm || !d_out) return fail(ATDE_ERR_INVALID, "null argument");
    if (S <= 0 || F <= 0) return fail(ATDE_ERR_INVALID, "empty batch (S=%d, F=%lld)", S, (long long)F);
    if (F > (1 << 21)) return fail(ATDE_ERR_INVALID, "too many frames per stream in one batch");
    CK(cudaSetDevice(e->cfg.device));
    int rc = ensure_state(e, S);
    if (rc) return rc;
    e->last_S = S; e->last_F = F;
    if (e->cfg.codec == ATDE_CODEC_ATRAC3PLUS) {
        const bool started = e->streams_started;
        e->last_out = F - (started ? 0 : 1);
        const char* why = "";
        atde::at3p::Profiler prof;
        prof.ctx = e; prof.begin = at3p_prof_begin; prof.end = at3p_prof_end;
        rc = atde::at3p::pipeline_run(e->at3p, d_pcm, 0, S, S, F, started, d_out, e->ws[0].stream, 0, &e->launches, &why, &prof);
        if (rc) return fail(rc == -3 ? ATDE_ERR_NOMEM : ATDE_ERR_CUDA, "ATRAC3plus pipeline: %s (%s)", why, cudaGetErrorString(cudaGetLastError()));
        atde::at3p::pipeline_commit(e->at3p);
        e->streams_started = true;
        return 0;
    }
    if (e->cfg.codec == ATDE_CODEC_ATRAC3) {
        const bool started = e->streams_started;
        e->last_out = F - (started ? 0 : 1);
        rc = run_at3(e, e->ws[0], d_pcm, 0, S, F, started, d_out);
        if (!rc) e->streams_started = true;
        return rc;
    }
    e->last_out = F;
    rc = run_at1(e, e->ws[0], d_pcm, 0, S, F, d_out, d_sizes);
    if (!rc) e->streams_started = true;
    return rc;
}

static int encode_batch_host(atde_encoder* e, const float* pcm, const int16_t* pcm16, int32_t S, int64_t F, uint8_t* out, int32_t* sizes);

int atde_encode_batch(atde_encoder* e, const float* pcm, int32_t S, int64_t F, uint8_t* out, int32_t* sizes)
{
    if (!e || !pcm || !out) return fail(ATDE_ERR_INVALID, "null argument");
    return encode_batch_host(e, pcm, nullptr, S, F, out, sizes);
}

int atde_encode_batch_i16(atde_encoder* e, const int16_t* pcm, int32_t S, int64_t F, uint8_t* out, int32_t* sizes)
{
    if (!e || !pcm || !out) return fail(ATDE_ERR_INVALID, "null argument");
    return encode_batch_host(e, nullptr, pcm, S, F, out, sizes);
}

static int encode_batch_host(atde_encoder* e, const float* pcm, const int16_t* pcm16, int32_t S, int64_t F, uint8_t* out, int32_t* sizes)
{
    if (S <= 0 || F <= 0) return fail(ATDE_ERR_INVALID, "empty batch (S=%d, F=%lld)", S, (long long)F);
    if (F > (1 << 21)) return fail(ATDE_ERR_INVALID, "too many frames per stream in one batch");
    CK(cudaSetDevice(e->cfg.device));
    int rc = ensure_state(e, S);
    if (rc) return rc;
    e->last_S = S; e->last_F = F;
    // atde_encode_batch_device() only enqueues on slot 0's stream: what it still has in flight must be done before
    // slot 1's stream touches the carried stream state
    CK(cudaStreamSynchronize(e->ws[0].stream));
    const int C = e->cfg.channels;
    const bool at3p = e->cfg.codec == ATDE_CODEC_ATRAC3PLUS;
    const bool at3 = e->cfg.codec == ATDE_CODEC_ATRAC3 || at3p;      // one-frame look-ahead, fixed-size units
    const bool started = e->streams_started;
    const long long n_out = F - ((at3 && !started) ? 1 : 0);
    e->last_out = n_out;
    const size_t pcm_per_stream = (size_t)F * e->frame_samples * C;            // floats
    const size_t out_per_stream = (size_t)n_out * e->units_per_frame * e->unit_bytes;
    const size_t units_per_stream = (size_t)n_out * e->units_per_frame;
    // chunk by streams so H2D of chunk k+1 overlaps compute of chunk k (two pipeline slots).  The compute-bound
    // codecs want large chunks (kernel tails; measured on the 10^6-frame ATRAC3 batch: 96 MiB 351 ms, 192 MiB
    // 326 ms, 384 MiB 294 ms, 1 GiB 286 ms per step), ATRAC1 is bound by the H2D copy and wants a short last
    // chunk.  The first chunk is a quarter of the size so the device starts early; the rest is split evenly.
    // (Not for ATRAC3plus: its tone search has a long per-launch tail that the next chunk's kernels fill, and
    // 191 + 191 + ... + 69 streams measured 350 ms per 250,880 frames against 373 ms with the short first chunk
    // and 494 ms with seven equal ones.)
    // ATRAC3 after the round-2 kernel work (10^6-frame batch, 168 ms of kernels against 149 ms of float copy-in): float PCM
    // 192 / 256 / 320 / 384 / 512 / 768 / 1024 / 1536 MiB per chunk -> 184.9 / 183.1 / 185.7 / 185.0 / 186.8 / 188.4 / 191.9 /
    // 195.5 ms per step; int16 PCM (half the bytes: compute-bound) 182.7 / 181.5 / 181.5 / 182.3 / 180.9 / 179.5 / 180.7 / 181.2
    // (profiles/tools/sweep_chunk.sh): 256 MiB chunks for float input, 768 MiB-equivalent ones for int16
    size_t target_floats = (size_t)(at3 ? (pcm16 || at3p ? 192 : 64) : 48) << 20;   // floats of PCM per chunk
    if (at3p) {
        // every tone-search launch ends in a ~14 ms tail: aim at about five chunks per batch, 768 MiB .. 3 GiB each
        // (10^6-frame batch: 21 chunks of 768 MiB 1479 ms per step against 1188 ms device-resident)
        const size_t fifth = (size_t)S * pcm_per_stream / 5;
        target_floats = fifth < ((size_t)192 << 20) ? ((size_t)192 << 20) : fifth > ((size_t)768 << 20) ? ((size_t)768 << 20) : fifth;
    }
    if (const char* env = getenv("ATDE_CHUNK_MIB")) {                           // tuning knob (MiB of PCM per chunk)
        const long v = atol(env);
        if (v > 0) target_floats = (size_t)v << 18;
    }
    int chunk = (int)(target_floats / pcm_per_stream);
    if (const char* env = getenv("ATDE_CHUNK_STREAMS")) {                       // test knob: streams per chunk
        const long v = atol(env);
        if (v > 0) chunk = (int)v;
    }
    if (chunk < 1) chunk = 1;
    if (chunk > S) chunk = S;
    // The plan: a quarter-size first chunk (the device starts early), equal middle chunks, and a half- and a
    // quarter-size chunk at the end — when the copy-in is as slow as the kernels (float PCM), the call ends one chunk's
    // compute after the last byte has arrived, so the last chunk should be short.  ATDE_CHUNK_TAPER=0 turns the taper off.
    std::vector<int> plan;
    {
        const int first = !at3p && chunk >= 4 && chunk < S ? chunk / 4 : chunk;
        plan.push_back(first);
        int rest = S - first;
        const char* tp = getenv("ATDE_CHUNK_TAPER");
        const bool taper = !at3p && chunk >= 8 && rest > 2 * chunk && !(tp && atoi(tp) == 0);
        const int t1 = taper ? chunk / 2 : 0, t2 = taper ? chunk / 4 : 0;
        rest -= t1 + t2;
        if (rest > 0) {
            const int pieces = at3p ? (rest + chunk - 1) / chunk : std::max(1, (rest + chunk - 1) / chunk);
            const int each = at3p ? chunk : (rest + pieces - 1) / pieces;
            for (int left = rest; left > 0; left -= each) plan.push_back(std::min(each, left));
        }
        if (t1) plan.push_back(t1);
        if (t2) plan.push_back(t2);
    }
    // A failure inside the loop must not return while earlier chunks' copies still use the caller's buffers, and it
    // leaves the carried stream state half advanced: drain both pipeline slots and invalidate the state
    // (the next batch needs atde_reset()).
    auto bail = [&](int code) {
        if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
        for (int i = 0; i < atde_encoder::kSlots; i++) cudaStreamSynchronize(e->ws[i].stream);
        e->have_state = false;
        e->n_state_streams = 0;
        e->streams_started = false;
        atde::at3p::pipeline_reset(e->at3p);
        return code;
    };
#define CKB(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return bail(fail(ATDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    int slot = 0;
    const int n_slots = at3p ? 2 : atde_encoder::kSlots;      // (the ATRAC3plus pipeline keeps two work areas)
    // The slots' streams run freely: the kernels of two chunks may share the SMs.  (ATDE_HOST_ORDERED=1 makes chunk k+1's
    // kernels wait for chunk k's while its copy-in still overlaps them — measured slower, 236 against 209 ms per 10^6
    // ATRAC3 frames: the small serial scan kernels then leave the device idle.  Kept for experiments.)
    const bool ordered = !at3p && getenv("ATDE_HOST_ORDERED") != nullptr;
    cudaEvent_t prev_done = nullptr;
    const bool ring = at3 && !at3p && e->copy_stream && getenv("ATDE_NO_PCM_RING") == nullptr;
    const float* d_chunk_pcm = nullptr;
    cudaEvent_t pcm_consumed = nullptr;
    int k = 0;
    if (ring) {
        // size the staging buffers before anything is in flight (a cudaMalloc in mid-pipeline synchronises the device)
        const size_t cnt = (size_t)*std::max_element(plan.begin(), plan.end()) * pcm_per_stream;
        for (int i = 0; i < std::min((int)plan.size(), (int)atde_encoder::kPcmRing); i++)
            if ((rc = pcm16 ? e->pcm16_ring[i].ensure(cnt + 8) : e->pcm_ring[i].ensure(cnt))) return bail(rc);
    }
    for (int s0 = 0; k < (int)plan.size(); s0 += plan[k], slot = (slot + 1) % n_slots, k++) {
        const int n = plan[k];
        Workspace& w = e->ws[slot];
        if ((rc = w.out.ensure((size_t)n * F * e->units_per_frame * e->unit_bytes + 1))) return bail(rc);   // (F frames: no growth on the continuation call)
        if (sizes && (rc = w.sizes.ensure((size_t)n * units_per_stream))) return bail(rc);
        if (ring) {
            const int pb = k % atde_encoder::kPcmRing;
            const size_t cnt = (size_t)n * pcm_per_stream;
            if (k >= atde_encoder::kPcmRing) CKB(cudaStreamWaitEvent(e->copy_stream, e->pcm_free[pb], 0));
            if (pcm16) {
                if ((rc = e->pcm16_ring[pb].ensure(cnt + 8))) return bail(rc);
                if ((rc = w.pcm.ensure(cnt))) return bail(rc);
                CKB(cudaMemcpyAsync(e->pcm16_ring[pb].p, pcm16 + (size_t)s0 * pcm_per_stream, cnt * sizeof(short), cudaMemcpyHostToDevice, e->copy_stream));
            } else {
                if ((rc = e->pcm_ring[pb].ensure(cnt))) return bail(rc);
                CKB(cudaMemcpyAsync(e->pcm_ring[pb].p, pcm + (size_t)s0 * pcm_per_stream, cnt * sizeof(float), cudaMemcpyHostToDevice, e->copy_stream));
            }
            CKB(cudaEventRecord(e->pcm_ready[pb], e->copy_stream));
            CKB(cudaStreamWaitEvent(w.stream, e->pcm_ready[pb], 0));
            if (pcm16) {
                const unsigned blocks = (unsigned)std::min<size_t>((cnt / 8 + 255) / 256 + 1, (size_t)148 * 16);
                ATDE_LAUNCH(pcm_i16_to_f32_kernel, blocks, 256, 0, w.stream, (const short*)e->pcm16_ring[pb].p, w.pcm.p, (long long)cnt);
                e->launches += 1;
                CKB(cudaEventRecord(e->pcm_free[pb], w.stream));           // the conversion is the staging buffer's one reader
                d_chunk_pcm = w.pcm.p;
                pcm_consumed = nullptr;
            } else {
                d_chunk_pcm = e->pcm_ring[pb].p;
                pcm_consumed = e->pcm_free[pb];
            }
        } else if ((rc = w.pcm.ensure((size_t)n * pcm_per_stream))) {
            return bail(rc);
        } else if (pcm16) {
            const size_t cnt = (size_t)n * pcm_per_stream;
            if ((rc = w.pcm16.ensure(cnt + 8))) return bail(rc);
            CKB(cudaMemcpyAsync(w.pcm16.p, pcm16 + (size_t)s0 * pcm_per_stream, cnt * sizeof(short), cudaMemcpyHostToDevice, w.stream));
            const unsigned blocks = (unsigned)std::min<size_t>((cnt / 8 + 255) / 256 + 1, (size_t)148 * 16);
            ATDE_LAUNCH(pcm_i16_to_f32_kernel, blocks, 256, 0, w.stream, (const short*)w.pcm16.p, w.pcm.p, (long long)cnt);
            e->launches += 1;
        } else {
            CKB(cudaMemcpyAsync(w.pcm.p, pcm + (size_t)s0 * pcm_per_stream, (size_t)n * pcm_per_stream * sizeof(float),
                               cudaMemcpyHostToDevice, w.stream));
        }
        if (ordered && prev_done) CKB(cudaStreamWaitEvent(w.stream, prev_done, 0));
        if (at3p) {
            const char* why = "";
            atde::at3p::Profiler prof;
            prof.ctx = e; prof.begin = at3p_prof_begin; prof.end = at3p_prof_end;
            rc = atde::at3p::pipeline_run(e->at3p, w.pcm.p, s0, n, S, F, started, w.out.p, w.stream, slot, &e->launches, &why, &prof);
            if (rc) return bail(fail(rc == -3 ? ATDE_ERR_NOMEM : ATDE_ERR_CUDA, "ATRAC3plus pipeline: %s (%s)", why, cudaGetErrorString(cudaGetLastError())));
        } else if (at3) rc = run_at3(e, w, ring ? d_chunk_pcm : w.pcm.p, s0, n, F, started, w.out.p, ring ? pcm_consumed : nullptr);
        else rc = run_at1(e, w, w.pcm.p, s0, n, F, w.out.p, sizes ? w.sizes.p : nullptr);
        if (rc) return bail(rc);
        if (ordered) {
            CKB(cudaEventRecord(e->chunk_done[slot], w.stream));
            prev_done = e->chunk_done[slot];
        }
        if (out_per_stream)
            CKB(cudaMemcpyAsync(out + (size_t)s0 * out_per_stream, w.out.p, (size_t)n * out_per_stream,
                               cudaMemcpyDeviceToHost, w.stream));
        if (sizes && !at3)
            CKB(cudaMemcpyAsync(sizes + (size_t)s0 * units_per_stream, w.sizes.p, (size_t)n * units_per_stream * sizeof(int),
                               cudaMemcpyDeviceToHost, w.stream));
    }
    if (e->copy_stream) CK(cudaStreamSynchronize(e->copy_stream));
    for (int i = 0; i < atde_encoder::kSlots; i++) CK(cudaStreamSynchronize(e->ws[i].stream));
#undef CKB
    if (sizes && at3)                                  // every WriteFrame payload is exactly FrameSz bytes
        for (size_t i = 0; i < (size_t)S * units_per_stream; i++) sizes[i] = e->unit_bytes;
    if (at3p) atde::at3p::pipeline_commit(e->at3p);
    e->streams_started = true;
    return 0;
}

int atde_set_profiling(atde_encoder* e, int32_t on)
{
    if (!e) return fail(ATDE_ERR_INVALID, "null handle");
    e->profiling = on != 0;
    e->ev_used = 0;
    return 0;
}

int atde_set_gain_trace(atde_encoder* e, int32_t on)
{
    if (!e) return fail(ATDE_ERR_INVALID, "null handle");
    if (e->cfg.codec != ATDE_CODEC_ATRAC3) return fail(ATDE_ERR_UNSUPPORTED, "the gain-control trace exists for ATRAC3 only");
    e->gain_trace = on != 0;
    if (on) e->taps_enabled = true;
    return 0;
}

int atde_kernel_times(atde_encoder* e, double* ms_sum, int64_t* count, int32_t n_kinds)
{
    if (!e || !ms_sum || !count) return fail(ATDE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    for (int k = 0; k < n_kinds; k++) { ms_sum[k] = 0; count[k] = 0; }
    for (size_t i = 0; i < e->ev_used; i++) {
        float ms = 0;
        const auto& p = e->ev_pool[i];
        if (p.kind < n_kinds && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { ms_sum[p.kind] += ms; count[p.kind]++; }
    }
    e->ev_used = 0;
    return 0;
}

int64_t atde_debug_tap(atde_encoder* e, int32_t what, void* dst, size_t capacity)
{
    if (!e || !dst) return fail(ATDE_ERR_INVALID, "null argument");
    if (what == 0) { e->taps_enabled = true; return 0; }      // arm taps for following batches
    const bool at3 = e->cfg.codec == ATDE_CODEC_ATRAC3;
    const size_t units = (size_t)e->last_S * e->last_out * e->cfg.channels;
    const Workspace& w = e->ws[0];
    const void* src = nullptr;
    size_t bytes = 0;
    if (at3) {
        const size_t L = (size_t)e->last_out + 1;
        switch (what) {
            case ATDE_TAP_SPECS: src = w.specs.p; bytes = units * 1024 * sizeof(float); break;      // scaled values
            case ATDE_TAP_CHLOUD: src = w.chloud.p; bytes = units * sizeof(float); break;
            case ATDE_TAP_LOUDNESS: src = w.loud.p; bytes = (size_t)e->last_S * e->last_out * sizeof(float); break;
            case ATDE_TAP_SFI: src = w.sfi.p; bytes = units * 32; break;
            case ATDE_TAP_WORDLEN: src = w.tap_wl.p; bytes = units * 32; break;
            case ATDE_TAP_BANDS: src = w.bands.p; bytes = (size_t)e->last_S * e->cfg.channels * 4 * (128 + 256 * L) * sizeof(float); break;
            case ATDE_TAP_CURVES: src = w.curves.p; bytes = (size_t)e->last_S * e->cfg.channels * 4 * e->last_out * sizeof(atde::at3::Curve); break;
            case ATDE_TAP_GSCALE: src = w.gscale.p; bytes = units * 16 * sizeof(float); break;
            case ATDE_TAP_ENERGY: src = w.energy.p; bytes = units * 32 * sizeof(float); break;
            case ATDE_TAP_TONAL: src = w.tonal.p; bytes = units * sizeof(atde::at3::TonalList); break;
            case ATDE_TAP_GAIN: src = w.gain.p; bytes = (size_t)e->last_S * e->cfg.channels * 3 * e->last_out * 96 * sizeof(float); break;
            case ATDE_TAP_TRACE_GAIN: src = w.trace_gain.p; bytes = (size_t)e->last_S * e->cfg.channels * 4 * e->last_out * 96 * sizeof(float); break;
            case ATDE_TAP_TRACE_STAT: src = w.trace_stat.p; bytes = (size_t)e->last_S * e->cfg.channels * 4 * e->last_out * 4 * sizeof(float); break;
            default: return fail(ATDE_ERR_INVALID, "unknown tap %d", what);
        }
    } else
    switch (what) {
        case ATDE_TAP_SPECS: src = w.specs.p; bytes = units * e->frame_samples * sizeof(float); break;
        case ATDE_TAP_MASKS: src = w.masks.p; bytes = units; break;
        case ATDE_TAP_CHLOUD: src = w.chloud.p; bytes = units * sizeof(float); break;
        case ATDE_TAP_LOUDNESS: src = w.loud.p; bytes = (size_t)e->last_S * e->last_F * sizeof(float); break;
        case ATDE_TAP_SFI: src = w.tap_sfi.p; bytes = units * 52; break;
        case ATDE_TAP_WORDLEN: src = w.tap_wl.p; bytes = units * 52; break;
        default: return fail(ATDE_ERR_INVALID, "unknown tap %d", what);
    }
    if (!src) return fail(ATDE_ERR_INVALID, "tap %d not captured (arm with what=0 before the batch)", what);
    if (bytes > capacity) return fail(ATDE_ERR_INVALID, "tap needs %zu bytes, capacity %zu", bytes, capacity);
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return (int64_t)bytes;
}

// ---- ATRAC1 decoder ----
struct atde_decoder {
    int channels = 0, device = 0;
    atde::at1::DecTables* d_tab = nullptr;
    DevBuf<unsigned char> units, hist, started;
    DevBuf<float> pcm;
    DevBuf<int> status;
    int n_state_streams = 0;
    cudaStream_t stream = nullptr;
};

int atde_decoder_create(int32_t channels, int32_t device, atde_decoder** out)
{
    using namespace atde;
    if (!out) return fail(ATDE_ERR_INVALID, "null argument");
    *out = nullptr;
    if (channels < 1 || channels > 2) return fail(ATDE_ERR_INVALID, "channels must be 1 or 2");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(ATDE_ERR_CUDA, "no CUDA device: libatde_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    atde_decoder* d = new (std::nothrow) atde_decoder();
    at1::DecTables* h = new (std::nothrow) at1::DecTables();
    if (!d || !h) { delete d; delete h; return fail(ATDE_ERR_NOMEM, "host alloc"); }
    d->channels = channels; d->device = device;
    memset(h, 0, sizeof(*h));
    qmf_window(h->qmf_window);
    for (uint32_t i = 0; i < 32; i++) h->sine_window[i] = sin((i + 0.5) * (M_PI / (2.0 * 32.0)));      // atrac1.h:128-132
    for (uint32_t i = 0; i < 64; i++) h->scale_table[i] = pow(2.0, (double)(i / 3.0 - 21.0));         // atrac1.h:122-127
    // TMIDCT<N>(N * 2) -> TMDCTBase(N, N) -> CalcSinCos(N, N) (atrac1denc.h:52-54, mdct.h:111-114)
    const std::vector<float> s512 = mdct_sincos(512, 512.0f), s256 = mdct_sincos(256, 256.0f), s64 = mdct_sincos(64, 64.0f);
    memcpy(h->isincos512, s512.data(), sizeof(h->isincos512));
    memcpy(h->isincos256, s256.data(), sizeof(h->isincos256));
    memcpy(h->isincos64, s64.data(), sizeof(h->isincos64));
    const auto tw128 = kiss_twiddles(128, false), tw64 = kiss_twiddles(64, false), tw16 = kiss_twiddles(16, false);
    memcpy(h->tw128, tw128.data(), sizeof(h->tw128));
    memcpy(h->tw64, tw64.data(), sizeof(h->tw64));
    memcpy(h->tw16, tw16.data(), sizeof(h->tw16));
    const auto p128 = kiss_perm(128), p64 = kiss_perm(64), p16 = kiss_perm(16);
    for (int i = 0; i < 128; i++) h->perm128[i] = (unsigned char)p128[i];
    for (int i = 0; i < 64; i++) h->perm64[i] = (unsigned char)p64[i];
    for (int i = 0; i < 16; i++) h->perm16[i] = (unsigned char)p16[i];
    cudaError_t ce = cudaMalloc(&d->d_tab, sizeof(at1::DecTables));
    if (ce == cudaSuccess) ce = cudaMemcpy(d->d_tab, h, sizeof(at1::DecTables), cudaMemcpyHostToDevice);
    delete h;
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { atde_decoder_destroy(d); return fail(ATDE_ERR_CUDA, "decoder setup failed: %s", cudaGetErrorString(ce)); }
    *out = d;
    return 0;
}

void atde_decoder_destroy(atde_decoder* d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    d->units.release(); d->hist.release(); d->started.release(); d->pcm.release(); d->status.release();
    if (d->d_tab) cudaFree(d->d_tab);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

int atde_decoder_reset(atde_decoder* d)
{
    if (!d) return fail(ATDE_ERR_INVALID, "null handle");
    d->n_state_streams = 0;
    return 0;
}

int atde_decode_batch(atde_decoder* d, const uint8_t* units, int32_t S, int64_t F, float* pcm)
{
    using namespace atde::at1;
    if (!d || !units || !pcm) return fail(ATDE_ERR_INVALID, "null argument");
    if (S <= 0 || F <= 0 || F > (1 << 21)) return fail(ATDE_ERR_INVALID, "bad batch (S=%d, F=%lld)", S, (long long)F);
    CK(cudaSetDevice(d->device));
    const int C = d->channels;
    if (d->n_state_streams && d->n_state_streams != S)
        return fail(ATDE_ERR_INVALID, "batch has %d streams but the handle carries state for %d; call atde_decoder_reset() first", S, d->n_state_streams);
    int rc;
    const size_t n_units = (size_t)S * F * C;
    if ((rc = d->units.ensure(n_units * kUnitBytes))) return rc;
    if ((rc = d->pcm.ensure(n_units * 512))) return rc;
    if ((rc = d->status.ensure(1))) return rc;
    if (!d->n_state_streams) {
        if ((rc = d->hist.ensure((size_t)S * C * kUnitBytes))) return rc;
        if ((rc = d->started.ensure(S))) return rc;
        CK(cudaMemsetAsync(d->started.p, 0, S, d->stream));
    }
    CK(cudaMemsetAsync(d->status.p, 0, sizeof(int), d->stream));
    CK(cudaMemcpyAsync(d->units.p, units, n_units * kUnitBytes, cudaMemcpyHostToDevice, d->stream));
    DecodeParams p;
    p.units = d->units.p; p.hist = d->hist.p; p.started = d->started.p;
    p.hist_out = d->hist.p; p.started_out = d->started.p;
    p.pcm = d->pcm.p; p.status = d->status.p; p.tab = d->d_tab;
    p.S = S; p.C = C; p.F = (int)F;
    launch_decode(p, d->stream);
    CK(cudaGetLastError());
    int status = 0;
    CK(cudaMemcpyAsync(pcm, d->pcm.p, n_units * 512 * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaMemcpyAsync(&status, d->status.p, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    d->n_state_streams = S;
    if (status & 1) {
        d->n_state_streams = 0;
        return fail(ATDE_ERR_UNSUPPORTED, "a sound unit uses a block-size code the reference encoder never writes (partial short-block modes)");
    }
    return 0;
}

// ---- encoder groups: one member per device, streams sharded contiguously, members run concurrently ----
struct atde_group {
    std::vector<atde_encoder*> members;
};

static void group_range(int S, int n, int r, int* lo, int* hi)
{
    const int base = S / n, extra = S % n;
    *lo = r * base + (r < extra ? r : extra);
    *hi = *lo + base + (r < extra ? 1 : 0);
}

int atde_create_group(const atde_settings* s, const int32_t* devices, int32_t n_devices, atde_group** out)
{
    if (!s || !devices || !out || n_devices <= 0) return fail(ATDE_ERR_INVALID, "bad group arguments");
    *out = nullptr;
    atde_group* g = new (std::nothrow) atde_group();
    if (!g) return fail(ATDE_ERR_NOMEM, "host alloc");
    for (int r = 0; r < n_devices; r++) {
        atde_settings m = *s;
        m.device = devices[r];
        atde_encoder* e = nullptr;
        const int rc = atde_create(&m, &e);
        if (rc) {
            const std::string why = g_err;
            atde_destroy_group(g);
            return fail(rc, "group member %d (device %d): %s", r, devices[r], why.c_str());
        }
        g->members.push_back(e);
    }
    *out = g;
    return 0;
}

void atde_destroy_group(atde_group* g)
{
    if (!g) return;
    for (atde_encoder* e : g->members) atde_destroy(e);
    delete g;
}

int atde_group_size(const atde_group* g) { return g ? (int)g->members.size() : ATDE_ERR_INVALID; }

int64_t atde_group_output_frames(const atde_group* g, int64_t n_frames)
{
    if (!g || g->members.empty()) return ATDE_ERR_INVALID;
    return atde_output_frames(g->members[0], n_frames);
}

int atde_group_reset(atde_group* g)
{
    if (!g) return fail(ATDE_ERR_INVALID, "null group");
    for (atde_encoder* e : g->members) {
        const int rc = atde_reset(e);
        if (rc) return rc;
    }
    return 0;
}

static int group_encode(atde_group* g, const float* pcm, const int16_t* pcm16, int32_t S, int64_t F, uint8_t* out, int32_t* sizes)
{
    if (!g || (!pcm && !pcm16) || !out) return fail(ATDE_ERR_INVALID, "null argument");
    const int n = (int)g->members.size();
    if (S < n) return fail(ATDE_ERR_INVALID, "a group of %d members needs at least %d streams per batch (got %d)", n, n, S);
    if (F <= 0) return fail(ATDE_ERR_INVALID, "empty batch");
    const atde_encoder* e0 = g->members[0];
    const int64_t fo = atde_output_frames(e0, F);
    const size_t pcm_per_stream = (size_t)F * e0->frame_samples * e0->cfg.channels;
    const size_t units_per_stream = (size_t)fo * e0->units_per_frame;
    const size_t out_per_stream = units_per_stream * e0->unit_bytes;
    std::vector<int> rcs(n, 0);
    std::vector<std::string> errs(n);
    std::vector<std::thread> workers;
    workers.reserve(n);
    for (int r = 0; r < n; r++) {
#ifdef ATDE_CPU_EMU
        // (the CPU-emulation shim of the tests runs one kernel at a time: members take turns there)
        if (!workers.empty()) { workers.back().join(); workers.pop_back(); }
#endif
        workers.emplace_back([&, r]() {
            int lo, hi;
            group_range(S, n, r, &lo, &hi);
            rcs[r] = encode_batch_host(g->members[r], pcm ? pcm + (size_t)lo * pcm_per_stream : nullptr,
                                       pcm16 ? pcm16 + (size_t)lo * pcm_per_stream : nullptr, hi - lo, F,
                                       out + (size_t)lo * out_per_stream, sizes ? sizes + (size_t)lo * units_per_stream : nullptr);
            if (rcs[r]) errs[r] = g_err;                     // g_err is per thread: hand the text to the caller's thread
        });
    }
    for (auto& w : workers) w.join();
    for (int r = 0; r < n; r++)
        if (rcs[r]) return fail(rcs[r], "group member %d: %s", r, errs[r].c_str());
    return 0;
}

int atde_group_encode_batch(atde_group* g, const float* pcm, int32_t S, int64_t F, uint8_t* out, int32_t* sizes)
{
    return group_encode(g, pcm, nullptr, S, F, out, sizes);
}

int atde_group_encode_batch_i16(atde_group* g, const int16_t* pcm, int32_t S, int64_t F, uint8_t* out, int32_t* sizes)
{
    return group_encode(g, nullptr, pcm, S, F, out, sizes);
}

} // extern "C"

namespace {
__global__ void math_kernel(int fn, const float* x, float* y, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    y[i] = fn == 0 ? atde::g_log10f(v) : (fn == 1 ? atde::g_log2f(v) : atde::g_logf(v));
}
} // namespace

extern "C" int atde_debug_math(int32_t device, int32_t fn, const float* x, float* y, int64_t n)
{
    if (!x || !y || n <= 0) return fail(ATDE_ERR_INVALID, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(ATDE_ERR_CUDA, "no CUDA device: libatde_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    float *dx = nullptr, *dy = nullptr;
    CK(cudaMalloc(&dx, n * sizeof(float)));
    CK(cudaMalloc(&dy, n * sizeof(float)));
    CK(cudaMemcpy(dx, x, n * sizeof(float), cudaMemcpyHostToDevice));
    ATDE_LAUNCH(math_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t) nullptr, (int)fn, (const float*)dx, dy, (long long)n);
    CK(cudaGetLastError());
    CK(cudaMemcpy(y, dy, n * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy);
    return 0;
}
