// at3_kernels.cuh — device-side interface of the ATRAC3 encode path (at3_analysis.cu, at3_pack.cu).
//
// Frame numbering inside one batch.  The reference encoder runs one frame behind its input
// (src/atrac3denc.cpp:697-718: the first lambda call only fills the look-ahead slot), so a batch is
// described by an EXTENDED frame sequence per stream:
//     ext frame 0      = the frame carried from the previous batch (started) or new frame 0 (fresh)
//     ext frame 1..L-1 = the remaining new frames
// Output frame f (0 <= f < n_out = L-1) encodes ext frame f and needs ext frames f-1 .. f+1.
// Band buffers are indexed u = 128 + 256*f + i for sample i of ext frame f; u < 128 is the tail of
// the frame before ext frame 0 (zeros for a fresh stream).
#pragma once
#include "atde_cuda.h"
#include "kissfft_dev.cuh"

namespace atde {
namespace at3 {

constexpr int kFrame = 1024;          // samples per channel-frame (atrac3.h:64)
constexpr int kBfus = 32;             // atrac3.h:63
constexpr int kBands = 4;
constexpr int kGainBands = 3;         // band 3 never carries a curve (atrac3denc.cpp:444-450)
constexpr float kLoudFactor = 0.006f; // atrac3denc.h:115
constexpr int kMaxTonal = 24;         // <= one run per BFU 8..28, runs may only merge
constexpr int kMaxUnitBytes = 1024;   // largest container frame (atrac3.h:211-220)

// Device tables (global memory), built on the host by atde_api.cu:build_at3_tables
struct DevTables {
    float scale_table[64];            // atrac3.h:179-183
    float encode_window[256];         // atrac3.h:184-186
    float gain_level[16];             // atrac3.h:192-194
    float gain_interp[31];            // atrac3.h:195-197
    float loud_curve[1024];           // CreateLoudnessCurve(1024)
    float ath[kBfus];                 // atrac3_bitstream.cpp:694-718
    float sincos512[256];             // TMDCT<512>(1)
    cpx tw128[128];                   // forward kissfft twiddles, MDCT-512's 128-point FFT
    unsigned char perm128[128];
    // gain control: TSpectralUpsampler (transient_spectral_upsampler.cpp)
    alignas(16) float planck[512];    // Planck-taper window, eps = 0.15 (read as float2)
    float hpf_h[2];                   // raised-cosine H at LowCutBin, LowCutBin+1
    int low_cut_bin;
    cpx tw256[256];                   // forward, kiss_fftr(512) -> complex FFT-256
    cpx super512[128];                // forward super twiddles of kiss_fftr(512)
    unsigned char perm256[256];
    cpx tw2048[2048];                 // inverse, kiss_fftri(4096) -> complex FFT-2048
    cpx super4096[1024];              // inverse super twiddles of kiss_fftri(4096)
    unsigned short iperm2048[2048];   // input index k -> gather slot of the digit-reversed order
    // the same twiddles regrouped per kernel pass so that consecutive lanes read consecutive elements
    cpx ftw[4][3][64];                // forward FFT-256, stage st (m = 4^st): tw256[(q+1) k (64/m)], k < m
    // (inverse passes 2 and 3 run on packed pairs; the twiddles are spread to (r, r | i, -i) on the fly, kissfft_dev.cuh)
    cpx gtw2[15][8];                  // inverse pass 2, lane group k: tw2048[64k(q+1)] | tw2048[16(k+8a)(q+1)]
    cpx gtw3a[3][128];                // inverse pass 3, m = 128: tw2048[4k(q+1)]
    cpx gtw3b[4][3][128];             // inverse pass 3, m = 512: tw2048[(k+128a)(q+1)]
};

// Gain curve of one (stream, channel, band, frame): n points, each level (4 bit) / location (5 bit)
struct Curve {
    unsigned char n;
    unsigned char level[7];
    unsigned char loc[7];
    unsigned char pad;
};

struct TonalBlock {
    unsigned short pos;               // first spectral line
    unsigned char bfu;                // BFU of the first line (TTonalVal::Bfu of ValPtr)
    unsigned char sfi;
    unsigned char len;
    unsigned char pad[3];
    float val[7];                     // scaled values
};

struct TonalList {
    int n;
    TonalBlock b[kMaxTonal];
};

struct Geometry {
    int S, C;
    int N;                            // new frames per stream in this batch
    int L;                            // extended frames per stream (N + started)
    int n_out;                        // = L - 1 output frames
    int BL;                           // band buffer length per (s,c,band) = 128 + 256*L
    int js;                           // joint stereo (container Js flag and C == 2)
    int js_mono;                      // container Js flag with ONE input channel: an empty second element is written
    int frame_sz;                     // container frame size in bytes
    int no_gain, no_tonal;
    int bfu_idx_const;
    float one;                        // 1.0f, opaque to the compiler (atde_cuda.h: add2)
};

struct Buffers {
    const float* pcm;                 // [S][N*1024][C] new frames of this batch
    unsigned char* started;           // [S] (all equal inside a batch; kept per stream for clarity)
    float* hist_tmp;                  // [S][2][1024][C] staging for the pcm_hist update
    // carried stream state
    float* pcm_hist;                  // [S][2][1024][C]: frame before the carried one, carried frame
    float* prevhalf;                  // [S][C][4][256] windowed + modulated half kept for ext frame 0
    float* next_scale;                // [S][C][4] PrevOverlapGainScale before ext frame 0
    float* ctx;                       // [S][C][3][4]: LastLevel, LastHpfEnergy, LastTarget, -
    float* loud_state;                // [S]
    // per-batch work buffers
    float* bands;                     // [S][C][4][BL]  (M/S already matrixed when js)
    float* gain;                      // [S][C][3][n_out][96]: gain[32], low[32], high[32]
    float* gstat;                     // [S][C][3][n_out][4]: hfr, curHpfEnergy, target, gain[31]
    float* gprev;                     // [S][C][3][n_out][4]: prevHpfEnergy, savedLastLevel, savedLastTarget, -
    Curve* curves;                    // [S][C][4][n_out]
    float* prevhalf_out;              // [S][C][4][256] staging: half left behind by the last output frame
    float* next_scale_out;            // [S][C][4]
    float* specs;                     // [S][n_out][C][1024] (scaled in place by the scale kernel)
    float* gscale;                    // [S][n_out][C][4][4]: PrevHalf, CurHalf, Frame, NextOverlapScale
    float* chloud;                    // [S][n_out][C]
    float* loud;                      // [S][n_out]
    unsigned char* sfi;               // [S][n_out][C][32]
    float* energy;                    // [S][n_out][C][32]
    TonalList* tonal;                 // [S][n_out][C]
    unsigned char* out;               // [S][n_out][frame_sz]
    unsigned char* tap_prec;          // [S][n_out][C][32] or nullptr (0xff beyond numBfu)
    // gain-control trace (`--yaml-log`), only allocated and written when atde_set_gain_trace() turned it on
    float* trace_gain;                // [S][C][4][n_out][96]: gain[32], low[32], high[32] — all FOUR bands
    float* trace_stat;                // [S][C][4][n_out][4]: hfr (sequential sums), curHpfEnergy, target, next_level
    const DevTables* tab;
};

void upload_qmf_window(const float w[48]);
void launch_qmf(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_analysis(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_trace(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_scan(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_curve(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_mdct(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_loudterm(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_loudness(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_scale_tonal(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_alloc_pack(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_carry(const Geometry& g, const Buffers& b, cudaStream_t st);

} // namespace at3
} // namespace atde
