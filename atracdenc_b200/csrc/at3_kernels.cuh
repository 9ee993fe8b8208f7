// at3_kernels.cuh — device-side interface of the ATRAC3 encode path (see at3_kernels.cu).
#pragma once
#include "atde_cuda.h"

namespace atde {
namespace at3 {

constexpr int kFrame = 1024;          // samples per channel-frame (atrac3.h:64)
constexpr int kBfus = 32;             // atrac3.h:63
constexpr int kBands = 4;
constexpr float kLoudFactor = 0.006f; // atrac3denc.h:114
constexpr int kMaxTonal = 32;         // the reference asserts < 32 tonal groups per channel

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
    float planck[512];
    float hpf_h[2];                   // H at LowCutBin, LowCutBin+1
    int low_cut_bin;
    cpx tw256[256];                   // forward, for kiss_fftr(512)
    cpx super512[128];                // forward super twiddles
    unsigned short perm256[256];
    cpx tw2048[2048];                 // inverse, for kiss_fftri(4096)
    cpx super4096[1024];              // inverse super twiddles
    unsigned short iperm2048[2048];   // input index -> gather slot
};

// A gain curve of one (channel, band, frame): n points, each level (4 bit) / location (5 bit)
struct Curve {
    unsigned short n;
    unsigned short pt[7];             // level << 8 | location
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
    int L;                            // extended frames per stream in this batch (carried + new)
    int n_out;                        // = L - 1 output frames
    int started;                      // 1: frame 0 of the extended sequence is the carried one
    int js;                           // joint stereo (LP4 with 2 channels)
    int frame_sz;                     // container frame size in bytes
    int no_gain, no_tonal;
    int bfu_idx_const;
};

struct Buffers {
    const float* pcm;                 // [S][N][1024][C] new frames of this batch (N = L - started)
    float* pcm_hist;                  // [S][2][1024][C]: frame before the carried one, carried frame
    float* bands;                     // [S][C][4][128 + L*256]
    Curve* curves;                    // [S][C][4][L]
    float* prevhalf;                  // carry [S][C][4][256]
    float* next_scale;                // carry [S][C][4]   PrevOverlapGainScale
    float* specs;                     // [S][n_out][C][1024] (scaled in place by the scale kernel)
    float* gscale;                    // [S][n_out][C][4][3] PrevHalf, CurHalf, Frame
    float* chloud;                    // [S][n_out][C]
    float* loud_state;                // carry [S]
    float* loud;                      // [S][n_out]
    unsigned char* sfi;               // [S][n_out][C][32]
    float* energy;                    // [S][n_out][C][32]
    TonalList* tonal;                 // [S][n_out][C]
    unsigned char* out;               // [S][n_out][frame_sz]
    // gain control scratch
    float* gain;                      // [S][C][4][n_out][96]: gain[32], low[32], high[32]
    float* gstat;                     // [S][C][4][n_out][4]: hfr, curHpf, target, reserved
    float* gprev;                     // [S][C][4][n_out][4]: prevHpf, savedLastLevel, savedLastTarget, reserved
    float* ctx;                       // carry [S][C][4][4]: LastLevel, LastHpfEnergy, LastTarget
    unsigned char* tap_prec;          // [S][n_out][C][32] or nullptr
    const DevTables* tab;
};

void upload_qmf_window(const float w[48]);
void launch_qmf(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_analysis(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_scan(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_gain_curve(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_mdct(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_loudness(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_scale_tonal(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_alloc_pack(const Geometry& g, const Buffers& b, cudaStream_t st);
void launch_carry(const Geometry& g, const Buffers& b, cudaStream_t st);

} // namespace at3
} // namespace atde
