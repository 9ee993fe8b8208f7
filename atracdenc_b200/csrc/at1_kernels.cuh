// at1_kernels.cuh — device-side interface of the ATRAC1 encode path (see at1_kernels.cu).
#pragma once
#include "atde_cuda.h"

namespace atde {
namespace at1 {

constexpr int kFrame = 512;          // samples per channel-frame (atrac1.h:121)
constexpr int kMaxBfus = 52;         // atrac1.h:86
constexpr int kUnitBytes = 212;      // sound unit size (atrac1.h:105)
constexpr float kLoudFactor = 0.006f; // atrac1denc.h:101

// Tables in global memory (read through L1).  Filled by the host from host_tables.cpp.
struct DevTables {
    float sine_window[32];       // atrac1.h:128-132
    float scale_table[64];       // atrac1.h:122-127
    float loud_curve[512];       // CreateLoudnessCurve(512)
    float ath_long[kMaxBfus];    // CalcAt1ATH, atrac1_bitalloc.cpp:118-135
    float sincos512[256];        // TMDCT<512>(1)
    float sincos256[128];        // TMDCT<256>(0.5)
    float sincos64[32];          // TMDCT<64>(0.5)
    cpx tw128[128], tw64[64], tw16[16];          // forward kissfft twiddles
    unsigned char perm128[128], perm64[64], perm16[16];
    // MDCT input weights (TAtrac1MDCT::Mdct, atrac1denc.cpp:80-90) as tables: position t of the windowed stretch of a
    // long block (32-sample sine slope | 1.0 | mirrored slope) and of a 64-sample short block (slope | mirrored slope)
    float win_long128[160], win_long256[288], win_short[64];
    // Fold map of a long block / of a short block (mdct.h:56-76), per output slot in kissfft's gather order: the four positions
    // t = j - joff of the slot's inputs in the windowed stretch, 10 bits each (a0 | a1 << 10 | b0 << 20 | b1 << 30;
    // 1023 = outside the stretch, the reference's buffer holds 0 there), and the slot's index i = n / 2 at bit 40
    alignas(8) unsigned long long fold128[64], fold256[128], fold64[16];   // long 128- / 256-sample band, short block
};

struct AnalysisParams {
    const float* pcm;            // [S][F*512][C] interleaved, normalised
    const float* hist;           // [S][512][C] previous frame of each stream, or nullptr (= zeros)
    const unsigned char* started;// [S] non-zero: stream already ran (frame 0 of this batch is not the stream's first); may be nullptr
    float* specs;                // [S][F][C][512]
    unsigned char* masks;        // [S][F][C]  window mask bits low=1 mid=2 hi=4
    float* chloud;               // [S][F][C]  per-channel loudness term
    const DevTables* tab;
    int S, C;
    int F;
    int window_auto;             // EWM_AUTO
    int window_mask;             // used when !window_auto
    float one;                   // 1.0f, opaque to the compiler (atde_cuda.h: add2)
};

struct LoudnessParams {
    const unsigned char* masks;  // [S][F][C]
    const float* chloud;         // [S][F][C]
    const float* loud_in;        // [S] carried loudness or nullptr (= LoudFactor)
    float* loud;                 // [S][F]  Loudness after frame f's update
    int S, C, F;
};

struct PackParams {
    const float* specs;          // [S][F][C][512]
    const unsigned char* masks;  // [S][F][C]
    const float* loud;           // [S][F]
    unsigned char* out;          // [S][F][C][212]
    int* sizes;                  // [S][F][C] WriteFrame payload length as the reference grows it, or nullptr
    unsigned char* tap_sfi;      // [S][F][C][52] or nullptr
    unsigned char* tap_wl;       // [S][F][C][52] or nullptr (after boost; 0xff beyond nbfu)
    const DevTables* tab;
    int S, C, F;
    int bfu_idx_const;
};

struct CarryParams {
    const float* pcm;            // [S][F*512][C]
    const float* loud;           // [S][F]
    float* hist;                 // [S][512][C]
    float* loud_state;           // [S]
    unsigned char* started;      // [S]
    int S, C, F;
};

// ---- decoder (at1_decode.cu) ----
struct DecTables {
    float sine_window[32];       // atrac1.h:128-132
    float scale_table[64];       // atrac1.h:122-127
    float qmf_window[48];        // QmfWindow, qmf.cpp:36-45
    float isincos512[256];       // TMIDCT<512>(512 * 2): CalcSinCos(512, 512)
    float isincos256[128];       // TMIDCT<256>(256 * 2)
    float isincos64[32];         // TMIDCT<64>(64 * 2)
    cpx tw128[128], tw64[64], tw16[16];          // forward kissfft twiddles (TMDCTBase always plans a forward FFT)
    unsigned char perm128[128], perm64[64], perm16[16];
};

struct DecodeParams {
    const unsigned char* units;      // [S][F][C][212] sound units
    const unsigned char* hist;       // [S][C][212] last sound unit of the previous batch
    const unsigned char* started;    // [S] non-zero: `hist` is valid
    unsigned char* hist_out;         // staging for the carry (may equal hist: written after the decode kernel)
    unsigned char* started_out;
    float* pcm;                      // [S][F*512][C] interleaved, clipped to [-1, 1]
    int* status;                     // bit 0: a frame used a block-size code this decoder refuses
    const DecTables* tab;
    int S, C, F;
};
void launch_decode(const DecodeParams& p, cudaStream_t st);

void upload_qmf_window(const float w[48]);
void launch_analysis(const AnalysisParams& p, cudaStream_t st);
void launch_loudness(const LoudnessParams& p, cudaStream_t st);
void launch_pack(const PackParams& p, cudaStream_t st);
void launch_carry(const CarryParams& p, cudaStream_t st);

constexpr int kTile = 4;             // frames per analysis block

} // namespace at1
} // namespace atde
