// at3p_kernels.cuh — device-side interface of the ATRAC3plus encode path (at3p_kernels.cu).
//
// Reference chain: at3plus_pqf_do_analyse (src/atrac/atrac3plus_pqf/atrac3plus_pqf.c:130-147)
// -> IGhaProcessor::DoAnalize (src/atrac/at3p/at3p_gha.cpp:692) -> TAt3pMDCT::Do (src/atrac/at3p/at3p_mdct.cpp:52-96)
// -> TScaler<NAt3p::TScaleTable>::ScaleFrame (src/atrac/atrac_scale.cpp:141-188)
// -> TAt3PBitStream::WriteFrame (src/atrac/at3p/at3p_bitstream.cpp:703-726), driven by
// TAt3PEnc::TImpl::EncodeFrame (src/atrac/at3p/at3p.cpp:88-194).
//
// Kernels: the PQF analysis filterbank, the tone filter, the 16-band MDCT-256 and the frame packer
// (scale, quantise, code-table choice, tonal block, bit writer) in at3p_kernels.cu; the GHA tone
// search in at3p_gha.cu; at3p_pipeline.cu strings them together behind atde_create(codec 4).
#pragma once
#include "atde_cuda.h"

namespace atde {
namespace at3p {

constexpr int kFrame = 2048;          // samples per channel-frame (src/atrac3p.h:61)
constexpr int kSubbands = 16;
constexpr int kSbSamples = 128;
constexpr int kPqfProto = 384;        // prototype length (atrac3plus_pqf.c:42)
constexpr int kPqfOverlap = kPqfProto - kSubbands;   // 368 history samples (atrac3plus_pqf.c:45)
constexpr int kQuantUnits = 32;
constexpr int kFrameBytes = 2048;     // at3p.cpp:41: BitStream(out, 2048)
constexpr int kMaxWaves = 48;         // at3p_gha.cpp:611

// Device tables, built on the host by build_tables() with the reference's expressions
struct DevTables {
    alignas(16) float sincos256[128]; // TMDCT<256>(1)   (mdct.cpp:25-36)
    alignas(16) float sine_win[128];  // SineWin128      (at3p_mdct.cpp:36-40)
    alignas(16) cpx tw64[64];         // forward kissfft twiddles of the 64-point FFT
    alignas(16) float scale_table[64];// NAt3p::TScaleTable::ScaleTable (at3p_tables.cpp:42-70)
    float inv_mant[8];                // 1 / atrac3p_mant_tab[wl] (at3p_tables.cpp:28-38)
    unsigned char spec_tab[56][4];    // group_size, num_coeffs, bits, is_signed (ff/atrac3plus_data.h:1427)
    unsigned vlc_off[57];
    unsigned spec_pack[56];           // the same per table in one word: group | coeffs << 4 | bits << 8 | signed << 12 | vlc_off << 16
    unsigned vlc[7812];               // code | len << 16 of THuffTables::VlcSpecs[0..55]
    unsigned wl_vlc[4][8];            // THuffTables::WordLens
    unsigned tone_bands_vlc[16];      // THuffTables::NumToneBands
    unsigned char qu_to_subband[32];
    unsigned char sb_to_powgrps[16];
    // tone synthesis (ff/atrac3plusdsp.c:49-66)
    alignas(16) float sine_table[2048];
    float hann_window[256];
    float amp_sf_tab[64];
};

// Flattened TAt3PGhaData (src/atrac/at3p/at3p_gha.h:29-66): what the tonal-block writer consumes.
struct ToneBlock {
    int present;                      // a non-null block with NumToneBands > 0
    int num_tone_bands;
    int second_is_leader;
    int tone_sharing[16];
    int n_sb[2];
    int sb[2][16][4];                 // WaveIndex, WaveNums, Envelope.first, Envelope.second
    int n_params[2];
    int params[2][64][4];             // FreqIndex, AmpSf, AmpIndex, PhaseIndex
};

const DevTables* device_tables();     // builds + uploads on first use; nullptr on failure

// where a unit's data lives when the kernels run inside the streaming pipeline (at3p_pipeline.cu):
// unit = (stream s, output q), q < fo
struct FilterLayout {
    int fo;                           // outputs per stream
    int tone_stride;                  // ToneBlock records per stream; the three tone pointers are pre-offset
    int in_frames, in_off;            // bands  [S][C][in_frames][2048], unit frame = in_off + q
    int out_frames, out_off;          // resid  [S][C][out_frames][2048], unit frame = out_off + q
};

// pcm [S][F*2048][C] interleaved -> bands [S][C][F][16][128]; every stream starts with a zero history
// pcm_tail: [S][368][C] samples preceding the batch (nullptr = zeros); bands [S][C][L][2048], frame f lands at joff + f
void launch_pqf(const float* pcm, const float* pcm_tail, float* bands, int S, int C, int F, int L, int joff, cudaStream_t st);
// Tone filter + MDCT input scaling (TGhaProcessorBase::ApplyFilter, at3p_gha.cpp:581-687;
// ff_atrac3p_generate_tones, ff/atrac3plusdsp.c:130-204; at3p.cpp:147-153): for unit u, frame bands
// [U][C][16][128] minus the tones of tb_now[u] / tb_next[u] (envelopes also need tb_old[u]) -> resid, same layout
void launch_tone_filter(const DevTables* T, const float* bands, const ToneBlock* tb_old, const ToneBlock* tb_now,
                        const ToneBlock* tb_next, float* resid, int units, int C, const FilterLayout& lay, cudaStream_t st);
// resid [S][C][F][16][128] (already scaled to the MDCT's input range) -> specs [S][F][C][2048];
// every stream starts with a zero overlap history
void launch_mdct(const DevTables* T, const float* resid, float* specs, int S, int C, int F, int lead, cudaStream_t st);
// specs [U][C][2048], tones [U] -> frames [U][2048]
// tones of unit (s, q) = tones[s * tone_stride + q]
void launch_pack(const DevTables* T, const float* specs, const ToneBlock* tones, unsigned char* frames,
                 int units, int C, int fo, int tone_stride, cudaStream_t st);

// tone search (at3p_gha.cu)
bool gha_tables_ready();
size_t gha_scratch_bytes(int blocks);
size_t gha_frame_out_bytes();
size_t gha_history_bytes();
int gha_blocks_for(long long n_analyses);
// analyses (s, f), f < nA, of band frames j0 + f (look-ahead j0 + f + 1) of bands [S][C][L][2048] -> frame_out [S][nA]
void launch_gha_search(const float* bands, int S, int C, int nA, int L, int j0, void* scratch, void* frame_out, int blocks, cudaStream_t st);
// per-stream FillResultBuf + history carry: frame_out [S][nA] -> tones[s * stride + off + f]
void launch_gha_result(const void* frame_out, int S, int C, int nA, void* hist_state, ToneBlock* tones, int stride, int off, cudaStream_t st);

// streaming pipeline (at3p_pipeline.cu): the body of TAt3PEnc::TImpl::EncodeFrame for S streams x N calls
struct StreamState;                   // per-handle carried state
StreamState* pipeline_create(int C, int gha_flags = 7);   // gha_flags: TAt3PEnc::TSettings::UseGha (bits 0..2)
void pipeline_destroy(StreamState*);
void pipeline_reset(StreamState*);
// Encodes N new frames per stream (d_pcm [S][N*2048][C]) continuing streams s0 .. s0+S-1 of `total_streams`;
// writes n_out = N (N - 1 on the first batch) frames per stream to d_out [S][n_out][2048].  Returns 0 or a
// negative status with *err set.
// optional per-kernel event timing supplied by the API layer; kinds: 0 PQF + MDCT, 2 scale/quantise/pack,
// 3 tone search, 4 tone filter
struct Profiler {
    void* ctx = nullptr;
    int (*begin)(void* ctx, cudaStream_t st, int kind) = nullptr;
    void (*end)(void* ctx, cudaStream_t st, int idx) = nullptr;
};
int pipeline_run(StreamState*, const float* d_pcm, int s0, int S, int total_streams, long long N, bool started,
                 unsigned char* d_out, cudaStream_t st, int slot, long long* launches, const char** err,
                 const Profiler* prof = nullptr);
// call once per batch after every chunk was enqueued
void pipeline_commit(StreamState*);

} // namespace at3p
} // namespace atde
