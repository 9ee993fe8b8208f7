/*
 * atde_b200.h — C ABI of libatde_b200.so, the B200-native drop-in for atracdenc's per-frame
 * encode hot path (QMF analysis -> MDCT -> transient detection / gain control -> bit allocation
 * -> quantisation -> frame bitstream).
 *
 * What it replaces in the reference (dcherednik/atracdenc):
 *   - the frame-processor lambdas returned by IProcessor::GetLambda()  (src/pcmengin.h:195-199)
 *       TAtrac1Encoder::GetLambda   src/atrac1denc.cpp:180-255
 *       TAtrac3Encoder::GetLambda   src/atrac3denc.cpp:679-867
 *       TAt3PEnc::GetLambda         src/atrac/at3p/at3p.cpp:202-206 (TImpl::EncodeFrame :88-194)
 *   - and, on the output side, produces exactly the byte vectors those lambdas hand to
 *       ICompressedOutput::WriteFrame(std::vector<char>)             (src/compressed_io.h:56-59)
 *     in the same order (ATRAC1: one 212-byte sound unit per channel per frame, channel 0 first,
 *     src/atrac/at1/atrac1_bitalloc.cpp:406; ATRAC3: one FrameSz-byte unit per frame,
 *     src/atrac/at3/atrac3_bitstream.cpp:845; ATRAC3plus: one 2048-byte unit per frame,
 *     src/atrac/at3p/at3p_bitstream.cpp:724-725).
 *
 * Unit of work: a batch of S independent streams x F consecutive frames.  A "stream" is what
 * the reference calls one encoder instance.  Streams continue across calls (the handle keeps
 * the cross-frame state of SURVEY.md §3.4 on the device) until atde_reset().
 *
 * Plain C: pointers and sizes only; no C++/torch types.  Every call returns 0 on success or a
 * negative atde_status; atde_last_error() describes the last failure on the calling thread.
 * There is NO CPU fallback: without a CUDA device every entry point that computes fails.
 */
#ifndef ATDE_B200_H
#define ATDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ATDE_OK = 0,
    ATDE_ERR_INVALID = -1,      /* bad argument / unsupported setting */
    ATDE_ERR_CUDA = -2,         /* CUDA runtime error (message in atde_last_error) */
    ATDE_ERR_NOMEM = -3,
    ATDE_ERR_UNSUPPORTED = -4   /* codec or option not built yet */
} atde_status;

typedef enum {
    ATDE_CODEC_ATRAC1 = 1,      /* -e atrac1        src/main.cpp:635-655 */
    ATDE_CODEC_ATRAC3 = 3,      /* -e atrac3 / atrac3_lp4  src/main.cpp:657-678 */
    ATDE_CODEC_ATRAC3PLUS = 4   /* -e atrac3plus    src/main.cpp:679-686; TAt3PEnc::TSettings incl. the ghadbg masks, no GHA_WIDEBAND */
} atde_codec;

/* Mirrors the reference's settings objects field by field. */
typedef struct {
    int32_t codec;              /* atde_codec */
    int32_t channels;           /* 1 or 2 (ICompressedOutput::GetChannelNum / TAtrac3EncoderSettings::SourceChannels) */
    /* NAtrac1::TAtrac1EncodeSettings (src/atrac/at1/atrac1.h:33-54) */
    uint32_t bfu_idx_const;     /* --bfuidxconst, 0 = automatic (also TAtrac3EncoderSettings::BfuIdxConst) */
    int32_t window_mode;        /* 0 = EWM_NOTRANSIENT, 1 = EWM_AUTO */
    uint32_t window_mask;       /* used when window_mode == 0 : bit0 low, bit1 mid, bit2 hi short */
    /* NAtrac3::TAtrac3EncoderSettings (src/atrac/at3/atrac3.h:260-277) */
    uint32_t bitrate;           /* bits/s as main.cpp passes it (kbit*1024); 0 = LP2 default */
    int32_t no_gain_control;    /* --nogaincontrol */
    int32_t no_tonal;           /* --notonal */
    int32_t device;             /* CUDA device ordinal */
    /* NAtracDEnc::TAt3PEnc::TSettings::UseGha (src/atrac3p.h:29-47; `--advanced ghadbg=N`): bit 0 GHA_PASS_INPUT,
     * bit 1 GHA_WRITE_TONAL, bit 2 GHA_WRITE_RESIUDAL; atde_default_settings sets GHA_ENABLED (7).  Bit 3
     * (GHA_WIDEBAND, an opt-in experiment of the reference) is refused. */
    uint32_t gha_flags;
    int32_t reserved[6];
} atde_settings;

typedef struct atde_encoder atde_encoder;

/* Fills *s with the reference's defaults for `codec` (main.cpp: TAtrac1EncodeSettings(0, EWM_AUTO, 0);
 * TAtrac3EncoderSettings(0, false, false, channels, 0)). */
void atde_default_settings(atde_settings* s, int32_t codec, int32_t channels);

/* Construct / destroy an encoder (== constructing TAtrac1Encoder / TAtrac3Encoder / TAt3PEnc). */
int atde_create(const atde_settings* s, atde_encoder** out);
void atde_destroy(atde_encoder* e);

/* Geometry of the codec the handle was built for. */
int atde_frame_samples(const atde_encoder* e);     /* sample-frames per lambda call: 512 / 1024 / 2048 */
int atde_units_per_frame(const atde_encoder* e);   /* WriteFrame calls per lambda call: ATRAC1 = channels, ATRAC3 = 1 */
int atde_unit_bytes(const atde_encoder* e);        /* bytes stored per unit in `out` (container frame size) */
int atde_lookahead_frames(const atde_encoder* e);  /* lambda calls that return LOOK_AHEAD before output starts (0 / 1) */
/* Output frames per stream the NEXT batch of n_frames input frames will produce: n_frames, except
 * that the first batch of an ATRAC3 / ATRAC3plus stream yields n_frames - 1 (the reference's first lambda call
 * only fills the look-ahead buffer and returns LOOK_AHEAD, src/atrac3denc.cpp:715-718). */
int64_t atde_output_frames(const atde_encoder* e, int64_t n_frames);

/*
 * Encode S streams x F frames from HOST memory (the reference-facing path: copies in, kernels,
 * copies out; synchronous).
 *   pcm   [S][F*frame_samples][channels] interleaved normalised float32 — the memory layout the
 *         reference lambda receives (data[i*Channels + ch], src/atrac1denc.cpp:206-209), one
 *         stream after another.
 *   out   [S][Fo][units_per_frame][unit_bytes], Fo = atde_output_frames(e, F); each unit is the WriteFrame payload zero-padded or
 *         truncated to the container frame size exactly as TAeaOutput/TRaw do (src/aea.cpp:182,
 *         src/raw.cpp:41-43).
 *   sizes optional [S][F][units_per_frame]: true payload length (std::vector<char>::size()) of
 *         each WriteFrame call; NULL to skip.
 * Streams continue from the previous call on this handle (same S required) unless atde_reset()
 * was called; the first batch after create/reset starts every stream from the reference's
 * initial encoder state.
 */
int atde_encode_batch(atde_encoder* e, const float* pcm, int32_t n_streams, int64_t n_frames,
                      uint8_t* out, int32_t* sizes);

/* Same, from interleaved 16-bit PCM [S][F*frame_samples][channels] (what a WAV file holds): half the host->device
 * bytes.  The samples are converted on the device the way the reference's reader converts them before the PCM
 * engine sees them — libsndfile's sf_readf_float on a PCM_16 file, value * (1 / 0x8000)
 * (src/pcm_io_sndfile.cpp:111-113, src/wav.cpp:46-61) — so the result equals atde_encode_batch() on those
 * floats.  SURVEY.md 8(f) rank 2. */
int atde_encode_batch_i16(atde_encoder* e, const int16_t* pcm, int32_t n_streams, int64_t n_frames,
                          uint8_t* out, int32_t* sizes);

/* Same, with pcm/out/sizes already resident in DEVICE memory of the handle's GPU; enqueued on the
 * handle's stream, returns without synchronising (use atde_sync). */
int atde_encode_batch_device(atde_encoder* e, const float* d_pcm, int32_t n_streams, int64_t n_frames,
                             uint8_t* d_out, int32_t* d_sizes);
int atde_sync(atde_encoder* e);

/* Forget all stream state: the next batch starts new streams. */
int atde_reset(atde_encoder* e);

/* cudaStream_t the handle launches on (as void*), for event timing by the caller. */
void* atde_cuda_stream(atde_encoder* e);
/* Number of kernel launches issued by this handle so far. */
int64_t atde_launch_count(const atde_encoder* e);

/* Per-kernel device timing.  With profiling on, every kernel launch of the handle is bracketed by
 * CUDA events on its stream; atde_kernel_times() synchronises, returns the summed milliseconds
 * and launch counts per kernel kind since the last query and clears them.
 * Kinds: 0 = QMF+MDCT analysis, 1 = loudness scan, 2 = (scale/)allocate/quantise/pack,
 *        ATRAC3 only: 3 = gain-control envelope analysis, 4 = gain curve scan + build, 5 = tonal + scale. */
int atde_set_profiling(atde_encoder* e, int32_t on);
int atde_kernel_times(atde_encoder* e, double* ms_sum, int64_t* count, int32_t n_kinds);

/* Test taps: copy an intermediate of the LAST batch to host memory.  Returns bytes copied or <0. */
typedef enum {
    ATDE_TAP_SPECS = 1,     /* float  [S][F][C][frame_samples] MDCT spectra */
    ATDE_TAP_MASKS = 2,     /* uint8  [S][F][C] ATRAC1 window masks */
    ATDE_TAP_CHLOUD = 3,    /* float  [S][F][C] per-channel loudness term */
    ATDE_TAP_LOUDNESS = 4,  /* float  [S][F] tracked loudness */
    ATDE_TAP_SFI = 5,       /* uint8  [S][F][C][52|32] scale factor indices */
    ATDE_TAP_WORDLEN = 6,   /* uint8  [S][F][C][52|32] word lengths / precisions (0xff beyond the coded BFUs) */
    /* ATRAC3 only (F = output frames of the batch) */
    ATDE_TAP_BANDS = 7,     /* float  [S][C][4][128 + 256*(F+1)] QMF bands incl. look-ahead frame (M/S when joint stereo) */
    ATDE_TAP_CURVES = 8,    /* 16-byte records [S][C][4][F]: n, level[7], loc[7], pad */
    ATDE_TAP_GSCALE = 9,    /* float  [S][F][C][4][4] PrevHalf, CurHalf, Frame, NextOverlapScale */
    ATDE_TAP_ENERGY = 10,   /* float  [S][F][C][32] BFU energies */
    ATDE_TAP_TONAL = 11,    /* tonal block lists [S][F][C] (at3_kernels.cuh: TonalList) */
    ATDE_TAP_GAIN = 12,     /* float  [S][C][3][F][96] sub-frame envelope: gain, low, high */
    /* ATRAC3 with the gain-control trace on (atde_set_gain_trace): the same analysis over all FOUR bands */
    ATDE_TAP_TRACE_GAIN = 13, /* float [S][C][4][F][96] gain, low, high */
    ATDE_TAP_TRACE_STAT = 14  /* float [S][C][4][F][4] high-frequency ratio, mean envelope, plateau target, next_level */
} atde_tap;
int64_t atde_debug_tap(atde_encoder* e, int32_t what, void* host_dst, size_t capacity);

/* ATRAC3: the data behind the reference's `--yaml-log <file>` gain-control trace (src/yaml_log.h:19-57; written by
 * src/atrac3denc.cpp:305-579,743-800 and src/transient_detector.cpp:298-446 through TAtrac3EncoderSettings::YamlLog,
 * src/atrac/at3/atrac3.h:263-276).  With the trace on, every following batch also runs the envelope analysis over band 3
 * (which the reference analyses and logs although it never carries a curve), keeps the exact high-frequency ratio and
 * the look-ahead level `next_level`, and keeps the taps above valid; the text itself is produced on the host from those
 * taps (atracdenc_b200/host/atde_gain_trace.{h,cpp}).  A debugging aid: the encode path is unchanged when it is off. */
int atde_set_gain_trace(atde_encoder* e, int32_t on);

/* Device-side math self-test hooks (tests only): evaluates the glibc replicas on n inputs. */
int atde_debug_math(int32_t device, int32_t fn /*0 log10f 1 log2f 2 logf*/, const float* x, float* y, int64_t n);

/*
 * Multi-GPU in ONE process: a group of encoders, one per device, that shards a batch BY STREAM (the reference's unit
 * of independence is the encoder instance == stream; frames of one stream never leave a device).
 * Stream s of an S-stream batch goes to member r with lo(r) <= s < lo(r+1), lo(r) = r*(S/n) + min(r, S%n) —
 * contiguous ranges whose sizes differ by at most one.  atde_group_encode_batch() runs the members concurrently (one
 * host thread per device); every device pulls its own shard of `pcm` from host memory over its own PCIe link and
 * writes its own slice of `out`, so there is no inter-GPU traffic and no collective on the data path.  Layouts,
 * stream-state and look-ahead semantics are those of atde_encode_batch() (the same S must be used from call to call
 * until atde_group_reset()).  `devices` may name a device more than once (several members on one GPU).
 * The reference has no counterpart (it is single-threaded, one encoder per process); this is what a caller that used
 * to fork one `atracdenc` per file and core calls instead.
 */
typedef struct atde_group atde_group;
int atde_create_group(const atde_settings* s, const int32_t* devices, int32_t n_devices, atde_group** out);
void atde_destroy_group(atde_group* g);
int atde_group_size(const atde_group* g);
int atde_group_encode_batch(atde_group* g, const float* pcm, int32_t n_streams, int64_t n_frames,
                            uint8_t* out, int32_t* sizes);
int atde_group_encode_batch_i16(atde_group* g, const int16_t* pcm, int32_t n_streams, int64_t n_frames,
                                uint8_t* out, int32_t* sizes);
int64_t atde_group_output_frames(const atde_group* g, int64_t n_frames);
int atde_group_reset(atde_group* g);

/*
 * ATRAC1 decoder (the step after the encode path): what TAtrac1Decoder's frame lambda writes
 * (src/atrac1denc.cpp:139-177 — dequantise, IMDCT, QMF synthesis, clip to [-1, 1], interleave), for S streams x F
 * frames at once, bit-exact.
 *   units [S][F][channels][212] sound units in ICompressedInput::ReadFrame order (channel 0 first, src/aea.cpp)
 *   pcm   [S][F*512][channels]  interleaved float, the layout the lambda fills (data[i*channels + ch])
 * Streams continue across calls (same S) until atde_decoder_reset().  A unit the reference would reject (negative
 * block-size code, mantissas running past the unit) decodes as silence, as there.  Block-size codes the reference
 * encoder never writes (2-of-4 / 2- or 4-of-8 short blocks), for which the reference decoder reads stale buffer
 * contents, are refused: ATDE_ERR_UNSUPPORTED.
 */
typedef struct atde_decoder atde_decoder;
int atde_decoder_create(int32_t channels, int32_t device, atde_decoder** out);
void atde_decoder_destroy(atde_decoder* d);
int atde_decode_batch(atde_decoder* d, const uint8_t* units, int32_t n_streams, int64_t n_frames, float* pcm);
int atde_decoder_reset(atde_decoder* d);

const char* atde_last_error(void);
const char* atde_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ATDE_B200_H */
