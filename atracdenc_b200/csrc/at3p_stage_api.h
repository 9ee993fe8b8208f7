/* at3p_stage_api.h — stage-level entry points of the ATRAC3plus kernels (host buffers in/out).
 * NOT part of the drop-in boundary (include/atde_b200.h; the product path is atde_create(ATDE_CODEC_ATRAC3PLUS)).
 * tests/ uses these to check every kernel of the ATRAC3plus chain on its own against the reference's taps.  All return 0 on success, -2 on a CUDA error, -3 on allocation failure. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
/* at3plus_pqf_do_analyse (src/atrac/atrac3plus_pqf/atrac3plus_pqf.c:130-147) over F frames of S fresh
 * streams: pcm [S][F*2048][C] interleaved -> bands [S][C][F][16][128]. */
int atde_at3p_stage_pqf(const float* pcm, int S, int C, int F, float* bands);
/* TAt3pMDCT::Do, sine windows (src/atrac/at3p/at3p_mdct.cpp:52-96), fresh history:
 * resid [S][C][F][16][128] -> specs [S][F][C][2048]. */
int atde_at3p_stage_mdct(const float* resid, int S, int C, int F, float* specs);
/* TScaler::ScaleFrame + TAt3PBitStream::WriteFrame (src/atrac/at3p/at3p_bitstream.cpp:703-726):
 * specs [U][C][2048], tones [U] flattened TAt3PGhaData (atde_at3p_tone_block_size() bytes each, layout of
 * at3p_kernels.cuh:ToneBlock) -> frames [U][2048]. */
int atde_at3p_stage_pack(const float* specs, const void* tones, int units, int C, unsigned char* frames);
int atde_at3p_tone_block_size(void);
/* TGhaProcessorBase::ApplyFilter + ff_atrac3p_generate_tones + the MDCT input scaling (at3p_gha.cpp:581-687,
 * ff/atrac3plusdsp.c:130-204, at3p.cpp:147-153): bands [U][C][2048] minus the tones of three consecutive GHA
 * results per unit (two calls ago, previous call, this call) -> resid [U][C][2048]. */
int atde_at3p_stage_tone_filter(const float* bands, const void* tb_old, const void* tb_now, const void* tb_next,
                                int units, int C, float* resid);
/* IGhaProcessor::DoAnalize without the filter (src/atrac/at3p/at3p_gha.cpp:692-743): bands [S][C][F][2048] of fresh
 * streams (frame f looks ahead into frame f+1; zeros past the end) -> tones [S][F] (ToneBlock records). */
int atde_at3p_stage_gha(const float* bands, int S, int C, int F, void* tones);
/* device replicas of glibc sin / cos / atan / sincosf (glibc_trig.cuh); fn: 0 sin, 1 cos, 2 atan, 3 sinf, 4 cosf */
int atde_at3p_debug_trig(int fn, const double* x, double* y, long long n);
#ifdef __cplusplus
}
#endif
