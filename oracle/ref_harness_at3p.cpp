/*
 * oracle/ref_harness_at3p.cpp — TEST INFRASTRUCTURE (see ref_harness.cpp header).
 * ATRAC3plus stage taps.  This translation unit takes the place of src/atrac/at3p/at3p.cpp in the
 * oracle build: it #includes that UNMODIFIED source where it lies (TAt3PEnc::TImpl is private to
 * it) and is compiled with -fno-access-control, so the harness can
 *   - wrap the encoder's IGhaProcessor in a recording proxy (PQF-domain input, tone result and the
 *     residual left in the work buffers after TGhaProcessorBase::ApplyFilter), and
 *   - read TImpl's per-channel Specs after every lambda call.
 * Nothing of the reference is changed; the proxy forwards every call to the real processor.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <string>

#include "atrac/at3p/at3p.cpp"          // the reference TU itself (src/atrac/at3p/at3p.cpp)
#include "atrac/at3p/at3p_mdct.h"
#include "atrac/atrac3plus_pqf/atrac3plus_pqf.h"

using namespace NAtracDEnc;

namespace {
class TNullOut3p : public ICompressedOutput {
    size_t Ch;
    std::vector<uint8_t>* Bytes;
public:
    TNullOut3p(size_t ch, std::vector<uint8_t>* b) : Ch(ch), Bytes(b) {}
    void WriteFrame(std::vector<char> data) override { Bytes->insert(Bytes->end(), data.begin(), data.end()); }
    std::string GetName() const override { return "null"; }
    size_t GetChannelNum() const override { return Ch; }
};
}

extern "C" {

/* Flattened TAt3PGhaData (src/atrac/at3p/at3p_gha.h:29-66). */
struct at3p_gha_rec {
    int32_t present;                 /* DoAnalize returned non-null */
    int32_t num_tone_bands;
    int32_t second_is_leader;
    int32_t tone_sharing[16];
    int32_t n_sb[2];                 /* Waves[ch].WaveSbInfos.size() */
    int32_t sb[2][16][4];            /* WaveIndex, WaveNums, Envelope.first, Envelope.second */
    int32_t n_params[2];
    int32_t params[2][64][4];        /* FreqIndex, AmpSf, AmpIndex, PhaseIndex */
};

int ref_at3p_gha_rec_size(void) { return (int)sizeof(at3p_gha_rec); }

} // extern "C"

namespace {

void Flatten(const TAt3PGhaData* d, int channels, at3p_gha_rec* r)
{
    memset(r, 0, sizeof(*r));
    if (!d) return;
    r->present = 1;
    r->num_tone_bands = d->NumToneBands;
    r->second_is_leader = d->SecondIsLeader;
    for (int i = 0; i < 16; i++) r->tone_sharing[i] = d->ToneSharing[i];
    for (int ch = 0; ch < channels; ch++) {
        const auto& w = d->Waves[ch];
        r->n_sb[ch] = (int32_t)w.WaveSbInfos.size();
        for (size_t i = 0; i < w.WaveSbInfos.size() && i < 16; i++) {
            r->sb[ch][i][0] = (int32_t)w.WaveSbInfos[i].WaveIndex;
            r->sb[ch][i][1] = (int32_t)w.WaveSbInfos[i].WaveNums;
            r->sb[ch][i][2] = (int32_t)w.WaveSbInfos[i].Envelope.first;
            r->sb[ch][i][3] = (int32_t)w.WaveSbInfos[i].Envelope.second;
        }
        r->n_params[ch] = (int32_t)w.WaveParams.size();
        for (size_t i = 0; i < w.WaveParams.size() && i < 64; i++) {
            r->params[ch][i][0] = (int32_t)w.WaveParams[i].FreqIndex;
            r->params[ch][i][1] = (int32_t)w.WaveParams[i].AmpSf;
            r->params[ch][i][2] = (int32_t)w.WaveParams[i].AmpIndex;
            r->params[ch][i][3] = (int32_t)w.WaveParams[i].PhaseIndex;
        }
    }
}

/* Forwards to the real processor and records what went in and what came out. */
class TTapGha : public IGhaProcessor {
public:
    TTapGha(std::unique_ptr<IGhaProcessor> real, int channels) : Real(std::move(real)), Channels(channels) {}
    const TAt3PGhaData* DoAnalize(TBufPtr b1, TBufPtr b2, float* w1, float* w2,
                                  const float* raw1Cur, const float* raw2Cur) override
    {
        for (int ch = 0; ch < Channels; ch++) {
            const TBufPtr& b = ch ? b2 : b1;
            memcpy(Cur[ch], b[0], sizeof(Cur[ch]));
            memcpy(Next[ch], b[1], sizeof(Next[ch]));
            memcpy(WorkIn[ch], ch ? w2 : w1, sizeof(WorkIn[ch]));
        }
        const TAt3PGhaData* res = Real->DoAnalize(b1, b2, w1, w2, raw1Cur, raw2Cur);
        for (int ch = 0; ch < Channels; ch++)
            memcpy(WorkOut[ch], ch ? w2 : w1, sizeof(WorkOut[ch]));
        Flatten(res, Channels, &Rec);
        return res;
    }
    std::unique_ptr<IGhaProcessor> Real;
    int Channels;
    float Cur[2][2048], Next[2][2048], WorkIn[2][2048], WorkOut[2][2048];
    at3p_gha_rec Rec;
};

} // namespace

extern "C" {

/*
 * Drives TAt3PEnc's lambda frame by frame (src/atrac/at3p/at3p.cpp:88-194).  Per OUTPUT frame o
 * (lambda call o+1; the first call returns LOOK_AHEAD), any pointer may be NULL:
 *   pqf_cur  [o][C][2048]  PQF-domain frame analysed in that call (CurBuf: 16 subbands x 128)
 *   pqf_next [o][C][2048]  look-ahead frame (NextBuf)
 *   work_in  [o][C][2048]  PrevBuf before DoAnalize (the frame that gets encoded, tones still in)
 *   work_out [o][C][2048]  PrevBuf after ApplyFilter (residual that goes to the MDCT)
 *   gha      [o]           result of DoAnalize in that call (it is WRITTEN one call later: `delay`)
 *   specs    [o][C][2048]  MDCT output of that call
 *   frames   [o][2048]     the frame bytes written in that call
 * Returns the number of output frames.
 */
long ref_at3p_stages(int channels, const float* pcm, long n_frames, int gha_flags,
                     float* pqf_cur, float* pqf_next, float* work_in, float* work_out,
                     at3p_gha_rec* gha, float* specs, unsigned char* frames)
{
    std::vector<uint8_t> bytes;
    TCompressedOutputPtr out(new TNullOut3p(channels, &bytes));
    TAt3PEnc::TSettings settings;
    if (gha_flags >= 0) settings.UseGha = (uint8_t)gha_flags;
    TAt3PEnc enc(std::move(out), channels, settings);
    auto* impl = enc.Impl.get();
    auto* tap = new TTapGha(std::move(impl->GhaProcessor), channels);
    impl->GhaProcessor.reset(tap);

    auto lambda = enc.GetLambda();
    TPCMEngine::ProcessMeta meta{(uint16_t)channels};
    std::vector<float> frame(2048 * channels);
    long produced = 0;
    for (long f = 0; f < n_frames; f++) {
        memcpy(frame.data(), pcm + (size_t)f * 2048 * channels, frame.size() * sizeof(float));
        const size_t before = bytes.size();
        auto res = lambda(frame.data(), meta);
        if (res != TPCMEngine::EProcessResult::PROCESSED)
            continue;
        const size_t o = (size_t)produced;
        for (int c = 0; c < channels; c++) {
            const size_t at = (o * channels + c) * 2048;
            if (pqf_cur) memcpy(pqf_cur + at, tap->Cur[c], 2048 * sizeof(float));
            if (pqf_next) memcpy(pqf_next + at, tap->Next[c], 2048 * sizeof(float));
            if (work_in) memcpy(work_in + at, tap->WorkIn[c], 2048 * sizeof(float));
            if (work_out) memcpy(work_out + at, tap->WorkOut[c], 2048 * sizeof(float));
            if (specs) memcpy(specs + at, impl->ChannelCtx[c].Specs.data(), 2048 * sizeof(float));
        }
        if (gha) gha[o] = tap->Rec;
        if (frames && bytes.size() - before == 2048) memcpy(frames + o * 2048, bytes.data() + before, 2048);
        produced++;
    }
    return produced;
}

/* TScaler<NAt3p::TScaleTable>::ScaleFrame + TAt3PBitStream::WriteFrame on caller-supplied spectra and
 * tone data (the tail of EncodeFrame, at3p.cpp:163-168): specs [U][C][2048], tones [U] -> frames [U][2048]. */
long ref_at3p_pack(int channels, const float* specs, const at3p_gha_rec* tones, long units, unsigned char* frames)
{
    std::vector<uint8_t> bytes;
    TNullOut3p out(channels, &bytes);
    TAt3PBitStream bs(&out, 2048);
    TScaler<NAt3p::TScaleTable> scaler;
    for (long u = 0; u < units; u++) {
        std::vector<TAt3PBitStream::TSingleChannelElement> sces(channels);
        for (int c = 0; c < channels; c++) {
            std::vector<float> sp(specs + ((size_t)u * channels + c) * 2048, specs + ((size_t)u * channels + c + 1) * 2048);
            sces[c].ScaledBlocks = scaler.ScaleFrame(sp, NAt3p::TScaleTable::TBlockSizeMod());
        }
        TAt3PGhaData d;
        const at3p_gha_rec& r = tones[u];
        const TAt3PGhaData* p = nullptr;
        if (r.present && r.num_tone_bands) {
            d.NumToneBands = (uint8_t)r.num_tone_bands;
            d.SecondIsLeader = r.second_is_leader != 0;
            for (int i = 0; i < 16; i++) d.ToneSharing[i] = r.tone_sharing[i] != 0;
            for (int ch = 0; ch < channels; ch++) {
                d.Waves[ch].WaveSbInfos.resize(r.n_sb[ch]);
                for (int i = 0; i < r.n_sb[ch]; i++) {
                    d.Waves[ch].WaveSbInfos[i].WaveIndex = (size_t)r.sb[ch][i][0];
                    d.Waves[ch].WaveSbInfos[i].WaveNums = (size_t)r.sb[ch][i][1];
                    d.Waves[ch].WaveSbInfos[i].Envelope = {(uint32_t)r.sb[ch][i][2], (uint32_t)r.sb[ch][i][3]};
                }
                d.Waves[ch].WaveParams.resize(r.n_params[ch]);
                for (int i = 0; i < r.n_params[ch]; i++)
                    d.Waves[ch].WaveParams[i] = TAt3PGhaData::TWaveParam{(uint32_t)r.params[ch][i][0], (uint32_t)r.params[ch][i][1],
                                                                         (uint32_t)r.params[ch][i][2], (uint32_t)r.params[ch][i][3]};
            }
            p = &d;
        }
        const size_t before = bytes.size();
        bs.WriteFrame(channels, p, sces);
        if (bytes.size() - before != 2048) return -1;
        memcpy(frames + (size_t)u * 2048, bytes.data() + before, 2048);
    }
    return units;
}

/* at3plus_pqf_do_analyse (src/atrac/atrac3plus_pqf/atrac3plus_pqf.c:130-147) over n_frames frames of
 * one channel, fresh context: out[f][2048] (16 subbands x 128). */
void ref_at3p_pqf(const float* pcm, long n_frames, float* out)
{
    at3plus_pqf_a_ctx_t ctx = at3plus_pqf_create_a_ctx();
    for (long f = 0; f < n_frames; f++)
        at3plus_pqf_do_analyse(ctx, pcm + f * 2048, out + f * 2048);
    at3plus_pqf_free_a_ctx(ctx);
}

/* TAt3pMDCT::Do (src/atrac/at3p/at3p_mdct.cpp:52-96) over n_frames frames of one channel, sine
 * windows, fresh history: bands[f][2048] -> specs[f][2048]. */
void ref_at3p_mdct(const float* bands, long n_frames, float* specs)
{
    TAt3pMDCT mdct;
    TAt3pMDCT::THistBuf hist = {{{0}}};
    for (long f = 0; f < n_frames; f++) {
        TAt3pMDCT::TPcmBandsData p;
        for (size_t b = 0; b < 16; b++) p[b] = bands + f * 2048 + b * 128;
        mdct.Do(specs + f * 2048, p, hist, TAt3pMDCTWin());
    }
}

} // extern "C"
