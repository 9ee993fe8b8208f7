// atde_encoders.cpp — see atde_encoders.h.  Pure host C++ over the C ABI; no CUDA types here.
#include "atde_encoders.h"

#include <string>

namespace NAtracDEnc {

static void Check(int rc)
{
    // C-ABI status -> the exception type main.cpp already catches (src/main.cpp:709-720)
    if (rc < 0)
        throw std::runtime_error(std::string("atde_b200: ") + atde_last_error());
}

TBatchedEncoderBase::TBatchedEncoderBase(TCompressedOutputPtr&& out, const atde_settings& settings)
    : Out(std::move(out))
{
    Check(atde_create(&settings, &Enc));
    Channels = settings.channels;
    FrameSamples = atde_frame_samples(Enc);
    Units = atde_units_per_frame(Enc);
    UnitBytes = atde_unit_bytes(Enc);
    LookAhead = atde_lookahead_frames(Enc);
}

TBatchedEncoderBase::~TBatchedEncoderBase()
{
    try {
        Flush();
    } catch (...) {
        // destructors must not throw; a failed flush loses the staged tail exactly like an
        // exception escaping the reference's lambda would have
    }
    atde_destroy(Enc);
}

TPCMEngine::EProcessResult TBatchedEncoderBase::Push(const float* data)
{
    // the PCM pointer is only valid during the call (it points into TPCMEngine's buffer): copy
    const size_t n = (size_t)FrameSamples * Channels;
    if (Stage.size() < BatchFrames * n)
        Stage.resize(BatchFrames * n);
    memcpy(&Stage[Staged * n], data, n * sizeof(float));
    Staged++;
    const bool lookAhead = Calls < (uint64_t)LookAhead;
    Calls++;
    if (Staged == BatchFrames)
        Flush();
    return lookAhead ? TPCMEngine::EProcessResult::LOOK_AHEAD : TPCMEngine::EProcessResult::PROCESSED;
}

void TBatchedEncoderBase::Flush()
{
    if (!Staged)
        return;
    // a look-ahead codec's first batch yields one frame less than it consumes
    const size_t units = (size_t)atde_output_frames(Enc, (int64_t)Staged) * Units;
    Bytes.resize(units * UnitBytes + 8);
    Sizes.resize(units + 1);
    const size_t staged = Staged;
    Staged = 0;
    Check(atde_encode_batch(Enc, Stage.data(), 1, (int64_t)staged, Bytes.data(), Sizes.data()));
    for (size_t u = 0; u < units; u++) {
        // same bytes, same length, same order as the reference's WriteFrame calls; payload bytes
        // beyond the container frame size are the bit writer's zero growth slack
        const char* p = reinterpret_cast<const char*>(&Bytes[u * UnitBytes]);
        std::vector<char> frame((size_t)Sizes[u], 0);
        memcpy(frame.data(), p, (size_t)Sizes[u] < (size_t)UnitBytes ? (size_t)Sizes[u] : (size_t)UnitBytes);
        Out->WriteFrame(std::move(frame));
    }
}

static atde_settings MakeAt1Settings(size_t channels, const NAtrac1::TAtrac1EncodeSettings& s)
{
    atde_settings c;
    atde_default_settings(&c, ATDE_CODEC_ATRAC1, (int32_t)channels);
    c.bfu_idx_const = s.GetBfuIdxConst();
    c.window_mode = s.GetWindowMode() == NAtrac1::TAtrac1EncodeSettings::EWindowMode::EWM_AUTO ? 1 : 0;
    c.window_mask = s.GetWindowMask();
    return c;
}

TAtrac1Encoder::TAtrac1Encoder(TCompressedOutputPtr&& aea, NAtrac1::TAtrac1EncodeSettings&& settings)
    : TBatchedEncoderBase(std::move(aea), MakeAt1Settings(aea->GetChannelNum(), settings))
{
}

TPCMEngine::TProcessLambda TAtrac1Encoder::GetLambda()
{
    return [this](float* data, const TPCMEngine::ProcessMeta& /*meta*/) { return Push(data); };
}

#ifndef ATDE_USE_REFERENCE_HEADERS
namespace NAtrac3 {
static const TContainerParams kContainerParams[8] = {        // atrac3.h:211-220
    {66150, 192, true},   {93713, 272, true},   {104738, 304, false}, {132300, 384, false},
    {146081, 424, false}, {176400, 512, false}, {264600, 768, false}, {352800, 1024, false}};
const TContainerParams* GetContainerParamsForBitrate(uint32_t bitrate)
{
    if (bitrate == 0) bitrate = 132300;                      // LP2 by default (atrac3.cpp:45-51)
    const TContainerParams* p = kContainerParams;
    while (p != kContainerParams + 8 && p->Bitrate < bitrate) ++p;    // std::lower_bound
    return p;
}
} // namespace NAtrac3
#endif

static atde_settings MakeAt3Settings(const NAtrac3::TAtrac3EncoderSettings& s)
{
    atde_settings c;
    atde_default_settings(&c, ATDE_CODEC_ATRAC3, (int32_t)s.SourceChannels);
    c.bitrate = s.ConteinerParams->Bitrate;                  // an exact table bitrate maps back to the same container
    c.no_gain_control = s.NoGainControll;
    c.no_tonal = s.NoTonalComponents;
    c.bfu_idx_const = s.BfuIdxConst;
    return c;
}

TAtrac3Encoder::TAtrac3Encoder(TCompressedOutputPtr&& oma, NAtrac3::TAtrac3EncoderSettings&& encoderSettings)
    : TBatchedEncoderBase(std::move(oma), MakeAt3Settings(encoderSettings))
{
}

TPCMEngine::TProcessLambda TAtrac3Encoder::GetLambda()
{
    return [this](float* data, const TPCMEngine::ProcessMeta& /*meta*/) { return Push(data); };
}

static atde_settings MakeAt3pSettings(int channels, const TAt3PEnc::TSettings& s)
{
    if (s.UseGha & ~(unsigned)TAt3PEnc::TSettings::GHA_ENABLED)
        throw std::runtime_error("atde_b200: the ATRAC3plus GHA_WIDEBAND experiment is not built");
    atde_settings c;
    atde_default_settings(&c, ATDE_CODEC_ATRAC3PLUS, (int32_t)channels);
    c.gha_flags = s.UseGha;
    return c;
}

TAt3PEnc::TAt3PEnc(TCompressedOutputPtr&& out, int channels, TSettings settings)
    : TBatchedEncoderBase(std::move(out), MakeAt3pSettings(channels, settings))
{
}

TPCMEngine::TProcessLambda TAt3PEnc::GetLambda()
{
    return [this](float* data, const TPCMEngine::ProcessMeta& /*meta*/) { return Push(data); };
}

} // namespace NAtracDEnc
