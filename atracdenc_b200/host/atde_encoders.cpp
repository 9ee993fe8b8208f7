// atde_encoders.cpp — see atde_encoders.h.  Pure host C++ over the C ABI; no CUDA types here.
#include "atde_encoders.h"

#include <atomic>
#include <iostream>
#include <string>

namespace NAtracDEnc {

#ifdef ATDE_USE_REFERENCE_HEADERS
namespace NAtdeMirror {
#endif

static std::atomic<int> g_flush_failures{0};

TBatchedEncoderBase::TBatchedEncoderBase(TCompressedOutputPtr&& out, const atde_settings& settings)
    : Out(std::move(out))
    , Batcher(settings)
{
}

TBatchedEncoderBase::~TBatchedEncoderBase()
{
    try {
        Flush();
    } catch (const std::exception& ex) {
        // destructors must not throw.  In the reference this error would have left the lambda and reached
        // main.cpp's catch (exit code 1): make it visible — callers that need the exception call Flush() first.
        g_flush_failures++;
        std::cerr << "atde_b200: final flush failed, " << Batcher.Pending() << " staged frame(s) lost: " << ex.what() << std::endl;
    } catch (...) {
        g_flush_failures++;
        std::cerr << "atde_b200: final flush failed, " << Batcher.Pending() << " staged frame(s) lost" << std::endl;
    }
}

int TBatchedEncoderBase::FlushFailures() { return g_flush_failures.load(); }

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
    Batcher.EnableGainTrace(encoderSettings.YamlLog, !encoderSettings.NoGainControll);
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

void TAt3PEnc::ParseAdvancedOpt(const char* opt, TSettings& settings)
{
    ParseAt3pAdvancedOpt(opt, settings.UseGha, settings.WidebandRefineMode);
}

#ifdef ATDE_USE_REFERENCE_HEADERS
} // namespace NAtdeMirror
#endif

} // namespace NAtracDEnc
