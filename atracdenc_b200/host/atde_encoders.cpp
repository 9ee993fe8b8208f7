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
    const size_t units = Staged * Units;
    Bytes.resize(units * UnitBytes + 8);
    Sizes.resize(units);
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

} // namespace NAtracDEnc
