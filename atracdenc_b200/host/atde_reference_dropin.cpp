// atde_reference_dropin.cpp — the encode hot path of atracdenc on B200, dropped INTO the atracdenc tree.
//
// Compiled as part of the reference's build (src/ on the include path) this file supplies the
// member functions the reference's encoder translation units define — for the reference's OWN, unmodified class
// declarations:
//     TAtrac1Encoder::TAtrac1Encoder / GetLambda               src/atrac1denc.h:105-106,  src/atrac1denc.cpp:60-68,180-254
//     TAtrac3Encoder::TAtrac3Encoder / ~TAtrac3Encoder / GetLambda   src/atrac3denc.h:131-133,  src/atrac3denc.cpp:93-106,694-866
//     TAt3PEnc::TAt3PEnc / GetLambda / ParseAdvancedOpt / TImpl      src/atrac3p.h:59-68,       src/atrac/at3p/at3p.cpp:36-284
//     TAtrac1Decoder::TAtrac1Decoder / GetLambda               src/atrac1denc.h:109-124,  src/atrac1denc.cpp:46-49,139-177
// so that src/main.cpp, src/pcmengin.h, the container writers and every header stay byte for byte what they are:
// main.cpp's `new TAtrac1Encoder(std::move(aeaIO), std::move(encoderSettings))` allocates the reference's class and
// runs THIS constructor.  The reference's encoder state members (filter banks, delay buffers ...) are constructed
// and never used; the frames go to libatde_b200.so through the C ABI.
//
// Where the batching state lives: the classes have no spare member, but each owns its container through a
// TCompressedOutputPtr.  The constructor wraps the container into a TDeferredOutput (an ICompressedOutput that owns
// the real container and the TFrameBatcher) and stores THAT in the member.  When main_() destroys the processor the
// member is destroyed, and ~TDeferredOutput delivers the staged tail to the real container before the container
// itself is closed — the only hook there is, since main.cpp never calls a flush.
//
// Build: the reference's atrac1denc.cpp / atrac3denc.cpp also hold code this path does not replace (TAtrac1MDCT,
// TAtrac1Decoder, TAtrac3MDCT); they stay in the build with the encoder class renamed away by a per-file compile
// definition (-DTAtrac1Encoder=TAtrac1EncoderCpu -DTAtrac1Decoder=TAtrac1DecoderCpu, -DTAtrac3Encoder=TAtrac3EncoderCpu);
// at3p.cpp is dropped.
#ifndef ATDE_USE_REFERENCE_HEADERS
#error "atde_reference_dropin.cpp is built inside the atracdenc tree: define ATDE_USE_REFERENCE_HEADERS and put src/ on the include path"
#endif

#include "atrac1denc.h"
#include "atrac/at1/atrac1_bitalloc.h"         // complete IAtrac1BitAlloc: TAtrac1Encoder's implicit destructor is emitted here
#include "atrac3denc.h"
#include "atrac3p.h"

#include "atde_batcher.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>

namespace NAtracDEnc {

namespace {

// The container as the encoder classes see it.  GetName / GetChannelNum forward to the real container; WriteFrame
// is what the batcher calls when a batch comes back from the GPU.
class TDeferredOutput : public ICompressedOutput {
public:
    TDeferredOutput(TCompressedOutputPtr&& real, const atde_settings& settings)
        : Real(std::move(real))
        , Batcher(settings)
    {}
    ~TDeferredOutput() override
    {
        try {
            Batcher.Flush(*Real);
        } catch (const std::exception& ex) {
            // In the reference this error would have left the lambda, reached main.cpp:718-721 and ended the process
            // with exit code 1.  A destructor cannot throw, and main_() is already returning 0: report it the way
            // main.cpp does and end the process with the reference's exit code (the output file is incomplete).
            std::cerr << "Encode/Decode error: " << ex.what() << " (atde_b200: " << Batcher.Pending()
                      << " staged frame(s) could not be encoded)" << std::endl;
            std::fflush(nullptr);
            std::_Exit(1);
        }
    }
    std::string GetName() const override { return Real->GetName(); }
    size_t GetChannelNum() const override { return Real->GetChannelNum(); }
    void WriteFrame(std::vector<char> data) override { Real->WriteFrame(std::move(data)); }
    TPCMEngine::EProcessResult Push(const float* data) { return Batcher.Push(data, *Real); }
    void EnableGainTrace(std::ostream* log, bool gainControl) { Batcher.EnableGainTrace(log, gainControl); }

private:
    TCompressedOutputPtr Real;                 // destroyed after the flush above
    TFrameBatcher Batcher;
};

TCompressedOutputPtr WrapAt1(TCompressedOutputPtr&& aea, const NAtrac1::TAtrac1EncodeSettings& s)
{
    atde_settings c;
    atde_default_settings(&c, ATDE_CODEC_ATRAC1, (int32_t)aea->GetChannelNum());     // src/atrac1denc.cpp:181
    c.bfu_idx_const = s.GetBfuIdxConst();
    c.window_mode = s.GetWindowMode() == NAtrac1::TAtrac1EncodeSettings::EWindowMode::EWM_AUTO ? 1 : 0;
    c.window_mask = s.GetWindowMask();
    return TCompressedOutputPtr(new TDeferredOutput(std::move(aea), c));
}

TCompressedOutputPtr WrapAt3(TCompressedOutputPtr&& oma, const NAtrac3::TAtrac3EncoderSettings& s)
{
    atde_settings c;
    atde_default_settings(&c, ATDE_CODEC_ATRAC3, (int32_t)s.SourceChannels);
    c.bitrate = s.ConteinerParams->Bitrate;    // an exact table bitrate maps back to the same container
    c.no_gain_control = s.NoGainControll;
    c.no_tonal = s.NoTonalComponents;
    c.bfu_idx_const = s.BfuIdxConst;
    TDeferredOutput* d = new TDeferredOutput(std::move(oma), c);
    TCompressedOutputPtr res(d);
    d->EnableGainTrace(s.YamlLog, !s.NoGainControll);       // `--yaml-log <file>` (src/main.cpp:661-673)
    return res;
}

TCompressedOutputPtr WrapAt3p(TCompressedOutputPtr&& out, int channels, const TAt3PEnc::TSettings& s)
{
    if (s.UseGha & ~(unsigned)TAt3PEnc::TSettings::GHA_ENABLED)
        throw std::runtime_error("atde_b200: the ATRAC3plus GHA_WIDEBAND experiment is not built");
    atde_settings c;
    atde_default_settings(&c, ATDE_CODEC_ATRAC3PLUS, (int32_t)channels);
    c.gha_flags = s.UseGha;
    return TCompressedOutputPtr(new TDeferredOutput(std::move(out), c));
}

TPCMEngine::TProcessLambda MakeLambda(ICompressedOutput* member)
{
    TDeferredOutput* d = static_cast<TDeferredOutput*>(member);
    return [d](float* data, const TPCMEngine::ProcessMeta& /*meta*/) { return d->Push(data); };
}

// The decoder side: the container as TAtrac1Decoder sees it.  The reference's lambda reads one sound unit per channel
// per call and decodes it on the spot; here the units are read AHEAD in batches, decoded on the GPU
// (atde_decode_batch) and handed out frame by frame.  A read error (the CLI's frame loop runs past the end of some
// files, src/main.cpp:697-705 with pcmengin.h's 4096-sample buffer) is kept and rethrown by the call that would have
// hit it in the reference.
class TPrefetchInput : public ICompressedInput {
public:
    explicit TPrefetchInput(TCompressedInputPtr&& real)
        : Real(std::move(real))
        , Channels((int)Real->GetChannelNum())
    {
        if (atde_decoder_create(Channels, 0, &Dec) < 0)
            throw std::runtime_error(std::string("atde_b200: ") + atde_last_error());
        if (const char* env = getenv("ATDE_BATCH_FRAMES")) {
            const long v = atol(env);
            if (v > 0) BatchFrames = (size_t)v;
        }
    }
    ~TPrefetchInput() override { atde_decoder_destroy(Dec); }
    std::unique_ptr<TFrame> ReadFrame() override { return Real->ReadFrame(); }
    uint64_t GetLengthInSamples() const override { return Real->GetLengthInSamples(); }
    std::string GetName() const override { return Real->GetName(); }
    size_t GetChannelNum() const override { return Real->GetChannelNum(); }

    TPCMEngine::EProcessResult Next(float* data)
    {
        if (Served == Ready) Refill();
        memcpy(data, &Pcm[Served * 512 * Channels], sizeof(float) * 512 * Channels);
        Served++;
        return TPCMEngine::EProcessResult::PROCESSED;
    }

private:
    void Refill()
    {
        if (Pending) std::rethrow_exception(Pending);
        // read ahead only inside the length the container announces (the caller decodes at least that much: src/main.cpp:351,
        // 697-705); beyond it, one frame per call like the reference, so that no read happens that it would not do
        const uint64_t announced = Real->GetLengthInSamples() / 512;
        size_t want = Consumed < announced ? (size_t)std::min<uint64_t>(BatchFrames, announced - Consumed) : 1;
        Units.resize(want * Channels * 212);
        size_t frames = 0;
        try {
            for (; frames < want; frames++)
                for (int ch = 0; ch < Channels; ch++) {
                    std::unique_ptr<TFrame> f(Real->ReadFrame());
                    memcpy(&Units[(frames * Channels + ch) * 212], f->Get(), 212);
                }
        } catch (...) {
            Pending = std::current_exception();      // the frames read so far are still delivered first
        }
        if (frames == 0) std::rethrow_exception(Pending);
        Pcm.resize(frames * 512 * Channels);
        if (atde_decode_batch(Dec, Units.data(), 1, (int64_t)frames, Pcm.data()) < 0)
            throw std::runtime_error(std::string("atde_b200: ") + atde_last_error());
        Served = 0;
        Ready = frames;
        Consumed += frames;
    }
    TCompressedInputPtr Real;
    int Channels;
    atde_decoder* Dec = nullptr;
    std::vector<uint8_t> Units;
    std::vector<float> Pcm;
    size_t Served = 0, Ready = 0, BatchFrames = 1024;
    uint64_t Consumed = 0;
    std::exception_ptr Pending;
};

} // namespace

// ---- ATRAC1 decoder (src/atrac1denc.h:109-124) ----
TAtrac1Decoder::TAtrac1Decoder(TCompressedInputPtr&& aea)
    : Aea(new TPrefetchInput(std::move(aea)))
{
}

TPCMEngine::TProcessLambda TAtrac1Decoder::GetLambda()
{
    TPrefetchInput* in = static_cast<TPrefetchInput*>(Aea.get());
    return [in](float* data, const TPCMEngine::ProcessMeta& /*meta*/) { return in->Next(data); };
}

// ---- ATRAC1 (src/atrac1denc.h:57-107) ----
TAtrac1Encoder::TAtrac1Encoder(TCompressedOutputPtr&& aea, NAtrac1::TAtrac1EncodeSettings&& settings)
    : Aea(WrapAt1(std::move(aea), settings))
    , Settings(std::move(settings))
    , LoudnessCurve()
{
}

TPCMEngine::TProcessLambda TAtrac1Encoder::GetLambda()
{
    return MakeLambda(Aea.get());
}

// ---- ATRAC3 (src/atrac3denc.h:95-134) ----
TAtrac3Encoder::TAtrac3Encoder(TCompressedOutputPtr&& oma, NAtrac3::TAtrac3EncoderSettings&& encoderSettings)
    : Oma(WrapAt3(std::move(oma), encoderSettings))
    , Params(std::move(encoderSettings))
    , LoudnessCurve()
    , Upsampler(11025.0f, 800.0f)
{
    YamlLog = Params.YamlLog;                  // written batch by batch by the TDeferredOutput's batcher (atde_gain_trace.cpp)
}

TAtrac3Encoder::~TAtrac3Encoder()
{
}

TPCMEngine::TProcessLambda TAtrac3Encoder::GetLambda()
{
    return MakeLambda(Oma.get());
}

// ---- ATRAC3plus (src/atrac3p.h:28-71): the reference keeps its state behind a pimpl; this path needs none ----
class TAt3PEnc::TImpl {
};

TAt3PEnc::TAt3PEnc(TCompressedOutputPtr&& out, int channels, TSettings settings)
    : Out(WrapAt3p(std::move(out), channels, settings))
    , Channels(channels)
{
}

TPCMEngine::TProcessLambda TAt3PEnc::GetLambda()
{
    return MakeLambda(Out.get());
}

void TAt3PEnc::ParseAdvancedOpt(const char* opt, TSettings& settings)
{
    ParseAt3pAdvancedOpt(opt, settings.UseGha, settings.WidebandRefineMode);
}

} // namespace NAtracDEnc
