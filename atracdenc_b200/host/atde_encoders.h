// atde_encoders.h — host-side mirror of the reference's frame processors, backed by the C ABI of
// libatde_b200.so (include/atde_b200.h).  Same class names, constructor signatures and lambda
// semantics as
//   NAtracDEnc::TAtrac1Encoder(TCompressedOutputPtr&&, NAtrac1::TAtrac1EncodeSettings&&)   src/atrac1denc.h:105
//   NAtracDEnc::TAtrac3Encoder(TCompressedOutputPtr&&, NAtrac3::TAtrac3EncoderSettings&&)  src/atrac3denc.h:131
//   NAtracDEnc::TAt3PEnc(TCompressedOutputPtr&&, int channels, TSettings)                  src/atrac3p.h:59
// for programs built WITHOUT the atracdenc tree (this repo's tests, a batch transcoder over the library).  Inside the
// atracdenc tree the reference's own class declarations are kept and atde_reference_dropin.cpp supplies their
// member functions instead — this header is not used there.
// Differences a caller can observe: WriteFrame calls are DEFERRED — frames are staged and encoded on the GPU in
// batches; payload bytes, lengths and call order are identical, and everything is flushed by Flush() or the
// destructor (the processor owns the container).  Call Flush() before destruction to get errors as exceptions; a
// failure inside the destructor is reported on stderr and through TBatchedEncoderBase::FlushFailures().
#pragma once
#include "atde_batcher.h"

#include <iosfwd>
#include <stdexcept>

#ifdef ATDE_USE_REFERENCE_HEADERS
#include "atrac/at1/atrac1.h"                  // NAtrac1::TAtrac1EncodeSettings
#include "atrac/at3/atrac3.h"                  // NAtrac3::TAtrac3EncoderSettings, TContainerParams
#endif

namespace NAtracDEnc {

#ifdef ATDE_USE_REFERENCE_HEADERS
// Built next to the reference's headers the mirror classes live in their own namespace: the reference's
// atrac1denc.h / atrac3denc.h / atrac3p.h declare classes of the same names.
namespace NAtdeMirror {
#endif

#ifndef ATDE_USE_REFERENCE_HEADERS
namespace NAtrac1 {
// src/atrac/at1/atrac1.h:33-54
class TAtrac1EncodeSettings {
public:
    enum class EWindowMode { EWM_NOTRANSIENT, EWM_AUTO };
    TAtrac1EncodeSettings() {}
    TAtrac1EncodeSettings(uint32_t bfuIdxConst, EWindowMode windowMode, uint32_t windowMask)
        : BfuIdxConst(bfuIdxConst), WindowMode(windowMode), WindowMask(windowMask) {}
    uint32_t GetBfuIdxConst() const { return BfuIdxConst; }
    EWindowMode GetWindowMode() const { return WindowMode; }
    uint32_t GetWindowMask() const { return WindowMask; }
private:
    const uint32_t BfuIdxConst = 0;
    EWindowMode WindowMode = EWindowMode::EWM_AUTO;
    const uint32_t WindowMask = 0;
};
} // namespace NAtrac1

namespace NAtrac3 {
// src/atrac/at3/atrac3.h:33-37,211-220,260-277
struct TContainerParams {
    const uint32_t Bitrate;
    const uint16_t FrameSz;
    const bool Js;
};
const TContainerParams* GetContainerParamsForBitrate(uint32_t bitrate);   // TAtrac3Data::GetContainerParamsForBitrate
struct TAtrac3EncoderSettings {
    TAtrac3EncoderSettings(uint32_t bitrate, bool noGainControll, bool noTonalComponents, uint8_t sourceChannels,
                           uint32_t bfuIdxConst, std::ostream* yamlLog = nullptr)
        : ConteinerParams(GetContainerParamsForBitrate(bitrate))
        , NoGainControll(noGainControll)
        , NoTonalComponents(noTonalComponents)
        , SourceChannels(sourceChannels)
        , BfuIdxConst(bfuIdxConst)
        , YamlLog(yamlLog)
    { }
    const TContainerParams* ConteinerParams;
    const bool NoGainControll;
    const bool NoTonalComponents;
    const uint8_t SourceChannels;
    const uint32_t BfuIdxConst;
    std::ostream* YamlLog;   // nullable; gain control debug log (`--yaml-log`), written batch by batch (atde_gain_trace.h)
};
} // namespace NAtrac3
#endif

// Common staging/flush machinery for all codecs.
class TBatchedEncoderBase : public IProcessor {
public:
    ~TBatchedEncoderBase() override;
    // Encode and deliver everything staged so far (idempotent).
    // Throws std::runtime_error on failure and keeps the staged frames (a later call retries them).
    void Flush() { Batcher.Flush(*Out); }
    // Frames staged before a batch is sent to the GPU (default 4096; tests use small values).
    void SetBatchFrames(size_t n) { Batcher.SetBatchFrames(n); }
    // Number of destructors (process-wide) whose final flush failed: the frames were lost and the error printed.
    static int FlushFailures();
protected:
    TBatchedEncoderBase(TCompressedOutputPtr&& out, const atde_settings& settings);
    TPCMEngine::EProcessResult Push(const float* data) { return Batcher.Push(data, *Out); }
    TCompressedOutputPtr Out;
    TFrameBatcher Batcher;
};

class TAtrac1Encoder : public TBatchedEncoderBase {
public:
    TAtrac1Encoder(TCompressedOutputPtr&& aea, NAtrac1::TAtrac1EncodeSettings&& settings);
    TPCMEngine::TProcessLambda GetLambda() override;
};

// The first lambda call returns LOOK_AHEAD (src/atrac3denc.cpp:715-718), every later one PROCESSED;
// one WriteFrame of exactly FrameSz bytes per PROCESSED call (src/atrac/at3/atrac3_bitstream.cpp:845).
class TAtrac3Encoder : public TBatchedEncoderBase {
public:
    TAtrac3Encoder(TCompressedOutputPtr&& oma, NAtrac3::TAtrac3EncoderSettings&& encoderSettings);
    TPCMEngine::TProcessLambda GetLambda() override;
};

// ATRAC3plus (src/atrac3p.h:28-71).  The first lambda call returns LOOK_AHEAD (at3p.cpp:109-111), every
// later one PROCESSED with one WriteFrame of 2048 bytes (at3p_bitstream.cpp:724-725).  The three
// processing flags of UseGha (`ghadbg`) are honoured; GHA_WIDEBAND (an opt-in experiment) throws.
class TAt3PEnc : public TBatchedEncoderBase {
public:
    struct TSettings {
        enum GhaProcessingFlags : uint8_t {
            GHA_PASS_INPUT = 1, GHA_WRITE_TONAL = 1 << 1, GHA_WRITE_RESIUDAL = 1 << 2, GHA_WIDEBAND = 1 << 3,
            GHA_ENABLED = GHA_PASS_INPUT | GHA_WRITE_TONAL | GHA_WRITE_RESIUDAL
        };
        uint8_t UseGha;
        uint8_t WidebandRefineMode;
        TSettings() : UseGha(GHA_ENABLED), WidebandRefineMode(0) {}
    };
    TAt3PEnc(TCompressedOutputPtr&& out, int channels, TSettings settings);
    TPCMEngine::TProcessLambda GetLambda() override;
    static constexpr int NumSamples = 2048;
    // src/atrac/at3p/at3p.cpp:224-284 (called at src/main.cpp:480)
    static void ParseAdvancedOpt(const char* opt, TSettings& settings);
};

#ifdef ATDE_USE_REFERENCE_HEADERS
} // namespace NAtdeMirror
#endif

} // namespace NAtracDEnc
