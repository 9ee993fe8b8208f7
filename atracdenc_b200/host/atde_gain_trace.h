// atde_gain_trace.h — the reference's `--yaml-log <file>` gain-control trace (§8(f) rank 4) for the GPU path.
//
// The reference writes the log while it encodes: one YAML document per processed frame (src/atrac3denc.cpp:743-750),
// per channel and band the envelope analysis and every decision of the curve builder (src/atrac3denc.cpp:305-579,
// src/transient_detector.cpp:298-446), then the energy scales (src/atrac3denc.cpp:786-798).  Format: src/yaml_log.h:19-57.
//
// Here the encoder runs in batches on the device, so the log is written after each batch from what the device left
// behind (include/atde_b200.h: atde_set_gain_trace and the ATDE_TAP_* buffers): QMF band samples, the sub-frame
// envelope of all four bands with the exact high-frequency ratio and `next_level`, the final curves and the energy
// scales.  The curve builder's INTERMEDIATE values (plateau, sticky ratios, pruned transitions, the point-0 guard's
// scores) never leave the curve kernel, so this file walks the same decisions once more on the host — scalar work,
// one stream, a debugging aid — and checks its final curve against the one the device encoded with: a disagreement
// is an error, not a silently different log.
// Pure host C++ above the C ABI; no CUDA types.
#pragma once
#include "../../include/atde_b200.h"

#include <cstdint>
#include <iosfwd>
#include <vector>

namespace NAtracDEnc {

class TGainTraceWriter {
public:
    // `out` is TAtrac3EncoderSettings::YamlLog (src/atrac/at3/atrac3.h:276); gainControl = !NoGainControll
    TGainTraceWriter(std::ostream* out, int channels, bool gainControl);
    // Call after atde_encode_batch() of ONE stream returned: appends one document per output frame of that batch.
    // Throws std::runtime_error when a tap is missing or the host's curve differs from the device's.
    void AppendBatch(atde_encoder* enc, int64_t outFrames);
    uint64_t FramesWritten() const { return FrameNum; }

private:
    struct TBandState {
        float LastLevel = 0.0f, LastHpfEnergy = 0.0f, LastTarget = 0.0f;   // TCurveBuilderCtx (transient_detector.h)
        float StoredHalf[256] = {};                                       // PcmBuffer.GetFirst(): windowed, modulated
    };
    struct TPoint { uint16_t Level; uint32_t Loc; };
    typedef std::vector<TPoint> TCurve;

    TCurve BuildCurve(const float* gain, const float* low, const float* high, TBandState& st, float minScore);
    void Band(int channel, int band, const float* env, const float* stat, const float* cur, const uint8_t* devCurve);

    std::ostream* Out;
    int Channels;
    bool GainControl;
    uint64_t FrameNum = 0;
    std::vector<TBandState> State;      // [channel][band]
    float Window[256], Level[16], Interp[31];
};

} // namespace NAtracDEnc
