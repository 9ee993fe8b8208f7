// atde_batcher.h — the staging / batching engine behind every host-side frame processor of this repo.
//
// One TFrameBatcher == one reference encoder instance (one stream).  The frame lambda the reference's PCM engine
// calls (src/pcmengin.h:150,152-192) hands it one frame of interleaved PCM at a time; frames are staged, encoded on
// the GPU in batches through the C ABI (include/atde_b200.h) and delivered to the container with the reference's
// WriteFrame calls — same bytes, same lengths, same order, only deferred:
//   ATRAC1      one call per channel per frame, channel 0 first (src/atrac1denc.cpp:249-251), raw TBitStream length
//   ATRAC3      one call of exactly FrameSz bytes per PROCESSED lambda call (src/atrac/at3/atrac3_bitstream.cpp:845)
//   ATRAC3plus  one call of 2048 bytes per PROCESSED call (src/atrac/at3p/at3p_bitstream.cpp:724-725)
// Pure host C++; no CUDA types.
#pragma once
#include "atde_boundary.h"
#include "../../include/atde_b200.h"

#include <iosfwd>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace NAtracDEnc {

class TGainTraceWriter;

class TFrameBatcher {
public:
    // Throws std::runtime_error (the type src/main.cpp:709-720 catches) when the C ABI refuses, e.g. no GPU.
    explicit TFrameBatcher(const atde_settings& settings);
    ~TFrameBatcher();                                  // releases the handle; does NOT flush (the owner decides how)
    TFrameBatcher(const TFrameBatcher&) = delete;
    TFrameBatcher& operator=(const TFrameBatcher&) = delete;

    // One lambda call: copies FrameSamples*Channels floats (the pointer is only valid during the call) and returns
    // what the reference's lambda returns for this call: LOOK_AHEAD for the first `lookahead` calls, else PROCESSED.
    // A full staging buffer is encoded and delivered to `out` before returning; errors throw and leave the staged
    // frames in place (a later Flush() retries them).
    TPCMEngine::EProcessResult Push(const float* data, ICompressedOutput& out);
    // Encodes and delivers everything staged so far (idempotent).  Throws on failure; staged frames are only
    // dropped after atde_encode_batch() succeeded.
    void Flush(ICompressedOutput& out);
    // Frames staged before a batch goes to the GPU (default 4096, or ATDE_BATCH_FRAMES from the environment).
    void SetBatchFrames(size_t n) { BatchFrames = n ? n : 1; }
    size_t Pending() const { return Staged; }
    // ATRAC3: the reference's `--yaml-log` gain-control trace (TAtrac3EncoderSettings::YamlLog, src/atrac/at3/atrac3.h:276):
    // every batch that comes back from the GPU appends its frames' documents to `log` (atde_gain_trace.h).
    void EnableGainTrace(std::ostream* log, bool gainControl);
    int GetChannels() const { return Channels; }

private:
    atde_encoder* Enc = nullptr;
    int Channels = 0, FrameSamples = 0, Units = 0, UnitBytes = 0, LookAhead = 0;
    std::vector<float> Stage;       // [BatchFrames][FrameSamples][Channels]
    std::vector<uint8_t> Bytes;
    std::vector<int32_t> Sizes;
    size_t Staged = 0, BatchFrames = 4096;
    uint64_t Calls = 0;
    std::unique_ptr<TGainTraceWriter> Trace;
};

// TAt3PEnc::ParseAdvancedOpt (src/atrac/at3p/at3p.cpp:196-284): "key=value[,key=value...]" with the keys `ghadbg`
// (GHA processing mask 0..15, flag names echoed to stderr) and `ghawbrefine` (0 | 1); same errors as the reference
// (std::runtime_error / std::invalid_argument from std::stoi).  Writes through two plain bytes so that it serves
// both the mirror class and the reference's own TAt3PEnc::TSettings.
void ParseAt3pAdvancedOpt(const char* opt, uint8_t& useGha, uint8_t& widebandRefineMode);

} // namespace NAtracDEnc
