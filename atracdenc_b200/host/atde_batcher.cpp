// atde_batcher.cpp — see atde_batcher.h.
#include "atde_batcher.h"
#include "atde_gain_trace.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

namespace NAtracDEnc {

static void Check(int rc)
{
    if (rc < 0)
        throw std::runtime_error(std::string("atde_b200: ") + atde_last_error());
}

TFrameBatcher::TFrameBatcher(const atde_settings& settings)
{
    Check(atde_create(&settings, &Enc));
    Channels = settings.channels;
    FrameSamples = atde_frame_samples(Enc);
    Units = atde_units_per_frame(Enc);
    UnitBytes = atde_unit_bytes(Enc);
    LookAhead = atde_lookahead_frames(Enc);
    if (const char* env = getenv("ATDE_BATCH_FRAMES")) {
        const long v = atol(env);
        if (v > 0) BatchFrames = (size_t)v;
    }
}

TFrameBatcher::~TFrameBatcher()
{
    atde_destroy(Enc);
}

void TFrameBatcher::EnableGainTrace(std::ostream* log, bool gainControl)
{
    if (!log)
        return;
    Check(atde_set_gain_trace(Enc, 1));
    Trace.reset(new TGainTraceWriter(log, Channels, gainControl));
}

TPCMEngine::EProcessResult TFrameBatcher::Push(const float* data, ICompressedOutput& out)
{
    const size_t n = (size_t)FrameSamples * Channels;
    if (Staged >= BatchFrames)                         // a previous flush failed and the caller went on
        Flush(out);
    if (Stage.size() < BatchFrames * n)
        Stage.resize(BatchFrames * n);
    memcpy(&Stage[Staged * n], data, n * sizeof(float));
    Staged++;
    const bool lookAhead = Calls < (uint64_t)LookAhead;
    Calls++;
    if (Staged == BatchFrames)
        Flush(out);
    return lookAhead ? TPCMEngine::EProcessResult::LOOK_AHEAD : TPCMEngine::EProcessResult::PROCESSED;
}

void TFrameBatcher::Flush(ICompressedOutput& out)
{
    if (!Staged)
        return;
    // a look-ahead codec's first batch yields one frame less than it consumes
    const size_t units = (size_t)atde_output_frames(Enc, (int64_t)Staged) * Units;
    Bytes.resize(units * UnitBytes + 8);
    Sizes.resize(units + 1);
    Check(atde_encode_batch(Enc, Stage.data(), 1, (int64_t)Staged, Bytes.data(), Sizes.data()));
    Staged = 0;                                        // only now: a failed batch stays staged
    if (Trace)
        Trace->AppendBatch(Enc, (int64_t)(units / Units));
    for (size_t u = 0; u < units; u++) {
        // same bytes, same length, same order as the reference's WriteFrame calls; payload bytes beyond the
        // container frame size are the bit writer's zero growth slack
        const char* p = reinterpret_cast<const char*>(&Bytes[u * UnitBytes]);
        std::vector<char> frame((size_t)Sizes[u], 0);
        memcpy(frame.data(), p, (size_t)Sizes[u] < (size_t)UnitBytes ? (size_t)Sizes[u] : (size_t)UnitBytes);
        out.WriteFrame(std::move(frame));
    }
}

namespace {
void SetGhaMask(const std::string& str, uint8_t& useGha, uint8_t&)
{
    const int mask = std::stoi(str);
    if (mask > 15 || mask < 0)
        throw std::runtime_error("invalud value of GHA processing mask");   // the reference's wording
    if (mask & 1) std::cerr << "GHA_PASS_INPUT" << std::endl;
    if (mask & 4) std::cerr << "GHA_WRITE_RESIUDAL" << std::endl;
    if (mask & 2) std::cerr << "GHA_WRITE_TONAL" << std::endl;
    if (mask & 8) std::cerr << "GHA_WIDEBAND" << std::endl;
    useGha = (uint8_t)mask;
}
void SetWidebandRefine(const std::string& str, uint8_t&, uint8_t& mode)
{
    const int v = std::stoi(str);
    if (v < 0 || v > 1)
        throw std::runtime_error("invalid ghawbrefine value (expected 0=subband or 1=raw)");
    mode = (uint8_t)v;
    std::cerr << "GHA_WIDEBAND_REFINE=" << (v == 1 ? "raw" : "subband") << std::endl;
}
} // namespace

void ParseAt3pAdvancedOpt(const char* opt, uint8_t& useGha, uint8_t& widebandRefineMode)
{
    if (opt == nullptr)
        return;
    // a key token runs up to '=', a value token up to ',' or the end; the grammar's error cases are the reference's
    const std::string s(opt);
    size_t pos = 0;
    for (;;) {
        size_t k = pos;
        while (k < s.size() && s[k] != '=' && s[k] != ',') k++;
        if (k == s.size()) throw std::runtime_error("unexpected end of key token");
        if (s[k] == ',') throw std::runtime_error("unexpected \",\" just after key.");
        const std::string key = s.substr(pos, k - pos);
        void (*handler)(const std::string&, uint8_t&, uint8_t&) = nullptr;
        if (key == "ghadbg") handler = &SetGhaMask;
        else if (key == "ghawbrefine") handler = &SetWidebandRefine;
        else throw std::runtime_error(std::string("unexpected advanced option \"") + key);
        size_t v = k + 1, e = v;
        while (e < s.size() && s[e] != ',') {
            if (s[e] == '=') throw std::runtime_error("unexpected \"=\" inside value token.");
            e++;
        }
        if (e > v) handler(s.substr(v, e - v), useGha, widebandRefineMode);
        if (e == s.size()) return;
        pos = e + 1;                                   // after ',': another key must follow
    }
}

} // namespace NAtracDEnc
