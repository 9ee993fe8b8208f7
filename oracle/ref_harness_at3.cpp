/*
 * oracle/ref_harness_at3.cpp — TEST INFRASTRUCTURE (see ref_harness.cpp header).
 * ATRAC3 stage taps.  Compiled with -fno-access-control so the harness can read the private
 * members of the UNMODIFIED reference TAtrac3Encoder after each lambda call
 * (src/atrac3denc.h:86-133): SingleChannelElements (gain points, tonal blocks, scaled blocks,
 * loudness term, gain energy scales), the tracked Loudness and the QMF look-ahead buffer.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <string>

#include "pcmengin.h"
#include "compressed_io.h"
#include "atrac3denc.h"

using namespace NAtracDEnc;

namespace {
class TNullOut : public ICompressedOutput {
    size_t Ch;
    std::vector<uint8_t>* Bytes;
public:
    TNullOut(size_t ch, std::vector<uint8_t>* b) : Ch(ch), Bytes(b) {}
    void WriteFrame(std::vector<char> data) override { Bytes->insert(Bytes->end(), data.begin(), data.end()); }
    std::string GetName() const override { return "null"; }
    size_t GetChannelNum() const override { return Ch; }
};
}

extern "C" {

/* Per OUTPUT frame f (lambda call f+1), per channel c (frame_rec[f][c]):
 *   float  loud_term; float gscale[4][3] (PrevHalf,CurHalf,Frame);
 *   int32  n_points[4]; int32 points[4][8][2] (level, location);
 *   int32  sfi[32]; float energy[32];
 *   int32  n_tonal; int32 tonal[64][4] (pos, bfu, sfi, len);
 *   float  bands[4][256]  (LookAheadBuf current slot = QMF output of the frame, pre-matrixing)
 * plus tracked[f] = Loudness after the frame.  Returns number of output frames. */
struct at3_rec {
    float loud_term;
    float gscale[4][3];
    int32_t n_points[4];
    int32_t points[4][8][2];
    int32_t sfi[32];
    float energy[32];
    int32_t n_tonal;
    int32_t tonal[64][4];
    float bands[4][256];
};

long ref_at3_stages(int channels, const float* pcm, long n_frames, int bitrate_kbit, int no_gain, int no_tonal,
                    at3_rec* recs, float* tracked, unsigned char* out_bytes, long out_cap, long* out_len)
{
    std::vector<uint8_t> bytes;
    TCompressedOutputPtr out(new TNullOut(channels, &bytes));
    TAtrac3Encoder enc(std::move(out), NAtrac3::TAtrac3EncoderSettings((uint32_t)bitrate_kbit * 1024, no_gain, no_tonal,
                                                                      channels, 0, nullptr));
    auto lambda = enc.GetLambda();
    TPCMEngine::ProcessMeta meta{(uint16_t)channels};
    long produced = 0;
    std::vector<float> frame(1024 * channels);
    for (long f = 0; f < n_frames; f++) {
        memcpy(frame.data(), pcm + (size_t)f * 1024 * channels, frame.size() * sizeof(float));
        /* LookAheadBuf is shifted at the end of the call; capture the current slot before it runs:
           after the previous call's shift, [128..383] is the frame about to be encoded. */
        float cur[2][4][256];
        for (int c = 0; c < channels; c++)
            for (int b = 0; b < 4; b++)
                memcpy(cur[c][b], &enc.LookAheadBuf[c][b][128], 256 * sizeof(float));
        auto res = lambda(frame.data(), meta);
        if (res != TPCMEngine::EProcessResult::PROCESSED)
            continue;
        for (int c = 0; c < channels; c++) {
            at3_rec& r = recs[produced * channels + c];
            memset(&r, 0, sizeof(r));
            const auto& sce = enc.SingleChannelElements[c];
            r.loud_term = sce.Loudness;
            for (int b = 0; b < 4; b++) {
                r.gscale[b][0] = sce.GainEnergyScale[b].PrevHalf;
                r.gscale[b][1] = sce.GainEnergyScale[b].CurHalf;
                r.gscale[b][2] = sce.GainEnergyScale[b].Frame;
                const auto& pts = sce.SubbandInfo.GetGainPoints(b);
                r.n_points[b] = (int32_t)pts.size();
                for (size_t i = 0; i < pts.size() && i < 8; i++) {
                    r.points[b][i][0] = (int32_t)pts[i].Level;
                    r.points[b][i][1] = (int32_t)pts[i].Location;
                }
                memcpy(r.bands[b], cur[c][b], 256 * sizeof(float));
            }
            for (size_t i = 0; i < sce.ScaledBlocks.size() && i < 32; i++) {
                r.sfi[i] = sce.ScaledBlocks[i].ScaleFactorIndex;
                r.energy[i] = sce.ScaledBlocks[i].Energy;
            }
            r.n_tonal = (int32_t)sce.TonalBlocks.size();
            for (size_t i = 0; i < sce.TonalBlocks.size() && i < 64; i++) {
                r.tonal[i][0] = sce.TonalBlocks[i].ValPtr->Pos;
                r.tonal[i][1] = sce.TonalBlocks[i].ValPtr->Bfu;
                r.tonal[i][2] = sce.TonalBlocks[i].ScaledBlock.ScaleFactorIndex;
                r.tonal[i][3] = (int32_t)sce.TonalBlocks[i].ScaledBlock.Values.size();
            }
        }
        tracked[produced] = enc.Loudness;
        produced++;
    }
    if ((long)bytes.size() <= out_cap) {
        memcpy(out_bytes, bytes.data(), bytes.size());
        *out_len = (long)bytes.size();
    } else {
        *out_len = -1;
    }
    return produced;
}

int ref_at3_rec_size(void) { return (int)sizeof(at3_rec); }

} // extern "C"
