/*
 * oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * Thin in-memory driver around the UNMODIFIED reference encoder sources in
 * /root/reference (compiled where they lie by oracle/Makefile into
 * oracle/_ref/libatde_ref.so).  It replaces only the two I/O edges of the
 * reference:
 *   - IPCMReader  (src/pcmengin.h:104-109)  -> memory reader that mirrors
 *     TWav::GetPCMReader (src/wav.cpp:46-61) including its TPCMBuffer::Zero call
 *   - ICompressedOutput (src/compressed_io.h:56-59) -> frame-capturing sink
 * and then runs exactly the loop of src/main.cpp:697-716.
 *
 * Nothing here is linked into, imported by, or reachable from the product
 * library (atracdenc_b200/csrc).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the resulting .so.
 *
 * The second half ("stage taps") drives the reference's public sub-objects
 * directly to expose intermediates (QMF bands, window masks, spectra, loudness,
 * scale factor indices) for bisecting a parity failure.
 */
#include <cstdint>
#include <cstdlib>
#include <new>
#include <cstring>
#include <memory>
#include <vector>
#include <string>
#include <stdexcept>

#include "pcmengin.h"
#include "compressed_io.h"
#include "atrac1denc.h"
#include "atrac3denc.h"
#include "atrac3p.h"
#include "atrac/at1/atrac1.h"
#include "atrac/at1/atrac1_qmf.h"
#include "atrac/at1/atrac1_bitalloc.h"
#include "atrac/at3/atrac3.h"
#include "atrac/atrac_scale.h"
#include "atrac/atrac_psy_common.h"
#include "transient_detector.h"
#include "util.h"
#include "lib/mdct/mdct.h"
#include <lib/fft/kissfft_impl/tools/kiss_fftr.h>

using namespace NAtracDEnc;

namespace {

/* Memory PCM source: same contract as TWav's reader (src/wav.cpp:46-61). */
class TMemReader : public IPCMReader {
    const float* Data;
    const uint64_t Total;       /* sample-frames available */
    const uint16_t Channels;
    mutable uint64_t Pos = 0;
public:
    TMemReader(const float* d, uint64_t total, uint16_t ch) : Data(d), Total(total), Channels(ch) {}
    bool Read(TPCMBuffer& buf, const uint32_t size) const override {
        uint64_t left = Total - Pos;
        uint32_t n = left < size ? (uint32_t)left : size;
        if (n == 0)
            return false;
        memcpy(buf[0], Data + Pos * Channels, sizeof(float) * (size_t)n * Channels);
        Pos += n;
        if (n != size)
            buf.Zero(n, size - n);   /* sic: the reference zeroes BYTES, keep its behaviour */
        return true;
    }
};

struct TSink {
    std::vector<uint8_t> Bytes;
    std::vector<int32_t> Sizes;
};

class TCapture : public ICompressedOutput {
    TSink* Sink;
    size_t Ch;
public:
    TCapture(TSink* s, size_t ch) : Sink(s), Ch(ch) {}
    void WriteFrame(std::vector<char> data) override {
        Sink->Sizes.push_back((int32_t)data.size());
        Sink->Bytes.insert(Sink->Bytes.end(), data.begin(), data.end());
    }
    std::string GetName() const override { return "capture"; }
    size_t GetChannelNum() const override { return Ch; }
};

} // namespace

extern "C" {

/*
 * Runs the reference encoder over an in-memory PCM stream.
 *  codec: 1 = ATRAC1, 3 = ATRAC3, 4 = ATRAC3plus
 *  pcm: interleaved normalised floats, n_avail sample-frames; loop runs while
 *       total_samples > processed (src/main.cpp:701).
 *  bitrate_kbit: ATRAC3 only (0 = LP2 default, 64 = LP4), passed *1024 like main.cpp:671.
 * Payloads are appended un-padded in WriteFrame call order; sizes[] receives each length.
 * Returns number of WriteFrame calls, or -1 on overflow of the caller's buffers / exception.
 */
long ref_encode(int codec, int channels, const float* pcm, long n_avail, long total_samples,
                int bitrate_kbit, int no_gain, int no_tonal, int bfu_idx_const,
                int at1_window_auto, int at1_window_mask,
                unsigned char* out, long out_cap, int* sizes, long sizes_cap, long* out_bytes)
{
    TSink sink;
    try {
        std::unique_ptr<IProcessor> proc;
        size_t step = 0;
        TCompressedOutputPtr cap(new TCapture(&sink, channels));
        if (codec == 1) {
            using NAtrac1::TAtrac1EncodeSettings;
            proc.reset(new TAtrac1Encoder(std::move(cap),
                TAtrac1EncodeSettings(bfu_idx_const,
                    at1_window_auto ? TAtrac1EncodeSettings::EWindowMode::EWM_AUTO
                                    : TAtrac1EncodeSettings::EWindowMode::EWM_NOTRANSIENT,
                    at1_window_mask)));
            step = 512;
        } else if (codec == 3) {
            proc.reset(new TAtrac3Encoder(std::move(cap),
                NAtrac3::TAtrac3EncoderSettings((uint32_t)bitrate_kbit * 1024, no_gain, no_tonal,
                                                channels, bfu_idx_const, nullptr)));
            step = 1024;
        } else if (codec == 4) {
            proc.reset(new TAt3PEnc(std::move(cap), channels, TAt3PEnc::TSettings()));
            step = 2048;
        } else {
            return -1;
        }
        TPCMEngine eng(4096, channels, TPCMEngine::TReaderPtr(new TMemReader(pcm, n_avail, channels)));
        auto lambda = proc->GetLambda();
        try {
            while ((uint64_t)total_samples > eng.ApplyProcess(step, lambda)) {}
        } catch (const TNoDataToRead&) {
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_encode: %s\n", e.what());
        return -1;
    }
    if ((long)sink.Bytes.size() > out_cap || (long)sink.Sizes.size() > sizes_cap)
        return -1;
    memcpy(out, sink.Bytes.data(), sink.Bytes.size());
    memcpy(sizes, sink.Sizes.data(), sink.Sizes.size() * sizeof(int));
    *out_bytes = (long)sink.Bytes.size();
    return (long)sink.Sizes.size();
}

/* ------------------------------------------------------------------------------------------
 * Stage taps, ATRAC1.  Drives Atrac1AnalysisFilterBank / TTransientDetector / TAtrac1MDCT /
 * TScaler / CreateLoudnessCurve / TrackLoudness in the order of the encoder lambda
 * (src/atrac1denc.cpp:201-254) over n_frames frames of `channels` interleaved channels.
 * Outputs (any may be NULL), all indexed [frame][channel]...:
 *   bands   [F][C][512]  low(128) | mid(128) | hi(256) as handed to Mdct (pre-window)
 *   masks   [F][C]       window mask bits (low=1, mid=2, hi=4)
 *   specs   [F][C][512]
 *   chloud  [F][C]       per-channel weighted energy
 *   loud    [F]          tracked Loudness after the frame's update
 *   sfi     [F][C][52]   scale factor indices
 */
void ref_at1_stages(int channels, const float* pcm, long n_frames, int window_auto, int window_mask,
                    float* bands, unsigned char* masks, float* specs, float* chloud, float* loud,
                    unsigned char* sfi)
{
    using NAtrac1::TAtrac1Data;
    std::vector<Atrac1AnalysisFilterBank> fb(channels);
    std::vector<TTransientDetector> dl(channels, TTransientDetector(16, 128));
    std::vector<TTransientDetector> dm(channels, TTransientDetector(16, 128));
    std::vector<TTransientDetector> dh(channels, TTransientDetector(16, 256));
    TAtrac1MDCT mdct;
    TScaler<TAtrac1Data> scaler;
    const std::vector<float> curve = CreateLoudnessCurve(512);
    std::vector<std::vector<float>> lo(channels, std::vector<float>(256 + 16, 0.f));
    std::vector<std::vector<float>> mi(channels, std::vector<float>(256 + 16, 0.f));
    std::vector<std::vector<float>> hi(channels, std::vector<float>(512 + 16, 0.f));
    float loudness = 0.006f;

    for (long f = 0; f < n_frames; f++) {
        uint32_t wm[2] = {0, 0};
        float cl[2] = {0, 0};
        for (int ch = 0; ch < channels; ch++) {
            float src[512];
            for (int i = 0; i < 512; i++)
                src[i] = pcm[((size_t)f * 512 + i) * channels + ch];
            fb[ch].Analysis(src, lo[ch].data(), mi[ch].data(), hi[ch].data());
            if (bands) {
                float* b = bands + ((size_t)f * channels + ch) * 512;
                memcpy(b, lo[ch].data(), 128 * sizeof(float));
                memcpy(b + 128, mi[ch].data(), 128 * sizeof(float));
                memcpy(b + 256, hi[ch].data(), 256 * sizeof(float));
            }
            if (window_auto) {
                wm[ch] |= (uint32_t)dl[ch].Detect(lo[ch].data());
                const std::vector<float>& im = InvertSpectr<128>(mi[ch].data());
                wm[ch] |= (uint32_t)dm[ch].Detect(im.data()) << 1;
                const std::vector<float>& ih = InvertSpectr<256>(hi[ch].data());
                wm[ch] |= (uint32_t)dh[ch].Detect(ih.data()) << 2;
            } else {
                wm[ch] = window_mask;
            }
            TAtrac1Data::TBlockSizeMod bs(wm[ch] & 1, wm[ch] & 2, wm[ch] & 4);
            std::vector<float> sp(512);
            mdct.Mdct(sp.data(), lo[ch].data(), mi[ch].data(), hi[ch].data(), bs);
            float l = 0.0;
            for (size_t i = 0; i < 512; i++) {
                float e = sp[i] * sp[i];
                l += e * curve[i];
            }
            cl[ch] = l;
            if (masks) masks[f * channels + ch] = (unsigned char)wm[ch];
            if (specs) memcpy(specs + ((size_t)f * channels + ch) * 512, sp.data(), 512 * sizeof(float));
            if (chloud) chloud[f * channels + ch] = l;
            if (sfi) {
                auto blocks = scaler.ScaleFrame(sp, bs);
                for (size_t i = 0; i < blocks.size(); i++)
                    sfi[((size_t)f * channels + ch) * 52 + i] = blocks[i].ScaleFactorIndex;
            }
        }
        if (channels == 2 && wm[0] == 0 && wm[1] == 0)
            loudness = TrackLoudness(loudness, cl[0], cl[1]);
        else if (wm[0] == 0)
            loudness = TrackLoudness(loudness, cl[0]);
        if (loud) loud[f] = loudness;
    }
}

/* Table taps: what the reference computed at start-up on THIS machine's libm. */
void ref_tables_at1(float* qmf_window48, float* sine_window32, float* scale_table64,
                    float* loudness_curve512, float* ath52)
{
    struct TQ : public TQmfCommon { TQ() : TQmfCommon() {} static const float* W() { return QmfWindow; } };
    TQ q;
    memcpy(qmf_window48, TQ::W(), 48 * sizeof(float));
    NAtrac1::TAtrac1Data d;
    memcpy(sine_window32, NAtrac1::TAtrac1Data::SineWindow, 32 * sizeof(float));
    memcpy(scale_table64, NAtrac1::TAtrac1Data::ScaleTable, 64 * sizeof(float));
    auto c = CreateLoudnessCurve(512);
    memcpy(loudness_curve512, c.data(), 512 * sizeof(float));
    /* ATH per BFU exactly as CalcAt1ATH (src/atrac/at1/atrac1_bitalloc.cpp:118-135) derives it */
    auto spec = CalcATH(512, 44100);
    for (int b = 0; b < 52; b++) {
        float x = 999;
        size_t s = NAtrac1::TAtrac1Data::SpecsStartLong[b];
        for (size_t line = s; line < s + NAtrac1::TAtrac1Data::SpecsPerBlock[b]; line++)
            x = fmin(x, spec[line]);
        x = pow(10, 0.1 * x);
        ath52[b] = x;
    }
}

/* MDCT tap: N in {64,256,512}; scale as the encoder uses (TAtrac1MDCT ctor). out has N/2 floats. */
void ref_mdct(int n, float scale, const float* in, float* out)
{
    if (n == 512) { NMDCT::TMDCT<512> m(scale); auto& r = m(in); memcpy(out, r.data(), 256 * 4); }
    else if (n == 256) { NMDCT::TMDCT<256> m(scale); auto& r = m(in); memcpy(out, r.data(), 128 * 4); }
    else if (n == 64) { NMDCT::TMDCT<64> m(scale); auto& r = m(in); memcpy(out, r.data(), 32 * 4); }
}

/* kissfft taps */
void ref_kiss_fft(int n, int inverse, const float* in_ri, float* out_ri)
{
    kiss_fft_cfg cfg = kiss_fft_alloc(n, inverse, nullptr, nullptr);
    kiss_fft(cfg, (const kiss_fft_cpx*)in_ri, (kiss_fft_cpx*)out_ri);
    kiss_fft_free(cfg);
}
void ref_kiss_fftr(int n, const float* in, float* out_ri /* n/2+1 complex */)
{
    kiss_fftr_cfg cfg = kiss_fftr_alloc(n, 0, nullptr, nullptr);
    kiss_fftr(cfg, in, (kiss_fft_cpx*)out_ri);
    kiss_fftr_free(cfg);
}
void ref_kiss_fftri(int n, const float* in_ri /* n/2+1 complex */, float* out)
{
    kiss_fftr_cfg cfg = kiss_fftr_alloc(n, 1, nullptr, nullptr);
    kiss_fftri(cfg, (const kiss_fft_cpx*)in_ri, out);
    kiss_fftr_free(cfg);
}

/* libm taps: the exact functions the reference resolves to on this box */
/* ATRAC1 decoder (src/atrac1denc.cpp:139-177): TAtrac1Decoder's lambda driven frame by frame over in-memory sound
 * units [F][C][212]; pcm receives what the lambda writes, [F*512][C] interleaved. */
namespace {
class TMemInput : public ICompressedInput {
    const unsigned char* Units;
    long Count, Pos = 0;
    size_t Ch;
public:
    TMemInput(const unsigned char* u, long count, size_t ch) : Units(u), Count(count), Ch(ch) {}
    std::unique_ptr<TFrame> ReadFrame() override {
        std::unique_ptr<TFrame> f(new TFrame(212));
        if (Pos < Count) memcpy(f->Get(), Units + 212 * Pos, 212); else memset(f->Get(), 0, 212);
        Pos++;
        return f;
    }
    uint64_t GetLengthInSamples() const override { return (uint64_t)(Count / (long)Ch) * 512; }
    std::string GetName() const override { return "mem"; }
    size_t GetChannelNum() const override { return Ch; }
};
} // namespace

long ref_at1_decode(int channels, const unsigned char* units, long n_frames, float* pcm)
{
    try {
        TCompressedInputPtr in(new TMemInput(units, n_frames * channels, channels));
        /* TAtrac1Decoder leaves its overlap buffers PcmBufLow/Mid/Hi uninitialised (src/atrac1denc.h:112-114); the
         * CLI allocates it with `new` on a fresh heap, i.e. in zero pages.  Give it zeroed storage here too, so that
         * the oracle does not depend on what happens to lie on the stack. */
        void* mem = calloc(1, sizeof(TAtrac1Decoder));
        TAtrac1Decoder* dec = new (mem) TAtrac1Decoder(std::move(in));
        {
            auto lambda = dec->GetLambda();
            TPCMEngine::ProcessMeta meta = {(uint16_t)channels};
            for (long f = 0; f < n_frames; f++) lambda(pcm + (size_t)f * 512 * channels, meta);
        }
        dec->~TAtrac1Decoder();
        free(mem);
        return n_frames;
    } catch (...) {
        return -1;
    }
}

float ref_log10f(float x) { return log10f(x); }
float ref_log2f(float x) { return log2f(x); }
double ref_log(double x) { return log(x); }
double ref_exp(double x) { return exp(x); }

} // extern "C"
