// tests/host_shim_driver.cpp — drives the C++ host shim exactly like src/main.cpp:697-716 drives
// the reference's processors: TPCMEngine(4096) + memory reader + GetLambda() loop, capturing the
// WriteFrame payloads to a file.
// usage: driver <pcm.f32> <channels> <total_samples> <out.bin> <batch_frames> [codec=1|3|4] [bitrate_kbit] [container]
// With a container name (aea, raw, oma, riff, rm) the output goes through this repo's container writers, created
// with the arguments main.cpp passes (src/main.cpp:312-326, :380-410, :438-462): the result is a whole file.
#include "../atracdenc_b200/host/atde_encoders.h"
#include "../atracdenc_b200/host/atde_containers.h"
#include <string>
#include <cstdio>
#include <cstdlib>

using namespace NAtracDEnc;

class TMemReader : public IPCMReader {
    const std::vector<float>& Data; size_t Ch; mutable size_t Pos = 0;
public:
    TMemReader(const std::vector<float>& d, size_t ch) : Data(d), Ch(ch) {}
    bool Read(TPCMBuffer& buf, const uint32_t size) const override {   // TWav::GetPCMReader, src/wav.cpp:46-61
        size_t total = Data.size() / Ch, left = total - Pos, n = left < size ? left : size;
        if (!n) return false;
        memcpy(buf[0], &Data[Pos * Ch], n * Ch * sizeof(float));
        Pos += n;
        if (n != size) buf.Zero(n, size - n);
        return true;
    }
};

class TFileSink : public ICompressedOutput {
    FILE* F; size_t Ch;
public:
    TFileSink(const char* path, size_t ch) : F(fopen(path, "wb")), Ch(ch) {}
    ~TFileSink() override { if (F) fclose(F); }
    void WriteFrame(std::vector<char> data) override {
        int32_t n = (int32_t)data.size();
        fwrite(&n, 4, 1, F);
        fwrite(data.data(), 1, data.size(), F);
    }
    std::string GetName() const override { return "file"; }
    size_t GetChannelNum() const override { return Ch; }
};

int main(int argc, char** argv)
{
    if (argc < 6) return 2;
    const size_t ch = atoi(argv[2]);
    const uint64_t total = strtoull(argv[3], nullptr, 10);
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> pcm(sz / 4);
    if (fread(pcm.data(), 4, pcm.size(), f) != pcm.size()) return 4;
    fclose(f);
    try {
        const int codec = argc > 6 ? atoi(argv[6]) : 1;
        const uint32_t kbit = argc > 7 ? (uint32_t)atoi(argv[7]) : 0;
        const std::string cont = argc > 8 ? argv[8] : "";
        TCompressedOutputPtr sink;
        if (cont.empty()) {
            sink.reset(new TFileSink(argv[4], ch));
        } else if (codec == 1) {
            const uint64_t numFrames = ch * total / 512;
            sink = cont == "raw" ? CreateRawOutput(argv[4], ch, 212) : CreateAeaOutput(argv[4], "test", ch, (uint32_t)numFrames);
        } else if (codec == 3) {
            const uint64_t numFrames = total / 1024;
            const NAtrac3::TContainerParams* cp = NAtrac3::GetContainerParamsForBitrate(kbit * 1024);
            if (cont == "riff") sink = CreateAt3Output(argv[4], 2, (uint32_t)numFrames, cp->FrameSz, cp->Js);
            else if (cont == "raw") sink = CreateRawOutput(argv[4], ch);
            else if (cont == "rm") sink = CreateRmOutput(argv[4], "test", ch, (uint32_t)numFrames, cp->FrameSz, cp->Js);
            else sink.reset(new TOma(argv[4], "test", ch, (int32_t)numFrames, OMAC_ID_ATRAC3, cp->FrameSz, cp->Js));
        } else {
            const uint64_t numFrames = total / 2048;
            if (cont == "riff") sink = CreateAt3POutput(argv[4], ch, (uint32_t)numFrames, 2048);
            else if (cont == "raw") sink = CreateRawOutput(argv[4], ch);
            else sink.reset(new TOma(argv[4], "test", ch, (int32_t)numFrames, OMAC_ID_ATRAC3PLUS, 2048, false));
        }
        std::unique_ptr<TBatchedEncoderBase> proc;
        size_t step = 512;
        if (codec == 1) {
            proc.reset(new TAtrac1Encoder(std::move(sink),
                NAtrac1::TAtrac1EncodeSettings(0, NAtrac1::TAtrac1EncodeSettings::EWindowMode::EWM_AUTO, 0)));
        } else if (codec == 4) {
            // src/main.cpp:478-482: TAt3PEnc(std::move(out), channels, settings)
            proc.reset(new TAt3PEnc(std::move(sink), (int)ch, TAt3PEnc::TSettings()));
            step = 2048;
        } else {
            // src/main.cpp:671: TAtrac3EncoderSettings(bitrate * 1024, noGainControl, noTonalComponents, channels, bfuIdxConst)
            proc.reset(new TAtrac3Encoder(std::move(sink),
                NAtrac3::TAtrac3EncoderSettings(kbit * 1024, false, false, (uint8_t)ch, 0)));
            step = 1024;
        }
        proc->SetBatchFrames(atoi(argv[5]));
        TPCMEngine eng(4096, ch, TPCMEngine::TReaderPtr(new TMemReader(pcm, ch)));
        auto lambda = proc->GetLambda();
        try {
            while (total > eng.ApplyProcess(step, lambda)) {}
        } catch (const TNoDataToRead&) {}
        // processor destroyed here -> flushes the staged tail, as in main.cpp's scope exit
    } catch (const std::exception& e) {
        fprintf(stderr, "driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
