// atde_containers.cpp — container writers behind ICompressedOutput (see atde_containers.h for the
// reference interfaces).  Built from one header-builder + one file sink instead of the reference's
// per-format packed structs; every quirk that shapes the bytes is noted where it is reproduced.
#include "atde_containers.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace {

// Growable byte image of a header with explicit-endianness appends.
class TBytes {
public:
    explicit TBytes(size_t reserve = 0) { B.reserve(reserve); }
    TBytes& U8(uint32_t v) { B.push_back((unsigned char)v); return *this; }
    TBytes& Le16(uint32_t v) { return U8(v).U8(v >> 8); }
    TBytes& Le32(uint32_t v) { return Le16(v).Le16(v >> 16); }
    TBytes& Be16(uint32_t v) { return U8(v >> 8).U8(v); }
    TBytes& Be32(uint32_t v) { return Be16(v >> 16).Be16(v); }
    TBytes& Tag(const char* s) { while (*s) U8((unsigned char)*s++); return *this; }
    TBytes& Raw(const void* p, size_t n) { const unsigned char* q = (const unsigned char*)p; B.insert(B.end(), q, q + n); return *this; }
    TBytes& Zeros(size_t n) { B.insert(B.end(), n, 0); return *this; }
    TBytes& PadTo(size_t n) { if (B.size() < n) B.resize(n, 0); return *this; }
    size_t Size() const { return B.size(); }
    const unsigned char* Data() const { return B.data(); }
private:
    std::vector<unsigned char> B;
};

// Owns the FILE; all writers go through it.
class TSink {
public:
    TSink(const std::string& filename)
        : Fp(fopen(filename.c_str(), "wb"))
    {
        if (!Fp) throw std::runtime_error("unable to open output file '" + filename + "'");
    }
    ~TSink() { if (Fp) fclose(Fp); }
    TSink(const TSink&) = delete;
    TSink& operator=(const TSink&) = delete;
    void Put(const void* p, size_t n, const char* what)
    {
        if (n && fwrite(p, 1, n, Fp) != n) throw std::runtime_error(std::string("Cannot write ") + what);
    }
    void Put(const TBytes& b, const char* what) { Put(b.Data(), b.Size(), what); }
    long Tell() { return ftell(Fp); }
    // overwrite a little/big-endian 32-bit field, then continue where the file ended
    void Patch32(long pos, uint32_t v, bool bigEndian)
    {
        TBytes b;
        if (bigEndian) b.Be32(v); else b.Le32(v);
        if (fseek(Fp, pos, SEEK_SET) == 0) fwrite(b.Data(), 1, 4, Fp);
        fseek(Fp, 0, SEEK_END);
    }
private:
    FILE* Fp;
};

class TContainerBase : public ICompressedOutput {
public:
    TContainerBase(const std::string& filename, size_t channels) : Out(filename), Channels(channels) {}
    std::string GetName() const override { return {}; }
    size_t GetChannelNum() const override { return Channels; }
protected:
    TSink Out;
    size_t Channels;
};

// ---- raw (src/raw.cpp:40-47): a non-zero frame size pads / truncates every payload ----------------
class TRawOut : public TContainerBase {
public:
    TRawOut(const std::string& filename, size_t channels, uint32_t frameSize)
        : TContainerBase(filename, channels), FrameSize(frameSize) {}
    void WriteFrame(std::vector<char> data) override
    {
        if (FrameSize) data.resize(FrameSize);
        Out.Put(data.data(), data.size(), "raw ATRAC data to file");
    }
private:
    uint32_t FrameSize;
};

// ---- AEA (src/aea.cpp:141-189) ------------------------------------------------------------------------
// 2048-byte header: magic 00 08 00 00, title (15 characters + NUL at 4..19), frame count (host order = little
// endian) at 260, channel count at 264; one all-zero 212-byte frame follows the header; the FIRST WriteFrame
// call is swallowed; every later payload is cut or zero-padded to 212 bytes.
class TAeaOut : public TContainerBase {
public:
    TAeaOut(const std::string& filename, const std::string& title, size_t channels, uint32_t numFrames)
        : TContainerBase(filename, channels), Title(title.substr(0, 15))
    {
        TBytes h(2048 + 212);
        h.U8(0x00).U8(0x08).U8(0x00).U8(0x00);
        h.Raw(Title.data(), Title.size()).PadTo(260);
        h.Le32(numFrames).U8((uint32_t)channels).PadTo(2048);
        h.Zeros(212);
        Out.Put(h, "AEA header");
    }
    std::string GetName() const override { return Title; }
    void WriteFrame(std::vector<char> data) override
    {
        if (!Started) { Started = true; return; }
        data.resize(212);
        Out.Put(data.data(), data.size(), "AEA frame");
    }
private:
    std::string Title;
    bool Started = false;
};

// ---- OMA (src/oma.cpp:28-52, liboma.c:128-236) ---------------------------------------------------------
// 96-byte "EA3" header, all zero but: [3] = 1, [5] = 96, [6] = [7] = 0xFF and the big-endian parameter word at 32:
//   ATRAC3      codec 0 << 24 | js << 17 | rate index << 13 | framesize / 8
//   ATRAC3plus  codec 1 << 24 |            rate index << 13 | (channel index + 1) << 10 | (framesize - 8) / 8
// (44100 Hz is rate index 1; mono / stereo are channel indices 0 / 1).  A field that does not fit leaves the word
// zero, the header is written regardless.  Every WriteFrame writes exactly `framesize` bytes.
class TOmaOut : public TContainerBase {
public:
    TOmaOut(const std::string& filename, size_t channels, int cid, uint32_t frameSize, bool jointStereo)
        : TContainerBase(filename, 2 /* "for ATRAC3" — what the reference answers for every OMA */), FrameSize(frameSize)
    {
        uint32_t params = 0;
        const uint32_t rateIdx = 1;
        if (cid == OMAC_ID_ATRAC3) {
            const uint32_t fs = frameSize / 8;
            if (fs <= 0x3FF) params = ((uint32_t)OMAC_ID_ATRAC3 << 24) | ((jointStereo ? 1u : 0u) << 17) | (rateIdx << 13) | fs;
        } else if (cid == OMAC_ID_ATRAC3PLUS) {
            const uint32_t fs = (frameSize - 8) / 8;
            const uint32_t chIdx = channels == 1 ? 0 : 1;
            if (fs <= 0x3FF) params = ((uint32_t)OMAC_ID_ATRAC3PLUS << 24) | (rateIdx << 13) | ((chIdx + 1) << 10) | fs;
        } else {
            throw std::runtime_error("unsupported OMA codec id");
        }
        TBytes h(96);
        h.Tag("EA3").U8(1).U8(0).U8(96).U8(0xFF).U8(0xFF).PadTo(32).Be32(params).PadTo(96);
        Out.Put(h, "OMA header");
    }
    void WriteFrame(std::vector<char> data) override
    {
        // the reference hands liboma a pointer and lets it read `framesize` bytes whatever the vector holds
        if (data.size() < FrameSize) throw std::runtime_error("OMA payload shorter than the frame size");
        Out.Put(data.data(), FrameSize, "OMA frame");
    }
private:
    uint32_t FrameSize;
};

// ---- RIFF/WAVE with ATRAC3 (0x270) or ATRAC3plus (extensible) payload (src/at3.cpp) --------------------
// All fields little endian.  The length fields are first written from the caller's frame estimate and rewritten
// from the number of frames actually delivered when the writer is destroyed (if any were, and the file is < 4 GiB).
class TRiffOut : public TContainerBase {
public:
    TRiffOut(const std::string& filename, size_t channels, uint32_t numFrames, uint32_t frameSize, bool plus, bool jointStereo)
        : TContainerBase(filename, plus ? channels : 2), FrameSize(frameSize), Plus(plus)
    {
        if (plus && frameSize > 0xFFFF) throw std::runtime_error("ATRAC3plus frame size is too large for WAV block_align");
        if (plus && channels > 0xFFFF) throw std::runtime_error("Too many channels for WAV output");
        const uint32_t samplesPerFrame = plus ? 2048 : 1024;
        const uint32_t extra = plus ? 22 : 14;
        HeaderSize = plus ? 80 : 76;
        const uint64_t fileSize = HeaderSize + (uint64_t)numFrames * frameSize;
        if (fileSize >= 0xFFFFFFFFull) throw std::runtime_error("File size is too big for this file format");
        TBytes h(HeaderSize);
        h.Tag("RIFF").Le32((uint32_t)(fileSize - 8)).Tag("WAVE");
        h.Tag("fmt ").Le32(18 + extra);
        h.Le16(plus ? 0xFFFE : 0x270).Le16((uint32_t)channels).Le32(44100);
        h.Le32(frameSize * 44100u / samplesPerFrame).Le16(frameSize).Le16(plus ? 16 : 0).Le16(extra);
        if (plus) {
            static const unsigned char guid[16] = {0xBF, 0xAA, 0x23, 0xE9, 0x58, 0xCB, 0x71, 0x44,
                                                   0xA1, 0x19, 0xFF, 0xFA, 0x01, 0xE4, 0xCE, 0x62};
            h.Le16(16).Le32(channels == 1 ? 0x4 : channels == 2 ? 0x3 : 0x0).Raw(guid, 16);
            h.Tag("fact").Le32(4);
            SamplesPos = (long)h.Size();
            h.Le32(numFrames * samplesPerFrame);
        } else {
            h.Le16(1).Le32(0x1000).Le16(jointStereo ? 1 : 0).Le16(jointStereo ? 1 : 0).Le16(1).Le16(0);
            h.Tag("fact").Le32(8);
            SamplesPos = (long)h.Size();
            h.Le32(numFrames * samplesPerFrame).Le32(samplesPerFrame);
        }
        h.Tag("data");
        DataSizePos = (long)h.Size();
        h.Le32(numFrames * frameSize);
        Out.Put(h, "WAV header to file");
    }
    ~TRiffOut() override
    {
        if (!Frames) return;
        const uint64_t fileSize = HeaderSize + Frames * (uint64_t)FrameSize;
        if (fileSize >= 0xFFFFFFFFull) return;
        Out.Patch32(4, (uint32_t)(fileSize - 8), false);
        Out.Patch32(SamplesPos, (uint32_t)Frames * (Plus ? 2048u : 1024u), false);
        Out.Patch32(DataSizePos, (uint32_t)Frames * FrameSize, false);
    }
    void WriteFrame(std::vector<char> data) override
    {
        if (Plus && data.size() != FrameSize) throw std::runtime_error("Unexpected ATRAC3plus frame size");
        Out.Put(data.data(), data.size(), "AT3 data to file");
        ++Frames;
    }
private:
    uint32_t FrameSize;
    bool Plus;
    uint32_t HeaderSize = 0;
    long SamplesPos = 0, DataSizePos = 0;
    uint64_t Frames = 0;
};

// ---- RealMedia, ra5 / "atrc" stream (src/rm.cpp) --------------------------------------------------------
// All fields big endian.  Chunks: .RMF (18) PROP (50) MDPR (168) DATA (18 + packets).  Three frames make one
// packet: a 12-byte packet header goes in front of frames 0, 3, 6, ...; the millisecond clock advances by three
// frame durations (as a double) after frames 2, 5, 8, ...  Payload words are XORed with 53 7F 61 03; bytes beyond
// the last whole word come out as zeros.  The DATA chunk size is patched when the writer is destroyed.
class TRmOut : public TContainerBase {
public:
    TRmOut(const std::string& filename, size_t channels, uint32_t numFrames, uint32_t frameSize, bool jointStereo)
        : TContainerBase(filename, 0), FrameSize(frameSize)
    {
        const double frameMs = 1000.0 * 1024.0 / 44100.0;
        FrameMs = frameMs;
        const uint32_t bitrate = (uint32_t)(8 * frameSize * 44100.0 / 1024.0);
        const uint32_t durationMs = (uint32_t)(numFrames * frameMs);
        static const char desc[] = "Audio Stream";            // 13 bytes with the terminator, as stored
        static const char mime[] = "audio/x-pn-realaudio";    // 21
        const uint32_t mdprSize = 42 + sizeof(mime) + sizeof(desc) + 92;

        TBytes h(18 + 50 + mdprSize + 18);
        h.Tag(".RMF").Be32(18).Be16(0).Be32(0).Be32(4);
        h.Tag("PROP").Be32(50).Be16(0).Be32(bitrate).Be32(bitrate).Be32(frameSize).Be32(frameSize).Be32(numFrames)
            .Be32(durationMs).Be32(0).Be32(0).Be32(18 + 50 + mdprSize).Be16(1).Be16(1 | 2);
        h.Tag("MDPR").Be32(mdprSize).Be16(0).Be16(0).Be32(bitrate).Be32(bitrate).Be32(frameSize).Be32(frameSize)
            .Be32(0).Be32(0).Be32(durationMs).U8(sizeof(desc)).Raw(desc, sizeof(desc)).U8(sizeof(mime)).Raw(mime, sizeof(mime));
        // codec data, 92 bytes
        h.Be32(92 - 4).Tag(".ra").U8(0xfd).Be16(5).Be16(0).Tag(".ra5").Be32(0x01b53530).Be16(5).Be32(0).Be16(2)
            .Be32(frameSize * 3).Be32(0x51540).Be32(bitrate / 8 * 60).Be32(bitrate / 8 * 60).Be16(1).Be16(frameSize * 3)
            .Be16(frameSize).Be16(0).U8(0).U8(0).Be16(44100).U8(0).U8(0).Be16(44100).Be16(0).Be16(16).Be16(2)
            .Tag("genr").Tag("atrc").U8(0x01).U8(0x07).U8(0).U8(0).Be32(10).Be32(4)
            .Be16((uint32_t)(1024 * channels)).Be16(0x88E).Be16(jointStereo ? 0x12 : 0x2);
        DataPos = (long)h.Size();
        h.Tag("DATA").Be32(0xffffffffu).Be16(0).Be32(numFrames).Be32(0);
        Out.Put(h, "RM headers");
    }
    ~TRmOut() override
    {
        const long long size = (long long)Out.Tell() - DataPos;
        if (size <= 0xffffffffLL) Out.Patch32(DataPos + 4, (uint32_t)size, true);
        else fprintf(stderr, "Too many data for RM container. Encoded data is writen, but format is incorrect.");
    }
    void WriteFrame(std::vector<char> data) override
    {
        // the reference scrambles into a function-static buffer sized by the first frame it ever sees
        if (data.size() != FrameSize) throw std::runtime_error("RealMedia frames must all have the container frame size");
        static const unsigned char key[4] = {0x53, 0x7F, 0x61, 0x03};
        std::vector<char> x(data.size(), 0);
        for (size_t i = 0; i + 4 <= data.size(); i += 4)
            for (int k = 0; k < 4; k++) x[i + k] = (char)((unsigned char)data[i + k] ^ key[k]);
        const uint32_t phase = FrameNum % 3;
        if (phase == 0) {
            TBytes p(12);
            p.Be16(0).Be16((uint32_t)(3 * data.size() + 12)).Be16(0).Be32((uint32_t)Clock).U8(0).U8(0x02);
            Out.Put(p, "packet header");
        } else if (phase == 2) {
            Clock += FrameMs * 3.0;
        }
        Out.Put(x.data(), x.size(), "codec data");
        FrameNum++;
    }
private:
    uint32_t FrameSize;
    double FrameMs = 0.0, Clock = 0.0;
    uint32_t FrameNum = 0;
    long DataPos = 0;
};

} // namespace

TCompressedOutputPtr CreateAeaOutput(const std::string& filename, const std::string& title, size_t numChannels, uint32_t numFrames)
{
    return TCompressedOutputPtr(new TAeaOut(filename, title, numChannels, numFrames));
}
TCompressedOutputPtr CreateRawOutput(const std::string& filename, size_t numChannels, uint32_t frameSize)
{
    return TCompressedOutputPtr(new TRawOut(filename, numChannels, frameSize));
}
TCompressedOutputPtr CreateOmaOutput(const std::string& filename, const std::string&, size_t numChannel, uint32_t, int cid,
                                     uint32_t framesize, bool jointStereo)
{
    return TCompressedOutputPtr(new TOmaOut(filename, numChannel, cid, framesize, jointStereo));
}
TCompressedOutputPtr CreateAt3Output(const std::string& filename, size_t numChannel, uint32_t numFrames, uint32_t framesize, bool jointStereo)
{
    return TCompressedOutputPtr(new TRiffOut(filename, numChannel, numFrames, framesize, false, jointStereo));
}
TCompressedOutputPtr CreateAt3POutput(const std::string& filename, size_t numChannel, uint32_t numFrames, uint32_t framesize)
{
    return TCompressedOutputPtr(new TRiffOut(filename, numChannel, numFrames, framesize, true, false));
}
TCompressedOutputPtr CreateRmOutput(const std::string& filename, const std::string&, size_t numChannel, uint32_t numFrames,
                                    uint32_t framesize, bool jointStereo)
{
    return TCompressedOutputPtr(new TRmOut(filename, numChannel, numFrames, framesize, jointStereo));
}

TOma::TOma(const std::string& filename, const std::string& title, size_t numChannel, uint32_t numFrames, int cid,
           uint32_t framesize, bool jointStereo)
    : Impl(CreateOmaOutput(filename, title, numChannel, numFrames, cid, framesize, jointStereo))
{}
TOma::~TOma() {}
void TOma::WriteFrame(std::vector<char> data) { Impl->WriteFrame(std::move(data)); }
std::string TOma::GetName() const { return {}; }      // the reference aborts here; nobody calls it
size_t TOma::GetChannelNum() const { return 2; }
