// oracle/pcm_io_testwav.cpp — TEST INFRASTRUCTURE: a PCM I/O backend for the reference CLI on a box without libsndfile.
//
// The reference's src/wav.cpp is backend-agnostic: it calls CreatePCMIOReadImpl / CreatePCMIOWriteImpl
// (src/wav.cpp:31-33) and talks to an IPCMProviderImpl (src/wav.h:54-63).  The stock backend is
// src/pcm_io_sndfile.cpp (libsndfile, absent in this image).  This file is a second backend for canonical 16-bit
// PCM RIFF/WAVE files so that the UNMODIFIED src/main.cpp + src/wav.cpp + src/pcmengin.h link into a working
// `atracdenc` for tests/test_reference_dropin.py.  Reading hands out what libsndfile's readf(float*) does for a
// PCM_16 file: sample * (1.0 / 0x8000).
#include "wav.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace {

uint32_t Le32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t Le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

class TTestWav : public IPCMProviderImpl {
public:
    explicit TTestWav(const std::string& path)
    {
        F = fopen(path.c_str(), "rb");
        if (!F) throw std::runtime_error("unable to open input file '" + path + "'");
        unsigned char h[12];
        if (fread(h, 1, 12, F) != 12 || memcmp(h, "RIFF", 4) || memcmp(h + 8, "WAVE", 4))
            throw std::runtime_error("not a RIFF/WAVE file: " + path);
        bool haveFmt = false;
        for (;;) {
            unsigned char ch[8];
            if (fread(ch, 1, 8, F) != 8) throw std::runtime_error("no data chunk: " + path);
            const uint32_t sz = Le32(ch + 4);
            if (!memcmp(ch, "fmt ", 4)) {
                std::vector<unsigned char> f(sz);
                if (fread(f.data(), 1, sz, F) != sz || sz < 16) throw std::runtime_error("bad fmt chunk");
                if (Le16(&f[0]) != 1 || Le16(&f[14]) != 16) throw std::runtime_error("only 16-bit PCM is supported by the test backend");
                Channels = Le16(&f[2]);
                Rate = Le32(&f[4]);
                haveFmt = true;
            } else if (!memcmp(ch, "data", 4)) {
                if (!haveFmt) throw std::runtime_error("data before fmt");
                Total = sz / (2 * Channels);
                break;
            } else {
                fseek(F, (long)(sz + (sz & 1)), SEEK_CUR);
            }
        }
    }
    TTestWav(const std::string& path, int channels, int sampleRate)
        : Channels((size_t)channels), Rate((size_t)sampleRate), Writing(true)
    {
        F = fopen(path.c_str(), "wb");
        if (!F) throw std::runtime_error("unable to open output file '" + path + "'");
        unsigned char h[44] = {0};
        fwrite(h, 1, 44, F);
    }
    ~TTestWav() override
    {
        if (F && Writing) {
            const uint32_t bytes = (uint32_t)(Total * 2 * Channels);
            unsigned char h[44];
            auto p32 = [&](int o, uint32_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; h[o + 2] = (v >> 16) & 255; h[o + 3] = v >> 24; };
            auto p16 = [&](int o, uint16_t v) { h[o] = v & 255; h[o + 1] = v >> 8; };
            memcpy(h, "RIFF", 4); p32(4, 36 + bytes); memcpy(h + 8, "WAVEfmt ", 8); p32(16, 16); p16(20, 1);
            p16(22, (uint16_t)Channels); p32(24, (uint32_t)Rate); p32(28, (uint32_t)(Rate * 2 * Channels));
            p16(32, (uint16_t)(2 * Channels)); p16(34, 16); memcpy(h + 36, "data", 4); p32(40, bytes);
            fseek(F, 0, SEEK_SET);
            fwrite(h, 1, 44, F);
        }
        if (F) fclose(F);
    }
    size_t GetChannelsNum() const override { return Channels; }
    size_t GetSampleRate() const override { return Rate; }
    size_t GetTotalSamples() const override { return Total; }
    size_t Read(TPCMBuffer& buf, size_t sz) override
    {
        const size_t want = std::min(sz, Total - Pos);
        Tmp.resize(want * Channels);
        const size_t got = want ? fread(Tmp.data(), 2 * Channels, want, F) : 0;
        float* dst = buf[0];
        for (size_t i = 0; i < got * Channels; i++) dst[i] = (float)Tmp[i] * (float)(1.0 / 0x8000);
        Pos += got;
        return got;
    }
    size_t Write(const TPCMBuffer& buf, size_t sz) override
    {
        const float* src = const_cast<TPCMBuffer&>(buf)[0];
        Tmp.resize(sz * Channels);
        for (size_t i = 0; i < sz * Channels; i++) {
            float v = src[i] * 32767.0f;
            v = v > 32767.0f ? 32767.0f : (v < -32768.0f ? -32768.0f : v);
            Tmp[i] = (int16_t)lrintf(v);
        }
        const size_t put = fwrite(Tmp.data(), 2 * Channels, sz, F);
        Total += put;
        return put;
    }

private:
    FILE* F = nullptr;
    size_t Channels = 0, Rate = 0, Total = 0, Pos = 0;
    bool Writing = false;
    std::vector<int16_t> Tmp;
};

} // namespace

IPCMProviderImpl* CreatePCMIOReadImpl(const std::string& path) { return new TTestWav(path); }
IPCMProviderImpl* CreatePCMIOWriteImpl(const std::string& path, int channels, int sampleRate) { return new TTestWav(path, channels, sampleRate); }
