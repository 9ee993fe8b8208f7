// atde_containers.h — the output side of the path: the containers the
// reference's main.cpp hands to the frame processors as ICompressedOutput, so that a program built on
// this repo's TAtrac1Encoder / TAtrac3Encoder / TAt3PEnc writes FILES byte-identical to atracdenc's.
// Same factory names and argument meaning as the reference:
//   CreateAeaOutput(filename, title, numChannels, numFrames)                       src/aea.h:45,   src/aea.cpp:141-199
//   CreateRawOutput(filename, numChannels, frameSize = 0)                          src/raw.h:27,   src/raw.cpp:28-69
//   TOma(filename, title, numChannel, numFrames, cid, framesize, jointStereo)      src/oma.h:24-33, src/oma.cpp:28-61,
//                                                                                  src/lib/liboma/src/liboma.c:190-236
//   CreateAt3Output(filename, numChannel, numFrames, framesize, jointStereo)       src/at3.h:24,   src/at3.cpp:158-262
//   CreateAt3POutput(filename, numChannel, numFrames, framesize)                   src/at3.h:27,   src/at3.cpp:264-376
//   CreateRmOutput(filename, title, numChannel, numFrames, framesize, jointStereo) src/rm.h:23,    src/rm.cpp:40-283
// Host-only C++ (no CUDA); checked byte for byte against the reference's own writers by
// tests/test_containers.py.  Where the reference's behaviour is undefined (an OMA payload shorter than the frame
// size, RealMedia frames of different sizes) these writers throw instead.
#pragma once
#include "atde_boundary.h"

#include <cstdint>
#include <string>

// (global namespace, like the reference's writers)

// codec ids of the OMA header (src/lib/liboma/include/oma.h:44-50)
enum { OMAC_ID_ATRAC3 = 0, OMAC_ID_ATRAC3PLUS = 1 };

TCompressedOutputPtr CreateAeaOutput(const std::string& filename, const std::string& title, size_t numChannels,
                                     uint32_t numFrames);
TCompressedOutputPtr CreateRawOutput(const std::string& filename, size_t numChannels, uint32_t frameSize = 0);
TCompressedOutputPtr CreateOmaOutput(const std::string& filename, const std::string& title, size_t numChannel,
                                     uint32_t numFrames, int cid, uint32_t framesize, bool jointStereo);
TCompressedOutputPtr CreateAt3Output(const std::string& filename, size_t numChannel, uint32_t numFrames,
                                     uint32_t framesize, bool jointStereo);
TCompressedOutputPtr CreateAt3POutput(const std::string& filename, size_t numChannel, uint32_t numFrames,
                                      uint32_t framesize);
TCompressedOutputPtr CreateRmOutput(const std::string& filename, const std::string& title, size_t numChannel,
                                    uint32_t numFrames, uint32_t framesize, bool jointStereo);

// main.cpp constructs the OMA writer directly (`new TOma(...)`, src/main.cpp:402, :456)
class TOma : public ICompressedOutput {
public:
    TOma(const std::string& filename, const std::string& title, size_t numChannel, uint32_t numFrames, int cid,
         uint32_t framesize, bool jointStereo);
    ~TOma() override;
    void WriteFrame(std::vector<char> data) override;
    std::string GetName() const override;
    size_t GetChannelNum() const override;
private:
    TCompressedOutputPtr Impl;
};

