// atde_boundary.h — the reference-side types at the drop-in boundary.
//
// When the shim is compiled INSIDE the atracdenc tree, define ATDE_USE_REFERENCE_HEADERS and the
// reference's own src/pcmengin.h and src/compressed_io.h are used unchanged.
// Stand-alone (this repo's tests, no reference sources on the box) the declarations below mirror
// that interface: same names, same signatures, same semantics —
//   TPCMEngine::ProcessMeta / EProcessResult / TProcessLambda   src/pcmengin.h:115-150
//   TPCMEngine::ApplyProcess                                    src/pcmengin.h:152-192
//   IProcessor                                                  src/pcmengin.h:195-199
//   ICompressedOutput::WriteFrame(std::vector<char>)            src/compressed_io.h:56-59
#pragma once

#ifdef ATDE_USE_REFERENCE_HEADERS
#include "pcmengin.h"
#include "compressed_io.h"
#else

#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <vector>

class TNoDataToRead : public std::exception {};
class TPCMBufferTooSmall : public std::exception {
    const char* what() const throw() override { return "PCM buffer too small"; }
};

class TPCMBuffer {
    std::vector<float> Buf_;
    size_t NumChannels;
public:
    TPCMBuffer(uint16_t bufSize, size_t numChannels) : NumChannels(numChannels) { Buf_.resize((size_t)bufSize * numChannels); }
    size_t Size() { return Buf_.size() / NumChannels; }
    float* operator[](size_t pos) { return &Buf_[pos * NumChannels]; }
    uint16_t Channels() const { return (uint16_t)NumChannels; }
    // The reference clears len*NumChannels BYTES, not floats (src/pcmengin.h:91-94); callers that
    // want the reference's end-of-stream behaviour must keep that.
    void Zero(size_t pos, size_t len) { memset(&Buf_[pos * NumChannels], 0, len * NumChannels); }
};

class IPCMReader {
public:
    virtual bool Read(TPCMBuffer& data, const uint32_t size) const = 0;
    virtual ~IPCMReader() {}
};

class TPCMEngine {
public:
    typedef std::unique_ptr<IPCMReader> TReaderPtr;
    struct ProcessMeta { const uint16_t Channels; };
    enum class EProcessResult { LOOK_AHEAD, PROCESSED };
    typedef std::function<EProcessResult(float* data, const ProcessMeta& meta)> TProcessLambda;

    TPCMEngine(uint16_t bufSize, size_t numChannels, TReaderPtr&& reader)
        : Buffer(bufSize, numChannels), Reader(std::move(reader)) {}

    uint64_t ApplyProcess(size_t step, TProcessLambda lambda)
    {
        if (step > Buffer.Size()) throw TPCMBufferTooSmall();
        bool drain = false;
        if (Reader) {
            const bool ok = Reader->Read(Buffer, (uint32_t)Buffer.Size());
            if (!ok) {
                if (ToDrain) drain = true;
                else throw TNoDataToRead();
            }
        }
        size_t lastPos = 0;
        ProcessMeta meta = {Buffer.Channels()};
        for (size_t i = 0; i + step <= Buffer.Size(); i += step) {
            auto res = lambda(Buffer[i], meta);
            if (res == EProcessResult::PROCESSED) {
                lastPos += step;
                if (drain && ToDrain--) break;
            } else {
                ToDrain++;
            }
        }
        Processed += lastPos;
        return Processed;
    }
private:
    TPCMBuffer Buffer;
    TReaderPtr Reader;
    uint64_t Processed = 0;
    uint64_t ToDrain = 0;
};

class IProcessor {
public:
    virtual typename TPCMEngine::TProcessLambda GetLambda() = 0;
    virtual ~IProcessor() {}
};

class ICompressedIO {
public:
    virtual std::string GetName() const = 0;
    virtual size_t GetChannelNum() const = 0;
    virtual ~ICompressedIO() {}
};

class ICompressedOutput : public ICompressedIO {
public:
    virtual void WriteFrame(std::vector<char> data) = 0;
};
typedef std::unique_ptr<ICompressedOutput> TCompressedOutputPtr;

#endif // ATDE_USE_REFERENCE_HEADERS
