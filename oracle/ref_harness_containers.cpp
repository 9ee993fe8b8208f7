// oracle/ref_harness_containers.cpp — TEST INFRASTRUCTURE (never linked into the product).
// Drives the reference's own container writers (src/aea.cpp, src/raw.cpp, src/oma.cpp + liboma,
// src/at3.cpp, src/rm.cpp), compiled unmodified into oracle/_ref/libatde_ref.so, so that
// tests/test_containers.py can compare this repo's writers with them byte for byte.
#include "aea.h"
#include "at3.h"
#include "oma.h"
#include "raw.h"
#include "rm.h"

#include <cstdint>
#include <cstdio>
#include <exception>

// kind: 0 AEA, 1 raw, 2 OMA/ATRAC3, 3 OMA/ATRAC3plus, 4 RIFF/ATRAC3, 5 RIFF/ATRAC3plus, 6 RealMedia.
// payload: the WriteFrame payloads back to back, sizes[i] bytes each.  Returns 0, or 1 when the writer threw.
extern "C" int ref_container_write(int kind, const char* path, const char* title, int channels, uint32_t num_frames,
                                   uint32_t frame_size, int js, const uint8_t* payload, const int32_t* sizes, int n)
{
    try {
        TCompressedOutputPtr out;
        switch (kind) {
        case 0: out = CreateAeaOutput(path, title, (size_t)channels, num_frames); break;
        case 1: out = CreateRawOutput(path, (size_t)channels, frame_size); break;
        case 2: out.reset(new TOma(path, title, (size_t)channels, num_frames, OMAC_ID_ATRAC3, frame_size, js != 0)); break;
        case 3: out.reset(new TOma(path, title, (size_t)channels, num_frames, OMAC_ID_ATRAC3PLUS, frame_size, js != 0)); break;
        case 4: out = CreateAt3Output(path, (size_t)channels, num_frames, frame_size, js != 0); break;
        case 5: out = CreateAt3POutput(path, (size_t)channels, num_frames, frame_size); break;
        case 6: out = CreateRmOutput(path, title, (size_t)channels, num_frames, frame_size, js != 0); break;
        default: return 2;
        }
        size_t off = 0;
        for (int i = 0; i < n; i++) {
            out->WriteFrame(std::vector<char>(payload + off, payload + off + sizes[i]));
            off += (size_t)sizes[i];
        }
    } catch (const std::exception&) {
        return 1;
    }
    return 0;
}
