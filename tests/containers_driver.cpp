// tests/containers_driver.cpp — C entry point over this repo's container writers
// (atracdenc_b200/host/atde_containers.h) with the same signature as oracle/ref_harness_containers.cpp's
// ref_container_write, for tests/test_containers.py.
#include "../atracdenc_b200/host/atde_containers.h"

#include <cstdint>
#include <exception>

extern "C" int atde_container_write(int kind, const char* path, const char* title, int channels, uint32_t num_frames,
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
