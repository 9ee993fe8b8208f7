// tests/cpuemu/cuda_emu.cpp — TEST TOOLING ONLY: storage for the emulation shim's thread-locals.
#include "cuda_emu.h"
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
namespace cuemu {
thread_local BlockCtx* ctx = nullptr;
thread_local int lin_tid = 0;
unsigned char* dyn_smem = nullptr;
}
