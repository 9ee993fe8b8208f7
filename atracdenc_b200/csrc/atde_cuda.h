// atde_cuda.h — single include for every kernel translation unit.
//
// Product build: nvcc, sm_100a, -fmad=false (the parity contract forbids FMA contraction: the
// reference is built for baseline x86-64).
//
// ATDE_CPU_EMU build: tests/cpuemu compiles the SAME kernel sources with g++ against a small
// thread-per-CUDA-thread shim so the kernels can be checked against the oracle on a box without
// a GPU.  That build is test tooling only; it is never part of libatde_b200.so and the product
// library has no CPU path.
#pragma once

#ifdef ATDE_CPU_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#define ATDE_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ATDE_HD __host__ __device__ __forceinline__
#define ATDE_D __device__ __forceinline__
#define ATDE_NOINLINE __device__ __noinline__
// named barrier over `count` threads (a multiple of 32) of the block
__device__ __forceinline__ void atde_named_barrier(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif

// Block-wide "parallel for": every phase between two __syncthreads() is a grid-stride loop over
// independent work items.
#define ATDE_PAR_FOR(i, n) for (int i = (int)threadIdx.x; i < (int)(n); i += (int)blockDim.x)

namespace atde {

struct cpx { float r, i; };

// Un-fused IEEE fp32 helpers.  With -fmad=false plain operators would do, but spelling the
// rounding out keeps the parity contract visible and survives a stray build flag.
ATDE_D float fmul(float a, float b) { return __fmul_rn(a, b); }
ATDE_D float fadd(float a, float b) { return __fadd_rn(a, b); }
ATDE_D float fsub(float a, float b) { return __fsub_rn(a, b); }

// C_MUL of kissfft (_kiss_fft_guts.h:87-89): 4 products, 1 sub, 1 add, each rounded.
ATDE_D cpx cmul(cpx a, cpx b)
{
    cpx m;
    m.r = fsub(fmul(a.r, b.r), fmul(a.i, b.i));
    m.i = fadd(fmul(a.r, b.i), fmul(a.i, b.r));
    return m;
}

// ---- packed fp32 pairs (Blackwell FMUL2 / FFMA2): two IEEE-rn operations per issue slot ----
// mul2 is a plain packed multiply.  add2 must stay an un-fused add: ptxas contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with --fmad=false, which would change the
// rounding.  So add2 is spelled fma(a, ONE, b) with ONE = (1.0f, 1.0f) taken from a kernel parameter
// (opaque to the compiler): a*1 + b rounds once, exactly like a + b, and nothing can be folded into it.
struct f32x2 { float x, y; };
#ifdef ATDE_CPU_EMU
ATDE_D f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; r.x = fmul(a.x, b.x); r.y = fmul(a.y, b.y); return r; }
ATDE_D f32x2 add2(f32x2 a, f32x2 b, f32x2 /*one*/) { f32x2 r; r.x = fadd(a.x, b.x); r.y = fadd(a.y, b.y); return r; }
#else
ATDE_D unsigned long long pack2(f32x2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
ATDE_D f32x2 unpack2(unsigned long long v)
{
    f32x2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
ATDE_D f32x2 mul2(f32x2 a, f32x2 b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(r);
}
ATDE_D f32x2 add2(f32x2 a, f32x2 b, f32x2 one)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(one)), "l"(pack2(b)));
    return unpack2(r);
}
#endif

} // namespace atde
