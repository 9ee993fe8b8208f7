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

struct __align__(8) cpx { float r, i; };   // 8-byte aligned: one 64-bit load / store per element

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
ATDE_D f32x2 sub2(f32x2 a, f32x2 b, f32x2 /*mone*/) { f32x2 r; r.x = fsub(a.x, b.x); r.y = fsub(a.y, b.y); return r; }
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
// a - b as b * (-1) + a with MONE = (-1.0f, -1.0f), opaque like ONE: the product is exact, one rounding
ATDE_D f32x2 sub2(f32x2 a, f32x2 b, f32x2 mone)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(b)), "l"(pack2(mone)), "l"(pack2(a)));
    return unpack2(r);
}
#endif

// ---- bulk asynchronous copies (1-D TMA, `cp.async.bulk`) with mbarrier completion ----
// Global -> shared: ONE lane arms the barrier with the byte count and issues the copies; the copy engine moves the
// tile while the warp computes; every consumer waits on the barrier's phase parity.  Shared -> global: the writers
// make their generic-proxy stores visible to the async proxy (async_proxy_fence), one lane issues the copy and commits
// the group; bulk_store_wait_read() tells when the shared source may be overwritten.  Addresses and sizes are
// multiples of 16 bytes.  SASS: UBLKCP (bulk copy), SYNCS (mbarrier).
//
// ATDE_DYN_SMEM(name): the block's dynamic shared memory as `unsigned char* name`.
#ifdef ATDE_CPU_EMU
#define ATDE_DYN_SMEM(name) unsigned char* name = cuemu::dyn_smem
struct mbar_t { volatile unsigned phase; volatile unsigned pending; };
ATDE_D void mbar_init(mbar_t* b, int /*arrivals*/) { b->phase = 0; b->pending = 0; }
ATDE_D void mbar_expect_tx(mbar_t* b, unsigned bytes) { b->pending = bytes; }
ATDE_D void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t* b)
{
    memcpy(dst, src, bytes);
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    b->pending -= bytes;
    if (b->pending == 0) { __atomic_thread_fence(__ATOMIC_SEQ_CST); b->phase = b->phase + 1; }
}
ATDE_D void mbar_wait(mbar_t* b, unsigned parity)
{
    while ((b->phase & 1u) == parity) sched_yield();
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
}
ATDE_D void async_proxy_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
ATDE_D void bulk_s2g(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
ATDE_D void bulk_store_commit() {}
ATDE_D void bulk_store_wait_read() {}
ATDE_D void bulk_store_wait_all() {}
#else
#define ATDE_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
typedef unsigned long long mbar_t;
ATDE_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
ATDE_D void mbar_init(mbar_t* b, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(arrivals) : "memory");
}
ATDE_D void mbar_expect_tx(mbar_t* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
ATDE_D void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
ATDE_D void mbar_wait(mbar_t* b, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
ATDE_D void async_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
ATDE_D void bulk_s2g(void* dst, const void* src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
ATDE_D void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
ATDE_D void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
ATDE_D void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

} // namespace atde
