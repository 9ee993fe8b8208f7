// tests/cpuemu/cuda_emu.h — TEST TOOLING ONLY.
//
// A deliberately small "CUDA on pthreads" shim: one OS thread per CUDA thread of a block, blocks
// executed one after another, __syncthreads()/warp collectives mapped onto pthread barriers.
// It lets the repo's CPU-only CI (no GPU in the authoring container) run the *actual* kernel
// sources of atracdenc_b200/csrc against the oracle at tiny sizes.  It is not a fallback: it is
// only ever compiled into tests/cpuemu/_build/libatde_emu.so, which nothing in the product loads.
//
// Supported subset = what the kernels use.  Warp collectives must be called by all 32 lanes.
#pragma once
#include <sched.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <vector>
#include <thread>
#include <pthread.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define ATDE_HD inline
#define ATDE_D inline
#define ATDE_NOINLINE inline

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace cuemu {
struct BlockCtx {
    int nthreads, nwarps;
    pthread_barrier_t bar;
    pthread_barrier_t named[8];     // bar.sync id, count (id 1..7), initialised on first use
    int named_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    pthread_mutex_t named_mu = PTHREAD_MUTEX_INITIALIZER;
    std::vector<pthread_barrier_t> wbar;
    std::vector<uint64_t> xchg;   // [nwarps][32]
    explicit BlockCtx(int nt) : nthreads(nt), nwarps((nt + 31) / 32), wbar(nwarps), xchg((size_t)nwarps * 32)
    {
        pthread_barrier_init(&bar, nullptr, nt);
        for (int w = 0; w < nwarps; w++) {
            int cnt = (w == nwarps - 1) ? nt - 32 * w : 32;
            pthread_barrier_init(&wbar[w], nullptr, cnt);
        }
    }
    ~BlockCtx()
    {
        pthread_barrier_destroy(&bar);
        for (int i = 0; i < 8; i++) if (named_count[i]) pthread_barrier_destroy(&named[i]);
        for (auto& b : wbar) pthread_barrier_destroy(&b);
    }
};
extern thread_local BlockCtx* ctx;
extern thread_local int lin_tid;
extern unsigned char* dyn_smem;
inline void warp_barrier() { pthread_barrier_wait(&ctx->wbar[lin_tid >> 5]); }
inline uint64_t* warp_slots() { return &ctx->xchg[(size_t)(lin_tid >> 5) * 32]; }
template <class T> inline T xchg(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload too wide");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    uint64_t* s = warp_slots();
    s[lin_tid & 31] = bits;
    warp_barrier();
    uint64_t got = s[src_lane & 31];
    warp_barrier();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
} // namespace cuemu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static const int warpSize = 32;

inline void __syncthreads() { pthread_barrier_wait(&cuemu::ctx->bar); }
// named barrier over `count` threads of the block (PTX bar.sync id, count)
inline void atde_named_barrier(int id, int count)
{
    cuemu::BlockCtx* c = cuemu::ctx;
    pthread_mutex_lock(&c->named_mu);
    if (!c->named_count[id]) { pthread_barrier_init(&c->named[id], nullptr, count); c->named_count[id] = count; }
    pthread_mutex_unlock(&c->named_mu);
    pthread_barrier_wait(&c->named[id]);
}
inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::warp_barrier(); }
inline void __threadfence_block() {}
inline void __threadfence() {}

template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cuemu::xchg(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cuemu::xchg(v, (cuemu::lin_tid & 31) ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32)
{
    int lane = cuemu::lin_tid & 31;
    int src = lane + (int)d;
    T r = cuemu::xchg(v, src > 31 ? lane : src);
    return r;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32)
{
    int lane = cuemu::lin_tid & 31;
    int src = lane - (int)d;
    return cuemu::xchg(v, src < 0 ? lane : src);
}
inline unsigned __ballot_sync(unsigned, int pred)
{
    uint64_t* s = cuemu::warp_slots();
    s[cuemu::lin_tid & 31] = pred ? 1 : 0;
    cuemu::warp_barrier();
    unsigned m = 0;
    int w = cuemu::lin_tid >> 5;
    int cnt = (w == cuemu::ctx->nwarps - 1) ? cuemu::ctx->nthreads - 32 * w : 32;
    for (int i = 0; i < cnt; i++) m |= (unsigned)(s[i] & 1) << i;
    cuemu::warp_barrier();
    return m;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    uint64_t* s = cuemu::warp_slots();
    s[cuemu::lin_tid & 31] = v;
    cuemu::warp_barrier();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r += (unsigned)s[i];
    cuemu::warp_barrier();
    return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    uint64_t* s = cuemu::warp_slots();
    s[cuemu::lin_tid & 31] = v;
    cuemu::warp_barrier();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r = (unsigned)s[i] > r ? (unsigned)s[i] : r;
    cuemu::warp_barrier();
    return r;
}
inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned)v); }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// CUDA's integer min/max overloads
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }

// ---- arithmetic intrinsics (host FP environment is round-to-nearest, no FMA contraction: the
// emu build uses -ffp-contract=off and no -march) ----
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __log2f(float a) { return log2f(a); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline int __float2int_rn(float x) { return (int)lrintf(x); }
inline float __double2float_rn(double x) { return (float)x; }
inline int __float2int_rz(float x) { return (int)x; }
inline int __double2int_rz(double x) { return (int)x; }
inline float __int2float_rn(int x) { return (float)x; }
inline float __uint2float_rn(unsigned x) { return (float)x; }
inline double __int2double_rn(int x) { return (double)x; }
inline double __ll2double_rn(long long x) { return (double)x; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned __brev(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 7;
        r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i);
    }
    return r;
}
template <class T> inline T __ldg(const T* p) { return *p; }

inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline void __nanosleep(unsigned) { sched_yield(); }
inline int atomicMax(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- runtime API subset ----
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct { double t; }* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; size_t totalGlobalMem; int major, minor; };
inline const char* cudaGetErrorString(cudaError_t) { return "cuda_emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 2; p->major = 10; p->minor = 0;
    snprintf(p->name, sizeof(p->name), "cpu-emu");
    return 0;
}
inline cudaError_t cudaDeviceSynchronize() { return 0; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n)
{
    *p = (T*)malloc(n ? n : 1);
    if (*p) memset((void*)*p, 0xCD, n);
    return *p ? 0 : 2;
}
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { *p = (T*)malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < h; r++) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return 0;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return 0; }
template <class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice)
{
    memcpy((char*)&sym + off, src, n);
    return 0;
}
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(sizeof(**e)); return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

// ---- launch ----
namespace cuemu {
template <class K, class... A> void launch(K kern, dim3 grid, dim3 block, size_t smem, A... args)
{
    const int nt = (int)(block.x * block.y * block.z);
    BlockCtx bctx(nt);
    std::vector<unsigned char> dyn(smem + 64);
    dyn_smem = dyn.data();
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; t++) {
        th.emplace_back([&, t]() {
            ctx = &bctx;
            lin_tid = t;
            blockDim = block;
            gridDim = grid;
            threadIdx.x = t % block.x;
            threadIdx.y = (t / block.x) % block.y;
            threadIdx.z = t / (block.x * block.y);
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                        kern(args...);
                        pthread_barrier_wait(&bctx.bar);
                    }
        });
    }
    for (auto& x : th) x.join();
}
} // namespace cuemu
#define ATDE_LAUNCH(kern, grid, block, smem, stream, ...) cuemu::launch(kern, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
