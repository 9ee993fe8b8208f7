// at1_decode.cu — ATRAC1 decoder on sm_100a (the step after the encode path).
//
// Replaces, for thousands of frames at once (reference: dcherednik/atracdenc):
//   TAtrac1Decoder::GetLambda          src/atrac1denc.cpp:139-177   (frame loop, clipping, interleaving)
//   TAtrac1Data::TBlockSizeMod::Parse  src/atrac/at1/atrac1.cpp:37-53
//   TAtrac1Dequantiser::Dequant        src/atrac/at1/atrac1_dequantiser.cpp:31-72
//   TAtrac1MDCT::IMdct                 src/atrac1denc.cpp:103-137  (+ vector_fmul_window :51-68)
//   NMDCT::TMIDCT<N>                   src/lib/mdct/mdct.h:107-180, kiss_fft (forward plan, N/4 points)
//   Atrac1SynthesisFilterBank          src/atrac/at1/atrac1_qmf.h:46-66, TQmf::Synthesis src/qmf/qmf.h:66-89
//
// Frames of a stream are decoded in parallel.  The decoder's cross-frame state — the 16-sample IMDCT overlap tail per
// band, the two 46-sample synthesis-QMF histories, the 39-sample delay line of the high band — is a finite function
// of the PREVIOUS frame's spectrum only (the positions it needs lie past that frame's own overlap zone), so a tile
// re-derives it from one halo frame: the frame before the tile, or the last sound unit carried from the previous
// batch, or nothing (a fresh stream: the reference's zero-initialised buffers).
// All arithmetic is un-fused IEEE fp32 in the reference's operation order (bit-exact contract).
//
// Block-size codes the reference ENCODER never writes (two of four / two or four of eight short blocks) make the
// reference decoder leave part of its band buffer stale from older frames; such frames are refused (status flag).
#include "at1_kernels.cuh"
#include "kissfft_dev.cuh"

namespace atde {
namespace at1 {

// BFU geometry, atrac1.h:89-106 (same tables as the packer in at1_kernels.cu)
__device__ const unsigned char kDecSpecsPerBlock[kMaxBfus] = {
    8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 7, 7, 7, 7, 9, 9, 9, 9, 10, 10, 10, 10,
    12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20, 20, 20, 20, 20, 20};
__device__ const unsigned short kDecStartLong[kMaxBfus] = {
    0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116, 122,
    128, 134, 140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226, 236, 246,
    256, 268, 280, 292, 304, 316, 328, 340, 352, 372, 392, 412, 432, 452, 472, 492};
__device__ const unsigned short kDecStartShort[kMaxBfus] = {
    0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90, 122,
    128, 160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182, 214, 246,
    256, 288, 320, 352, 384, 416, 448, 480, 268, 300, 332, 364, 396, 428, 460, 492};

constexpr int kDecTile = 4;                         // frames per block (+ one halo frame)
constexpr int kDecFrames = kDecTile + 1;
constexpr int kDecThreads = 128;

// MSB-first bit field of a sound unit (TBitStream::Read, bitstream.cpp:63-92); n <= 16
ATDE_D unsigned get_bits(const unsigned char* u, int pos, int n)
{
    const int byte = pos >> 3, off = pos & 7;
    const unsigned w = ((unsigned)u[byte] << 16) | ((unsigned)u[byte + 1] << 8) | (unsigned)u[byte + 2];
    return (w >> (24 - off - n)) & ((1u << n) - 1u);
}

__global__ void __launch_bounds__(kDecThreads) at1_decode_kernel(DecodeParams p)
{
    __shared__ __align__(16) float sp[kDecFrames][512];          // spectra; later the IMDCT outputs (invBuf)
    __shared__ __align__(16) cpx fft[kDecFrames][256];
    __shared__ float lowS[kDecFrames * 128], midS[kDecFrames * 128], hiS[kDecFrames * 256];
    __shared__ float midlow[kDecFrames * 256];
    __shared__ unsigned char ub[kDecFrames][216];
    __shared__ unsigned char smode[kDecFrames];                  // bit b: band b uses short blocks; 0x80: frame absent
    __shared__ float W[32];
    __shared__ float qw[48];

    const DecTables* __restrict__ T = p.tab;
    const int s = blockIdx.y, c = blockIdx.z;
    const int t0 = blockIdx.x * kDecTile;
    const int C = p.C, F = p.F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool started = p.started && p.started[s];

    ATDE_PAR_FOR(i, 32) W[i] = T->sine_window[i];
    ATDE_PAR_FOR(i, 48) qw[i] = T->qmf_window[i];
    // ---- sound units of the halo frame and the tile ----
    for (int k = 0; k < kDecFrames; k++) {
        const int ft = t0 - 1 + k;
        const unsigned char* src = nullptr;
        if (ft >= 0 && ft < F) src = p.units + (((size_t)s * F + ft) * C + c) * kUnitBytes;
        else if (ft < 0 && started) src = p.hist + ((size_t)s * C + c) * kUnitBytes;
        ATDE_PAR_FOR(i, 216) ub[k][i] = (src && i < kUnitBytes) ? src[i] : 0;
        if (tid == 0) smode[k] = src ? 0 : 0x80;
    }
    __syncthreads();
    // ---- block-size mode + dequantisation: one warp per frame, lane per BFU ----
    for (int k = warp; k < kDecFrames; k += kDecThreads / 32) {
        float* specs = sp[k];
        for (int i = lane; i < 512; i += 32) specs[i] = 0.0f;
        __syncwarp();
        if (smode[k] & 0x80) continue;
        const unsigned char* u = ub[k];
        const int lc0 = 2 - (int)get_bits(u, 0, 2), lc1 = 2 - (int)get_bits(u, 2, 2), lc2 = 3 - (int)get_bits(u, 4, 2);
        static const unsigned char amount[8] = {20, 28, 32, 36, 40, 44, 48, 52};
        const int nbfu = (get_bits(u, 8, 3) == 0) ? 20 : (get_bits(u, 8, 3) == 1 ? 28 : 28 + 4 * ((int)get_bits(u, 8, 3) - 1));
        (void)amount;
        int wl[2], sf[2], nb[2];
        int mine = 0;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int b = lane + 32 * r;
            wl[r] = sf[r] = nb[r] = 0;
            if (b < nbfu) {
                const int raw = (int)get_bits(u, 16 + 4 * b, 4);
                wl[r] = raw ? raw + 1 : 0;                               // !!wordLens + wordLens
                sf[r] = (int)get_bits(u, 16 + 4 * nbfu + 6 * b, 6);
            }
            if (b < kMaxBfus) nb[r] = wl[r] * kDecSpecsPerBlock[b];
            mine += nb[r];
        }
        // bit position of every BFU's first mantissa: BFUs are read in index order
        int incl0 = nb[0];
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl0, d);
            if (lane >= d) incl0 += up;
        }
        const int tot0 = __shfl_sync(0xffffffffu, incl0, 31);
        int incl1 = nb[1];
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl1, d);
            if (lane >= d) incl1 += up;
        }
        const int total = 16 + 10 * nbfu + tot0 + __shfl_sync(0xffffffffu, incl1, 31);
        // a negative LogCount or a read past the end of the unit throws in the reference: the frame decodes as
        // silence with the neutral block size (atrac1denc.cpp:153-163)
        if (lc0 < 0 || lc1 < 0 || lc2 < 0 || total > kUnitBytes * 8) continue;
        if (lc0 == 1 || lc1 == 1 || lc2 == 1 || lc2 == 2) {              // never written by the reference encoder
            if (lane == 0) atomicOr(reinterpret_cast<unsigned*>(p.status), 1u);
            continue;
        }
        if (lane == 0) smode[k] = (unsigned char)((lc0 ? 1 : 0) | (lc1 ? 2 : 0) | (lc2 ? 4 : 0));
        const int start_bits[2] = {16 + 10 * nbfu + incl0 - nb[0], 16 + 10 * nbfu + tot0 + incl1 - nb[1]};
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int b = lane + 32 * r;
            if (b >= kMaxBfus || wl[r] == 0) continue;
            const int band = b < 20 ? 0 : (b < 36 ? 1 : 2);
            const bool shrt = band == 0 ? lc0 != 0 : (band == 1 ? lc1 != 0 : lc2 != 0);
            const int pos0 = shrt ? kDecStartShort[b] : kDecStartLong[b];
            const float scale = T->scale_table[sf[r]];
            const float max_quant = __double2float_rn(__ddiv_rn(1.0, (double)(float)((1 << (wl[r] - 1)) - 1)));
            const float mul = fmul(scale, max_quant);
            int bp = start_bits[r];
            for (int i = 0; i < kDecSpecsPerBlock[b]; i++) {
                const int v = (int)get_bits(u, bp, wl[r]);
                const int sv = (v << (32 - wl[r])) >> (32 - wl[r]);      // MakeSign
                specs[pos0 + i] = fmul(mul, (float)sv);
                bp += wl[r];
            }
        }
    }
    __syncthreads();

    // ---- IMDCT pre-twiddle (mdct.h:127-137) into kissfft's gather order; bands 1, 2 mirrored first (SwapArray) ----
    ATDE_PAR_FOR(uu, kDecFrames * 256) {
        const int k = uu >> 8, slot = uu & 255;
        const int b = slot < 64 ? 0 : (slot < 128 ? 1 : 2);
        const int bslot = slot - (b == 0 ? 0 : (b == 1 ? 64 : 128));
        const int size = (b == 2) ? 256 : 128;
        const bool shrt = (smode[k] >> b) & 1;
        const float* in = sp[k] + (b == 0 ? 0 : (b == 1 ? 128 : 256));
        int n2, i, kb = 0;
        const float* cs;
        if (!shrt) {
            n2 = size;
            i = (b == 2) ? T->perm128[bslot] : T->perm64[bslot];
            cs = (b == 2) ? T->isincos512 : T->isincos256;
        } else {
            n2 = 32;
            kb = bslot >> 4;
            i = T->perm16[bslot & 15];
            cs = T->isincos64;
        }
        in += 32 * kb;
        const int n = 2 * i;
        const float r0 = b ? in[n2 - 1 - n] : in[n];
        const float i0 = b ? in[n] : in[n2 - 1 - n];
        const float cc = cs[n], ss = cs[n + 1];
        cpx X;
        X.r = fmul(-2.0f, fadd(fmul(i0, ss), fmul(r0, cc)));
        X.i = fmul(-2.0f, fsub(fmul(i0, cc), fmul(r0, ss)));
        fft[k][slot] = X;
    }
    __syncthreads();
    // ---- FFT stages, innermost first.  Per frame: 16 (low) + 16 (mid) + 64 (hi) butterfly slots ----
    for (int st = 0; st < 4; st++) {
        ATDE_PAR_FOR(uu, kDecFrames * 96) {
            const int k = uu / 96, w = uu - k * 96;
            const int b = w < 16 ? 0 : (w < 32 ? 1 : 2);
            const int v = w - (b == 0 ? 0 : (b == 1 ? 16 : 32));
            const bool shrt = (smode[k] >> b) & 1;
            cpx* buf = fft[k] + (b == 0 ? 0 : (b == 1 ? 64 : 128));
            if (shrt) {
                const int ninst4 = (b == 2) ? 32 : 16;
                if (st < 2 && v < ninst4) {
                    const int inst = v >> 2, vv = v & 3;
                    kf_stage4<false>(buf + 16 * inst, T->tw16, vv, st == 0 ? 1 : 4, st == 0 ? 4 : 1);
                }
            } else if (b != 2) {
                if (st < 3 && v < 16) {
                    const int m = st == 0 ? 1 : (st == 1 ? 4 : 16);
                    kf_stage4<false>(buf, T->tw64, v, m, 16 / m);
                }
            } else {
                if (st == 0) {
                    kf_stage2(buf, T->tw128, v, 1, 64);
                } else if (v < 32) {
                    const int m = st == 1 ? 2 : (st == 2 ? 8 : 32);
                    kf_stage4<false>(buf, T->tw128, v, m, 32 / m);
                }
            }
        }
        __syncthreads();
    }
    // ---- post-twiddle (mdct.h:144-177): the middle half of TMIDCT's output, inv[n] = i1, inv[n2-1-n] = r1 ----
    ATDE_PAR_FOR(uu, kDecFrames * 256) {
        const int k = uu >> 8, slot = uu & 255;
        const int b = slot < 64 ? 0 : (slot < 128 ? 1 : 2);
        const int bslot = slot - (b == 0 ? 0 : (b == 1 ? 64 : 128));
        const int size = (b == 2) ? 256 : 128;
        const bool shrt = (smode[k] >> b) & 1;
        int n2, i, kb = 0;
        const float* cs;
        if (!shrt) { n2 = size; i = bslot; cs = (b == 2) ? T->isincos512 : T->isincos256; }
        else { n2 = 32; kb = bslot >> 4; i = bslot & 15; cs = T->isincos64; }
        const int n = 2 * i;
        const cpx z = fft[k][slot];
        const float cc = cs[n], ss = cs[n + 1];
        const float r1 = fadd(fmul(z.r, cc), fmul(z.i, ss));
        const float i1 = fsub(fmul(z.r, ss), fmul(z.i, cc));
        float* inv = sp[k] + (b == 0 ? 0 : (b == 1 ? 128 : 256)) + 32 * kb;
        inv[n] = i1;
        inv[n2 - 1 - n] = r1;
    }
    __syncthreads();
    // ---- overlap + window (vector_fmul_window, 16-sample sine slopes) -> band sample streams ----
    ATDE_PAR_FOR(uu, kDecFrames * 512) {
        const int k = uu >> 9, q = uu & 511;
        const int b = q < 128 ? 0 : (q < 256 ? 1 : 2);
        const int base = b == 0 ? 0 : (b == 1 ? 128 : 256);
        const int size = (b == 2) ? 256 : 128;
        const int pos = q - base;
        float* dst = (b == 0 ? lowS : (b == 1 ? midS : hiS)) + k * size;
        float v = 0.0f;
        if (!(smode[k] & 0x80)) {
            const bool shrt = (smode[k] >> b) & 1;
            const float* inv = sp[k] + base;
            const int start = shrt ? (pos & ~31) : 0;
            const int pp = pos - start;
            if (pp >= 32) {
                v = inv[pos - 16];                                   // memcpy(dstBuf + 32, &invBuf[16], ...)
            } else {
                // the 16 samples before this block: the previous block's second half, or the previous frame's tail
                float prev;
                const int pi = pp < 16 ? pp : 31 - pp;
                if (start > 0) prev = inv[start - 16 + pi];
                else if (k > 0 && !(smode[k - 1] & 0x80)) prev = sp[k - 1][base + size - 16 + pi];
                else prev = 0.0f;
                if (pp < 16) v = fsub(fmul(prev, W[31 - pp]), fmul(inv[start + 15 - pp], W[pp]));
                else v = fadd(fmul(prev, W[31 - pp]), fmul(inv[start + pp - 16], W[pp]));
            }
        }
        dst[pos] = v;
    }
    __syncthreads();
    // ---- synthesis QMF 2: (low, mid) -> low+mid band, samples 232 .. of the block's stream (qmf.h:66-89) ----
    ATDE_PAR_FOR(pr, kDecFrames * 128) {
        if (pr < 116) continue;
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int t = 0; t < 24; t++) {
            const int m = pr - 23 + t;
            const float l = lowS[m], u = midS[m];
            s1 = fadd(s1, fmul(fadd(l, u), qw[2 * t]));
            s2 = fadd(s2, fmul(fsub(l, u), qw[2 * t + 1]));
        }
        midlow[2 * pr] = s2;
        midlow[2 * pr + 1] = s1;
    }
    __syncthreads();
    // ---- synthesis QMF 1: (low+mid, hi delayed by 39) -> PCM of the tile's frames; clip; interleave ----
    ATDE_PAR_FOR(pj, kDecTile * 256) {
        const int pr = 256 + pj;                                     // output pair of the block's stream
        const int ft = t0 + (pj >> 8);
        if (ft >= F) continue;
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int t = 0; t < 24; t++) {
            const int m = pr - 23 + t;
            const float l = midlow[m], u = hiS[m - 39];
            s1 = fadd(s1, fmul(fadd(l, u), qw[2 * t]));
            s2 = fadd(s2, fmul(fsub(l, u), qw[2 * t + 1]));
        }
        float o0 = s2, o1 = s1;
        o0 = o0 > 1.0f ? 1.0f : o0; o0 = o0 < -1.0f ? -1.0f : o0;   // PcmValueMax / PcmValueMin (atrac1denc.cpp:166-170)
        o1 = o1 > 1.0f ? 1.0f : o1; o1 = o1 < -1.0f ? -1.0f : o1;
        float* out = p.pcm + ((size_t)s * F * 512 + (size_t)ft * 512 + 2 * (pj & 255)) * C + c;
        out[0] = o0;
        out[C] = o1;
    }
}

__global__ void at1_decode_carry_kernel(DecodeParams p)
{
    const int s = blockIdx.x;
    const int n = p.C * kUnitBytes;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        p.hist_out[(size_t)s * n + i] = p.units[((size_t)s * p.F + (p.F - 1)) * n + i];
    if (threadIdx.x == 0) p.started_out[s] = 1;
}

void launch_decode(const DecodeParams& p, cudaStream_t st)
{
    dim3 grid((p.F + kDecTile - 1) / kDecTile, p.S, p.C);
    ATDE_LAUNCH(at1_decode_kernel, grid, kDecThreads, 0, st, p);
    ATDE_LAUNCH(at1_decode_carry_kernel, p.S, 128, 0, st, p);
}

} // namespace at1
} // namespace atde
