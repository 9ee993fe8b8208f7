// Host check of at3_pack.cu's warp-cooperative quantiser (compute_units: fast re-rounding walk and the
// exact fallback) against the reference's own QuantMantisas (linked from oracle/_ref/libatde_ref.so):
// random scaled spectra with heavy value duplication (ties in |delta|), random word length per BFU.
// The kernel code runs on the pthread CUDA shim, one 32-thread warp per launch.  TEST TOOLING.
#include "at3_pack.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace NAtracDEnc { float QuantMantisas(const float* in, uint32_t first, uint32_t last, float mul, bool ea, int* mantisas); }
using namespace atde::at3;
static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float urand() { return (float)((rnd() >> 40) * (1.0 / 16777216.0)); }

static PackShared g_sh;
static unsigned g_cv[32];
static float g_err[32];
static int g_wl[32];

static void unit_kernel(int dummy)
{
    (void)dummy;
    const int lane = threadIdx.x;
    const int start = kBlockStart[lane], len = kBlockStart[lane + 1] - kBlockStart[lane];
    float e1 = 0.0f;
    for (int j = 0; j < len; j++) e1 = atde::fadd(e1, atde::fmul(g_sh.sv[start + j], g_sh.sv[start + j]));
    float er = 0.0f;
    const unsigned cv = compute_units(g_sh, lane, g_wl[lane] != 0, g_wl[lane], start, len, e1, er);
    g_cv[lane] = cv; g_err[lane] = er;
}

int main(int argc, char** argv)
{
    const long iters = argc > 1 ? atol(argv[1]) : 3000;
    const float mq[8] = {0.0f, 1.5f, 2.5f, 3.5f, 4.5f, 7.5f, 15.5f, 31.5f};
    long bad = 0, units = 0;
    for (long it = 0; it < iters; it++) {
        const int mode = (int)(rnd() % 5);
        const int palette = 1 + (int)(rnd() % 12);
        float pal[12];
        for (int i = 0; i < 12; i++) pal[i] = (2 * urand() - 1) * 0.99999f;
        const float amp = mode == 3 ? urand() * 0.1f : 1.0f;
        for (int j = 0; j < 1024; j++) {
            float v;
            if (mode == 0) v = (2 * urand() - 1) * 0.99999f;
            else if (mode == 1) v = pal[rnd() % palette];                    // many exact duplicates -> ties
            else if (mode == 2) v = (float)((int)(rnd() % 41) - 20) / 20.0f * 0.99999f;
            else if (mode == 3) v = (2 * urand() - 1) * amp;
            else v = (rnd() & 1) ? pal[rnd() % palette] : (2 * urand() - 1) * 0.5f;
            g_sh.sv[j] = v;
        }
        for (int b = 0; b < 32; b++) g_wl[b] = (rnd() % 8 == 0) ? 0 : 1 + (int)(rnd() % 7);
        cuemu::launch(unit_kernel, dim3(1), dim3(32), 0, 0);
        for (int b = 0; b < 32; b++) {
            if (!g_wl[b]) continue;
            units++;
            const int start = kBlockStart[b], len = kBlockStart[b + 1] - start, wl = g_wl[b];
            int ref_m[128];
            const float ref_e = NAtracDEnc::QuantMantisas(g_sh.sv + start, 0, len, mq[wl], b > 18, ref_m);
            unsigned vlc = 0;
            if (wl > 1) { for (int j = 0; j < len; j++) vlc += vlc_bits_of(wl, ref_m[j]); }
            else { for (int j = 0; j < len / 2; j++) vlc += vlc_pair_bits(ref_m[2 * j], ref_m[2 * j + 1]); }
            const unsigned clc = wl > 1 ? (unsigned)kClcLen[wl] * len : 4u * len / 2;
            bool ok = (memcmp(&g_err[b], &ref_e, 4) == 0) || (g_err[b] != g_err[b] && ref_e != ref_e);
            ok = ok && g_cv[b] == (clc | (vlc << 16));
            for (int j = 0; j < len; j++) ok = ok && (ref_m[j] == g_sh.mant[start + j]);
            if (!ok) { bad++; if (bad < 6) printf("mismatch it=%ld bfu=%d wl=%d mode=%d\n", it, b, wl, mode); }
        }
    }
    printf("iters=%ld units=%ld mismatches=%ld\n", iters, units, bad);
    return bad != 0;
}
