// Host check of at3_pack.cu's quant_mantissas / quant_unit_exact against the reference's own
// QuantMantisas (linked from oracle/_ref/libatde_ref.so): random scaled spectra with heavy value
// duplication (ties in |delta|), all word lengths, all ATRAC3 block sizes.  TEST TOOLING.
#include "at3_pack.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace NAtracDEnc { float QuantMantisas(const float* in, uint32_t first, uint32_t last, float mul, bool ea, int* mantisas); }
static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float urand() { return (float)((rnd() >> 40) * (1.0 / 16777216.0)); }
int main(int argc, char** argv)
{
    const long iters = argc > 1 ? atol(argv[1]) : 200000;
    const float mq[8] = {0.0f, 1.5f, 2.5f, 3.5f, 4.5f, 7.5f, 15.5f, 31.5f};
    const int sizes[5] = {8, 16, 32, 64, 128};
    long bad = 0, exact_calls = 0;
    for (long it = 0; it < iters; it++) {
        const int len = sizes[rnd() % 5];
        const int wl = 1 + (int)(rnd() % 7);
        const int mode = (int)(rnd() % 5);
        float in[128];
        const int palette = 1 + (int)(rnd() % 12);
        float pal[12];
        for (int i = 0; i < 12; i++) pal[i] = (2 * urand() - 1) * 0.99999f;
        const float amp = mode == 3 ? urand() * 0.1f : 1.0f;
        for (int j = 0; j < len; j++) {
            if (mode == 0) in[j] = (2 * urand() - 1) * 0.99999f;
            else if (mode == 1) in[j] = pal[rnd() % palette];                 // many exact duplicates -> ties
            else if (mode == 2) in[j] = (float)((int)(rnd() % 41) - 20) / 20.0f * 0.99999f;
            else if (mode == 3) in[j] = (2 * urand() - 1) * amp;
            else in[j] = (rnd() & 1) ? pal[rnd() % palette] : (2 * urand() - 1) * 0.5f;
        }
        const float mul = mq[wl];
        const float inv2 = 1.0 / (mul * mul);
        for (int ea = 0; ea < 2; ea++) {
            int ref_m[128];
            const float ref_e = NAtracDEnc::QuantMantisas(in, 0, len, mul, ea, ref_m);
            signed char m[128]; float ckey[128]; unsigned char cidx[128];
            const float e = atde::at3::quant_mantissas(in, len, ea, mul, inv2, m, ckey, cidx);
            bool ok = (memcmp(&e, &ref_e, 4) == 0) || (e != e && ref_e != ref_e);
            for (int j = 0; j < len; j++) ok = ok && (ref_m[j] == m[j]);
            signed char m2[128];
            const float e2 = ea ? atde::at3::quant_unit_exact(in, len, mul, inv2, m2) : e;
            bool ok2 = (memcmp(&e2, &ref_e, 4) == 0) || (e2 != e2 && ref_e != ref_e);
            if (ea) for (int j = 0; j < len; j++) ok2 = ok2 && (ref_m[j] == m2[j]);
            if (!ok || !ok2) { bad++; if (bad < 6) printf("mismatch len=%d wl=%d mode=%d ea=%d fast=%d exact=%d\n", len, wl, mode, ea, ok, ok2); }
        }
    }
    printf("iters=%ld mismatches=%ld\n", iters, bad);
    return bad != 0;
}
