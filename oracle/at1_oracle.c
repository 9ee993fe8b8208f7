/*
 * oracle/at1_oracle.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * Plain-C, streaming restatement of the reference's ATRAC1 per-frame encode path:
 *   TAtrac1Encoder::GetLambda            src/atrac1denc.cpp:180-255
 *   Atrac1AnalysisFilterBank::Analysis   src/atrac/at1/atrac1_qmf.h:37-43
 *   TTransientDetector::Detect/HPFilter  src/transient_detector.cpp:52-93
 *   TAtrac1MDCT::Mdct                    src/atrac1denc.cpp:70-102
 *   TScaler::Scale/ScaleFrame            src/atrac/atrac_scale.cpp:141-188
 *   TAt1BitAlloc::Write + parts          src/atrac/at1/atrac1_bitalloc.cpp:80-409
 * Pinned bit-for-bit against oracle/_ref (tests/test_oracle_vs_ref.py).
 */
#include "oracle_common.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants, src/atrac/at1/atrac1.h:86-104 ---- */
#define MAX_BFUS 52
static const uint32_t SpecsPerBlock[MAX_BFUS] = {
    8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 7, 7, 7, 7, 9, 9, 9, 9, 10, 10, 10, 10,
    12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20, 20, 20, 20, 20, 20
};
static const uint32_t BlocksPerBand[4] = {0, 20, 36, 52};
static const uint32_t SpecsStartLong[MAX_BFUS] = {
    0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116, 122,
    128, 134, 140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226, 236, 246,
    256, 268, 280, 292, 304, 316, 328, 340, 352, 372, 392, 412, 432, 452, 472, 492,
};
static const uint32_t SpecsStartShort[MAX_BFUS] = {
    0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90, 122,
    128, 160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182, 214, 246,
    256, 288, 320, 352, 384, 416, 448, 480, 268, 300, 332, 364, 396, 428, 460, 492
};
static const uint32_t BfuAmountTab[8] = {20, 28, 32, 36, 40, 44, 48, 52};

/* src/atrac/at1/atrac1_bitalloc.cpp:37-67 */
static const float FixedBitAllocTableLong[MAX_BFUS] = {
    7, 7, 7, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4,
    4, 4, 3, 3, 3, 3, 3, 3, 2, 1, 1, 1, 1, 0, 0, 0
};
static const float FixedBitAllocTableShort[MAX_BFUS] = {
    6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
    4, 4, 4, 4, 4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0
};
static const float BitAllocSpread = 0.4f;
static const uint32_t BitBoostMask[MAX_BFUS] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
    1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1,
    1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0
};

static uint32_t bfu_to_band(uint32_t i) { return i < 20 ? 0 : (i < 36 ? 1 : 2); }

typedef struct {
    int init;
    float qmf_w[48];
    float sine_window[32];
    float scale_table[64];
    float loud_curve[512];
    float ath_long[MAX_BFUS];
    omdct *m512, *m256, *m64;
    /* boost map: (bits, pos) sorted by bits, insertion order within equal keys */
    uint32_t boost_bits[MAX_BFUS], boost_pos[MAX_BFUS];
    int n_boost;
} at1_tables;
static at1_tables T;

static void at1_init_tables(void)
{
    if (T.init) return;
    oqmf_window(T.qmf_w);
    for (uint32_t i = 0; i < 64; i++)                       /* atrac1.h:122-127 */
        T.scale_table[i] = pow(2.0, (double)(i / 3.0 - 21.0));
    for (uint32_t i = 0; i < 32; i++)                       /* atrac1.h:128-132 */
        T.sine_window[i] = sin((i + 0.5) * (M_PI / (2.0 * 32.0)));
    ocreate_loudness_curve(512, T.loud_curve);              /* atrac1denc.cpp:39 */
    /* CalcAt1ATH, atrac1_bitalloc.cpp:118-135 */
    float ath_spec[512];
    ocalc_ath(512, 44100, ath_spec);
    for (int b = 0; b < MAX_BFUS; b++) {
        float x = 999;
        for (uint32_t line = SpecsStartLong[b]; line < SpecsStartLong[b] + SpecsPerBlock[b]; line++)
            x = fmin(x, ath_spec[line]);
        x = pow(10, 0.1 * x);
        T.ath_long[b] = x;
    }
    T.m512 = omdct_alloc(512, 1);                           /* atrac1denc.h:49-51 */
    T.m256 = omdct_alloc(256, 0.5);
    T.m64 = omdct_alloc(64, 0.5);
    /* TBitsBooster ctor, atrac1_bitalloc.cpp:69-78: multimap ordered by nBits */
    T.n_boost = 0;
    for (uint32_t key = 1; key <= 20; key++)
        for (uint32_t i = 0; i < MAX_BFUS; i++)
            if (BitBoostMask[i] && SpecsPerBlock[i] == key) {
                T.boost_bits[T.n_boost] = key;
                T.boost_pos[T.n_boost++] = i;
            }
    T.init = 1;
}

/* ---------------------------------------------------------------------------------------- */
typedef struct {
    float hpf_buf[256 + 21];
    float last_energy;
    int block_sz;
} at1_detector;

typedef struct {
    float q1_hist[46], q2_hist[46];
    float delay[39 + 512];
    float low[256 + 16], mid[256 + 16], hi[512 + 16];
    at1_detector det[3];
} at1_channel;

typedef struct {
    int channels;
    int window_auto, window_mask, bfu_idx_const;
    float loudness;
    at1_channel ch[2];
} at1_encoder;

/* TTransientDetector::HPFilter, transient_detector.cpp:52-70 */
static void hp_filter(at1_detector* d, const float* in, float* out)
{
    static const float fircoef[] = {
        -8.65163e-18 * 2.0, -0.00851586 * 2.0, -6.74764e-18 * 2.0, 0.0209036 * 2.0,
        -3.36639e-17 * 2.0, -0.0438162 * 2.0, -1.54175e-17 * 2.0, 0.0931738 * 2.0,
        -5.52212e-17 * 2.0, -0.313819 * 2.0
    };
    const int B = d->block_sz;
    memcpy(d->hpf_buf + 20, in, B * sizeof(float));
    const float* inBuf = d->hpf_buf;          /* note: hpf_buf[B+20] is never written -> stays 0 */
    for (int i = 0; i < B; ++i) {
        float s = inBuf[i + 10];
        float s2 = 0;
        for (int j = 0; j < 9; j += 2) {
            s += fircoef[j] * (inBuf[i + j] + inBuf[i + 21 - j]);
            s2 += fircoef[j + 1] * (inBuf[i + j + 1] + inBuf[i + 21 - j - 1]);
        }
        out[i] = (s + s2) / 2;
    }
    memcpy(d->hpf_buf, in + (B - 20), 20 * sizeof(float));
}

/* calculateRMS, transient_detector.cpp:33-40 */
static float calc_rms(const float* in, uint32_t n)
{
    float s = 0;
    for (uint32_t i = 0; i < n; i++)
        s += (in[i] * in[i]);
    s /= n;
    return sqrtf(s);
}

/* TTransientDetector::Detect, transient_detector.cpp:73-93 (ShortSz = 16) */
static int detect(at1_detector* d, const float* buf)
{
    const int nshort = d->block_sz / 16;
    float rms[17];
    float filtered[256];
    hp_filter(d, buf, filtered);
    int trans = 0;
    rms[0] = d->last_energy;
    for (int i = 1; i < nshort + 1; ++i) {
        rms[i] = 19.0 * og_log10f(calc_rms(&filtered[(i - 1) * 16], 16));
        if (rms[i] - rms[i - 1] > 16)
            trans = 1;
        if (rms[i - 1] - rms[i] > 20)
            trans = 1;
    }
    d->last_energy = rms[nshort];
    return trans;
}

/* TAtrac1MDCT::Mdct, atrac1denc.cpp:70-102.  logcount = {0|2, 0|2, 0|3} */
static void at1_mdct(float specs[512], float* low, float* mid, float* hi, const int logcount[3])
{
    uint32_t pos = 0;
    for (uint32_t band = 0; band < 3; band++) {
        const uint32_t nblocks = 1u << logcount[band];
        float* src = (band == 0) ? low : (band == 1) ? mid : hi;
        const uint32_t buf_sz = (band == 2) ? 256 : 128;
        const uint32_t block_sz = (nblocks == 1) ? buf_sz : 32;
        const uint32_t win_start = (nblocks == 1) ? ((band == 2) ? 112 : 48) : 0;
        const float multiple = (nblocks != 1 && band == 2) ? 2.0 : 1.0;
        float tmp[512];
        memset(tmp, 0, sizeof(tmp));
        uint32_t block_pos = 0;
        for (uint32_t k = 0; k < nblocks; ++k) {
            memcpy(&tmp[win_start], &src[buf_sz], 32 * sizeof(float));
            for (uint32_t i = 0; i < 32; i++) {
                src[buf_sz + i] = T.sine_window[i] * src[block_pos + block_sz - 32 + i];
                src[block_pos + block_sz - 32 + i] = T.sine_window[31 - i] * src[block_pos + block_sz - 32 + i];
            }
            memcpy(&tmp[win_start + 32], &src[block_pos], block_sz * sizeof(float));
            float sp[256];
            uint32_t n_sp;
            if (nblocks == 1) {
                if (band == 2) { omdct_run(T.m512, tmp, sp); n_sp = 256; }
                else { omdct_run(T.m256, tmp, sp); n_sp = 128; }
            } else {
                omdct_run(T.m64, tmp, sp); n_sp = 32;
            }
            for (uint32_t i = 0; i < n_sp; i++)
                specs[block_pos + pos + i] = sp[i] * multiple;
            if (band) {                                       /* SwapArray, util.h:42-49 */
                float* p = &specs[block_pos + pos];
                for (uint32_t i = 0, j = n_sp - 1; i < n_sp / 2; ++i, --j) {
                    float t = p[i]; p[i] = p[j]; p[j] = t;
                }
            }
            block_pos += 32;
        }
        pos += buf_sz;
    }
}

typedef struct {
    uint8_t sfi;
    float values[20];
    float energy;
} at1_block;

/* TScaler::Scale, atrac_scale.cpp:141-172 */
static void scale_block(const float* in, uint32_t len, at1_block* res)
{
    float max_abs = 0;
    for (uint32_t i = 0; i < len; ++i) {
        const float a = fabsf(in[i]);
        if (a > max_abs)
            max_abs = a;
    }
    if (max_abs > 1.0f)
        max_abs = 1.0f;
    int idx = 0;                       /* std::map::lower_bound over ScaleTable */
    while (idx < 63 && T.scale_table[idx] < max_abs)
        idx++;
    const float sf = T.scale_table[idx];
    res->sfi = (uint8_t)idx;
    res->energy = 0.0;
    for (uint32_t i = 0; i < len; ++i) {
        float v = in[i] / sf;
        float e = in[i] * in[i];
        res->energy += e;
        if (fabsf(v) >= 1.0)
            v = (v > 0) ? 0.99999 : -0.99999;
        res->values[i] = v;
    }
}

/* CalcLowToMidTilt, atrac1_bitalloc.cpp:147-161 */
static float low_to_mid_tilt(const at1_block* blocks, uint32_t nbfu)
{
    float sum_low = 0.0f, sum_mid = 0.0f;
    uint32_t n_low = 0, n_mid = 0;
    for (uint32_t i = 0; i < nbfu; ++i) {
        switch (bfu_to_band(i)) {
            case 0: sum_low += blocks[i].sfi; n_low++; break;
            case 1: sum_mid += blocks[i].sfi; n_mid++; break;
            default: break;
        }
    }
    if (!n_low || !n_mid)
        return 0.0f;
    return sum_low / n_low - sum_mid / n_mid;
}

/* CalcBitsAllocation, atrac1_bitalloc.cpp:163-205 */
static void calc_bits_allocation(const at1_block* blocks, uint32_t nbfu, float spread, float shift,
                                 const int logcount[3], float loudness, uint32_t* bits)
{
    const float tilt = low_to_mid_tilt(blocks, nbfu);
    const float mid_bias = fminf(1.5f, 0.3f * fmaxf(0.0f, tilt - 7.0f));
    const float band_bias[3] = {0.0f, mid_bias, mid_bias * 0.5f};
    for (uint32_t i = 0; i < nbfu; ++i) {
        int short_block = logcount[bfu_to_band(i)] != 0;
        const float fix = short_block ? FixedBitAllocTableShort[i] : FixedBitAllocTableLong[i];
        float ath = T.ath_long[i] * loudness;
        if (!short_block && blocks[i].energy < ath) {
            bits[i] = 0;
        } else {
            int tmp = spread * ((float)blocks[i].sfi / 3.2f) + (1.0f - spread) * fix - shift
                      + band_bias[bfu_to_band(i)];
            if (tmp > 16) bits[i] = 16;
            else if (tmp < 2) bits[i] = 0;
            else bits[i] = tmp;
        }
    }
}

/* GetMaxUsedBfuId, atrac1_bitalloc.cpp:207-230 */
static uint32_t max_used_bfu_id(const uint32_t* bits, uint32_t size)
{
    uint32_t idx = 7;
    for (;;) {
        uint32_t nbfu = BfuAmountTab[idx];
        if (nbfu > size) {
            idx--;
        } else if (idx != 0) {
            uint32_t i = 0;
            while (idx && bits[nbfu - 1 - i] == 0) {
                if (++i >= (BfuAmountTab[idx] - BfuAmountTab[idx - 1])) {
                    idx--;
                    nbfu -= i;
                    i = 0;
                }
            }
            break;
        } else {
            break;
        }
    }
    return idx;
}

/* TBitsBooster::ApplyBoost, atrac1_bitalloc.cpp:80-114 */
static uint32_t apply_boost(uint32_t* bits, uint32_t size, uint32_t cur, uint32_t target)
{
    const uint32_t max_per_iter = T.boost_bits[T.n_boost - 1];
    const uint32_t min_key = T.boost_bits[0];
    uint32_t surplus = target - cur;
    uint32_t key = (surplus > max_per_iter) ? max_per_iter : surplus;
    int max_it = 0;                                 /* upper_bound(key) */
    while (max_it < T.n_boost && T.boost_bits[max_it] <= key)
        max_it++;
    if (max_it == 0)
        return surplus;
    while (surplus >= min_key) {
        int done = 1;
        for (int it = 0; it < max_it; ++it) {
            const uint32_t cur_bits = T.boost_bits[it];
            const uint32_t cur_pos = T.boost_pos[it];
            if (cur_pos >= size)
                break;
            if (bits[cur_pos] == 16u)
                continue;
            const uint32_t per_spec = bits[cur_pos] ? 1 : 2;
            if (bits[cur_pos] == 0u && cur_bits * 2 > surplus)
                continue;
            if (cur_bits * per_spec > surplus)
                continue;
            bits[cur_pos] += per_spec;
            surplus -= cur_bits * per_spec;
            done = 0;
        }
        if (done)
            break;
    }
    return surplus;
}

static uint32_t avail_bits(uint32_t nbfu) { return 212 * 8 - 3 - 32 - 2 - 3 - nbfu * (4 + 6); }

/* TAt1BitAlloc::Write -> TBitStreamEncoder::Do over {TConfigure, TBfuAlloc}
 * (atrac1_bitalloc.cpp:240-409, encode.cpp:100-129) */
static int at1_write(const at1_block* blocks, const int logcount[3], float loudness, uint32_t bfu_idx_const,
                     uint8_t* out /* >= 256 */)
{
    uint32_t bfu_idx = bfu_idx_const ? bfu_idx_const - 1 : 7;
    const int auto_bfu = !bfu_idx_const;
    uint32_t alloc[MAX_BFUS];
    uint32_t nbfu;
    obisect ba;
    for (;;) {                                          /* Repeat => restart from TConfigure */
        nbfu = BfuAmountTab[bfu_idx];
        obisect_start(&ba, avail_bits(nbfu), -3, 15);
        uint32_t bits_used;
        for (;;) {                                      /* NeedRepeat => re-enter TBfuAlloc::Encode */
            float shift = obisect_continue(&ba);
            calc_bits_allocation(blocks, nbfu, BitAllocSpread, shift, logcount, loudness, alloc);
            bits_used = 0;
            for (uint32_t i = 0; i < nbfu; i++)
                bits_used += SpecsPerBlock[i] * alloc[i];
            if (obisect_submit(&ba, bits_used))
                break;
        }
        if (auto_bfu) {
            uint32_t used = max_used_bfu_id(alloc, nbfu);
            if (used < bfu_idx) {
                bfu_idx--;
                continue;
            }
        }
        apply_boost(alloc, nbfu, bits_used, avail_bits(nbfu));
        break;
    }
    /* TBfuAlloc::Dump, atrac1_bitalloc.cpp:279-327 */
    obits bs;
    obits_init(&bs);
    obits_write(&bs, 0x2 - logcount[0], 2);
    obits_write(&bs, 0x2 - logcount[1], 2);
    obits_write(&bs, 0x3 - logcount[2], 2);
    obits_write(&bs, 0, 2);
    obits_write(&bs, bfu_idx, 3);
    obits_write(&bs, 0, 2);
    obits_write(&bs, 0, 3);
    for (uint32_t i = 0; i < nbfu; i++)
        obits_write(&bs, alloc[i] ? (alloc[i] - 1) : 0, 4);
    for (uint32_t i = 0; i < nbfu; i++)
        obits_write(&bs, blocks[i].sfi, 6);
    for (uint32_t i = 0; i < nbfu; i++) {
        const uint32_t wl = alloc[i];
        if (wl == 0 || wl == 1)
            continue;
        const float multiple = ((1 << (wl - 1)) - 1);
        for (uint32_t j = 0; j < SpecsPerBlock[i]; j++) {
            const int tmp = lrintf(blocks[i].values[j] * multiple);
            obits_write(&bs, omake_sign(tmp, wl), wl);
        }
    }
    obits_write(&bs, 0x0, 8);
    obits_write(&bs, 0x0, 8);
    obits_write(&bs, 0x0, 8);
    memcpy(out, bs.buf, bs.size);
    return bs.size;
}

/* ---------------------------------------------------------------------------------------- */
void* oat1_create(int channels, int window_auto, int window_mask, int bfu_idx_const)
{
    at1_init_tables();
    at1_encoder* e = (at1_encoder*)calloc(1, sizeof(at1_encoder));
    e->channels = channels;
    e->window_auto = window_auto;
    e->window_mask = window_mask;
    e->bfu_idx_const = bfu_idx_const;
    e->loudness = 0.006f;                                  /* atrac1denc.h:101-102 */
    for (int c = 0; c < 2; c++) {
        e->ch[c].det[0].block_sz = 128;
        e->ch[c].det[1].block_sz = 128;
        e->ch[c].det[2].block_sz = 256;
    }
    return e;
}

void oat1_destroy(void* h) { free(h); }

/*
 * One call of the encoder lambda (atrac1denc.cpp:201-254): 512 interleaved sample-frames in,
 * `channels` sound units out (each zero-padded to 212 bytes in `out`, true WriteFrame payload
 * lengths in sizes[]).  Optional taps: specs [C][512], masks [C], chloud [C], sfi [C][52].
 */
void oat1_frame(void* h, const float* data, uint8_t* out, int* sizes,
                float* tap_specs, uint8_t* tap_masks, float* tap_chloud, float* tap_loud, uint8_t* tap_sfi)
{
    at1_encoder* e = (at1_encoder*)h;
    const int C = e->channels;
    int logcount[2][3];
    uint32_t masks[2] = {0, 0};
    float specs[2][512];
    float chl[2] = {0, 0};
    for (int c = 0; c < C; c++) {
        at1_channel* ch = &e->ch[c];
        float src[512];
        for (int i = 0; i < 512; ++i)
            src[i] = data[i * C + c];
        /* Atrac1AnalysisFilterBank::Analysis, atrac1_qmf.h:37-43 */
        float midlow[256];
        memcpy(&ch->delay[0], &ch->delay[256], sizeof(float) * 39);
        oqmf_analysis(T.qmf_w, ch->q1_hist, src, 512, midlow, &ch->delay[39]);
        oqmf_analysis(T.qmf_w, ch->q2_hist, midlow, 256, ch->low, ch->mid);
        memcpy(ch->hi, &ch->delay[0], sizeof(float) * 256);

        if (e->window_auto) {                              /* atrac1denc.cpp:214-229 */
            float inv[256];
            masks[c] |= (uint32_t)detect(&ch->det[0], ch->low);
            memcpy(inv, ch->mid, 128 * sizeof(float));
            for (int i = 0; i < 128; i += 2) inv[i] *= -1;   /* InvertSpectr, util.h:51-63 */
            masks[c] |= (uint32_t)detect(&ch->det[1], inv) << 1;
            memcpy(inv, ch->hi, 256 * sizeof(float));
            for (int i = 0; i < 256; i += 2) inv[i] *= -1;
            masks[c] |= (uint32_t)detect(&ch->det[2], inv) << 2;
        } else {
            masks[c] = e->window_mask;
        }
        logcount[c][0] = (masks[c] & 1) ? 2 : 0;           /* TBlockSizeMod::Create, atrac1.h:60-66 */
        logcount[c][1] = (masks[c] & 2) ? 2 : 0;
        logcount[c][2] = (masks[c] & 4) ? 3 : 0;
        at1_mdct(specs[c], ch->low, ch->mid, ch->hi, logcount[c]);
        float l = 0.0;
        for (int i = 0; i < 512; i++) {                    /* atrac1denc.cpp:235-240 */
            float en = specs[c][i] * specs[c][i];
            l += en * T.loud_curve[i];
        }
        chl[c] = l;
    }
    if (C == 2 && masks[0] == 0 && masks[1] == 0)          /* atrac1denc.cpp:243-247 */
        e->loudness = otrack_loudness2(e->loudness, chl[0], chl[1]);
    else if (masks[0] == 0)
        e->loudness = otrack_loudness1(e->loudness, chl[0]);

    for (int c = 0; c < C; c++) {
        at1_block blocks[MAX_BFUS];
        for (uint32_t band = 0; band < 3; ++band) {        /* ScaleFrame, atrac_scale.cpp:174-188 */
            const int short_win = logcount[c][band] != 0;
            for (uint32_t b = BlocksPerBand[band]; b < BlocksPerBand[band + 1]; ++b) {
                const uint32_t start = short_win ? SpecsStartShort[b] : SpecsStartLong[b];
                scale_block(&specs[c][start], SpecsPerBlock[b], &blocks[b]);
            }
        }
        uint8_t buf[512];
        memset(buf, 0, sizeof(buf));
        int n = at1_write(blocks, logcount[c], e->loudness / 0.006f, e->bfu_idx_const, buf);
        memcpy(out + c * 212, buf, 212);
        if (sizes) sizes[c] = n;
        if (tap_specs) memcpy(tap_specs + c * 512, specs[c], 512 * sizeof(float));
        if (tap_masks) tap_masks[c] = (uint8_t)masks[c];
        if (tap_chloud) tap_chloud[c] = chl[c];
        if (tap_sfi) for (int b = 0; b < MAX_BFUS; b++) tap_sfi[c * 52 + b] = blocks[b].sfi;
    }
    if (tap_loud) *tap_loud = e->loudness;
}

/* Whole-stream convenience: n_frames frames of interleaved PCM -> [n_frames][C][212] + sizes */
void oat1_encode(int channels, int window_auto, int window_mask, int bfu_idx_const,
                 const float* pcm, long n_frames, uint8_t* out, int* sizes)
{
    void* h = oat1_create(channels, window_auto, window_mask, bfu_idx_const);
    for (long f = 0; f < n_frames; f++)
        oat1_frame(h, pcm + (size_t)f * 512 * channels, out + (size_t)f * channels * 212,
                   sizes ? sizes + f * channels : NULL, NULL, NULL, NULL, NULL, NULL);
    oat1_destroy(h);
}

void oat1_tables(float* qmf48, float* sine32, float* scale64, float* loud512, float* ath52)
{
    at1_init_tables();
    memcpy(qmf48, T.qmf_w, sizeof(T.qmf_w));
    memcpy(sine32, T.sine_window, sizeof(T.sine_window));
    memcpy(scale64, T.scale_table, sizeof(T.scale_table));
    memcpy(loud512, T.loud_curve, sizeof(T.loud_curve));
    memcpy(ath52, T.ath_long, sizeof(T.ath_long));
}
