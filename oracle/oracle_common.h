/*
 * oracle/oracle_common.h — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * Plain-C restatement of the primitives of dcherednik/atracdenc's encode hot path.
 * Every function cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
 * library built from oracle/*.c.  Parity is PINNED: tests/test_oracle_vs_ref.py checks this
 * restatement bit-for-bit against oracle/_ref/libatde_ref.so (the unmodified reference sources
 * compiled here) — the reference's own tests hold no bit-exact golden vectors for this path
 * (SURVEY.md §8c), so outputs of the reference itself are the pin.
 *
 * All arithmetic is IEEE fp32/fp64, round-to-nearest-even, NO fused multiply-add
 * (build with -ffp-contract=off, no -march), operation order exactly as the reference.
 */
#ifndef ATDE_ORACLE_COMMON_H
#define ATDE_ORACLE_COMMON_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float r, i; } ocpx;

/* ---- kissfft restatement (src/lib/fft/kissfft_impl/kiss_fft.c) ---- */
typedef struct {
    int nfft, inverse;
    int factors[64];
    ocpx* tw;
} okiss;
okiss* okiss_alloc(int nfft, int inverse);              /* kiss_fft.c:339-369 */
void okiss_free(okiss*);
void okiss_fft(const okiss*, const ocpx* in, ocpx* out); /* kiss_fft.c:385-388, kf_work :237-302 */

typedef struct {
    okiss* sub;
    ocpx* super_tw;
    ocpx* tmp;
    int ncfft;
} okissr;
okissr* okissr_alloc(int nfft, int inverse);                    /* tools/kiss_fftr.c:29-59 */
void okissr_free(okissr*);
void okiss_fftr(const okissr*, const float* in, ocpx* out);     /* tools/kiss_fftr.c:61-115 */
void okiss_fftri(const okissr*, const ocpx* in, float* out);    /* tools/kiss_fftr.c:117-153 */

/* ---- MDCT restatement (src/lib/mdct/mdct.h:51-104, mdct.cpp:25-45) ---- */
typedef struct {
    int n;
    float* sincos;   /* n/2 floats */
    okiss* fft;      /* n/4 points */
    ocpx *fin, *fout;
} omdct;
omdct* omdct_alloc(int n, float scale);
void omdct_free(omdct*);
void omdct_run(const omdct*, const float* in, float* out /* n/2 */);

/* ---- 48-tap QMF (src/qmf/qmf.h:47-64, qmf.cpp:25-45) ---- */
void oqmf_window(float w[48]);
/* hist: 46 floats of state (in/out); in: n_in samples; lower/upper: n_in/2 each */
void oqmf_analysis(const float w[48], float* hist46, const float* in, int n_in, float* lower, float* upper);

/* ---- MSB-first bit writer (src/lib/bitstream/bitstream.cpp:40-63) ---- */
typedef struct {
    uint8_t buf[4096];
    int size;       /* std::vector<char>::size() as the reference grows it */
    int bits_used;
} obits;
void obits_init(obits*);
void obits_write(obits*, uint32_t val, int n);
int omake_sign(int val, unsigned bits);                 /* bitstream.h:27-31 */

/* ---- bisection driver (src/lib/bs_encode/encode.cpp:57-93) ---- */
typedef struct {
    size_t target;
    float min_l, max_l, cur, last;
    int need_repeat;
} obisect;
void obisect_start(obisect*, size_t target, float mn, float mx);
float obisect_continue(obisect*);
int obisect_submit(obisect*, size_t got);   /* returns 1 when finished */

/* ---- glibc 2.39 libm replicas (x86-64 ifunc "fma" variants, see glibc_replica.c) ---- */
float og_logf(float x);
float og_log10f(float x);
float og_log2f(float x);
double og_log(double x);
double og_exp(double x);

/* ---- psycho-acoustic tables (src/atrac/atrac_psy_common.cpp) ---- */
void ocalc_ath(int len, int sample_rate, float* out);            /* :126-140 */
void ocreate_loudness_curve(int sz, float* out);                 /* :142-156 */
float otrack_loudness2(float prev, float l0, float l1);          /* atrac_psy_common.h:46-49 */
float otrack_loudness1(float prev, float l);                     /* atrac_psy_common.h:51-54 */

#ifdef __cplusplus
}
#endif
#endif
