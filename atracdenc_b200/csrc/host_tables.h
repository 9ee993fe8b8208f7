// host_tables.h — tables the kernels consume, computed ON THE HOST with the same expressions,
// types and libm entry points the reference uses at start-up, then uploaded once.  They are never
// recomputed with device libm (device sin/cos/pow are not bit-identical to glibc's).
#pragma once
#include <cstdint>
#include <vector>

namespace atde {

struct cpxh { float r, i; };

// kissfft twiddles: tw[i] = (float)cos/sin(-2*pi*i/n) (kiss_fft.c:357-363)
std::vector<cpxh> kiss_twiddles(int n, bool inverse);
// kiss_fftr super twiddles (tools/kiss_fftr.c:50-56), n = real FFT size
std::vector<cpxh> kiss_super_twiddles(int n, bool inverse);
// mixed-radix digit reversal implied by kf_work's recursion (kiss_fft.c:237-302) with the
// factorisation of kf_factor (:308-330): out slot o reads input perm[o]
std::vector<uint16_t> kiss_perm(int n);
// radix/sub-length pairs of kf_factor, outermost first
std::vector<int> kiss_factors(int n);
// MDCT pre/post twiddle table (mdct.cpp:25-36): n/2 floats (cos,sin interleaved)
std::vector<float> mdct_sincos(int n, float scale);
// 48-tap QMF window (qmf.cpp:25-45)
void qmf_window(float w[48]);
// loudness weighting curve (atrac_psy_common.cpp:142-156)
std::vector<float> loudness_curve(int sz);
// absolute threshold of hearing per spectral line, dB (atrac_psy_common.cpp:126-140)
std::vector<float> calc_ath(int len, int sample_rate);

} // namespace atde
