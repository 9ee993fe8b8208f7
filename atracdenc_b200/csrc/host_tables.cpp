// host_tables.cpp — see host_tables.h.  Plain host C++, built WITHOUT -march / -ffast-math /
// FP contraction so every expression rounds exactly as in the reference's start-up code.
#include "host_tables.h"
#include <cmath>
#include <math.h>

namespace atde {

std::vector<int> kiss_factors(int n)
{
    // kf_factor (kiss_fft.c:308-330): strip 4s, then 2s, then odd primes
    std::vector<int> f;
    int p = 4;
    const double root = std::floor(std::sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2;
            else if (p == 2) p = 3;
            else p += 2;
            if (p > root) p = n;
        }
        n /= p;
        f.push_back(p);
        f.push_back(n);
    } while (n > 1);
    return f;
}

std::vector<cpxh> kiss_twiddles(int n, bool inverse)
{
    std::vector<cpxh> tw(n);
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < n; ++i) {
        double phase = -2 * pi * i / n;
        if (inverse) phase *= -1;
        tw[i].r = (float)::cos(phase);
        tw[i].i = (float)::sin(phase);
    }
    return tw;
}

std::vector<cpxh> kiss_super_twiddles(int n, bool inverse)
{
    const int nc = n >> 1;
    std::vector<cpxh> tw(nc / 2);
    for (int i = 0; i < nc / 2; ++i) {
        double phase = -3.14159265358979323846264338327 * ((double)(i + 1) / nc + .5);
        if (inverse) phase *= -1;
        tw[i].r = (float)::cos(phase);
        tw[i].i = (float)::sin(phase);
    }
    return tw;
}

std::vector<uint16_t> kiss_perm(int n)
{
    const std::vector<int> f = kiss_factors(n);
    const int ns = (int)f.size() / 2;
    std::vector<uint16_t> perm(n);
    for (int o = 0; o < n; o++) {
        int rem = o, idx = 0, stride = 1;
        for (int s = 0; s < ns; s++) {
            const int p = f[2 * s], m = f[2 * s + 1];
            const int q = rem / m;
            rem -= q * m;
            idx += q * stride;
            stride *= p;
        }
        perm[o] = (uint16_t)idx;
    }
    return perm;
}

std::vector<float> mdct_sincos(int n, float scale)
{
    // CalcSinCos (mdct.cpp:25-36): alpha/omiga/scale are float variables, the trig argument is a
    // float expression, cos/sin resolve to the float overloads.
    std::vector<float> t(n >> 1);
    const float alpha = 2.0 * M_PI / (8.0 * n);
    const float omiga = 2.0 * M_PI / n;
    scale = std::sqrt(scale / n);
    for (int i = 0; i < (n >> 2); ++i) {
        t[2 * i + 0] = scale * std::cos(omiga * i + alpha);
        t[2 * i + 1] = scale * std::sin(omiga * i + alpha);
    }
    return t;
}

void qmf_window(float w[48])
{
    static const float half[24] = {
        -0.00001461907,  -0.00009205479, -0.000056157569,  0.00030117269,
        0.0002422519,    -0.00085293897, -0.0005205574,    0.0020340169,
        0.00078333891,   -0.0042153862,  -0.00075614988,   0.0078402944,
        -0.000061169922, -0.01344162,    0.0024626821,     0.021736089,
        -0.007801671,    -0.034090221,   0.01880949,       0.054326009,
        -0.043596379,    -0.099384367,   0.13207909,       0.46424159
    };
    for (int i = 0; i < 24; i++)
        w[i] = w[47 - i] = half[i] * 2.0;
}

std::vector<float> loudness_curve(int sz)
{
    std::vector<float> res(sz);
    for (int i = 0; i < sz; i++) {
        float f = (float)(i + 3) * 0.5 * 44100 / (float)sz;
        float t = std::log10(f) - 3.5;
        t = -10 * t * t + 3 - f / 3000;
        t = std::pow(10, (0.1 * t));
        res[i] = t;
    }
    return res;
}

// ATH curve: table of the Musepack model the reference borrows (atrac_psy_common.cpp:33-95),
// values in millibel re 20 uPa, 4 steps per third starting at 10 Hz.
static float ath_frank(float freq)
{
    static const short tab[] = {
        9669, 9669, 9626, 9512, 9353, 9113, 8882, 8676, 8469, 8243, 7997, 7748, 7492, 7239, 7000, 6762,
        6529, 6302, 6084, 5900, 5717, 5534, 5351, 5167, 5004, 4812, 4638, 4466, 4310, 4173, 4050, 3922,
        3723, 3577, 3451, 3281, 3132, 3036, 2902, 2760, 2658, 2591, 2441, 2301, 2212, 2125, 2018, 1900,
        1770, 1682, 1594, 1512, 1430, 1341, 1260, 1198, 1136, 1057,  998,  943,  887,  846,  744,  712,
         693,  668,  637,  606,  580,  555,  529,  502,  475,  448,  422,  398,  375,  351,  327,  322,
         312,  301,  291,  268,  246,  215,  182,  146,  107,   61,   13,  -35,  -96, -156, -179, -235,
        -295, -350, -401, -421, -446, -499, -532, -535, -513, -476, -431, -313, -179,    8,  203,  403,
         580,  736,  881, 1022, 1154, 1251, 1348, 1421, 1479, 1399, 1285, 1193, 1287, 1519, 1914, 2369,
        3352, 4352, 5352, 6352, 7352, 8352, 9352, 9999, 9999, 9999, 9999, 9999,
    };
    if (freq < 10.) freq = 10.;
    if (freq > 29853.) freq = 29853.;
    const double freq_log = 40. * ::log10(0.1 * freq);
    const unsigned index = (unsigned)freq_log;
    return 0.01 * (tab[index] * (1 + index - freq_log) + tab[index + 1] * (freq_log - index));
}

std::vector<float> calc_ath(int len, int sample_rate)
{
    std::vector<float> res(len);
    const float mf = (float)sample_rate / 2000.0;
    for (size_t i = 0; i < res.size(); i++) {
        const float f = (float)(i + 1) * mf / len;
        float trh = ath_frank(1.e3 * f) - 100;
        trh -= f * f * 0.015;
        res[i] = trh;
    }
    return res;
}

} // namespace atde
