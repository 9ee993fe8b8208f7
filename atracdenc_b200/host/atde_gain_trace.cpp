// atde_gain_trace.cpp — see atde_gain_trace.h.
//
// Every number printed is either a device value (taps) or a few float operations on device values in the order the
// reference performs them (this file is built with -ffp-contract=off like the rest of the host side; std::log2 and
// std::sqrt are the libm the reference calls).  The text is produced with the same iostream manipulators the reference
// uses, so equal values give equal bytes.
#include "atde_gain_trace.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <ostream>
#include <stdexcept>
#include <string>

namespace NAtracDEnc {

namespace {

constexpr int kSub = 32;              // sub-frames per band frame (AnalyzeGain(..., 2048, 32), atrac3denc.cpp:331)
constexpr int kNeutral = 4;           // TAtrac3Data::ExponentOffset: GainLevel[4] == 1
constexpr int kPointBudget = 6;       // kMaxCurvePoints (transient_detector.cpp:462)
constexpr int kScoreWindow = 3;       // kTransientScoreWindow

struct TFix { int Digits; };
std::ostream& operator<<(std::ostream& os, TFix f) { return os << std::fixed << std::setprecision(f.Digits); }

void Sequence(std::ostream& os, const float* v, int n, int digits)      // YamlWriteFloatSeq (yaml_log.h:76-85)
{
    os << TFix{digits} << "[";
    for (int i = 0; i < n; i++)
        os << (i ? ", " : "") << v[i];
    os << "]";
}

int TopBit(uint32_t x)                 // GetFirstSetBit (util.h:65-76): position of the highest set bit, 0 for 0
{
    int n = 0;
    while (x >>= 1) n++;
    return n;
}

int LevelOfRatio(float x)              // RelationToIdx (transient_detector.cpp:139-147)
{
    if (x <= 0.5f) {
        x = 1.0f / std::max(x, 0.00048828125f);
        return kNeutral + TopBit(static_cast<uint32_t>(x));
    }
    x = std::min(x, 16.0f);
    return kNeutral - TopBit(static_cast<uint32_t>(x));
}

// MedianFilter<1> (transient_detector.cpp:149-166): the middle of the sorted window; a two-element window at the
// edges gives its larger element
void Median3(const float* in, float* out)
{
    for (int i = 0; i < kSub; i++) {
        const float a = in[std::max(i - 1, 0)], b = in[i], c = in[std::min(i + 1, kSub - 1)];
        if (i == 0 || i == kSub - 1)
            out[i] = std::max(a, c);
        else
            out[i] = std::max(std::min(a, b), std::min(std::max(a, b), c));
    }
}

struct TPlateau { float Level, MaxRaw; bool Release; };

TPlateau Plateau(const float* in)     // FindPlateau(in, 3) (transient_detector.cpp:175-236)
{
    TPlateau r{0.0f, 0.0f, false};
    for (int i = 0; i < kSub; i++) r.MaxRaw = std::max(r.MaxRaw, in[i]);
    float filt[kSub];
    Median3(in, filt);
    float best = 0.0f;
    int end = -1;
    for (int j = 0; j + 3 <= kSub; j++) {
        const float m = std::min(std::min(filt[j], filt[j + 1]), filt[j + 2]);
        if (m > best) { best = m; end = j + 2; }
    }
    if (best < 1e-6f)
        return r;
    while (end + 1 < kSub && filt[end + 1] >= best) end++;
    if (end < kSub - 1) {
        if (in[kSub - 1] < best * 0.1f) {
            r.Release = true;
        } else {
            bool recovers = false;
            for (int i = end + 1; i < kSub && !recovers; i++) recovers = in[i] >= best * 0.7f;
            r.Release = !recovers && in[kSub - 1] < best * 0.5f;
        }
    }
    r.Level = best;
    return r;
}

float EdgeScore(const float* env, int loc)     // BoundaryTransientScore(env, loc, 3) (transient_detector.cpp:251-274)
{
    float left = 0.0f, right = 0.0f;
    for (int i = std::max(0, loc - kScoreWindow); i < loc; i++) left = std::max(left, env[i]);
    for (int i = loc; i < std::min(kSub, loc + kScoreWindow); i++) right = std::max(right, env[i]);
    const float eps = 1e-9f;
    return std::max((right + eps) / (left + eps), (left + eps) / (right + eps));
}

void Need(int64_t rc, const char* what)
{
    if (rc < 0)
        throw std::runtime_error(std::string("atde_b200: gain trace: ") + what + ": " + atde_last_error());
}

} // namespace

TGainTraceWriter::TGainTraceWriter(std::ostream* out, int channels, bool gainControl)
    : Out(out)
    , Channels(channels)
    , GainControl(gainControl)
    , State((size_t)channels * 4)
{
    // TAtrac3Data's tables (src/atrac/at3/atrac3.h:184-197), same expressions and types
    for (int i = 0; i < 256; i++) Window[i] = (sin(((i + 0.5) / 256.0 - 0.5) * M_PI) + 1.0);
    for (int i = 0; i < 16; i++) Level[i] = pow(2.0, kNeutral - i);
    for (int i = 0; i < 31; i++) Interp[i] = pow(2.0, -1.0 / 8 * (i - 15));
}

// CalcCurve (transient_detector.cpp:276-482) with its log lines
TGainTraceWriter::TCurve TGainTraceWriter::BuildCurve(const float* in, const float* low, const float* high, TBandState& st,
                                                      float minScore)
{
    std::ostream& os = *Out;
    TCurve curve;
    const TPlateau pl = Plateau(in);
    const bool usePlateau = pl.Level > 1e-6f && !pl.Release && pl.Level >= pl.MaxRaw * 0.4f;
    const float target = usePlateau ? pl.Level : in[kSub - 1];
    os << TFix{6}
       << "        plateau_level: " << pl.Level << "\n"
       << "        plateau_max_raw: " << pl.MaxRaw << "\n"
       << "        plateau_release: " << (pl.Release ? "true" : "false") << "\n"
       << "        target: " << target << "  # source: " << (usePlateau ? "plateau" : "in.back") << "\n";
    const float prevLevel = st.LastLevel, prevTarget = st.LastTarget;
    st.LastLevel = in[kSub - 1];
    st.LastTarget = target;
    if (target < 1e-6f || prevLevel < 1e-6f)
        return curve;

    float filt[kSub];
    Median3(in, filt);
    float peak = 0.0f;
    for (int i = 0; i < kSub; i++) peak = std::max(peak, in[i]);
    const float intra = peak / std::max(target, 1e-9f);
    float inter = 1.0f;
    if (prevTarget > 1e-6f)
        inter = std::max(prevTarget, target) / std::max(std::min(prevTarget, target), 1e-9f);
    const bool sticky = intra <= 7.0f && inter <= 10.0f;
    os << TFix{4}
       << "        sticky_frame_eligible: " << (sticky ? "true" : "false") << "\n"
       << "        sticky_intra_ratio: " << intra << "\n"
       << "        sticky_inter_ratio: " << inter << "\n";

    int lev[kSub];
    for (int i = 0; i < kSub; i++) {
        int level = LevelOfRatio(filt[i] / target);
        if (i > 0 && sticky) {
            float lo = low[i] / target, hi = high[i] / target;
            if (lo > hi) std::swap(lo, hi);
            const int a = LevelOfRatio(lo), b = LevelOfRatio(hi);
            const int mn = std::min(a, b), mx = std::max(a, b), prev = lev[i - 1];
            if (mx - mn <= 1 && std::abs(level - prev) == 1 && prev >= mn && prev <= mx)
                level = prev;
        }
        lev[i] = level;
    }
    int tail = 0;                                      // targetSf: everything from here on is neutral
    for (int sf = kSub - 2; sf >= 0; sf--)
        if (lev[sf] != kNeutral) { tail = sf + 1; break; }
    if (tail == 0)
        return curve;
    os << TFix{4}
       << "        transient_min_score: " << minScore << "\n"
       << "        transient_window: " << kScoreWindow << "\n";

    struct TStep { int Loc, Level, Delta; };
    std::vector<TStep> steps;                          // right to left
    int anchor = kNeutral;
    for (int sf = tail - 1; sf >= 0; sf--) {
        if (lev[sf] == anchor)
            continue;
        const int loc = sf + 1, delta = std::abs(lev[sf] - anchor);
        const float score = EdgeScore(filt, loc);
        if (loc == tail || delta >= 2 || score >= minScore) {
            steps.push_back({loc, lev[sf], delta});
            anchor = lev[sf];
        } else {
            os << TFix{4} << "        transition_pruned: {loc: " << loc << ", delta: " << delta << ", score: " << score << "}\n";
        }
    }
    if (steps.empty())
        return curve;
    if ((int)steps.size() > kPointBudget) {
        // the six largest jumps, the rightmost first among equals (locations are distinct: a strict order)
        std::sort(steps.begin(), steps.end(), [](const TStep& a, const TStep& b) {
            return a.Delta != b.Delta ? a.Delta > b.Delta : a.Loc > b.Loc;
        });
        steps.resize(kPointBudget);
    }
    std::sort(steps.begin(), steps.end(), [](const TStep& a, const TStep& b) { return a.Loc < b.Loc; });
    for (const TStep& s : steps)
        curve.push_back({(uint16_t)s.Level, (uint32_t)s.Loc});
    return curve;
}

namespace {

// BuildSubframeDivisors (atrac3denc.cpp:228-255): mean divisor of each 8-sample sub-frame
template <class TCurve>
void SubframeDivisors(const TCurve& pts, const float* level, const float* interp, float out[kSub])
{
    float div[256];
    std::fill(div, div + 256, 1.0f);
    uint32_t pos = 0;
    for (size_t i = 0; i < pts.size(); i++) {
        const uint32_t ramp = pts[i].Loc << 3;
        float l = level[pts[i].Level];
        const float step = interp[(i + 1 < pts.size() ? (int)pts[i + 1].Level : kNeutral) - (int)pts[i].Level + 15];
        for (; pos < ramp && pos < 256; pos++) div[pos] = l;
        for (; pos < ramp + 8 && pos < 256; pos++) { div[pos] = l; l *= step; }
    }
    for (int sf = 0; sf < kSub; sf++) {
        float sum = 0.0f;
        for (int k = 0; k < 8; k++) sum += div[sf * 8 + k];
        out[sf] = sum / 8.0f;
    }
}

// CalcCurveEarlyMismatchScore (atrac3denc.cpp:259-297)
template <class TCurve>
float MismatchScore(const float* gain, float target, const TCurve& pts, const float* level, const float* interp)
{
    if (target <= 1e-9f)
        return 0.0f;
    float div[kSub];
    SubframeDivisors(pts, level, interp, div);
    uint32_t last = 0;
    for (const auto& p : pts) last = std::max(last, p.Loc);
    const uint32_t n = std::min<uint32_t>(kSub, std::max<uint32_t>(3, last + 3));
    const float eps = 1e-9f;
    float fit = 0.0f;
    for (uint32_t sf = 0; sf < n; sf++) {
        const float mod = gain[sf] / std::max(div[sf], eps);
        const float e = std::log2(std::max(mod, eps) / std::max(target, eps));
        fit += e * e;
    }
    fit /= n;
    float leak = 0.0f, weight = 0.0f;
    for (uint32_t sf = 0; sf + 1 < n; sf++) {
        const float a = std::log2(std::max(div[sf], eps));
        const float b = std::log2(std::max(div[sf + 1], eps));
        const float d = b - a;
        const float w = 0.5f * (gain[sf] + gain[sf + 1]);
        leak += d * d * w;
        weight += w;
    }
    if (weight > eps)
        leak /= weight;
    return fit + 0.25f * leak;
}

} // namespace

// One band of CreateSubbandInfo (atrac3denc.cpp:311-578).  env: gain[32] low[32] high[32]; stat: hfr, the device's mean
// envelope and plateau target (checked against the host's), next_level;
// cur: the frame's 256 band samples; devCurve: the 16-byte curve record the device encoded with (bands 0..2).
void TGainTraceWriter::Band(int channel, int band, const float* env, const float* stat, const float* cur, const uint8_t* devCurve)
{
    std::ostream& os = *Out;
    TBandState& st = State[(size_t)channel * 4 + band];
    const float* gain = env;
    const float hfr = stat[0], nextLevel = stat[3];
    os << "      - band: " << band << "\n";
    TCurve curve;
    bool analysed = false;
    if (hfr < 0.05f) {
        os << TFix{4} << "        skip: low_hfr  # high_freq_ratio " << hfr << " < threshold\n";
        st.LastLevel = 0.0f;
    } else {
        analysed = true;
        float mean = 0.0f;
        for (int i = 0; i < kSub; i++) mean += gain[i];
        mean /= static_cast<float>(kSub);
        const float before = st.LastHpfEnergy;
        st.LastHpfEnergy = mean;
        const float hpfRatio = (mean > 1e-9f && before > 1e-9f) ? (before / mean) : 1.0f;

        os << "        pcm_qmf:  # 256 raw QMF samples, non-modulated, non-windowed\n          ";
        Sequence(os, cur, 256, 6);
        os << "\n";
        float eStored = 0.0f, eCur = 0.0f;
        for (int i = 0; i < 256; i++) {
            eStored += st.StoredHalf[i] * st.StoredHalf[i];
            eCur += cur[i] * cur[i];
        }
        const float overlapRatio = eStored / (eCur + 1e-9f);
        const float minScore = 1.9f * std::min(1.5f, std::max(1.0f, hpfRatio));
        os << TFix{4}
           << "        high_freq_ratio: " << hfr << "\n"
           << "        overlap_ratio: " << overlapRatio << "  # prev_E/cur_E full-band; >1 means prev frame louder\n"
           << "        hpf_overlap_ratio: " << hpfRatio << "  # prev_HPF/cur_HPF; used for transient suppression decisions\n"
           << "        dynamic_min_score: " << minScore << "\n"
           << "        next_level: " << nextLevel << "\n"
           << "        gain: ";
        Sequence(os, gain, kSub, 4);
        os << "  # 32 subframe RMS values\n";

        const float prevTarget = st.LastTarget;
        curve = BuildCurve(gain, env + 32, env + 64, st, minScore);
        const float curTarget = st.LastTarget;
        if (std::memcmp(&curTarget, &stat[2], sizeof(float)) != 0 || std::memcmp(&mean, &stat[1], sizeof(float)) != 0)
            throw std::runtime_error("atde_b200: gain trace: the host's envelope statistics differ from the device's");
        if (curve.empty()) {
            os << "        skip: no_curve\n";
            analysed = false;                               // `continue`: no curve_final line, no curve
        } else {
            os << "        curve_raw:\n";
            for (const TPoint& p : curve) os << "          - {level: " << p.Level << ", loc: " << p.Loc << "}\n";
            float peak = 0.0f;
            for (int i = 0; i < kSub; i++) peak = std::max(peak, gain[i]);
            if (peak < 1e-4f) {
                os << TFix{6} << "        skip: below_min_signal  # maxGain " << peak << "\n";
                curve.clear();
            }
            if (hfr < 0.3f) {
                os << "        skip: amplify_low_hfr\n";
                curve.clear();
            }
            os << TFix{4} << "        max_gain: " << peak << "\n";
            if (band >= 3) {
                os << "        skip: band_ge_3  # inaudible HF; gain modulation disabled\n";
                curve.clear();
            } else {
                // explicit point 0 (atrac3denc.cpp:457-554)
                const TCurve raw = curve;
                bool changed = false, valid = false;
                float mod = 0.0f;
                if (!curve.empty() && curve[0].Loc > 0) {
                    float sum = 0.0f;
                    for (uint32_t sf = 0; sf < curve[0].Loc; sf++) sum += gain[sf];
                    mod = (sum / curve[0].Loc) / Level[curve[0].Level];
                    valid = true;
                } else if (curve.empty()) {
                    float sum = 0.0f;
                    for (int i = 0; i < kSub; i++) sum += gain[i];
                    mod = sum / kSub;
                    valid = true;
                }
                os << TFix{6} << "        prev_target: " << prevTarget << "\n"
                   << "        hpf_rms_next_mod: " << mod << "\n";
                const bool usable = valid && prevTarget > 1e-6f && mod > 1e-6f;
                if (usable) {
                    const uint16_t l0 = (uint16_t)LevelOfRatio(prevTarget / mod);
                    os << "        point0_level: " << l0 << "  # RelationToIdx(prev_target/hpf_rms_next_mod)\n";
                    auto at = std::find_if(curve.begin(), curve.end(), [](const TPoint& p) { return p.Loc == 0; });
                    if (at != curve.end()) {
                        if (at->Level != l0) { at->Level = l0; changed = true; }
                    } else if (l0 != kNeutral || !curve.empty()) {
                        curve.insert(curve.begin(), TPoint{l0, 0});
                        changed = true;
                    }
                }
                if (changed) {
                    const float scoreRaw = MismatchScore(gain, curTarget, raw, Level, Interp);
                    const float scoreNew = MismatchScore(gain, curTarget, curve, Level, Interp);
                    bool keepForEdge = false;
                    float errRaw = 0.0f, errNew = 0.0f;
                    if (usable) {
                        const float want = std::min(std::max(prevTarget / mod, Level[15]), Level[0]);      // LimitRel
                        const float sRaw = Level[raw.empty() ? kNeutral : raw[0].Level];
                        const float sNew = Level[curve.empty() ? kNeutral : curve[0].Level];
                        const float eps = 1e-9f;
                        errRaw = std::abs(std::log2(std::max(sRaw, eps) / std::max(want, eps)));
                        errNew = std::abs(std::log2(std::max(sNew, eps) / std::max(want, eps)));
                        keepForEdge = errNew + 0.20f < errRaw;
                        os << TFix{6} << "        point0_guard_boundary_err_before: " << errRaw << "\n"
                           << "        point0_guard_boundary_err_after: " << errNew << "\n";
                    }
                    if (!keepForEdge && scoreNew > scoreRaw * (1.0f + 0.02f)) {
                        curve = raw;
                        os << TFix{6} << "        point0_guard: reverted  # score_after " << scoreNew
                           << " > score_before " << scoreRaw << "\n";
                    } else {
                        os << TFix{6} << "        point0_guard: kept  # score_before " << scoreRaw << ", score_after " << scoreNew;
                        if (keepForEdge)
                            os << ", boundary_err_before " << errRaw << ", boundary_err_after " << errNew;
                        os << "\n";
                    }
                }
            }
            if (curve.size() >= 2 && curve[0].Loc == 0 && curve[0].Level == curve[1].Level)
                curve.erase(curve.begin());
            os << "        curve_final:\n";
            for (const TPoint& p : curve) os << "          - {level: " << p.Level << ", loc: " << p.Loc << "}\n";
        }
    }
    if (!analysed)
        curve.clear();
    if (devCurve) {
        // 16-byte record: n, level[7], loc[7], pad (include/atde_b200.h: ATDE_TAP_CURVES)
        bool same = devCurve[0] == curve.size();
        for (size_t i = 0; same && i < curve.size(); i++)
            same = devCurve[1 + i] == curve[i].Level && devCurve[8 + i] == curve[i].Loc;
        if (!same)
            throw std::runtime_error("atde_b200: gain trace: the host's curve differs from the one the device encoded with (frame "
                                     + std::to_string(FrameNum) + ", channel " + std::to_string(channel) + ", band "
                                     + std::to_string(band) + ")");
    }
    // what TAtrac3MDCT::Mdct leaves in PcmBuffer.GetFirst() for the next frame (atrac3denc.cpp:33-58 with
    // TGainProcessor::Modulate, gain_processor.h:87-121): the frame divided by its curve, windowed
    float mod[256];
    std::memcpy(mod, cur, sizeof(mod));
    uint32_t pos = 0;
    for (size_t i = 0; i < curve.size(); i++) {
        const uint32_t ramp = curve[i].Loc << 3;
        float l = Level[curve[i].Level];
        const float step = Interp[(i + 1 < curve.size() ? (int)curve[i + 1].Level : kNeutral) - (int)curve[i].Level + 15];
        for (; pos < ramp; pos++) mod[pos] /= l;
        for (; pos < ramp + 8; pos++) { mod[pos] /= l; l *= step; }
    }
    for (int i = 0; i < 256; i++) st.StoredHalf[i] = Window[i] * mod[i];
}

void TGainTraceWriter::AppendBatch(atde_encoder* enc, int64_t outFrames)
{
    if (!Out || outFrames <= 0)
        return;
    std::ostream& os = *Out;
    const size_t F = (size_t)outFrames, C = (size_t)Channels;
    const size_t BL = 128 + 256 * (F + 1);
    std::vector<float> bands, env, stat, scale;
    std::vector<uint8_t> curves;
    if (GainControl) {
        bands.resize(C * 4 * BL);
        env.resize(C * 4 * F * 96);
        stat.resize(C * 4 * F * 4);
        scale.resize(F * C * 16);
        curves.resize(C * 4 * F * 16);
        Need(atde_debug_tap(enc, ATDE_TAP_BANDS, bands.data(), bands.size() * sizeof(float)), "band samples");
        Need(atde_debug_tap(enc, ATDE_TAP_TRACE_GAIN, env.data(), env.size() * sizeof(float)), "envelope");
        Need(atde_debug_tap(enc, ATDE_TAP_TRACE_STAT, stat.data(), stat.size() * sizeof(float)), "envelope statistics");
        Need(atde_debug_tap(enc, ATDE_TAP_GSCALE, scale.data(), scale.size() * sizeof(float)), "energy scales");
        Need(atde_debug_tap(enc, ATDE_TAP_CURVES, curves.data(), curves.size()), "curves");
    }
    for (size_t f = 0; f < F; f++, FrameNum++) {
        const float seconds = static_cast<float>(FrameNum) * 1024u / 44100.0f;
        os << "---\nframe: " << FrameNum << "\n" << TFix{3} << "time: " << seconds << "  # seconds\n" << "channels:\n";
        if (!GainControl)
            continue;
        for (size_t c = 0; c < C; c++) {
            os << "  - channel: " << c << "\n" << "    bands:\n";
            for (size_t b = 0; b < 4; b++) {
                const size_t item = (c * 4 + b) * F + f;
                Band((int)c, (int)b, &env[item * 96], &stat[item * 4], &bands[(c * 4 + b) * BL + 128 + 256 * f],
                     b < 3 ? &curves[item * 16] : nullptr);
            }
            os << TFix{6} << "    gain_energy_scale:\n";
            for (size_t b = 0; b < 4; b++) {
                const float* s = &scale[((f * C + c) * 4 + b) * 4];
                os << "      - {band: " << b << ", prev_half: " << s[0] << ", cur_half: " << s[1] << ", frame: " << s[2]
                   << ", next_overlap: " << s[3] << "}\n";
            }
        }
    }
}

} // namespace NAtracDEnc
