// qmf_dev.cuh — register-tiled 48-tap half-band split, TQmf::Analysis (reference src/qmf/qmf.h:47-64).
//
// One task = R consecutive output pairs j0 .. j0 + R - 1 of one source array:
//     lower_sum = sum_{i=0..23} W[2i]   * src[2j + 49 - 2i]        (sequential, i ascending)
//     upper_sum = sum_{i=0..23} W[2i+1] * src[2j + 48 - 2i]
//     upper = lower_sum - upper_sum;  lower = lower_sum + upper_sum              (qmf.h:60-62)
// Tap pair i multiplies the ALIGNED sample pair p = j + 24 - i = (src[2p], src[2p+1]) by (W[2i+1], W[2i]): the two
// running sums advance with one packed multiply and one packed add per tap pair (Blackwell FMUL2 / FFMA2, IEEE-rn,
// un-fused: see atde_cuda.h:add2).  The R outputs of a task share their R + 23 sample pairs, loaded once.
//
// This variant reads an UNPADDED array with 64-bit loads.  With an odd R the 2R-float stride between the tasks of
// consecutive lanes visits all bank pairs once per half warp, so the loads are conflict-free without padding.
#pragma once
#include "atde_cuda.h"

namespace atde {

// cw: tap pairs (W[2i+1], W[2i]), i = 0..23, 8-byte aligned (shared or constant memory)
template <int R>
ATDE_D void qmf_task64(const float* src, const float* cw, int j0, f32x2 one, float* lower, float* upper)
{
    f32x2 xp[R + 23];                                  // sample pairs j0 + 1 .. j0 + R + 23
    const float2* q = reinterpret_cast<const float2*>(src) + (j0 + 1);
#pragma unroll
    for (int u = 0; u < R + 23; u++) {
        const float2 v = q[u];
        xp[u].x = v.x; xp[u].y = v.y;
    }
    f32x2 acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { acc[r].x = 0.0f; acc[r].y = 0.0f; }
#pragma unroll
    for (int i = 0; i < 24; i++) {
        const float2 c = *reinterpret_cast<const float2*>(cw + 2 * i);
        f32x2 cc;
        cc.x = c.x; cc.y = c.y;
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = add2(acc[r], mul2(cc, xp[r + 23 - i]), one);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {                      // acc.x = upper sum, acc.y = lower sum
        upper[r] = fsub(acc[r].y, acc[r].x);
        lower[r] = fadd(acc[r].y, acc[r].x);
    }
}

} // namespace atde
