/*
 * llr_dev.h — device functions of the soft-decision map shared by the demodulator kernel and the
 * stand-alone fsk_rx_filt_to_llrs entry point ([U mpdecode_core.c::FskDemod / Somap / max_star0 /
 * bessi0]).  Same IEEE operation order as the CPU evaluation.
 */
#ifndef PIRIP_LLR_DEV_H
#define PIRIP_LLR_DEV_H

#include "exact_math.h"

/* log I0(x), Numerical Recipes polynomial in the log domain ([U mpdecode_core.c::bessi0]) */
__device__ __forceinline__ float logbessi0(float x) {
    const float ax = fabsf(x);
    double y, ans;
    if ((double)ax < 3.75) {
        y = (double)x / 3.75;
        y *= y;
        ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
        return (float)em_log(ans);
    }
    y = 3.75 / (double)ax;
    ans = 0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2)))))));
    return (float)(((double)ax - 0.5 * em_log((double)ax)) + em_log(ans));
}

/* linear-log-MAP max* ([U mpdecode_core.c::max_star0]) */
__device__ __forceinline__ float max_star0(float delta1, float delta2) {
    const double AJIAN = -0.24904163195436, TJIAN = 2.50681740420944;
    const float diff = delta2 - delta1;
    if ((double)diff > TJIAN) return delta2;
    if ((double)diff < -TJIAN) return delta1;
    if (diff > 0) return (float)((double)delta2 + AJIAN * ((double)diff - TJIAN));
    return (float)((double)delta1 - AJIAN * ((double)diff + TJIAN));
}


/* LLRs of one symbol from its M filter magnitudes (tone-major stride `stride`); llr > 0 means bit 0 */
__device__ __forceinline__ void symbol_llrs(const float *rx_filt, int stride, int M, float v_est, float SNRest, float *out) {
    const int bps = M == 2 ? 1 : 2;
    const float scale = 2.0f * SNRest;
    float num[2] = { -1000000.0f, -1000000.0f }, den[2] = { -1000000.0f, -1000000.0f };
    for (int m = 0; m < M; ++m) {
        const float yr = rx_filt[m * stride];
        const float env = (float)sqrt((double)(yr * yr / (v_est * v_est)));
        const float metric = logbessi0(scale * env);
        int mask = 1 << (bps - 1);
        for (int k = 0; k < bps; ++k) {
            if (mask & m)
                num[k] = max_star0(num[k], metric);
            else
                den[k] = max_star0(den[k], metric);
            mask >>= 1;
        }
    }
    for (int k = 0; k < bps; ++k) out[k] = -(num[k] - den[k]);
}

#endif
