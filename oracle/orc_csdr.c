/*
 * orc_csdr.c — CPU oracle: csdr sample conditioning.  TEST INFRASTRUCTURE (see oracle.h).
 * PARITY UNPINNED: restatement of ha7ilm/csdr libcsdr.c (un-vendored, un-pinned
 * dependency of pirip, build_csdr.sh:4-5) as used by
 *   csdr convert_u8_f | csdr fir_decimate_cc 45 | csdr convert_f_s16   [R README.md:109,162]
 * SURVEY.md 8a rows A1-A4, App. A.5.
 */
#include "oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>

/* csdr defines PI as a float constant */
#define CSDR_PI ((float)3.14159265358979323846)

/* [U libcsdr.c::convert_u8_f] */
void orc_convert_u8_f(const uint8_t *in, float *out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = ((float)in[i]) / (UCHAR_MAX / 2.0) - 1.0;
}

/* [U libcsdr.c::convert_f_s16] C truncation toward zero */
void orc_convert_f_s16(const float *in, int16_t *out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = (int16_t)(in[i] * SHRT_MAX);
}

/* [U libcsdr.c::firdes_filter_len].  NOTE: the C expression is 4.0/transition_bw
   with a FLOAT transition_bw; for the CLI default 0.05f that is
   4.0/0.0500000007... = 79.9999988 -> (int) 79, already odd -> 79 taps
   (SURVEY.md 8a A2 derives 81 from exact decimal arithmetic; DESIGN.md discusses). */
int orc_firdes_filter_len(float transition_bw) {
    int result = 4.0 / transition_bw;
    if (result % 2 == 0) result++;
    return result;
}

/* [U libcsdr.c::firdes_wkernel_hamming] */
static float wkernel_hamming(float rate) {
    rate = 0.5 + rate / 2;
    return 0.54 - 0.46 * cos(2 * CSDR_PI * rate);
}

/* [U libcsdr.c::firdes_lowpass_f + normalize_fir_f], Hamming window */
void orc_firdes_lowpass_f(float *taps, int length, float cutoff_rate) {
    int middle = length / 2;
    taps[middle] = 2 * CSDR_PI * cutoff_rate * wkernel_hamming(0);
    for (int i = 1; i <= middle; ++i)
        taps[middle - i] = taps[middle + i] = (sin(2 * CSDR_PI * cutoff_rate * i) / i) * wkernel_hamming((float)i / middle);
    float sum = 0;
    for (int i = 0; i < length; ++i) sum += taps[i];
    for (int i = 0; i < length; ++i) taps[i] = taps[i] / sum;
}

/* [U libcsdr.c::fir_decimate_cc] real taps on a complex stream, accumulate in tap order */
size_t orc_fir_decimate_cc(const orc_comp *in, orc_comp *out, size_t n, int decimation,
                           const float *taps, int taps_length) {
    size_t oi = 0;
    for (size_t i = 0; i < n; i += (size_t)decimation) {
        if (i + (size_t)taps_length > n) break;
        float acci = 0;
        float accq = 0;
        for (int ti = 0; ti < taps_length; ++ti) acci += in[i + ti].real * taps[ti];
        for (int ti = 0; ti < taps_length; ++ti) accq += in[i + ti].imag * taps[ti];
        out[oi].real = acci;
        out[oi].imag = accq;
        oi++;
    }
    return oi;
}

/* The csdr CLI keeps the unconsumed tail of its buffer and refills, i.e. one
   continuous causal FIR anchored at stream sample 0 [U csdr.c fir_decimate_cc main
   loop].  The CLI drops the last partial buffer at EOF (buffer-size dependent);
   this stream form emits every output whose taps fit in the stream. */
size_t orc_csdr_decimate_stream(const uint8_t *iq_u8, size_t nsamples, int decimation,
                                float transition_bw, orc_comp *out_f, int16_t *out_s16) {
    int L = orc_firdes_filter_len(transition_bw);
    float *taps = (float *)malloc(sizeof(float) * (size_t)L);
    orc_firdes_lowpass_f(taps, L, 0.5 / (float)decimation);
    orc_comp *x = (orc_comp *)malloc(sizeof(orc_comp) * (nsamples ? nsamples : 1));
    orc_convert_u8_f(iq_u8, (float *)x, 2 * nsamples);
    size_t nout_max = nsamples / (size_t)decimation + 1;
    orc_comp *y = out_f ? out_f : (orc_comp *)malloc(sizeof(orc_comp) * nout_max);
    size_t nout = orc_fir_decimate_cc(x, y, nsamples, decimation, taps, L);
    if (out_s16) orc_convert_f_s16((const float *)y, out_s16, 2 * nout);
    if (!out_f) free(y);
    free(x);
    free(taps);
    return nout;
}
