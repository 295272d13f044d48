/*
 * pirip_csdr.h — C-ABI of the sample-conditioning stage (csdr) of pirip's receive chain:
 *     csdr convert_u8_f | csdr fir_decimate_cc 45 | csdr convert_f_s16      [R README.md:109,162]
 *
 * Layer 1 keeps the libcsdr function names and argument meaning (ha7ilm/csdr libcsdr.h,
 * un-vendored dependency cloned by [R build_csdr.sh:4-5]) so code
 * that links libcsdr — rtl_fsk does, via -DCSDR_BUILD_DIR [R build_rtlsdr.sh:9] — can link
 * this library instead.  The FIR runs on the GPU; the tap design is host code.
 *
 * Layer 2 is the batched, fused form used by the B200 receive chain: u8 IQ in HBM ->
 * convert + FIR + decimate (+ s16 requantise) in one kernel, many streams per launch.
 *
 * Results are bit-identical to the scalar algorithm (same IEEE operations in tap order).
 * No CPU fallback: compute entry points fail when no CUDA device is usable.
 */
#ifndef PIRIP_CSDR_H
#define PIRIP_CSDR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct complexf_s {
    float i;
    float q;
} complexf;

/* ---- layer 1: libcsdr names ------------------------------------------------------------ */

/* replaces libcsdr.h::convert_u8_f — out[i] = in[i]/(UCHAR_MAX/2.0) - 1.0 (host) */
void convert_u8_f(unsigned char *input, float *output, int input_size);
/* replaces libcsdr.h::convert_f_s16 — out[i] = (short)(in[i]*SHRT_MAX) (host) */
void convert_f_s16(float *input, short *output, int input_size);
/* replaces libcsdr.h::firdes_filter_len */
int firdes_filter_len(float transition_bw);
/* replaces libcsdr.h::firdes_lowpass_f with window = WINDOW_HAMMING (the CLI default) */
void firdes_lowpass_f(float *output, int length, float cutoff_rate, int window);
/* replaces libcsdr.h::fir_decimate_cc — returns the number of outputs; the FIR runs on the GPU */
int fir_decimate_cc(complexf *input, complexf *output, int input_size, int decimation, float *taps, int taps_length);

/* ---- layer 2: batched fused decimator -------------------------------------------------- */

typedef struct pirip_decim_plan pirip_decim_plan;

/* taps: host array of taps_length floats (e.g. from firdes_lowpass_f).  NULL on failure. */
pirip_decim_plan *pirip_decim_plan_create(int decimation, const float *taps, int taps_length, int device);
void pirip_decim_plan_destroy(pirip_decim_plan *plan);
/* outputs produced for a stream of n input samples: every k with k*decimation + taps_length <= n */
size_t pirip_decim_out_count(const pirip_decim_plan *plan, size_t n_input_samples);

/* n_streams u8 IQ streams already in device memory (stream s at d_iq_u8 + s*in_stride_bytes,
 * samples_per_stream complex samples each).  Writes complex float outputs to d_out_f32 and/or
 * truncated int16 pairs to d_out_s16 (either may be NULL), out_stride complex samples apart.
 * Asynchronous on `stream` (cudaStream_t).  Returns 0 or a negative error code. */
int pirip_decim_u8_batch_dev(pirip_decim_plan *plan, const uint8_t *d_iq_u8, size_t n_streams,
                             size_t samples_per_stream, size_t in_stride_bytes, complexf *d_out_f32,
                             int16_t *d_out_s16, size_t out_stride, void *stream);
/* same with host buffers (copies inside, synchronous) */
int pirip_decim_u8_batch(pirip_decim_plan *plan, const uint8_t *iq_u8, size_t n_streams, size_t samples_per_stream,
                         complexf *out_f32, int16_t *out_s16);

#ifdef __cplusplus
}
#endif
#endif /* PIRIP_CSDR_H */
