/*
 * oracle.h — CPU ORACLE for the pirip IQ->bits hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under pirip_b200/, include/ or tools/ may include, link or call
 * anything in oracle/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline.
 *
 * PARITY UNPINNED.  pirip (/root/reference) holds none of the arithmetic of
 * this path; it git-clones drowe67/codec2 (no pin, build_codec2.sh:3-5),
 * ha7ilm/csdr (no pin, build_csdr.sh:4-5) and drowe67/librtlsdr@development
 * (build_rtlsdr.sh:4-6) at build time and none of them is in this container
 * (SURVEY.md section 0 / 8c).  This oracle is a scalar, single-threaded C
 * restatement of the PUBLISHED algorithms of those projects as of the pirip
 * era (2020-07 .. 2021-10), written from the description in SURVEY.md
 * section 8a / Appendix A, anchored on pirip's own call sites
 * (README.md:105,109,184,200-208; test/loopback_rtl_sdr.sh:16;
 * tx/rpitx_fsk.cpp:34-39,75-83,313-336,395; tx/frame_repeater.c:55-88).
 * The reference's own tests hold no numeric vector for this path; what can be
 * pinned independently of the reference (CRC-16/CCITT-FALSE check value, UW
 * pattern, LCG/glibc test frames, FIR tap properties, noncoherent FSK BER
 * theory) is pinned in tests/test_oracle_*.py.
 *
 * Each function names the upstream unit it restates as
 * [U file::function] (SURVEY.md citation convention) and the pirip call
 * site [R file:line] that puts it on the path.
 */
#ifndef PIRIP_ORACLE_H
#define PIRIP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float real, imag; } orc_comp;

/* ------------------------------------------------------------------ FFT */
/* [U kiss_fft.c] mixed radix 4/2 decimation-in-time complex FFT, float */
typedef struct orc_fft orc_fft;
orc_fft *orc_fft_alloc(int nfft);
void orc_fft_free(orc_fft *cfg);
void orc_fft_run(const orc_fft *cfg, const orc_comp *in, orc_comp *out);
/* copy of the float twiddle table (nfft entries) for tests */
void orc_fft_twiddles(const orc_fft *cfg, orc_comp *out);

/* ------------------------------------------------------------------ FSK */
#define ORC_MODE_M_MAX 4

/* [U fsk.h::struct FSK] */
struct orc_fsk {
    int Ndft, Fs, N, Rs, Ts, Nmem, P, Nsym, Nbits, f1_tx, tone_spacing, mode;
    float tc;
    int est_min, est_max, est_space;
    float *hann_table;
    float *Sf;
    orc_comp phi_c[ORC_MODE_M_MAX];
    orc_comp *f_dc;
    orc_fft *fft_cfg;
    float norm_rx_timing;
    orc_comp tx_phase_c;
    float EbNodB;
    float f_est[ORC_MODE_M_MAX];
    float f2_est[ORC_MODE_M_MAX];
    int freq_est_type;
    float ppm;
    float SNRest, v_est, rx_sig_pow, rx_nse_pow;
    int nin;
    int burst_mode;
    int lock_nin;
    /* the fields of struct MODEM_STATS fsk_demod_core writes */
    float stats_snr_est, stats_foff, stats_rx_timing, stats_clock_offset;
    int normalise_eye;
};

struct orc_fsk *orc_fsk_create(int Fs, int Rs, int M, int tx_f1, int tx_fs);
struct orc_fsk *orc_fsk_create_hbr(int Fs, int Rs, int M, int P, int Nsym, int f1_tx, int tone_spacing);
void orc_fsk_destroy(struct orc_fsk *fsk);
uint32_t orc_fsk_nin(struct orc_fsk *fsk);
void orc_fsk_set_freq_est_limits(struct orc_fsk *fsk, int est_min, int est_max);
void orc_fsk_set_freq_est_alg(struct orc_fsk *fsk, int est_type);
void orc_fsk_clear_estimators(struct orc_fsk *fsk);
void orc_fsk_enable_burst_mode(struct orc_fsk *fsk);
void orc_fsk_demod(struct orc_fsk *fsk, uint8_t rx_bits[], const orc_comp fsk_in[]);
void orc_fsk_demod_sd(struct orc_fsk *fsk, float rx_filt[], const orc_comp fsk_in[]);
/* [U fsk.c::fsk_mod_c] continuous-phase M-FSK modulator, complex output */
void orc_fsk_mod_c(struct orc_fsk *fsk, orc_comp fsk_out[], const uint8_t tx_bits[], int nbits);

/* per-frame record of every float the demod leaves in struct FSK; the GPU path
   writes the same record so tests can compare float for float */
struct orc_frame_stats {
    float f_est[4];
    float f2_est[4];
    float norm_rx_timing;
    float ppm;
    float SNRest;
    float v_est;
    float rx_sig_pow;
    float rx_nse_pow;
    float EbNodB;
    float snr_est;
    float foff;
    float rx_timing;
    int32_t nin_next;
    int32_t sample_offset; /* first input sample of this frame */
};

/* input sample formats of the fsk_demod CLI [U fsk_demod.c main loop] */
enum { ORC_FMT_U8 = 0, ORC_FMT_S16 = 1, ORC_FMT_F32 = 2 };

struct orc_demod_cfg {
    int Fs, Rs, M, P, Nsym;
    int f1_tx, tone_spacing;
    int est_min, est_max;   /* used when have_limits */
    int have_limits;
    int freq_est_type;      /* 0 peak, 1 mask */
    int fmt;                /* ORC_FMT_* , complex interleaved */
    int soft;               /* 0: hard bits; 1: also LLRs and rx_filt */
};

/* Whole-stream demod = the fsk_demod CLI main loop [U fsk_demod.c] as used at
   [R README.md:105,109], [R test/loopback_rtl_sdr.sh:16]: read fsk_nin()
   complex samples, convert, demodulate, emit Nbits; stop on a short read.
   bits_out: max_frames*Nbits bytes (may be NULL); llr_out: max_frames*Nbits
   floats (NULL unless soft); rx_filt_out: max_frames*M*Nsym floats (or NULL);
   stats_out: max_frames records (or NULL).  Returns the number of frames. */
size_t orc_fsk_demod_stream(const struct orc_demod_cfg *cfg, const void *iq, size_t nsamples,
                            uint8_t *bits_out, float *llr_out, float *rx_filt_out,
                            struct orc_frame_stats *stats_out, size_t max_frames);

/* ------------------------------------------------------------------ csdr */
/* [U libcsdr.c] as used at [R README.md:109,162] */
void orc_convert_u8_f(const uint8_t *in, float *out, size_t n);
void orc_convert_f_s16(const float *in, int16_t *out, size_t n);
int orc_firdes_filter_len(float transition_bw);
void orc_firdes_lowpass_f(float *taps, int length, float cutoff_rate);
/* block function: returns number of outputs written */
size_t orc_fir_decimate_cc(const orc_comp *in, orc_comp *out, size_t n, int decimation,
                           const float *taps, int taps_length);
/* the `csdr convert_u8_f | csdr fir_decimate_cc D | csdr convert_f_s16` chain over a
   whole u8 IQ stream; out_s16 gets 2*n_out int16 (may be NULL), out_f gets n_out
   complex floats (may be NULL).  Returns n_out. */
size_t orc_csdr_decimate_stream(const uint8_t *iq_u8, size_t nsamples, int decimation,
                                float transition_bw, orc_comp *out_f, int16_t *out_s16);

/* ------------------------------------------------------------------ LLR + LDPC */
/* [U mpdecode_core.c::fsk_rx_filt_to_llrs] */
void orc_fsk_rx_filt_to_llrs(float llr[], const float rx_filt[], float v_est, float SNRest, int M, int nsyms);

/* [U mpdecode_core.h::struct LDPC] — the table layout of codec2's H_*.c files */
struct orc_ldpc {
    int max_iter;
    int CodeLength;
    int NumberParityBits;
    int NumberRowsHcols;
    int max_row_weight;
    int max_col_weight;
    const uint16_t *H_rows;
    const uint16_t *H_cols;
};
/* [U mpdecode_core.c::encode] RA accumulate */
void orc_ldpc_encode(const struct orc_ldpc *ldpc, const uint8_t ibits[], uint8_t pbits[]);
/* [U mpdecode_core.c::run_ldpc_decoder / SumProduct]; returns iterations used */
int orc_run_ldpc_decoder(const struct orc_ldpc *ldpc, uint8_t out_char[], const float input[], int *parityCheckCount);
float orc_phi0(float x);

/* ------------------------------------------------------------------ framing */
/* [U freedv_api.c] as declared at [R tx/rpitx_fsk.cpp:34-39] */
uint16_t orc_gen_crc16(const uint8_t *data, int length);
void orc_pack(uint8_t *bytes, const uint8_t *bits, int nbits);
void orc_unpack(uint8_t *bits, const uint8_t *bytes, int nbits);
int orc_check_crc16_unpacked(const uint8_t *bits, int nbits);
void orc_uw_bits(uint8_t uw[32]);
/* [U ofdm.c::ofdm_generate_payload_data_bits] [R tx/rpitx_fsk.cpp:367] */
void orc_generate_payload_data_bits(uint8_t bits[], int n);
/* [U fsk_get_test_bits.c] [R README.md:101] */
void orc_test_bits_frame(uint8_t bits[], int frame_len);
/* [R tx/rpitx_fsk.cpp:75-83] */
void orc_insert_crc(uint8_t data_bits[], int data_bits_per_frame);
/* [U freedv_fsk.c::freedv_tx_fsk_ldpc_framer] [R tx/rpitx_fsk.cpp:395] */
void orc_fsk_ldpc_framer(const struct orc_ldpc *ldpc, uint8_t frame[], const uint8_t payload[]);

#define ORC_RX_TRIAL_SYNC 0x1
#define ORC_RX_SYNC 0x2
#define ORC_RX_BITS 0x4
#define ORC_RX_BIT_ERRORS 0x8

/* [U freedv_fsk.c::freedv_rx_fsk_ldpc_data] deframer state */
struct orc_deframer {
    struct orc_ldpc ldpc;
    int bits_per_frame;  /* 32 + CodeLength */
    int data_bits;       /* CodeLength - NumberParityBits */
    float *twoframes_llr;
    uint8_t *twoframes_hard;
    float *frame_llr;
    int frame_llr_size;
    int frame_llr_nbits;
    int state;
    int best_location;
    int baduw;
    int thresh1, thresh2, baduw_thresh;
    /* last processed frame, for the rtl_fsk -v log line [R README.md:200] */
    int last_uw_errors, last_iter, last_pcc;
};
struct orc_deframer *orc_deframer_create(const struct orc_ldpc *ldpc, int max_nbits_per_call);
void orc_deframer_destroy(struct orc_deframer *d);
/* push nbits new LLRs; returns rx_status; payload_bits (data_bits bytes, 1 bit/byte)
   is written when ORC_RX_BITS is set */
int orc_deframer_push(struct orc_deframer *d, const float llr[], int nbits, uint8_t payload_bits[]);
void orc_deframer_info(const struct orc_deframer *d, int info[6]);

/* ------------------------------------------------------------------ misc */
/* which math backend the oracle was built with: 0 = exact_math.h, 1 = glibc libm */
int orc_math_backend(void);
float orc_cosf(float x);
float orc_sinf(float x);
float orc_atan2f(float y, float x);
float orc_expf(float x);
float orc_logf(float x);
float orc_log10f(float x);

#ifdef __cplusplus
}
#endif
#endif
