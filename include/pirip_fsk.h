/*
 * pirip_fsk.h — C-ABI of the B200 FSK demodulator (libpirip_b200.so).
 *
 * Two layers:
 *
 *  1. The codec2 `struct FSK` API, symbol for symbol, so the library drops in
 *     where pirip links libcodec2.so ([R tx/Makefile:9]) and where rtl_fsk /
 *     the fsk_demod CLI call it ([R README.md:105,109,114],
 *     [R test/loopback_rtl_sdr.sh:16], [R test/loopback_rtl_fsk.sh:10]).
 *     Each prototype replaces the codec2 src/fsk.h declaration of the same name
 *     (un-vendored dependency cloned by [R build_codec2.sh:3-5]; SURVEY.md 8b
 *     lists the interface).  The demodulator state lives in GPU memory; every
 *     fsk_demod*() call is synchronous from the caller's point of view.
 *
 *  2. A batched extension (not in codec2) for many independent captures:
 *     one persistent CTA per capture runs the whole fsk_demod main loop on the
 *     device.  Per capture it produces exactly what layer 1 produces when fed that
 *     capture from a fresh fsk_create_hbr() by the fsk_demod CLI loop.
 *
 * There is no CPU fallback: every entry point that computes fails (returns a
 * negative code, or aborts for the void codec2 calls, as codec2's asserts do)
 * when no CUDA device is usable.
 */
#ifndef PIRIP_FSK_H
#define PIRIP_FSK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- layer 1: codec2 fsk.h ------------------------------------------------ */

#ifndef __COMP__
#define __COMP__
/* codec2 comp.h */
typedef struct {
    float real;
    float imag;
} COMP;
#endif

#define MODE_2FSK 2
#define MODE_4FSK 4
#define MODE_M_MAX 4
#define FSK_DEFAULT_P 8
#define FSK_DEFAULT_NSYM 50
#define FSK_SCALE 16383

/* codec2 modem_stats.h, the members fsk.c writes */
#define MODEM_STATS_NC_MAX 50
#define MODEM_STATS_NR_MAX 160
#define MODEM_STATS_ET_MAX 8
#define MODEM_STATS_EYE_IND_MAX 160
struct MODEM_STATS {
    int Nc;
    float snr_est;
    COMP rx_symbols[MODEM_STATS_NR_MAX][MODEM_STATS_NC_MAX + 1];
    int nr;
    int sync;
    float foff;
    float rx_timing;
    float clock_offset;
    float sync_metric;
    float rx_eye[MODEM_STATS_ET_MAX][MODEM_STATS_EYE_IND_MAX];
    int neyetr;
    int neyesamp;
    float f_est[MODE_M_MAX];
    float fft_buf[2 * 1024];
    void *fft_cfg;
};

/* codec2 fsk.h::struct FSK — public field names and types kept;
   hann_table / Sf are host mirrors refreshed after every demod call; f_dc and
   fft_cfg of upstream are device-side here and reached through `dev`. */
struct FSK {
    int Ndft;
    int Fs;
    int N;
    int Rs;
    int Ts;
    int Nmem;
    int P;
    int Nsym;
    int Nbits;
    int f1_tx;
    int tone_spacing;
    int mode;
    float tc;
    int est_min;
    int est_max;
    int est_space;
    float *hann_table;
    float *Sf;
    COMP phi_c[MODE_M_MAX];
    COMP *f_dc;    /* NULL: lives on the device */
    void *fft_cfg; /* NULL: the FFT runs inside the kernel */
    float norm_rx_timing;
    COMP tx_phase_c;
    float EbNodB;
    float f_est[MODE_M_MAX];
    float f2_est[MODE_M_MAX];
    int freq_est_type;
    float ppm;
    float SNRest;
    float v_est;
    float rx_sig_pow;
    float rx_nse_pow;
    int nin;
    int burst_mode;
    int lock_nin;
    struct MODEM_STATS *stats;
    int normalise_eye;
    void *dev; /* opaque: device state, plan and stream */
};

/* replaces codec2 src/fsk.h::fsk_create */
struct FSK *fsk_create(int Fs, int Rs, int M, int tx_f1, int tx_fs);
/* replaces codec2 src/fsk.h::fsk_create_hbr */
struct FSK *fsk_create_hbr(int Fs, int Rs, int M, int P, int Nsym, int f1_tx, int tone_spacing);
/* replaces codec2 src/fsk.h::fsk_destroy */
void fsk_destroy(struct FSK *fsk);
/* replaces codec2 src/fsk.h::fsk_nin */
uint32_t fsk_nin(struct FSK *fsk);
/* replaces codec2 src/fsk.h::fsk_demod — rx_bits: Nbits bytes, one bit per byte */
void fsk_demod(struct FSK *fsk, uint8_t rx_bits[], COMP fsk_in[]);
/* replaces codec2 src/fsk.h::fsk_demod_sd — rx_filt: M*Nsym floats, tone-major */
void fsk_demod_sd(struct FSK *fsk, float rx_filt[], COMP fsk_in[]);
/* replaces codec2 src/fsk.h::fsk_set_freq_est_limits */
void fsk_set_freq_est_limits(struct FSK *fsk, int est_min, int est_max);
/* replaces codec2 src/fsk.h::fsk_set_freq_est_alg */
void fsk_set_freq_est_alg(struct FSK *fsk, int est_type);
/* replaces codec2 src/fsk.h::fsk_clear_estimators */
void fsk_clear_estimators(struct FSK *fsk);
/* replaces codec2 src/fsk.h::fsk_enable_burst_mode */
void fsk_enable_burst_mode(struct FSK *fsk);
/* replaces codec2 src/fsk.h::fsk_get_demod_stats */
void fsk_get_demod_stats(struct FSK *fsk, struct MODEM_STATS *stats);
/* replaces codec2 src/fsk.h::fsk_setup_modem_stats */
void fsk_setup_modem_stats(struct FSK *fsk, struct MODEM_STATS *stats);
/* replaces codec2 src/fsk.h::fsk_stats_normalise_eye */
void fsk_stats_normalise_eye(struct FSK *fsk, int normalise_enable);
/* Stream extension: as many whole modem frames as `n_avail` samples hold, at most max_frames, in ONE launch -
 * what a receiver front end that hands over fixed-size chunks ([U rtl_fsk.c] USB callback, [R README.md:172]) calls instead of
 * fsk_demod once per frame.  The struct FSK advances exactly as by that many fsk_demod / fsk_demod_sd calls, outputs are the same
 * floats.  rx_bits [frames][Nbits], rx_filt [frames][mode * Nsym], stats [frames] may each be NULL.  Returns the frames
 * demodulated, *consumed the samples they took (the caller keeps the rest for the next call). */
struct pirip_frame_stats;
int pirip_fsk_demod_frames(struct FSK *fsk, const COMP *in, size_t n_avail, int max_frames, uint8_t *rx_bits, float *rx_filt,
                           struct pirip_frame_stats *stats, size_t *consumed);
/* replaces codec2 src/fsk.h::fsk_mod / fsk_mod_c (Tx side, host only) */
void fsk_mod(struct FSK *fsk, float fsk_out[], uint8_t tx_bits[], int nbits);
void fsk_mod_c(struct FSK *fsk, COMP fsk_out[], uint8_t tx_bits[], int nbits);
/* replaces codec2 src/fsk.h::fsk_mod_ext_vco: per-sample tone frequency (Hz) for an external VCO, Ts values per symbol */
void fsk_mod_ext_vco(struct FSK *fsk, float vco_out[], uint8_t tx_bits[], int nbits);

/* ---- layer 2: batched extension ------------------------------------------- */

/* complex interleaved input sample formats of the fsk_demod CLI
   (-d u8 [R README.md:105], -c s16 [R README.md:109], float for the COMP API) */
enum { PIRIP_FMT_U8 = 0, PIRIP_FMT_S16 = 1, PIRIP_FMT_F32 = 2 };

struct pirip_fsk_cfg {
    int Fs, Rs, M, P, Nsym;
    int f1_tx, tone_spacing;
    int est_min, est_max; /* used when have_limits != 0 (--fsk_lower / --fsk_upper) */
    int have_limits;
    int freq_est_type; /* 0 peak picking, 1 mask (--mask) */
    int fmt;           /* PIRIP_FMT_* */
    int soft;          /* != 0: LLRs (fsk_demod -s) are produced as well */
};

/* every float fsk_demod_core leaves in struct FSK, one record per frame */
struct pirip_frame_stats {
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
    int32_t sample_offset;
};

typedef struct pirip_fsk_plan pirip_fsk_plan;

/* Validate cfg (same conditions codec2 asserts), upload the window / twiddle /
   timing-oscillator tables to `device`.  NULL on failure (see pirip_last_error). */
pirip_fsk_plan *pirip_fsk_plan_create(const struct pirip_fsk_cfg *cfg, int device);
void pirip_fsk_plan_destroy(pirip_fsk_plan *plan);
/* derived sizes: Nbits per frame, M*Nsym, and an upper bound on frames per capture */
int pirip_fsk_plan_nbits(const pirip_fsk_plan *plan);
int pirip_fsk_plan_nfilt(const pirip_fsk_plan *plan);
size_t pirip_fsk_plan_max_frames(const pirip_fsk_plan *plan, size_t samples_per_capture);

/* Demodulate n_captures independent captures that are ALREADY IN DEVICE MEMORY.
 * d_iq: capture c starts at byte c*capture_stride_bytes and holds
 * samples_per_capture complex samples of cfg.fmt.  Outputs (device pointers, any may
 * be NULL except d_nframes): d_bits [n_captures][max_frames*Nbits] one bit per byte,
 * d_llr same shape float (soft only), d_rx_filt [n_captures][max_frames*M*Nsym],
 * d_stats [n_captures][max_frames], d_nframes [n_captures].  `stream` is a
 * cudaStream_t (NULL = default stream); the call is asynchronous on that stream.
 * Returns 0, or a negative error code. */
int pirip_fsk_demod_batch_dev(pirip_fsk_plan *plan, const void *d_iq, size_t n_captures,
                              size_t samples_per_capture, size_t capture_stride_bytes,
                              uint8_t *d_bits, float *d_llr, float *d_rx_filt,
                              struct pirip_frame_stats *d_stats, int32_t *d_nframes,
                              size_t max_frames, void *stream);

/* Same with HOST buffers: copies the captures in, runs the kernel, copies the results
 * out, returns when they are in the caller's buffers. */
int pirip_fsk_demod_batch(pirip_fsk_plan *plan, const void *iq, size_t n_captures,
                          size_t samples_per_capture, uint8_t *bits, float *llr, float *rx_filt,
                          struct pirip_frame_stats *stats, int32_t *nframes, size_t max_frames);

/* One process, several GPUs (BASELINE.json config 5: independent captures shard across the GPUs of a box): plans[r]
 * was created on device r' of the caller's choice; the batch is cut into n_plans contiguous blocks whose sizes differ by at most
 * one, every device demodulates its block from and to the caller's host buffers at the same time, and the decoded bit streams
 * arrive in `bits` in capture order (the gather is the device-to-host copy of every shard).  No data-path collective. */
int pirip_fsk_demod_batch_multi(pirip_fsk_plan *const *plans, int n_plans, const void *iq, size_t n_captures,
                                size_t samples_per_capture, uint8_t *bits, float *llr, int32_t *nframes, size_t max_frames);

/* Numerics mode of the batch entry points.
 *   PIRIP_FSK_MODE_FAST (default where the configuration allows it): the tolerance-mode kernel — tone estimates exact,
 *     integrators / timing / LLRs within the float tolerance stated in DESIGN.md, every discrete decision guarded by a
 *     margin; captures with a decision inside the margin are re-run by the exact kernel in the same call, so hard bits,
 *     nin and frame counts are the scalar algorithm's.
 *   PIRIP_FSK_MODE_EXACT: the exact-order kernel for every capture (every float bit-identical to the CPU evaluation).
 * The environment variable PIRIP_FSK_MODE=exact selects the exact mode for plans created afterwards. */
#define PIRIP_FSK_MODE_EXACT 0
#define PIRIP_FSK_MODE_FAST 1
int pirip_fsk_plan_set_mode(pirip_fsk_plan *plan, int mode);
int pirip_fsk_plan_mode(const pirip_fsk_plan *plan);
/* after a batch call in fast mode: number of captures the guards sent to the exact kernel (synchronises the device);
   flags_out (may be NULL) receives the per-capture flag words of the first n captures: 0 = clean, else
   reason bits (1 symbol margin, 2 nin threshold, 4 unused since round 2 - see pirip_fsk_plan_ppm_near_ties, 8 NaN, 16 ingest time-out) | (1 + first flagged frame) << 8 */
long pirip_fsk_plan_flagged(pirip_fsk_plan *plan, int32_t *flags_out, size_t n);
/* frames the fast kernel replayed exactly inside the kernel since the plan was created (out[0]) and replays it had to
   give up (out[1]: those captures are flagged and re-run by the exact kernel); synchronises the device */
int pirip_fsk_plan_replays(pirip_fsk_plan *plan, int32_t out[2]);
/* frames whose ppm update decision (|change of norm_rx_timing| < 0.2, [U fsk.c::fsk_demod_core]) lay within the estimates' own
   error of the threshold and was taken from the fast values (not replayed: it only feeds the ppm statistic, and the scalar
   algorithm's own decision there depends on the libm in use); synchronises the device */
long pirip_fsk_plan_ppm_near_ties(pirip_fsk_plan *plan);

/* one-line description of the kernel and launch geometry the batch entry points use for this plan
   (static storage inside the plan; valid until the plan is destroyed) */
const char *pirip_fsk_plan_kernel(const pirip_fsk_plan *plan);

/* Pack nbits one-bit-per-byte values (device memory, nbits a multiple of 8, d_bits 16-byte aligned) into nbits/8 bytes, MSB
   first - the bit order of codec2's freedv_pack as pirip uses it ([R tx/rpitx_fsk.cpp:38,75-83]).  Used to gather the decoded
   bit streams of a sharded batch with an eighth of the bytes.  Asynchronous on `stream`. */
int pirip_pack_bits_dev(const uint8_t *d_bits, size_t nbits, uint8_t *d_packed, void *stream);

/* number of kernel launches issued by this library so far (bench.py gpu_launches) */
uint64_t pirip_kernel_launches(void);
/* last error message of the calling thread ("" if none) */
const char *pirip_last_error(void);
/* library / build identification */
const char *pirip_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PIRIP_FSK_H */
