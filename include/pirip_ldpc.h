/*
 * pirip_ldpc.h — C-ABI of the FEC / framing stage of pirip's receive chain:
 * soft-decision mapping, LDPC sum-product decoding, FSK_LDPC framing (UW | data+CRC16 | parity).
 *
 * Layer 1 keeps the codec2 names pirip itself links against or that rtl_fsk / ldpc_dec call:
 *   - [R tx/rpitx_fsk.cpp:34-39] declares freedv_tx_fsk_ldpc_bits_per_frame,
 *     freedv_tx_fsk_ldpc_framer, ofdm_generate_payload_data_bits, freedv_gen_crc16,
 *     freedv_pack, freedv_unpack; [R tx/rpitx_fsk.cpp:225,319-324,541] use ldpc_codes_list,
 *     freedv_open_advanced(FREEDV_MODE_FSK_LDPC), freedv_get_bits_per_modem_frame, freedv_close;
 *     [R tx/frame_repeater.c:71-88] uses FREEDV_RX_SYNC / FREEDV_RX_BITS.
 *   - codec2 mpdecode_core.h: run_ldpc_decoder, fsk_rx_filt_to_llrs; ldpc_codes.h:
 *     ldpc_codes_find / ldpc_codes_setup.
 * The decoder and the LLR map run on the GPU; table set-up, framing and CRC are host code.
 *
 * Layer 2 is the batched decoder: many codewords per launch, device-resident LLRs.
 *
 * Code tables: codec2's H_*.c files are not available offline; the built-in codes are stand-in
 * repeat-accumulate codes with the same names, sizes and table layout (tools/gen_ldpc_codes.py).
 * ldpc_codes_add() registers a real codec2 table at run time.
 */
#ifndef PIRIP_LDPC_H
#define PIRIP_LDPC_H

#include <stddef.h>
#include <stdint.h>

#include "pirip_fsk.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- codec2 mpdecode_core.h::struct LDPC ---------------------------------------------- */
struct LDPC {
    char name[32];
    int max_iter;
    int dec_type;
    float q_scale_factor;
    float r_scale_factor;
    int CodeLength;
    int NumberParityBits;
    int NumberRowsHcols;
    int max_row_weight;
    int max_col_weight;
    uint16_t *H_rows;
    uint16_t *H_cols;
    int ldpc_data_bits_per_frame;
    int ldpc_coded_bits_per_frame;
    int protection_mode;
    int data_bits_per_frame;
    int coded_bits_per_frame;
};

/* replaces codec2 ldpc_codes.h */
int ldpc_codes_num(void);
void ldpc_codes_list(void);
int ldpc_codes_find(char name[]);
void ldpc_codes_setup(struct LDPC *ldpc, char name[]);
/* new: register a table in codec2's layout (copied); returns its index or -1 */
int ldpc_codes_add(const char *name, int max_iter, int CodeLength, int NumberParityBits, int NumberRowsHcols,
                   int max_row_weight, int max_col_weight, const uint16_t *H_rows, const uint16_t *H_cols);

/* replaces codec2 mpdecode_core.h::encode (renamed: `encode` is too generic a symbol to export) —
   RA accumulate: pbits[p] = (sum_i ibits[H_rows[p + i*NPB] - 1] + pbits[p-1]) & 1 */
void ldpc_encode_frame(struct LDPC *ldpc, unsigned char ibits[], unsigned char pbits[]);
/* replaces codec2 mpdecode_core.h::run_ldpc_decoder — returns iterations used (GPU) */
int run_ldpc_decoder(struct LDPC *ldpc, uint8_t out_char[], float input[], int *parityCheckCount);
/* replaces codec2 mpdecode_core.h::fsk_rx_filt_to_llrs (GPU) */
void fsk_rx_filt_to_llrs(float llr[], float rx_filt[], float v_est, float SNRest, int M, int nsyms);

/* ---- codec2 freedv_api.h subset on pirip's path ----------------------------------------- */
#define FREEDV_MODE_FSK_LDPC 9
#define FREEDV_RX_TRIAL_SYNC 0x1
#define FREEDV_RX_SYNC 0x2
#define FREEDV_RX_BITS 0x4
#define FREEDV_RX_BIT_ERRORS 0x8

struct freedv_advanced {
    int interleave_frames;
    int M;
    int Rs;
    int Fs;
    int first_tone;
    int tone_spacing;
    char *codename;
};
struct freedv; /* opaque */

/* replaces freedv_api.h::freedv_open_advanced (FREEDV_MODE_FSK_LDPC only; NULL otherwise) */
struct freedv *freedv_open_advanced(int mode, struct freedv_advanced *adv);
void freedv_close(struct freedv *f);
int freedv_get_bits_per_modem_frame(struct freedv *f);
/* replaces the functions declared at [R tx/rpitx_fsk.cpp:34-39] */
int freedv_tx_fsk_ldpc_bits_per_frame(struct freedv *f);
void freedv_tx_fsk_ldpc_framer(struct freedv *f, uint8_t frame[], uint8_t payload_data[]);
void ofdm_generate_payload_data_bits(uint8_t payload_data_bits[], int n);
unsigned short freedv_gen_crc16(unsigned char *data_p, int length);
void freedv_pack(unsigned char *bytes, unsigned char *bits, int nbits);
void freedv_unpack(unsigned char *bits, unsigned char *bytes, int nbits);
int freedv_check_crc16_unpacked(unsigned char *bits, int nbits);
/* FSK_LDPC receive side ([U freedv_fsk.c::freedv_rx_fsk_ldpc_data]): push one modem frame of
   LLRs (nbits of them); returns rx_status (FREEDV_RX_*); payload_bits gets
   freedv_get_bits_per_modem_frame() unpacked bits when FREEDV_RX_BITS is set */
int freedv_fsk_ldpc_rx_llrs(struct freedv *f, const float llr[], int nbits, uint8_t payload_bits[]);
/* fields of the rtl_fsk -v log line [R README.md:200]: state, uw_loc, uw_err, iter, pcc, nbits queued */
void freedv_fsk_ldpc_rx_info(struct freedv *f, int info[6]);

/* ---- layer 2: batched decoder ------------------------------------------------------------ */
typedef struct pirip_ldpc_plan pirip_ldpc_plan;
pirip_ldpc_plan *pirip_ldpc_plan_create(const struct LDPC *ldpc, int device);
void pirip_ldpc_plan_destroy(pirip_ldpc_plan *plan);
/* Check-node update of the batched decoder (all modes honour the code's iteration cap - 15 for FSK_LDPC [R README.md:211] -
 * and stop at the first iteration whose parity checks all pass):
 *   PIRIP_LDPC_SUM_PRODUCT (default): [U mpdecode_core.c::SumProduct] with its phi0 in double - decoded bits, `iter` and
 *     `pcc` identical to the CPU evaluation;
 *   PIRIP_LDPC_SUM_PRODUCT_FAST: the same update with phi0 in single precision (hardware exp / log): same bits on every
 *     word that converges, `iter` may differ by one now and then;
 *   PIRIP_LDPC_MIN_SUM: normalised min-sum (scale 0.75): no transcendentals; same payload on converged words, a few tenths
 *     of a dB earlier breakdown at threshold.
 * PIRIP_LDPC_MODE=minsum|fast in the environment selects the mode of plans created afterwards. */
#define PIRIP_LDPC_SUM_PRODUCT 0
#define PIRIP_LDPC_SUM_PRODUCT_FAST 1
#define PIRIP_LDPC_MIN_SUM 2
int pirip_ldpc_plan_set_mode(pirip_ldpc_plan *plan, int mode);
int pirip_ldpc_plan_mode(const pirip_ldpc_plan *plan);
/* d_llr [n_codewords][CodeLength] float (positive = bit 0); outputs (device): d_bits
 * [n_codewords][CodeLength] one bit per byte, d_iter [n_codewords], d_pcc [n_codewords]
 * (either count may be NULL).  Asynchronous on `stream`.  Returns 0 or a negative code. */
int pirip_ldpc_decode_batch_dev(pirip_ldpc_plan *plan, const float *d_llr, size_t n_codewords, uint8_t *d_bits,
                                int32_t *d_iter, int32_t *d_pcc, void *stream);
/* host buffers, synchronous */
int pirip_ldpc_decode_batch(pirip_ldpc_plan *plan, const float *llr, size_t n_codewords, uint8_t *bits, int32_t *iter,
                            int32_t *pcc);

/* ---- layer 2: the coded receive chain of a batch of captures on the device (BASELINE.json config 4) ---------------
 * u8 / s16 / f32 IQ captures -> FSK demodulation with LLR output -> per capture: 32-bit UW search over a two-frame
 * window, sync state machine, LDPC decode of the frame behind the UW, CRC-16 -> per MODEM FRAME one rx_status byte
 * (FREEDV_RX_*) and, where FREEDV_RX_BITS is set, the packed payload: the record stream `rtl_fsk --code <name> -b` writes
 * ([R README.md:184,292]; consumer [R tx/frame_repeater.c:55-62]), for every capture of the batch in two launches (the
 * demodulator, then ONE launch that runs framing + LDPC + CRC for all captures).  The LDPC update is the ldpc plan's mode
 * (pirip_ldpc_plan_set_mode on pirip_fsk_ldpc_rx_plan_ldpc()); the demodulator's numerics mode likewise. */
typedef struct pirip_rx_plan pirip_rx_plan;
pirip_rx_plan *pirip_fsk_ldpc_rx_plan_create(const struct pirip_fsk_cfg *cfg, const char *codename, int device);
void pirip_fsk_ldpc_rx_plan_destroy(pirip_rx_plan *plan);
pirip_fsk_plan *pirip_fsk_ldpc_rx_plan_fsk(pirip_rx_plan *plan);
pirip_ldpc_plan *pirip_fsk_ldpc_rx_plan_ldpc(pirip_rx_plan *plan);
int pirip_fsk_ldpc_rx_plan_payload_bytes(const pirip_rx_plan *plan);
/* host buffers: status [n][max_frames], payload [n][max_frames][payload_bytes], nframes [n] (may be NULL), info (may be
 * NULL) [n][max_frames][4] = uw_loc (-1: no frame step at this modem frame), uw_err, iter, pcc - the fields of the
 * `rtl_fsk -v` log line [R README.md:200].  max_frames from pirip_fsk_plan_max_frames().  Synchronous. */
int pirip_fsk_ldpc_rx_batch(pirip_rx_plan *plan, const void *iq, size_t n_captures, size_t samples_per_capture, uint8_t *status,
                            uint8_t *payload, int32_t *nframes, int32_t *info, size_t max_frames);
/* device buffers (d_llr [n][max_frames * Nbits] scratch), asynchronous on `stream` */
int pirip_fsk_ldpc_rx_batch_dev(pirip_rx_plan *plan, const void *d_iq, size_t n_captures, size_t samples_per_capture,
                                size_t capture_stride_bytes, float *d_llr, int32_t *d_nframes, uint8_t *d_status, uint8_t *d_payload,
                                int32_t *d_info, size_t max_frames, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PIRIP_LDPC_H */
