/*
 * orc_frame.c — CPU oracle: FSK_LDPC framing, CRC, test-bit sources, and the
 * exported math probes.  TEST INFRASTRUCTURE (see oracle.h).
 * PARITY UNPINNED: restatement of codec2 src/freedv_fsk.c, src/freedv_api.c,
 * src/ofdm.c (ofdm_rand), src/fsk_get_test_bits.c, whose prototypes pirip itself
 * declares at [R tx/rpitx_fsk.cpp:34-39] and whose frame layout pirip's
 * transmitter builds at [R tx/rpitx_fsk.cpp:75-83,313-336,382-396].
 * SURVEY.md 8a rows A13, A16-A18, App. A.7.
 */
#include "oracle.h"
#include "orc_math.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>

/* [U freedv_api.c::freedv_gen_crc16] CRC-16/CCITT-FALSE: poly 0x1021, init 0xFFFF,
   no reflection, no final xor.  Check value "123456789" -> 0x29B1. */
uint16_t orc_gen_crc16(const uint8_t *data, int length) {
    uint16_t crc = 0xFFFF;
    for (int n = 0; n < length; ++n) {
        crc ^= (uint16_t)((uint16_t)data[n] << 8);
        for (int b = 0; b < 8; ++b) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}

/* [U freedv_api.c::freedv_pack / freedv_unpack] MSB first, consistent with
   (source_byte >> (7-i)) & 1 at [R tx/rpitx_fsk.cpp:384] */
void orc_pack(uint8_t *bytes, const uint8_t *bits, int nbits) {
    memset(bytes, 0, (size_t)((nbits + 7) / 8));
    for (int i = 0; i < nbits; ++i) bytes[i >> 3] |= (uint8_t)((bits[i] & 1) << (7 - (i & 7)));
}

void orc_unpack(uint8_t *bits, const uint8_t *bytes, int nbits) {
    for (int i = 0; i < nbits; ++i) bits[i] = (bytes[i >> 3] >> (7 - (i & 7))) & 1;
}

/* [U freedv_api.c::freedv_check_crc16_unpacked] */
int orc_check_crc16_unpacked(const uint8_t *bits, int nbits) {
    assert((nbits % 8) == 0);
    int nbytes = nbits / 8;
    uint8_t *bytes = (uint8_t *)malloc((size_t)nbytes);
    orc_pack(bytes, bits, nbits);
    uint16_t tx_crc16 = (uint16_t)((bytes[nbytes - 2] << 8) | bytes[nbytes - 1]);
    uint16_t rx_crc16 = orc_gen_crc16(bytes, nbytes - 2);
    free(bytes);
    return tx_crc16 == rx_crc16;
}

/* [R tx/rpitx_fsk.cpp:75-83] calculate_and_insert_crc */
void orc_insert_crc(uint8_t data_bits[], int data_bits_per_frame) {
    assert((data_bits_per_frame % 8) == 0);
    int nbytes = data_bits_per_frame / 8;
    uint8_t *bytes = (uint8_t *)calloc((size_t)nbytes, 1);
    orc_pack(bytes, data_bits, data_bits_per_frame - 16);
    uint16_t crc16 = orc_gen_crc16(bytes, nbytes - 2);
    uint8_t crc_bytes[2] = { (uint8_t)(crc16 >> 8), (uint8_t)(crc16 & 0xff) };
    orc_unpack(data_bits + data_bits_per_frame - 16, crc_bytes, 16);
    free(bytes);
}

/* [U freedv_fsk.c::fsk_ldpc_uw] 0x5186fe15, MSB first */
void orc_uw_bits(uint8_t uw[32]) {
    const uint32_t word = 0x5186fe15u;
    for (int i = 0; i < 32; ++i) uw[i] = (word >> (31 - i)) & 1;
}

/* [U ofdm.c::ofdm_rand + ofdm_generate_payload_data_bits] [R tx/rpitx_fsk.cpp:36,367] */
void orc_generate_payload_data_bits(uint8_t bits[], int n) {
    uint64_t seed = 1;
    for (int i = 0; i < n; ++i) {
        seed = (1103515245ull * seed + 12345) % 32768;
        bits[i] = (uint16_t)seed > 16384;
    }
}

/* [U fsk_get_test_bits.c] one frame of glibc rand() bits from srand(158324)
   [R README.md:101], [R test/include.sh:6-8] */
void orc_test_bits_frame(uint8_t bits[], int frame_len) {
    srand(158324);
    for (int i = 0; i < frame_len; ++i) bits[i] = rand() & 0x1;
}

/* [U freedv_fsk.c::freedv_tx_fsk_ldpc_framer] frame = UW | data | parity */
void orc_fsk_ldpc_framer(const struct orc_ldpc *ldpc, uint8_t frame[], const uint8_t payload[]) {
    int K = ldpc->CodeLength - ldpc->NumberParityBits;
    orc_uw_bits(frame);
    memcpy(frame + 32, payload, (size_t)K);
    orc_ldpc_encode(ldpc, frame + 32, frame + 32 + K);
}

/* ---------------------------------------------------------------- deframer
 * [U freedv_fsk.c::freedv_rx_fsk_ldpc_data]: modem-frame sized chunks of LLRs are
 * queued; every time a whole frame's worth is queued the two-frame buffer slides
 * by one frame and the sync state machine runs once.  Log fields of `rtl_fsk -v`
 * [R README.md:200-208]: nbits = LLRs left queued, state, uw_loc, uw_err, bad_uw,
 * iter, pcc, rxst.
 */
struct orc_deframer *orc_deframer_create(const struct orc_ldpc *ldpc, int max_nbits_per_call) {
    struct orc_deframer *d = (struct orc_deframer *)calloc(1, sizeof *d);
    d->ldpc = *ldpc;
    d->bits_per_frame = 32 + ldpc->CodeLength;
    d->data_bits = ldpc->CodeLength - ldpc->NumberParityBits;
    d->twoframes_llr = (float *)calloc((size_t)(2 * d->bits_per_frame), sizeof(float));
    d->twoframes_hard = (uint8_t *)calloc((size_t)(2 * d->bits_per_frame), 1);
    d->frame_llr_size = d->bits_per_frame + max_nbits_per_call;
    d->frame_llr = (float *)calloc((size_t)d->frame_llr_size, sizeof(float));
    d->frame_llr_nbits = 0;
    d->state = 0;
    d->best_location = 0;
    d->baduw = 0;
    d->thresh1 = 5;
    d->thresh2 = 6;
    d->baduw_thresh = 1;
    return d;
}

void orc_deframer_destroy(struct orc_deframer *d) {
    if (!d) return;
    free(d->twoframes_llr);
    free(d->twoframes_hard);
    free(d->frame_llr);
    free(d);
}

int orc_deframer_push(struct orc_deframer *d, const float llr[], int nbits, uint8_t payload_bits[]) {
    const int bpf = d->bits_per_frame;
    int rx_status = 0;
    uint8_t uw[32];
    orc_uw_bits(uw);

    assert(d->frame_llr_nbits + nbits <= d->frame_llr_size);
    memcpy(d->frame_llr + d->frame_llr_nbits, llr, sizeof(float) * (size_t)nbits);
    d->frame_llr_nbits += nbits;

    if (d->frame_llr_nbits >= bpf) {
        /* slide the two-frame window by one frame and take the oldest queued frame */
        memmove(d->twoframes_llr, d->twoframes_llr + bpf, sizeof(float) * (size_t)bpf);
        memcpy(d->twoframes_llr + bpf, d->frame_llr, sizeof(float) * (size_t)bpf);
        memmove(d->frame_llr, d->frame_llr + bpf, sizeof(float) * (size_t)(d->frame_llr_nbits - bpf));
        d->frame_llr_nbits -= bpf;
        for (int i = 0; i < 2 * bpf; ++i) d->twoframes_hard[i] = d->twoframes_llr[i] < 0;

        int next_state = d->state;
        int errors = 0;
        switch (d->state) {
            case 0: {
                /* out of sync: best UW match over one frame of start positions */
                int best_errors = 32;
                d->best_location = 0;
                for (int i = 0; i < bpf; ++i) {
                    errors = 0;
                    for (int u = 0; u < 32; ++u) errors += d->twoframes_hard[i + u] ^ uw[u];
                    if (errors < best_errors) {
                        best_errors = errors;
                        d->best_location = i;
                    }
                }
                if (best_errors <= d->thresh1) {
                    errors = best_errors;
                    next_state = 1;
                    d->baduw = 0;
                }
                break;
            }
            case 1:
                /* in sync: UW must keep matching at the same location */
                for (int u = 0; u < 32; ++u) errors += d->twoframes_hard[d->best_location + u] ^ uw[u];
                if (errors > d->thresh2) {
                    d->baduw++;
                    if (d->baduw >= d->baduw_thresh) next_state = 0;
                } else {
                    d->baduw = 0;
                }
                break;
        }
        d->last_uw_errors = errors;
        d->last_iter = 0;
        d->last_pcc = 0;
        if (next_state == 1) {
            uint8_t *decoded = (uint8_t *)malloc((size_t)d->ldpc.CodeLength);
            int pcc = 0;
            int iter = orc_run_ldpc_decoder(&d->ldpc, decoded, &d->twoframes_llr[d->best_location + 32], &pcc);
            d->last_iter = iter;
            d->last_pcc = pcc;
            if (pcc != d->ldpc.NumberParityBits) rx_status |= ORC_RX_BIT_ERRORS;
            if (orc_check_crc16_unpacked(decoded, d->data_bits)) {
                rx_status |= ORC_RX_BITS;
                memcpy(payload_bits, decoded, (size_t)d->data_bits);
            } else if (d->state == 0) {
                /* CRC failed on the first frame after a UW hit: dud UW, keep searching */
                next_state = 0;
            }
            free(decoded);
        }
        d->state = next_state;
    }
    if (d->state == 1) rx_status |= ORC_RX_SYNC;
    return rx_status;
}

/* fields of the `rtl_fsk -v` log line of the last push [R README.md:200]: state, uw_loc, uw_err, iter, pcc, nbits queued */
void orc_deframer_info(const struct orc_deframer *d, int info[6]) {
    info[0] = d->state;
    info[1] = d->best_location;
    info[2] = d->last_uw_errors;
    info[3] = d->last_iter;
    info[4] = d->last_pcc;
    info[5] = d->frame_llr_nbits;
}

/* ---------------------------------------------------------------- math probes */
int orc_math_backend(void) { return ORC_MATH_BACKEND; }
float orc_cosf(float x) { return om_cosf(x); }
float orc_sinf(float x) { return om_sinf(x); }
float orc_atan2f(float y, float x) { return om_atan2f(y, x); }
float orc_expf(float x) { return (float)om_exp((double)x); }
float orc_logf(float x) { return (float)om_log((double)x); }
float orc_log10f(float x) { return om_log10f(x); }
