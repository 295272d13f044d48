/*
 * ldpc_enc — codec2's stand-alone LDPC encoder command line ([U codec2 src/ldpc_enc.c]): host-only, the
 * RA accumulate encoder behind ldpc_encode_frame() (include/pirip_ldpc.h).
 *
 *   ldpc_enc InOneBitPerByte OutOneBitPerByte [--sd] [--code CodeName] [--testframes Nframes]
 *
 * Reads data_bits_per_frame bits (one per byte) per frame, writes the CodeLength-bit codeword (data, then parity), one
 * bit per byte, or with --sd as BPSK soft decisions 1 - 2 bit (float).  --testframes N ignores the input and sends N
 * frames of the ofdm_rand payload ldpc_dec --testframes scores.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/pirip_ldpc.h"

static int opt(int argc, char **argv, const char *name) {
    for (int i = 1; i < argc; ++i)
        if (!strcmp(argv[i], name)) return i;
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s InOneBitPerByte OutOneBitPerByte [--sd] [--code CodeName] [--testframes Nframes]\n", argv[0]);
        ldpc_codes_list();
        return 1;
    }
    char code[64] = "HRA_112_112";
    int a;
    if ((a = opt(argc, argv, "--code")) && a + 1 < argc) snprintf(code, sizeof code, "%s", argv[a + 1]);
    if (ldpc_codes_find(code) < 0) {
        fprintf(stderr, "ldpc_enc: unknown code %s\n", code);
        return 1;
    }
    struct LDPC ldpc;
    ldpc_codes_setup(&ldpc, code);
    const bool sd = opt(argc, argv, "--sd");
    long nframes = -1;
    if ((a = opt(argc, argv, "--testframes")) && a + 1 < argc) nframes = atol(argv[a + 1]);
    FILE *fin = !strcmp(argv[1], "-") ? stdin : fopen(argv[1], "rb");
    FILE *fout = !strcmp(argv[2], "-") ? stdout : fopen(argv[2], "wb");
    if (!fin || !fout) {
        fprintf(stderr, "ldpc_enc: cannot open %s\n", !fin ? argv[1] : argv[2]);
        return 1;
    }
    const int N = ldpc.CodeLength, NPB = ldpc.NumberParityBits, K = ldpc.data_bits_per_frame ? ldpc.data_bits_per_frame : N - NPB;
    std::vector<uint8_t> cw((size_t)N);
    std::vector<float> sdout((size_t)N);
    for (long f = 0; nframes < 0 || f < nframes; ++f) {
        if (nframes >= 0)
            ofdm_generate_payload_data_bits(cw.data(), K);
        else if (fread(cw.data(), 1, (size_t)K, fin) != (size_t)K)
            break;
        ldpc_encode_frame(&ldpc, cw.data(), cw.data() + K);
        if (sd) {
            for (int i = 0; i < N; ++i) sdout[(size_t)i] = 1.0f - 2.0f * (float)cw[(size_t)i];
            fwrite(sdout.data(), sizeof(float), (size_t)N, fout);
        } else {
            fwrite(cw.data(), 1, (size_t)N, fout);
        }
    }
    if (fin != stdin) fclose(fin);
    if (fout != stdout) fclose(fout);
    return 0;
}
