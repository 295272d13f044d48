/*
 * ldpc_dec — codec2's stand-alone LDPC decoder command line ([U codec2 src/ldpc_dec.c], SURVEY.md 8b / U13, the
 * FEC stage of BASELINE config 4) on libpirip_b200.so: every frame is decoded by the sum-product kernel behind
 * run_ldpc_decoder() (include/pirip_ldpc.h).
 *
 *   ldpc_dec InOneSymbolPerFloat OutOneBitPerByte [--sd] [--half] [--code CodeName] [--testframes] [--max_iter n] [--mute]
 *
 * Input: float LLRs, CodeLength per frame (positive = bit 0), or with --sd BPSK soft decisions (+1 = bit 0) that
 * sd_to_llr() scales by the Es/N0 it estimates from the frame ([U mpdecode_core.c::sd_to_llr]).  Output: the
 * data_bits_per_frame decoded data bits, one bit per byte.  --half loads CodeLength/2 new samples per decode attempt and
 * only emits frames that converged (codec2's poor man's frame sync); --mute drops frames with >= 10 % failed checks;
 * --testframes counts raw and coded errors against the ofdm_rand payload the encoder sends with the same switch and
 * prints codec2's two BER lines on stderr.  `-` is stdin / stdout.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/pirip_ldpc.h"

/* [U mpdecode_core.c::sd_to_llr]: BPSK soft decisions of unknown scale -> LLRs, noise variance estimated around +/- mean */
static void sd_to_llr(float llr[], const float sd[], int n) {
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += fabs((double)sd[i]);
    const double mean = sum / n;
    double sumsq = 0.0;
    sum = 0.0;
    for (int i = 0; i < n; i++) {
        const double sign = (sd[i] > 0.0f) - (sd[i] < 0.0f);
        const double x = (double)sd[i] / mean - sign;
        sum += x;
        sumsq += x * x;
    }
    const double estvar = (n * sumsq - sum * sum) / ((double)n * (n - 1));
    const double estEsN0 = 1.0 / (2.0 * estvar + 1E-3);
    for (int i = 0; i < n; i++) llr[i] = (float)(4.0 * estEsN0 * (double)sd[i]);
}

static int opt(int argc, char **argv, const char *name) {
    for (int i = 1; i < argc; ++i)
        if (!strcmp(argv[i], name)) return i;
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s InOneSymbolPerFloat OutOneBitPerByte [--sd] [--half] [--code CodeName] [--testframes] [--max_iter n] [--mute]\n",
                argv[0]);
        ldpc_codes_list();
        return 1;
    }
    char code[64] = "HRA_112_112";
    int a;
    if ((a = opt(argc, argv, "--code")) && a + 1 < argc) snprintf(code, sizeof code, "%s", argv[a + 1]);
    if (ldpc_codes_find(code) < 0) {
        fprintf(stderr, "ldpc_dec: unknown code %s\n", code);
        return 1;
    }
    struct LDPC ldpc;
    ldpc_codes_setup(&ldpc, code);
    if ((a = opt(argc, argv, "--max_iter")) && a + 1 < argc) ldpc.max_iter = atoi(argv[a + 1]);
    const bool sd = opt(argc, argv, "--sd"), half = opt(argc, argv, "--half"), mute = opt(argc, argv, "--mute"),
               testframes = opt(argc, argv, "--testframes");
    FILE *fin = !strcmp(argv[1], "-") ? stdin : fopen(argv[1], "rb");
    FILE *fout = !strcmp(argv[2], "-") ? stdout : fopen(argv[2], "wb");
    if (!fin || !fout) {
        fprintf(stderr, "ldpc_dec: cannot open %s\n", !fin ? argv[1] : argv[2]);
        return 1;
    }
    const int N = ldpc.CodeLength, NPB = ldpc.NumberParityBits, K = ldpc.data_bits_per_frame ? ldpc.data_bits_per_frame : N - NPB;
    std::vector<float> in((size_t)N, 0.0f), llr((size_t)N);
    std::vector<uint8_t> out((size_t)N), tx((size_t)N);
    if (testframes) { /* the frame ldpc_enc --testframes sends: ofdm_rand data bits + parity */
        ofdm_generate_payload_data_bits(tx.data(), K);
        ldpc_encode_frame(&ldpc, tx.data(), tx.data() + K);
    }
    long Tbits = 0, Terrs = 0, Tbits_raw = 0, Terrs_raw = 0, frames = 0;
    const int nread = half ? N / 2 : N;
    for (;;) {
        if (half) memmove(in.data(), in.data() + nread, sizeof(float) * (size_t)(N - nread));
        if (fread(in.data() + (N - nread), sizeof(float), (size_t)nread, fin) != (size_t)nread) break;
        if (sd)
            sd_to_llr(llr.data(), in.data(), N);
        else
            memcpy(llr.data(), in.data(), sizeof(float) * (size_t)N);
        int pcc = 0;
        const int iter = run_ldpc_decoder(&ldpc, out.data(), llr.data(), &pcc);
        ++frames;
        if (testframes) {
            int e_raw = 0, e = 0;
            for (int i = 0; i < N; ++i) e_raw += (llr[(size_t)i] < 0.0f) != (tx[(size_t)i] != 0);
            for (int i = 0; i < K; ++i) e += out[(size_t)i] != tx[(size_t)i];
            if (!half || pcc == NPB) { /* half-frame mode only scores the attempts that lined up */
                Tbits_raw += N, Terrs_raw += e_raw, Tbits += K, Terrs += e;
            }
        }
        const bool emit = half ? pcc == NPB : (!mute || (NPB - pcc) * 10 < NPB);
        if (emit) fwrite(out.data(), 1, (size_t)K, fout);
        if (opt(argc, argv, "-v")) fprintf(stderr, "frame: %ld iter: %d pcc: %d\n", frames, iter, pcc);
        if (fout == stdout) fflush(stdout);
    }
    if (testframes) {
        fprintf(stderr, "Raw   Tbits: %6ld Terr: %6ld BER: %4.3f\n", Tbits_raw, Terrs_raw, Tbits_raw ? (double)Terrs_raw / Tbits_raw : 0.0);
        fprintf(stderr, "Coded Tbits: %6ld Terr: %6ld BER: %4.3f\n", Tbits, Terrs, Tbits ? (double)Terrs / Tbits : 0.0);
    }
    if (fin != stdin) fclose(fin);
    if (fout != stdout) fclose(fout);
    return 0;
}
