/*
 * fsk_put_test_bits — test-frame sink with the command line of codec2's tool, as used at
 * [R README.md:105,114], [R test/loopback_rtl_sdr.sh:16], [R test/loopback_rtl_fsk.sh:10]:
 *     fsk_put_test_bits [-q] [-p passPackets] [-b berLimit] [-f frameLen] [-t threshold] in
 * Slides a frameLen window over the received bit stream (one bit per byte), counts a packet
 * whenever the window is within `threshold` (default 0.1) bit-error ratio of the known test
 * frame, prints BER, and exits 0 on PASS / 1 on FAIL.  Host-only tool.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char *argv[]) {
    int quiet = 0, framesize = 100, pass_packets = 0;
    double ber_limit = -1.0, thresh = 0.1;
    const char *in = NULL;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-q")) quiet = 1;
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) pass_packets = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-b") && i + 1 < argc) ber_limit = atof(argv[++i]);
        else if (!strcmp(argv[i], "-f") && i + 1 < argc) framesize = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-t") && i + 1 < argc) thresh = atof(argv[++i]);
        else in = argv[i];
    }
    if (!in || framesize <= 0) {
        fprintf(stderr, "usage: %s [-q] [-p passPackets] [-b berLimit] [-f frameLen] [-t threshold] InputBitsOnePerByte\n", argv[0]);
        return 1;
    }
    FILE *fin = strcmp(in, "-") == 0 ? stdin : fopen(in, "rb");
    if (!fin) {
        fprintf(stderr, "Error opening input file: %s\n", in);
        return 1;
    }
    uint8_t *tx = (uint8_t *)malloc((size_t)framesize), *rx = (uint8_t *)calloc((size_t)framesize, 1);
    srand(158324);
    for (int i = 0; i < framesize; ++i) tx[i] = rand() & 0x1;
    long bitcnt = 0, biterr = 0, packets = 0, nread = 0;
    int c, since_last = 0;
    while ((c = fgetc(fin)) != EOF) {
        memmove(rx, rx + 1, (size_t)framesize - 1);
        rx[framesize - 1] = (uint8_t)(c & 1);
        nread++;
        since_last++;
        if (nread < framesize) continue;
        int errs = 0;
        for (int j = 0; j < framesize; ++j) errs += rx[j] != tx[j];
        if (errs < thresh * framesize && since_last >= framesize) {
            packets++;
            bitcnt += framesize;
            biterr += errs;
            since_last = 0;
            if (!quiet) fprintf(stderr, "[%04ld] BER %5.3f, bits tested %6ld, bit errors %6ld errs: %4d\n", packets,
                                (double)biterr / bitcnt, bitcnt, biterr, errs);
        }
    }
    double ber = bitcnt ? (double)biterr / bitcnt : 0.5;
    fprintf(stderr, "[%04ld] BER %5.3f, bits tested %6ld, bit errors %6ld\n", packets, ber, bitcnt, biterr);
    int pass = 1;
    if (pass_packets > 0 && packets < pass_packets) pass = 0;
    if (ber_limit >= 0 && (bitcnt == 0 || ber > ber_limit)) pass = 0;
    if (pass_packets > 0 || ber_limit >= 0) fprintf(stderr, pass ? "PASS\n" : "FAIL\n");
    free(tx);
    free(rx);
    return pass ? 0 : 1;
}
