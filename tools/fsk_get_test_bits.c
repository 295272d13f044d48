/*
 * fsk_get_test_bits — test-frame source with the command line and stream format of codec2's
 * tool of the same name, as used at [R README.md:101,179] and [R test/loopback_rtl_fsk.sh:17]:
 *     fsk_get_test_bits out numBits [frameLen]
 * One frame of frameLen (default 100) bits from glibc rand() seeded with 158324 is repeated;
 * output is one bit per byte ("-" = stdout).  Host-only tool.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TEST_FRAME_SIZE 100

int main(int argc, char *argv[]) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s OutputBitsOnePerByte numBits [frameLen]\n", argv[0]);
        return 1;
    }
    FILE *fout = strcmp(argv[1], "-") == 0 ? stdout : fopen(argv[1], "wb");
    if (!fout) {
        fprintf(stderr, "Error opening output file: %s\n", argv[1]);
        return 1;
    }
    long numBits = atol(argv[2]);
    int framesize = argc > 3 ? atoi(argv[3]) : TEST_FRAME_SIZE;
    if (numBits < 0 || framesize <= 0) {
        fprintf(stderr, "bad numBits / frameLen\n");
        return 1;
    }
    uint8_t *frame = (uint8_t *)malloc((size_t)framesize);
    srand(158324);
    for (int i = 0; i < framesize; ++i) frame[i] = rand() & 0x1;
    /* whole frames only, like upstream: numBits rounded up to a multiple of the frame */
    for (long sent = 0; sent < numBits; sent += framesize) {
        fwrite(frame, 1, (size_t)framesize, fout);
        if (fout == stdout) fflush(stdout);
    }
    free(frame);
    if (fout != stdout) fclose(fout);
    return 0;
}
