/*
 * csdr — the three csdr sub-commands on pirip's receive path ([R README.md:109,162]) with raw
 * native-endian stdin/stdout streams:
 *     csdr convert_u8_f | csdr fir_decimate_cc <N> [transition_bw] | csdr convert_f_s16
 * fir_decimate_cc keeps the unconsumed tail of its buffer, so the stream sees one continuous
 * causal FIR anchored at sample 0 ([U csdr.c]); the FIR itself runs on the GPU
 * (fir_decimate_cc() of libpirip_b200.so).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/pirip_csdr.h"

int main(int argc, char *argv[]) {
    if (argc < 2) {
        fprintf(stderr, "usage: csdr convert_u8_f | fir_decimate_cc <N> [transition_bw] | convert_f_s16\n");
        return 1;
    }
    const size_t bufsize = 16384;
    if (!strcmp(argv[1], "convert_u8_f")) {
        std::vector<unsigned char> in(bufsize);
        std::vector<float> out(bufsize);
        size_t n;
        while ((n = fread(in.data(), 1, bufsize, stdin)) > 0) {
            convert_u8_f(in.data(), out.data(), (int)n);
            fwrite(out.data(), sizeof(float), n, stdout);
            fflush(stdout);
        }
        return 0;
    }
    if (!strcmp(argv[1], "convert_f_s16")) {
        std::vector<float> in(bufsize);
        std::vector<short> out(bufsize);
        size_t n;
        while ((n = fread(in.data(), sizeof(float), bufsize, stdin)) > 0) {
            convert_f_s16(in.data(), out.data(), (int)n);
            fwrite(out.data(), sizeof(short), n, stdout);
            fflush(stdout);
        }
        return 0;
    }
    if (!strcmp(argv[1], "fir_decimate_cc")) {
        if (argc < 3) {
            fprintf(stderr, "fir_decimate_cc: decimation factor required\n");
            return 1;
        }
        const int D = atoi(argv[2]);
        float tbw = 0.05;
        if (argc >= 4) tbw = (float)atof(argv[3]);
        if (D < 1) return 1;
        const int L = firdes_filter_len(tbw);
        std::vector<float> taps((size_t)L);
        firdes_lowpass_f(taps.data(), L, 0.5 / (float)D, 0);
        fprintf(stderr, "fir_decimate_cc: taps_length = %d\n", L);
        const size_t cap = bufsize > (size_t)(4 * L) ? bufsize : (size_t)(4 * L);
        std::vector<complexf> in(cap), out(cap / (size_t)D + 2);
        size_t have = 0;
        for (;;) {
            const size_t got = fread(in.data() + have, sizeof(complexf), cap - have, stdin);
            have += got;
            if (have >= (size_t)L) {
                const int no = fir_decimate_cc(in.data(), out.data(), (int)have, D, taps.data(), L);
                fwrite(out.data(), sizeof(complexf), (size_t)no, stdout);
                fflush(stdout);
                const size_t skip = (size_t)D * (size_t)no;
                memmove(in.data(), in.data() + skip, (have - skip) * sizeof(complexf));
                have -= skip;
            }
            if (got == 0) break;
        }
        return 0;
    }
    fprintf(stderr, "csdr: unknown function %s\n", argv[1]);
    return 1;
}
