/*
 * ldpc_noise — codec2's AWGN stage between ldpc_enc --sd and ldpc_dec --sd ([U codec2 src/ldpc_noise.c]): adds Gaussian
 * noise of variance No / 2 to a stream of float soft decisions.  Host-only test plumbing.
 *
 *   ldpc_noise InputFile OutputFile NodB        (Es = 1, so Es/No = -NodB)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s InputFile OutputFile NodB\n", argv[0]);
        return 1;
    }
    FILE *fin = !strcmp(argv[1], "-") ? stdin : fopen(argv[1], "rb");
    FILE *fout = !strcmp(argv[2], "-") ? stdout : fopen(argv[2], "wb");
    if (!fin || !fout) {
        fprintf(stderr, "ldpc_noise: cannot open %s\n", !fin ? argv[1] : argv[2]);
        return 1;
    }
    const double No = pow(10.0, atof(argv[3]) / 10.0), sigma = sqrt(No / 2.0);
    float x;
    double sum_xx = 0.0, sum_nn = 0.0;
    long n = 0;
    while (fread(&x, sizeof x, 1, fin) == 1) {
        double u = 0.0; /* sum of 12 uniforms: unit-variance Gaussian to +/- 6 sigma */
        for (int i = 0; i < 12; ++i) u += (double)rand() / RAND_MAX;
        const double noise = sigma * (u - 6.0);
        const float y = (float)((double)x + noise);
        sum_xx += (double)x * x, sum_nn += noise * noise, ++n;
        fwrite(&y, sizeof y, 1, fout);
    }
    if (n) fprintf(stderr, "ldpc_noise: %ld samples, measured Es/No %.2f dB\n", n, 10.0 * log10(sum_xx / (2.0 * sum_nn + 1e-30)));
    if (fin != stdin) fclose(fin);
    if (fout != stdout) fclose(fout);
    return 0;
}
