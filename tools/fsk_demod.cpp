/*
 * fsk_demod — the codec2 `fsk_demod` command line on the B200 library, drop-in for the pipelines at
 * [R README.md:105,109] and [R test/loopback_rtl_sdr.sh:16]:
 *     fsk_demod [-c|-d|-f] [-p P] [--nsym N] [--fsk_lower Hz] [--fsk_upper Hz] [--mask spacing] [-s]
 *               [--batch] M Fs Rs InputFile OutputFile          ("-" = stdin / stdout)
 *   -d  u8 complex input (rtl_sdr format)      -c  s16 complex input      -f  float complex input
 *   -s  soft decisions: Nbits float32 LLRs per frame instead of Nbits bytes (one bit per byte)
 *   -t[N] / --stats[=N]  modem statistics as one JSON object per N frames (default 8) on stderr, upstream's keys
 *       "secs", "EbNodB", "ppm", "f1_est" .. "fM_est" ([U fsk_demod.c]); the eye diagram and input spectrum arrays of
 *       upstream's line are not produced ("eye_diagram": [], "samp_fft": [])
 * Default mode follows upstream's main loop call for call: fread fsk_nin() samples, convert,
 * fsk_demod()/fsk_demod_sd(), fwrite — the state stays on the GPU between calls.  --batch reads the
 * whole input first and runs it as one capture through the persistent kernel (same output bytes).
 */
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <vector>

#include "../include/pirip_fsk.h"
#include "../include/pirip_ldpc.h"

#define FDMDV_SCALE 825

static void usage(const char *p) {
    fprintf(stderr,
            "usage: %s [-c|-d|-f] [-p P] [--nsym N] [--fsk_lower Hz] [--fsk_upper Hz] [--mask spacing] [-s] [-t[N]|--stats[=N]] [--batch]\n"
            "          (2|4) SampleRate SymbolRate InputModemRawFile OutputFile\n", p);
    exit(1);
}

int main(int argc, char *argv[]) {
    int fmt = PIRIP_FMT_S16, real_input = 1, P = 0, nsym = FSK_DEFAULT_NSYM, lower = 0, upper = 0, user_lower = 0, user_upper = 0, mask = 0, soft = 0, batch = 0;
    int stats_every = 0;
    static struct option lo[] = { { "fsk_lower", required_argument, 0, 'b' }, { "fsk_upper", required_argument, 0, 'u' },
                                  { "mask", required_argument, 0, 'm' },      { "nsym", required_argument, 0, 'n' },
                                  { "batch", no_argument, 0, 'B' },           { "help", no_argument, 0, 'h' },
                                  { "stats", optional_argument, 0, 't' },
                                  { 0, 0, 0, 0 } };
    int o;
    while ((o = getopt_long(argc, argv, "cdfp:sht::", lo, NULL)) != -1) {
        switch (o) {
            case 'c': fmt = PIRIP_FMT_S16; real_input = 0; break;
            case 'd': fmt = PIRIP_FMT_U8; real_input = 0; break;
            case 'f': fmt = PIRIP_FMT_F32; real_input = 0; break;
            case 'p': P = atoi(optarg); break;
            case 's': soft = 1; break;
            case 't': stats_every = optarg ? atoi(optarg) : 8; if (stats_every < 1) stats_every = 8; break;
            case 'b': lower = atoi(optarg); user_lower = 1; break;
            case 'u': upper = atoi(optarg); user_upper = 1; break;
            case 'm': mask = atoi(optarg); break;
            case 'n': nsym = atoi(optarg); break;
            case 'B': batch = 1; break;
            default: usage(argv[0]);
        }
    }
    if (argc - optind < 5) usage(argv[0]);
    const int M = atoi(argv[optind]), Fs = atoi(argv[optind + 1]), Rs = atoi(argv[optind + 2]);
    const char *inname = argv[optind + 3], *outname = argv[optind + 4];
    if ((M != 2 && M != 4) || Fs <= 0 || Rs <= 0 || Fs % Rs) usage(argv[0]);
    if (P == 0) P = FSK_DEFAULT_P;
    if ((Fs / Rs) % P || P < 4) {
        fprintf(stderr, "Ts/P must be an integer and P >= 4\n");
        return 1;
    }
    FILE *fin = strcmp(inname, "-") == 0 ? stdin : fopen(inname, "rb");
    FILE *fout = strcmp(outname, "-") == 0 ? stdout : fopen(outname, "wb");
    if (!fin || !fout) {
        fprintf(stderr, "Couldn't open files\n");
        return 1;
    }
    /* the estimator window is always set, as upstream's fsk_demod does ([U fsk_demod.c], M-confidence): without
       --fsk_lower the whole negative half is searched as well when the input is complex (-c / -d / -f), --fsk_upper defaults
       to Fs/2 */
    const int have_limits = 1;
    if (!user_lower) lower = real_input ? 0 : -Fs / 2;
    if (!user_upper) upper = Fs / 2;
    if (lower < -Fs / 2 || upper > Fs / 2 || upper <= lower) {
        fprintf(stderr, "fsk_demod: estimator limits must satisfy -Fs/2 <= lower < upper <= Fs/2\n");
        return 1;
    }
    const int tone_spacing = mask ? mask : 400;
    const size_t bps = real_input ? 2 : (fmt == PIRIP_FMT_U8 ? 2 : (fmt == PIRIP_FMT_S16 ? 4 : 8));

    if (batch && !real_input) {
        std::vector<unsigned char> raw;
        unsigned char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fin)) > 0) raw.insert(raw.end(), buf, buf + n);
        struct pirip_fsk_cfg cfg = { Fs, Rs, M, P, nsym, 1200, tone_spacing, lower, upper, have_limits, mask ? 1 : 0, fmt, soft };
        pirip_fsk_plan *plan = pirip_fsk_plan_create(&cfg, 0);
        if (!plan) {
            fprintf(stderr, "fsk_demod: %s\n", pirip_last_error());
            return 1;
        }
        const size_t ns = raw.size() / bps, mf = pirip_fsk_plan_max_frames(plan, ns), nb = (size_t)pirip_fsk_plan_nbits(plan);
        std::vector<uint8_t> bits(mf * nb);
        std::vector<float> llr(soft ? mf * nb : 0);
        int32_t nf = 0;
        if (pirip_fsk_demod_batch(plan, raw.data(), 1, ns, bits.data(), soft ? llr.data() : NULL, NULL, NULL, &nf, mf)) {
            fprintf(stderr, "fsk_demod: %s\n", pirip_last_error());
            return 1;
        }
        if (soft)
            fwrite(llr.data(), sizeof(float), (size_t)nf * nb, fout);
        else
            fwrite(bits.data(), 1, (size_t)nf * nb, fout);
        pirip_fsk_plan_destroy(plan);
        fclose(fout);
        return 0;
    }

    struct FSK *fsk = fsk_create_hbr(Fs, Rs, M, P, nsym, 1200, tone_spacing);
    if (have_limits) fsk_set_freq_est_limits(fsk, lower, upper);
    if (mask) fsk_set_freq_est_alg(fsk, 1);
    const int nmax = fsk->N + fsk->Ts * 2;
    std::vector<unsigned char> raw((size_t)nmax * 8);
    std::vector<COMP> mod((size_t)nmax);
    std::vector<uint8_t> bits((size_t)fsk->Nbits);
    std::vector<float> filt((size_t)(fsk->mode * fsk->Nsym)), llrs((size_t)fsk->Nbits);
    long frames_seen = 0;
    for (;;) {
        const size_t nin = fsk_nin(fsk);
        if (fread(raw.data(), bps, nin, fin) != nin) break; /* short read ends the loop, as upstream */
        if (real_input) {
            const int16_t *r = (const int16_t *)raw.data();
            for (size_t i = 0; i < nin; ++i) {
                mod[i].real = ((float)r[i]) / FDMDV_SCALE;
                mod[i].imag = 0.0;
            }
        } else if (fmt == PIRIP_FMT_U8) {
            for (size_t i = 0; i < nin; ++i) {
                mod[i].real = ((float)raw[2 * i] - 127.0) / 128.0;
                mod[i].imag = ((float)raw[2 * i + 1] - 127.0) / 128.0;
            }
        } else if (fmt == PIRIP_FMT_S16) {
            const int16_t *r = (const int16_t *)raw.data();
            for (size_t i = 0; i < nin; ++i) {
                mod[i].real = ((float)r[2 * i]) / FDMDV_SCALE;
                mod[i].imag = ((float)r[2 * i + 1] / FDMDV_SCALE);
            }
        } else {
            memcpy(mod.data(), raw.data(), nin * sizeof(COMP));
        }
        if (soft) {
            fsk_demod_sd(fsk, filt.data(), mod.data());
            fsk_rx_filt_to_llrs(llrs.data(), filt.data(), fsk->v_est, fsk->SNRest, fsk->mode, fsk->Nsym);
            fwrite(llrs.data(), sizeof(float), (size_t)fsk->Nbits, fout);
        } else {
            fsk_demod(fsk, bits.data(), mod.data());
            fwrite(bits.data(), 1, (size_t)fsk->Nbits, fout);
        }
        if (fin == stdin || fout == stdout) fflush(fout);
        if (stats_every && ++frames_seen % stats_every == 0) {
            struct MODEM_STATS st;
            fsk_get_demod_stats(fsk, &st);
            fprintf(stderr, "{\"secs\": %ld, \"EbNodB\": %5.1f, \"ppm\": %4d,", (long)time(NULL), st.snr_est, (int)fsk->ppm);
            const float *fe = fsk->freq_est_type ? fsk->f2_est : fsk->f_est;
            fprintf(stderr, " \"f1_est\":%.1f, \"f2_est\":%.1f", fe[0], fe[1]);
            if (fsk->mode == 4) fprintf(stderr, ", \"f3_est\":%.1f, \"f4_est\":%.1f", fe[2], fe[3]);
            fprintf(stderr, ", \"eye_diagram\":[], \"samp_fft\":[]}\n");
        }
    }
    fsk_destroy(fsk);
    fclose(fout);
    return 0;
}
