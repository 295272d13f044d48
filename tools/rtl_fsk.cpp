/*
 * rtl_fsk — file-source build of the integrated receiver pirip runs everywhere ([R README.md:114,172,184,292],
 * [R test/loopback_rtl_fsk.sh:10], [R script/frame_repeater:36], [R script/ping:47]): u8 IQ in -> optional csdr
 * decimation (-s -> -a) -> fsk_demod, or with --code the FSK_LDPC receive side (fsk_demod_sd -> LLRs -> unique-word
 * sync -> LDPC -> CRC) -> stdout.  The USB tuner of drowe67/librtlsdr's rtl_fsk (out of scope) is
 * replaced by `-i file` ("-" = stdin, the byte stream `rtl_sdr` writes); everything downstream of the samples keeps
 * rtl_fsk's command line and output formats, so `rtl_fsk ... -b | frame_repeater 256 0x2` and
 * `rtl_fsk ... | fsk_put_test_bits -` run as in pirip's scripts:
 *
 *   rtl_fsk [-i iqfile] [-s rtlFs] [-a modemFs] -r Rs -m M [-p P] [--code NAME] [--mask spacing]
 *           [--fsk_lower Hz] [--fsk_upper Hz] [--filter addr] [--testframes] [-b] [-q] [-v] [-u host] [-n samples]
 *           [-f Hz] [-g gain] [-w bw] [-e reg] [-L logfile] out
 *
 *   uncoded          Nbits bytes per modem frame, one bit per byte                       [R README.md:114]
 *   --code NAME      packed payload bytes of every frame whose CRC checks                [R README.md:184,297]
 *   -b               records of 1 rx_status byte + K/8 payload bytes at EVERY modem frame, zeros when there is no
 *                    frame — the clock of tx/frame_repeater.c                            [R tx/frame_repeater.c:55-62]
 *   --filter addr    drop frames whose first payload byte (source address) equals addr   [R README.md:303-304]
 *   -v               per-frame log line with the fields of                               [R README.md:200-208]
 *   --chunk K        modem frames per ingest chunk and kernel launch (default 32; 1 = one frame per call as fsk_demod.c)
 *   -u host          once per second of signal one JSON line over UDP to host:8001 with the keys script/dash.py reads
 *                    (SfdB, f_est_Hz, fsk_lower_Hz, fsk_upper_Hz, SNRest_lin, norm_rx_timing, Fs_Hz) [R script/dash.py:26-46]
 *
 * All arithmetic runs in libpirip_b200.so (GPU); there is no CPU fallback.
 */
#include <arpa/inet.h>
#include <getopt.h>
#include <math.h>
#include <netdb.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../include/pirip_csdr.h"
#include "../include/pirip_fsk.h"
#include "../include/pirip_ldpc.h"

static void usage(const char *p) {
    fprintf(stderr,
            "usage: %s [-i iqfile] [-s rtlFs] [-a modemFs] -r Rs -m M [-p P] [--code NAME] [--mask spacing] [--fsk_lower Hz]\n"
            "          [--fsk_upper Hz] [--filter addr] [--testframes] [-b] [-q] [-v] [-u host] [-n samples] [-f Hz] [-g gain] [-w bw] [-e reg]\n"
            "          [-L logfile] OutputFile\n",
            p);
    exit(1);
}

int main(int argc, char *argv[]) {
    int Fs_rtl = 0, Fs = 0, Rs = 10000, M = 2, P = 0, lower = -1, upper = -1, mask = 0, filter_addr = -1;
    int burst_records = 0, quiet = 0, verbose = 0, testframes = 0;
    int chunk_frames = 32; /* --chunk: modem frames per ingest chunk = per launch */
    long long max_samples = -1; /* -n: stop after this many input samples (at the tuner rate) [R README.md:114] */
    const char *code = nullptr, *inname = "-", *dash_host = nullptr;
    static struct option lo[] = { { "code", required_argument, 0, 'c' },      { "mask", required_argument, 0, 'k' },
                                  { "fsk_lower", required_argument, 0, 'l' }, { "fsk_upper", required_argument, 0, 'U' },
                                  { "filter", required_argument, 0, 'F' },    { "testframes", no_argument, 0, 't' },
                                  { "chunk", required_argument, 0, 'C' },     { "help", no_argument, 0, 'h' },
                                  { 0, 0, 0, 0 } };
    int o;
    while ((o = getopt_long(argc, argv, "i:s:a:r:m:p:f:g:u:n:e:w:L:d:bqvh", lo, NULL)) != -1) {
        switch (o) {
            case 'i': inname = optarg; break;
            case 's': Fs_rtl = atoi(optarg); break;
            case 'a': Fs = atoi(optarg); break;
            case 'r': Rs = atoi(optarg); break;
            case 'm': M = atoi(optarg); break;
            case 'p': P = atoi(optarg); break;
            case 'c': code = optarg; break;
            case 'k': mask = atoi(optarg); break;
            case 'l': lower = atoi(optarg); break;
            case 'U': upper = atoi(optarg); break;
            case 'F': filter_addr = (int)strtol(optarg, NULL, 0); break;
            case 't': testframes = 1; break;
            case 'C': chunk_frames = atoi(optarg); break;
            case 'b': burst_records = 1; break;
            case 'q': quiet = 1; break;
            case 'v': verbose = 1; break;
            case 'u': dash_host = optarg; break;
            case 'n': max_samples = (long long)atof(optarg); break;
            /* tuner frequency / gain / device, tuner bandwidth (-w) and register tweak (-e) [R README.md:152,172,239], log file
               (-L file [R script/frame_repeater:23]): accepted as upstream spells them, nothing to tune in a file */
            case 'f': case 'g': case 'd': case 'e': case 'w': case 'L': break;
            default: usage(argv[0]);
        }
    }
    if (argc - optind < 1) usage(argv[0]);
    const char *outname = argv[optind];
    if (Fs_rtl == 0 && Fs == 0) Fs_rtl = 240000;
    if (Fs == 0) {
        /* no -a: the modem runs at the tuner rate ([R README.md:114]) - unless that puts more than 100 samples into a symbol
           (`-r 1000` at the default 240 kHz, [R README.md:152,184]: a 12000-sample modem frame and a 4096-point estimator, beyond
           what one capture group of the GPU kernels holds).  Then the csdr stage in front decimates to the lowest rate that keeps
           40 samples per symbol, the rate pirip's own scripts pick for that symbol rate (-a 40000 -r 1000, [R script/ping:47]). */
        Fs = Fs_rtl;
        if (Rs > 0 && Fs_rtl / Rs > 100) {
            for (int D2 = Fs_rtl / (40 * Rs); D2 >= 1; --D2)
                if (Fs_rtl % D2 == 0 && (Fs_rtl / D2) % Rs == 0 && Fs_rtl / D2 >= 40 * Rs) {
                    Fs = Fs_rtl / D2;
                    break;
                }
            if (Fs != Fs_rtl && !quiet) fprintf(stderr, "rtl_fsk: modem sample rate %d Hz (tuner %d Hz / %d); -a overrides\n", Fs, Fs_rtl, Fs_rtl / Fs);
        }
    }
    if (Fs_rtl == 0) Fs_rtl = Fs;
    if ((M != 2 && M != 4) || Rs <= 0 || Fs % Rs || Fs_rtl % Fs) usage(argv[0]);
    const int D = Fs_rtl / Fs; /* csdr decimation ratio, e.g. 1.8M -> 40k = 45 [R README.md:109], 2.4M -> 80k = 30 [R README.md:172] */
    if (P == 0) {
        /* upstream takes no -p: the oversampling is derived from Ts the way the FSK_LDPC modem open does ([U freedv_fsk.c::
           freedv_fsk_open], M-confidence: Ts halved while it is even and above 10): Ts 24 -> 6, 20 -> 10, 40 -> 10, 240 -> 15, 8 -> 8 */
        P = Fs / Rs;
        while (P > 10 && (P % 2) == 0) P /= 2;
    }
    if ((Fs / Rs) % P || P < 4) {
        fprintf(stderr, "rtl_fsk: Ts/P must be an integer and P >= 4\n");
        return 1;
    }
    FILE *fin = strcmp(inname, "-") == 0 ? stdin : fopen(inname, "rb");
    FILE *fout = strcmp(outname, "-") == 0 ? stdout : fopen(outname, "wb");
    if (!fin || !fout) {
        fprintf(stderr, "rtl_fsk: couldn't open files\n");
        return 1;
    }

    struct FSK *fsk = fsk_create_hbr(Fs, Rs, M, P, FSK_DEFAULT_NSYM, 1200, mask ? mask : 400);
    /* coded mode follows the FSK_LDPC modem's estimator window (0 .. Fs/2, [U freedv_fsk.c]) unless told otherwise */
    if (lower >= 0 || upper >= 0 || code) {
        if (lower < 0) lower = 0;
        if (upper < 0) upper = Fs / 2;
        fsk_set_freq_est_limits(fsk, lower, upper);
    }
    if (mask) fsk_set_freq_est_alg(fsk, 1);

    struct freedv *fdv = nullptr;
    int K_bits = 0;
    if (code) {
        struct freedv_advanced adv;
        memset(&adv, 0, sizeof adv);
        adv.codename = (char *)code;
        adv.Rs = Rs;
        adv.Fs = Fs;
        adv.M = M;
        fdv = freedv_open_advanced(FREEDV_MODE_FSK_LDPC, &adv);
        if (!fdv) {
            fprintf(stderr, "rtl_fsk: unknown LDPC code %s\n", code);
            return 1;
        }
        K_bits = freedv_get_bits_per_modem_frame(fdv);
    }
    std::vector<uint8_t> test_payload((size_t)(K_bits > 0 ? K_bits : 1));
    if (testframes && K_bits) ofdm_generate_payload_data_bits(test_payload.data(), K_bits);

    /* csdr front end: convert_u8_f + fir_decimate_cc D with the CLI's default taps (transition 0.05, Hamming) */
    std::vector<float> taps;
    if (D > 1) {
        taps.resize((size_t)firdes_filter_len(0.05f));
        firdes_lowpass_f(taps.data(), (int)taps.size(), 0.5f / (float)D, 1 /* WINDOW_HAMMING */);
    }
    const int L = (int)taps.size();

    int sock = -1;
    struct sockaddr_in dash;
    memset(&dash, 0, sizeof dash);
    if (dash_host) {
        struct hostent *he = gethostbyname(dash_host);
        sock = socket(AF_INET, SOCK_DGRAM, 0);
        if (he && sock >= 0) {
            dash.sin_family = AF_INET;
            dash.sin_port = htons(8001);
            memcpy(&dash.sin_addr, he->h_addr_list[0], (size_t)he->h_length);
        } else {
            fprintf(stderr, "rtl_fsk: can't reach dashboard host %s\n", dash_host);
            sock = -1;
        }
    }

    /* The sample source delivers CHUNKS (upstream: the USB callback's buffers; here: reads of K frames' worth of the file or
       pipe).  A chunk goes to the device once and pirip_fsk_demod_frames() demodulates every whole frame in it in ONE launch
       (SURVEY.md 8f N2); what is left over waits for the next chunk.  --chunk 1 is the frame-per-call loop of fsk_demod.c. */
    const int nmax = fsk->N + fsk->Ts * 2;
    const size_t K = (size_t)(chunk_frames < 1 ? 1 : chunk_frames);
    const size_t target = K * (size_t)fsk->N + (size_t)fsk->Ts;
    std::vector<COMP> cbuf;                 /* modem-rate samples not yet demodulated */
    cbuf.reserve(target + (size_t)nmax);
    std::vector<uint8_t> bits_all(K * (size_t)fsk->Nbits), payload((size_t)(K > 0 && code ? freedv_get_bits_per_modem_frame(fdv) : 1)),
        packed((size_t)(code ? freedv_get_bits_per_modem_frame(fdv) / 8 : 1));
    std::vector<float> filt_all(K * (size_t)(fsk->mode * fsk->Nsym)), llrs((size_t)fsk->Nbits);
    std::vector<pirip_frame_stats> st_all(K);
    std::vector<unsigned char> raw;       /* undecimated u8 samples not yet consumed */
    std::vector<float> fbuf;              /* their float form (csdr convention) */
    std::vector<complexf> dec;
    std::vector<float> timing_1s;
    long frames = 0, packets = 0, test_errs = 0, test_bits_n = 0;
    double since_dash = 0.0;
    long long consumed_in = 0; /* input samples (tuner rate) taken from the source so far */
    bool eof = false;

    /* append up to `want` modem-rate samples to cbuf; fewer at the end of the input */
    auto pull = [&](size_t want) {
        if (eof || want == 0) return;
        if (D == 1) {
            if (max_samples >= 0 && consumed_in + (long long)want > max_samples) want = (size_t)(max_samples - consumed_in), eof = true;
            raw.resize(want * 2);
            const size_t got = fread(raw.data(), 2, want, fin);
            if (got < want) eof = true;
            consumed_in += (long long)got;
            const size_t o = cbuf.size();
            cbuf.resize(o + got);
            for (size_t i = 0; i < got; ++i) { /* the fsk_demod -d convention [U fsk_demod.c] */
                cbuf[o + i].real = ((float)raw[2 * i] - 127.0) / 128.0;
                cbuf[o + i].imag = ((float)raw[2 * i + 1] - 127.0) / 128.0;
            }
        } else {
            /* one causal FIR over the continuous stream, phase anchored at stream sample 0 (csdr main loop):
               n outputs need D*(n-1)+L inputs; D*n of them are consumed */
            if (max_samples >= 0 && consumed_in + (long long)(D * want) > max_samples) want = (size_t)((max_samples - consumed_in) / D), eof = true;
            if (want == 0) return;
            const size_t need = (size_t)D * (want - 1) + (size_t)L;
            const size_t have = raw.size() / 2;
            if (have < need) {
                raw.resize(need * 2);
                const size_t got = fread(raw.data() + have * 2, 2, need - have, fin);
                if (got < need - have) {
                    eof = true;
                    raw.resize((have + got) * 2);
                }
            }
            const size_t avail = raw.size() / 2;
            if (avail < (size_t)L) return;
            const size_t nout = (avail - (size_t)L) / (size_t)D + 1 < want ? (avail - (size_t)L) / (size_t)D + 1 : want;
            const size_t nin_f = (size_t)D * (nout - 1) + (size_t)L;
            fbuf.resize(nin_f * 2);
            dec.resize(nout + 1);
            convert_u8_f(raw.data(), fbuf.data(), (int)(nin_f * 2));
            const int got = fir_decimate_cc((complexf *)fbuf.data(), dec.data(), (int)nin_f, D, taps.data(), L);
            const size_t o = cbuf.size();
            cbuf.resize(o + (size_t)got);
            for (int i = 0; i < got; ++i) {
                cbuf[o + (size_t)i].real = dec[(size_t)i].i;
                cbuf[o + (size_t)i].imag = dec[(size_t)i].q;
            }
            consumed_in += (long long)D * got;
            raw.erase(raw.begin(), raw.begin() + (std::ptrdiff_t)((size_t)D * (size_t)got * 2));
        }
    };

    for (;;) {
        if (cbuf.size() < target) pull(target - cbuf.size());
        if (cbuf.size() < (size_t)fsk_nin(fsk)) break; /* a short read ends the loop, as upstream */
        size_t used = 0;
        const int nf = pirip_fsk_demod_frames(fsk, cbuf.data(), cbuf.size(), (int)K, bits_all.data(), filt_all.data(), st_all.data(), &used);
        if (nf <= 0) break;
        for (int f = 0; f < nf; ++f) {
            const pirip_frame_stats &fs = st_all[(size_t)f];
            const uint8_t *bits = bits_all.data() + (size_t)f * (size_t)fsk->Nbits;
            const float *filt = filt_all.data() + (size_t)f * (size_t)(fsk->mode * fsk->Nsym);
            if (!code) {
                fwrite(bits, 1, (size_t)fsk->Nbits, fout);
            } else {
                fsk_rx_filt_to_llrs(llrs.data(), (float *)filt, fs.v_est, fs.SNRest, fsk->mode, fsk->Nsym);
                int st = freedv_fsk_ldpc_rx_llrs(fdv, llrs.data(), fsk->Nbits, payload.data());
                bool deliver = (st & FREEDV_RX_BITS) != 0;
                if (deliver) freedv_pack(packed.data(), payload.data(), K_bits);
                if (deliver && filter_addr >= 0 && packed[0] == (uint8_t)filter_addr) { /* our own repeater talking */
                    deliver = false;
                    st &= ~FREEDV_RX_BITS;
                }
                if (deliver) {
                    packets++;
                    if (testframes) {
                        for (int i = 0; i < K_bits - 16; ++i) test_errs += payload[(size_t)i] != test_payload[(size_t)i];
                        test_bits_n += K_bits - 16;
                    }
                }
                if (verbose && (st & (FREEDV_RX_BITS | FREEDV_RX_BIT_ERRORS))) {
                    int info[6];
                    freedv_fsk_ldpc_rx_info(fdv, info);
                    fprintf(stderr, "%3d nbits: %3d state: %d uw_loc: %3d uw_err: %2d snrdB: %4.1f iter: %2d pcc: %3d rxst: %c%c%c%c\n",
                            (int)packets, info[5], info[0], info[1], info[2], 10.0 * log10((double)fs.SNRest + 1.0), info[3], info[4],
                            st & FREEDV_RX_BIT_ERRORS ? 'E' : '-', st & FREEDV_RX_BITS ? 'B' : '-', st & FREEDV_RX_SYNC ? 'S' : '-',
                            st & FREEDV_RX_TRIAL_SYNC ? 'T' : '-');
                }
                if (burst_records) {
                    const uint8_t sb = (uint8_t)st;
                    if (!deliver) memset(packed.data(), 0, packed.size());
                    fwrite(&sb, 1, 1, fout);
                    fwrite(packed.data(), 1, (size_t)K_bits / 8, fout);
                } else if (deliver) {
                    fwrite(packed.data(), 1, (size_t)K_bits / 8, fout);
                }
            }
            frames++;
            /* dashboard telemetry, once per second of signal */
            timing_1s.push_back(fs.norm_rx_timing);
            since_dash += (double)(f + 1 < nf ? st_all[(size_t)f + 1].sample_offset - fs.sample_offset : fs.nin_next) / (double)Fs;
            if (sock >= 0 && since_dash >= 1.0) {
                std::string js = "{\"SfdB\":[";
                char num[256];
                for (int i = 0; i < fsk->Ndft; ++i) {
                    snprintf(num, sizeof num, "%s%.1f", i ? "," : "", 20.0 * log10((double)fsk->Sf[i] + 1e-9));
                    js += num;
                }
                js += "],\"f_est_Hz\":[";
                for (int m = 0; m < M; ++m) {
                    snprintf(num, sizeof num, "%s%.0f", m ? "," : "", (double)(fsk->freq_est_type ? fs.f2_est[m] : fs.f_est[m]));
                    js += num;
                }
                snprintf(num, sizeof num, "],\"fsk_lower_Hz\":%d,\"fsk_upper_Hz\":%d,\"Fs_Hz\":%d,\"SNRest_lin\":%.3f,\"norm_rx_timing\":[",
                         fsk->est_min, fsk->est_max, Fs, (double)fs.SNRest);
                js += num;
                for (size_t i = 0; i < timing_1s.size(); ++i) {
                    snprintf(num, sizeof num, "%s%.3f", i ? "," : "", (double)timing_1s[i]);
                    js += num;
                }
                js += "]}\n";
                sendto(sock, js.data(), js.size(), 0, (struct sockaddr *)&dash, sizeof dash);
                timing_1s.clear();
                since_dash = 0.0;
            }
        }
        if (fin == stdin || fout == stdout) fflush(fout);
        cbuf.erase(cbuf.begin(), cbuf.begin() + (std::ptrdiff_t)used);
    }
    if (!quiet) {
        fprintf(stderr, "rtl_fsk: %ld modem frames", frames);
        if (code) fprintf(stderr, ", %ld packets", packets);
        if (testframes && test_bits_n) fprintf(stderr, ", test frame BER %.4f (%ld/%ld)", (double)test_errs / (double)test_bits_n, test_errs, test_bits_n);
        fprintf(stderr, "\n");
    }
    if (fdv) freedv_close(fdv);
    fsk_destroy(fsk);
    fclose(fout);
    if (sock >= 0) close(sock);
    return 0;
}
