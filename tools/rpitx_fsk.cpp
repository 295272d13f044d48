/*
 * rpitx_fsk — IQ-file build of pirip's transmitter: everything tx/rpitx_fsk.cpp does to the bits
 * — CRC16 insert [R tx/rpitx_fsk.cpp:75-83], FSK_LDPC framing through the freedv calls it declares [R :34-39,317-325,395],
 * preamble [R :313-336], bursts with burst-control bytes 1 / 0 / 2 from stdin [R :429-497], --testframes / --bursts /
 * --seq / --source [R :369-423] — but the M-FSK waveform goes to a u8 IQ file (the format rtl_sdr produces) instead of
 * the Raspberry Pi's PLL (librpitx, GPIO and FIFO handling are hardware and out of scope).  With it the receive chain
 * can be driven from pirip's own command lines:
 *
 *   rpitx_fsk /dev/zero --code H_256_512_4 -r 10000 -s 10000 --testframes 3 --bursts 2 -o burst.u8     [R README.md:191]
 *   fsk_get_test_bits - 10000 | rpitx_fsk - -r 10000 -s 10000 -o test.u8                                [R README.md:101]
 *   rtl_fsk ... -b | frame_repeater 256 0x2 | rpitx_fsk - --code H_256_512_4 -r 10000 -s 10000 --packed -o rpt.u8
 *                                                                                                        [R script/frame_repeater:36-38]
 * Same options as the reference (-m -r -s -t -c --code --listcodes --testframes --bursts --delay --seq --packed --source;
 * -f -g --fifo are accepted and ignored), plus where the samples go:
 *   -o file       output u8 I,Q file ("-" = stdout)          --fs Hz      sample rate of the file (default 8 * Rs)
 *   --f1 Hz       frequency of tone 0 as the receiver sees it (default = shift: the rx is tuned one shift low)
 *   --ebno dB     add complex AWGN for this Eb/N0 (uncoded bit energy)    --amp A   peak amplitude in LSB (default 45)
 *   --lead ms     silence before the first burst (default 2 frame times)   --seed N  noise seed
 * Host-only code: uses the framer / CRC / pack entry points of libpirip_b200.so, no kernel.
 */
#include <getopt.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <vector>

#include "../include/pirip_ldpc.h"

namespace {

struct Wave {
    double fs, rs, f1, shift, amp;
    int M;
    double phase = 0.0, sigma = 0.0;
    bool carrier = false;
    std::mt19937_64 rng;
    std::normal_distribution<double> gauss{ 0.0, 1.0 };
    FILE *out = nullptr;
    double tsym_acc = 0.0; /* fractional samples carried between symbols */
    std::vector<unsigned char> buf;

    void sample(double re, double im) {
        if (sigma > 0.0) {
            re += sigma * gauss(rng);
            im += sigma * gauss(rng);
        }
        const double a = 127.5 + amp * re, b = 127.5 + amp * im;
        buf.push_back((unsigned char)(a < 0 ? 0 : a > 255 ? 255 : lrint(a)));
        buf.push_back((unsigned char)(b < 0 ? 0 : b > 255 ? 255 : lrint(b)));
        if (buf.size() >= 1 << 16) flush();
    }
    void flush() {
        fwrite(buf.data(), 1, buf.size(), out);
        buf.clear();
    }
    /* one symbol of continuous-phase FSK: tone index = symbol value, ascending frequency [R tx/rpitx_fsk.cpp:127-143] */
    void symbol(int sym) {
        tsym_acc += fs / rs;
        const int n = (int)floor(tsym_acc);
        tsym_acc -= n;
        const double dph = 2.0 * M_PI * (f1 + shift * sym) / fs;
        for (int i = 0; i < n; ++i) {
            phase += dph;
            if (phase > M_PI) phase -= 2.0 * M_PI;
            sample(cos(phase), sin(phase));
        }
    }
    void silence(double seconds) {
        const long n = lrint(seconds * fs);
        for (long i = 0; i < n; ++i) sample(0.0, 0.0);
    }
    /* log2(M) bits MSB first -> symbol [R tx/rpitx_fsk.cpp:133-138] */
    void frame(const uint8_t *bits, int nbits) {
        const int bps = M == 2 ? 1 : 2;
        for (int i = 0; i + bps <= nbits; i += bps) {
            int sym = 0;
            for (int k = 0; k < bps; ++k) sym = (sym << 1) | (bits[i + k] & 1);
            symbol(sym);
        }
    }
};

/* last 16 data bits <- CRC16 of the packed bytes before them, big endian [R tx/rpitx_fsk.cpp:75-83] */
void insert_crc(uint8_t *bits, int nbits) {
    std::vector<unsigned char> bytes((size_t)nbits / 8);
    freedv_pack(bytes.data(), bits, nbits - 16);
    const unsigned short crc = freedv_gen_crc16(bytes.data(), nbits / 8 - 2);
    unsigned char tail[2] = { (unsigned char)(crc >> 8), (unsigned char)(crc & 0xff) };
    freedv_unpack(bits + nbits - 16, tail, 16);
}

void usage(const char *p) {
    fprintf(stderr,
            "usage: %s [-m 2|4] [-r symbRateHz] [-s shiftHz] [-t] [-c] [--code CodeName] [--listcodes] [--testframes N] [--bursts B]\n"
            "          [--delay ms] [--seq] [--packed] [--source Byte] [-o iqFile] [--fs Hz] [--f1 Hz] [--ebno dB] [--amp A]\n"
            "          [--lead ms] [--seed N] InputOneBitPerCharFile\n", p);
    exit(1);
}

}  // namespace

int main(int argc, char *argv[]) {
    int Rs = 10000, M = 2, testframes = 0, Nframes = 0, Nbursts = 1, seq = 0, packed = 0, delay_ms = 0, carrier_test = 0, one_zero = 0;
    int source_byte = 0, lead_ms = -1;
    double shift = -1, fs = 0, f1 = -1, ebno = 1e9, amp = 45.0;
    unsigned long seed = 1;
    const char *code = nullptr, *outname = "-";
    static struct option lo[] = { { "code", required_argument, 0, 'a' },   { "listcodes", no_argument, 0, 'b' },
                                  { "testframes", required_argument, 0, 'u' }, { "bursts", required_argument, 0, 'e' },
                                  { "seq", no_argument, 0, 'q' },          { "fifo", required_argument, 0, 'i' },
                                  { "packed", no_argument, 0, 'l' },       { "source", required_argument, 0, 'd' },
                                  { "delay", required_argument, 0, 'j' },  { "fs", required_argument, 0, 'F' },
                                  { "f1", required_argument, 0, 'T' },     { "ebno", required_argument, 0, 'E' },
                                  { "amp", required_argument, 0, 'A' },    { "lead", required_argument, 0, 'L' },
                                  { "seed", required_argument, 0, 'S' },   { 0, 0, 0, 0 } };
    int o;
    while ((o = getopt_long(argc, argv, "a:bcd:e:f:g:i:m:o:qr:s:tu:h", lo, NULL)) != -1) {
        switch (o) {
            case 'a': code = optarg; break;
            case 'b': ldpc_codes_list(); return 0;
            case 'c': carrier_test = 1; break;
            case 'd': source_byte = (int)strtol(optarg, NULL, 0); break;
            case 'e': Nbursts = atoi(optarg); break;
            case 'm': M = atoi(optarg); break;
            case 'l': packed = 1; break;
            case 'q': seq = 1; break;
            case 'r': Rs = (int)atof(optarg); break;
            case 's': shift = atof(optarg); break;
            case 't': one_zero = 1; break;
            case 'u': testframes = 1; Nframes = atoi(optarg); break;
            case 'j': delay_ms = atoi(optarg); break;
            case 'o': outname = optarg; break;
            case 'F': fs = atof(optarg); break;
            case 'T': f1 = atof(optarg); break;
            case 'E': ebno = atof(optarg); break;
            case 'A': amp = atof(optarg); break;
            case 'L': lead_ms = atoi(optarg); break;
            case 'S': seed = strtoul(optarg, NULL, 0); break;
            case 'f': case 'g': case 'i': break; /* carrier frequency, antenna GPIO, stats FIFO: Pi hardware */
            default: usage(argv[0]);
        }
    }
    if (testframes && !code) {
        fprintf(stderr, "internal testframe mode only supported in --code coded mode!\n");
        return 1;
    }
    if (!(carrier_test || one_zero) && argc - optind < 1) usage(argv[0]);
    if (M != 2 && M != 4) usage(argv[0]);
    if (shift < 0) shift = 2.0 * Rs; /* [R tx/rpitx_fsk.cpp:347-348] */
    if (fs <= 0) fs = 8.0 * Rs;
    if (f1 < 0) f1 = shift;
    FILE *fin = stdin;
    if (!(carrier_test || one_zero) && strcmp(argv[optind], "-") != 0) {
        fin = fopen(argv[optind], "rb");
        if (!fin) {
            fprintf(stderr, "Error opening input file: %s\n", argv[optind]);
            return 1;
        }
    }
    Wave w;
    w.fs = fs, w.rs = Rs, w.f1 = f1, w.shift = shift, w.amp = amp, w.M = M;
    w.rng.seed(seed);
    w.out = strcmp(outname, "-") == 0 ? stdout : fopen(outname, "wb");
    if (!w.out) {
        fprintf(stderr, "Error opening output file: %s\n", outname);
        return 1;
    }
    if (ebno < 1e8) { /* unit-amplitude complex signal: Es = Ts samples, N0 = complex noise variance */
        const double esno = pow(10.0, ebno / 10.0) * (M == 2 ? 1.0 : 2.0);
        w.sigma = sqrt((fs / Rs) / esno / 2.0);
    }

    const int npre_sym = 50 * (M >> 1), npre_bits = npre_sym * (M >> 1);
    std::vector<uint8_t> preamble((size_t)npre_bits);
    struct freedv *fdv = nullptr;
    int K = M == 2 ? 1 : 2, nframe_bits = K;
    if (code) {
        struct freedv_advanced adv;
        memset(&adv, 0, sizeof adv);
        adv.codename = (char *)code;
        adv.Rs = Rs;
        adv.Fs = 8 * Rs; /* only to satisfy the FSK_LDPC open, as in the reference: no demodulator here */
        adv.M = M;
        fdv = freedv_open_advanced(FREEDV_MODE_FSK_LDPC, &adv);
        if (!fdv) {
            fprintf(stderr, "rpitx_fsk: unknown LDPC code %s\n", code);
            return 1;
        }
        K = freedv_get_bits_per_modem_frame(fdv);
        nframe_bits = freedv_tx_fsk_ldpc_bits_per_frame(fdv);
        int sym = 0; /* cycle through the symbols: bit pairs 00 01 10 11 00 ... */
        for (int i = 0; i + 1 < npre_bits; i += 2) {
            preamble[(size_t)i] = (sym >> 1) & 1;
            preamble[(size_t)i + 1] = sym & 1;
            sym += 1;
        }
        fprintf(stderr, "rpitx_fsk: FSK LDPC mode code: %s data_bits_per_frame: %d\n", code, K);
    }
    std::vector<uint8_t> data((size_t)K), txf((size_t)nframe_bits);
    const double frame_s = (double)nframe_bits / (M >> 1) / Rs;
    const double gap_s = delay_ms ? delay_ms / 1e3 : 2.0 * frame_s; /* [R tx/rpitx_fsk.cpp:415-418] */
    w.silence(lead_ms >= 0 ? lead_ms / 1e3 : 2.0 * frame_s);

    if (carrier_test || one_zero) {
        for (int i = 0; i < 4 * Rs; ++i) w.symbol(carrier_test ? 0 : i & 1);
    } else if (testframes) {
        ofdm_generate_payload_data_bits(data.data(), K);
        for (int b = 0; b < Nbursts; ++b) {
            w.frame(preamble.data(), npre_bits);
            for (int f = 0; f < Nframes; ++f) {
                if (source_byte)
                    for (int i = 0; i < 8; ++i) data[(size_t)i] = (source_byte >> (7 - i)) & 1;
                if (seq) {
                    const int s = (f + 1) & 0xff;
                    for (int i = 0; i < 8; ++i) data[8 + (size_t)i] = (s >> (7 - i)) & 1;
                }
                insert_crc(data.data(), K);
                freedv_tx_fsk_ldpc_framer(fdv, txf.data(), data.data());
                w.frame(txf.data(), nframe_bits);
            }
            fprintf(stderr, "rpitx_fsk: End of burst %d\n", b);
            w.silence(gap_s);
        }
    } else {
        for (;;) {
            uint8_t burst_control = 0;
            if (code && fread(&burst_control, 1, 1, fin) != 1) break;
            size_t got;
            if (packed) {
                std::vector<unsigned char> bytes((size_t)K / 8);
                got = fread(bytes.data(), 1, bytes.size(), fin) * 8;
                freedv_unpack(data.data(), bytes.data(), K);
            } else {
                got = fread(data.data(), 1, (size_t)K, fin);
            }
            if (got != (size_t)K) break;
            if (!code) { /* uncoded: the bits as they come */
                w.frame(data.data(), K);
                continue;
            }
            if (burst_control == 1) w.frame(preamble.data(), npre_bits); /* start of burst */
            if (burst_control == 0 || burst_control == 1) {
                insert_crc(data.data(), K); /* the last two bytes of the frame are replaced by the CRC */
                freedv_tx_fsk_ldpc_framer(fdv, txf.data(), data.data());
                w.frame(txf.data(), nframe_bits);
            }
            if (burst_control == 2) w.silence(gap_s); /* end of burst: dummy frame, carrier off */
        }
        w.silence(gap_s);
    }
    w.flush();
    if (w.out != stdout) fclose(w.out);
    if (fdv) freedv_close(fdv);
    return 0;
}
