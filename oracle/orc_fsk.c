/*
 * orc_fsk.c — CPU oracle: FFT + FSK demodulator.  TEST INFRASTRUCTURE (see oracle.h).
 * PARITY UNPINNED: restatement of codec2 src/kiss_fft.c, src/fsk.c, src/fsk_demod.c
 * (un-vendored dependency of pirip, build_codec2.sh:3-5) from SURVEY.md 8a/App. A.
 *
 * Float discipline: every expression is evaluated in the C type the upstream
 * expression has (float fields, double literals such as .9 / 2*M_PI promote to
 * double), so the GPU kernel can mirror the rounding sequence op for op.
 * Compile with -O2 -ffp-contract=off and WITHOUT -ffast-math.
 */
#include "oracle.h"
#include "orc_math.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================
 * FFT  [U kiss_fft.c::kiss_fft_alloc / kf_factor / kf_work / kf_bfly2 / kf_bfly4]
 * On the path via fsk_demod_freq_est.
 * Only power-of-two sizes are ever requested (Ndft = 2^k), so the factor list
 * is 4,4,...,(2) and only the radix-4 and radix-2 butterflies are restated.
 * ====================================================================== */
#define ORC_FFT_MAXFACTORS 32
struct orc_fft {
    int nfft;
    int factors[2 * ORC_FFT_MAXFACTORS];
    orc_comp *tw;
};

static orc_comp c_mul(orc_comp a, orc_comp b) {
    orc_comp m;
    m.real = a.real * b.real - a.imag * b.imag;
    m.imag = a.real * b.imag + a.imag * b.real;
    return m;
}
static orc_comp c_add(orc_comp a, orc_comp b) { orc_comp r = { a.real + b.real, a.imag + b.imag }; return r; }
static orc_comp c_sub(orc_comp a, orc_comp b) { orc_comp r = { a.real - b.real, a.imag - b.imag }; return r; }

orc_fft *orc_fft_alloc(int nfft) {
    orc_fft *cfg = (orc_fft *)calloc(1, sizeof *cfg);
    cfg->nfft = nfft;
    cfg->tw = (orc_comp *)malloc(sizeof(orc_comp) * (size_t)nfft);
    for (int i = 0; i < nfft; ++i) {
        /* twiddles are evaluated in double and rounded to float once */
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / nfft;
        cfg->tw[i].real = (float)cos(phase);
        cfg->tw[i].imag = (float)sin(phase);
    }
    /* factor: peel 4s first, then 2, then odd factors */
    int n = nfft, p = 4, *f = cfg->factors;
    double root = floor(sqrt((double)n));
    do {
        while (n % p) {
            switch (p) {
                case 4: p = 2; break;
                case 2: p = 3; break;
                default: p += 2; break;
            }
            if (p > root) p = n;
        }
        n /= p;
        *f++ = p;
        *f++ = n;
    } while (n > 1);
    return cfg;
}

void orc_fft_free(orc_fft *cfg) {
    if (!cfg) return;
    free(cfg->tw);
    free(cfg);
}

void orc_fft_twiddles(const orc_fft *cfg, orc_comp *out) { memcpy(out, cfg->tw, sizeof(orc_comp) * (size_t)cfg->nfft); }

static void bfly2(orc_comp *F, int fstride, const orc_fft *cfg, int m) {
    orc_comp *F2 = F + m;
    const orc_comp *tw = cfg->tw;
    for (int k = 0; k < m; ++k) {
        orc_comp t = c_mul(F2[k], *tw);
        tw += fstride;
        F2[k] = c_sub(F[k], t);
        F[k] = c_add(F[k], t);
    }
}

static void bfly4(orc_comp *F, int fstride, const orc_fft *cfg, int m) {
    const orc_comp *tw1 = cfg->tw, *tw2 = cfg->tw, *tw3 = cfg->tw;
    const int m2 = 2 * m, m3 = 3 * m;
    for (int k = 0; k < m; ++k) {
        orc_comp s0 = c_mul(F[k + m], *tw1);
        orc_comp s1 = c_mul(F[k + m2], *tw2);
        orc_comp s2 = c_mul(F[k + m3], *tw3);
        orc_comp s5 = c_sub(F[k], s1);
        F[k] = c_add(F[k], s1);
        orc_comp s3 = c_add(s0, s2);
        orc_comp s4 = c_sub(s0, s2);
        F[k + m2] = c_sub(F[k], s3);
        tw1 += fstride;
        tw2 += fstride * 2;
        tw3 += fstride * 3;
        F[k] = c_add(F[k], s3);
        /* forward transform */
        F[k + m].real = s5.real + s4.imag;
        F[k + m].imag = s5.imag - s4.real;
        F[k + m3].real = s5.real - s4.imag;
        F[k + m3].imag = s5.imag + s4.real;
    }
}

static void fft_work(orc_comp *Fout, const orc_comp *f, int fstride, const int *factors, const orc_fft *cfg) {
    const int p = factors[0], m = factors[1];
    if (m == 1) {
        for (int k = 0; k < p; ++k) Fout[k] = f[k * fstride];
    } else {
        for (int k = 0; k < p; ++k) fft_work(Fout + k * m, f + k * fstride, fstride * p, factors + 2, cfg);
    }
    switch (p) {
        case 2: bfly2(Fout, fstride, cfg, m); break;
        case 4: bfly4(Fout, fstride, cfg, m); break;
        default: assert(0 && "oracle FFT: power-of-two sizes only");
    }
}

void orc_fft_run(const orc_fft *cfg, const orc_comp *in, orc_comp *out) {
    assert(in != out);
    fft_work(out, in, 1, cfg->factors, cfg);
}

/* ======================================================================
 * complex helpers  [U comp_prim.h]
 * ====================================================================== */
static orc_comp comp_exp_j(float phi) {
    orc_comp r;
    r.real = om_cosf(phi);
    r.imag = om_sinf(phi);
    return r;
}
static orc_comp comp0(void) { orc_comp r = { 0.0f, 0.0f }; return r; }
static orc_comp cconj(orc_comp a) { orc_comp r = { a.real, -a.imag }; return r; }
static orc_comp fcmult(float a, orc_comp b) { orc_comp r = { a * b.real, a * b.imag }; return r; }
static orc_comp comp_normalize(orc_comp a) {
    float av = sqrtf((a.real * a.real) + (a.imag * a.imag));
    orc_comp r = { a.real / av, a.imag / av };
    return r;
}

/* ======================================================================
 * struct FSK lifecycle  [U fsk.c::fsk_create_core / fsk_create_hbr / fsk_create]
 * SURVEY.md 8a row A6, App. A.1.  Reached from [R README.md:105] through the CLI.
 * ====================================================================== */
static void generate_hann_table(struct orc_fsk *fsk) {
    /* [U fsk.c::fsk_generate_hann_table] recursive oscillator form of
       0.5 - 0.5 cos(2 pi i / (Ndft-1)) */
    int Ndft = fsk->Ndft;
    orc_comp dphi = comp_exp_j((float)((2 * M_PI) / ((float)Ndft - 1)));
    orc_comp rphi = { .5f, 0.0f };
    rphi = c_mul(cconj(dphi), rphi);
    for (int i = 0; i < Ndft; ++i) {
        rphi = c_mul(dphi, rphi);
        fsk->hann_table[i] = (float)(.5 - rphi.real);
    }
}

struct orc_fsk *orc_fsk_create_hbr(int Fs, int Rs, int M, int P, int Nsym, int f1_tx, int tone_spacing) {
    assert(Fs > 0);
    assert(Rs > 0);
    assert(P > 0);
    assert(Nsym > 0);
    assert((Fs % Rs) == 0);        /* Ts must be an integer */
    assert(((Fs / Rs) % P) == 0);  /* Ts/P must be an integer */
    assert(P >= 4);                /* enough timing offsets to choose from */
    assert(M == 2 || M == 4);

    struct orc_fsk *fsk = (struct orc_fsk *)calloc(1, sizeof *fsk);
    assert(fsk);
    /* bins within 10% of the symbol rate, rounded up to a power of two */
    float bin_width_Hz = 0.1 * Rs;
    float Ndft = (float)Fs / bin_width_Hz;
    Ndft = pow(2.0, ceil(log2(Ndft)));

    fsk->Fs = Fs;
    fsk->Rs = Rs;
    fsk->Ts = Fs / Rs;
    fsk->burst_mode = 0;
    fsk->P = P;
    fsk->Nsym = Nsym;
    fsk->N = fsk->Ts * fsk->Nsym;
    fsk->Ndft = (int)Ndft;
    fsk->tc = 0.1;
    fsk->Nmem = fsk->N + (2 * fsk->Ts);
    fsk->f1_tx = f1_tx;
    fsk->tone_spacing = tone_spacing;
    fsk->nin = fsk->N;
    fsk->lock_nin = 0;
    fsk->mode = M;
    fsk->Nbits = M == 2 ? fsk->Nsym : fsk->Nsym * 2;
    fsk->est_min = 0;
    fsk->est_max = Fs;
    fsk->est_space = 0.75 * Rs;
    fsk->freq_est_type = 0;

    for (int i = 0; i < M; ++i) fsk->phi_c[i] = comp_exp_j(0);
    fsk->f_dc = (orc_comp *)malloc(sizeof(orc_comp) * (size_t)(M * fsk->Nmem));
    for (int i = 0; i < M * fsk->Nmem; ++i) fsk->f_dc[i] = comp0();
    fsk->fft_cfg = orc_fft_alloc(fsk->Ndft);
    fsk->Sf = (float *)malloc(sizeof(float) * (size_t)fsk->Ndft);
    for (int i = 0; i < fsk->Ndft; ++i) fsk->Sf[i] = 0;
    fsk->hann_table = (float *)malloc(sizeof(float) * (size_t)fsk->Ndft);
    generate_hann_table(fsk);

    fsk->norm_rx_timing = 0;
    fsk->tx_phase_c = comp_exp_j(0);
    fsk->EbNodB = 0;
    for (int i = 0; i < M; ++i) fsk->f_est[i] = 0;
    fsk->ppm = 0;
    fsk->stats_snr_est = 0;
    fsk->stats_foff = 0;
    fsk->stats_rx_timing = 0;
    fsk->stats_clock_offset = 0;
    fsk->normalise_eye = 1;
    return fsk;
}

struct orc_fsk *orc_fsk_create(int Fs, int Rs, int M, int tx_f1, int tx_fs) {
    /* FSK_DEFAULT_P = 8, FSK_DEFAULT_NSYM = 50 */
    return orc_fsk_create_hbr(Fs, Rs, M, 8, 50, tx_f1, tx_fs);
}

void orc_fsk_destroy(struct orc_fsk *fsk) {
    if (!fsk) return;
    free(fsk->f_dc);
    free(fsk->Sf);
    free(fsk->hann_table);
    orc_fft_free(fsk->fft_cfg);
    free(fsk);
}

uint32_t orc_fsk_nin(struct orc_fsk *fsk) { return (uint32_t)fsk->nin; }

void orc_fsk_set_freq_est_limits(struct orc_fsk *fsk, int est_min, int est_max) {
    assert(fsk != NULL);
    assert(est_min >= -fsk->Fs / 2);
    assert(est_max <= fsk->Fs / 2);
    assert(est_max > est_min);
    fsk->est_min = est_min;
    fsk->est_max = est_max;
}

void orc_fsk_set_freq_est_alg(struct orc_fsk *fsk, int est_type) {
    assert(fsk != NULL);
    fsk->freq_est_type = est_type;
}

void orc_fsk_clear_estimators(struct orc_fsk *fsk) {
    for (int i = 0; i < fsk->Ndft; ++i) fsk->Sf[i] = 0;
    fsk->nin = fsk->N;
}

void orc_fsk_enable_burst_mode(struct orc_fsk *fsk) {
    fsk->burst_mode = 1;
    fsk->nin = fsk->N;
}

/* ======================================================================
 * tone frequency estimator  [U fsk.c::fsk_demod_freq_est]
 * SURVEY.md 8a row A7, App. A.2.  `--fsk_lower/--fsk_upper` [R README.md:105],
 * `--mask` [R README.md:239].
 * ====================================================================== */
static void demod_freq_est(struct orc_fsk *fsk, const orc_comp fsk_in[], float *freqs, int M) {
    const int Ndft = fsk->Ndft;
    const int Fs = fsk->Fs;
    const int nin = fsk->nin;
    int freqi[ORC_MODE_M_MAX];

    orc_comp *fftin = (orc_comp *)malloc(sizeof(orc_comp) * (size_t)Ndft);
    orc_comp *fftout = (orc_comp *)malloc(sizeof(orc_comp) * (size_t)Ndft);
    for (int i = 0; i < Ndft; ++i) fftout[i] = comp0();

    int st = (fsk->est_min * Ndft) / Fs + Ndft / 2;
    if (st < 0) st = 0;
    int en = (fsk->est_max * Ndft) / Fs + Ndft / 2;
    if (en > Ndft) en = Ndft;
    int f_zero = (fsk->est_space * Ndft) / Fs;

    int numffts = (int)floor((float)nin / (Ndft / 2)) - 1;
    for (int j = 0; j < numffts; ++j) {
        int a = j * Ndft / 2;
        /* Hann-window one Ndft-sample block, 50% overlap between blocks */
        for (int i = 0; i < Ndft; ++i) {
            float hann = fsk->hann_table[i];
            fftin[i].real = hann * fsk_in[i + a].real;
            fftin[i].imag = hann * fsk_in[i + a].imag;
        }
        orc_fft_run(fsk->fft_cfg, fftin, fftout);
        /* FFT shift: DC to bin Ndft/2 */
        for (int i = 0; i < Ndft / 2; ++i) {
            orc_comp tmp = fftout[i];
            fftout[i] = fftout[i + Ndft / 2];
            fftout[i + Ndft / 2] = tmp;
        }
        /* magnitude^2 of each bin */
        for (int i = 0; i < Ndft; ++i)
            fftout[i].real = (fftout[i].real * fftout[i].real) + (fftout[i].imag * fftout[i].imag);
        /* single-pole IIR of the MAGNITUDE across blocks (and frames) */
        float tc = fsk->tc;
        for (int i = 0; i < Ndft; ++i) {
            fsk->Sf[i] = (fsk->Sf[i] * (1 - tc)) + (sqrtf(fftout[i].real) * tc);
            fftout[i].imag = fsk->Sf[i];
        }
    }
    /* [restatement note] with numffts == 0 upstream searches whatever the scratch
       buffer held; here it is defined as the current Sf */
    if (numffts <= 0)
        for (int i = 0; i < Ndft; ++i) fftout[i].imag = fsk->Sf[i];

    /* estimator 1: pick M peaks, blanking +/- f_zero bins around each */
    for (int i = 0; i < M; ++i) {
        int imax = 0;
        float max = 0;
        for (int j = st; j < en; ++j) {
            if (fftout[j].imag > max) {
                max = fftout[j].imag;
                imax = j;
            }
        }
        int f_min = imax - f_zero;
        f_min = f_min < 0 ? 0 : f_min;
        int f_max = imax + f_zero;
        f_max = f_max > Ndft ? Ndft : f_max;
        for (int j = f_min; j < f_max; ++j) fftout[j].imag = 0;
        freqi[i] = imax - Ndft / 2;
    }
    /* ascending sort of the M bin indices */
    {
        int i = 1;
        while (i < M) {
            if (freqi[i] >= freqi[i - 1]) {
                i++;
            } else {
                int j = freqi[i];
                freqi[i] = freqi[i - 1];
                freqi[i - 1] = j;
                if (i > 1) i--;
            }
        }
    }
    for (int i = 0; i < M; ++i) freqs[i] = (float)(freqi[i]) * ((float)Fs / (float)Ndft);

    /* estimator 2: slide a mask with 3-bin-wide teeth at the tone spacing over Sf */
    float *mask = (float *)malloc(sizeof(float) * (size_t)Ndft);
    for (int i = 0; i < Ndft; ++i) mask[i] = 0.0;
    for (int i = 0; i < 3; ++i) mask[i] = 1.0;
    int bin = 0;
    for (int m = 1; m <= M - 1; ++m) {
        bin = round((float)m * fsk->tone_spacing * Ndft / Fs) - 1;
        for (int i = bin; i <= bin + 2; ++i)
            if (i >= 0 && i < Ndft) mask[i] = 1.0;
    }
    int len_mask = bin + 2 + 1;
    int b_max = st;
    float corr_max = 0.0;
    const float *Sf = fsk->Sf;
    for (int b = st; b < en - len_mask; ++b) {
        float corr = 0.0;
        for (int i = 0; i < len_mask; ++i) corr += mask[i] * Sf[b + i];
        if (corr > corr_max) {
            corr_max = corr;
            b_max = b;
        }
    }
    float foff = (b_max - Ndft / 2) * Fs / Ndft; /* integer arithmetic, then to float */
    for (int m = 0; m < M; ++m) fsk->f2_est[m] = foff + m * fsk->tone_spacing;

    free(mask);
    free(fftin);
    free(fftout);
}

/* ======================================================================
 * demodulator core  [U fsk.c::fsk_demod_core]
 * SURVEY.md 8a rows A8-A11, App. A.3.
 * ====================================================================== */
static void demod_core(struct orc_fsk *fsk, uint8_t rx_bits[], float rx_filt[], const orc_comp fsk_in[]) {
    const int N = fsk->N;
    const int Ts = fsk->Ts;
    const int Rs = fsk->Rs;
    const int Fs = fsk->Fs;
    const int nsym = fsk->Nsym;
    const int nin = fsk->nin;
    const int P = fsk->P;
    const int Nmem = fsk->Nmem;
    const int M = fsk->mode;
    orc_comp *phi_c = fsk->phi_c;
    orc_comp *f_dc = fsk->f_dc;
    const int nold = Nmem - nin;

    /* A7: tone frequencies */
    demod_freq_est(fsk, fsk_in, fsk->f_est, M);
    const float *f_est = fsk->freq_est_type ? fsk->f2_est : fsk->f_est;

    /* A8: slide the per-tone baseband memory left by nin samples */
    for (int m = 0; m < M; ++m)
        for (int i = 0, j = Nmem - nold; i < nold; ++i, ++j) f_dc[m * Nmem + i] = f_dc[m * Nmem + j];

    /* A8: mix each tone to DC, continuing the oscillator phase of the last frame */
    for (int m = 0; m < M; ++m) {
        orc_comp dphi_m = comp_exp_j((float)(2 * M_PI * ((f_est[m]) / (float)(Fs))));
        for (int i = nold, j = 0; i < Nmem; ++i, ++j) {
            phi_c[m] = c_mul(phi_c[m], dphi_m);
            f_dc[m * Nmem + i] = c_mul(fsk_in[j], cconj(phi_c[m]));
        }
        phi_c[m] = comp_normalize(phi_c[m]);
    }

    /* A9: integrate over one symbol period at every one of the (Nsym+1)*P offsets */
    const int nint = (nsym + 1) * P;
    orc_comp *f_int = (orc_comp *)malloc(sizeof(orc_comp) * (size_t)(M * nint));
    for (int i = 0; i < nint; ++i) {
        int st = i * Ts / P;
        int en = st + Ts;
        for (int m = 0; m < M; ++m) {
            orc_comp acc = comp0();
            for (int j = st; j < en; ++j) acc = c_add(acc, f_dc[m * Nmem + j]);
            f_int[m * nint + i] = acc;
        }
    }

    /* A10: fine timing — the spectral line at Rs of the summed squared envelopes */
    orc_comp dphift = comp_exp_j((float)(2 * M_PI * ((float)(Rs) / (float)(P * Rs))));
    orc_comp phi_ft = { 1.0f, 0.0f };
    orc_comp t_c = comp0();
    for (int i = 0; i < nint; ++i) {
        float ft1 = 0;
        for (int m = 0; m < M; ++m) {
            orc_comp v = f_int[m * nint + i];
            ft1 += (v.real * v.real) + (v.imag * v.imag);
        }
        t_c = c_add(t_c, fcmult(ft1, phi_ft));
        phi_ft = c_mul(phi_ft, dphift);
    }
    if (isnan(t_c.real) || isnan(t_c.imag)) {
        free(f_int);
        return;
    }
    float norm_rx_timing = om_atan2f(t_c.imag, t_c.real) / (2 * M_PI);
    float rx_timing = norm_rx_timing * (float)P;
    float old_norm_rx_timing = fsk->norm_rx_timing;
    fsk->norm_rx_timing = norm_rx_timing;

    /* sample clock offset; ignore the jumps a change of nin causes */
    float d_norm_rx_timing = norm_rx_timing - old_norm_rx_timing;
    if (fabsf(d_norm_rx_timing) < .2) {
        float appm = 1e6 * d_norm_rx_timing / (float)nsym;
        fsk->ppm = .9 * fsk->ppm + .1 * appm;
    }

    /* how many samples the next call wants */
    if (!fsk->burst_mode && !fsk->lock_nin) {
        if (norm_rx_timing > 0.25)
            fsk->nin = N + Ts / 4;
        else if (norm_rx_timing < -0.25)
            fsk->nin = N - Ts / 4;
        else
            fsk->nin = N;
    }

    /* A11: resample the integrator outputs at the timing instant (linear interp) */
    int low_sample = (int)floorf(rx_timing);
    float fract = rx_timing - (float)low_sample;
    int high_sample = (int)ceilf(rx_timing);

    float meanebno = 0, stdebno = 0;
    float rx_nse_pow = 1E-12;
    float rx_sig_pow = 0.0;
    for (int i = 0; i < nsym; ++i) {
        int st = (i + 1) * P;
        float tmax[ORC_MODE_M_MAX];
        for (int m = 0; m < M; ++m) {
            orc_comp t = fcmult(1 - fract, f_int[m * nint + st + low_sample]);
            t = c_add(t, fcmult(fract, f_int[m * nint + st + high_sample]));
            tmax[m] = (t.real * t.real) + (t.imag * t.imag);
        }
        /* hard decision: largest tone energy, first one wins ties */
        float max = tmax[0];
        int sym = 0;
        for (int m = 0; m < M; ++m) {
            if (tmax[m] > max) {
                max = tmax[m];
                sym = m;
            }
        }
        if (rx_bits != NULL) {
            if (M == 2) {
                rx_bits[i] = sym == 1;
            } else {
                rx_bits[(i * 2) + 1] = (sym & 0x1);
                rx_bits[(i * 2)] = (sym & 0x2) >> 1;
            }
        }
        /* tone-major filter magnitudes for the LLR stage; always update S and N */
        float sum = 0.0;
        for (int m = 0; m < M; ++m) {
            if (rx_filt != NULL) rx_filt[m * nsym + i] = sqrtf(tmax[m]);
            sum += tmax[m];
        }
        rx_sig_pow += max;
        rx_nse_pow += (sum - max) / (M - 1);
        /* EbNodB estimator statistics */
        stdebno += max;
        meanebno += sqrtf(max);
    }
    fsk->rx_sig_pow = rx_sig_pow = rx_sig_pow / nsym;
    fsk->rx_nse_pow = rx_nse_pow = rx_nse_pow / nsym;
    fsk->v_est = sqrt(rx_sig_pow - rx_nse_pow);
    fsk->SNRest = rx_sig_pow / rx_nse_pow;

    meanebno = meanebno / (float)nsym;
    stdebno = (stdebno / (float)nsym) - (meanebno * meanebno);
    if (stdebno > 0.0)
        stdebno = sqrt(stdebno);
    else
        stdebno = 0.0;
    fsk->EbNodB = -6 + (20 * om_log10f((1e-6 + meanebno) / (1e-6 + stdebno)));

    fsk->stats_clock_offset = fsk->ppm;
    fsk->stats_snr_est = .5 * fsk->stats_snr_est + .5 * fsk->EbNodB;
    fsk->stats_rx_timing = (float)rx_timing;
    float fc_avg = (f_est[0] + f_est[1]) / 2;
    float fc_tx = (fsk->f1_tx + fsk->f1_tx + fsk->tone_spacing) / 2; /* integer expression */
    fsk->stats_foff = fc_tx - fc_avg;
    /* eye-diagram samples: display only, not restated */
    free(f_int);
}

void orc_fsk_demod(struct orc_fsk *fsk, uint8_t rx_bits[], const orc_comp fsk_in[]) { demod_core(fsk, rx_bits, NULL, fsk_in); }
void orc_fsk_demod_sd(struct orc_fsk *fsk, float rx_filt[], const orc_comp fsk_in[]) { demod_core(fsk, NULL, rx_filt, fsk_in); }

/* ======================================================================
 * modulator  [U fsk.c::fsk_mod_c] — synthetic input generation only
 * ====================================================================== */
void orc_fsk_mod_c(struct orc_fsk *fsk, orc_comp fsk_out[], const uint8_t tx_bits[], int nbits) {
    orc_comp tx_phase_c = fsk->tx_phase_c;
    const int Ts = fsk->Ts, Fs = fsk->Fs, M = fsk->mode;
    orc_comp dosc_f[ORC_MODE_M_MAX];
    for (int m = 0; m < M; ++m)
        dosc_f[m] = comp_exp_j((float)(2 * M_PI * ((float)(fsk->f1_tx + (fsk->tone_spacing * m)) / (float)(Fs))));
    int bit_i = 0;
    int nsym = nbits / (M >> 1);
    for (int i = 0; i < nsym; ++i) {
        int sym = 0;
        for (int m = M; m >>= 1;) {
            uint8_t bit = tx_bits[bit_i] == 1 ? 1 : 0;
            sym = (sym << 1) | bit;
            bit_i++;
        }
        orc_comp dph = dosc_f[sym];
        for (int j = 0; j < Ts; ++j) {
            tx_phase_c = c_mul(tx_phase_c, dph);
            fsk_out[i * Ts + j] = fcmult(2.0, tx_phase_c);
        }
    }
    fsk->tx_phase_c = comp_normalize(tx_phase_c);
}

/* ======================================================================
 * whole-stream demod  [U fsk_demod.c main loop], SURVEY.md App. A.4
 * ====================================================================== */
#define ORC_FDMDV_SCALE 825

size_t orc_fsk_demod_stream(const struct orc_demod_cfg *cfg, const void *iq, size_t nsamples,
                            uint8_t *bits_out, float *llr_out, float *rx_filt_out,
                            struct orc_frame_stats *stats_out, size_t max_frames) {
    struct orc_fsk *fsk = orc_fsk_create_hbr(cfg->Fs, cfg->Rs, cfg->M, cfg->P, cfg->Nsym, cfg->f1_tx, cfg->tone_spacing);
    if (cfg->have_limits) orc_fsk_set_freq_est_limits(fsk, cfg->est_min, cfg->est_max);
    orc_fsk_set_freq_est_alg(fsk, cfg->freq_est_type);

    const int Nbits = fsk->Nbits, M = fsk->mode, Nsym = fsk->Nsym;
    orc_comp *modbuf = (orc_comp *)malloc(sizeof(orc_comp) * (size_t)(fsk->N + fsk->Ts * 2));
    uint8_t *bitbuf = (uint8_t *)malloc((size_t)Nbits);
    float *rx_filt = (float *)malloc(sizeof(float) * (size_t)(M * Nsym));
    float *llrs = (float *)malloc(sizeof(float) * (size_t)Nbits);
    size_t pos = 0, nframes = 0;

    while (nframes < max_frames) {
        size_t nin = orc_fsk_nin(fsk);
        if (pos + nin > nsamples) break; /* short read ends the loop */
        if (cfg->fmt == ORC_FMT_U8) {
            const uint8_t *raw = (const uint8_t *)iq + 2 * pos;
            for (size_t i = 0; i < nin; ++i) {
                modbuf[i].real = ((float)raw[2 * i] - 127.0) / 128.0;
                modbuf[i].imag = ((float)raw[2 * i + 1] - 127.0) / 128.0;
            }
        } else if (cfg->fmt == ORC_FMT_S16) {
            const int16_t *raw = (const int16_t *)iq + 2 * pos;
            for (size_t i = 0; i < nin; ++i) {
                modbuf[i].real = ((float)raw[2 * i]) / ORC_FDMDV_SCALE;
                modbuf[i].imag = ((float)raw[2 * i + 1] / ORC_FDMDV_SCALE);
            }
        } else {
            const orc_comp *raw = (const orc_comp *)iq + pos;
            for (size_t i = 0; i < nin; ++i) modbuf[i] = raw[i];
        }
        demod_core(fsk, bitbuf, rx_filt, modbuf);
        if (cfg->soft) orc_fsk_rx_filt_to_llrs(llrs, rx_filt, fsk->v_est, fsk->SNRest, M, Nsym);

        if (bits_out) memcpy(bits_out + nframes * (size_t)Nbits, bitbuf, (size_t)Nbits);
        if (llr_out && cfg->soft) memcpy(llr_out + nframes * (size_t)Nbits, llrs, sizeof(float) * (size_t)Nbits);
        if (rx_filt_out) memcpy(rx_filt_out + nframes * (size_t)(M * Nsym), rx_filt, sizeof(float) * (size_t)(M * Nsym));
        if (stats_out) {
            struct orc_frame_stats *s = &stats_out[nframes];
            memset(s, 0, sizeof *s);
            for (int m = 0; m < M; ++m) {
                s->f_est[m] = fsk->f_est[m];
                s->f2_est[m] = fsk->f2_est[m];
            }
            s->norm_rx_timing = fsk->norm_rx_timing;
            s->ppm = fsk->ppm;
            s->SNRest = fsk->SNRest;
            s->v_est = fsk->v_est;
            s->rx_sig_pow = fsk->rx_sig_pow;
            s->rx_nse_pow = fsk->rx_nse_pow;
            s->EbNodB = fsk->EbNodB;
            s->snr_est = fsk->stats_snr_est;
            s->foff = fsk->stats_foff;
            s->rx_timing = fsk->stats_rx_timing;
            s->nin_next = fsk->nin;
            s->sample_offset = (int32_t)pos;
        }
        pos += nin;
        nframes++;
    }
    free(modbuf);
    free(bitbuf);
    free(rx_filt);
    free(llrs);
    orc_fsk_destroy(fsk);
    return nframes;
}
