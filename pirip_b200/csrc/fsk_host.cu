/*
 * fsk_host.cu — host side of the FSK demodulator: configuration validation, table
 * generation, the batched C-ABI and codec2's `struct FSK` API on top of the
 * persistent-CTA kernel (fsk_kernel.cu).  See include/pirip_fsk.h for the contract.
 *
 * The derived sizes follow codec2 fsk.c::fsk_create_core (SURVEY.md 8a row A6,
 * App. A.1); invalid configurations abort through assert() in the codec2-named
 * entry points, exactly as upstream does, and return NULL / a negative code in the
 * batched extension.
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "exact_math.h"
#include "fsk_dev.h"
#include "host_util.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------------------------------------------------------------- tables */

namespace {

struct cpx { float re, im; };
static inline cpx cmulf(cpx a, cpx b) { return cpx{ a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; }
static inline cpx cexpj(float phi) { return cpx{ em_cosf(phi), em_sinf(phi) }; }

/* Hann window through the same recursive oscillator codec2 uses for its table
   ([U fsk.c::fsk_generate_hann_table]) */
static std::vector<float> make_hann(int Ndft) {
    std::vector<float> h((size_t)Ndft);
    const cpx d = cexpj((float)((2 * M_PI) / ((float)Ndft - 1)));
    cpx r = { .5f, 0.0f };
    r = cmulf(cpx{ d.re, -d.im }, r);
    for (int i = 0; i < Ndft; ++i) {
        r = cmulf(d, r);
        h[(size_t)i] = (float)(.5 - r.re);
    }
    return h;
}

/* forward FFT twiddles, evaluated in double and rounded once ([U kiss_fft.c::kiss_fft_alloc]) */
static std::vector<float2> make_twiddles(int n) {
    std::vector<float2> t((size_t)n);
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < n; ++i) {
        const double phase = -2 * pi * i / n;
        t[(size_t)i] = make_float2((float)cos(phase), (float)sin(phase));
    }
    return t;
}

/* the timing-line oscillator e^{j 2 pi i / P} as the float recursion of
   [U fsk.c::fsk_demod_core] produces it; it does not depend on the signal */
static std::vector<float2> make_phi_ft(int Rs, int P, int nint) {
    std::vector<float2> t((size_t)nint);
    const cpx d = cexpj((float)(2 * M_PI * ((float)(Rs) / (float)(P * Rs))));
    cpx ph = { 1.0f, 0.0f };
    for (int i = 0; i < nint; ++i) {
        t[(size_t)i] = make_float2(ph.re, ph.im);
        ph = cmulf(ph, d);
    }
    return t;
}

/* slot of FFT input n before the first butterfly level: kiss_fft's decimation-in-time digit
   reversal for the factor list, then the bank swizzle of fsk_kernel.cu::fswz */
static std::vector<unsigned short> make_perm(const FskDevCfg &d) {
    std::vector<unsigned short> p((size_t)d.Ndft);
    for (int n = 0; n < d.Ndft; ++n) {
        int t = n, pos = 0;
        for (int s = 0; s < d.nfac; ++s) {
            const int digit = t % d.fac_p[s];
            t /= d.fac_p[s];
            pos += digit * d.fac_m[s];
        }
        p[(size_t)n] = (unsigned short)(pos ^ (5 * ((pos >> 4) & 3)));
    }
    return p;
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

/* ---------------------------------------------------------------- plan */

struct pirip_fsk_plan {
    pirip_fsk_cfg user;
    FskDevCfg dev;
    int device;
    float *d_hann = nullptr;
    float2 *d_tw = nullptr;
    float2 *d_phi_ft = nullptr;
    unsigned short *d_perm = nullptr;
    std::vector<float> h_hann;
    std::string kernel_desc;
    FskFastPlan *fast = nullptr; /* tolerance-mode batch kernel (fsk_fast.cu); NULL when the configuration is outside its range */
    int mode = 1;                /* 1: fast kernel + exact re-run of flagged captures; 0: exact kernel only */
    size_t last_n = 0;           /* captures of the last fast launch sequence (for pirip_fsk_plan_flagged) */
    size_t last_chunks = 0;
    /* reusable device buffers of the host-buffer entry point */
    struct Work {
        static const int kStreams = 3;
        cudaStream_t streams[kStreams] = { nullptr, nullptr, nullptr };
        unsigned char *d_iq = nullptr;
        uint8_t *d_bits = nullptr;
        float *d_llr = nullptr, *d_filt = nullptr;
        pirip_frame_stats *d_stats = nullptr;
        int32_t *d_nf = nullptr;
        unsigned long long *d_landed = nullptr; /* streamed ingest: samples per capture already on the device */
        unsigned long long *h_marks = nullptr;  /* pinned: the values d_landed takes, one per time slice */
        cudaEvent_t ev_reset = nullptr;
        size_t cap_iq = 0, cap_bits = 0, cap_llr = 0, cap_filt = 0, cap_stats = 0, cap_nf = 0;
    } work;
};

static bool cfg_valid(const pirip_fsk_cfg &c, std::string &why) {
    char buf[160];
#define REQ(cond)                                                   \
    if (!(cond)) {                                                  \
        snprintf(buf, sizeof buf, "invalid FSK config: %s", #cond); \
        why = buf;                                                  \
        return false;                                               \
    }
    REQ(c.Fs > 0);
    REQ(c.Rs > 0);
    REQ(c.P > 0);
    REQ(c.Nsym > 0);
    REQ((c.Fs % c.Rs) == 0);
    REQ(((c.Fs / c.Rs) % c.P) == 0);
    REQ(c.P >= 4);
    REQ(c.M == 2 || c.M == 4);
    REQ(c.fmt == PIRIP_FMT_U8 || c.fmt == PIRIP_FMT_S16 || c.fmt == PIRIP_FMT_F32);
    if (c.have_limits) {
        REQ(c.est_min >= -c.Fs / 2);
        REQ(c.est_max <= c.Fs / 2);
        REQ(c.est_max > c.est_min);
    }
#undef REQ
    return true;
}

/* Geometry of the batch kernel (fsk_kernel2.cu): one persistent CTA of 768 threads per SM, as many
   captures in lockstep as fit 227 KB.  Supported when the FFT block count does not depend on nin
   (the next frame's frequency estimate then runs one frame ahead); otherwise, or with
   PIRIP_FSK_K2=0, the batch entry points use fsk_kernel.cu. */
static void derive_k2(const pirip_fsk_cfg &c, FskDevCfg &d) {
    FskDevCfg::K2 &k = d.k2;
    memset(&k, 0, sizeof k);
    if (const char *e = getenv("PIRIP_FSK_K2"))
        if (atoi(e) == 0) return;
    const int half = d.Ndft / 2, nin_max = d.N + d.Ts / 4, nin_min = d.N - d.Ts / 4;
    const bool fixed_nin = d.burst_mode || d.lock_nin;
    const int n0 = nin_min / half - 1, n1 = d.N / half - 1, n2 = nin_max / half - 1;
    if (n1 < 1 || (!fixed_nin && !(n0 == n1 && n1 == n2))) return;
    if ((n1 + 1) * half > nin_min) return; /* the blocks must lie inside the shortest frame */
    k.numffts = n1;
    k.threads = 512; /* 16 warps (14 workers, timing, oscillator), 128 registers per thread */
    if (const char *e = getenv("PIRIP_FSK_K2_NW")) {
        const int v = atoi(e);
        if (v == 16 || v == 20 || v == 24) k.threads = 32 * v;
    }
    k.chunk = 128;
    while ((nin_max + k.chunk - 1) / k.chunk > FSK2_MAX_CHUNKS) k.chunk *= 2;
    if (k.chunk < d.Ts + 16) return;       /* the 4-chunk ring must cover a task's history */
    k.ring = 4 * k.chunk;
    k.ering = 4 * k.chunk;
    k.ckpt_every = FSK2_CKPT_EVERY;
    k.nckpt = nin_max / k.ckpt_every + 2;
    k.tail = align_up(2 * d.Ts + d.Ts / 4 + 2, 8);
    const int bps = c.fmt == PIRIP_FMT_U8 ? 2 : (c.fmt == PIRIP_FMT_S16 ? 4 : 8);
    k.stage_bytes = align_up(nin_max * bps + 32, 16);
    const int step = d.Ts / c.P;
    k.fast_ts = (step == 1 && (d.Ts == 24 || d.Ts == 20 || d.Ts == 40)) ? 1 : 0;
    if ((3 + c.M) * c.Nsym > c.M * k.ering) return; /* decision scratch lives in the energy ring */
    int off = 0;
    k.off_stage = off;
    off += 2 * k.stage_bytes;
    k.off_ring = off;
    off += align_up(c.M * (k.ring + 2) * 8, 16);
    k.off_tail = off;
    off += align_up(c.M * k.tail * 8, 16);
    k.off_ckpt = off;
    off += align_up(c.M * k.nckpt * 8, 16);
    k.off_ering = off;
    off += align_up(c.M * k.ering * 4, 16);
    k.off_Sf = off;
    off += align_up(2 * d.Ndft * 4, 16);
    k.off_scal = off;
    off += align_up((int)sizeof(Fsk2CapShared), 16);
    while ((off & 127) != 32) off += 16; /* capture blocks 32 (mod 128) bytes apart: spreads the per-capture lanes over banks */
    k.cap_stride = off;
    k.num_sms = 148; /* plan_build replaces it with the device's SM count */
    const int limit = 227 * 1024;
    int gcap = FSK2_MAX_G;
    if (gcap * c.M > 32) gcap = 32 / c.M; /* oscillator lanes (g, m) of one warp */
    if (const char *e = getenv("PIRIP_FSK_K2_G")) {
        const int v = atoi(e);
        if (v >= 1 && v < gcap) gcap = v;
    }
    for (int g = gcap; g >= 1; --g) {
        int o = 0;
        k.off_fft = o;
        o += align_up((k.threads / 64) * FSK2_FFT_SLOT_ELEMS(d.Ndft) * 8, 16); /* a dedicated FFT scratch slot for half the warps */
        k.off_tw = o;
        o += align_up(d.Ndft * 8, 16);
        k.off_hann = o;
        o += align_up(d.Ndft * 4, 16);
        k.off_perm = o;
        o += align_up(d.Ndft * 2, 16);
        k.off_phift = o;
        o += align_up(d.nint * 8, 16);
        k.off_ctl = o;
        o += align_up((int)sizeof(Fsk2CtaShared) + g * 2 * (int)sizeof(void *), 16);
        if (g * k.cap_stride + o <= limit) {
            k.gmax = g;
            k.smem_bytes = g * k.cap_stride + o;
            k.ok = 1;
            /* the other half of the warps borrow FFT scratch from the energy rings of captures 0 .. NW/2-1, above the
               decision scratch, when there are enough capture blocks and room */
            const int nw = k.threads / 32;
            k.ering_fft_off = align_up((3 + c.M) * c.Nsym, 64);
            const bool room = k.ering_fft_off + 2 * FSK2_FFT_SLOT_ELEMS(d.Ndft) <= c.M * k.ering && g >= nw / 2;
            k.fft_warps = room ? nw : nw / 2;
            /* opt-in: measured SLOWER on B200 (1250 x 1 s captures: 0 -> 10.82 ms, 5 -> 11.17 ms, 7 -> 12.32 ms, 9 -> 11.93 ms):
               FFT code running beside the streaming code slows the worker chain more than the shorter FFT phase saves */
            k.fft_early = 0;
            /* the queue needs every warp to be able to run an FFT task */
            k.task_queue = (k.fft_warps == nw && g * (k.numffts + 2) <= FSK2_MAX_Q) ? 1 : 0;
            if (const char *e = getenv("PIRIP_FSK_K2_QUEUE")) k.task_queue = k.task_queue ? atoi(e) : 0; /* 0 static, 1 queue, 2 mixed order */
            if (const char *e = getenv("PIRIP_FSK_K2_FFT_EARLY")) k.fft_early = atoi(e) < 0 ? 0 : atoi(e);
            return;
        }
    }
}

/* fill every derived field of FskDevCfg except the device pointers */
static bool derive_cfg(const pirip_fsk_cfg &c, int burst_mode, int lock_nin, int max_g, FskDevCfg &d, std::string &why) {
    memset(&d, 0, sizeof d);
    d.Fs = c.Fs;
    d.Rs = c.Rs;
    d.M = c.M;
    d.P = c.P;
    d.Nsym = c.Nsym;
    d.Ts = c.Fs / c.Rs;
    d.N = d.Ts * c.Nsym;
    d.Nmem = d.N + 2 * d.Ts;
    {
        /* bins within 10% of the symbol rate, next power of two */
        float bin_width_Hz = 0.1 * c.Rs;
        float Ndft = (float)c.Fs / bin_width_Hz;
        Ndft = pow(2.0, ceil(log2(Ndft)));
        d.Ndft = (int)Ndft;
    }
    d.Nbits = c.M == 2 ? c.Nsym : 2 * c.Nsym;
    d.nint = (c.Nsym + 1) * c.P;
    d.tc = 0.1;
    d.fs_over_ndft = (float)c.Fs / (float)d.Ndft;
    d.log2_ndft = ilog2(d.Ndft);
    d.est_type = c.freq_est_type;
    d.tone_spacing = c.tone_spacing;
    d.f1_tx = c.f1_tx;
    d.burst_mode = burst_mode;
    d.lock_nin = lock_nin;
    d.fmt = c.fmt;
    d.soft = c.soft;

    if (d.Ndft < 32 || d.Ndft > 1024) {
        why = "Ndft outside the supported 32..1024 range";
        return false;
    }
    if ((long long)c.Fs * d.Ndft > 0x7fffffffLL) {
        why = "Fs*Ndft overflows int (as it would upstream)";
        return false;
    }
    /* estimator window: defaults are codec2's est_min = 0, est_max = Fs */
    const int est_min = c.have_limits ? c.est_min : 0;
    const int est_max = c.have_limits ? c.est_max : c.Fs;
    const int est_space = 0.75 * c.Rs;
    d.st = (est_min * d.Ndft) / c.Fs + d.Ndft / 2;
    if (d.st < 0) d.st = 0;
    d.en = (est_max * d.Ndft) / c.Fs + d.Ndft / 2;
    if (d.en > d.Ndft) d.en = d.Ndft;
    d.f_zero = (est_space * d.Ndft) / c.Fs;

    /* tone mask: 3-bin teeth at multiples of the tone spacing */
    std::vector<char> mask((size_t)d.Ndft, 0);
    for (int i = 0; i < 3; ++i) mask[(size_t)i] = 1;
    int bin = 0;
    for (int m = 1; m <= c.M - 1; ++m) {
        bin = round((float)m * c.tone_spacing * d.Ndft / c.Fs) - 1;
        for (int i = bin; i <= bin + 2; ++i)
            if (i >= 0 && i < d.Ndft) mask[(size_t)i] = 1;
    }
    d.len_mask = bin + 2 + 1;
    d.nmask = 0;
    for (int i = 0; i < d.len_mask && i < d.Ndft; ++i)
        if (mask[(size_t)i]) {
            if (d.nmask >= FSK_MAX_MASK) {
                why = "tone mask has too many teeth";
                return false;
            }
            d.mask_pos[d.nmask++] = i;
        }

    /* FFT factors in kiss_fft order: 4s first, then a 2 */
    {
        int n = d.Ndft, fstride = 1;
        d.nfac = 0;
        while (n > 1) {
            const int p = (n % 4 == 0) ? 4 : 2;
            n /= p;
            d.fac_p[d.nfac] = p;
            d.fac_m[d.nfac] = n;
            d.fac_log2m[d.nfac] = ilog2(n);
            d.fac_fstride[d.nfac] = fstride;
            fstride *= p;
            d.nfac++;
        }
    }

    /* launch geometry and shared memory layout (see fsk_kernel.cu "PARALLEL STRUCTURE") */
    const int nin_max = d.N + d.Ts / 4;
    d.max_numffts = nin_max / (d.Ndft / 2) - 1;
    if (d.max_numffts < 1) d.max_numffts = 1;
    /* tones of one capture are 2 (mod 16) elements apart: even (16-byte vector loads) and on
       different banks for the per-tone oscillator lanes; at least 2 spare elements */
    d.fdc_tone_stride = d.Nmem + 2;
    while ((d.fdc_tone_stride & 15) != 2) d.fdc_tone_stride++;
    {
        /* frequency estimate one frame ahead (front team inside the CTA): possible when the FFT block
           count is the same for the three values nin can take.  Measured on B200 it does not pay
           while two CTAs per SM already overlap each other's phases (13.5 vs 15.0 Gsamples/s), so it
           is opt-in (PIRIP_FSK_PIPE=1) and covered by the GPU tests. */
        const int half = d.Ndft / 2;
        const int n0 = (d.N - d.Ts / 4) / half - 1, n1 = d.N / half - 1, n2 = nin_max / half - 1;
        const bool fixed_nin = burst_mode || lock_nin;
        d.numffts_const = n1;
        d.pipeline = ((fixed_nin || (n0 == n1 && n1 == n2)) && n1 >= 1 && getenv("PIRIP_FSK_PIPE") != nullptr) ? 1 : 0;
    }
    d.chunk = 256; /* measured best on B200 (64: 24.3 ms, 128: 19.6, 256: 19.2, 512: 20.1 for the bench batch) */
    if (const char *e = getenv("PIRIP_FSK_CHUNK")) {
        const int v = atoi(e);
        if (v >= 16 && v <= 1024 && (v & (v - 1)) == 0) d.chunk = v;
    }
    while ((nin_max + d.chunk - 1) / d.chunk > FSK_MAX_CHUNKS) d.chunk *= 2;
    /* Per-capture block for a given magnitude-buffer capacity; returns its size. */
    auto layout_capture = [&](int mag_blocks) {
        int off = 0;
        d.off_fdc = off;
        off += align_up(c.M * d.fdc_tone_stride * 8, 16);
        d.off_xin = off;
        {
            const int in_elems = nin_max > d.nint ? nin_max : d.nint; /* frame input, then timing products */
            off += align_up((in_elems + 2) * 8, 16);
        }
        d.off_mag = off;
        {
            const int fft_floats = mag_blocks * d.Ndft;
            const int dec_floats = d.Ndft + (3 + c.M) * c.Nsym; /* tmpS | sig | nse | sqm | rx_filt */
            off += align_up((fft_floats > dec_floats ? fft_floats : dec_floats) * 4, 16);
        }
        d.off_Sf = off;
        off += align_up(d.Ndft * 4, 16);
        d.off_stage = off;
        d.stage_bytes = align_up(nin_max * (c.fmt == PIRIP_FMT_U8 ? 2 : (c.fmt == PIRIP_FMT_S16 ? 4 : 8)) + 32, 16);
        off += d.stage_bytes;
        d.off_phead = off;
        off += align_up((d.Ts + d.Ts / 4 + 1) * 8, 16);
        d.off_scal = off;
        off += align_up((int)sizeof(FskCapShared), 16);
        /* capture blocks 32 (mod 128) bytes apart: with tones 16 (mod 128) bytes apart the lanes
           (capture, tone) of one quarter-warp hit distinct 16-byte bank groups (M = 2, G <= 4) */
        while ((off & 127) != (c.M == 2 ? 32 : 64)) off += 16;
        d.cap_stride = off;
        d.mag_blocks = mag_blocks;
        return off;
    };
    auto layout_cta = [&](int threads) {
        int o = 0;
        d.off_fft = o;
        o += align_up((threads / 32) * d.Ndft * 8, 16);
        d.off_tw = o;
        o += align_up(d.Ndft * 8, 16);
        d.off_hann = o;
        o += align_up(d.Ndft * 4, 16);
        d.off_phift = 0; /* the timing oscillator table stays in global memory */
        d.off_perm = o;
        o += align_up(d.Ndft * 2, 16);
        d.off_ctl = o;
        o += align_up((int)sizeof(FskCtaShared), 16);
        return o;
    };

    /* Geometry.  Preferred: two 256-thread CTAs of G = 2 captures per SM (their phases overlap);
       otherwise one 512-thread CTA with as many captures as fit 227 KB.  PIRIP_FSK_G forces G. */
    int force_g = 0;
    if (const char *e = getenv("PIRIP_FSK_G")) force_g = atoi(e);
    if (force_g > FSK_MAX_G) force_g = FSK_MAX_G;
    if (force_g == 7) force_g = 6;
    d.G = 0;
    const int sm_bytes = 228 * 1024, cta_reserve = 1024;
    if (max_g >= 2 && (force_g == 0 || force_g == 2)) {
        const int budget = sm_bytes / 2 - cta_reserve;
        const int cta_bytes = layout_cta(256);
        for (int mb = d.max_numffts; mb >= 1; --mb) {
            if (2 * layout_capture(mb) + cta_bytes <= budget) {
                d.G = 2;
                d.threads = 256;
                d.smem_bytes = 2 * d.cap_stride + cta_bytes;
                break;
            }
        }
    }
    if (d.G == 0) {
        const int limit = 227 * 1024;
        const int gmax = force_g ? force_g : (max_g < 5 ? max_g : 5);
        for (int g = gmax; g >= 1; --g) {
            if (g == 7) continue;
            const int threads = g <= 2 ? 256 : 512;
            const int cta_bytes = layout_cta(threads);
            if (g * layout_capture(d.max_numffts) + cta_bytes <= limit) {
                d.G = g;
                d.threads = threads;
                d.smem_bytes = g * d.cap_stride + cta_bytes;
                break;
            }
        }
    }
    if (d.G == 0) {
        why = "configuration needs more than 227 KB of shared memory per capture";
        return false;
    }
    derive_k2(c, d);
    return true;
}

static void plan_free(pirip_fsk_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    fsk_fast_plan_destroy(p->fast);
    cudaFree(p->d_hann);
    cudaFree(p->d_tw);
    cudaFree(p->d_phi_ft);
    cudaFree(p->d_perm);
    cudaFree(p->work.d_iq);
    cudaFree(p->work.d_bits);
    cudaFree(p->work.d_llr);
    cudaFree(p->work.d_filt);
    cudaFree(p->work.d_stats);
    cudaFree(p->work.d_nf);
    cudaFree(p->work.d_landed);
    if (p->work.h_marks) cudaFreeHost(p->work.h_marks);
    if (p->work.ev_reset) cudaEventDestroy(p->work.ev_reset);
    for (int i = 0; i < pirip_fsk_plan::Work::kStreams; ++i)
        if (p->work.streams[i]) cudaStreamDestroy(p->work.streams[i]);
    delete p;
}

static pirip_fsk_plan *plan_build(const pirip_fsk_cfg &cfg, int device, int burst_mode, int lock_nin, int max_g = FSK_MAX_G,
                                  bool want_fast = false) {
    std::string why;
    if (!cfg_valid(cfg, why)) {
        pirip::set_error(why);
        return nullptr;
    }
    pirip_fsk_plan *p = new pirip_fsk_plan();
    p->user = cfg;
    p->device = device;
    if (!derive_cfg(cfg, burst_mode, lock_nin, max_g, p->dev, why)) {
        pirip::set_error(why);
        delete p;
        return nullptr;
    }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) p->dev.k2.num_sms = sms;
    }
    p->h_hann = make_hann(p->dev.Ndft);
    const std::vector<float2> tw = make_twiddles(p->dev.Ndft);
    const std::vector<float2> pf = make_phi_ft(cfg.Rs, cfg.P, p->dev.nint);
    const std::vector<unsigned short> pm = make_perm(p->dev);
    if (!pirip::cuda_ok(cudaSetDevice(device), "cudaSetDevice") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_hann, sizeof(float) * p->h_hann.size()), "cudaMalloc hann") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_tw, sizeof(float2) * tw.size()), "cudaMalloc twiddles") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_phi_ft, sizeof(float2) * pf.size()), "cudaMalloc phi_ft") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_hann, p->h_hann.data(), sizeof(float) * p->h_hann.size(), cudaMemcpyHostToDevice), "copy hann") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice), "copy twiddles") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_phi_ft, pf.data(), sizeof(float2) * pf.size(), cudaMemcpyHostToDevice), "copy phi_ft") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_perm, sizeof(unsigned short) * pm.size()), "cudaMalloc perm") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_perm, pm.data(), sizeof(unsigned short) * pm.size(), cudaMemcpyHostToDevice), "copy perm") ||
        !pirip::cuda_ok(fsk_kernel_prepare(p->dev), "cudaFuncSetAttribute") ||
        (p->dev.k2.ok && !pirip::cuda_ok(fsk2_kernel_prepare(p->dev), "cudaFuncSetAttribute (batch kernel)"))) {
        plan_free(p);
        return nullptr;
    }
    p->dev.hann = p->d_hann;
    p->dev.twiddle = p->d_tw;
    p->dev.phi_ft = p->d_phi_ft;
    p->dev.perm = p->d_perm;
    /* the fast kernel hands flagged captures to the batch kernel's list mode, so it needs that kernel too */
    if (want_fast && p->dev.k2.ok) {
        std::string why_not;
        p->fast = fsk_fast_plan_create(cfg, p->dev, p->h_hann.data(), pf.data(), device, why_not);
        if (!p->fast) pirip::set_error("");
    }
    if (const char *e = getenv("PIRIP_FSK_MODE")) p->mode = (strcmp(e, "exact") == 0 || strcmp(e, "0") == 0) ? 0 : 1;
    return p;
}

extern "C" pirip_fsk_plan *pirip_fsk_plan_create(const pirip_fsk_cfg *cfg, int device) {
    if (!cfg) {
        pirip::set_error("cfg is NULL");
        return nullptr;
    }
    return plan_build(*cfg, device, 0, 0, FSK_MAX_G, true);
}

extern "C" int pirip_fsk_plan_set_mode(pirip_fsk_plan *plan, int mode) {
    if (!plan) return pirip::fail(-1, "plan is NULL");
    if (mode != PIRIP_FSK_MODE_EXACT && mode != PIRIP_FSK_MODE_FAST) return pirip::fail(-2, "unknown mode");
    if (mode == PIRIP_FSK_MODE_FAST && !plan->fast) return pirip::fail(-3, "configuration outside the fast kernel's range");
    plan->mode = mode;
    return 0;
}

/* flags of the last batch call (fast mode): synchronises the device, copies the per-capture flag words and returns how
   many captures were re-run by the exact kernel; 0 in exact mode */
extern "C" long pirip_fsk_plan_flagged(pirip_fsk_plan *plan, int32_t *flags_out, size_t n) {
    if (!plan) return pirip::fail(-1, "plan is NULL");
    if (!(plan->fast && plan->mode) || plan->last_n == 0) return 0;
    if (!pirip::cuda_ok(cudaSetDevice(plan->device), "cudaSetDevice")) return -10;
    if (!pirip::cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return -13;
    int *d_flags = nullptr, *d_list = nullptr, *d_count = nullptr;
    if (!pirip::cuda_ok(fsk_fast_flag_buffers(plan->fast, plan->last_n, &d_flags, &d_list, &d_count), "flag buffers")) return -12;
    int counts[64];
    const size_t nch = plan->last_chunks < 64 ? plan->last_chunks : 64;
    if (!pirip::cuda_ok(cudaMemcpy(counts, d_count, sizeof(int) * nch, cudaMemcpyDeviceToHost), "D2H flag counts")) return -12;
    long total = 0;
    for (size_t i = 0; i < nch; ++i) total += counts[i];
    if (flags_out) {
        const size_t m = n < plan->last_n ? n : plan->last_n;
        if (!pirip::cuda_ok(cudaMemcpy(flags_out, d_flags, sizeof(int32_t) * m, cudaMemcpyDeviceToHost), "D2H flags")) return -12;
    }
    return total;
}

/* exact in-kernel frame replays of this plan so far (synchronises): out[0] replayed, out[1] replays that fell back to flagging */
extern "C" int pirip_fsk_plan_replays(pirip_fsk_plan *plan, int32_t out[2]) {
    if (!plan || !out) return pirip::fail(-1, "plan / out is NULL");
    out[0] = out[1] = 0;
    if (!plan->fast) return 0;
    if (!pirip::cuda_ok(cudaSetDevice(plan->device), "cudaSetDevice")) return -10;
    if (!pirip::cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return -13;
    int v[64];
    if (!pirip::cuda_ok(fsk_fast_replay_counts(plan->fast, v), "D2H replay counts")) return -12;
    out[0] = v[0];
    out[1] = v[1];
    if (getenv("PIRIP_FSK_CHAIN_DIAG"))
        fprintf(stderr, "side chain: replays %d failed %d, ppm-threshold near-ties %d, kcycles waiting for checkpoints %d, kcycles in replays %d; mailbox lag: mean %.1f us, max %d us over %d frames\n", v[0], v[1],
                v[2], v[3], v[4], v[7] ? (double)v[5] / v[7] : 0.0, v[6], v[7]);
    if (getenv("PIRIP_FSK_CHAIN_DIAG"))
        fprintf(stderr, "side chain warps: %d iterations with work (%.0f cycles each), %d without; %d kcycles in all\n", v[8],
                v[8] ? 1024.0 * v[10] / v[8] : 0.0, v[9], v[11]);
    if (getenv("PIRIP_FSK_CHAIN_DIAG")) {
        fprintf(stderr, "replay sections, k cycles per replay: wait %.1f mix %.1f timing line %.1f atan2/lane 0 %.1f decisions %.1f symbol sums %.1f rest %.1f\n",
                v[0] ? v[24] / 16.0 / v[0] : 0.0, v[0] ? v[25] / 16.0 / v[0] : 0.0, v[0] ? v[26] / 16.0 / v[0] : 0.0, v[0] ? v[27] / 16.0 / v[0] : 0.0,
                v[0] ? v[28] / 16.0 / v[0] : 0.0, v[0] ? v[29] / 16.0 / v[0] : 0.0, v[0] ? v[30] / 16.0 / v[0] : 0.0);
        fprintf(stderr, "side chain cycles per iteration by warp slot:");
        for (int w = 0; w < 18; ++w)
            if (v[40 + w]) fprintf(stderr, " %d:%.0f", w, 1024.0 * v[16 + w] / v[40 + w]);
        fprintf(stderr, "\n");
    }
    return 0;
}

/* frames whose ppm update decision (|d norm_rx_timing| < 0.2) lay inside the estimates' own error and was taken from the fast
   values, since the plan was created (synchronises the device); the ppm statistic of those captures may differ from the exact
   evaluation for the following ~40 frames, nothing else does */
extern "C" long pirip_fsk_plan_ppm_near_ties(pirip_fsk_plan *plan) {
    if (!plan) return pirip::fail(-1, "plan is NULL");
    if (!plan->fast) return 0;
    if (!pirip::cuda_ok(cudaSetDevice(plan->device), "cudaSetDevice")) return -10;
    if (!pirip::cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return -13;
    int v[64];
    if (!pirip::cuda_ok(fsk_fast_replay_counts(plan->fast, v), "D2H replay counts")) return -12;
    return v[2];
}

extern "C" int pirip_fsk_plan_mode(const pirip_fsk_plan *plan) { return plan ? ((plan->fast && plan->mode) ? 1 : 0) : -1; }

extern "C" void pirip_fsk_plan_destroy(pirip_fsk_plan *plan) { plan_free(plan); }

extern "C" const char *pirip_fsk_plan_kernel(const pirip_fsk_plan *plan) {
    if (!plan) return "";
    pirip_fsk_plan *p = const_cast<pirip_fsk_plan *>(plan);
    char buf[256];
    if (plan->fast && plan->mode) {
        p->kernel_desc = fsk_fast_plan_desc(plan->fast);
        return p->kernel_desc.c_str();
    }
    if (plan->dev.k2.ok)
        snprintf(buf, sizeof buf, "fsk2_kernel<M=%d, %d warps>: one persistent CTA per SM (%d SMs), up to %d captures in lockstep, %d B dynamic smem",
                 plan->dev.M, plan->dev.k2.threads / 32, plan->dev.k2.num_sms, plan->dev.k2.gmax, plan->dev.k2.smem_bytes);
    else
        snprintf(buf, sizeof buf, "fsk_demod_kernel<G=%d>: %d threads per CTA, %d B dynamic smem", plan->dev.G, plan->dev.threads,
                 plan->dev.smem_bytes);
    p->kernel_desc = buf;
    return p->kernel_desc.c_str();
}

extern "C" int pirip_fsk_plan_nbits(const pirip_fsk_plan *plan) { return plan ? plan->dev.Nbits : -1; }
extern "C" int pirip_fsk_plan_nfilt(const pirip_fsk_plan *plan) { return plan ? plan->dev.M * plan->dev.Nsym : -1; }

extern "C" size_t pirip_fsk_plan_max_frames(const pirip_fsk_plan *plan, size_t samples_per_capture) {
    if (!plan) return 0;
    const size_t nin_min = (size_t)(plan->dev.N - plan->dev.Ts / 4);
    return samples_per_capture / nin_min + 1;
}

static size_t bytes_per_sample(int fmt) { return fmt == PIRIP_FMT_U8 ? 2 : (fmt == PIRIP_FMT_S16 ? 4 : 8); }

static int batch_dev(pirip_fsk_plan *plan, const void *d_iq, size_t n_captures, size_t samples_per_capture,
                     size_t capture_stride_bytes, uint8_t *d_bits, float *d_llr, float *d_rx_filt,
                     struct pirip_frame_stats *d_stats, int32_t *d_nframes, size_t max_frames, void *stream,
                     const unsigned long long *d_landed, size_t flag_off = 0, size_t flag_total = 0, size_t chunk = 0, int phases = 3);

extern "C" int pirip_fsk_demod_batch_dev(pirip_fsk_plan *plan, const void *d_iq, size_t n_captures,
                                         size_t samples_per_capture, size_t capture_stride_bytes, uint8_t *d_bits,
                                         float *d_llr, float *d_rx_filt, struct pirip_frame_stats *d_stats,
                                         int32_t *d_nframes, size_t max_frames, void *stream) {
    if (plan) {
        plan->last_n = n_captures;
        plan->last_chunks = 1;
    }
    return batch_dev(plan, d_iq, n_captures, samples_per_capture, capture_stride_bytes, d_bits, d_llr, d_rx_filt, d_stats,
                     d_nframes, max_frames, stream, nullptr);
}

static int batch_dev(pirip_fsk_plan *plan, const void *d_iq, size_t n_captures, size_t samples_per_capture,
                     size_t capture_stride_bytes, uint8_t *d_bits, float *d_llr, float *d_rx_filt,
                     struct pirip_frame_stats *d_stats, int32_t *d_nframes, size_t max_frames, void *stream,
                     const unsigned long long *d_landed, size_t flag_off, size_t flag_total, size_t chunk, int phases) {
    if (!plan) return pirip::fail(-1, "plan is NULL");
    if (n_captures == 0) return 0;
    if (!d_iq) return pirip::fail(-2, "d_iq is NULL");
    if (d_llr && !plan->dev.soft) return pirip::fail(-3, "LLR output requested but cfg.soft == 0");
    if (capture_stride_bytes < samples_per_capture * bytes_per_sample(plan->dev.fmt))
        return pirip::fail(-4, "capture_stride_bytes smaller than one capture");
    if (n_captures > 0x7fffffffu) return pirip::fail(-5, "too many captures for one launch");
    if (!pirip::cuda_ok(cudaSetDevice(plan->device), "cudaSetDevice")) return -10;
    FskBatchArgs a;
    memset(&a, 0, sizeof a);
    a.iq = (const unsigned char *)d_iq;
    a.capture_stride_bytes = capture_stride_bytes;
    a.samples_per_capture = samples_per_capture;
    a.max_frames = max_frames;
    a.bits = d_bits;
    a.llr = d_llr;
    a.rx_filt = d_rx_filt;
    a.stats = d_stats;
    a.nframes = d_nframes;
    a.n_captures = n_captures;
    a.tma_ok = (reinterpret_cast<size_t>(d_iq) & 15) == 0 && getenv("PIRIP_FSK_NO_TMA") == nullptr;
    a.landed = plan->dev.k2.ok ? d_landed : nullptr;
    if (plan->fast && plan->mode && flag_total == 0 && n_captures > fsk_fast_round_captures(plan->fast)) {
        /* more captures than one launch keeps resident: rounds of launches that share ONE list of flagged captures (batch-wide
           indices), then ONE re-run of that list by the exact kernel - a re-run per round cost ~1.4 ms of per-frame latency
           for every round that had a single flagged capture */
        const size_t R = fsk_fast_round_captures(plan->fast);
        const size_t nb = (size_t)plan->dev.Nbits, nfl = (size_t)plan->dev.M * plan->dev.Nsym;
        int *d_flags = nullptr, *d_list = nullptr, *d_count = nullptr;
        if (!pirip::cuda_ok(fsk_fast_flag_buffers(plan->fast, n_captures, &d_flags, &d_list, &d_count), "cudaMalloc flags")) return -12;
        plan->last_chunks = 1;
        if (phases & 1) {
            if (!pirip::cuda_ok(cudaMemsetAsync(d_count, 0, sizeof(int), (cudaStream_t)stream), "reset flag count")) return -11;
            for (size_t c0 = 0; c0 < n_captures; c0 += R) {
                FskBatchArgs r = a;
                r.n_captures = c0 + R <= n_captures ? R : n_captures - c0;
                r.iq = a.iq + c0 * capture_stride_bytes;
                r.bits = d_bits ? d_bits + c0 * max_frames * nb : nullptr;
                r.llr = d_llr ? d_llr + c0 * max_frames * nb : nullptr;
                r.rx_filt = d_rx_filt ? d_rx_filt + c0 * max_frames * nfl : nullptr;
                r.stats = d_stats ? d_stats + c0 * max_frames : nullptr;
                r.nframes = d_nframes + c0;
                r.landed = d_landed;
                if (!pirip::cuda_ok(fsk_fast_launch(plan->fast, r, d_flags + c0, d_list, d_count, (cudaStream_t)stream, c0, n_captures, c0),
                                    "fsk_fast_kernel launch"))
                    return -11;
                pirip::count_launch();
            }
        }
        if (!(phases & 2) || getenv("PIRIP_FSK_FAST_NO_RERUN") != nullptr) return 0;
        FskBatchArgs x = a;
        x.landed = nullptr;
        x.cap_list = d_list;
        x.cap_count = d_count;
        x.total_captures = n_captures;
        if (!pirip::cuda_ok(fsk2_launch_demod(plan->dev, x, n_captures, (cudaStream_t)stream), "fsk2_kernel (flagged captures) launch")) return -11;
        pirip::count_launch();
        return 0;
    }
    if (plan->fast && plan->mode) {
        /* fast kernel over every capture, then the exact kernel over the captures it flagged (list on the device) */
        if (chunk >= 64) return pirip::fail(-6, "too many chunks");
        int *d_flags = nullptr, *d_list = nullptr, *d_count = nullptr;
        if (!pirip::cuda_ok(fsk_fast_flag_buffers(plan->fast, flag_total ? flag_total : n_captures, &d_flags, &d_list, &d_count), "cudaMalloc flags"))
            return -12;
        d_flags += flag_off;
        d_list += flag_off;
        d_count += chunk;
        a.landed = d_landed;
        /* phases: 1 = the fast kernel, 2 = the exact kernel over the captures it flagged (the host-buffer call puts the end of
           its streamed copies between the two) */
        if (phases & 1) {
            if (!pirip::cuda_ok(cudaMemsetAsync(d_count, 0, sizeof(int), (cudaStream_t)stream), "reset flag count")) return -11;
            if (!pirip::cuda_ok(fsk_fast_launch(plan->fast, a, d_flags, d_list, d_count, (cudaStream_t)stream, flag_off, flag_total, chunk),
                                "fsk_fast_kernel launch"))
                return -11;
            pirip::count_launch();
        }
        if (!(phases & 2)) return 0;
        FskBatchArgs x = a;
        x.landed = nullptr; /* the host makes the re-run wait for the last copy */
        x.cap_list = d_list;
        x.cap_count = d_count;
        x.total_captures = n_captures;
        /* launch bound of the re-run: one resident group per SM is more than the guards ever flag; the rest of a longer
           list (never seen) would still be processed, group after group */
        if (getenv("PIRIP_FSK_FAST_NO_RERUN") != nullptr) return 0; /* diagnostics: the fast kernel's own results for every capture */
        if (!pirip::cuda_ok(fsk2_launch_demod(plan->dev, x, n_captures, (cudaStream_t)stream), "fsk2_kernel (flagged captures) launch")) return -11;
        pirip::count_launch();
        return 0;
    }
    if (plan->dev.k2.ok) {
        if (!pirip::cuda_ok(fsk2_launch_demod(plan->dev, a, n_captures, (cudaStream_t)stream), "fsk2_kernel launch")) return -11;
    } else if (!pirip::cuda_ok(fsk_launch_demod(plan->dev, a, n_captures, (cudaStream_t)stream), "fsk_demod_kernel launch"))
        return -11;
    pirip::count_launch();
    return 0;
}

/* Host-buffer variant.  Device buffers, streams and events are cached in the plan and
   reused across calls; the captures are split into chunks that are copied in, demodulated
   and copied out on alternating streams so PCIe transfers overlap the kernel. */
namespace {
template <typename T>
bool ensure_cap(T *&ptr, size_t &cap_bytes, size_t need_bytes, const char *what) {
    if (need_bytes <= cap_bytes && ptr) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap_bytes = 0;
    if (!pirip::cuda_ok(cudaMalloc(&ptr, need_bytes), what)) return false;
    cap_bytes = need_bytes;
    return true;
}
}  // namespace

extern "C" int pirip_fsk_demod_batch(pirip_fsk_plan *plan, const void *iq, size_t n_captures,
                                     size_t samples_per_capture, uint8_t *bits, float *llr, float *rx_filt,
                                     struct pirip_frame_stats *stats, int32_t *nframes, size_t max_frames) {
    if (!plan) return pirip::fail(-1, "plan is NULL");
    if (n_captures == 0) return 0;
    if (!iq) return pirip::fail(-2, "iq is NULL");
    if (llr && !plan->dev.soft) return pirip::fail(-3, "LLR output requested but cfg.soft == 0");
    if (!pirip::cuda_ok(cudaSetDevice(plan->device), "cudaSetDevice")) return -10;
    pirip_fsk_plan::Work &w = plan->work;
    const size_t cap_bytes = samples_per_capture * bytes_per_sample(plan->dev.fmt);
    const size_t stride = (cap_bytes + 15) / 16 * 16;
    const size_t nb = (size_t)plan->dev.Nbits, nf = (size_t)plan->dev.M * plan->dev.Nsym;
    int rc = 0;
#define CK(call, what) \
    if (rc == 0 && !pirip::cuda_ok((call), what)) rc = -12;
    for (int i = 0; i < pirip_fsk_plan::Work::kStreams; ++i)
        if (!w.streams[i]) CK(cudaStreamCreateWithFlags(&w.streams[i], cudaStreamNonBlocking), "cudaStreamCreate");
    if (rc == 0 && !ensure_cap(w.d_iq, w.cap_iq, stride * n_captures, "cudaMalloc iq")) rc = -12;
    if (rc == 0 && !ensure_cap(w.d_nf, w.cap_nf, sizeof(int32_t) * n_captures, "cudaMalloc nframes")) rc = -12;
    if (rc == 0 && bits && !ensure_cap(w.d_bits, w.cap_bits, n_captures * max_frames * nb, "cudaMalloc bits")) rc = -12;
    if (rc == 0 && llr && !ensure_cap(w.d_llr, w.cap_llr, sizeof(float) * n_captures * max_frames * nb, "cudaMalloc llr")) rc = -12;
    if (rc == 0 && rx_filt && !ensure_cap(w.d_filt, w.cap_filt, sizeof(float) * n_captures * max_frames * nf, "cudaMalloc rx_filt")) rc = -12;
    if (rc == 0 && stats && !ensure_cap(w.d_stats, w.cap_stats, sizeof(pirip_frame_stats) * n_captures * max_frames, "cudaMalloc stats")) rc = -12;

    /* chunking: about one resident wave of CTAs per chunk, at most 8 chunks */
    size_t nchunks = (n_captures + 295) / 296;
    if (nchunks > 8) nchunks = 8;
    if (plan->dev.k2.ok) {
        /* the batch kernel keeps every capture of a launch resident (gmax per SM) and its time is dominated by the
           per-frame latency, not by the capture count (measured: 148 captures 6.0 ms, 1250 captures 11.2 ms), so
           splitting a batch that fits one launch only adds launches */
        const size_t resident = (size_t)plan->dev.k2.num_sms * (size_t)plan->dev.k2.gmax;
        nchunks = (n_captures + resident - 1) / resident;
    }
    const bool fast = plan->fast && plan->mode;
    if (fast) {
        /* one launch for the whole batch (a launch costs 200 frames of per-frame latency however few captures it has), the
           captures streamed in by time slices beside it - see below.  PIRIP_FSK_E2E_CHUNKS > 1 selects plain chunking by
           captures instead (copy chunk k + 1 while chunk k is demodulated). */
        nchunks = (n_captures + fsk_fast_round_captures(plan->fast) - 1) / fsk_fast_round_captures(plan->fast);
        if (const char *e = getenv("PIRIP_FSK_E2E_CHUNKS"))
            if (atoi(e) >= 1 && atoi(e) <= 64) nchunks = (size_t)atoi(e);
        while (nchunks > 1 && n_captures / nchunks < 64) nchunks /= 2;
    }
    if (nchunks < 1) nchunks = 1;
    const size_t per = (n_captures + nchunks - 1) / nchunks;
    plan->last_n = n_captures;
    plan->last_chunks = (n_captures + per - 1) / per;
    /* Streamed ingest (batch kernel, one launch): the captures go to the device in time slices on the copy stream while
       the kernel already demodulates the first ones on the compute stream; the kernel reads d_landed before it fetches
       a frame.  PCIe and the SMs work at the same time although the batch is a single launch. */
    const size_t bps = bytes_per_sample(plan->dev.fmt);
    const size_t nin_max = (size_t)(plan->dev.N + plan->dev.Ts / 4);
    /* the kernel runs about as fast as PCIe delivers (C1: 10.4 ms of kernel, 11 ms of copy for 600 MB), so it trails the copy
       by one slice: what is left to do when the last byte has landed is one slice of work */
    size_t nslices = 24; /* measured, C1 batch of 1250 x 1 s: 8 slices 12.63 ms, 16: 12.05, 24: 12.00, 32: 12.01 per call */
    if (const char *e = getenv("PIRIP_FSK_INGEST_SLICES"))
        if (atoi(e) >= 1 && atoi(e) <= 64) nslices = (size_t)atoi(e);
    while (nslices > 1 && samples_per_capture / nslices < 4 * nin_max) nslices /= 2;
    /* A kernel that waits for copies issued after its launch needs launches to be asynchronous.  Tools that inject into
       the CUDA runtime (Nsight Compute's kernel replay, compute-sanitizer: CUDA_INJECTION64_PATH is set for the process)
       serialise the launch on the host, which would leave the kernel waiting for ever: copy first, then run. */
    const bool injected = getenv("CUDA_INJECTION64_PATH") != nullptr || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr ||
                          (getenv("CUDA_LAUNCH_BLOCKING") != nullptr && atoi(getenv("CUDA_LAUNCH_BLOCKING")) != 0);
    /* Only the fast kernel runs behind the copies by default: when its samples do not arrive (a second client on the GPU, a tool
       that serialises copies behind the kernel) it flags the capture after ~2 s and the capture is re-run by the exact kernel
       once the copies are done - the call still succeeds.  The exact kernel's device-side wait can only fail the launch
       (correctness by time-out), so the exact mode copies first unless PIRIP_FSK_EXACT_STREAM_INGEST=1 asks for the overlap. */
    const bool streamed = plan->dev.k2.ok && nchunks == 1 && nslices > 1 && !injected && getenv("PIRIP_FSK_NO_STREAM_INGEST") == nullptr &&
                          (fast || getenv("PIRIP_FSK_EXACT_STREAM_INGEST") != nullptr);
    if (rc == 0 && streamed) {
        if (!w.d_landed) CK(cudaMalloc(&w.d_landed, sizeof(unsigned long long)), "cudaMalloc landed");
        if (!w.h_marks) CK(cudaHostAlloc(&w.h_marks, 72 * sizeof(unsigned long long), cudaHostAllocDefault), "cudaHostAlloc marks");
        if (!w.ev_reset) CK(cudaEventCreateWithFlags(&w.ev_reset, cudaEventDisableTiming), "cudaEventCreate");
    }
    for (size_t c0 = 0, ci = 0; rc == 0 && c0 < n_captures; c0 += per, ++ci) {
        const size_t nc = (c0 + per <= n_captures) ? per : n_captures - c0;
        cudaStream_t s = w.streams[ci % pirip_fsk_plan::Work::kStreams];
        unsigned char *d_iq = w.d_iq + c0 * stride;
        uint8_t *d_bits = bits ? w.d_bits + c0 * max_frames * nb : nullptr;
        float *d_llr = llr ? w.d_llr + c0 * max_frames * nb : nullptr;
        float *d_filt = rx_filt ? w.d_filt + c0 * max_frames * nf : nullptr;
        pirip_frame_stats *d_stats = stats ? w.d_stats + c0 * max_frames : nullptr;
        int32_t *d_nf = w.d_nf + c0;
        /* frames past the end of a capture are never written by the kernel: define them as zero */
        if (bits) CK(cudaMemsetAsync(d_bits, 0, nc * max_frames * nb, s), "memset bits");
        if (llr) CK(cudaMemsetAsync(d_llr, 0, sizeof(float) * nc * max_frames * nb, s), "memset llr");
        if (rx_filt) CK(cudaMemsetAsync(d_filt, 0, sizeof(float) * nc * max_frames * nf, s), "memset rx_filt");
        if (stats) CK(cudaMemsetAsync(d_stats, 0, sizeof(pirip_frame_stats) * nc * max_frames, s), "memset stats");
        if (streamed) {
            cudaStream_t cs = w.streams[(ci + 1) % pirip_fsk_plan::Work::kStreams]; /* copy stream */
            CK(cudaMemsetAsync(w.d_landed, 0, sizeof(unsigned long long), s), "reset landed");
            CK(cudaEventRecord(w.ev_reset, s), "event record");
            bool launched = false;
            if (rc == 0) {
                rc = batch_dev(plan, d_iq, nc, samples_per_capture, stride, d_bits, d_llr, d_filt, d_stats, d_nf, max_frames, s,
                               w.d_landed, c0, n_captures, ci, fast ? 1 : 3);
                launched = rc == 0;
            }
            CK(cudaStreamWaitEvent(cs, w.ev_reset, 0), "stream wait");
            const size_t per_slice = ((samples_per_capture + nslices - 1) / nslices + 7) / 8 * 8;
            for (size_t k = 0, s0 = 0; rc == 0 && s0 < samples_per_capture; ++k, s0 += per_slice) {
                const size_t s1 = s0 + per_slice < samples_per_capture ? s0 + per_slice : samples_per_capture;
                CK(cudaMemcpy2DAsync(d_iq + s0 * bps, stride, (const unsigned char *)iq + c0 * cap_bytes + s0 * bps, cap_bytes,
                                     (s1 - s0) * bps, nc, cudaMemcpyHostToDevice, cs),
                   "H2D iq slice");
                w.h_marks[k] = (unsigned long long)s1;
                CK(cudaMemcpyAsync(w.d_landed, &w.h_marks[k], sizeof(unsigned long long), cudaMemcpyHostToDevice, cs), "publish landed");
            }
            if (rc != 0 && launched) {
                /* a copy failed after the kernel was launched: release the kernel (it would otherwise wait for samples that never
                   come until its own time-out); the call still reports the error */
                w.h_marks[71] = (unsigned long long)samples_per_capture;
                cudaMemcpyAsync(w.d_landed, &w.h_marks[71], sizeof(unsigned long long), cudaMemcpyHostToDevice, cs);
            }
            if (fast && launched) {
                /* the exact re-run of flagged captures (normally none) reads the captures: after the last copy */
                cudaError_t e1 = cudaEventRecord(w.ev_reset, cs);
                cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(s, w.ev_reset, 0) : e1;
                if (rc == 0 && !pirip::cuda_ok(e2, "wait for the streamed copies")) rc = -12;
                if (rc == 0)
                    rc = batch_dev(plan, d_iq, nc, samples_per_capture, stride, d_bits, d_llr, d_filt, d_stats, d_nf, max_frames, s, nullptr,
                                   c0, n_captures, ci, 2);
            }
        } else {
            CK(cudaMemcpy2DAsync(d_iq, stride, (const unsigned char *)iq + c0 * cap_bytes, cap_bytes, cap_bytes, nc,
                                 cudaMemcpyHostToDevice, s),
               "H2D iq");
            if (rc == 0)
                rc = batch_dev(plan, d_iq, nc, samples_per_capture, stride, d_bits, d_llr, d_filt, d_stats, d_nf, max_frames, s, nullptr,
                               c0, n_captures, ci);
        }
        if (bits) CK(cudaMemcpyAsync(bits + c0 * max_frames * nb, d_bits, nc * max_frames * nb, cudaMemcpyDeviceToHost, s), "D2H bits");
        if (llr) CK(cudaMemcpyAsync(llr + c0 * max_frames * nb, d_llr, sizeof(float) * nc * max_frames * nb, cudaMemcpyDeviceToHost, s), "D2H llr");
        if (rx_filt) CK(cudaMemcpyAsync(rx_filt + c0 * max_frames * nf, d_filt, sizeof(float) * nc * max_frames * nf, cudaMemcpyDeviceToHost, s), "D2H rx_filt");
        if (stats) CK(cudaMemcpyAsync(stats + c0 * max_frames, d_stats, sizeof(pirip_frame_stats) * nc * max_frames, cudaMemcpyDeviceToHost, s), "D2H stats");
        if (nframes) CK(cudaMemcpyAsync(nframes + c0, d_nf, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s), "D2H nframes");
    }
    for (int i = 0; i < pirip_fsk_plan::Work::kStreams; ++i)
        if (w.streams[i] && !pirip::cuda_ok(cudaStreamSynchronize(w.streams[i]), "kernel execution") && rc == 0) rc = -13;
#undef CK
    return rc;
}

/* ---------------------------------------------------------------- one process, several GPUs */

/* Captures are independent, so a batch shards over the devices with no data-path collective (BASELINE.json
   config 5): plan r (one per device, each created with its own device index) demodulates the contiguous block
   [r n / W + min(r, n mod W), ...) - block sizes differ by at most one - from and to the caller's host buffers, all devices at
   once (one host thread per device: the host-buffer call is synchronous).  The "gather" of the decoded bit streams is the
   device-to-host copy of every shard into its place in `bits`.  Returns 0 or the first failing device's code. */
extern "C" int pirip_fsk_demod_batch_multi(pirip_fsk_plan *const *plans, int n_plans, const void *iq, size_t n_captures,
                                           size_t samples_per_capture, uint8_t *bits, float *llr, int32_t *nframes, size_t max_frames) {
    if (!plans || n_plans < 1) return pirip::fail(-1, "no plans");
    for (int r = 0; r < n_plans; ++r) {
        if (!plans[r]) return pirip::fail(-1, "plan is NULL");
        if (plans[r]->dev.Nbits != plans[0]->dev.Nbits || plans[r]->dev.fmt != plans[0]->dev.fmt || plans[r]->dev.N != plans[0]->dev.N)
            return pirip::fail(-3, "the plans of a multi-device call must share one configuration");
    }
    if (n_captures == 0) return 0;
    if (!iq) return pirip::fail(-2, "iq is NULL");
    const size_t cap_bytes = samples_per_capture * bytes_per_sample(plans[0]->dev.fmt);
    const size_t nb = (size_t)plans[0]->dev.Nbits;
    const size_t base = n_captures / (size_t)n_plans, extra = n_captures % (size_t)n_plans;
    std::vector<int> rcs((size_t)n_plans, 0);
    std::vector<std::string> errs((size_t)n_plans);
    std::vector<std::thread> ths;
    for (int r = 0; r < n_plans; ++r) {
        const size_t c0 = (size_t)r * base + ((size_t)r < extra ? (size_t)r : extra);
        const size_t nc = base + ((size_t)r < extra ? 1 : 0);
        if (nc == 0) continue;
        ths.emplace_back([=, &rcs, &errs] {
            rcs[(size_t)r] = pirip_fsk_demod_batch(plans[r], (const unsigned char *)iq + c0 * cap_bytes, nc, samples_per_capture,
                                                   bits ? bits + c0 * max_frames * nb : nullptr, llr ? llr + c0 * max_frames * nb : nullptr, nullptr,
                                                   nullptr, nframes ? nframes + c0 : nullptr, max_frames);
            if (rcs[(size_t)r]) errs[(size_t)r] = pirip_last_error(); /* the error text is per thread */
        });
    }
    for (std::thread &t : ths) t.join();
    for (int r = 0; r < n_plans; ++r)
        if (rcs[(size_t)r]) return pirip::fail(rcs[(size_t)r], "device " + std::to_string(plans[r]->device) + ": " + errs[(size_t)r]);
    return 0;
}

/* ---------------------------------------------------------------- codec2 struct FSK API */

namespace {

struct FskDevice {
    pirip_fsk_plan *plan = nullptr;
    pirip_fsk_cfg cfg;
    cudaStream_t stream = nullptr;
    float *d_state = nullptr;
    COMP *d_in = nullptr;
    uint8_t *d_bits = nullptr;
    float *d_filt = nullptr;
    float *d_Sf = nullptr;
    pirip_frame_stats *d_stats = nullptr;
    /* pinned host staging */
    COMP *h_in = nullptr;
    uint8_t *h_bits = nullptr;
    float *h_filt = nullptr;
    float *h_Sf = nullptr;
    pirip_frame_stats *h_stats = nullptr;
    FskDevStateHdr *h_hdr = nullptr;
    int nin_max = 0;
    bool dirty_cfg = false;
    /* pirip_fsk_demod_frames: many frames per launch */
    COMP *m_in = nullptr, *mh_in = nullptr; /* device samples, pinned staging ring slot */
    uint8_t *m_bits = nullptr, *mh_bits = nullptr;
    float *m_filt = nullptr, *mh_filt = nullptr;
    pirip_frame_stats *m_stats = nullptr, *mh_stats = nullptr;
    int32_t *m_nf = nullptr, *mh_nf = nullptr;
    size_t m_cap_samples = 0, m_cap_frames = 0;
};

[[noreturn]] void die(const char *what) {
    fprintf(stderr, "libpirip_b200: %s: %s\n", what, pirip_last_error());
    abort();
}

void ck(cudaError_t e, const char *what) {
    if (!pirip::cuda_ok(e, what)) die(what);
}

void rebuild_plan(struct FSK *fsk) {
    FskDevice *d = (FskDevice *)fsk->dev;
    pirip_fsk_cfg c = d->cfg;
    c.est_min = fsk->est_min;
    c.est_max = fsk->est_max;
    c.have_limits = !(fsk->est_min == 0 && fsk->est_max == fsk->Fs);
    c.freq_est_type = fsk->freq_est_type;
    c.tone_spacing = fsk->tone_spacing;
    c.f1_tx = fsk->f1_tx;
    pirip_fsk_plan *np = plan_build(c, d->plan ? d->plan->device : pirip::current_device(), fsk->burst_mode, fsk->lock_nin, 1);
    if (!np) die("fsk plan");
    plan_free(d->plan);
    d->plan = np;
    d->cfg = c;
    d->dirty_cfg = false;
}

void write_state_nin(struct FSK *fsk) {
    /* push nin / Sf changes made through the API (clear_estimators, burst mode) to the device */
    FskDevice *d = (FskDevice *)fsk->dev;
    ck(cudaMemcpyAsync(d->h_hdr, d->d_state, sizeof(FskDevStateHdr), cudaMemcpyDeviceToHost, d->stream), "read state");
    ck(cudaStreamSynchronize(d->stream), "sync");
    d->h_hdr->nin = fsk->nin;
    ck(cudaMemcpyAsync(d->d_state, d->h_hdr, sizeof(FskDevStateHdr), cudaMemcpyHostToDevice, d->stream), "write state");
    ck(cudaStreamSynchronize(d->stream), "sync");
}

void demod_common(struct FSK *fsk, uint8_t *rx_bits, float *rx_filt, COMP *fsk_in) {
    FskDevice *d = (FskDevice *)fsk->dev;
    /* fields a caller may poke directly between calls */
    if (d->dirty_cfg || d->plan->dev.burst_mode != fsk->burst_mode || d->plan->dev.lock_nin != fsk->lock_nin ||
        d->plan->dev.est_type != fsk->freq_est_type || d->plan->dev.tone_spacing != fsk->tone_spacing)
        rebuild_plan(fsk);
    ck(cudaSetDevice(d->plan->device), "cudaSetDevice");
    const int nin = fsk->nin;
    const int M = fsk->mode, Nsym = fsk->Nsym, Nbits = fsk->Nbits;
    memcpy(d->h_in, fsk_in, sizeof(COMP) * (size_t)nin);
    ck(cudaMemcpyAsync(d->d_in, d->h_in, sizeof(COMP) * (size_t)nin, cudaMemcpyHostToDevice, d->stream), "H2D samples");
    FskBatchArgs a;
    memset(&a, 0, sizeof a);
    a.iq = (const unsigned char *)d->d_in;
    a.capture_stride_bytes = sizeof(COMP) * (size_t)d->nin_max;
    a.samples_per_capture = (size_t)nin;
    a.max_frames = 1;
    a.n_captures = 1;
    a.tma_ok = 0;
    a.bits = d->d_bits;
    a.rx_filt = d->d_filt;
    a.stats = d->d_stats;
    a.state = d->d_state;
    a.state_stride_floats = fsk_state_floats(fsk->Ndft, M, fsk->Nmem);
    a.Sf_out = d->d_Sf;
    ck(fsk_launch_demod(d->plan->dev, a, 1, d->stream), "fsk_demod_kernel launch");
    pirip::count_launch();
    ck(cudaMemcpyAsync(d->h_bits, d->d_bits, (size_t)Nbits, cudaMemcpyDeviceToHost, d->stream), "D2H bits");
    ck(cudaMemcpyAsync(d->h_filt, d->d_filt, sizeof(float) * (size_t)(M * Nsym), cudaMemcpyDeviceToHost, d->stream), "D2H rx_filt");
    ck(cudaMemcpyAsync(d->h_stats, d->d_stats, sizeof(pirip_frame_stats), cudaMemcpyDeviceToHost, d->stream), "D2H stats");
    ck(cudaMemcpyAsync(d->h_Sf, d->d_Sf, sizeof(float) * (size_t)fsk->Ndft, cudaMemcpyDeviceToHost, d->stream), "D2H Sf");
    ck(cudaMemcpyAsync(d->h_hdr, d->d_state, sizeof(FskDevStateHdr), cudaMemcpyDeviceToHost, d->stream), "D2H state");
    ck(cudaStreamSynchronize(d->stream), "fsk_demod kernel");

    if (rx_bits) memcpy(rx_bits, d->h_bits, (size_t)Nbits);
    if (rx_filt) memcpy(rx_filt, d->h_filt, sizeof(float) * (size_t)(M * Nsym));
    memcpy(fsk->Sf, d->h_Sf, sizeof(float) * (size_t)fsk->Ndft);
    const pirip_frame_stats &s = *d->h_stats;
    for (int m = 0; m < M; ++m) {
        fsk->f_est[m] = s.f_est[m];
        fsk->f2_est[m] = s.f2_est[m];
        fsk->phi_c[m].real = d->h_hdr->phi_c[m].x;
        fsk->phi_c[m].imag = d->h_hdr->phi_c[m].y;
    }
    fsk->norm_rx_timing = s.norm_rx_timing;
    fsk->ppm = s.ppm;
    fsk->SNRest = s.SNRest;
    fsk->v_est = s.v_est;
    fsk->rx_sig_pow = s.rx_sig_pow;
    fsk->rx_nse_pow = s.rx_nse_pow;
    fsk->EbNodB = s.EbNodB;
    fsk->nin = s.nin_next;
    fsk->stats->snr_est = s.snr_est;
    fsk->stats->foff = s.foff;
    fsk->stats->rx_timing = s.rx_timing;
    fsk->stats->clock_offset = s.ppm;
}

}  // namespace

extern "C" struct FSK *fsk_create_hbr(int Fs, int Rs, int M, int P, int Nsym, int f1_tx, int tone_spacing) {
    /* same conditions upstream asserts ([U fsk.c::fsk_create_core]) */
    assert(Fs > 0);
    assert(Rs > 0);
    assert(P > 0);
    assert(Nsym > 0);
    assert((Fs % Rs) == 0);
    assert(((Fs / Rs) % P) == 0);
    assert(P >= 4);
    assert(M == 2 || M == 4);

    struct FSK *fsk = (struct FSK *)calloc(1, sizeof(struct FSK));
    assert(fsk != NULL);
    FskDevice *d = new FskDevice();
    fsk->dev = d;
    pirip_fsk_cfg c;
    memset(&c, 0, sizeof c);
    c.Fs = Fs;
    c.Rs = Rs;
    c.M = M;
    c.P = P;
    c.Nsym = Nsym;
    c.f1_tx = f1_tx;
    c.tone_spacing = tone_spacing;
    c.have_limits = 0;
    c.fmt = PIRIP_FMT_F32;
    d->cfg = c;
    d->plan = plan_build(c, pirip::current_device(), 0, 0, 1);
    if (!d->plan) die("fsk_create_hbr");
    const FskDevCfg &v = d->plan->dev;
    fsk->Ndft = v.Ndft;
    fsk->Fs = Fs;
    fsk->N = v.N;
    fsk->Rs = Rs;
    fsk->Ts = v.Ts;
    fsk->Nmem = v.Nmem;
    fsk->P = P;
    fsk->Nsym = Nsym;
    fsk->Nbits = v.Nbits;
    fsk->f1_tx = f1_tx;
    fsk->tone_spacing = tone_spacing;
    fsk->mode = M;
    fsk->tc = v.tc;
    fsk->est_min = 0;
    fsk->est_max = Fs;
    fsk->est_space = 0.75 * Rs;
    fsk->freq_est_type = 0;
    fsk->nin = v.N;
    fsk->burst_mode = 0;
    fsk->lock_nin = 0;
    fsk->normalise_eye = 1;
    fsk->hann_table = (float *)malloc(sizeof(float) * (size_t)v.Ndft);
    memcpy(fsk->hann_table, d->plan->h_hann.data(), sizeof(float) * (size_t)v.Ndft);
    fsk->Sf = (float *)calloc((size_t)v.Ndft, sizeof(float));
    for (int m = 0; m < M; ++m) {
        fsk->phi_c[m].real = 1.0f;
        fsk->phi_c[m].imag = 0.0f;
    }
    fsk->tx_phase_c.real = 1.0f;
    fsk->tx_phase_c.imag = 0.0f;
    fsk->stats = (struct MODEM_STATS *)calloc(1, sizeof(struct MODEM_STATS));
    fsk_setup_modem_stats(fsk, fsk->stats);

    d->nin_max = v.N + v.Ts / 4 + 16;
    const size_t sfl = fsk_state_floats(v.Ndft, M, v.Nmem);
    ck(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    ck(cudaMalloc(&d->d_state, sizeof(float) * sfl), "cudaMalloc state");
    ck(cudaMalloc(&d->d_in, sizeof(COMP) * (size_t)d->nin_max), "cudaMalloc in");
    ck(cudaMalloc(&d->d_bits, (size_t)v.Nbits), "cudaMalloc bits");
    ck(cudaMalloc(&d->d_filt, sizeof(float) * (size_t)(M * Nsym)), "cudaMalloc filt");
    ck(cudaMalloc(&d->d_Sf, sizeof(float) * (size_t)v.Ndft), "cudaMalloc Sf");
    ck(cudaMalloc(&d->d_stats, sizeof(pirip_frame_stats)), "cudaMalloc stats");
    ck(cudaMallocHost(&d->h_in, sizeof(COMP) * (size_t)d->nin_max), "cudaMallocHost in");
    ck(cudaMallocHost(&d->h_bits, (size_t)v.Nbits), "cudaMallocHost bits");
    ck(cudaMallocHost(&d->h_filt, sizeof(float) * (size_t)(M * Nsym)), "cudaMallocHost filt");
    ck(cudaMallocHost(&d->h_Sf, sizeof(float) * (size_t)v.Ndft), "cudaMallocHost Sf");
    ck(cudaMallocHost(&d->h_stats, sizeof(pirip_frame_stats)), "cudaMallocHost stats");
    ck(cudaMallocHost(&d->h_hdr, sizeof(FskDevStateHdr)), "cudaMallocHost hdr");
    /* initial state: zero memory, unit oscillators, nin = N */
    std::vector<float> st(sfl, 0.0f);
    FskDevStateHdr *h = (FskDevStateHdr *)st.data();
    for (int m = 0; m < 4; ++m) h->phi_c[m] = make_float2(1.0f, 0.0f);
    h->nin = v.N;
    ck(cudaMemcpy(d->d_state, st.data(), sizeof(float) * sfl, cudaMemcpyHostToDevice), "init state");
    ck(cudaMemset(d->d_bits, 0, (size_t)v.Nbits), "memset bits");
    return fsk;
}

extern "C" struct FSK *fsk_create(int Fs, int Rs, int M, int tx_f1, int tx_fs) {
    return fsk_create_hbr(Fs, Rs, M, FSK_DEFAULT_P, FSK_DEFAULT_NSYM, tx_f1, tx_fs);
}

extern "C" void fsk_destroy(struct FSK *fsk) {
    if (!fsk) return;
    FskDevice *d = (FskDevice *)fsk->dev;
    if (d) {
        if (d->plan) cudaSetDevice(d->plan->device);
        cudaFree(d->d_state);
        cudaFree(d->d_in);
        cudaFree(d->d_bits);
        cudaFree(d->d_filt);
        cudaFree(d->d_Sf);
        cudaFree(d->d_stats);
        cudaFreeHost(d->h_in);
        cudaFreeHost(d->h_bits);
        cudaFreeHost(d->h_filt);
        cudaFreeHost(d->h_Sf);
        cudaFreeHost(d->h_stats);
        cudaFreeHost(d->h_hdr);
        cudaFree(d->m_in);
        cudaFree(d->m_bits);
        cudaFree(d->m_filt);
        cudaFree(d->m_stats);
        cudaFree(d->m_nf);
        cudaFreeHost(d->mh_in);
        cudaFreeHost(d->mh_bits);
        cudaFreeHost(d->mh_filt);
        cudaFreeHost(d->mh_stats);
        cudaFreeHost(d->mh_nf);
        if (d->stream) cudaStreamDestroy(d->stream);
        plan_free(d->plan);
        delete d;
    }
    free(fsk->hann_table);
    free(fsk->Sf);
    free(fsk->stats);
    free(fsk);
}

extern "C" uint32_t fsk_nin(struct FSK *fsk) { return (uint32_t)fsk->nin; }

extern "C" void fsk_demod(struct FSK *fsk, uint8_t rx_bits[], COMP fsk_in[]) { demod_common(fsk, rx_bits, NULL, fsk_in); }

extern "C" void fsk_demod_sd(struct FSK *fsk, float rx_filt[], COMP fsk_in[]) { demod_common(fsk, NULL, rx_filt, fsk_in); }

extern "C" void fsk_set_freq_est_limits(struct FSK *fsk, int est_min, int est_max) {
    assert(fsk != NULL);
    assert(est_min >= -fsk->Fs / 2);
    assert(est_max <= fsk->Fs / 2);
    assert(est_max > est_min);
    fsk->est_min = est_min;
    fsk->est_max = est_max;
    ((FskDevice *)fsk->dev)->dirty_cfg = true;
}

extern "C" void fsk_set_freq_est_alg(struct FSK *fsk, int est_type) {
    assert(fsk != NULL);
    fsk->freq_est_type = est_type;
    ((FskDevice *)fsk->dev)->dirty_cfg = true;
}

extern "C" void fsk_clear_estimators(struct FSK *fsk) {
    FskDevice *d = (FskDevice *)fsk->dev;
    for (int i = 0; i < fsk->Ndft; ++i) fsk->Sf[i] = 0;
    fsk->nin = fsk->N;
    ck(cudaMemsetAsync(d->d_state + sizeof(FskDevStateHdr) / sizeof(float), 0, sizeof(float) * (size_t)fsk->Ndft, d->stream), "clear Sf");
    write_state_nin(fsk);
}

extern "C" void fsk_enable_burst_mode(struct FSK *fsk) {
    fsk->burst_mode = 1;
    fsk->nin = fsk->N;
    ((FskDevice *)fsk->dev)->dirty_cfg = true;
    write_state_nin(fsk);
}

extern "C" void fsk_setup_modem_stats(struct FSK *fsk, struct MODEM_STATS *stats) {
    /* [U fsk.c::stats_init] the members the demodulator maintains */
    (void)fsk;
    stats->Nc = 0;
    stats->snr_est = 0;
    stats->sync = 0;
    stats->foff = 0;
    stats->rx_timing = 0;
    stats->clock_offset = 0;
    stats->neyetr = 0;
    stats->neyesamp = 0;
}

extern "C" void fsk_get_demod_stats(struct FSK *fsk, struct MODEM_STATS *stats) {
    stats->clock_offset = fsk->stats->clock_offset;
    stats->snr_est = fsk->stats->snr_est;
    stats->rx_timing = fsk->stats->rx_timing;
    stats->foff = fsk->stats->foff;
    stats->neyesamp = fsk->stats->neyesamp;
    stats->neyetr = fsk->stats->neyetr;
    stats->sync = 0;
    stats->nr = 0;
    stats->Nc = 0;
    for (int m = 0; m < MODE_M_MAX; ++m) stats->f_est[m] = fsk->f_est[m];
}

extern "C" void fsk_stats_normalise_eye(struct FSK *fsk, int normalise_enable) {
    assert(fsk != NULL);
    fsk->normalise_eye = normalise_enable;
}

/* Stream extension (the front end of a live receiver hands over fixed-size chunks, not single frames): as many
   whole modem frames as `n_avail` samples hold, at most max_frames, in ONE launch - the frame loop, the nin decisions and the
   state carried from frame to frame stay on the device; samples go through a pinned staging buffer.  The struct FSK advances
   exactly as by the same number of fsk_demod / fsk_demod_sd calls (same kernel, same state record), so the outputs are the
   same floats.  rx_bits [frames][Nbits], rx_filt [frames][M * Nsym], stats [frames] may each be NULL.  Returns the number of
   frames demodulated (0 when n_avail < fsk_nin()), *consumed the samples they took. */
extern "C" int pirip_fsk_demod_frames(struct FSK *fsk, const COMP *in, size_t n_avail, int max_frames, uint8_t *rx_bits, float *rx_filt,
                                      struct pirip_frame_stats *stats, size_t *consumed) {
    assert(fsk != NULL && in != NULL);
    FskDevice *d = (FskDevice *)fsk->dev;
    if (consumed) *consumed = 0;
    if (max_frames < 1 || n_avail < (size_t)fsk->nin) return 0;
    if (d->dirty_cfg || d->plan->dev.burst_mode != fsk->burst_mode || d->plan->dev.lock_nin != fsk->lock_nin ||
        d->plan->dev.est_type != fsk->freq_est_type || d->plan->dev.tone_spacing != fsk->tone_spacing)
        rebuild_plan(fsk);
    ck(cudaSetDevice(d->plan->device), "cudaSetDevice");
    const int M = fsk->mode, Nsym = fsk->Nsym, Nbits = fsk->Nbits;
    const size_t nin_min = (size_t)(fsk->N - fsk->Ts / 4);
    size_t K = n_avail / nin_min + 1;
    if (K > (size_t)max_frames) K = (size_t)max_frames;
    if (n_avail > d->m_cap_samples) {
        cudaFree(d->m_in);
        cudaFreeHost(d->mh_in);
        ck(cudaMalloc(&d->m_in, sizeof(COMP) * n_avail), "cudaMalloc samples");
        ck(cudaHostAlloc(&d->mh_in, sizeof(COMP) * n_avail, cudaHostAllocDefault), "cudaHostAlloc samples");
        d->m_cap_samples = n_avail;
    }
    if (K > d->m_cap_frames) {
        cudaFree(d->m_bits), cudaFree(d->m_filt), cudaFree(d->m_stats);
        cudaFreeHost(d->mh_bits), cudaFreeHost(d->mh_filt), cudaFreeHost(d->mh_stats);
        ck(cudaMalloc(&d->m_bits, K * (size_t)Nbits), "cudaMalloc bits");
        ck(cudaMalloc(&d->m_filt, sizeof(float) * K * (size_t)(M * Nsym)), "cudaMalloc rx_filt");
        ck(cudaMalloc(&d->m_stats, sizeof(pirip_frame_stats) * K), "cudaMalloc stats");
        ck(cudaHostAlloc(&d->mh_bits, K * (size_t)Nbits, cudaHostAllocDefault), "cudaHostAlloc bits");
        ck(cudaHostAlloc(&d->mh_filt, sizeof(float) * K * (size_t)(M * Nsym), cudaHostAllocDefault), "cudaHostAlloc rx_filt");
        ck(cudaHostAlloc(&d->mh_stats, sizeof(pirip_frame_stats) * K, cudaHostAllocDefault), "cudaHostAlloc stats");
        d->m_cap_frames = K;
    }
    if (!d->m_nf) {
        ck(cudaMalloc(&d->m_nf, sizeof(int32_t)), "cudaMalloc nframes");
        ck(cudaHostAlloc(&d->mh_nf, sizeof(int32_t), cudaHostAllocDefault), "cudaHostAlloc nframes");
    }
    const int nin0 = fsk->nin;
    memcpy(d->mh_in, in, sizeof(COMP) * n_avail);
    ck(cudaMemcpyAsync(d->m_in, d->mh_in, sizeof(COMP) * n_avail, cudaMemcpyHostToDevice, d->stream), "H2D samples");
    FskBatchArgs a;
    memset(&a, 0, sizeof a);
    a.iq = (const unsigned char *)d->m_in;
    a.capture_stride_bytes = sizeof(COMP) * n_avail;
    a.samples_per_capture = n_avail;
    a.max_frames = K;
    a.n_captures = 1;
    a.tma_ok = 0;
    a.bits = d->m_bits;
    a.rx_filt = d->m_filt;
    a.stats = d->m_stats;
    a.nframes = d->m_nf;
    a.state = d->d_state;
    a.state_stride_floats = fsk_state_floats(fsk->Ndft, M, fsk->Nmem);
    a.Sf_out = d->d_Sf;
    ck(fsk_launch_demod(d->plan->dev, a, 1, d->stream), "fsk_demod_kernel launch");
    pirip::count_launch();
    ck(cudaMemcpyAsync(d->mh_nf, d->m_nf, sizeof(int32_t), cudaMemcpyDeviceToHost, d->stream), "D2H nframes");
    ck(cudaMemcpyAsync(d->mh_bits, d->m_bits, K * (size_t)Nbits, cudaMemcpyDeviceToHost, d->stream), "D2H bits");
    ck(cudaMemcpyAsync(d->mh_filt, d->m_filt, sizeof(float) * K * (size_t)(M * Nsym), cudaMemcpyDeviceToHost, d->stream), "D2H rx_filt");
    ck(cudaMemcpyAsync(d->mh_stats, d->m_stats, sizeof(pirip_frame_stats) * K, cudaMemcpyDeviceToHost, d->stream), "D2H stats");
    ck(cudaMemcpyAsync(d->h_Sf, d->d_Sf, sizeof(float) * (size_t)fsk->Ndft, cudaMemcpyDeviceToHost, d->stream), "D2H Sf");
    ck(cudaMemcpyAsync(d->h_hdr, d->d_state, sizeof(FskDevStateHdr), cudaMemcpyDeviceToHost, d->stream), "D2H state");
    ck(cudaStreamSynchronize(d->stream), "fsk_demod kernel");
    const int nf = *d->mh_nf;
    if (nf <= 0) return 0;
    if (rx_bits) memcpy(rx_bits, d->mh_bits, (size_t)nf * (size_t)Nbits);
    if (rx_filt) memcpy(rx_filt, d->mh_filt, sizeof(float) * (size_t)nf * (size_t)(M * Nsym));
    if (stats) memcpy(stats, d->mh_stats, sizeof(pirip_frame_stats) * (size_t)nf);
    memcpy(fsk->Sf, d->h_Sf, sizeof(float) * (size_t)fsk->Ndft);
    const pirip_frame_stats &s = d->mh_stats[nf - 1];
    for (int m = 0; m < M; ++m) {
        fsk->f_est[m] = s.f_est[m];
        fsk->f2_est[m] = s.f2_est[m];
        fsk->phi_c[m].real = d->h_hdr->phi_c[m].x;
        fsk->phi_c[m].imag = d->h_hdr->phi_c[m].y;
    }
    fsk->norm_rx_timing = s.norm_rx_timing;
    fsk->ppm = s.ppm;
    fsk->SNRest = s.SNRest;
    fsk->v_est = s.v_est;
    fsk->rx_sig_pow = s.rx_sig_pow;
    fsk->rx_nse_pow = s.rx_nse_pow;
    fsk->EbNodB = s.EbNodB;
    fsk->nin = s.nin_next;
    fsk->stats->snr_est = s.snr_est;
    fsk->stats->foff = s.foff;
    fsk->stats->rx_timing = s.rx_timing;
    fsk->stats->clock_offset = s.ppm;
    if (consumed) *consumed = (size_t)s.sample_offset + (size_t)(nf > 1 ? d->mh_stats[nf - 2].nin_next : nin0);
    return nf;
}

/* Tx side, host only: continuous-phase M-FSK by a recursive complex oscillator
   ([U fsk.c::fsk_mod_c]); tone index = symbol value as at [R tx/rpitx_fsk.cpp:127-143] */
extern "C" void fsk_mod_c(struct FSK *fsk, COMP fsk_out[], uint8_t tx_bits[], int nbits) {
    const int Ts = fsk->Ts, Fs = fsk->Fs, M = fsk->mode;
    cpx ph = { fsk->tx_phase_c.real, fsk->tx_phase_c.imag };
    cpx step[MODE_M_MAX];
    for (int m = 0; m < M; ++m)
        step[m] = cexpj((float)(2 * M_PI * ((float)(fsk->f1_tx + (fsk->tone_spacing * m)) / (float)(Fs))));
    int bit_i = 0;
    const int nsym = nbits / (M >> 1);
    for (int i = 0; i < nsym; ++i) {
        int sym = 0;
        for (int m = M; m >>= 1;) {
            sym = (sym << 1) | (tx_bits[bit_i] == 1 ? 1 : 0);
            bit_i++;
        }
        for (int j = 0; j < Ts; ++j) {
            ph = cmulf(ph, step[sym]);
            fsk_out[i * Ts + j].real = 2.0f * ph.re;
            fsk_out[i * Ts + j].imag = 2.0f * ph.im;
        }
    }
    const float av = sqrtf((ph.re * ph.re) + (ph.im * ph.im));
    fsk->tx_phase_c.real = ph.re / av;
    fsk->tx_phase_c.imag = ph.im / av;
}

extern "C" void fsk_mod(struct FSK *fsk, float fsk_out[], uint8_t tx_bits[], int nbits) {
    /* real output = real part of the complex modulator */
    const int n = (nbits / (fsk->mode >> 1)) * fsk->Ts;
    std::vector<COMP> tmp((size_t)n);
    fsk_mod_c(fsk, tmp.data(), tx_bits, nbits);
    for (int i = 0; i < n; ++i) fsk_out[i] = tmp[(size_t)i].real;
}

/* Tx side, host only: the drive signal of an external VCO - the tone frequency of every sample, f1_tx + tone_spacing * symbol, Ts
   values per symbol ([U fsk.c::fsk_mod_ext_vco]).  pirip's own transmitter derives its tone table the same way
   ([R tx/rpitx_fsk.cpp:127-143]). */
extern "C" void fsk_mod_ext_vco(struct FSK *fsk, float vco_out[], uint8_t tx_bits[], int nbits) {
    const int Ts = fsk->Ts, M = fsk->mode;
    assert(fsk->f1_tx > 0);
    assert(fsk->tone_spacing > 0);
    int bit_i = 0;
    const int nsym = nbits / (M >> 1);
    for (int i = 0; i < nsym; ++i) {
        int sym = 0;
        for (int m = M; m >>= 1;) {
            sym = (sym << 1) | (tx_bits[bit_i] == 1 ? 1 : 0);
            bit_i++;
        }
        for (int j = 0; j < Ts; ++j) vco_out[i * Ts + j] = (float)fsk->f1_tx + (float)fsk->tone_spacing * (float)sym;
    }
}
