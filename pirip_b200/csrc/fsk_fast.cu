/*
 * fsk_fast.cu — tolerance-mode batch demodulator ("fast path") for the IQ -> bits hot path.
 *
 * One CTA of 64 threads demodulates one capture for its whole length (frame loop inside the kernel); nine CTAs
 * are resident per SM for the benchmark shape, so 1250 captures are one wave.  What it computes per frame is
 * codec2's fsk_demod_core ([U fsk.c::fsk_demod_freq_est / fsk_demod_core])
 * on the samples the fsk_demod CLI loop would hand it ([U fsk_demod.c], [R README.md:105,109],
 * [R test/loopback_rtl_sdr.sh:16]).
 *
 * Numerics contract (north_star: "hard-decision bits and frame indices bit-exact, freq/timing estimates and LLRs
 * within a stated float tolerance"):
 *   EXACT    - input conversion, Hann window, FFT (kiss_fft's butterflies, operand order and float twiddles, no
 *              FMA), magnitudes, the IIR spectrum, peak picking / mask correlation: the tone estimates f_est /
 *              f2_est are the same floats the scalar algorithm produces, for every frame, by construction.
 *   FAST     - down-conversion with a closed-form oscillator (phase kept as a 32-bit fraction of a turn, amplitude
 *              drift of the scalar float recursion modelled), sliding-window integrators, parallel timing-line
 *              sum, FMA everywhere.  Decision variables agree with the scalar algorithm to ~1e-6 relative.
 *   GUARDED  - every discrete decision taken from a FAST quantity (symbol arg-max, the nin slip thresholds at
 *              |norm_rx_timing| = 0.25, the ppm update threshold) is checked against a margin; a capture with a
 *              decision inside the margin is FLAGGED and re-run by the exact kernel (fsk_kernel2.cu) from the
 *              host side of this file's caller, so hard bits / nin / frame counts stay bit-exact.
 *
 * Exact sections use __fmul_rn / __fadd_rn / __fsqrt_rn (never contracted); fast sections call fmaf explicitly.
 * The translation unit is still built with -fmad=false so that nothing is fused by accident.
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "exact_math.h"
#include "fsk_dev.h"
#include "host_util.h"

namespace {

typedef unsigned long long u64;
constexpr int T = 64;          /* threads per CTA (one capture) */
constexpr int TAILPAD_MIN = 8;

struct OscEntry {
    float2 dphi;     /* the float pair the scalar algorithm multiplies by every sample */
    float2 dstride;  /* dphi^(2T) evaluated in double: phase step of a thread between its sample pairs */
    unsigned step32; /* angle of dphi as a 32-bit fraction of a turn */
    float lng;       /* ln |dphi| : per-sample amplitude drift of the scalar float recursion */
    float f;         /* the frequency estimate this entry stands for (Hz) */
    float pad;
};

struct FastCfg {
    int Fs, Rs, P, Nsym, Ts, N, Nmem, Ndft, Nbits, nint, step;
    int st, en, nb, f_zero, est_type, tone_spacing, f1_tx, fixed_nin, soft;
    int nmask, len_mask, mask_pos[FSK_MAX_MASK];
    int log2_ndft, nfac, fac_p[FSK_MAX_FFT_FACTORS], fac_m[FSK_MAX_FFT_FACTORS], fac_fstride[FSK_MAX_FFT_FACTORS];
    int nin_max, nin_min, tailpad, RO, bps;
    float fs_over_ndft, in_scale, g_sym, g_tim;
    float2 tw_l2[9]; /* 256-point FFT: twiddles of the second butterfly level, [k'-1][j-1] = tw[16 k' j] */
    const float *hannS;     /* [Ndft] Hann window times in_scale */
    const float2 *twiddle;  /* [Ndft] kiss_fft twiddles */
    const OscEntry *osc;    /* [Ndft][4] */
    const float2 *ftideal;  /* [P] e^{j 2 pi k / P} */
    const float2 *ftcorr;   /* [T] mean of (timing oscillator of the scalar recursion) / ideal over a thread's run of outputs */
    const float2 *phi_ft;   /* [nint] the timing oscillator exactly as the scalar float recursion produces it (exact replay) */
    int off_raw, raw_bytes, off_fdc, fdcS, off_Sf, off_hann, off_filt, off_ideal, off_misc, smem_bytes;
    int group_stride;         /* shared memory per capture group (bytes, 128-byte multiple) */
    int fft_blocks_per_round; /* blocks transformed at a time: as many as fit the scratch row */
    int prune;                /* 256-point FFT: the estimator window lies inside output bins k < 32 */
    int fft_merged;           /* 256-point FFT: the magnitude row holds every block of a frame, one IIR pass per frame */
    int exit_on_flag;
    int diag; /* PIRIP_FSK_FAST_DIAG: the stats record carries the timing-line strength in `foff` */
};

struct FastArgs {
    const unsigned char *iq;
    size_t stride, nsamples, max_frames, n_captures;
    size_t alloc_end; /* first byte address past the input allocation that bulk copies may touch (16-byte aligned down) */
    unsigned char *bits;
    float *llr, *rx_filt;
    pirip_frame_stats *stats;
    int32_t *nframes;
    int *flags;      /* [n_captures] 0 = clean, else reason bits | (first flagged frame << 8) */
    int *flag_list;  /* compacted indices of flagged captures */
    int *flag_count;
    int tma_ok;
    const u64 *landed; /* optional: samples of every capture already in device memory (the host streams the captures in while the
                          kernel runs); a frame is fetched only when it has landed */
    /* exact-oscillator side chain (fsk_osc_chain_kernel): NULL when it does not run beside this launch */
    u64 *mbox;          /* [n_captures][max_frames + 1] per-frame mailbox: what the side chain needs to run frame f */
    float4 *ckpt;       /* [n_captures][2][ckpt_per_frame][4] records {phi.x, phi.y, tag, ~tag}: oscillator state before every 64th
                           sample, by frame parity; tag = (frame + 1) << 8 | chunk makes a record self-validating (no fences) */
    int *afin;          /* [n_captures] set by the fast kernel when it has finished a capture */
    int *replays;       /* [2] counters: frames replayed exactly, replays that failed (side chain not available in time) */
    int ckpt_per_frame;
    /* launch geometry: the first chain_ctas CTAs run the exact-oscillator chains (chain_wpc warps each), every other CTA
       holds G capture groups */
    int G, chain_ctas, chain_wpc;
    int chain_cpi;      /* PIRIP_FSK_CHAIN_CPI: checkpoint chunks per side-chain iteration (0: the built-in choice) */
    int chain_diag;     /* diagnostics (PIRIP_FSK_CHAIN_DIAG): mailbox words carry their posting time, the chain sums its lag */
    long long *trace;   /* diagnostics (PIRIP_FSK_FAST_TRACE): clock64 at the phase boundaries of frames 20..51 of the first group */
    int cap_index_base; /* added to a capture's index in flag_list: rounds of one call share one list of batch-wide indices */
};

enum { FLAG_SYM = 1, FLAG_TIM = 2, FLAG_PPM = 4, FLAG_NAN = 8, FLAG_INGEST = 16 };

struct Misc {
    u64 bar;
    u64 pos;
    float2 dphi[4], dstride[4];
    unsigned step32[4], acc[4];
    float lng[4], f_est[4], f2_est[4];
    int freqi[4];
    int bmax, nin, nin_next, nold, numffts, low, high, roff, cont, pending, frame, flags, use_tma, parity;
    float fract, omf, norm_rx_timing, ppm, snr_est, rx_timing, SNRest, v_est, rx_sig_pow, rx_nse_pow, EbNodB, foff;
    float red[2][8];
    int nflag_sym;
    /* exact replay */
    int fflags;         /* flags raised in the current frame */
    int nin_prev, rp_ok;
    u64 pos_prev;
    float2 dphi_prev[4];
    float2 xtc;         /* exact timing-line sum */
    int osc_idx[4];
    float rho, dfr;
    int near;
};

#define MBOX_VALID (1ull << 63)
#define MBOX_END (1ull << 62)
#define CKPT_EVERY 128 /* nominal samples per checkpoint chunk: a frame of nin samples has ceil(nin / 128) chunks of nearly equal length.
                          A replay mixes a frame with one thread per (tone, chunk), so the chunk is the serial length of that phase
                          (40 k cycles; 22 k with 61-sample chunks) - but every chunk costs the side chain ~100 cycles beside its
                          steps, and the chain has no cycles to spare: with 20 chunks a frame its lag doubled and the batch took
                          3.50 ms instead of 3.15. */
/* checkpoint chunks of a frame of nin samples: K = ceil(nin / CKPT_EVERY) chunks, the first nin % K one sample longer */
__host__ __device__ __forceinline__ int ck_chunks(int nin) { return (nin + CKPT_EVERY - 1) / CKPT_EVERY; }
__host__ __device__ __forceinline__ int ck_start(int nin, int k) {
    const int K = ck_chunks(nin), base = nin / K, rem = nin - base * K;
    return k * base + (k < rem ? k : rem);
}
__host__ __device__ __forceinline__ int ck_of(int nin, int j) { /* chunk holding sample j */
    const int K = ck_chunks(nin), base = nin / K, rem = nin - base * K;
    const int split = rem * (base + 1);
    return j < split ? j / (base + 1) : rem + (j - split) / base;
}

/* ------------------------------------------------------------------ small device helpers */
/* One CTA holds up to 9 capture GROUPS of T = 64 threads; the threads of a group meet at the group's own hardware barrier
   (bar.sync id, 64), so the groups of a CTA run independently of each other. */
__device__ __forceinline__ void group_sync() { asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x >> 6) + 1), "r"(64) : "memory"); }
#define GSYNC() group_sync()
/* phase time stamps of one group (diagnostics; args.trace is NULL otherwise) */
#define TR(i)                                                                                                         \
    do {                                                                                                              \
        if (args.trace && tid == 0 && grp == 0 && cta_a == 0 && frame >= 20 && frame < 52) args.trace[(frame - 20) * 16 + (i)] = clock64(); \
    } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(u64 *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    if (mbar_try(bar, parity)) return;
    unsigned ns = 32;
    while (!mbar_try(bar, parity)) {
        __nanosleep(ns);
        ns = ns < 256 ? ns * 2 : 256;
    }
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

/* packed f32x2 add / sub (Blackwell FADD2): bit-identical to two scalar IEEE additions (tools/micro/f32x2_exact.cu), half
   the issue slots.  Only additions are packed: ptxas would contract a packed multiply feeding a packed add into FFMA2. */
__device__ __forceinline__ u64 pk(float2 a) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 upk(u64 r) {
    float2 a;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ float2 xadd(float2 a, float2 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return upk(r);
}
__device__ __forceinline__ float2 xsub(float2 a, float2 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return upk(r);
}
/* exact complex multiply, the operation sequence of kiss_fft's C_MUL (a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x): four scalar
   multiplies (never fused), the subtraction written as an addition of the exactly negated product */
__device__ __forceinline__ float2 xmul(float2 a, float2 b) {
    return xadd(make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.x, b.y)), make_float2(__fmul_rn(a.y, -b.y), __fmul_rn(a.y, b.x)));
}

/* [U kiss_fft.c::kf_bfly4], forward transform, one butterfly; s0..s2 are the twiddled inputs */
__device__ __forceinline__ void bfly4_core(float2 &f0, float2 &f1, float2 &f2, float2 &f3, float2 s0, float2 s1, float2 s2) {
    const float2 s5 = xsub(f0, s1);
    f0 = xadd(f0, s1);
    const float2 s3 = xadd(s0, s2);
    const float2 s4 = xsub(s0, s2);
    f2 = xsub(f0, s3);
    f0 = xadd(f0, s3);
    f1 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
    f3 = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
}
__device__ __forceinline__ void bfly4_tw(float2 &f0, float2 &f1, float2 &f2, float2 &f3, float2 t1, float2 t2, float2 t3) {
    bfly4_core(f0, f1, f2, f3, xmul(f1, t1), xmul(f2, t2), xmul(f3, t3));
}
/* twiddles (1, -0): x*(1,-0) == x for every finite x (up to the sign of a zero), so the products are skipped */
__device__ __forceinline__ void bfly4_unit(float2 &f0, float2 &f1, float2 &f2, float2 &f3) { bfly4_core(f0, f1, f2, f3, f1, f2, f3); }
/* the same butterfly when only outputs 0 and 1 (f0, f1) are needed */
__device__ __forceinline__ void bfly4_tw_01(float2 &f0, float2 &f1, float2 f2, float2 f3, float2 t1, float2 t2, float2 t3) {
    const float2 s0 = xmul(f1, t1), s1 = xmul(f2, t2), s2 = xmul(f3, t3);
    const float2 s5 = xsub(f0, s1);
    f0 = xadd(f0, s1);
    const float2 s3 = xadd(s0, s2);
    const float2 s4 = xsub(s0, s2);
    f0 = xadd(f0, s3);
    f1 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
}
/* ... and when only output 0 is needed */
__device__ __forceinline__ float2 bfly4_tw_0(float2 f0, float2 f1, float2 f2, float2 f3, float2 t1, float2 t2, float2 t3) {
    const float2 s0 = xmul(f1, t1), s1 = xmul(f2, t2), s2 = xmul(f3, t3);
    return xadd(xadd(f0, s1), xadd(s0, s2));
}

/* fast complex helpers (FMA) */
__device__ __forceinline__ float2 fcmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
/* x * conj(p) */
__device__ __forceinline__ float2 fmix(float2 x, float2 p) { return make_float2(fmaf(x.x, p.x, x.y * p.y), fmaf(x.y, p.x, -x.x * p.y)); }

/* (float)b - 127 exactly, without the conversion pipe: 0x4B000000 | b is the float 2^23 + b */
__device__ __forceinline__ float u8m127(unsigned w, unsigned sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | sel)) - 8388735.0f; }

/* complex sample at byte pointer p (shared or global), converted as the fsk_demod CLI does ([U fsk_demod.c]); the u8
   value is returned WITHOUT its exact power-of-two scale 1/128 (cfg.in_scale), which the callers fold into a table */
template <int FMT>
__device__ __forceinline__ float2 raw_cvt(const unsigned char *p) {
    if (FMT == PIRIP_FMT_U8) {
        const unsigned w = *reinterpret_cast<const unsigned short *>(p);
        return make_float2(u8m127(w, 0), u8m127(w, 1));
    } else if (FMT == PIRIP_FMT_S16) {
        const short2 s = *reinterpret_cast<const short2 *>(p);
        return make_float2(__fdiv_rn((float)s.x, 825.0f), __fdiv_rn((float)s.y, 825.0f));
    } else {
        return *reinterpret_cast<const float2 *>(p);
    }
}

/* square root for FAST quantities (MUFU, relative error ~1e-7): the translation unit is built with -prec-sqrt=true, where
   sqrtf() is a ~20-instruction IEEE sequence */
__device__ __forceinline__ float fsqrt_fast(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* ------------------------------------------------------------------ exact FFT, 256 points, registers */
/* One 256-point block by 16 threads: two radix-4 levels per register pass (a 16-point sub-transform), one exchange
   through shared memory.  The butterflies, their operands and the order of the additions are kiss_fft's
   ([U kiss_fft.c::kf_work] with factors 4,4,4,4), so the outputs are the floats kiss_fft produces.
   t16: thread within the block's group; src: first raw sample of the block; xch: the group's exchange buffer
   (16 x 17 float2); magout: magnitudes of bins [st, en) of this block. */
template <int FMT>
__device__ __forceinline__ void fft256_block(const FastCfg &cfg, int t16, const unsigned char *src, const float *hannS, float2 *xch,
                                             float *magout, const float2 (&t3)[3], const float2 (&t4)[4][3]) {
    float2 F[16];
    /* leaf order: F[4 k3 + k4] = x[n + 16 k3 + 64 k4] */
#pragma unroll
    for (int k3 = 0; k3 < 4; ++k3)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            const int n = t16 + 16 * k3 + 64 * k4;
            const float2 x = raw_cvt<FMT>(src + (size_t)n * cfg.bps);
            const float h = hannS[n];
            F[4 * k3 + k4] = make_float2(__fmul_rn(h, x.x), __fmul_rn(h, x.y));
        }
#pragma unroll
    for (int k3 = 0; k3 < 4; ++k3) bfly4_unit(F[4 * k3], F[4 * k3 + 1], F[4 * k3 + 2], F[4 * k3 + 3]);
    bfly4_unit(F[0], F[4], F[8], F[12]);
#pragma unroll
    for (int k = 1; k < 4; ++k) bfly4_tw(F[k], F[k + 4], F[k + 8], F[k + 12], cfg.tw_l2[3 * (k - 1)], cfg.tw_l2[3 * (k - 1) + 1], cfg.tw_l2[3 * (k - 1) + 2]);
    /* exchange: thread n holds Y_n[q]; thread q needs Y_n[q] for all n */
#pragma unroll
    for (int q = 0; q < 16; ++q) xch[q * 17 + t16] = F[q];
    const unsigned gmask = 0xffffu << (threadIdx.x & 16); /* the 16 threads of this block's group */
    __syncwarp(gmask);
#pragma unroll
    for (int n = 0; n < 16; ++n) F[n] = xch[t16 * 17 + n];
    __syncwarp(gmask);
    if (cfg.prune) {
        /* the estimator window lies inside bins k < 32 (l = 0, jj < 2): the butterflies that only feed other bins are skipped;
           the ones that remain are the same operations on the same operands */
#pragma unroll
        for (int kA = 0; kA < 4; ++kA) bfly4_tw_01(F[kA], F[kA + 4], F[kA + 8], F[kA + 12], t3[0], t3[1], t3[2]);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const float2 v = bfly4_tw_0(F[4 * jj], F[4 * jj + 1], F[4 * jj + 2], F[4 * jj + 3], t4[jj][0], t4[jj][1], t4[jj][2]);
            const int i = t16 + 16 * jj + 128;
            if (i >= cfg.st && i < cfg.en) magout[i - cfg.st] = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
        }
        return;
    }
    /* level 3 (m = 16, fstride = 4): per kA over kB, twiddles tw[4q], tw[8q], tw[12q] */
#pragma unroll
    for (int kA = 0; kA < 4; ++kA) bfly4_tw(F[kA], F[kA + 4], F[kA + 8], F[kA + 12], t3[0], t3[1], t3[2]);
    /* level 4 (m = 64, fstride = 1): per r = q + 16 jj over kA, twiddles tw[r], tw[2r], tw[3r] */
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) bfly4_tw(F[4 * jj], F[4 * jj + 1], F[4 * jj + 2], F[4 * jj + 3], t4[jj][0], t4[jj][1], t4[jj][2]);
    /* F[l + 4 jj] = X[q + 16 jj + 64 l]; fft shift, magnitude of the bins the estimator looks at */
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int k = t16 + 16 * jj + 64 * l;
            const int i = (k + 128) & 255;
            if (i >= cfg.st && i < cfg.en) {
                const float2 v = F[l + 4 * jj];
                magout[i - cfg.st] = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
            }
        }
}

/* ------------------------------------------------------------------ exact FFT, any power of two, shared memory */
/* nblk blocks (block j at raw sample a0 + j*Ndft/2 of the frame) transformed in buf (nblk x Ndft float2), all
   CTA threads; same butterfly sequence as [U kiss_fft.c::kf_work].  Ends with the magnitudes of bins [st, en). */
template <int FMT>
__device__ void fft_generic(const FastCfg &cfg, int tid, const unsigned char *frame_raw, int j0, int nblk, const float *hannS, float2 *buf,
                            float *mag) {
    const int Ndft = cfg.Ndft, half = Ndft >> 1;
    for (int e = tid; e < nblk * Ndft; e += T) {
        const int b = e >> cfg.log2_ndft, pos = e & (Ndft - 1);
        int n = 0;
        for (int s = 0; s < cfg.nfac; ++s) n += ((pos / cfg.fac_m[s]) % cfg.fac_p[s]) * cfg.fac_fstride[s];
        const float2 x = raw_cvt<FMT>(frame_raw + ((size_t)(j0 + b) * half + n) * cfg.bps);
        const float h = hannS[n];
        buf[e] = make_float2(__fmul_rn(h, x.x), __fmul_rn(h, x.y));
    }
    GSYNC();
    for (int s = cfg.nfac - 1; s >= 0; --s) {
        const int p = cfg.fac_p[s], m = cfg.fac_m[s], fs = cfg.fac_fstride[s];
        const int per_blk = Ndft / p;
        for (int e = tid; e < nblk * per_blk; e += T) {
            const int b = e / per_blk, bi = e - b * per_blk;
            const int grp = bi / m, k = bi - grp * m;
            float2 *F = buf + (size_t)b * Ndft + grp * p * m + k;
            if (p == 4) {
                float2 f0 = F[0], f1 = F[m], f2 = F[2 * m], f3 = F[3 * m];
                bfly4_tw(f0, f1, f2, f3, cfg.twiddle[k * fs], cfg.twiddle[2 * k * fs], cfg.twiddle[3 * k * fs]);
                F[0] = f0;
                F[m] = f1;
                F[2 * m] = f2;
                F[3 * m] = f3;
            } else { /* [U kiss_fft.c::kf_bfly2] */
                const float2 t = xmul(F[m], cfg.twiddle[k * fs]);
                const float2 f0 = F[0];
                F[m] = xsub(f0, t);
                F[0] = xadd(f0, t);
            }
        }
        GSYNC();
    }
    for (int e = tid; e < nblk * cfg.nb; e += T) {
        const int b = e / cfg.nb, i = cfg.st + (e - b * cfg.nb);
        const float2 v = buf[(size_t)b * Ndft + ((i + half) & (Ndft - 1))];
        mag[(size_t)b * cfg.nb + (i - cfg.st)] = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
    }
    GSYNC();
}

/* float log I0(x) ([U mpdecode_core.c::bessi0] polynomial in the log domain), tolerance mode */
__device__ __forceinline__ float logbessi0_f(float x) {
    const float ax = fabsf(x);
    if (ax < 3.75f) {
        float y = x * (1.0f / 3.75f);
        y *= y;
        const float ans = 1.0f + y * (3.5156229f + y * (3.0899424f + y * (1.2067492f + y * (0.2659732f + y * (0.360768e-1f + y * 0.45813e-2f)))));
        return logf(ans);
    }
    const float y = 3.75f / ax;
    const float ans = 0.39894228f + y * (0.1328592e-1f + y * (0.225319e-2f + y * (-0.157565e-2f + y * (0.916281e-2f + y * (-0.2057706e-1f + y * (0.2635537e-1f + y * (-0.1647633e-1f + y * 0.392377e-2f)))))));
    return (ax - 0.5f * logf(ax)) + logf(ans);
}
__device__ __forceinline__ float max_star0_f(float d1, float d2) {
    const float AJIAN = -0.24904163195436f, TJIAN = 2.50681740420944f;
    const float diff = d2 - d1;
    if (diff > TJIAN) return d2;
    if (diff < -TJIAN) return d1;
    if (diff > 0) return d2 + AJIAN * (diff - TJIAN);
    return d1 - AJIAN * (diff + TJIAN);
}

/* ------------------------------------------------------------------ exact oscillator side chain */
/* release / acquire accesses to global memory shared by the two kernels of a launch (gpu scope) */
__device__ __forceinline__ void st_release_u64(u64 *p, u64 v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_s32(int *p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
/* a checkpoint record, read past L1; valid when its tag words are the expected ones */
__device__ __forceinline__ bool ld_ckpt(const float4 *p, int tag, float2 *phi) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    *phi = make_float2(v.x, v.y);
    return __float_as_int(v.z) == tag && __float_as_int(v.w) == ~tag;
}
__device__ __forceinline__ void st_release_s32(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_s32(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

/* [U comp_prim.h::cmult] as fsk_demod_core applies it to the oscillator: never fused */
__device__ __forceinline__ float2 ocmul(float2 a, float2 b) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

/* the same product with its four multiplications as two packed ones: dxx = (d.x, d.x), dyn = (d.y, -d.y);
   (a.x d.x, a.y d.x) and (a.x d.y, -a.y d.y), then x = p1.x + p2.y, y = p1.y + p2.x.  Packed multiplies round each half like the
   scalar instruction (tools/micro/f32x2_exact.cu), a * (-b) is -(a * b) exactly, and the additions stay SCALAR: ptxas would
   contract a packed add with the multiply that feeds it into FFMA2 (tests/test_sass_contract_cpu.py: none in this object).
   Four instructions instead of six - measured on the side chain's SMs: 7.7 cycles of a scheduler per warp and step against 9.4
   for the scalar form, but ~14 cycles of dependent latency against 8.  (A third form - two packed multiplies and a packed fma by
   a run-time 1.0, exact as well, tools/micro/ocmul_pk3.cu - is three instructions and slower than both: 8.3 cycles, 21 of latency.) */
__device__ __forceinline__ float2 ocmul_pk(float2 a, float2 dxx, float2 dyn) {
    u64 p1, p2;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p1) : "l"(pk(a)), "l"(pk(dxx)));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p2) : "l"(pk(a)), "l"(pk(dyn)));
    const float2 q1 = upk(p1), q2 = upk(p2);
    return make_float2(__fadd_rn(q1.x, q2.y), __fadd_rn(q1.y, q2.x));
}

/* The one computation of the path that cannot be reformulated without changing its floats: the recursive oscillator
   phi_c[m] *= dphi_m of [U fsk.c::fsk_demod_core] (one rounding per sample, renormalised once per frame).  The fast kernel
   mixes with a closed-form phase and needs the scalar algorithm's exact phases only when a frame has to be replayed
   exactly - but then from the capture's first sample on.  This kernel runs beside the fast kernel, ONE LANE PER CAPTURE
   (32 independent serial chains per warp), follows the frame sequence the fast kernel publishes in `mbox` (frame length and
   the oscillator table entries chosen by the exact tone estimate) and leaves a checkpoint of phi_c every 64 samples.
   It never blocks the fast kernel; a lane without a mailbox entry idles, the warp sleeps when no lane has work, and the
   kernel gives up after ~100 ms without any (a tool serialising the two kernels): the fast kernel then falls back to
   flagging, i.e. to the exact kernel. */
template <int M>
__device__ __noinline__ void osc_chain_role(const FastCfg &cfg, const FastArgs &args) {
    /* lane = (capture, tone): 32 / M captures per warp, every lane one serial chain of dependent multiply-adds.  The chain
       CTAs have their SMs to themselves (one CTA per SM): a chain warp that shares its scheduler with demodulating warps
       falls behind the frames it is meant to shadow.  A lane takes captures lane0, lane0 + LB, ... one after the other, in
       the order the demodulating groups take them. */
    if ((int)(threadIdx.x >> 5) >= args.chain_wpc) return;
    const int lane = threadIdx.x & 31;
    const int m = lane % M;
    const size_t gw = (size_t)blockIdx.x * args.chain_wpc + (threadIdx.x >> 5);
    const size_t LB = (size_t)args.chain_ctas * args.chain_wpc * (32 / M);
    size_t cap = gw * (32 / M) + lane / M;
    bool done = cap >= args.n_captures, have = false;
    float2 phi = make_float2(1.0f, 0.0f), dphi = make_float2(1.0f, 0.0f);
    int f = 0, k = 0, K = 0, base = 0, rem = 0;
    const u64 *box = args.mbox + (done ? 0 : cap * (args.max_frames + 1));
    float4 *ck = args.ckpt + (done ? 0 : cap * (size_t)(2 * args.ckpt_per_frame * 4)) + m;
    const int *afin = args.afin + (done ? 0 : cap);
    unsigned idle = 0;
    int d_work = 0, d_idle = 0;
    long long d_cw = 0;
    const long long d_t0 = clock64();
    /* step counts of a chunk: base or base + 1 for each of the three frame lengths; every lane runs the smallest count with no
       predicate at all and the few steps by which the counts differ under a per-lane select - no vote across the warp */
    int nlo, nhi;
    {
        const int b0 = cfg.nin_min / ck_chunks(cfg.nin_min), b1 = cfg.N / ck_chunks(cfg.N), b2 = cfg.nin_max / ck_chunks(cfg.nin_max);
        nlo = min(b0, min(b1, b2));
        nhi = max(b0, max(b1, b2)) + 1;
    }
    /* the CTA's nine active warps sit three, two, two and two on the four schedulers, and the chain is bound by the issue rate of
       the fma pipe: measured per oscillator step, two warps of scalar steps share a scheduler at 18.8 cycles, three would at 28;
       the three run the packed form (ocmul_pk) at 23.  The demodulating groups ask for a frame of 1200 steps every ~26 k cycles,
       21.7 cycles a step: the lanes of those three warps still fall behind slowly, which is what a replay waits for. */
    const bool packed = ((threadIdx.x >> 5) & 3) == 0 && args.chain_wpc > 8;
    /* chunks per iteration: half a frame (of the longest form; a lane whose frame ends before the iteration does waits for it) */
    int CPI;
    {
        const int k0 = ck_chunks(cfg.nin_min), k1 = ck_chunks(cfg.N), k2 = ck_chunks(cfg.nin_max);
        CPI = (max(k0, max(k1, k2)) + 1) / 2;
        if (args.chain_cpi > 0) CPI = args.chain_cpi;
    }
    while (true) {
        if (__all_sync(0xffffffffu, done)) break;
        const long long d_ti = clock64();
        /* Poll first, look at the answers last: a lane without a frame reads its mailbox word, a lane on the last chunk of a frame
           the word of the NEXT frame (a lane that is behind goes straight on), any other working lane the capture's "finished"
           flag; the words are consumed after the steps below, whose operands (dphi, have) they would change - so the loads'
           latency hides behind a chunk of steps.  The words are self-contained: relaxed loads. */
        u64 w = 0;
        int fin = 0;
        const bool last = have && K - k <= CPI; /* the frame ends inside this iteration */
        const bool poll_box = !done && (!have || last), poll_fin = !done && have && !last;
        if (poll_box) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(box + f + (last ? 1 : 0)) : "memory");
        if (poll_fin) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(fin) : "l"(afin) : "memory");
        if (__any_sync(0xffffffffu, have)) {
            idle = 0;
            /* CPI chunks per iteration: the polls, the votes and the frame-end normalisation (an IEEE square root and two
               divisions, ~200 issue cycles that the whole warp pays whenever one lane ends a frame) cost 265 cycles per
               iteration - with one chunk per iteration that was a tenth of the chain's time, and the slowest warps could not
               keep the demodulators' pace */
#pragma unroll 1
            for (int kk = 0; kk < CPI; ++kk) {
                if (have) {
                    const int tag = ((f + 1) << 8) | k;
                    ck[((size_t)(f & 1) * args.ckpt_per_frame + k) * 4] = make_float4(phi.x, phi.y, __int_as_float(tag), __int_as_float(~tag));
                }
                const int n = have ? base + (k < rem ? 1 : 0) : 0;
                float2 wp = phi;
                if (packed) {
                    const float2 dxx = make_float2(dphi.x, dphi.x), dyn = make_float2(dphi.y, -dphi.y);
#pragma unroll 8
                    for (int it = 0; it < nlo; ++it) wp = ocmul_pk(wp, dxx, dyn);
                    for (int it = nlo; it < nhi; ++it) {
                        const float2 r = ocmul_pk(wp, dxx, dyn);
                        if (it < n) wp = r;
                    }
                } else {
#pragma unroll 8
                    for (int it = 0; it < nlo; ++it) wp = ocmul(wp, dphi);
                    for (int it = nlo; it < nhi; ++it) {
                        const float2 r = ocmul(wp, dphi);
                        if (it < n) wp = r;
                    }
                }
                if (have) {
                    phi = wp;
                    if (++k == K) {
                        /* [U comp_prim.h::comp_normalize] */
                        const float av = __fsqrt_rn(__fadd_rn(__fmul_rn(phi.x, phi.x), __fmul_rn(phi.y, phi.y)));
                        phi = make_float2(__fdiv_rn(phi.x, av), __fdiv_rn(phi.y, av));
                        ++f;
                        have = false;
                    }
                }
                if (!__any_sync(0xffffffffu, have)) break;
            }
            ++d_work;
            d_cw += clock64() - d_ti;
        } else {
            __nanosleep(500);
            ++d_idle;
            if (++idle > 200000u) break;
        }
        /* the answers.  Tied to the chain's result (never a NaN) so that neither compiler hoists the tests - and with them the
           wait for the loads - above the steps. */
        {
            const int never = (phi.x != phi.x) ? 1 : 0;
            fin += never;
            w += (u64)never;
        }
        bool next = false;
        if (poll_fin && fin) next = true, have = false; /* the demodulator has finished this capture: nobody will ask any more */
        if (poll_box && (w & MBOX_VALID)) {
            if (w & MBOX_END)
                next = true;
            else {
                const int code = (int)((w >> 48) & 3);
                const int nin = code == 0 ? cfg.nin_min : (code == 1 ? cfg.N : cfg.nin_max);
                K = ck_chunks(nin);
                base = nin / K;
                rem = nin - base * K;
                dphi = cfg.osc[(w >> (12 * m)) & 0xfff].dphi;
                if (args.chain_diag && m == 0) {
                    /* diagnostics: how long the word waited for this lane (replays[5] sum, [6] max, [7] count; ~microseconds) */
                    u64 gt;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                    const int lag = (int)(((gt >> 10) - (w >> 50)) & 0xfffull);
                    atomicAdd(args.replays + 5, lag);
                    atomicMax(args.replays + 6, lag);
                    atomicAdd(args.replays + 7, 1);
                }
                have = true;
                k = 0;
            }
        }
        if (next) {
            have = false;
            cap += LB;
            if (cap >= args.n_captures)
                done = true;
            else {
                phi = make_float2(1.0f, 0.0f);
                f = 0;
                box = args.mbox + cap * (args.max_frames + 1);
                ck = args.ckpt + cap * (size_t)(2 * args.ckpt_per_frame * 4) + m;
                afin = args.afin + cap;
            }
        }
    }
    if (args.chain_diag && lane == 0) { /* diagnostics: iterations with / without work, kcycles in the former / in all */
        atomicAdd(args.replays + 8, d_work);
        atomicAdd(args.replays + 9, d_idle);
        atomicAdd(args.replays + 10, (int)(d_cw >> 10));
        atomicAdd(args.replays + 11, (int)((clock64() - d_t0) >> 10));
        atomicAdd(args.replays + 16 + (threadIdx.x >> 5), (int)(d_cw >> 10)); /* per warp slot of the CTA (<= 18) */
        atomicAdd(args.replays + 40 + (threadIdx.x >> 5), d_work);
    }
}

/* ------------------------------------------------------------------ exact replay of one frame */
/* Recompute frame `frame` of capture `cap` with the scalar algorithm's own operations ([U fsk.c::fsk_demod_core],
   SURVEY.md App. A.3): mixing with the recursive oscillator (from the side chain's checkpoints), integrators summed from
   zero in ascending order, the timing-line sum accumulated serially, atan2 / log10 through exact_math.h.  Given the exact
   tone estimates (always exact in this kernel) the results are the floats of the CPU evaluation, so the decisions the
   fast path could not take safely are taken here.  Rare (a few frames per 10^4); all CTA threads take part.
   On success: bits / rx_filt of the frame rewritten, ms-> timing, nin_next, interpolation and per-frame statistics set
   from the exact values, `filt` holds the exact magnitudes; the fast baseband memory of the frame's last `tailpad` samples
   is kept (the next frame continues the FAST oscillator phase).  Returns false when the side chain did not deliver. */
template <int M, int FMT, int SPEC>
__device__ __noinline__ bool exact_frame_replay(const FastCfg &cfg, const FastArgs &args, unsigned char *smem, size_t cap, int frame,
                                                int nin, u64 pos) {
    const int tid = threadIdx.x & (T - 1), lane = tid & 31, warp = tid >> 5;
    float2 *fdc = reinterpret_cast<float2 *>(smem + cfg.off_fdc);
    float *filt = reinterpret_cast<float *>(smem + cfg.off_filt);
    Misc *ms = reinterpret_cast<Misc *>(smem + cfg.off_misc);
    /* the compiled-in shapes (see the kernel): with Ts a constant the 24-term window sums unroll, their loads go out together and
       only the additions stay serial - a replay cost 295 k cycles with run-time sizes, and the slowest capture's replays are the
       tail of the whole launch */
    constexpr bool S1 = SPEC != 0;
    constexpr int SP = SPEC == 1 ? 24 : (SPEC == 2 ? 8 : 1);
    const int Ts = S1 ? 24 : cfg.Ts, P = S1 ? SP : cfg.P, Nsym = S1 ? 50 : cfg.Nsym, Nmem = S1 ? 1248 : cfg.Nmem, fdcS = cfg.fdcS;
    const int step = S1 ? 24 / SP : cfg.step, nint = S1 ? 51 * SP : cfg.nint, bps = S1 ? 2 : cfg.bps;
    const int nold = Nmem - nin;
    const unsigned char *cap_base = args.iq + cap * args.stride;
    /* diagnostics (PIRIP_FSK_CHAIN_DIAG): kcycles per section of the replay, replays[24 + section] */
    long long rp_t = clock64();
#define RP(i)                                                                        \
    if (args.chain_diag && tid == 0) {                                               \
        const long long rp_n = clock64();                                            \
        atomicAdd(args.replays + 24 + (i), (int)((rp_n - rp_t) >> 6));               \
        rp_t = rp_n;                                                                 \
    }

    /* the checkpoint records this frame needs (its own chunks and the chunks of the previous frame's tail) must have been
       written by the side chain: every thread polls one record; all or nothing */
    const float4 *ck = args.ckpt + cap * (size_t)(2 * args.ckpt_per_frame * 4);
    const int nprev = ms->nin_prev;
    const int jt0 = nprev - nold;                         /* first tail sample, in the previous frame's numbering */
    const int ct0 = frame > 0 ? ck_of(nprev, jt0) : 0;    /* its checkpoint chunk */
    const int nct = frame > 0 ? ck_chunks(nprev) - ct0 : 0;
    const int ncn = ck_chunks(nin);
    {
        if (tid == 0) ms->rp_ok = 1;
        GSYNC();
        int ok = (args.ckpt && Nsym <= T && cfg.tailpad <= 128) ? 1 : 0;
        const long long t0 = clock64();
        for (int task = tid; ok && task < M * (nct + ncn); task += T) {
            const int m = task / (nct + ncn), seg = task - m * (nct + ncn);
            const bool prev = seg < nct;
            const int chunk = prev ? ct0 + seg : seg - nct, fr_ = prev ? frame - 1 : frame;
            const float4 *rec = ck + ((size_t)(fr_ & 1) * args.ckpt_per_frame + chunk) * 4 + m;
            float2 dummy;
            while (!ld_ckpt(rec, ((fr_ + 1) << 8) | chunk, &dummy)) {
                if (clock64() - t0 > 6000000ll) { /* ~3 ms */
                    ok = 0;
                    break;
                }
                __nanosleep(200);
            }
        }
        if (tid == 0 && args.replays) atomicAdd(args.replays + 3, (int)((clock64() - t0) >> 10));
        if (!ok) ms->rp_ok = 0;
        GSYNC();
        if (!ms->rp_ok) return false;
    }
    RP(0);

    /* keep the fast baseband memory the next frame's tail is taken from */
    constexpr int KEEP = (M * 128 + T - 1) / T; /* tailpad <= 128 */
    float2 keep[KEEP];
#pragma unroll
    for (int i = 0; i < KEEP; ++i) {
        const int e = tid + i * T;
        const int m = e / cfg.tailpad, k = e - m * cfg.tailpad;
        keep[i] = (m < M) ? fdc[m * fdcS + Nmem - cfg.tailpad + k] : make_float2(0.0f, 0.0f);
    }
    GSYNC();

    /* exact f_dc: the tail comes from the previous frame's last nold samples (its oscillator), then the nin new samples */
    {
        if (frame == 0)
            for (int i = tid; i < M * nold; i += T) fdc[(i / nold) * fdcS + (i % nold)] = make_float2(0.0f, 0.0f);
        for (int task = tid; task < M * (nct + ncn); task += T) {
            const int m = task / (nct + ncn), seg = task - m * (nct + ncn);
            const bool prev = seg < nct;
            const int chunk = prev ? ct0 + seg : seg - nct;
            const int fr_ = prev ? frame - 1 : frame;
            const int jend = prev ? nprev : nin;
            const int jc = ck_start(jend, chunk), jhi = ck_start(jend, chunk + 1); /* the chunk's samples */
            const int jlo = (prev && chunk == ct0) ? jt0 : jc;
            const float2 d = prev ? ms->dphi_prev[m] : ms->dphi[m];
            float2 phi;
            ld_ckpt(ck + ((size_t)(fr_ & 1) * args.ckpt_per_frame + chunk) * 4 + m, 0, &phi); /* validated above */
            const unsigned char *src = cap_base + ((prev ? ms->pos_prev : pos)) * bps;
            /* eight samples are fetched (global memory: the frame's staging buffer already holds the next frame) before their
               eight oscillator steps run */
            for (int j8 = jc; j8 < jhi; j8 += 8) {
                float2 xs[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) xs[u] = raw_cvt<FMT>(src + (size_t)(j8 + u < jhi ? j8 + u : jhi - 1) * bps);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j8 + u;
                    if (j < jhi) {
                        phi = ocmul(phi, d);
                        if (j >= jlo) {
                            float2 x = xs[u];
                            x.x = __fmul_rn(x.x, cfg.in_scale); /* exact: power of two (1 for s16 / f32) */
                            x.y = __fmul_rn(x.y, cfg.in_scale);
                            /* in * conj(phi) as cmult(in, cconj(phi)) evaluates it */
                            const float2 y = make_float2(__fadd_rn(__fmul_rn(x.x, phi.x), __fmul_rn(x.y, phi.y)),
                                                         __fadd_rn(__fmul_rn(x.x, -phi.y), __fmul_rn(x.y, phi.x)));
                            const int idx = prev ? j - jt0 : nold + j;
                            fdc[m * fdcS + idx] = y;
                        }
                    }
                }
            }
        }
    }
    GSYNC();
    RP(1);

    /* window energies and the timing line, accumulated in the scalar order: both warps sum 32 windows each per round, the second
       hands its products over through `filt` (free until the decisions below), warp 0 adds all 64 in index order (lane 0 carries
       the sum).  A round costs ~4 k cycles whoever runs it - the replaying group shares its schedulers with eight busy groups and
       its code is cold - so halving the rounds (39 with warp 0 alone) took 58 k cycles off a replay.  (All 64 threads writing the
       products to a row in global memory and warp 0 adding them from there with 13 loads in flight was slower: the SM's L1 is
       ~5 KB beside 223 KB of shared memory, the row comes back from L2.) */
    {
        float2 tc = make_float2(0.0f, 0.0f);
        float2 *hand = reinterpret_cast<float2 *>(filt);
        const bool two = (size_t)M * Nsym * sizeof(float) >= 32 * sizeof(float2);
        const int per = two ? 64 : 32;
        for (int i0 = 0; i0 < nint; i0 += per) {
            const int i = i0 + (two ? tid : lane);
            float2 prod = make_float2(0.0f, 0.0f);
            if (i < nint && (two || warp == 0)) {
                float ft1 = 0.0f;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 *f = fdc + m * fdcS + i * step;
                    float2 a = make_float2(0.0f, 0.0f);
                    if (S1) {
#pragma unroll
                        for (int k = 0; k < 24; ++k) a = xadd(a, f[k]);
                    } else {
#pragma unroll 4
                        for (int k = 0; k < Ts; ++k) a = xadd(a, f[k]);
                    }
                    ft1 = __fadd_rn(ft1, __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)));
                }
                const float2 w = cfg.phi_ft[i];
                prod = make_float2(__fmul_rn(ft1, w.x), __fmul_rn(ft1, w.y));
            }
            if (two) {
                if (warp == 1) hand[lane] = prod;
                GSYNC();
            }
            if (warp == 0) {
                const int n = nint - i0 < 32 ? nint - i0 : 32;
                for (int l = 0; l < n; ++l) {
                    const float px = __shfl_sync(0xffffffffu, prod.x, l), py = __shfl_sync(0xffffffffu, prod.y, l);
                    tc = xadd(tc, make_float2(px, py));
                }
                if (two) {
                    const int n2 = nint - i0 - 32 < 32 ? nint - i0 - 32 : 32;
                    for (int l = 0; l < n2; ++l) tc = xadd(tc, hand[l]);
                }
            }
            if (two) GSYNC(); /* `hand` is rewritten by the next round */
        }
        RP(2);
        if (warp == 0)
        if (lane == 0) {
            ms->xtc = tc;
            int fl = 0;
            if (isnan(tc.x) || isnan(tc.y)) fl = FLAG_NAN;
            const float norm = (float)((double)em_atan2f(tc.y, tc.x) / (2 * EM_PI));
            const float rx_timing = __fmul_rn(norm, (float)P);
            const float d = __fsub_rn(norm, ms->red[0][7]); /* red[0][7]: norm_rx_timing of the previous frame */
            /* the ppm update threshold |d| < 0.2: d involves the PREVIOUS frame's estimate, a fast value, so a near-tie cannot be
               settled here either; it is counted (see the fast path), not flagged */
            float ppm = ms->red[0][6];
            if ((double)fabsf(d) < .2) {
                const float appm = (float)(1e6 * (double)d / (double)(float)Nsym);
                ppm = (float)(.9 * (double)ppm + .1 * (double)appm);
            }
            ms->ppm = ppm;
            ms->norm_rx_timing = norm;
            int nn = nin;
            if (!cfg.fixed_nin) nn = (double)norm > 0.25 ? cfg.N + Ts / 4 : ((double)norm < -0.25 ? cfg.N - Ts / 4 : cfg.N);
            ms->nin_next = nn;
            const float lowf = floorf(rx_timing);
            ms->low = (int)lowf;
            ms->fract = __fsub_rn(rx_timing, (float)ms->low);
            ms->high = (int)ceilf(rx_timing);
            ms->omf = __fsub_rn(1.0f, ms->fract);
            ms->rx_timing = rx_timing;
            ms->fflags = fl;
        }
    }
    GSYNC();
    RP(3);
    if (ms->fflags) {
        /* not settled (NaN input, ppm threshold): leave the frame flagged */
#pragma unroll
        for (int i = 0; i < KEEP; ++i) {
            const int e = tid + i * T;
            const int m = e / cfg.tailpad, k = e - m * cfg.tailpad;
            if (m < M) fdc[m * fdcS + Nmem - cfg.tailpad + k] = keep[i];
        }
        GSYNC();
        return false;
    }

    /* exact decisions, one thread per symbol; the per-frame sums in symbol order through the warps */
    float mx = 0.0f, nse_t = 0.0f, sq = 0.0f;
    const bool vs = tid < Nsym;
    if (tid < ((Nsym + 31) & ~31) && Nsym <= T) {
        if (vs) {
            const int low = ms->low, high = ms->high;
            const float fr_w = ms->fract, omf = ms->omf;
            const int st = (tid + 1) * P;
            float tmax[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float2 *fa = fdc + m * fdcS + (st + low) * step, *fb = fdc + m * fdcS + (st + high) * step;
                float2 a = make_float2(0.0f, 0.0f), b = make_float2(0.0f, 0.0f);
                if (S1) {
#pragma unroll
                    for (int k = 0; k < 24; ++k) {
                        a = xadd(a, fa[k]);
                        b = xadd(b, fb[k]);
                    }
                } else {
#pragma unroll 4
                    for (int k = 0; k < Ts; ++k) {
                        a = xadd(a, fa[k]);
                        b = xadd(b, fb[k]);
                    }
                }
                float2 t = make_float2(__fmul_rn(omf, a.x), __fmul_rn(omf, a.y));
                t = xadd(t, make_float2(__fmul_rn(fr_w, b.x), __fmul_rn(fr_w, b.y)));
                tmax[m] = __fadd_rn(__fmul_rn(t.x, t.x), __fmul_rn(t.y, t.y));
            }
            mx = tmax[0];
            int symv = 0;
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (tmax[m] > mx) {
                    mx = tmax[m];
                    symv = m;
                }
            if (args.bits) {
                unsigned char *bo = args.bits + (cap * args.max_frames + (size_t)frame) * (size_t)cfg.Nbits;
                if (M == 2)
                    bo[tid] = (unsigned char)(symv == 1);
                else {
                    bo[2 * tid + 1] = (unsigned char)(symv & 1);
                    bo[2 * tid] = (unsigned char)((symv & 2) >> 1);
                }
            }
            float sum = 0.0f;
            float *fo = args.rx_filt ? args.rx_filt + (cap * args.max_frames + (size_t)frame) * (size_t)(M * Nsym) : nullptr;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float rf = __fsqrt_rn(tmax[m]);
                filt[m * Nsym + tid] = rf;
                if (fo) fo[m * Nsym + tid] = rf;
                sum = __fadd_rn(sum, tmax[m]);
            }
            nse_t = __fdiv_rn(__fsub_rn(sum, mx), (float)(M - 1));
            sq = __fsqrt_rn(mx);
        }
    }
    RP(4);
    /* serial sums in symbol order: warp 0 (symbols 0..31), then warp 1 */
    for (int w = 0; w < T / 32; ++w) {
        if (warp == w) {
            float s_sig = w == 0 ? 0.0f : ms->red[1][0], s_nse = w == 0 ? (float)1E-12 : ms->red[1][1];
            float s_std = w == 0 ? 0.0f : ms->red[1][2], s_mean = w == 0 ? 0.0f : ms->red[1][3];
            const int n = Nsym - 32 * w < 32 ? Nsym - 32 * w : 32;
            for (int l = 0; l < n; ++l) {
                const float a = __shfl_sync(0xffffffffu, mx, l), b = __shfl_sync(0xffffffffu, nse_t, l), c = __shfl_sync(0xffffffffu, sq, l);
                s_sig = __fadd_rn(s_sig, a);
                s_nse = __fadd_rn(s_nse, b);
                s_std = __fadd_rn(s_std, a);
                s_mean = __fadd_rn(s_mean, c);
            }
            if (lane == 0) {
                ms->red[1][0] = s_sig;
                ms->red[1][1] = s_nse;
                ms->red[1][2] = s_std;
                ms->red[1][3] = s_mean;
            }
        }
        GSYNC();
    }
    RP(5);
    if (tid == 0) {
        const float nsymf = (float)Nsym;
        const float sig = __fdiv_rn(ms->red[1][0], nsymf), nse = __fdiv_rn(ms->red[1][1], nsymf);
        ms->rx_sig_pow = sig;
        ms->rx_nse_pow = nse;
        ms->v_est = (float)sqrt((double)__fsub_rn(sig, nse));
        ms->SNRest = __fdiv_rn(sig, nse);
        const float meanebno = __fdiv_rn(ms->red[1][3], nsymf);
        float stdebno = __fsub_rn(__fdiv_rn(ms->red[1][2], nsymf), __fmul_rn(meanebno, meanebno));
        stdebno = (double)stdebno > 0.0 ? (float)sqrt((double)stdebno) : 0.0f;
        ms->EbNodB = __fadd_rn(-6.0f, __fmul_rn(20.0f, em_log10f((float)((1e-6 + (double)meanebno) / (1e-6 + (double)stdebno)))));
        ms->snr_est = (float)(.5 * (double)ms->snr_est + .5 * (double)ms->EbNodB); /* still the value before this frame */
        if (args.replays) atomicAdd(args.replays, 1);
    }
    /* the fast baseband memory of the frame's end, for the next frame's tail */
#pragma unroll
    for (int i = 0; i < KEEP; ++i) {
        const int e = tid + i * T;
        const int m = e / cfg.tailpad, k = e - m * cfg.tailpad;
        if (m < M) fdc[m * fdcS + Nmem - cfg.tailpad + k] = keep[i];
    }
    GSYNC();
    RP(6);
#undef RP
    return true;
}

/* ------------------------------------------------------------------ the kernel */
/* SPEC = 1: the benchmark shape (config 1 / 5: Ts = P = 24, Nsym = 50, Ndft = 256) with its sizes as compile-time constants, so
   that the integrator, decision and mixing loops unroll; SPEC = 2: the same with P = 8 (configs 3 / 4 at Fs = 240 k, Rs = 10 k);
   SPEC = 0: every size from cfg. */
template <int M, int FMT, bool STEP1, int SPEC>
__global__ void __launch_bounds__((M == 2 ? 9 : 5) * T, 1) fsk_fast_kernel(const __grid_constant__ FastCfg cfg, const __grid_constant__ FastArgs args) {
    extern __shared__ __align__(128) unsigned char smem_all[];
    /* roles: the first chain_ctas CTAs run the exact-oscillator chains, the others demodulate (one CTA per SM either way) */
    if ((int)blockIdx.x < args.chain_ctas) {
        osc_chain_role<M>(cfg, args);
        return;
    }
    const int grp = threadIdx.x >> 6;
    if (grp >= args.G) return;
    unsigned char *smem = smem_all + (size_t)grp * cfg.group_stride;
    const int tid = threadIdx.x & (T - 1), lane = tid & 31, warp = tid >> 5;
    const int cta_a = (int)blockIdx.x - args.chain_ctas, n_a = (int)gridDim.x - args.chain_ctas;

    unsigned char *raw = smem + cfg.off_raw;
    float2 *fdc = reinterpret_cast<float2 *>(smem + cfg.off_fdc);
    float *Sf = reinterpret_cast<float *>(smem + cfg.off_Sf);
    float *hannS = reinterpret_cast<float *>(smem + cfg.off_hann);
    float *filt = reinterpret_cast<float *>(smem + cfg.off_filt);
    float2 *ideal = reinterpret_cast<float2 *>(smem + cfg.off_ideal);
    Misc *ms = reinterpret_cast<Misc *>(smem + cfg.off_misc);

    constexpr bool S1 = SPEC != 0;                                   /* Ts = 24, Nsym = 50, Ndft = 256, u8 input compiled in */
    constexpr int SP = SPEC == 1 ? 24 : (SPEC == 2 ? 8 : 1);         /* ... with P = 24 (config 1 / 5) or P = 8 (configs 3 / 4) */
    const int Ndft = S1 ? 256 : cfg.Ndft, half = Ndft >> 1, nb = cfg.nb, Ts = S1 ? 24 : cfg.Ts, P = S1 ? SP : cfg.P, Nsym = S1 ? 50 : cfg.Nsym;
    const int Nmem = S1 ? 1248 : cfg.Nmem, fdcS = cfg.fdcS;
    const int step = S1 ? 24 / SP : cfg.step, nint = S1 ? 51 * SP : cfg.nint, bps = S1 ? 2 : cfg.bps;

    /* per-thread constants: FFT twiddles of the last two levels (256-point path), timing-line correction */
    float2 t3[3], t4[4][3];
    if (cfg.log2_ndft == 8) {
        const int q = tid & 15;
#pragma unroll
        for (int j = 0; j < 3; ++j) t3[j] = cfg.twiddle[4 * q * (j + 1)];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int j = 0; j < 3; ++j) t4[jj][j] = cfg.twiddle[(q + 16 * jj) * (j + 1)];
    }
    const float2 ftc = cfg.ftcorr[tid];

    for (int i = tid; i < Ndft; i += T) hannS[i] = cfg.hannS[i];
    for (int i = tid; i <= P; i += T) ideal[i] = cfg.ftideal[i % P];
    /* group (cta_a, grp) takes captures grp * n_a + cta_a, then every n_a * G-th one after it */
    for (size_t cap = (size_t)grp * n_a + cta_a; cap < args.n_captures; cap += (size_t)n_a * args.G) {
    const unsigned char *cap_base = args.iq + cap * args.stride;
    GSYNC();
    for (int i = tid; i < nb; i += T) Sf[i] = 0.0f;
    for (int i = tid; i < M * fdcS; i += T) fdc[i] = make_float2(0.0f, 0.0f);
    if (tid == 0) {
        /* a group takes several captures one after the other when the batch is larger than the resident groups: the barrier
           object of the previous capture is invalidated before it is initialised again (PTX: initialising a live mbarrier is undefined) */
        if (cap != (size_t)grp * n_a + cta_a) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&ms->bar)) : "memory");
        mbar_init(&ms->bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ms->pos = 0;
        ms->nin = cfg.N;
        ms->nin_next = cfg.N;
        ms->frame = 0;
        ms->flags = 0;
        ms->fflags = 0;
        ms->nin_prev = 0;
        ms->pos_prev = 0;
        for (int m = 0; m < 4; ++m) ms->dphi[m] = ms->dphi_prev[m] = make_float2(1.0f, 0.0f);
        ms->nflag_sym = 0;
        ms->parity = 0;
        ms->pending = 0;
        ms->norm_rx_timing = ms->ppm = ms->snr_est = 0.0f;
        for (int m = 0; m < 4; ++m) ms->acc[m] = 0u;
    }
    GSYNC();

    /* issue the copy of raw samples [pos, pos + nin_max) (16-byte aligned superset) into `raw`; thread 0 only */
    auto issue_raw = [&](u64 pos) {
        if (args.landed) {
            /* streamed ingest: wait until the copy engine has delivered this frame (longest form) for every capture.  The wait
               sleeps, costs no issue slots to speak of, and gives up after ~2 s without the samples (a tool that serialises
               copies behind the kernel): the capture is then flagged and re-run by the exact kernel after the copies. */
            u64 need = pos + (u64)cfg.nin_max;
            if (need > args.nsamples) need = args.nsamples;
            const volatile u64 *lw = args.landed;
            if (*lw < need) {
                const long long t0 = clock64();
                while (*lw < need) {
                    __nanosleep(1000);
                    if (clock64() - t0 > 4000000000ll) {
                        ms->flags |= FLAG_INGEST;
                        break;
                    }
                }
            }
        }
        const size_t a = reinterpret_cast<size_t>(cap_base) + pos * bps;
        size_t want_end = a + (size_t)cfg.nin_max * bps;
        const size_t cap_end = reinterpret_cast<size_t>(cap_base) + args.nsamples * bps;
        if (want_end > cap_end) want_end = cap_end;
        const size_t lo = a & ~(size_t)15;
        const size_t hi = (want_end + 15) & ~(size_t)15;
        ms->roff = (int)(a - lo);
        if (args.tma_ok && hi <= args.alloc_end && hi > lo) {
            const unsigned bytes = (unsigned)(hi - lo);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&ms->bar, bytes);
            tma_load_1d(raw, reinterpret_cast<const void *>(lo), bytes, &ms->bar);
            ms->use_tma = 1;
            ms->pending = 1;
        } else {
            ms->use_tma = 0;
            ms->pending = 0;
        }
    };
    if (tid == 0) {
        ms->cont = (args.max_frames > 0 && (u64)cfg.N <= args.nsamples) ? 1 : 0;
        if (ms->cont) issue_raw(0);
    }
    GSYNC();

    while (ms->cont) {
        const int nin = ms->nin, nold = Nmem - nin, frame = ms->frame;
        const u64 pos = ms->pos;
        TR(0);
        /* ---- raw samples of this frame ---- */
        if (ms->use_tma) {
            if (tid == 0) mbar_wait(&ms->bar, (unsigned)ms->parity); /* the others block in the barrier below: no issue slots */
        } else {
            /* generic copy (unaligned input, or the superset would cross the end of the allocation) */
            const unsigned char *g = cap_base + pos * bps;
            const size_t have = (args.nsamples - pos) * bps;
            const size_t nbytes = (size_t)cfg.nin_max * bps < have ? (size_t)cfg.nin_max * bps : have;
            /* past L1: with streamed ingest the copy engine may still be writing the rest of a 32-byte sector that an earlier
               frame's read brought into L1 */
            for (size_t i = tid; i < nbytes; i += T) raw[ms->roff + i] = __ldcg(g + i);
        }
        GSYNC();
        TR(1);
        const unsigned char *fr = raw + ms->roff; /* sample 0 of the frame */

        /* ---- A7: tone frequency estimate (exact) ---- */
        const int numffts = nin / half - 1;
        {
            float2 *scratch = fdc + cfg.tailpad;                       /* row 0 beyond the tail */
            float *mag = reinterpret_cast<float *>(fdc + fdcS + cfg.tailpad); /* row 1 beyond the tail */
            const float tc = 0.1f, c9 = __fsub_rn(1.0f, tc);
            if (cfg.log2_ndft == 8) {
                const int gpr = cfg.fft_blocks_per_round; /* thread groups (blocks) per round: as many as the scratch row holds */
                if (cfg.fft_merged) {
                    /* every block of the frame first (a half-warp reuses its exchange buffer round after round, ordered by the
                       block routine's own __syncwarp), then ONE pass of the IIR over the blocks in order */
                    for (int j0 = 0; j0 < numffts; j0 += gpr) {
                        const int j = j0 + (tid >> 4);
                        if ((tid >> 4) < gpr && j < numffts)
                            fft256_block<FMT>(cfg, tid & 15, fr + (size_t)j * half * bps, hannS, scratch + (tid >> 4) * 272, mag + (size_t)j * nb, t3, t4);
                    }
                    GSYNC();
                    TR(2);
                    for (int b = tid; b < nb; b += T) {
                        float s = Sf[b];
                        for (int r = 0; r < numffts; ++r) s = __fadd_rn(__fmul_rn(s, c9), __fmul_rn(mag[r * nb + b], tc));
                        Sf[b] = s;
                    }
                    GSYNC();
                    TR(3);
                } else
                for (int j0 = 0; j0 < numffts; j0 += gpr) {
                    const int j = j0 + (tid >> 4);
                    if ((tid >> 4) < gpr && j < numffts)
                        fft256_block<FMT>(cfg, tid & 15, fr + (size_t)j * half * bps, hannS, scratch + (tid >> 4) * 272, mag + (size_t)(tid >> 4) * nb, t3, t4);
                    GSYNC();
                    const int nr = numffts - j0 < gpr ? numffts - j0 : gpr;
                    for (int b = tid; b < nb; b += T) {
                        float s = Sf[b];
                        for (int r = 0; r < nr; ++r) s = __fadd_rn(__fmul_rn(s, c9), __fmul_rn(mag[r * nb + b], tc));
                        Sf[b] = s;
                    }
                    GSYNC();
                }
            } else {
                const int per = cfg.fft_blocks_per_round;
                for (int j0 = 0; j0 < numffts; j0 += per) {
                    const int nr = numffts - j0 < per ? numffts - j0 : per;
                    fft_generic<FMT>(cfg, tid, fr, j0, nr, hannS, scratch, mag);
                    for (int b = tid; b < nb; b += T) {
                        float s = Sf[b];
                        for (int r = 0; r < nr; ++r) s = __fadd_rn(__fmul_rn(s, c9), __fmul_rn(mag[r * nb + b], tc));
                        Sf[b] = s;
                    }
                    GSYNC();
                }
            }
            /* peak picking with blanking (warp 0) and the mask correlation (warp 1), on Sf[st, en) */
            float *tmp = mag; /* scratch copy for the blanking */
            /* This block runs once per frame, straight through, and the frame loop's code does not fit the instruction cache: every
               128-byte line of it (8 instructions) is fetched from L2 again each frame, ~350 cycles a line - the first version
               (shuffle trees, an unrolled scan for wide windows) took 6 700 of a frame's 19 000 cycles for 150 executed
               instructions (PIRIP_FSK_FAST_TRACE).  Hence as few instructions as possible: warp-wide REDUX on the float bit
               patterns (the values are >= 0, so they order like unsigned integers), ballot + ffs for "first maximum wins". */
            if (warp == 0) {
                const bool inreg = nb <= 32; /* a window of at most 32 bins: one bin per lane, blanking in registers */
                float myv = 0.0f;
                if (inreg)
                    myv = lane < nb ? Sf[lane] : 0.0f;
                else
                    for (int b = lane; b < nb; b += 32) tmp[b] = Sf[b];
                __syncwarp();
                int fq[4];
#pragma unroll
                for (int it = 0; it < M; ++it) {
                    float best = 0.0f;
                    int bi = 0;
                    if (inreg) {
                        if (myv > 0.0f) best = myv, bi = lane;
                    } else {
#pragma unroll 1
                        for (int b = lane; b < nb; b += 32) {
                            const float v = tmp[b];
                            if (v > best) {
                                best = v;
                                bi = b;
                            }
                        }
                    }
                    /* first maximum wins; nothing above zero leaves bin 0 of the whole spectrum, as the scalar scan does */
                    const unsigned top = __reduce_max_sync(0xffffffffu, __float_as_uint(best));
                    const unsigned who = top ? __reduce_min_sync(0xffffffffu, __float_as_uint(best) == top ? (unsigned)bi : 0xffffffffu) : 0u;
                    const int imax = top ? cfg.st + (int)who : 0;
                    int lo = imax - cfg.f_zero, hi = imax + cfg.f_zero;
                    lo = lo < cfg.st ? cfg.st : lo;
                    hi = hi > cfg.en ? cfg.en : hi;
                    if (inreg) {
                        if (cfg.st + lane >= lo && cfg.st + lane < hi) myv = 0.0f;
                    } else {
#pragma unroll 1
                        for (int b = lo + lane; b < hi; b += 32) tmp[b - cfg.st] = 0.0f;
                        __syncwarp();
                    }
                    fq[it] = imax - half;
                }
                if (lane == 0) {
                    /* ascending sort (upstream writes an insertion sort; the result is the same), static indices: registers */
                    auto cswap = [](int &a, int &b) {
                        const int lo = a < b ? a : b, hi = a < b ? b : a;
                        a = lo;
                        b = hi;
                    };
                    if (M == 2)
                        cswap(fq[0], fq[1]);
                    else {
                        cswap(fq[0], fq[1]);
                        cswap(fq[2], fq[3]);
                        cswap(fq[0], fq[2]);
                        cswap(fq[1], fq[3]);
                        cswap(fq[1], fq[2]);
                    }
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        ms->freqi[m] = fq[m];
                        ms->f_est[m] = __fmul_rn((float)fq[m], cfg.fs_over_ndft);
                    }
                }
                TR(12);
            } else if (warp == 1) {
                float best = 0.0f;
                int bi = 0x7fffffff;
#pragma unroll 1
                for (int b = cfg.st + lane; b < cfg.en - cfg.len_mask; b += 32) {
                    float corr = 0.0f;
#pragma unroll 1
                    for (int k = 0; k < cfg.nmask; ++k) corr = __fadd_rn(corr, Sf[b - cfg.st + cfg.mask_pos[k]]);
                    if (corr > best) {
                        best = corr;
                        bi = b;
                    }
                }
                /* the largest correlation above zero, the lowest bin among equals; none: the window's first bin */
                const unsigned top = __reduce_max_sync(0xffffffffu, __float_as_uint(best));
                const unsigned who = __reduce_min_sync(0xffffffffu, (top && __float_as_uint(best) == top) ? (unsigned)bi : 0x7fffffffu);
                if (lane == 0) {
                    ms->bmax = top ? (int)who : cfg.st;
                    const float foff = (float)((ms->bmax - half) * cfg.Fs / Ndft);
#pragma unroll
                    for (int m = 0; m < M; ++m) ms->f2_est[m] = __fadd_rn(foff, (float)(m * cfg.tone_spacing));
                }
            }
            GSYNC();
            TR(4);
            if (tid < M) {
                const int m = tid;
                const int ei = cfg.est_type ? ms->bmax * 4 + m : (ms->freqi[m] + half) * 4;
                const OscEntry e = cfg.osc[ei];
                ms->dphi_prev[m] = ms->dphi[m];
                ms->dphi[m] = e.dphi;
                ms->dstride[m] = e.dstride;
                ms->step32[m] = e.step32;
                ms->lng[m] = e.lng;
                ms->osc_idx[m] = ei;
            }
            GSYNC();
            TR(5);
            if (tid == 0 && args.mbox) {
                /* tell the exact-oscillator side chain what this frame is */
                u64 w = MBOX_VALID | ((u64)(nin == cfg.nin_min ? 0 : (nin == cfg.N ? 1 : 2)) << 48);
                for (int m = 0; m < M; ++m) w |= (u64)ms->osc_idx[m] << (12 * m);
                if (args.chain_diag) {
                    /* diagnostics: when this word was posted, microseconds modulo 4096 (bits 50..61) */
                    u64 gt;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                    w |= ((gt >> 10) & 0xfffull) << 50;
                }
                st_relaxed_u64(args.mbox + cap * (args.max_frames + 1) + frame, w); /* self-contained word: no fence */
            }
        }

        /* ---- A8: mix the nin new samples down with each tone (fast): pairs of samples, a thread steps 2T samples ---- */
        {
            const int par = nold & 1; /* pairs start where nold + j is even: 16-byte stores */
            const int npairs = (nin + par + 1) >> 1;
            float2 p0[M], p1[M], D[M];
            const int j0 = 2 * tid - par;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const unsigned ph = ms->acc[m] + (unsigned)(j0 + 1) * ms->step32[m];
                float sn, cs;
                sincospif((float)(int)ph * 4.656612873077393e-10f, &sn, &cs); /* 2^-31: pi*x = 2 pi * turns */
                const float amp = cfg.in_scale * fmaf((float)(j0 + 1), ms->lng[m], 1.0f);
                p0[m] = make_float2(cs * amp, sn * amp);
                p1[m] = fcmul(p0[m], ms->dphi[m]);
                D[m] = ms->dstride[m];
            }
            const bool al4 = FMT == PIRIP_FMT_U8 && (((ms->roff - 2 * par) & 3) == 0);
            const unsigned char *p = fr + (ptrdiff_t)j0 * bps;
            float2 *dst = fdc + nold + j0;
            if (al4) {
                /* u8 pairs on 4-byte boundaries (the usual case): the complete pairs in a loop without predicates, the next
                   pair's bytes loaded before the current pair is mixed; the pair before sample 0 (odd nold, thread 0) and the
                   half pair at the end on their own */
                const int qfull = (nin + par) >> 1; /* pairs q < qfull hold two samples of the frame (but for q = 0 when par) */
                /* mode 0: both samples, 1: the first only, 2: the second only */
                auto mixpair = [&](unsigned w, float2 *dd, int mode) {
                    const float2 x0 = make_float2(u8m127(w, 0), u8m127(w, 1)), x1 = make_float2(u8m127(w, 2), u8m127(w, 3));
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float2 y0 = fmix(x0, p0[m]), y1 = fmix(x1, p1[m]);
                        if (mode == 0)
                            *reinterpret_cast<float4 *>(dd + m * fdcS) = make_float4(y0.x, y0.y, y1.x, y1.y);
                        else if (mode == 1)
                            dd[m * fdcS] = y0;
                        else
                            dd[m * fdcS + 1] = y1;
                        p0[m] = fcmul(p0[m], D[m]);
                        p1[m] = fcmul(p1[m], D[m]);
                    }
                };
                int q = tid;
                if (par && tid == 0) {
                    mixpair(*reinterpret_cast<const unsigned *>(p), dst, 2);
                    q += T, p += 2 * T * 2, dst += 2 * T;
                }
                if (q < qfull) {
                    unsigned w = *reinterpret_cast<const unsigned *>(p);
#pragma unroll 2
                    for (; q + T < qfull; q += T, dst += 2 * T) {
                        p += 2 * T * 2;
                        const unsigned wn = *reinterpret_cast<const unsigned *>(p);
                        mixpair(w, dst, 0);
                        w = wn;
                    }
                    mixpair(w, dst, 0);
                    q += T, p += 2 * T * 2, dst += 2 * T;
                }
                if (q == qfull && q < npairs) mixpair(*reinterpret_cast<const unsigned *>(p), dst, 1);
            } else
            for (int q = tid; q < npairs; q += T, p += 2 * T * bps, dst += 2 * T) {
                const int j = 2 * q - par;
                const float2 x0 = j >= 0 ? raw_cvt<FMT>(p) : make_float2(0.0f, 0.0f);
                const float2 x1 = raw_cvt<FMT>(p + bps);
                const bool full = j >= 0 && j + 1 < nin;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 y0 = fmix(x0, p0[m]), y1 = fmix(x1, p1[m]);
                    if (full)
                        *reinterpret_cast<float4 *>(dst + m * fdcS) = make_float4(y0.x, y0.y, y1.x, y1.y);
                    else if (j >= 0)
                        dst[m * fdcS] = y0;
                    else if (j + 1 < nin)
                        dst[m * fdcS + 1] = y1;
                    p0[m] = fcmul(p0[m], D[m]);
                    p1[m] = fcmul(p1[m], D[m]);
                }
            }
        }
        GSYNC();
        TR(6);
        /* raw is free: fetch the next frame (its start is known, its length is not: fetch the longest) */
        if (tid == 0) {
            const u64 npos = pos + (u64)nin;
            ms->parity ^= ms->use_tma;
            if ((size_t)frame + 1 < args.max_frames && npos + (u64)cfg.nin_min <= args.nsamples)
                issue_raw(npos);
            else {
                ms->pending = 0;
                ms->use_tma = 0;
            }
        }

        /* ---- A9/A10: sliding-window integrators, window energies, timing line (fast) ---- */
        {
            float tr = 0.0f, ti = 0.0f, esum = 0.0f;
            const int RO = cfg.RO;
            const int i0 = tid * RO;
            if (i0 < nint) {
                const int cnt = nint - i0 < RO ? nint - i0 : RO;
                if (STEP1) {
                    /* step == 1, Ts even, RO even: two outputs per iteration, 16-byte loads (window starts are even) */
                    const float4 *r4[M];
                    float2 S[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        r4[m] = reinterpret_cast<const float4 *>(fdc + m * fdcS + i0);
                        float2 a = make_float2(0.0f, 0.0f), b = make_float2(0.0f, 0.0f);
                        for (int k = 0; k < (Ts >> 1); ++k) {
                            const float4 v = r4[m][k];
                            a = xadd(a, make_float2(v.x, v.y));
                            b = xadd(b, make_float2(v.z, v.w));
                        }
                        S[m] = xadd(a, b);
                    }
                    const int hTs = Ts >> 1;
                    int ip = i0 % P;
                    /* the samples of the NEXT pair of outputs are fetched before this pair is worked on: the 16-byte loads were
                       consumed right after issue (half of this loop's stall samples were short-scoreboard) */
                    float4 ovv[M], nvv[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) ovv[m] = r4[m][0], nvv[m] = r4[m][hTs];
#pragma unroll 2
                    for (int o = 0; o < cnt; o += 2) {
                        float e0, e1;
                        float4 ovn[M], nvn[M];
                        const int on = o + 2 < cnt ? (o + 2) >> 1 : o >> 1;
#pragma unroll
                        for (int m = 0; m < M; ++m) ovn[m] = r4[m][on], nvn[m] = r4[m][hTs + on];
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            const float4 ov = ovv[m], nv = nvv[m];
                            const float ea = fmaf(S[m].x, S[m].x, S[m].y * S[m].y);
                            const float2 Sn = xadd(S[m], xsub(make_float2(nv.x, nv.y), make_float2(ov.x, ov.y)));
                            const float eb = fmaf(Sn.x, Sn.x, Sn.y * Sn.y);
                            S[m] = xadd(Sn, xsub(make_float2(nv.z, nv.w), make_float2(ov.z, ov.w)));
                            e0 = m ? e0 + ea : ea;
                            e1 = m ? e1 + eb : eb;
                        }
                        const float2 w0 = ideal[ip], w1 = ideal[ip + 1]; /* the table holds P + 1 entries: no wrap inside a pair */
                        if (o + 1 >= cnt) e1 = 0.0f;
                        esum += e0 + e1;
                        tr = fmaf(e0, w0.x, tr);
                        ti = fmaf(e0, w0.y, ti);
                        tr = fmaf(e1, w1.x, tr);
                        ti = fmaf(e1, w1.y, ti);
                        ip += 2;
                        if (ip >= P) ip -= P;
#pragma unroll
                        for (int m = 0; m < M; ++m) ovv[m] = ovn[m], nvv[m] = nvn[m];
                    }
                } else {
                    float2 S[M];
                    const float2 *r[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        r[m] = fdc + m * fdcS + i0 * step;
                        float2 a = make_float2(0.0f, 0.0f);
                        for (int k = 0; k < Ts; ++k) a = xadd(a, r[m][k]);
                        S[m] = a;
                    }
                    int ip = i0 % P;
                    for (int o = 0; o < cnt; ++o) {
                        float e = 0.0f;
#pragma unroll
                        for (int m = 0; m < M; ++m) e += fmaf(S[m].x, S[m].x, S[m].y * S[m].y);
                        const float2 w = ideal[ip];
                        esum += e;
                        tr = fmaf(e, w.x, tr);
                        ti = fmaf(e, w.y, ti);
                        ip = ip + 1 == P ? 0 : ip + 1;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            for (int k = 0; k < step; ++k) S[m] = xadd(S[m], xsub(r[m][Ts + k], r[m][k]));
                            r[m] += step;
                        }
                    }
                }
                /* the scalar recursion's timing oscillator over this run, relative to the ideal one */
                const float cr = fmaf(tr, ftc.x, -ti * ftc.y), ci = fmaf(tr, ftc.y, ti * ftc.x);
                tr = cr;
                ti = ci;
            }
            tr = warp_sum(tr);
            ti = warp_sum(ti);
            esum = warp_sum(esum);
            if (lane == 0) {
                ms->red[warp][0] = tr;
                ms->red[warp][1] = ti;
                ms->red[warp][5] = esum;
            }
        }
        GSYNC();
        TR(7);

        /* ---- A10: timing estimate, nin for the next call, interpolation weights (thread 0) ---- */
        if (tid == 0) {
            const float tcr = ms->red[0][0] + ms->red[1][0], tci = ms->red[0][1] + ms->red[1][1];
            const float esum = ms->red[0][5] + ms->red[1][5];
            ms->rho = __fdividef(fsqrt_fast(tcr * tcr + tci * tci), fmaxf(esum, 1e-30f)); /* strength of the timing line: conditioning of the estimate */
            int fl = 0;
            if (isnan(tcr) || isnan(tci)) fl |= FLAG_NAN;
            /* How far the scalar algorithm's norm_rx_timing can be from this one: both timing-line sums carry an absolute
               error of a few 1e-7 of the summed window energies, so the angle is uncertain by that over the line's relative
               strength rho (measured: |d norm| * rho <= 9.2e-8 over every test configuration; 4e-7 used). */
            const float dn = __fdividef(cfg.g_tim, fmaxf(ms->rho, 1e-9f)) + 2e-7f;
            const float norm = atan2f(tci, tcr) * 0.15915494309189535f;
            const float rx_timing = norm * (float)P;
            const float d = norm - ms->norm_rx_timing;
            ms->red[0][7] = ms->norm_rx_timing; /* for an exact replay of this frame */
            ms->red[0][6] = ms->ppm;
            ms->norm_rx_timing = norm;
            /* |d| within the estimates' own error of the 0.2 update threshold of ppm: the decision only feeds the ppm statistic (a
               tolerance-mode output), and it is less certain than this margin in the scalar algorithm itself - the oracle built
               with glibc's atan2f differs from the exact_math build by 2e-5 in norm_rx_timing.  It is taken from the fast values
               and counted (replays[2]); flagging it cost a whole-capture re-run by the exact kernel, +6 ms for the batch, in
               one batch out of four. */
            if (fabsf(fabsf(d) - 0.2f) < 2.0f * dn && args.replays) atomicAdd(args.replays + 2, 1);
            if (fabsf(d) < 0.2f) ms->ppm = 0.9f * ms->ppm + 0.1f * (1e6f * d / (float)Nsym);
            int nn = nin;
            /* the branch cut of atan2: an angle within the error of +-pi may come out with the other sign in the scalar
               algorithm - norm_rx_timing +0.5 or -0.5, a sampling instant a whole symbol away and the opposite nin slip (three
               captures in 1.9 million; with a locked nin the sampling instant still jumps, so this test stands outside) */
            if (0.5f - fabsf(norm) < dn) fl |= FLAG_TIM;
            if (!cfg.fixed_nin) {
                if (fabsf(fabsf(norm) - 0.25f) < dn) fl |= FLAG_TIM;
                nn = norm > 0.25f ? cfg.N + Ts / 4 : (norm < -0.25f ? cfg.N - Ts / 4 : cfg.N);
            }
            ms->nin_next = nn;
            const float lowf = floorf(rx_timing);
            ms->low = (int)lowf;
            ms->high = (int)ceilf(rx_timing);
            ms->fract = rx_timing - lowf;
            ms->omf = 1.0f - ms->fract;
            /* uncertainty of the interpolation weight.  Next to an integer timing offset the scalar algorithm may interpolate
               between the neighbouring pair of integrator outputs (same value to first order: the resampled output is
               continuous in the timing); the symbol guard then also covers that pair's difference (near = -1 / +1). */
            ms->dfr = (float)P * dn;
            ms->near = ms->fract < ms->dfr ? -1 : (ms->fract > 1.0f - ms->dfr ? 1 : 0);
            ms->rx_timing = rx_timing;
            ms->fflags = fl;
        }
        GSYNC();
        TR(8);

        /* ---- A11: resample at the timing instant, decide, per-symbol statistics (fast, guarded) ---- */
        {
            const int low = ms->low;
            const float fr_w = ms->fract;
            /* t = (1 - fract) W(low) + fract W(high) = sum_{k < Ts} v_k + fract (sum_{Ts <= k < Ts + step} v_k - sum_{k < step} v_k):
               the four lanes of a symbol add a quarter of the window each (packed adds), lane 3 the correction.  The sums of a
               pass (16 symbols per group of 64 threads) end in the first lane of every symbol; four passes are gathered so that
               every lane of a warp holds ONE symbol - lane L of warp w the symbol 16 (L >> 3) + 8 w + (L & 7) of the round - and
               the decision, its guard and the statistics run once per round with full warps. */
            const int per = (Ts + 3) >> 2;
            float s_sig = 0.0f, s_nse = 0.0f, s_mean = 0.0f;
            int fl = 0;
            unsigned char *bo = args.bits ? args.bits + (cap * args.max_frames + (size_t)frame) * (size_t)cfg.Nbits : nullptr;
            float *fo = args.rx_filt ? args.rx_filt + (cap * args.max_frames + (size_t)frame) * (size_t)(M * Nsym) : nullptr;
            const bool two = ms->high != low; /* an integer timing offset: W(high) is W(low), fract is 0 */
            const int near = ms->near;
            const float dfr = ms->dfr;
            const int sub = tid & 3;
            const int k0 = sub * per, k1 = (k0 + per < Ts) ? k0 + per : Ts;
            for (int sym0 = 0; sym0 < Nsym; sym0 += 64) {
                float E[M], U[M], Cc[M]; /* this lane's symbol of the round: tone energies, |dt/dfract|^2, Re(t conj dt/dfract) */
#pragma unroll
                for (int m = 0; m < M; ++m) E[m] = U[m] = Cc[m] = 0.0f;
#pragma unroll 1 /* a real loop: four unrolled copies, each run once per frame, cost more in instruction fetch than they saved */
                for (int pass = 0; pass < 4; ++pass) {
                    if (sym0 + 16 * pass >= Nsym) break; /* uniform */
                    const int sym = sym0 + 16 * pass + (tid >> 2);
                    const bool valid = sym < Nsym;
                    const int base = ((sym + 1) * P + low) * step;
                    float2 t[M], u[M];
                    float uu[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) t[m] = u[m] = make_float2(0.0f, 0.0f), uu[m] = 0.0f;
                    if (valid) {
                        const float2 *f = fdc + base;
                        if (S1 && (low & 1) == 0) {
                            /* an even timing offset: the quarter window starts on a 16-byte boundary - three 16-byte loads per
                               tone (conflict-free: eight lanes cover eight bank groups), the same additions in the same order */
#pragma unroll
                            for (int k = 0; k < 3; ++k)
#pragma unroll
                                for (int m = 0; m < M; ++m) {
                                    const float4 v4 = *reinterpret_cast<const float4 *>(f + m * fdcS + k0 + 2 * k);
                                    t[m] = xadd(xadd(t[m], make_float2(v4.x, v4.y)), make_float2(v4.z, v4.w));
                                }
                        } else if (S1) {
#pragma unroll
                            for (int k = 0; k < 6; ++k)
#pragma unroll
                                for (int m = 0; m < M; ++m) t[m] = xadd(t[m], f[m * fdcS + k0 + k]);
                        } else {
#pragma unroll 2
                            for (int k = k0; k < k1; ++k)
#pragma unroll
                                for (int m = 0; m < M; ++m) t[m] = xadd(t[m], f[m * fdcS + k]);
                        }
                        if (sub == 3 && two) {
                            for (int k = 0; k < step; ++k)
#pragma unroll
                                for (int m = 0; m < M; ++m) u[m] = xadd(u[m], xsub(f[m * fdcS + Ts + k], f[m * fdcS + k]));
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                t[m].x = fmaf(fr_w, u[m].x, t[m].x);
                                t[m].y = fmaf(fr_w, u[m].y, t[m].y);
                                uu[m] = fmaf(u[m].x, u[m].x, u[m].y * u[m].y); /* |dt/dfract|^2 */
                            }
                        }
                        if (sub == 3 && near != 0) {
                            /* the neighbouring pair: W(low) - W(low-1) below, W(low+2) - W(low+1) above */
                            const float2 *g = near < 0 ? f - step : f + (two ? step : 0);
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                float2 u2 = make_float2(0.0f, 0.0f);
                                for (int k = 0; k < step; ++k) u2 = xadd(u2, xsub(g[m * fdcS + Ts + k], g[m * fdcS + k]));
                                uu[m] = fmaxf(uu[m], fmaf(u2.x, u2.x, u2.y * u2.y));
                            }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        t[m].x += __shfl_xor_sync(0xffffffffu, t[m].x, 1);
                        t[m].y += __shfl_xor_sync(0xffffffffu, t[m].y, 1);
                        t[m].x += __shfl_xor_sync(0xffffffffu, t[m].x, 2);
                        t[m].y += __shfl_xor_sync(0xffffffffu, t[m].y, 2);
                    }
                    /* gather: the symbol of lanes 4k .. 4k+3 goes to lane 8 pass + k (energies from any of the four lanes, the
                       correction term from the symbol's lane 3) */
                    const int src = 4 * (lane & 7);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float e = fmaf(t[m].x, t[m].x, t[m].y * t[m].y);
                        /* Re(t conj u) = half of d|t|^2 / dfract: only the symbol's lane 3 holds u (zero elsewhere) */
                        const float c = fmaf(t[m].x, u[m].x, t[m].y * u[m].y);
                        const float ge = __shfl_sync(0xffffffffu, e, src), gu = __shfl_sync(0xffffffffu, uu[m], src + 3);
                        const float gc = __shfl_sync(0xffffffffu, c, src + 3);
                        if ((lane >> 3) == pass) E[m] = ge, U[m] = gu, Cc[m] = gc;
                    }
                }
                const int sym = sym0 + 16 * (lane >> 3) + 8 * warp + (lane & 7);
                if (sym < Nsym) {
                    float mx = E[0], sum = 0.0f, second = -1.0f;
                    int symv = 0;
#pragma unroll
                    for (int m = 0; m < M; ++m) sum += E[m];
#pragma unroll
                    for (int m = 1; m < M; ++m)
                        if (E[m] > mx) {
                            mx = E[m];
                            symv = m;
                        }
                    float uub = U[0], cb = Cc[0];
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        if (m == symv) uub = U[m], cb = Cc[m];
                        if (m != symv && E[m] > second) second = E[m];
                    }
                    /* the scalar algorithm's energies differ from these by rounding (g_sym, relative) and by the uncertainty dfr
                       of the interpolation weight.  All tones are resampled with the SAME weight, so what a decision sees is the
                       change of a DIFFERENCE: |t + d u|^2 = |t|^2 + 2 d Re(t conj u) + d^2 |u|^2 gives, against every other tone m,
                       |change of (e_best - e_m)| <= 2 dfr |c_best - c_m| + dfr^2 (|u_best|^2 + |u_m|^2) with c = Re(t conj u) - far
                       below the sum of the tones' own worst cases 2 dfr |t| |u|, which flagged several times as many frames.
                       Next to an integer timing offset (near != 0) the scalar algorithm may interpolate the neighbouring pair;
                       there the tones' own worst cases stay (2 sqrt(a b) <= a + b first, without square roots). */
                    bool tie = false;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        if (m == symv) continue;
                        const float gs = cfg.g_sym * (mx + E[m]);
                        const float lin = near == 0 ? 2.0f * fabsf(cb - Cc[m]) : (mx + E[m]) + (uub + U[m]);
                        if (mx - E[m] <= gs + dfr * (lin + (uub + U[m]) * dfr) && mx > 0.0f) { /* a 0 == 0 tie (zero input) is exact */
                            const float tight = near == 0 ? lin : 2.0f * (sqrtf(mx * uub) + sqrtf(E[m] * U[m]));
                            tie = tie || mx - E[m] <= gs + dfr * (tight + (uub + U[m]) * dfr);
                        }
                    }
                    if (tie) {
                        fl |= FLAG_SYM;
                        atomicAdd(&ms->nflag_sym, 1);
                    }
                    if (bo) {
                        if (M == 2)
                            bo[sym] = (unsigned char)(symv == 1);
                        else {
                            bo[2 * sym + 1] = (unsigned char)(symv & 1);
                            bo[2 * sym] = (unsigned char)((symv & 2) >> 1);
                        }
                    }
                    float rmx = 0.0f;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float rf = fsqrt_fast(E[m]);
                        filt[m * Nsym + sym] = rf;
                        if (fo) fo[m * Nsym + sym] = rf;
                        if (m == symv) rmx = rf;
                    }
                    s_sig += mx;
                    s_nse += (sum - mx) * (1.0f / (float)(M - 1));
                    s_mean += rmx;
                }
            }
            s_sig = warp_sum(s_sig);
            s_nse = warp_sum(s_nse);
            s_mean = warp_sum(s_mean);
            if (lane == 0) {
                ms->red[warp][2] = s_sig;
                ms->red[warp][3] = s_nse;
                ms->red[warp][4] = s_mean;
            }
            if (fl) atomicOr(&ms->fflags, fl);
        }
        GSYNC();
        TR(9);

        /* ---- a decision inside its margin: replay the frame exactly (rare) ---- */
        bool replayed = false;
        if (ms->fflags && args.mbox) {
            const long long t_rp = clock64();
            replayed = exact_frame_replay<M, FMT, SPEC>(cfg, args, smem, cap, frame, nin, pos);
            if (!replayed && tid == 0 && args.replays) atomicAdd(args.replays + 1, 1);
            if (tid == 0 && args.replays) atomicAdd(args.replays + 4, (int)((clock64() - t_rp) >> 10)); /* diagnostics: kcycles spent in replays */
        }

        /* ---- stats of the frame, the record, loop control (thread 0) ---- */
        if (tid == 0 && !replayed) {
            const float rnsym = 1.0f / (float)Nsym;
            const float sig = (ms->red[0][2] + ms->red[1][2]) * rnsym;
            const float nse = (1e-12f + ms->red[0][3] + ms->red[1][3]) * rnsym;
            float meanebno = (ms->red[0][4] + ms->red[1][4]) * rnsym;
            float stdebno = sig - meanebno * meanebno;
            stdebno = stdebno > 0.0f ? fsqrt_fast(stdebno) : 0.0f;
            ms->rx_sig_pow = sig;
            ms->rx_nse_pow = nse;
            ms->v_est = fsqrt_fast(sig - nse);
            ms->SNRest = __fdividef(sig, nse);
            ms->EbNodB = -6.0f + 20.0f * log10f(__fdividef(1e-6f + meanebno, 1e-6f + stdebno));
            ms->snr_est = 0.5f * ms->snr_est + 0.5f * ms->EbNodB;
        }
        if (tid == 0) {
            if (ms->fflags && !replayed) ms->flags |= ms->fflags;
            ms->fflags = 0;
            const float sig = ms->rx_sig_pow, nse = ms->rx_nse_pow;
            const float *fe = cfg.est_type ? ms->f2_est : ms->f_est;
            ms->foff = (float)((cfg.f1_tx + cfg.f1_tx + cfg.tone_spacing) / 2) - (fe[0] + fe[1]) / 2.0f;
            if (args.stats) {
                pirip_frame_stats *st = &args.stats[cap * args.max_frames + (size_t)frame];
                for (int m = 0; m < 4; ++m) {
                    st->f_est[m] = m < M ? ms->f_est[m] : 0.0f;
                    st->f2_est[m] = m < M ? ms->f2_est[m] : 0.0f;
                }
                st->norm_rx_timing = ms->norm_rx_timing;
                st->ppm = ms->ppm;
                st->SNRest = ms->SNRest;
                st->v_est = ms->v_est;
                st->rx_sig_pow = sig;
                st->rx_nse_pow = nse;
                st->EbNodB = ms->EbNodB;
                st->snr_est = ms->snr_est;
                st->foff = cfg.diag ? ms->rho : ms->foff;
                st->rx_timing = ms->rx_timing;
                st->nin_next = ms->nin_next;
                st->sample_offset = (int32_t)pos;
            }
            /* advance */
            for (int m = 0; m < M; ++m) ms->acc[m] += (unsigned)nin * ms->step32[m];
            ms->nin_prev = nin;
            ms->pos_prev = pos;
            ms->pos = pos + (u64)nin;
            ms->nin = ms->nin_next;
            ms->frame = frame + 1;
            int cont = ((size_t)(frame + 1) < args.max_frames && ms->pos + (u64)ms->nin <= args.nsamples) ? 1 : 0;
            if ((ms->flags && cfg.exit_on_flag) || (ms->flags & FLAG_INGEST)) cont = 0;
            if (ms->flags && (ms->flags >> 8) == 0) ms->flags |= ((frame + 1) & 0xfff) << 8;
            ms->cont = cont;
        }
        GSYNC();
        TR(10);

        /* ---- A12: LLRs (fast) ---- */
        if (cfg.soft && args.llr) {
            float *lo = args.llr + (cap * args.max_frames + (size_t)frame) * (size_t)cfg.Nbits;
            const float scale = 2.0f * ms->SNRest, v = ms->v_est;
            for (int sym = tid; sym < Nsym; sym += T) {
                constexpr int BPS = M == 2 ? 1 : 2;
                float num[2] = { -1000000.0f, -1000000.0f }, den[2] = { -1000000.0f, -1000000.0f };
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float yr = filt[m * Nsym + sym];
                    const float env = sqrtf(yr * yr / (v * v));
                    const float metric = logbessi0_f(scale * env);
                    int mask = 1 << (BPS - 1);
#pragma unroll
                    for (int k = 0; k < BPS; ++k) {
                        if (mask & m)
                            num[k] = max_star0_f(num[k], metric);
                        else
                            den[k] = max_star0_f(den[k], metric);
                        mask >>= 1;
                    }
                }
#pragma unroll
                for (int k = 0; k < BPS; ++k) lo[sym * BPS + k] = -(num[k] - den[k]);
            }
        }
        /* ---- keep the tail of the baseband memory for the next frame ---- */
        {
            const int nold2 = Nmem - ms->nin; /* ms->nin is already the next frame's nin */
#pragma unroll
            for (int m = 0; m < M; ++m)
                for (int k = tid; k < nold2; k += T) fdc[m * fdcS + k] = fdc[m * fdcS + (Nmem - nold2) + k];
        }
        GSYNC();
        TR(11);
    }

    /* a bulk copy still in flight must land before the CTA retires */
    if (ms->pending && ms->use_tma) mbar_wait(&ms->bar, (unsigned)ms->parity);
    if (tid == 0) {
        if (args.mbox) {
            st_relaxed_u64(args.mbox + cap * (args.max_frames + 1) + ms->frame, MBOX_VALID | MBOX_END);
            st_relaxed_s32(args.afin + cap, 1);
        }
        if (args.nframes) args.nframes[cap] = ms->frame;
        int fl = ms->flags;
        if (fl) fl |= (ms->nflag_sym > 2047 ? 2047 : ms->nflag_sym) << 20; /* bits 20..30: symbols inside the margin */
        if (args.flags) args.flags[cap] = fl;
        if (fl && args.flag_list && args.flag_count) {
            const int slot = atomicAdd(args.flag_count, 1);
            args.flag_list[slot] = (int)cap + args.cap_index_base;
        }
    }
    } /* captures of this group */
}

/* ------------------------------------------------------------------ host side */
struct cpx {
    float re, im;
};
inline cpx cmulf(cpx a, cpx b) { return cpx{ a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; }

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

struct FskFastPlan {
    FastCfg cfg;
    int M, fmt, device;
    float *d_hannS = nullptr;
    float2 *d_ideal = nullptr, *d_corr = nullptr;
    OscEntry *d_osc = nullptr;
    int *d_flags = nullptr, *d_list = nullptr, *d_count = nullptr; /* d_count: one counter per chunk of a host-buffer call */
    size_t cap_flags = 0;
    /* exact-oscillator side chain */
    u64 *d_mbox = nullptr;
    float4 *d_ckpt = nullptr;
    int *d_afin = nullptr, *d_replays = nullptr;
    long long *d_trace = nullptr; /* PIRIP_FSK_FAST_TRACE */
    size_t cap_mbox = 0, cap_chain = 0;
    int ckpt_per_frame = 0;
    bool chain = true;
    int gmax = 1;   /* capture groups per CTA: what shared memory and registers allow (at most 9 / 5) */
    int num_sms = 148;
    std::string desc;
};

template <int M, int FMT, bool STEP1, int SPEC>
static cudaError_t fast_launch_t(const FskFastPlan *p, const FastArgs &a, cudaStream_t stream, int mode, unsigned grid, unsigned block, size_t smem,
                                 int *regs) {
    auto kern = fsk_fast_kernel<M, FMT, STEP1, SPEC>;
    if (mode == 1) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (mode == 2) {
        cudaFuncAttributes at;
        cudaError_t e = cudaFuncGetAttributes(&at, kern);
        if (e == cudaSuccess) *regs = at.numRegs;
        return e;
    }
    kern<<<grid, block, smem, stream>>>(p->cfg, a);
    return cudaGetLastError();
}

/* mode 0: launch, 1: set the shared-memory attribute, 2: query registers per thread */
static cudaError_t fast_dispatch(const FskFastPlan *p, const FastArgs &a, cudaStream_t stream, int mode, unsigned grid = 0, unsigned block = 0,
                                 size_t smem = 0, int *regs = nullptr) {
    /* the paired integrator needs step == 1 and even Ts, RO, nint */
    const bool s1 = p->cfg.step == 1 && (p->cfg.Ts & 1) == 0 && (p->cfg.RO & 1) == 0 && (p->cfg.nint & 1) == 0;
    /* the benchmark shape has a kernel of its own with the sizes compiled in */
    const FastCfg &c = p->cfg;
    if (s1 && p->M == 2 && p->fmt == PIRIP_FMT_U8 && c.Ts == 24 && c.P == 24 && c.Nsym == 50 && c.Ndft == 256 && c.Nmem == 1248 && c.nint == 1224 &&
        c.bps == 2 && !getenv("PIRIP_FSK_FAST_NO_SPEC"))
        return fast_launch_t<2, PIRIP_FMT_U8, true, 1>(p, a, stream, mode, grid, block, smem, regs);
    if (!s1 && p->fmt == PIRIP_FMT_U8 && c.Ts == 24 && c.P == 8 && c.Nsym == 50 && c.Ndft == 256 && c.Nmem == 1248 && c.nint == 408 && c.bps == 2 &&
        !getenv("PIRIP_FSK_FAST_NO_SPEC")) {
        if (p->M == 2) return fast_launch_t<2, PIRIP_FMT_U8, false, 2>(p, a, stream, mode, grid, block, smem, regs);
        if (p->M == 4) return fast_launch_t<4, PIRIP_FMT_U8, false, 2>(p, a, stream, mode, grid, block, smem, regs);
    }
#define D(MM, FF)                                                                                                   \
    if (p->M == MM && p->fmt == FF)                                                                                 \
        return s1 ? fast_launch_t<MM, FF, true, 0>(p, a, stream, mode, grid, block, smem, regs)                    \
                  : fast_launch_t<MM, FF, false, 0>(p, a, stream, mode, grid, block, smem, regs);
    D(2, PIRIP_FMT_U8)
    D(2, PIRIP_FMT_S16)
    D(2, PIRIP_FMT_F32)
    D(4, PIRIP_FMT_U8)
    D(4, PIRIP_FMT_S16)
    D(4, PIRIP_FMT_F32)
#undef D
    return cudaErrorInvalidValue;
}

void fsk_fast_plan_destroy(FskFastPlan *p) {
    if (!p) return;
    cudaFree(p->d_hannS);
    cudaFree(p->d_ideal);
    cudaFree(p->d_corr);
    cudaFree(p->d_osc);
    cudaFree(p->d_flags);
    cudaFree(p->d_list);
    cudaFree(p->d_count);
    cudaFree(p->d_mbox);
    cudaFree(p->d_ckpt);
    cudaFree(p->d_afin);
    cudaFree(p->d_replays);
    cudaFree(p->d_trace);
    delete p;
}

FskFastPlan *fsk_fast_plan_create(const pirip_fsk_cfg &c, const FskDevCfg &d, const float *h_hann, const float2 *h_phi_ft, int device,
                                  std::string &why) {
    FskFastPlan *p = new FskFastPlan();
    p->M = c.M;
    p->fmt = c.fmt;
    p->device = device;
    FastCfg &f = p->cfg;
    memset(&f, 0, sizeof f);
    f.Fs = d.Fs; f.Rs = d.Rs; f.P = d.P; f.Nsym = d.Nsym; f.Ts = d.Ts; f.N = d.N; f.Nmem = d.Nmem; f.Ndft = d.Ndft;
    f.Nbits = d.Nbits; f.nint = d.nint; f.step = d.Ts / d.P;
    f.st = d.st; f.en = d.en; f.nb = d.en - d.st; f.f_zero = d.f_zero; f.est_type = d.est_type;
    f.tone_spacing = d.tone_spacing; f.f1_tx = d.f1_tx; f.fixed_nin = (d.burst_mode || d.lock_nin) ? 1 : 0; f.soft = d.soft;
    f.nmask = d.nmask; f.len_mask = d.len_mask;
    for (int i = 0; i < FSK_MAX_MASK; ++i) f.mask_pos[i] = d.mask_pos[i];
    f.log2_ndft = d.log2_ndft; f.nfac = d.nfac;
    for (int i = 0; i < FSK_MAX_FFT_FACTORS; ++i) {
        f.fac_p[i] = d.fac_p[i];
        f.fac_m[i] = d.fac_m[i];
        f.fac_fstride[i] = d.fac_fstride[i];
    }
    f.nin_max = d.N + d.Ts / 4;
    f.nin_min = d.N - d.Ts / 4;
    f.bps = c.fmt == PIRIP_FMT_U8 ? 2 : (c.fmt == PIRIP_FMT_S16 ? 4 : 8);
    f.fs_over_ndft = d.fs_over_ndft;
    f.in_scale = c.fmt == PIRIP_FMT_U8 ? 0.0078125f : 1.0f;
    f.g_sym = 2e-5f; /* the energies at a given timing: fast mixing (closed-form phasors, FMA) and sliding integrators against the
                        scalar algorithm's recursive oscillator, whose amplitude drifts within a frame.  4e-6 was the first value:
                        22 single wrong bits in 1.9e10 (tools/validate_seeds.py, 1.9 million captures against the exact kernel),
                        hidden until the interpolation term of the guard was tightened; all 22 are flagged at 1e-5 */
    f.g_tim = 4e-7f; /* timing-line sums: absolute error relative to the summed window energies (observed <= 9.2e-8) */
    if (const char *e = getenv("PIRIP_FSK_GUARD_SYM")) f.g_sym = (float)atof(e);
    if (const char *e = getenv("PIRIP_FSK_GUARD_TIM")) f.g_tim = (float)atof(e);
    f.diag = getenv("PIRIP_FSK_FAST_DIAG") ? 1 : 0;
    f.exit_on_flag = 1;
    if (const char *e = getenv("PIRIP_FSK_FAST_NO_EXIT")) f.exit_on_flag = atoi(e) ? 0 : 1;
    if (f.nb <= 0 || f.step < 1 || d.Nsym < 3 || f.nint > T * 256) {
        why = "configuration outside the fast kernel's range";
        delete p;
        return nullptr;
    }
    /* integrator outputs per thread: even (16-byte aligned window starts), half of it odd (bank spread) */
    f.RO = (f.nint + T - 1) / T;
    if (f.RO & 1) f.RO++;
    /* step 1: a thread reads its run with 16-byte loads, eight lanes per shared-memory wavefront.  Runs RO * 8 bytes apart fall
       on distinct 16-byte bank groups only when RO / 2 is odd (RO = 20: every load is a 2-way conflict; 22: none) */
    if (f.step == 1 && ((f.RO >> 1) & 1) == 0 && !getenv("PIRIP_FSK_FAST_RO_EVEN")) f.RO += 2;
    f.tailpad = align_up(2 * d.Ts + d.Ts / 4 + 2, TAILPAD_MIN);
    if (f.tailpad < 16) f.tailpad = 16;

    /* shared memory layout */
    f.raw_bytes = align_up(f.nin_max * f.bps + 32, 16);
    f.fdcS = align_up(d.Nmem, 8) + 8;
    while ((f.fdcS & 15) != 8) f.fdcS += 2;
    /* scratch row must hold the FFT exchange / generic FFT blocks, the magnitude row the magnitudes of a round */
    const int row_free = (f.fdcS - f.tailpad) * 8;
    int off = 0;
    off = 16; /* a pair of samples may start one sample before the frame */
    f.off_raw = off; off += f.raw_bytes;
    f.off_fdc = off; off += c.M * f.fdcS * 8;
    f.off_Sf = off; off += align_up(f.nb * 4, 16);
    f.off_hann = off; off += align_up(d.Ndft * 4, 16);
    f.off_filt = off; off += align_up(c.M * d.Nsym * 4, 16);
    f.off_ideal = off; off += align_up((d.P + 1) * 8, 16);
    f.off_misc = off; off += align_up((int)sizeof(Misc), 16);
    f.smem_bytes = off;
    f.group_stride = align_up(off, 128);
    const int numffts_max = f.nin_max / (d.Ndft / 2) - 1;
    f.fft_merged = (d.log2_ndft == 8 && numffts_max * f.nb * 4 <= row_free && !getenv("PIRIP_FSK_FAST_NO_MERGE")) ? 1 : 0;
    if (d.log2_ndft == 8) {
        int per = row_free / (272 * 8);
        const int permag = row_free / (f.nb * 4);
        if (permag < per) per = permag;
        if (per > T / 16) per = T / 16;
        if (per < 1) {
            why = "fast kernel: scratch rows too small";
            delete p;
            return nullptr;
        }
        f.fft_blocks_per_round = per;
        f.prune = (f.st >= 128 && f.en <= 128 + 32) ? 1 : 0;
        if (getenv("PIRIP_FSK_FAST_NO_PRUNE")) f.prune = 0;
    } else {
        int per = row_free / (d.Ndft * 8);
        const int permag = f.nb > 0 ? row_free / (f.nb * 4) : per;
        if (permag < per) per = permag;
        if (per < 1 && numffts_max > 0) {
            why = "fast kernel: an FFT block does not fit the scratch row";
            delete p;
            return nullptr;
        }
        f.fft_blocks_per_round = per < 1 ? 1 : per;
    }
    if (f.smem_bytes > 227 * 1024) {
        why = "fast kernel: more than 227 KB of shared memory";
        delete p;
        return nullptr;
    }

    /* tables */
    std::vector<float> hs((size_t)d.Ndft);
    for (int i = 0; i < d.Ndft; ++i) hs[(size_t)i] = h_hann[i] * f.in_scale; /* exact: power-of-two scale */
    std::vector<OscEntry> osc((size_t)d.Ndft * 4);
    const double two_pi = 6.283185307179586476925286766559;
    for (int b = 0; b < d.Ndft; ++b)
        for (int m = 0; m < 4; ++m) {
            float fest;
            if (d.est_type) {
                const float foff = (float)((b - d.Ndft / 2) * d.Fs / d.Ndft);
                fest = foff + (float)(m * d.tone_spacing);
            } else
                fest = (float)(b - d.Ndft / 2) * d.fs_over_ndft;
            /* [U fsk.c::fsk_demod_core] dphi_m = comp_exp_j(2*M_PI*(f_est/(float)Fs)) */
            const float th = (float)(2 * M_PI * ((fest) / (float)(d.Fs)));
            const float cf = em_cosf(th), sf = em_sinf(th);
            const double ang = atan2((double)sf, (double)cf);
            const double lng = 0.5 * log((double)cf * cf + (double)sf * sf);
            OscEntry e;
            e.dphi = make_float2(cf, sf);
            const double turns = ang / two_pi;
            const double tfrac = turns - floor(turns);
            e.step32 = (unsigned)(unsigned long long)llround(tfrac * 4294967296.0);
            const double n2 = 2.0 * T;
            const double gs = exp(lng * n2);
            e.dstride = make_float2((float)(gs * cos(ang * n2)), (float)(gs * sin(ang * n2)));
            e.lng = (float)lng;
            e.f = fest;
            e.pad = 0.0f;
            osc[(size_t)b * 4 + m] = e;
        }
    std::vector<float2> ideal((size_t)d.P), corr((size_t)T);
    for (int k = 0; k < d.P; ++k) ideal[(size_t)k] = make_float2((float)cos(two_pi * k / d.P), (float)sin(two_pi * k / d.P));
    (void)cmulf;
    for (int r = 0; r < T; ++r) {
        double sr = 0, si = 0;
        int cnt = 0;
        for (int i = r * f.RO; i < (r + 1) * f.RO && i < f.nint; ++i, ++cnt) {
            const double ir = cos(two_pi * (i % d.P) / d.P), ii = sin(two_pi * (i % d.P) / d.P);
            /* phi_ft[i] * conj(ideal) */
            sr += (double)h_phi_ft[i].x * ir + (double)h_phi_ft[i].y * ii;
            si += (double)h_phi_ft[i].y * ir - (double)h_phi_ft[i].x * ii;
        }
        corr[(size_t)r] = cnt ? make_float2((float)(sr / cnt), (float)(si / cnt)) : make_float2(1.0f, 0.0f);
    }
    if (d.log2_ndft == 8) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        for (int k = 1; k < 4; ++k)
            for (int j = 1; j < 4; ++j) {
                const double phase = -2 * pi * (16 * k * j) / 256;
                f.tw_l2[3 * (k - 1) + (j - 1)] = make_float2((float)cos(phase), (float)sin(phase));
            }
    }
    bool ok = pirip::cuda_ok(cudaSetDevice(device), "cudaSetDevice") &&
              pirip::cuda_ok(cudaMalloc(&p->d_hannS, sizeof(float) * hs.size()), "cudaMalloc hannS") &&
              pirip::cuda_ok(cudaMalloc(&p->d_osc, sizeof(OscEntry) * osc.size()), "cudaMalloc osc") &&
              pirip::cuda_ok(cudaMalloc(&p->d_ideal, sizeof(float2) * ideal.size()), "cudaMalloc ideal") &&
              pirip::cuda_ok(cudaMalloc(&p->d_corr, sizeof(float2) * corr.size()), "cudaMalloc corr") &&
              pirip::cuda_ok(cudaMalloc(&p->d_count, sizeof(int) * 64), "cudaMalloc flag count") &&
              pirip::cuda_ok(cudaMemset(p->d_count, 0, sizeof(int) * 64), "memset flag count") &&
              pirip::cuda_ok(cudaMemcpy(p->d_hannS, hs.data(), sizeof(float) * hs.size(), cudaMemcpyHostToDevice), "copy hannS") &&
              pirip::cuda_ok(cudaMemcpy(p->d_osc, osc.data(), sizeof(OscEntry) * osc.size(), cudaMemcpyHostToDevice), "copy osc") &&
              pirip::cuda_ok(cudaMemcpy(p->d_ideal, ideal.data(), sizeof(float2) * ideal.size(), cudaMemcpyHostToDevice), "copy ideal") &&
              pirip::cuda_ok(cudaMemcpy(p->d_corr, corr.data(), sizeof(float2) * corr.size(), cudaMemcpyHostToDevice), "copy corr");
    f.hannS = p->d_hannS;
    f.twiddle = d.twiddle;
    f.osc = p->d_osc;
    f.ftideal = p->d_ideal;
    f.ftcorr = p->d_corr;
    f.phi_ft = d.phi_ft;
    p->ckpt_per_frame = ck_chunks(f.nin_max) + 1;
    p->chain = getenv("PIRIP_FSK_NO_CHAIN") == nullptr;
    ok = ok && pirip::cuda_ok(cudaMalloc(&p->d_replays, sizeof(int) * 64), "cudaMalloc replays") &&
         pirip::cuda_ok(cudaMemset(p->d_replays, 0, sizeof(int) * 64), "memset replays");
    FastArgs dummy;
    memset(&dummy, 0, sizeof dummy);
    ok = ok && pirip::cuda_ok(fast_dispatch(p, dummy, nullptr, 1), "cudaFuncSetAttribute (fast kernel)");
    int regs = 128;
    ok = ok && pirip::cuda_ok(fast_dispatch(p, dummy, nullptr, 2, 0, 0, 0, &regs), "cudaFuncGetAttributes (fast kernel)");
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) p->num_sms = sms;
        int g = (227 * 1024) / f.group_stride;
        const int greg = 65536 / (T * (regs > 0 ? regs : 128));
        if (greg < g) g = greg;
        const int gcap = c.M == 2 ? 9 : 5; /* __launch_bounds__ of the kernel */
        if (g > gcap) g = gcap;
        if (const char *ev = getenv("PIRIP_FSK_FAST_G"))
            if (atoi(ev) >= 1 && atoi(ev) < g) g = atoi(ev);
        p->gmax = g;
        if (g < 1) {
            pirip::set_error("fast kernel: a capture group does not fit an SM");
            ok = false;
        }
    }
    if (!ok) {
        why = pirip::last_error();
        fsk_fast_plan_destroy(p);
        return nullptr;
    }
    char buf[256];
    snprintf(buf, sizeof buf,
             "fsk_fast_kernel<M=%d>: one CTA per SM, up to %d capture groups of %d threads (%d B smem each), exact-oscillator chain CTAs %s; guards sym %.1e "
             "tim %.1e; decisions inside a margin are replayed exactly in the kernel",
             c.M, p->gmax, T, f.group_stride, p->chain ? "on" : "off", (double)f.g_sym, (double)f.g_tim);
    p->desc = buf;
    return p;
}

const char *fsk_fast_plan_desc(const FskFastPlan *p) { return p ? p->desc.c_str() : ""; }

/* device buffers for the flags of a launch: [n] flags, [n] list, count */
cudaError_t fsk_fast_flag_buffers(FskFastPlan *p, size_t n, int **flags, int **list, int **count) {
    if (n > p->cap_flags) {
        cudaFree(p->d_flags);
        cudaFree(p->d_list);
        p->d_flags = p->d_list = nullptr;
        p->cap_flags = 0;
        cudaError_t e = cudaMalloc(&p->d_flags, sizeof(int) * n);
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&p->d_list, sizeof(int) * n);
        if (e != cudaSuccess) return e;
        p->cap_flags = n;
    }
    *flags = p->d_flags;
    *list = p->d_list;
    *count = p->d_count;
    return cudaSuccess;
}

/* replays of the plan so far: [0] frames replayed exactly, [1] replays that could not be done (side chain late / absent) */
cudaError_t fsk_fast_replay_counts(FskFastPlan *p, int out[64]) { return cudaMemcpy(out, p->d_replays, sizeof(int) * 64, cudaMemcpyDeviceToHost); }

static cudaError_t chain_buffers(FskFastPlan *p, size_t n, size_t max_frames) {
    const size_t need_mbox = n * (max_frames + 1);
    if (need_mbox > p->cap_mbox) {
        cudaFree(p->d_mbox);
        p->d_mbox = nullptr;
        p->cap_mbox = 0;
        cudaError_t e = cudaMalloc(&p->d_mbox, sizeof(u64) * need_mbox);
        if (e != cudaSuccess) return e;
        p->cap_mbox = need_mbox;
    }
    if (n > p->cap_chain) {
        cudaFree(p->d_ckpt);
        cudaFree(p->d_afin);
        p->d_ckpt = nullptr;
        p->d_afin = nullptr;
        p->cap_chain = 0;
        cudaError_t e = cudaMalloc(&p->d_ckpt, sizeof(float4) * n * 2 * p->ckpt_per_frame * 4);
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&p->d_afin, sizeof(int) * n);
        if (e != cudaSuccess) return e;
        p->cap_chain = n;
    }
    return cudaSuccess;
}

/* `off` / `total`: position of this launch's captures inside the buffers of a chunked host-buffer call */
cudaError_t fsk_fast_launch(FskFastPlan *p, const FskBatchArgs &b, int *d_flags, int *d_list, int *d_count, cudaStream_t stream,
                            size_t off, size_t total, size_t slot) {
    /* slot = index base of this launch's captures in a flag list shared by the rounds of one call (0: the list is the launch's own) */
    FastArgs a;
    memset(&a, 0, sizeof a);
    a.iq = b.iq;
    a.stride = b.capture_stride_bytes;
    a.nsamples = b.samples_per_capture;
    a.max_frames = b.max_frames;
    a.n_captures = b.n_captures;
    a.alloc_end = (reinterpret_cast<size_t>(b.iq) + (b.n_captures - 1) * b.capture_stride_bytes + b.samples_per_capture * (size_t)p->cfg.bps) & ~(size_t)15;
    a.bits = b.bits;
    a.llr = b.llr;
    a.rx_filt = b.rx_filt;
    a.stats = b.stats;
    a.nframes = b.nframes;
    a.flags = d_flags;
    a.flag_list = d_list;
    a.flag_count = d_count;
    a.tma_ok = b.tma_ok;
    a.landed = b.landed;
    a.ckpt_per_frame = p->ckpt_per_frame;
    a.cap_index_base = (int)slot;
    if (total < off + b.n_captures) total = off + b.n_captures;
    const size_t n = b.n_captures;
    /* geometry: one CTA per SM.  Demodulating CTAs hold up to gmax capture groups; with fewer captures than groups the
       captures are spread over the SMs first.  Chain CTAs: one lane per (capture, tone) that can be in flight at once,
       about 9 warps per CTA (a warp needs ~60% of a scheduler's issue slots to keep pace: two per scheduler is the most). */
    const int sms = p->num_sms;
    int chain_ctas = 0, wpc = 0;
    if (p->chain) {
        size_t inflight = (size_t)sms * p->gmax;
        if (inflight > n) inflight = n;
        const size_t warps = (inflight * p->M + 31) / 32;
        wpc = 9;
        if (const char *ev = getenv("PIRIP_FSK_CHAIN_WPC"))
            if (atoi(ev) >= 1 && atoi(ev) <= 18) wpc = atoi(ev);
        if ((size_t)wpc > warps) wpc = (int)warps;
        if (wpc > 2 * p->gmax) wpc = 2 * p->gmax; /* the CTA has 2 * G warps */
        chain_ctas = (int)((warps + wpc - 1) / wpc);
        if (chain_ctas > sms / 4) chain_ctas = sms / 4; /* a longer list is worked off lane by lane */
    }
    int n_a = sms - chain_ctas;
    if ((size_t)n_a > n) n_a = (int)n;
    int G = (int)((n + n_a - 1) / n_a);
    if (G > p->gmax) G = p->gmax;
    int block_groups = G;
    if (chain_ctas && 2 * block_groups < wpc) block_groups = (wpc + 1) / 2;
    a.G = G;
    a.chain_ctas = chain_ctas;
    a.chain_wpc = wpc;
    cudaError_t e;
    if (chain_ctas) {
        if ((e = chain_buffers(p, total, b.max_frames)) != cudaSuccess) return e;
        a.mbox = p->d_mbox + off * (b.max_frames + 1);
        a.ckpt = p->d_ckpt + off * (size_t)(2 * p->ckpt_per_frame * 4);
        a.afin = p->d_afin + off;
        a.replays = p->d_replays;
        a.chain_diag = getenv("PIRIP_FSK_CHAIN_DIAG") ? 1 : 0;
        if (const char *ev = getenv("PIRIP_FSK_CHAIN_CPI")) a.chain_cpi = atoi(ev) > 0 && atoi(ev) <= 64 ? atoi(ev) : 0;
        if ((e = cudaMemsetAsync(a.mbox, 0, sizeof(u64) * n * (b.max_frames + 1), stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(a.ckpt, 0, sizeof(float4) * n * 2 * p->ckpt_per_frame * 4, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(a.afin, 0, sizeof(int) * n, stream)) != cudaSuccess) return e;
    }
    /* more than half an SM's shared memory per CTA: never two CTAs on one SM, so every CTA of the grid is resident at once
       (the chain lanes and the replays wait for each other across CTAs) */
    size_t smem = (size_t)block_groups * p->cfg.group_stride;
    if (smem < 116 * 1024) smem = 116 * 1024;
    if (getenv("PIRIP_FSK_FAST_TRACE")) {
        /* diagnostics: cycles between the phase boundaries of frames 20..51 of the first group, averaged, to stderr (synchronous) */
        if (!p->d_trace && (e = cudaMalloc(&p->d_trace, sizeof(long long) * 32 * 16)) != cudaSuccess) return e;
        cudaMemsetAsync(p->d_trace, 0, sizeof(long long) * 32 * 16, stream);
        a.trace = p->d_trace;
        e = fast_dispatch(p, a, stream, 0, (unsigned)(n_a + chain_ctas), (unsigned)(block_groups * T), smem);
        if (e != cudaSuccess) return e;
        long long h[32 * 16];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, p->d_trace, sizeof h, cudaMemcpyDeviceToHost);
        static const char *names[12] = { "", "raw wait", "FFT", "IIR", "peak / mask", "osc set-up", "mix", "integrate", "timing (tid 0)", "decide",
                                         "replay / stats / record", "LLR + tail copy" };
        double sum[12] = { 0 };
        int nfr = 0;
        for (int f = 0; f < 32; ++f) {
            if (!h[f * 16 + 11] || !h[f * 16]) continue;
            ++nfr;
            for (int i = 1; i < 12; ++i) sum[i] += (double)(h[f * 16 + i] - h[f * 16 + i - 1]);
        }
        if (nfr) {
            double tot = 0;
            for (int i = 1; i < 12; ++i) tot += sum[i] / nfr;
            double w0 = 0;
            for (int f = 0; f < 32; ++f)
                if (h[f * 16 + 11] && h[f * 16] && h[f * 16 + 12]) w0 += (double)(h[f * 16 + 12] - h[f * 16 + 3]) / nfr;
            fprintf(stderr, "fast kernel trace (%d frames of group 0, G = %d): %.0f cycles per frame; peak pick (warp 0) alone %.0f\n", nfr, G, tot, w0);
            for (int i = 1; i < 12; ++i) fprintf(stderr, "  %-26s %7.0f cycles  %5.1f %%\n", names[i], sum[i] / nfr, 100.0 * sum[i] / nfr / tot);
        }
        return cudaSuccess;
    }
    return fast_dispatch(p, a, stream, 0, (unsigned)(n_a + chain_ctas), (unsigned)(block_groups * T), smem);
}

int fsk_fast_uses_chain(const FskFastPlan *p) { return p && p->chain ? 1 : 0; }

/* captures one launch keeps resident (every group one capture): larger batches are cut into launches of this size by the
   callers - a group that takes a second capture in the same launch left the next launch waiting on its bulk-copy barrier
   (seen with 1400 and 2600 captures on B200), and a round of launches costs no more than the in-kernel loop did */
size_t fsk_fast_round_captures(const FskFastPlan *p) {
    /* the largest n that fsk_fast_launch's geometry holds with one capture per group: (SMs - chain CTAs for n) * gmax >= n */
    const int sms = p->num_sms;
    for (size_t n = (size_t)sms * p->gmax; n > 1; --n) {
        int chain_ctas = 0;
        if (p->chain) {
            const size_t warps = (n * p->M + 31) / 32;
            int wpc = 9;
            if ((size_t)wpc > warps) wpc = (int)warps;
            if (wpc > 2 * p->gmax) wpc = 2 * p->gmax;
            chain_ctas = (int)((warps + wpc - 1) / wpc);
            if (chain_ctas > sms / 4) chain_ctas = sms / 4;
        }
        if ((size_t)(sms - chain_ctas) * (size_t)p->gmax >= n) return n;
    }
    return 1;
}
