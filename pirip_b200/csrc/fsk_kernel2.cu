/*
 * fsk_kernel2.cu — the batched FSK demodulator: one persistent CTA per SM, up to `gmax`
 * captures demodulated in lockstep, the whole `fsk_demod` main loop on the device
 * ([U codec2 fsk_demod.c] loop, [U fsk.c::fsk_demod_freq_est / fsk_demod_core],
 * [U mpdecode_core.c::fsk_rx_filt_to_llrs]; pirip call sites [R README.md:105,109],
 * [R test/loopback_rtl_sdr.sh:16]).
 *
 * Same NUMERICS CONTRACT as fsk_kernel.cu: identical IEEE-754 operations in the same order as
 * the scalar algorithm (no FMA contraction, exact_math.h transcendentals), so bits, nin
 * decisions and every float are bit-identical to the CPU evaluation.  What changed is where
 * the data lives and who computes what:
 *
 *  - All captures of a launch are resident at once (1250 captures = 8..9 per SM on 148 SMs):
 *    no wave quantisation.  That needs < 25 KB of shared memory per capture, so the kernel
 *    never holds a whole frame of baseband samples.  The per-tone f_dc memory is a RING of
 *    4 chunks indexed by absolute stream position (no per-frame slide), the integrator
 *    outputs are consumed on the fly, and the two outputs per symbol that the decision needs
 *    are RE-SUMMED after the timing estimate from the raw u8 samples (still staged), the
 *    oscillator phase checkpoints (every 16th phase kept) and the f_dc tail of the previous
 *    frame — the same float operations in the same order, so the same floats.
 *  - Streaming phase, per 128-sample chunk, synchronised with mbarriers only:
 *      oscillator warp  lane (g, m): serial phase recursion of tone m of capture g (up to
 *                       32 recurrences per instruction), phases into the ring;
 *      worker warps     mix in place (u8 converted on the fly); integrate 16 adjacent windows
 *                       per thread with packed f32x2 adds (every window summed from zero in
 *                       ascending order; 20 16-byte loads per 384 adds), window energies to a
 *                       small ring;
 *      timing warp      lane g: serial timing-line sum of capture g right behind the workers;
 *      FFT warps        warp g: the frequency estimate of the NEXT frame of capture g (its
 *                       position is already known and the TMA copy was issued one frame
 *                       ahead), Hann + radix-4/2 FFT blocks, magnitude IIR held in registers,
 *                       peak pick / mask, dphi.
 *  - Decision phase: lane (g, symbol) re-sums both tones' windows, interpolates, decides;
 *    then the four serial per-frame sums, SNR / EbNo, optional LLRs, stats record.
 *  - The f_dc ring is XOR-swizzled in 16-byte units by the 128-byte row index, so the
 *    16-window integrator tasks of a quarter-warp (one row apart) hit distinct bank groups.
 */
#include <stdio.h>

#include "exact_math.h"
#include "fsk_dev.h"
#include "llr_dev.h"

namespace {

typedef unsigned long long u64;

/* ---- packed f32x2 arithmetic (FADD2 / FMUL2): two independent IEEE fp32 operations ---- */
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    float2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

/* in * conj(phi).  Scalar on purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with
   -fmad=false (measured: 1-ulp differences), so packed f32x2 operations are used in this file ONLY for
   additions whose operands are not products (integrator and window sums). */
__device__ __forceinline__ float2 mixc(float2 x, float2 o) { return make_float2(x.x * o.x + x.y * o.y, x.y * o.x - x.x * o.y); }

/* ---- mbarrier / TMA wrappers (shared::cta) ---- */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

/* FFT scratch swizzle (same as fsk_kernel.cu) */
__device__ __forceinline__ int fswz(int e) { return e ^ (5 * ((e >> 4) & 3)); }

/* u8 -> float exactly as ((float)b - 127.0)/128.0 */
__device__ __forceinline__ float u8f(unsigned b) { return (float)((int)b - 127) * 0.0078125f; }

/* complex sample i of a raw interleaved stream, converted as the fsk_demod CLI does ([U fsk_demod.c]) */
__device__ __forceinline__ float2 raw_sample(const unsigned char *src, int fmt, int i) {
    if (fmt == PIRIP_FMT_U8) {
        const unsigned w = *reinterpret_cast<const unsigned short *>(src + 2 * i);
        return make_float2(u8f(w & 0xff), u8f(w >> 8));
    }
    if (fmt == PIRIP_FMT_S16) {
        const short2 v = *reinterpret_cast<const short2 *>(src + 4 * i);
        return make_float2((float)v.x / 825.0f, (float)v.y / 825.0f); /* FDMDV_SCALE */
    }
    return *reinterpret_cast<const float2 *>(src + 8 * (size_t)i);
}

__device__ __forceinline__ size_t bytes_per_sample(int fmt) { return fmt == PIRIP_FMT_U8 ? 2 : (fmt == PIRIP_FMT_S16 ? 4 : 8); }

/* One warp: Hann-window N = 2^LOGN samples, FFT them in the warp's scratch slot with kiss_fft's
   decimation-in-time factor order 4,4,..,(2); fft-shifted magnitudes of bins lane + 32 q in mg[q]. */
template <int LOGN>
__device__ __forceinline__ void warp_fft_mag(const unsigned char *src, int fmt, int first, float2 *buf, float (&mg)[(1 << LOGN) / 32],
                                             const float2 *tw, const float *hann, const unsigned short *perm, int lane) {
    constexpr int N = 1 << LOGN;
    constexpr int N4 = LOGN / 2;
    constexpr bool HAS2 = (LOGN & 1) != 0;
#pragma unroll
    for (int q = 0; q < N / 32; ++q) {
        const int n = lane + (q << 5);
        const float h = hann[n];
        const float2 v = raw_sample(src, fmt, first + n);
        buf[perm[n]] = make_float2(h * v.x, h * v.y);
    }
    __syncwarp();
    if (HAS2) {
        const float2 w0 = tw[0];
#pragma unroll
        for (int q = 0; q < N / 64; ++q) {
            const int b = lane + (q << 5);
            const int i0 = fswz(2 * b), i1 = fswz(2 * b + 1);
            const float2 t = cmul(buf[i1], w0);
            const float2 F0 = buf[i0];
            buf[i1] = csub(F0, t);
            buf[i0] = cadd(F0, t);
        }
        __syncwarp();
    }
#pragma unroll
    for (int L = N4 - 1; L >= 0; --L) {
        const int lm = (HAS2 ? 1 : 0) + 2 * (N4 - 1 - L);
        const int m = 1 << lm;
        const int fs = N >> (lm + 2);
        constexpr int NB = N / 128 > 0 ? N / 128 : 1;
        constexpr int PB = NB >= 2 ? 2 : 1;
#pragma unroll
        for (int q0 = 0; q0 < NB; q0 += PB) {
            int ix[PB][4];
            float2 F[PB][4], T[PB][3];
            bool ok[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int b = lane + ((q0 + u) << 5);
                ok[u] = N / 4 >= 32 || b < N / 4;
                const int blk = b >> lm, k = b & (m - 1);
                const int base = (blk << (lm + 2)) + k;
                ix[u][0] = fswz(base);
                ix[u][1] = fswz(base + m);
                ix[u][2] = fswz(base + 2 * m);
                ix[u][3] = fswz(base + 3 * m);
                if (ok[u]) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) F[u][r] = buf[ix[u][r]];
                    T[u][0] = tw[k * fs];
                    T[u][1] = tw[2 * k * fs];
                    T[u][2] = tw[3 * k * fs];
                }
            }
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                if (ok[u]) {
                    float2 F0 = F[u][0];
                    const float2 s0 = cmul(F[u][1], T[u][0]);
                    const float2 s1 = cmul(F[u][2], T[u][1]);
                    const float2 s2 = cmul(F[u][3], T[u][2]);
                    const float2 s5 = csub(F0, s1);
                    F0 = cadd(F0, s1);
                    const float2 s3 = cadd(s0, s2);
                    const float2 s4 = csub(s0, s2);
                    buf[ix[u][2]] = csub(F0, s3);
                    buf[ix[u][0]] = cadd(F0, s3);
                    buf[ix[u][1]] = make_float2(s5.x + s4.y, s5.y - s4.x);
                    buf[ix[u][3]] = make_float2(s5.x - s4.y, s5.y + s4.x);
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < N / 32; ++q) {
        const int i = lane + (q << 5);
        const float2 X = buf[fswz((i + (N >> 1)) & (N - 1))];
        mg[q] = sqrtf((X.x * X.x) + (X.y * X.y));
    }
    __syncwarp();
}

/* numffts Hann-windowed half-overlapped FFT blocks of one frame; the single-pole magnitude IIR runs
   over the blocks in order with the spectrum held in registers ([U fsk.c::fsk_demod_freq_est]) */
template <int LOGN>
__device__ __noinline__ void frame_spectrum(const unsigned char *src, int fmt, int numffts, float2 *buf, const float *Sf_in,
                                            float *Sf_out, const float2 *tw, const float *hann, const unsigned short *perm, float tc,
                                            int lane) {
    constexpr int N = 1 << LOGN, Q = N / 32;
    const float omtc = 1.0f - tc;
    float s[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) s[q] = Sf_in[lane + (q << 5)];
    for (int j = 0; j < numffts; ++j) {
        float mg[Q];
        warp_fft_mag<LOGN>(src, fmt, j * (N >> 1), buf, mg, tw, hann, perm, lane);
#pragma unroll
        for (int q = 0; q < Q; ++q) s[q] = (s[q] * omtc) + (mg[q] * tc);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) Sf_out[lane + (q << 5)] = s[q];
    __syncwarp();
}

struct Cap2 {
    float2 *ring;  /* [M][ring + 2] f_dc ring by absolute stream position, swizzled */
    float2 *tail;  /* [M][tail] last f_dc samples of the previous frame */
    float2 *ckpt;  /* [M][nckpt] oscillator phase before sample ckpt_every * c of the frame */
    float *ering;  /* [M][ering] window energies; the decision phase reuses it as scratch */
    float *Sf;     /* [2][Ndft] IIR spectrum by frame parity */
    unsigned char *stage; /* [2][stage_bytes] raw samples of frame f / f + 1 */
    Fsk2CapShared *cs;
};

__device__ __forceinline__ Cap2 cap2(unsigned char *smem, const FskDevCfg &cfg, int g) {
    unsigned char *cb = smem + (size_t)g * cfg.k2.cap_stride;
    Cap2 p;
    p.ring = reinterpret_cast<float2 *>(cb + cfg.k2.off_ring);
    p.tail = reinterpret_cast<float2 *>(cb + cfg.k2.off_tail);
    p.ckpt = reinterpret_cast<float2 *>(cb + cfg.k2.off_ckpt);
    p.ering = reinterpret_cast<float *>(cb + cfg.k2.off_ering);
    p.Sf = reinterpret_cast<float *>(cb + cfg.k2.off_Sf);
    p.stage = cb + cfg.k2.off_stage;
    p.cs = reinterpret_cast<Fsk2CapShared *>(cb + cfg.k2.off_scal);
    return p;
}

/* ring element of absolute stream position a: 16-byte units XOR-swizzled by the 128-byte row index */
__device__ __forceinline__ unsigned rsw(unsigned a, unsigned rmask) {
    const unsigned s = a & rmask;
    return s ^ ((s >> 3) & 14u);
}

/* per-capture view of the frame being demodulated, read once per role */
struct FrameView {
    int active, nin, nold, fast;
    unsigned p0;  /* absolute position of the frame's first new sample (low 32 bits) */
    unsigned B;   /* absolute position of f_dc index 0 of the frame */
};

__device__ __forceinline__ FrameView load_view(const Fsk2CapShared *cs, bool exists) {
    FrameView v;
    v.active = exists && cs->active;
    v.nin = cs->nin;
    v.nold = cs->nold;
    v.p0 = (unsigned)cs->pos;
    v.B = v.p0 - (unsigned)v.nold;
    v.fast = cs->fast;
    return v;
}

/* Issue the TMA bulk copy of raw samples [pos, pos + nin_max) of capture cs->cap into staging buffer b
   (16-byte aligned superset, clipped to the batch allocation) and record where the frame's samples
   will be read from: the staged copy when it covers nin_max samples, global memory otherwise. */
__device__ void stage_issue(const FskDevCfg &cfg, const FskBatchArgs &args, const Cap2 &c, int b, unsigned long long pos, u64 *bar,
                            const unsigned char **src_out) {
    Fsk2CapShared *cs = c.cs;
    const size_t bps = bytes_per_sample(cfg.fmt);
    const size_t base = reinterpret_cast<size_t>(args.iq);
    const size_t a = base + (size_t)cs->cap * args.capture_stride_bytes + pos * bps;
    cs->stg_valid[b] = 0;
    *src_out = reinterpret_cast<const unsigned char *>(a);
    if (!args.tma_ok || pos >= args.samples_per_capture) return;
    const size_t want = (size_t)(cfg.N + cfg.Ts / 4) * bps;
    const size_t lo = a & ~(size_t)15;
    size_t hi = (a + want + 15) & ~(size_t)15;
    const size_t safe = (base + (args.n_captures - 1) * args.capture_stride_bytes + args.samples_per_capture * bps) & ~(size_t)15;
    if (hi > safe) hi = safe;
    if (hi <= lo) return;
    const unsigned bytes = (unsigned)(hi - lo);
    unsigned char *dst = c.stage + (size_t)b * cfg.k2.stage_bytes;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_1d(dst, reinterpret_cast<const void *>(lo), bytes, bar);
    cs->stg_lo[b] = (unsigned long long)lo;
    cs->stg_bytes[b] = bytes;
    cs->stg_valid[b] = 1;
    cs->stg_fills[b] += 1;
    if (a + want <= lo + bytes) *src_out = dst + (a - lo);
}

/* integrator outputs complete once the first `jc` new samples of the frame are mixed (generic path) */
__device__ __forceinline__ int outputs_ready(int nold, int jc, int Ts, int step, int nint) {
    const int have = nold + jc - Ts;
    if (have < 0) return 0;
    const int r = (step == 1 ? have : have / step) + 1;
    return r < nint ? r : nint;
}

/* windows of capture view v whose energies exist once chunk c has been integrated.  Fast path: whole
   16-window tasks; a task is run in the chunk that completes its last window. */
__device__ __forceinline__ int windows_after_chunk(const FrameView &v, int c, int CH, int Ts, int step, int nint) {
    if (c < 0 || !v.active) return 0;
    const int jc = (c + 1) * CH;
    if (jc >= v.nin) return nint;
    if (v.fast) {
        const int num = jc - 15 - Ts + v.nold;
        if (num < 0) return 0;
        const int w = ((num >> 4) + 1) << 4;
        return w < nint ? w : nint;
    }
    return outputs_ready(v.nold, jc, Ts, step, nint);
}

/* 16 adjacent integrator windows of one tone: window k sums ring samples s0 + k .. s0 + k + TS - 1 in
   ascending order from zero; |sum|^2 of window k to dst[k].  s0 even. */
template <int TS>
__device__ __forceinline__ void integrate16(const float2 *ring_m, unsigned s0, unsigned umask, float *dst) {
    u64 acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0ull;
    const unsigned u0 = s0 >> 1;
    constexpr int NU = (16 + TS) / 2;
#pragma unroll
    for (int r = 0; r < NU; ++r) {
        const unsigned u = (u0 + r) & umask;
        const unsigned pu = u ^ ((u >> 3) & 7u);
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(ring_m + 2 * pu);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (2 * r >= k && 2 * r < k + TS) acc[k] = add2(acc[k], v.x);
            if (2 * r + 1 >= k && 2 * r + 1 < k + TS) acc[k] = add2(acc[k], v.y);
        }
    }
    float en[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float2 a = upk(acc[k]);
        en[k] = (a.x * a.x) + (a.y * a.y);
    }
#pragma unroll
    for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4 *>(dst + k) = make_float4(en[k], en[k + 1], en[k + 2], en[k + 3]);
}

/* one window, any Ts / alignment */
__device__ __forceinline__ float window_energy(const float2 *ring_m, unsigned s0, unsigned rmask, int Ts) {
    u64 acc = 0ull;
    for (int j = 0; j < Ts; ++j) acc = add2(acc, *reinterpret_cast<const u64 *>(ring_m + rsw(s0 + j, rmask)));
    const float2 a = upk(acc);
    return (a.x * a.x) + (a.y * a.y);
}

#ifdef FSK2_TIMING
#define TMARK(var)                      \
    do {                                \
        const long long tn_ = clock64(); \
        var += tn_ - tl_;               \
        tl_ = tn_;                      \
    } while (0)
#else
#define TMARK(var)
#endif

template <int M, int NW>
__global__ void __launch_bounds__(NW * 32, 1) fsk2_kernel(const __grid_constant__ FskDevCfg cfg, const __grid_constant__ FskBatchArgs args) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nthreads = NW * 32;
    const int Ndft = cfg.Ndft, nint = cfg.nint, Nsym = cfg.Nsym, P = cfg.P, Ts = cfg.Ts, Nmem = cfg.Nmem, Nbits = cfg.Nbits;
    const int step = Ts / P;
    const int CH = cfg.k2.chunk, R = cfg.k2.ring, ER = cfg.k2.ering, CKE = cfg.k2.ckpt_every, NCK = cfg.k2.nckpt, TAILN = cfg.k2.tail;
    const unsigned rmask = (unsigned)R - 1u, umask = ((unsigned)R >> 1) - 1u;
    const int RT = R + 2; /* tone stride of the ring: 16 bytes of skew between tones */
    const int fmt = cfg.fmt;
    unsigned char *cta = smem + (size_t)cfg.k2.gmax * cfg.k2.cap_stride;
    Fsk2CtaShared *ctl = reinterpret_cast<Fsk2CtaShared *>(cta + cfg.k2.off_ctl);
    float2 *tw = reinterpret_cast<float2 *>(cta + cfg.k2.off_tw);
    float *hann = reinterpret_cast<float *>(cta + cfg.k2.off_hann);
    unsigned short *perm = reinterpret_cast<unsigned short *>(cta + cfg.k2.off_perm);
    float2 *fft_slots = reinterpret_cast<float2 *>(cta + cfg.k2.off_fft);
    const float2 *phi_ft = cfg.phi_ft;
    constexpr int osc_warp = NW - 1, tim_warp = NW - 2;

    for (int i = tid; i < Ndft; i += nthreads) {
        tw[i] = cfg.twiddle[i];
        hann[i] = cfg.hann[i];
        perm[i] = cfg.perm[i];
    }

    const size_t ncap = args.n_captures;
    const int grid = (int)gridDim.x;
    const int ncap_cta = ncap > blockIdx.x ? (int)((ncap - blockIdx.x + grid - 1) / grid) : 0;
#ifdef FSK2_TIMING
    long long t_osc = 0, t_oscw = 0, t_wosc = 0, t_mix = 0, t_wmix = 0, t_int = 0, t_wtim = 0, t_tim = 0, t_timw = 0, t_fft = 0,
              t_fftw = 0, t_stream = 0, t_dec = 0, t_sum = 0, t_llr = 0, t_adv = 0, t_pre = 0;
    long long tl_ = clock64();
#endif

    for (int k0 = 0; k0 < ncap_cta; k0 += cfg.k2.gmax) {
        const int G = ncap_cta - k0 < cfg.k2.gmax ? ncap_cta - k0 : cfg.k2.gmax;
        const int W = NW - 2 - G;     /* worker warps: G .. NW-3 */
        const int nwt = W * 32;
        const int wtid = (warp - G) * 32 + lane;
        const bool is_fft = warp < G, is_worker = warp >= G && warp < NW - 2;
        __syncthreads();
        /* ---- barriers and initial state of the group ---- */
        if (tid == 0) {
            for (int i = 0; i < FSK2_MAX_CHUNKS; ++i) {
                mbar_init(&ctl->bar_osc[i], 1);
                mbar_init(&ctl->bar_mix[i], W);
                mbar_init(&ctl->bar_int[i], W);
                mbar_init(&ctl->bar_tim[i], 1);
            }
            for (int g = 0; g < FSK2_MAX_G; ++g) {
                mbar_init(&ctl->bar_in[g][0], 1);
                mbar_init(&ctl->bar_in[g][1], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            ctl->G = G;
        }
        for (int g = 0; g < G; ++g) {
            const Cap2 c = cap2(smem, cfg, g);
            for (int i = tid; i < M * RT; i += nthreads) c.ring[i] = make_float2(0.0f, 0.0f);
            for (int i = tid; i < M * TAILN; i += nthreads) c.tail[i] = make_float2(0.0f, 0.0f);
            for (int i = tid; i < 2 * Ndft; i += nthreads) c.Sf[i] = 0.0f;
            if (tid == 0) {
                Fsk2CapShared *cs = c.cs;
                for (int m = 0; m < 4; ++m) {
                    cs->phi_c[m] = make_float2(1.0f, 0.0f);
                    cs->f_est[0][m] = cs->f_est[1][m] = 0.0f;
                    cs->f2_est[0][m] = cs->f2_est[1][m] = 0.0f;
                    cs->dphi[0][m] = cs->dphi[1][m] = make_float2(1.0f, 0.0f);
                }
                cs->norm_rx_timing = cs->ppm = cs->snr_est = cs->EbNodB = 0.0f;
                cs->SNRest = cs->v_est = cs->rx_sig_pow = cs->rx_nse_pow = cs->foff = cs->rx_timing = 0.0f;
                cs->nin = cfg.N;
                cs->nin_next = cfg.N;
                cs->nold = Nmem - cfg.N;
                cs->pos = 0;
                cs->frame = 0;
                cs->nan_flag = 0;
                cs->cap = (int)(blockIdx.x + (size_t)(k0 + g) * grid);
                cs->stg_valid[0] = cs->stg_valid[1] = 0;
                cs->stg_fills[0] = cs->stg_fills[1] = 0;
                cs->fast = 0;
                /* frame 0 exists for some nin iff its shortest form fits */
                cs->fon = args.max_frames > 0 && (unsigned long long)(cfg.N - Ts / 4) <= args.samples_per_capture;
                cs->active = args.max_frames > 0 && (unsigned long long)cfg.N <= args.samples_per_capture;
            }
        }
        __syncthreads();
        const unsigned char *src_cur = nullptr, *src_nxt = nullptr; /* meaningful in threads tid < G */
        if (tid < G) {
            const Cap2 c = cap2(smem, cfg, tid);
            if (c.cs->fon) stage_issue(cfg, args, c, 0, 0, &ctl->bar_in[tid][0], &src_cur);
        }
        /* the raw-sample pointers of the two staging buffers live in shared memory for all roles */
        const unsigned char **srcp = reinterpret_cast<const unsigned char **>(ctl + 1); /* [gmax][2], after the control block */
        if (tid < G) srcp[2 * tid + 0] = src_cur;
        __syncthreads();

        /* ---- prologue: frequency estimate of frame 0 ---- */
        auto freq_estimate = [&](int g, int fb) {
            /* FFT warp of capture g: estimate for the frame whose parity is fb, from staging buffer fb */
            const Cap2 c = cap2(smem, cfg, g);
            Fsk2CapShared *cs = c.cs;
            if (cs->stg_valid[fb]) mbar_wait(&ctl->bar_in[g][fb], (unsigned)(cs->stg_fills[fb] - 1) & 1u);
            const unsigned char *src = srcp[2 * g + fb];
            float2 *buf = fft_slots + (size_t)g * Ndft;
            const float *Sf_in = c.Sf + (size_t)(fb ^ 1) * Ndft;
            float *Sf_out = c.Sf + (size_t)fb * Ndft;
            switch (cfg.log2_ndft) {
                case 6: frame_spectrum<6>(src, fmt, cfg.k2.numffts, buf, Sf_in, Sf_out, tw, hann, perm, cfg.tc, lane); break;
                case 7: frame_spectrum<7>(src, fmt, cfg.k2.numffts, buf, Sf_in, Sf_out, tw, hann, perm, cfg.tc, lane); break;
                case 8: frame_spectrum<8>(src, fmt, cfg.k2.numffts, buf, Sf_in, Sf_out, tw, hann, perm, cfg.tc, lane); break;
                case 9: frame_spectrum<9>(src, fmt, cfg.k2.numffts, buf, Sf_in, Sf_out, tw, hann, perm, cfg.tc, lane); break;
                default: frame_spectrum<10>(src, fmt, cfg.k2.numffts, buf, Sf_in, Sf_out, tw, hann, perm, cfg.tc, lane); break;
            }
            /* estimator 1: M rounds of argmax (first maximum wins) + blanking, on a scratch copy */
            float *tmpS = reinterpret_cast<float *>(buf);
            for (int i = lane; i < Ndft; i += 32) tmpS[i] = Sf_out[i];
            __syncwarp();
            for (int r = 0; r < M; ++r) {
                float best = 0.0f;
                int ibest = 0;
                for (int j = cfg.st + lane; j < cfg.en; j += 32) {
                    const float v = tmpS[j];
                    if (v > best) {
                        best = v;
                        ibest = j;
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, ibest, off);
                    if (ov > best || (ov == best && oi < ibest)) {
                        best = ov;
                        ibest = oi;
                    }
                }
                int f_min = ibest - cfg.f_zero;
                f_min = f_min < 0 ? 0 : f_min;
                int f_max = ibest + cfg.f_zero;
                f_max = f_max > Ndft ? Ndft : f_max;
                for (int j = f_min + lane; j < f_max; j += 32) tmpS[j] = 0.0f;
                if (lane == 0) cs->freqi[r] = ibest - (Ndft >> 1);
                __syncwarp();
            }
            if (lane == 0) {
                int fi[4];
                for (int m = 0; m < M; ++m) fi[m] = cs->freqi[m];
                for (int a = 1; a < M; ++a) {
                    const int key = fi[a];
                    int b = a - 1;
                    while (b >= 0 && fi[b] > key) {
                        fi[b + 1] = fi[b];
                        --b;
                    }
                    fi[b + 1] = key;
                }
                for (int m = 0; m < M; ++m) cs->f_est[fb][m] = (float)fi[m] * cfg.fs_over_ndft;
            }
            /* estimator 2: slide the tone mask over Sf, first maximum wins */
            {
                float best = 0.0f;
                int bbest = 0x7fffffff;
                for (int b = cfg.st + lane; b < cfg.en - cfg.len_mask; b += 32) {
                    float corr = 0.0f;
                    for (int q = 0; q < cfg.nmask; ++q) corr += Sf_out[b + cfg.mask_pos[q]];
                    if (corr > best) {
                        best = corr;
                        bbest = b;
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bbest, off);
                    if (ov > best || (ov == best && oi < bbest)) {
                        best = ov;
                        bbest = oi;
                    }
                }
                if (lane == 0) {
                    const int b_max = (best > 0.0f) ? bbest : cfg.st;
                    const float foff = (float)((b_max - (Ndft >> 1)) * cfg.Fs / Ndft);
                    for (int m = 0; m < M; ++m) cs->f2_est[fb][m] = foff + (float)(m * cfg.tone_spacing);
                }
            }
            __syncwarp();
            if (lane < M) {
                const float f = cfg.est_type ? cs->f2_est[fb][lane] : cs->f_est[fb][lane];
                const float arg = (float)(EM_TWO_PI * (double)(f / (float)cfg.Fs));
                double sn, csn;
                em_sincos((double)arg, &sn, &csn);
                cs->dphi[fb][lane] = make_float2((float)csn, (float)sn);
            }
            __syncwarp();
        };
        if (is_fft && cap2(smem, cfg, warp).cs->fon) freq_estimate(warp, 0);

        /* frame bookkeeping by one thread: which captures have a whole frame left */
        auto bookkeeping = [&]() {
            int any = 0, mx = 0, gen = 0;
            for (int g = 0; g < G; ++g) {
                Fsk2CapShared *cs = cap2(smem, cfg, g).cs;
                int act = cs->active;
                if (act && ((size_t)cs->frame >= args.max_frames || cs->pos + (unsigned long long)cs->nin > args.samples_per_capture))
                    act = 0; /* short read ends the loop */
                cs->active = act;
                cs->nold = Nmem - cs->nin;
                /* the 16-window integrator needs 16-byte aligned window starts: f_dc index 0 at an even position */
                cs->fast = cfg.k2.fast_ts && step == 1 && ((((unsigned)cs->pos - (unsigned)cs->nold) & 1u) == 0);
                if (act) {
                    any = 1;
                    mx = cs->nin > mx ? cs->nin : mx;
                    if (!cs->fast) gen = 1;
                }
            }
            ctl->any_active = any;
            ctl->max_nin = mx;
            ctl->any_generic = gen;
        };
        __syncthreads();
        if (tid == 0) bookkeeping();
        __syncthreads();
        TMARK(t_pre);

        unsigned par = 0; /* phase parity of the per-chunk mbarriers, bit c */
        for (int f = 0;; ++f) {
            if (!ctl->any_active) break;
            const int fb = f & 1;
            const int max_nin = ctl->max_nin;
            const int nchunks = (max_nin + CH - 1) / CH;
            /* ---- visibility of frame f's staged samples; plan and fetch frame f + 1 ---- */
            for (int g = 0; g < G; ++g) {
                Fsk2CapShared *cs = cap2(smem, cfg, g).cs;
                if (cs->active && cs->stg_valid[fb]) mbar_wait(&ctl->bar_in[g][fb], (unsigned)(cs->stg_fills[fb] - 1) & 1u);
            }
            if (tid < G) {
                const Cap2 c = cap2(smem, cfg, tid);
                Fsk2CapShared *cs = c.cs;
                cs->fon = 0;
                if (cs->active) {
                    const unsigned long long npos = cs->pos + (unsigned long long)cs->nin;
                    if ((size_t)f + 1 < args.max_frames && npos + (unsigned long long)(cfg.N - Ts / 4) <= args.samples_per_capture) {
                        const unsigned char *s = nullptr;
                        stage_issue(cfg, args, c, fb ^ 1, npos, &ctl->bar_in[tid][fb ^ 1], &s);
                        srcp[2 * tid + (fb ^ 1)] = s;
                        cs->fon = 1;
                    }
                }
            }
            __syncthreads();
            TMARK(t_adv);

            /* ================= streaming phase ================= */
            if (warp == osc_warp) {
                /* A8: oscillator recursion, lane (g, m) */
                const int g = lane / M, m = lane - g * M;
                const bool in_range = lane < G * M;
                const Cap2 c = cap2(smem, cfg, in_range ? g : 0);
                const bool on = in_range && c.cs->active;
                const int my_nin = on ? c.cs->nin : 0;
                const float2 d = on ? c.cs->dphi[fb][m] : make_float2(1.0f, 0.0f);
                float2 ph = on ? c.cs->phi_c[m] : make_float2(1.0f, 0.0f);
                float2 *ring_m = c.ring + m * RT;
                float2 *ck_m = c.ckpt + m * NCK;
                const unsigned p0 = (unsigned)c.cs->pos;
                if (on) ck_m[0] = ph;
                int lo = on ? my_nin : 0x7fffffff;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const int o = __shfl_xor_sync(0xffffffffu, lo, off);
                    lo = o < lo ? o : lo;
                }
                for (int ck = 0; ck < nchunks; ++ck) {
                    if (ck >= 3) {
                        TMARK(t_osc);
                        mbar_wait(&ctl->bar_int[ck - 3], (par >> (ck - 3)) & 1u);
                        TMARK(t_oscw);
                    }
                    const int j0 = ck * CH;
                    const int jend = j0 + CH < max_nin ? j0 + CH : max_nin;
                    if (jend <= lo && (CKE & 7) == 0) {
                        if (on) {
                            for (int j = j0; j < jend; j += 8) {
                                const int nb = jend - j < 8 ? jend - j : 8;
                                if (nb == 8) {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        ph = cmul(ph, d);
                                        ring_m[rsw(p0 + (unsigned)(j + q), rmask)] = ph;
                                    }
                                    if (((j + 8) & (CKE - 1)) == 0) ck_m[(j + 8) / CKE] = ph;
                                } else {
                                    for (int q = 0; q < nb; ++q) {
                                        ph = cmul(ph, d);
                                        ring_m[rsw(p0 + (unsigned)(j + q), rmask)] = ph;
                                        if (((j + q + 1) & (CKE - 1)) == 0) ck_m[(j + q + 1) / CKE] = ph;
                                    }
                                }
                            }
                        }
                    } else {
                        for (int j = j0; j < jend; ++j) {
                            if (j < my_nin) {
                                ph = cmul(ph, d);
                                ring_m[rsw(p0 + (unsigned)j, rmask)] = ph;
                                if (((j + 1) & (CKE - 1)) == 0) ck_m[(j + 1) / CKE] = ph;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctl->bar_osc[ck]);
                }
                if (on) {
                    const float av = sqrtf((ph.x * ph.x) + (ph.y * ph.y));
                    c.cs->phi_c[m] = make_float2(ph.x / av, ph.y / av);
                }
                TMARK(t_osc);
            } else if (warp == tim_warp) {
                /* A10: serial timing-line sum, lane g */
                const int g = lane;
                const Cap2 c = cap2(smem, cfg, g < G ? g : 0);
                Fsk2CapShared *cs = c.cs;
                const FrameView v = load_view(cs, g < G);
                const bool on = v.active;
                float tre = 0.0f, tim = 0.0f;
                int done = 0;
                const float *er = c.ering;
                for (int ck = 0; ck < nchunks; ++ck) {
                    TMARK(t_tim);
                    mbar_wait(&ctl->bar_int[ck], (par >> ck) & 1u);
                    TMARK(t_timw);
                    const int avail = on ? windows_after_chunk(v, ck, CH, Ts, step, nint) : 0;
                    /* four windows at a time while aligned: 16-byte energy loads, the products are off the chain */
                    while (done < avail) {
                        if ((done & 3) == 0 && done + 4 <= avail) {
                            const int e = done & (ER - 1);
                            float ft[4];
                            {
                                const float4 e0 = *reinterpret_cast<const float4 *>(er + e);
                                ft[0] = e0.x, ft[1] = e0.y, ft[2] = e0.z, ft[3] = e0.w;
                            }
#pragma unroll
                            for (int m = 1; m < M; ++m) {
                                const float4 em = *reinterpret_cast<const float4 *>(er + m * ER + e);
                                ft[0] += em.x, ft[1] += em.y, ft[2] += em.z, ft[3] += em.w;
                            }
                            const float4 pa = __ldg(reinterpret_cast<const float4 *>(phi_ft + done));
                            const float4 pb = __ldg(reinterpret_cast<const float4 *>(phi_ft + done + 2));
                            const float r0 = ft[0] * pa.x, i0 = ft[0] * pa.y, r1 = ft[1] * pa.z, i1 = ft[1] * pa.w;
                            const float r2 = ft[2] * pb.x, i2 = ft[2] * pb.y, r3 = ft[3] * pb.z, i3 = ft[3] * pb.w;
                            tre += r0, tim += i0;
                            tre += r1, tim += i1;
                            tre += r2, tim += i2;
                            tre += r3, tim += i3;
                            done += 4;
                        } else {
                            const int e = done & (ER - 1);
                            float ft = er[e];
#pragma unroll
                            for (int m = 1; m < M; ++m) ft += er[m * ER + e];
                            const float2 ph = __ldg(phi_ft + done);
                            tre += ft * ph.x;
                            tim += ft * ph.y;
                            done += 1;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctl->bar_tim[ck]);
                }
                if (on) {
                    if (tre != tre || tim != tim) {
                        cs->nan_flag = 1;
                        cs->nin_next = cs->nin;
                    } else {
                        cs->nan_flag = 0;
                        const float norm = (float)((double)em_atan2f(tim, tre) / EM_TWO_PI);
                        const float rx_timing = norm * (float)P;
                        const float old = cs->norm_rx_timing;
                        cs->norm_rx_timing = norm;
                        const float dn = norm - old;
                        if ((double)fabsf(dn) < .2) {
                            const float appm = (float)(1e6 * (double)dn / (double)(float)Nsym);
                            cs->ppm = (float)(.9 * (double)cs->ppm + .1 * (double)appm);
                        }
                        int nn = cs->nin;
                        if (!cfg.burst_mode && !cfg.lock_nin) {
                            if ((double)norm > 0.25)
                                nn = cfg.N + Ts / 4;
                            else if ((double)norm < -0.25)
                                nn = cfg.N - Ts / 4;
                            else
                                nn = cfg.N;
                        }
                        cs->nin_next = nn;
                        const int low = (int)floorf(rx_timing);
                        cs->low = low;
                        cs->fract = rx_timing - (float)low;
                        cs->one_minus_fract = 1.0f - cs->fract;
                        cs->high = (int)ceilf(rx_timing);
                        cs->rx_timing = rx_timing;
                    }
                }
                TMARK(t_tim);
            } else if (is_worker) {
                const int chshift = 31 - __clz(CH);
                for (int ck = 0; ck < nchunks; ++ck) {
                    const int j0 = ck * CH;
                    TMARK(t_int);
                    mbar_wait(&ctl->bar_osc[ck], (par >> ck) & 1u);
                    TMARK(t_wosc);
                    /* A8 mix: f_dc = in * conj(phi) in place over the oscillator values, one sample (all tones) per lane */
                    for (int t = wtid; t < (G << chshift); t += nwt) {
                        const int g = t >> chshift, j = j0 + (t & (CH - 1));
                        const Cap2 c = cap2(smem, cfg, g);
                        if (!c.cs->active || j >= c.cs->nin) continue;
                        const float2 x = raw_sample(srcp[2 * g + fb], fmt, j);
                        const unsigned sl = rsw((unsigned)c.cs->pos + (unsigned)j, rmask);
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            float2 *p = c.ring + m * RT + sl;
                            *p = mixc(x, *p);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctl->bar_mix[ck]);
                    TMARK(t_mix);
                    mbar_wait(&ctl->bar_mix[ck], (par >> ck) & 1u);
                    if (ck >= 3) mbar_wait(&ctl->bar_tim[ck - 3], (par >> (ck - 3)) & 1u);
                    TMARK(t_wmix);
                    /* A9 integrate + |.|^2 per tone for every window completed by this chunk */
                    if (cfg.k2.fast_ts && step == 1) {
                        for (int t = wtid; t < G * M * 8; t += nwt) {
                            const int q = t & 7, gm = t >> 3;
                            const int g = gm / M, m = gm - g * M;
                            const Cap2 c = cap2(smem, cfg, g);
                            const FrameView v = load_view(c.cs, true);
                            if (!v.active || !v.fast) continue;
                            const int w_lo = windows_after_chunk(v, ck - 1, CH, Ts, step, nint);
                            const int w_hi = windows_after_chunk(v, ck, CH, Ts, step, nint);
                            const float2 *ring_m = c.ring + m * RT;
                            float *er = c.ering + m * ER;
                            for (int e0 = w_lo + 16 * q; e0 < w_hi; e0 += 128) {
                                float *dst = er + (e0 & (ER - 1));
                                const unsigned s0 = v.B + (unsigned)e0;
                                switch (Ts) {
                                    case 24: integrate16<24>(ring_m, s0, umask, dst); break;
                                    case 20: integrate16<20>(ring_m, s0, umask, dst); break;
                                    default: integrate16<40>(ring_m, s0, umask, dst); break;
                                }
                            }
                        }
                    }
                    if (ctl->any_generic) {
                        /* generic windows: captures the fast path does not take */
                        const int wpc = CH / step + 2; /* upper bound of windows completed per chunk and capture */
                        for (int t = wtid; t < G * M * wpc; t += nwt) {
                            const int gm = t / wpc, k = t - gm * wpc;
                            const int g = gm / M, m = gm - g * M;
                            const Cap2 c = cap2(smem, cfg, g);
                            const FrameView v = load_view(c.cs, true);
                            if (!v.active || v.fast) continue;
                            const int w_lo = windows_after_chunk(v, ck - 1, CH, Ts, step, nint);
                            const int w_hi = windows_after_chunk(v, ck, CH, Ts, step, nint);
                            for (int i = w_lo + k; i < w_hi; i += wpc)
                                c.ering[m * ER + (i & (ER - 1))] = window_energy(c.ring + m * RT, v.B + (unsigned)(i * step), rmask, Ts);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctl->bar_int[ck]);
                }
                TMARK(t_int);
            } else if (is_fft) {
                /* A7 for frame f + 1 of capture `warp`, from the staging buffer fetched this iteration */
                if (cap2(smem, cfg, warp).cs->fon) freq_estimate(warp, fb ^ 1);
                TMARK(t_fft);
            }
            par ^= nchunks >= 32 ? ~0u : ((1u << nchunks) - 1u);
            __syncthreads();
#ifdef FSK2_TIMING
            if (warp == osc_warp) TMARK(t_oscw);
            else if (warp == tim_warp) TMARK(t_timw);
            else if (is_worker) TMARK(t_wosc);
            else TMARK(t_fftw);
#endif

            /* ---- A11: re-sum the two integrator outputs around the timing instant of every symbol, interpolate,
                        decide.  Lane (g, symbol), all tones. ---- */
            for (int it = tid; it < G * Nsym; it += nthreads) {
                const int g = it / Nsym, i = it - g * Nsym;
                const Cap2 c = cap2(smem, cfg, g);
                Fsk2CapShared *cs = c.cs;
                if (!cs->active || cs->nan_flag) continue;
                const int nold = cs->nold;
                const float fract = cs->fract, omf = cs->one_minus_fract;
                const int st = (i + 1) * P;
                const int olo = (st + cs->low) * step, ohi = (st + cs->high) * step;
                const int last = ohi + Ts; /* exclusive */
                u64 a[M], b[M];
#pragma unroll
                for (int m = 0; m < M; ++m) a[m] = b[m] = 0ull;
                int idx = olo;
                for (; idx < last && idx < nold; ++idx) {
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const u64 v = *reinterpret_cast<const u64 *>(c.tail + m * TAILN + (idx - nold + TAILN));
                        if (idx < olo + Ts) a[m] = add2(a[m], v);
                        if (idx >= ohi) b[m] = add2(b[m], v);
                    }
                }
                if (idx < last) {
                    const int jf = idx - nold; /* first new sample needed */
                    const int cq = jf / CKE;
                    float2 ph[M], d[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        ph[m] = c.ckpt[m * NCK + cq];
                        d[m] = cs->dphi[fb][m];
                    }
                    const unsigned char *src = srcp[2 * g + fb];
                    for (int jj = cq * CKE; jj < last - nold; ++jj) {
#pragma unroll
                        for (int m = 0; m < M; ++m) ph[m] = cmul(ph[m], d[m]);
                        if (jj >= jf) {
                            const float2 x = raw_sample(src, fmt, jj);
                            const int id = jj + nold;
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float2 fv = mixc(x, ph[m]);
                                const u64 v = pk(fv.x, fv.y);
                                if (id < olo + Ts) a[m] = add2(a[m], v);
                                if (id >= ohi) b[m] = add2(b[m], v);
                            }
                        }
                    }
                }
                float tmax[M];
                float *sig = c.ering, *nse = sig + Nsym, *sqm = nse + Nsym, *rfl = sqm + Nsym;
                const size_t cap = (size_t)cs->cap;
                const size_t frame = (size_t)f;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float2 fa = upk(a[m]), fb2 = upk(b[m]);
                    float2 t = make_float2(omf * fa.x, omf * fa.y);
                    t = cadd(t, make_float2(fract * fb2.x, fract * fb2.y));
                    tmax[m] = (t.x * t.x) + (t.y * t.y);
                    const float rf = sqrtf(tmax[m]);
                    rfl[m * Nsym + i] = rf; /* tone-major, as rx_filt */
                    if (args.rx_filt) args.rx_filt[(cap * args.max_frames + frame) * (size_t)(M * Nsym) + m * Nsym + i] = rf;
                }
                float mx = tmax[0];
                int sym = 0;
#pragma unroll
                for (int m = 1; m < M; ++m)
                    if (tmax[m] > mx) {
                        mx = tmax[m];
                        sym = m;
                    }
                if (args.bits) {
                    unsigned char *bo = args.bits + (cap * args.max_frames + frame) * (size_t)Nbits;
                    if (M == 2) {
                        bo[i] = sym == 1;
                    } else {
                        bo[2 * i + 1] = sym & 1;
                        bo[2 * i] = (sym & 2) >> 1;
                    }
                }
                float sum = 0.0f;
#pragma unroll
                for (int m = 0; m < M; ++m) sum += tmax[m];
                sig[i] = mx;
                nse[i] = (sum - mx) / (float)(M - 1);
                sqm[i] = sqrtf(mx);
            }
            if (args.bits) {
                /* the reference returns early on a NaN timing estimate and the CLI re-emits its bit buffer */
                for (int g = 0; g < G; ++g) {
                    Fsk2CapShared *cs = cap2(smem, cfg, g).cs;
                    if (!cs->active || !cs->nan_flag) continue;
                    unsigned char *bo = args.bits + ((size_t)cs->cap * args.max_frames + (size_t)f) * (size_t)Nbits;
                    for (int i = tid; i < Nbits; i += nthreads) bo[i] = f ? bo[i - Nbits] : 0;
                }
            }
            __syncthreads();
            TMARK(t_dec);
            if (warp < G) {
                /* four serial sums over the symbols of capture `warp`, lanes 0..3 */
                const Cap2 c = cap2(smem, cfg, warp);
                Fsk2CapShared *cs = c.cs;
                const bool on = cs->active && !cs->nan_flag;
                const int s = lane;
                if (on && s < 4) {
                    const float *sig = c.ering;
                    const float *src = s == 1 ? sig + Nsym : (s == 3 ? sig + 2 * Nsym : sig);
                    float acc = s == 1 ? 1E-12f : 0.0f;
#pragma unroll 10
                    for (int i = 0; i < Nsym; ++i) acc += src[i];
                    cs->sums[s] = acc;
                }
                __syncwarp();
                if (on && s == 0) {
                    float rx_sig_pow = cs->sums[0], rx_nse_pow = cs->sums[1], stdebno = cs->sums[2], meanebno = cs->sums[3];
                    rx_sig_pow = rx_sig_pow / (float)Nsym;
                    rx_nse_pow = rx_nse_pow / (float)Nsym;
                    cs->rx_sig_pow = rx_sig_pow;
                    cs->rx_nse_pow = rx_nse_pow;
                    cs->v_est = (float)sqrt((double)(rx_sig_pow - rx_nse_pow));
                    cs->SNRest = rx_sig_pow / rx_nse_pow;
                    meanebno = meanebno / (float)Nsym;
                    stdebno = (stdebno / (float)Nsym) - (meanebno * meanebno);
                    if ((double)stdebno > 0.0)
                        stdebno = (float)sqrt((double)stdebno);
                    else
                        stdebno = 0.0f;
                    const float ratio = (float)((1e-6 + (double)meanebno) / (1e-6 + (double)stdebno));
                    cs->EbNodB = -6.0f + (20.0f * em_log10f(ratio));
                    cs->snr_est = (float)(.5 * (double)cs->snr_est + .5 * (double)cs->EbNodB);
                    const float *fe = cfg.est_type ? cs->f2_est[fb] : cs->f_est[fb];
                    const float fc_avg = (fe[0] + fe[1]) / 2.0f;
                    const float fc_tx = (float)((cfg.f1_tx + cfg.f1_tx + cfg.tone_spacing) / 2);
                    cs->foff = fc_tx - fc_avg;
                }
            }
            /* f_dc tail of this frame for the next frame's decisions (the ring keeps it until then) */
            for (int t = tid; t < G * M * TAILN; t += nthreads) {
                const int gm = t / TAILN, k = t - gm * TAILN;
                const int g = gm / M, m = gm - g * M;
                const Cap2 c = cap2(smem, cfg, g);
                if (!c.cs->active) continue;
                const unsigned pend = (unsigned)c.cs->pos + (unsigned)c.cs->nin;
                c.tail[m * TAILN + k] = c.ring[m * RT + rsw(pend - (unsigned)TAILN + (unsigned)k, rmask)];
            }
            __syncthreads();
            TMARK(t_sum);

            /* ---- A12: soft decisions ([U mpdecode_core.c::fsk_rx_filt_to_llrs]) ---- */
            if (args.llr) {
                const int bps = M == 2 ? 1 : 2;
                for (int it = tid; it < G * Nsym; it += nthreads) {
                    const int g = it / Nsym, i = it - g * Nsym;
                    const Cap2 c = cap2(smem, cfg, g);
                    Fsk2CapShared *cs = c.cs;
                    if (!cs->active || cs->nan_flag) continue;
                    const float *rfl = c.ering + 3 * Nsym;
                    float l2[2];
                    symbol_llrs(rfl + i, Nsym, M, cs->v_est, cs->SNRest, l2);
                    float *lo = args.llr + ((size_t)cs->cap * args.max_frames + (size_t)f) * (size_t)Nbits;
                    for (int k = 0; k < bps; ++k) lo[bps * i + k] = l2[k];
                }
                __syncthreads();
            }
            TMARK(t_llr);

            /* ---- per-frame record, advance, next frame's bookkeeping ---- */
            if (tid == 0) {
                for (int g = 0; g < G; ++g) {
                    Fsk2CapShared *cs = cap2(smem, cfg, g).cs;
                    if (!cs->active) continue;
                    if (args.stats) {
                        pirip_frame_stats st;
                        for (int m = 0; m < 4; ++m) {
                            st.f_est[m] = m < M ? cs->f_est[fb][m] : 0.0f;
                            st.f2_est[m] = m < M ? cs->f2_est[fb][m] : 0.0f;
                        }
                        st.norm_rx_timing = cs->norm_rx_timing;
                        st.ppm = cs->ppm;
                        st.SNRest = cs->SNRest;
                        st.v_est = cs->v_est;
                        st.rx_sig_pow = cs->rx_sig_pow;
                        st.rx_nse_pow = cs->rx_nse_pow;
                        st.EbNodB = cs->EbNodB;
                        st.snr_est = cs->snr_est;
                        st.foff = cs->foff;
                        st.rx_timing = cs->rx_timing;
                        st.nin_next = cs->nin_next;
                        st.sample_offset = (int32_t)cs->pos;
                        args.stats[(size_t)cs->cap * args.max_frames + (size_t)f] = st;
                    }
                    cs->pos += (unsigned long long)cs->nin;
                    cs->nin = cs->nin_next;
                    cs->frame += 1;
                }
                bookkeeping();
            }
            __syncthreads();
            TMARK(t_adv);
        }

        /* ---- results of the group ---- */
        for (int g = 0; g < G; ++g) {
            const Cap2 c = cap2(smem, cfg, g);
            const size_t cap = (size_t)c.cs->cap;
            const int nfr = c.cs->frame;
            if (args.nframes && tid == 0) args.nframes[cap] = (int32_t)nfr;
            if (args.Sf_out) {
                const float *Sf = c.Sf + (size_t)((nfr + 1) & 1) * Ndft; /* spectrum after the last demodulated frame */
                for (int i = tid; i < Ndft; i += nthreads) args.Sf_out[cap * (size_t)Ndft + i] = Sf[i];
            }
        }
    }
#ifdef FSK2_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == NW - 3 || warp == tim_warp || warp == osc_warp))
        printf("w%02d pre %lld adv %lld | osc %lld oscwait %lld | tim %lld timwait %lld | w-osc %lld mix %lld w-mix %lld int %lld | fft %lld "
               "fftwait %lld | dec %lld sum %lld llr %lld\n",
               warp, t_pre, t_adv, t_osc, t_oscw, t_tim, t_timw, t_wosc, t_mix, t_wmix, t_int, t_fft, t_fftw, t_dec, t_sum, t_llr);
    (void)t_wtim;
    (void)t_stream;
#endif
}

template <int M, int NW>
cudaError_t launch_mn(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream, bool prepare_only) {
    if (prepare_only)
        return cudaFuncSetAttribute(fsk2_kernel<M, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.k2.smem_bytes);
    const size_t sms = (size_t)cfg.k2.num_sms;
    const unsigned grid = (unsigned)(n_captures < sms ? n_captures : sms);
    fsk2_kernel<M, NW><<<grid, NW * 32, cfg.k2.smem_bytes, stream>>>(cfg, args);
    return cudaGetLastError();
}

cudaError_t dispatch(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n, cudaStream_t stream, bool prep) {
    if (cfg.k2.threads != 768) return cudaErrorInvalidValue;
    if (cfg.M == 2) return launch_mn<2, 24>(cfg, args, n, stream, prep);
    return launch_mn<4, 24>(cfg, args, n, stream, prep);
}

}  // namespace

cudaError_t fsk2_kernel_prepare(const FskDevCfg &cfg) {
    FskBatchArgs dummy;
    return dispatch(cfg, dummy, 0, nullptr, true);
}

cudaError_t fsk2_launch_demod(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream) {
    if (n_captures == 0) return cudaSuccess;
    return dispatch(cfg, args, n_captures, stream, false);
}
