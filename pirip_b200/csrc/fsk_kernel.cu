/*
 * fsk_kernel.cu — the FSK demodulator hot path on sm_100a.
 *
 * One persistent CTA demodulates G independent captures in lockstep, running the whole
 * `fsk_demod` main loop ([U codec2 fsk_demod.c], pirip call sites [R README.md:105,109],
 * [R test/loopback_rtl_sdr.sh:16]) on the device.  Per frame and capture it converts nin
 * IQ samples, estimates the tone frequencies (windowed FFT blocks + IIR spectrum + peak
 * pick / mask, [U fsk.c::fsk_demod_freq_est]), mixes every tone to DC, integrates over a
 * symbol at (Nsym+1)*P offsets, extracts the symbol-rate timing line, resamples, decides
 * bits and maps soft decisions ([U fsk.c::fsk_demod_core],
 * [U mpdecode_core.c::fsk_rx_filt_to_llrs]).  All demodulator state (f_dc memory, Sf,
 * oscillator phases, nin) stays in shared memory between frames; HBM traffic is the u8 IQ
 * in (TMA bulk copies, one frame ahead) and the bits out.
 *
 * NUMERICS CONTRACT.  The kernel performs the same IEEE-754 operations in the same order
 * as the reference algorithm: sequential accumulation where the reference accumulates
 * sequentially (integrators, timing line, oscillator recursion), the same radix-4/2
 * decimation-in-time butterflies as kiss_fft, and exact_math.h for every transcendental.
 * Compile with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false.  The output — bits
 * AND floats — is bit-identical to a CPU evaluation of the algorithm (tests/ check this).
 *
 * PARALLEL STRUCTURE.  Three recurrences of the algorithm are inherently serial in
 * floating point: the per-tone oscillator phi *= dphi (nin steps), the timing-line sum
 * over (Nsym+1)*P integrator outputs, and the per-frame power sums.  They run on single
 * lanes, but (a) the lanes of one warp carry the recurrences of all G captures x M tones
 * at once, and (b) the frame is a streaming pipeline: the oscillator warp publishes its
 * progress chunk by chunk through mbarriers, the worker warps mix and integrate right
 * behind it, and the timing warp accumulates the timing line right behind them, so
 * everything downstream of the oscillator hides under its latency.  The two serial warps
 * are the highest-numbered warps of the CTA (the issue arbiter favours high warp ids).
 * FFTs, IIR, peak search, decisions and LLRs are data-parallel phases between CTA barriers.
 */
#include <stdio.h>

#include "exact_math.h"
#include "fsk_dev.h"
#include "llr_dev.h"

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    float2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

/* mbarrier (shared::cta) wrappers: arrive has release, try_wait has acquire semantics at CTA
   scope, and a waiting warp is suspended by the hardware instead of spinning on issue slots */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
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
/* TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier */
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

/* FFT scratch index swizzle: XOR bits [5:4] of the element index into bits [3:2] and
   [1:0], so the radix-4 stages with m = 1 and m = 4 (16 lanes, 8-byte elements) touch 16
   distinct bank pairs; unit-stride stages are unaffected (constant XOR per 16-block) */
__device__ __forceinline__ int fswz(int e) { return e ^ (5 * ((e >> 4) & 3)); }

/* One warp: Hann-window N = 2^LOGN samples of x, FFT them in the warp's scratch slot with
   kiss_fft's decimation-in-time factor order 4,4,..,(2), write the fft-shifted magnitudes.
   perm[n] is the (swizzled) slot of input n before the first butterfly level. */
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

template <int LOGN>
__device__ __noinline__ void warp_fft_mag(const float2 *xs, const unsigned char *src, int fmt, int first, float2 *buf,
                                             float *out, const float2 *tw, const float *hann, const unsigned short *perm,
                                             int lane) {
    constexpr int N = 1 << LOGN;
    constexpr int N4 = LOGN / 2;          /* radix-4 levels */
    constexpr bool HAS2 = (LOGN & 1) != 0; /* innermost radix-2 level */
#pragma unroll
    for (int q = 0; q < N / 32; ++q) {
        const int n = lane + (q << 5);
        const float h = hann[n];
        const float2 v = xs ? xs[first + n] : raw_sample(src, fmt, first + n);
        buf[perm[n]] = make_float2(h * v.x, h * v.y);
    }
    __syncwarp();
    if (HAS2) {
        /* m = 1, fstride = N/2: the only twiddle is tw[0] */
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
        const int lm = (HAS2 ? 1 : 0) + 2 * (N4 - 1 - L); /* log2 m */
        const int m = 1 << lm;
        const int fs = N >> (lm + 2); /* fstride */
        constexpr int NB = N / 128 > 0 ? N / 128 : 1; /* butterflies per lane at this level */
        constexpr int PB = NB >= 2 ? 2 : 1;           /* processed two at a time: loads, math, stores */
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
    /* fft shift (DC -> bin N/2) and magnitude */
#pragma unroll
    for (int q = 0; q < N / 32; ++q) {
        const int i = lane + (q << 5);
        const float2 X = buf[fswz((i + (N >> 1)) & (N - 1))];
        out[i] = sqrtf((X.x * X.x) + (X.y * X.y));
    }
    __syncwarp();
}

/* number of integrator outputs whose whole window has been mixed once `j` new samples of
   this frame are in place (window i covers f_dc[i*step, i*step+Ts), new samples start at nold) */
__device__ __forceinline__ int outputs_ready(int nold, int j, int Ts, int step, int nint) {
    const int have = nold + j - Ts;
    if (have < 0) return 0;
    const int r = (step == 1 ? have : have / step) + 1;
    return r < nint ? r : nint;
}

/* sum of Ts consecutive samples starting at src, ascending order from zero */
__device__ __forceinline__ float2 window_sum(const float2 *src, int Ts) {
    float ar = 0.0f, ai = 0.0f;
#pragma unroll 4
    for (int j = 0; j < Ts; ++j) {
        const float2 v = src[j];
        ar += v.x;
        ai += v.y;
    }
    return make_float2(ar, ai);
}

/* two window sums at once (independent chains interleaved for latency) */
__device__ __forceinline__ void window_sum2(const float2 *sa, const float2 *sb, int Ts, float2 &a, float2 &b) {
    float ar = 0.0f, ai = 0.0f, br = 0.0f, bi = 0.0f;
#pragma unroll 4
    for (int j = 0; j < Ts; ++j) {
        const float2 v = sa[j], w = sb[j];
        ar += v.x;
        ai += v.y;
        br += w.x;
        bi += w.y;
    }
    a = make_float2(ar, ai);
    b = make_float2(br, bi);
}

struct CapPtrs {
    float2 *fdc;
    float2 *xin; /* frame input; later the timing-line products */
    float *mag;  /* FFT magnitudes; later tmpS | sig | nse | sqm | rx_filt */
    float *Sf;
    unsigned char *stage; /* TMA landing zone for the next frame's raw samples */
    float2 *phead;        /* first D timing-line products */
    FskCapShared *cs;
};

__device__ __forceinline__ CapPtrs cap_ptrs(unsigned char *smem, const FskDevCfg &cfg, int g) {
    unsigned char *cb = smem + (size_t)g * cfg.cap_stride;
    CapPtrs p;
    p.fdc = reinterpret_cast<float2 *>(cb + cfg.off_fdc);
    p.xin = reinterpret_cast<float2 *>(cb + cfg.off_xin);
    p.mag = reinterpret_cast<float *>(cb + cfg.off_mag);
    p.Sf = reinterpret_cast<float *>(cb + cfg.off_Sf);
    p.stage = cb + cfg.off_stage;
    p.phead = reinterpret_cast<float2 *>(cb + cfg.off_phead);
    p.cs = reinterpret_cast<FskCapShared *>(cb + cfg.off_scal);
    return p;
}

/* Timing-line product of integrator output e.  Output e becomes ready nold - Ts + 1 samples
   ahead of the input index, so the products live in the consumed part of the input buffer
   shifted back by D = Ts + Ts/4 + 1 (>= nold_max - Ts + 1); the first D have their own slots. */
__device__ __forceinline__ float2 *prod_slot(const CapPtrs &c, int e, int D) {
    return e < D ? c.phead + e : c.xin + (e - D);
}

__device__ __forceinline__ size_t bytes_per_sample(int fmt) {
    return fmt == PIRIP_FMT_U8 ? 2 : (fmt == PIRIP_FMT_S16 ? 4 : 8);
}

/* Issue the TMA bulk copy of the raw samples [pos, pos + nin_max) of capture `cap` into its
   staging buffer (16-byte aligned superset, clipped to the batch allocation). */
__device__ void stage_issue(const FskDevCfg &cfg, const FskBatchArgs &args, const CapPtrs &c, size_t cap,
                            unsigned long long pos, unsigned long long *bar) {
    FskCapShared *cs = c.cs;
    cs->stg_valid = 0;
    if (!args.tma_ok) return;
    const size_t bps = bytes_per_sample(cfg.fmt);
    if (pos >= args.samples_per_capture) return;
    const size_t base = reinterpret_cast<size_t>(args.iq);
    const size_t a = base + cap * args.capture_stride_bytes + pos * bps;
    const size_t lo = a & ~(size_t)15;
    size_t hi = (a + (size_t)(cfg.N + cfg.Ts / 4) * bps + 15) & ~(size_t)15;
    const size_t safe = (base + (args.n_captures - 1) * args.capture_stride_bytes + args.samples_per_capture * bps) & ~(size_t)15;
    if (hi > safe) hi = safe;
    if (hi <= lo) return;
    const unsigned bytes = (unsigned)(hi - lo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_1d(c.stage, reinterpret_cast<const void *>(lo), bytes, bar);
    cs->stg_lo = (unsigned long long)lo;
    cs->stg_bytes = bytes;
    cs->stg_valid = 1;
}


/* A team of warps that synchronises on its own barrier: the whole CTA (id 0, __syncthreads) or a
   subset on a named barrier */
struct Team {
    int nw, w, lane, bar;
};
__device__ __forceinline__ void team_sync(const Team &t) {
    if (t.bar == 0)
        __syncthreads();
    else
        named_bar_sync(t.bar, t.nw * 32);
}

struct CtaTables {
    float2 *fft_slots; /* one Ndft slot per warp of the CTA */
    const float2 *tw;
    const float *hann;
    const unsigned short *perm;
};

/* A7: tone frequency estimate of one frame for every capture of the CTA, by a team of warps.
   Reads the raw samples of the frame through cs->fsrc (TMA-staged copy or global memory), leaves
   f_est / f2_est / dphi in estimate slot `slot` and the updated IIR spectrum in Sf.
   FFT blocks are transformed mag_blocks at a time; the magnitude IIR runs over each batch in block
   order, so a small magnitude buffer serves any number of blocks. */
template <int G, bool FROM_XIN>
__device__ void front_stage(const FskDevCfg &cfg, unsigned char *smem, const CtaTables &tb, const Team &tm, int slot, int cta_warp) {
    const int Ndft = cfg.Ndft, M = cfg.M, MB = cfg.mag_blocks;
    const int lane = tm.lane, tthreads = tm.nw * 32, ttid = tm.w * 32 + lane;
    const float tc = cfg.tc;
    const float omtc = 1.0f - tc;
    float2 *fft_slot = tb.fft_slots + (size_t)cta_warp * Ndft;
    for (int b0 = 0; b0 < cfg.max_numffts; b0 += MB) {
        for (int job = tm.w; job < G * MB; job += tm.nw) {
            const int g = job / MB, jb = job - g * MB, j = b0 + jb;
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            if (!c.cs->fon || j >= c.cs->fnumffts) continue;
            /* the frame already converted into shared memory, or the raw (staged / global) samples */
            const float2 *xs = FROM_XIN ? c.xin : nullptr;
            const unsigned char *src = FROM_XIN ? nullptr : c.cs->fsrc;
            const int fmt = cfg.fmt;
            float *out = c.mag + (size_t)jb * Ndft;
            const int first = j * (Ndft >> 1);
            switch (cfg.log2_ndft) {
                case 6: warp_fft_mag<6>(xs, src, fmt, first, fft_slot, out, tb.tw, tb.hann, tb.perm, lane); break;
                case 7: warp_fft_mag<7>(xs, src, fmt, first, fft_slot, out, tb.tw, tb.hann, tb.perm, lane); break;
                case 8: warp_fft_mag<8>(xs, src, fmt, first, fft_slot, out, tb.tw, tb.hann, tb.perm, lane); break;
                case 9: warp_fft_mag<9>(xs, src, fmt, first, fft_slot, out, tb.tw, tb.hann, tb.perm, lane); break;
                default: warp_fft_mag<10>(xs, src, fmt, first, fft_slot, out, tb.tw, tb.hann, tb.perm, lane); break;
            }
        }
        team_sync(tm);
        const bool last = b0 + MB >= cfg.max_numffts;
        for (int g = 0; g < G; ++g) {
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            if (!c.cs->fon) continue;
            const int jn = c.cs->fnumffts - b0 < MB ? c.cs->fnumffts - b0 : MB; /* blocks of this batch */
            for (int i = ttid; i < Ndft; i += tthreads) {
                float s = c.Sf[i];
                for (int j = 0; j < jn; ++j) s = (s * omtc) + (c.mag[j * Ndft + i] * tc);
                c.Sf[i] = s;
                if (last) c.mag[i] = s; /* tmpS: scratch copy the peak search blanks */
            }
        }
        team_sync(tm);
    }
    for (int g = tm.w; g < G; g += tm.nw) {
        const CapPtrs c = cap_ptrs(smem, cfg, g);
        if (!c.cs->fon) continue;
        float *tmpS = c.mag;
        /* estimator 1: M rounds of argmax (first maximum wins) + blanking */
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
            if (lane == 0) c.cs->freqi[r] = ibest - (Ndft >> 1);
            __syncwarp();
        }
        if (lane == 0) {
            int fi[4];
            for (int m = 0; m < M; ++m) fi[m] = c.cs->freqi[m];
            for (int a = 1; a < M; ++a) { /* ascending insertion sort */
                const int key = fi[a];
                int b = a - 1;
                while (b >= 0 && fi[b] > key) {
                    fi[b + 1] = fi[b];
                    --b;
                }
                fi[b + 1] = key;
            }
            for (int m = 0; m < M; ++m) c.cs->f_est[slot][m] = (float)fi[m] * cfg.fs_over_ndft;
        }
        /* estimator 2: slide the tone mask over Sf, first maximum wins */
        {
            float best = 0.0f;
            int bbest = 0x7fffffff;
            for (int b = cfg.st + lane; b < cfg.en - cfg.len_mask; b += 32) {
                float corr = 0.0f;
                for (int q = 0; q < cfg.nmask; ++q) corr += c.Sf[b + cfg.mask_pos[q]];
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
                for (int m = 0; m < M; ++m) c.cs->f2_est[slot][m] = foff + (float)(m * cfg.tone_spacing);
            }
        }
        __syncwarp();
        if (lane < M) {
            const float f = cfg.est_type ? c.cs->f2_est[slot][lane] : c.cs->f_est[slot][lane];
            const float arg = (float)(EM_TWO_PI * (double)(f / (float)cfg.Fs));
            double sn, csn;
            em_sincos((double)arg, &sn, &csn);
            c.cs->dphi[slot][lane] = make_float2((float)csn, (float)sn);
        }
    }
}

/* Raw samples of the frame starting at sample `pos` of capture `cap`: the TMA-staged copy when it
   covers the range, global memory otherwise */
__device__ __forceinline__ const unsigned char *frame_src(const FskDevCfg &cfg, const FskBatchArgs &args, const CapPtrs &c,
                                                          size_t cap, unsigned long long pos, int nsamples) {
    const size_t bps = bytes_per_sample(cfg.fmt);
    const size_t a = reinterpret_cast<size_t>(args.iq) + cap * args.capture_stride_bytes + pos * bps;
    if (c.cs->stg_valid && a >= c.cs->stg_lo && a + (size_t)nsamples * bps <= c.cs->stg_lo + c.cs->stg_bytes)
        return c.stage + (a - c.cs->stg_lo);
    return reinterpret_cast<const unsigned char *>(a);
}

/* G <= 2: 256-thread CTAs, two resident per SM, so the barrier-separated phases of one CTA
   overlap the other's; G >= 3: one 512-thread CTA per SM */
template <int G>
__global__ void __launch_bounds__(G <= 2 ? 256 : 512, G <= 2 ? 2 : 1) fsk_demod_kernel(const __grid_constant__ FskDevCfg cfg,
                                                           const __grid_constant__ FskBatchArgs args) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int M = cfg.M, Nmem = cfg.Nmem, Ndft = cfg.Ndft, nint = cfg.nint, Nsym = cfg.Nsym, P = cfg.P, Ts = cfg.Ts;
    const int Nbits = cfg.Nbits, TS = cfg.fdc_tone_stride, CH = cfg.chunk, step = Ts / P;
    const size_t cap0 = (size_t)blockIdx.x * G;
    unsigned char *cta = smem + (size_t)G * cfg.cap_stride;
    FskCtaShared *ctl = reinterpret_cast<FskCtaShared *>(cta + cfg.off_ctl);
    float2 *tw = reinterpret_cast<float2 *>(cta + cfg.off_tw);
    float *hann = reinterpret_cast<float *>(cta + cfg.off_hann);
    const float2 *phi_ft = cfg.phi_ft; /* global (L2-resident); loads are issued ahead of their use */
    unsigned short *perm = reinterpret_cast<unsigned short *>(cta + cfg.off_perm);
    const int D = Ts + Ts / 4 + 1;
    /* Roles of the streaming phase.  Warp w issues on sub-partition w % 4.  The two serial
       warps (oscillators, timing line) share sub-partition 3 and no worker runs there, so the
       latency-bound recurrences never wait for an issue slot; the workers are the warps of
       sub-partitions 0..2. */
    const int osc_warp = nwarps - 1, tim_warp = nwarps - 5;
    /* When the frequency estimate of the next frame can run one frame ahead (its FFT block count
       does not depend on the timing decision and no state is carried out of the launch), the warps
       of sub-partition 2 and the spare warps of sub-partition 3 form the front team that computes
       it during the streaming phase; the workers are then the warps of sub-partitions 0 and 1. */
    const bool pipelined = cfg.pipeline && args.state == nullptr;
    const bool is_serial = warp == osc_warp || warp == tim_warp;
    const bool is_worker = pipelined ? (warp & 3) < 2 : (warp & 3) != 3;
    const bool is_front = pipelined && !is_worker && !is_serial;
    const int nww = pipelined ? nwarps / 2 : nwarps - nwarps / 4; /* worker warps */
    const int nwt = nww * 32;                                      /* worker threads */
    const int wtid = (pipelined ? ((warp >> 2) * 2 + (warp & 3)) : (warp - (warp >> 2))) * 32 + lane; /* dense index */
    const int nfront = nwarps / 2 - 2; /* front warps: w%4 == 2, and w%4 == 3 except the two serial warps */
    int front_idx = 0;
    for (int w = 0; w < warp; ++w)
        if ((w & 3) >= 2 && w != osc_warp && w != tim_warp) front_idx++;
    const CtaTables tb = { reinterpret_cast<float2 *>(cta + cfg.off_fft), tw, hann, perm };
    int slot = 0; /* estimate slot of the frame being demodulated */
    unsigned long long par = 0;        /* phase parity of every per-chunk mbarrier, bit ck */
    unsigned par_in = 0;               /* phase parity of the staging mbarriers, bit g */

    /* ---- tables, barriers, initial state -------------------------------------------- */
    for (int i = tid; i < Ndft; i += nthreads) {
        tw[i] = cfg.twiddle[i];
        hann[i] = cfg.hann[i];
        perm[i] = cfg.perm[i];
    }
    if (tid == 0) {
        for (int i = 0; i < FSK_MAX_CHUNKS; ++i) {
            mbar_init(&ctl->bar_osc[i], 1);
            mbar_init(&ctl->bar_int[i], nww);
        }
        for (int g = 0; g < FSK_MAX_G; ++g) mbar_init(&ctl->bar_in[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int g = 0; g < G; ++g) {
        const CapPtrs c = cap_ptrs(smem, cfg, g);
        const size_t cap = cap0 + g;
        const bool exists = cap < args.n_captures;
        float *state = (args.state && exists) ? args.state + cap * args.state_stride_floats : nullptr;
        if (state) {
            const FskDevStateHdr *h = reinterpret_cast<const FskDevStateHdr *>(state);
            const float *sSf = state + sizeof(FskDevStateHdr) / sizeof(float);
            const float2 *sfdc = reinterpret_cast<const float2 *>(sSf + Ndft);
            for (int i = tid; i < Ndft; i += nthreads) c.Sf[i] = sSf[i];
            for (int m = 0; m < M; ++m)
                for (int i = tid; i < Nmem; i += nthreads) c.fdc[m * TS + i] = sfdc[m * Nmem + i];
            if (tid == 0) {
                for (int m = 0; m < 4; ++m) c.cs->phi_c[m] = h->phi_c[m];
                c.cs->norm_rx_timing = h->norm_rx_timing;
                c.cs->ppm = h->ppm;
                c.cs->snr_est = h->snr_est;
                c.cs->EbNodB = h->EbNodB;
                c.cs->nin = h->nin;
            }
        } else {
            for (int i = tid; i < Ndft; i += nthreads) c.Sf[i] = 0.0f;
            for (int i = tid; i < M * TS; i += nthreads) c.fdc[i] = make_float2(0.0f, 0.0f);
            if (tid == 0) {
                for (int m = 0; m < 4; ++m) c.cs->phi_c[m] = make_float2(1.0f, 0.0f);
                c.cs->norm_rx_timing = 0.0f;
                c.cs->ppm = 0.0f;
                c.cs->snr_est = 0.0f;
                c.cs->EbNodB = 0.0f;
                c.cs->nin = cfg.N;
            }
        }
        if (tid == 0) {
            for (int m = 0; m < 4; ++m) {
                c.cs->f_est[0][m] = c.cs->f_est[1][m] = 0.0f;
                c.cs->f2_est[0][m] = c.cs->f2_est[1][m] = 0.0f;
            }
            c.cs->SNRest = c.cs->v_est = c.cs->rx_sig_pow = c.cs->rx_nse_pow = c.cs->foff = c.cs->rx_timing = 0.0f;
            c.cs->pos = 0;
            c.cs->frame = 0;
            c.cs->active = exists ? 1 : 0;
            c.cs->nan_flag = 0;
            c.cs->stg_valid = 0;
            c.cs->fon = 0;
        }
    }
    __syncthreads();
    if (tid < G) {
        const CapPtrs c = cap_ptrs(smem, cfg, tid);
        if (c.cs->active) stage_issue(cfg, args, c, cap0 + tid, 0, &ctl->bar_in[tid]);
    }
    __syncthreads();
    if (pipelined) {
        /* prologue: the whole CTA computes the frequency estimate of frame 0 */
        if (tid < G) {
            const CapPtrs c = cap_ptrs(smem, cfg, tid);
            FskCapShared *cs = c.cs;
            cs->fon = cs->active && args.max_frames > 0 && (unsigned long long)(cfg.N - Ts / 4) <= args.samples_per_capture;
            cs->fnumffts = cfg.numffts_const;
            if (cs->fon) cs->fsrc = frame_src(cfg, args, c, cap0 + tid, 0, cfg.N - Ts / 4);
        }
        __syncthreads();
        for (int g = 0; g < G; ++g) {
            FskCapShared *cs = cap_ptrs(smem, cfg, g).cs;
            if (cs->fon && cs->stg_valid) mbar_wait(&ctl->bar_in[g], (par_in >> g) & 1u);
        }
        const Team all = { nwarps, warp, lane, 0 };
        front_stage<G, false>(cfg, smem, tb, all, 0, warp);
        __syncthreads();
    }

#ifdef FSK_PHASE_TIMING
    long long tph[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    long long tlast = clock64();
    int tsink = 0;
#define PHASE_MARK(k)                     \
    do {                                  \
        /* a barrier only blocks at the next access to protected state: touch shared memory first */ \
        tsink += *reinterpret_cast<volatile int *>(&ctl->max_nin); \
        const long long tnow = clock64() + (tsink & 0); \
        tph[k] += tnow - tlast;           \
        tlast = tnow;                     \
    } while (0)
    long long wph[6] = { 0, 0, 0, 0, 0, 0 };
    long long wlast = 0;
#define WMARK(k)                                         \
    do {                                                 \
        tsink += *reinterpret_cast<volatile int *>(&ctl->max_nin); \
        const long long tnow = clock64() + (tsink & 0);  \
        if (k) wph[k] += tnow - wlast;                   \
        wlast = tnow;                                    \
    } while (0)
#else
#define PHASE_MARK(k)
#define WMARK(k)
#endif
    /* frame bookkeeping: which captures still have a whole frame to read.  Done by one thread
       before the first frame and at the end of every frame, under that frame's last barrier. */
    auto bookkeeping = [&]() {
        int any = 0, mx = 0;
        for (int g = 0; g < G; ++g) {
            FskCapShared *cs = cap_ptrs(smem, cfg, g).cs;
            int act = cs->active;
            if (act && ((size_t)cs->frame >= args.max_frames ||
                        cs->pos + (unsigned long long)cs->nin > args.samples_per_capture))
                act = 0; /* short read ends the loop */
            cs->active = act;
            cs->nold = Nmem - cs->nin;
            cs->numffts = cs->nin / (Ndft >> 1) - 1;
            if (!pipelined) {
                cs->fon = act;
                cs->fnumffts = cs->numffts;
            }
            if (act) {
                any = 1;
                mx = cs->nin > mx ? cs->nin : mx;
            }
        }
        ctl->any_active = any;
        ctl->max_nin = mx;
    };
    if (tid == 0) bookkeeping();
    __syncthreads();
    for (;;) {
        if (!ctl->any_active) break;
        const int max_nin = ctl->max_nin;
        PHASE_MARK(0);

        /* ---- A5: input conversion ([U fsk_demod.c]) + A8: slide the f_dc memory ---------- */
        for (int g = 0; g < G; ++g) {
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            FskCapShared *cs = c.cs;
            if (cs->stg_valid) {
                /* the staged copy issued one frame ago must have landed before anyone moves on,
                   whether or not the capture still has a frame to demodulate */
                mbar_wait(&ctl->bar_in[g], (par_in >> g) & 1u);
                par_in ^= 1u << g;
            }
            if (!cs->active) continue;
            const int nin = cs->nin, nold = cs->nold;
            const unsigned char *src = frame_src(cfg, args, c, cap0 + g, cs->pos, nin);
            if (cfg.fmt == PIRIP_FMT_U8) {
                /* 4-byte reads (two samples) over the aligned body, single samples at the ends */
                const int mis = (int)((reinterpret_cast<size_t>(src) >> 1) & 1);
                if (mis && tid == 0) c.xin[0] = make_float2(u8f(src[0]), u8f(src[1]));
                const int npair = (nin - mis) >> 1;
                const unsigned *s4 = reinterpret_cast<const unsigned *>(src + 2 * mis);
                for (int k = tid; k < npair; k += nthreads) {
                    const unsigned w = s4[k];
                    float2 *dst = c.xin + mis + 2 * k;
                    dst[0] = make_float2(u8f(w & 0xff), u8f((w >> 8) & 0xff));
                    dst[1] = make_float2(u8f((w >> 16) & 0xff), u8f(w >> 24));
                }
                if (((nin - mis) & 1) && tid == 32) {
                    const int j = nin - 1;
                    c.xin[j] = make_float2(u8f(src[2 * j]), u8f(src[2 * j + 1]));
                }
            } else if (cfg.fmt == PIRIP_FMT_S16) {
                const short *s2 = reinterpret_cast<const short *>(src);
                float *xf = reinterpret_cast<float *>(c.xin);
                for (int i = tid; i < 2 * nin; i += nthreads) xf[i] = (float)s2[i] / 825.0f; /* FDMDV_SCALE */
            } else {
                const float2 *s8 = reinterpret_cast<const float2 *>(src);
                for (int i = tid; i < nin; i += nthreads) c.xin[i] = s8[i];
            }
            if (nold <= nin) {
                for (int m = 0; m < M; ++m)
                    for (int i = tid; i < nold; i += nthreads) c.fdc[m * TS + i] = c.fdc[m * TS + i + nin];
            } else if (tid < M) {
                for (int i = 0; i < nold; ++i) c.fdc[tid * TS + i] = c.fdc[tid * TS + i + nin];
            }
        }
        __syncthreads();
        PHASE_MARK(1);

        if (!pipelined) {
            /* the staging buffers are free again (the frame now lives in xin): fetch the next frame's
               samples behind the compute, then A7 for this frame by the whole CTA */
            if (tid < G) {
                const CapPtrs c = cap_ptrs(smem, cfg, tid);
                if (c.cs->active)
                    stage_issue(cfg, args, c, cap0 + tid, c.cs->pos + (unsigned long long)c.cs->nin, &ctl->bar_in[tid]);
                else
                    c.cs->stg_valid = 0;
            }
            const Team all = { nwarps, warp, lane, 0 };
            front_stage<G, true>(cfg, smem, tb, all, slot, warp);
        } else if (tid < G) {
            /* front team mode: fetch the next frame and tell the team where its raw samples are */
            const CapPtrs c = cap_ptrs(smem, cfg, tid);
            FskCapShared *cs = c.cs;
            cs->fon = 0;
            if (cs->active) {
                const unsigned long long npos = cs->pos + (unsigned long long)cs->nin;
                stage_issue(cfg, args, c, cap0 + tid, npos, &ctl->bar_in[tid]);
                /* the next frame exists for some nin iff its shortest form fits the capture */
                if ((size_t)cs->frame + 1 < args.max_frames &&
                    npos + (unsigned long long)(cfg.N - Ts / 4) <= args.samples_per_capture) {
                    cs->fon = 1;
                    cs->fnumffts = cfg.numffts_const;
                    cs->fsrc = frame_src(cfg, args, c, cap0 + tid, npos, cfg.N - Ts / 4);
                }
            } else {
                cs->stg_valid = 0;
            }
        }
        __syncthreads();
        PHASE_MARK(3);

        /* ================= streaming phase: oscillators -> mix -> integrate -> timing line ===== */
        const int nchunks = (max_nin + CH - 1) / CH;
        if (warp == osc_warp) {
            /* A8: the oscillator recursion phi *= dphi is serial; lane (g, m) carries tone m of
               capture g, all lanes step together and publish progress once per chunk */
            const int g = lane / M, m = lane - g * M;
            const bool on = lane < G * M && cap_ptrs(smem, cfg, g < G ? g : 0).cs->active;
            const CapPtrs c = cap_ptrs(smem, cfg, on ? g : 0);
            const int my_nin = on ? c.cs->nin : 0;
            const float2 d = on ? c.cs->dphi[slot][m] : make_float2(1.0f, 0.0f);
            float2 ph = on ? c.cs->phi_c[m] : make_float2(1.0f, 0.0f);
            float2 *dst = c.fdc + m * TS + c.cs->nold;
            /* smallest nin over the active lanes: below it no lane needs a bounds check */
            int lo = on ? my_nin : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int o = __shfl_xor_sync(0xffffffffu, lo, off);
                lo = o < lo ? o : lo;
            }
            for (int ck = 0; ck < nchunks; ++ck) {
                const int j0 = ck * CH;
                const int jend = j0 + CH < max_nin ? j0 + CH : max_nin;
                if (jend <= lo) {
                    if (on) {
                        /* two steps per 16-byte store once the destination is 16-byte aligned */
                        int j = j0;
                        if ((reinterpret_cast<size_t>(dst + j) & 15) != 0) {
                            ph = cmul(ph, d);
                            dst[j++] = ph;
                        }
#pragma unroll 4
                        for (; j + 1 < jend; j += 2) {
                            const float2 p1 = cmul(ph, d);
                            ph = cmul(p1, d);
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(p1.x, p1.y, ph.x, ph.y);
                        }
                        if (j < jend) {
                            ph = cmul(ph, d);
                            dst[j] = ph;
                        }
                    }
                } else {
                    for (int j = j0; j < jend; ++j) {
                        if (j < my_nin) {
                            ph = cmul(ph, d);
                            dst[j] = ph;
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
        } else if (warp == tim_warp) {
            /* A10: the timing-line sum is a serial float accumulation; lane g carries capture g
               (real and imaginary chains interleaved) right behind the integrating workers */
            const int g = lane;
            const bool on = g < G && cap_ptrs(smem, cfg, g < G ? g : 0).cs->active;
            const CapPtrs c = cap_ptrs(smem, cfg, on ? g : 0);
            FskCapShared *cs = c.cs;
            const int my_nin = on ? cs->nin : 0, nold = cs->nold;
            float tre = 0.0f, tim = 0.0f;
            int done = 0;
            for (int ck = 0; ck < nchunks; ++ck) {
                mbar_wait(&ctl->bar_int[ck], (unsigned)(par >> ck) & 1u);
                int jc = (ck + 1) * CH;
                jc = jc < my_nin ? jc : my_nin;
                const int avail = on ? outputs_ready(nold, jc, Ts, step, nint) : 0;
                /* the two product regions are contiguous: batch eight independent loads ahead of
                   the sixteen dependent adds so the chain runs at FADD latency, not LDS latency */
                while (done < avail) {
                    const bool head = done < D;
                    const float2 *src = head ? c.phead + done : c.xin + (done - D);
                    const int lim = head ? (avail < D ? avail : D) : avail;
                    int n = lim - done;
                    for (; n >= 8; n -= 8, src += 8, done += 8) {
                        float2 v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = src[k];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            tre += v[k].x;
                            tim += v[k].y;
                        }
                    }
                    for (; n > 0; --n, ++src, ++done) {
                        const float2 v = *src;
                        tre += v.x;
                        tim += v.y;
                    }
                }
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
        } else if (is_worker) {
            /* workers: A8 mix and A9 integrate, chunk by chunk behind the oscillator warp.
               Lanes wt .. wt+M-1 share one (capture, slot): one tone each. */
            const int wt = wtid;
            const int m = wt & (M - 1);
            const int q = wt / M;
            const int g = q % G, slot = q / G;
            const int nslots = (nwt / M) / G; /* slots per capture; threads beyond G*nslots*M idle */
            const bool have_slot = slot < nslots;
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            const bool on = have_slot && c.cs->active;
            const int nin = c.cs->nin, nold = c.cs->nold;
            float2 *ftone = c.fdc + m * TS;
            const unsigned lane_base = lane & ~(M - 1);
            for (int ck = 0; ck < nchunks; ++ck) {
                const int j0 = ck * CH;
                const int jend = j0 + CH < max_nin ? j0 + CH : max_nin;
                WMARK(0);
                mbar_wait(&ctl->bar_osc[ck], (unsigned)(par >> ck) & 1u);
                WMARK(1);
                /* mix: f_dc = in * conj(phi), in place over the oscillator values */
                if (on) {
                    const int je = jend < nin ? jend : nin;
                    for (int j = j0 + slot; j < je; j += nslots) {
                        const float2 x = c.xin[j];
                        const float2 o = ftone[nold + j];
                        ftone[nold + j] = make_float2(x.x * o.x + x.y * o.y, x.y * o.x - x.x * o.y);
                    }
                }
                WMARK(2);
                named_bar_sync(1, nwt);
                WMARK(3);
                /* integrate every output whose window became complete with this chunk, form
                   sum_m |f_int|^2 and multiply by the timing oscillator */
                const int lo = ck == 0 ? 0 : outputs_ready(nold, j0 < nin ? j0 : nin, Ts, step, nint);
                const int hi = on ? outputs_ready(nold, jend < nin ? jend : nin, Ts, step, nint) : 0;
                if (step == 1) {
                    /* four adjacent windows per thread share their 16-byte loads; every window is
                       still accumulated in ascending sample order from zero */
                    for (int e = (lo & ~1) + 4 * slot; __any_sync(0xffffffffu, e < hi); e += 4 * nslots) {
                        float en[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                        float2 phv[4];
                        if (m == 0 && e < hi) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) phv[k] = e + k < nint ? __ldg(phi_ft + e + k) : make_float2(0.0f, 0.0f);
                        }
                        if (e < hi) {
                            const float2 *s2 = ftone + e;
                            float2 a0 = make_float2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0;
                            /* head: samples 0..3 enter the windows one by one */
                            {
                                const float4 v = *reinterpret_cast<const float4 *>(s2);
                                const float4 w = *reinterpret_cast<const float4 *>(s2 + 2);
                                a0 = cadd(a0, make_float2(v.x, v.y));
                                a0 = cadd(a0, make_float2(v.z, v.w));
                                a1 = cadd(a1, make_float2(v.z, v.w));
                                a0 = cadd(a0, make_float2(w.x, w.y));
                                a1 = cadd(a1, make_float2(w.x, w.y));
                                a2 = cadd(a2, make_float2(w.x, w.y));
                                a0 = cadd(a0, make_float2(w.z, w.w));
                                a1 = cadd(a1, make_float2(w.z, w.w));
                                a2 = cadd(a2, make_float2(w.z, w.w));
                                a3 = cadd(a3, make_float2(w.z, w.w));
                            }
                            int r = 4;
#pragma unroll 2
                            for (; r + 1 < Ts; r += 2) {
                                const float4 v = *reinterpret_cast<const float4 *>(s2 + r);
                                a0 = cadd(a0, make_float2(v.x, v.y));
                                a1 = cadd(a1, make_float2(v.x, v.y));
                                a2 = cadd(a2, make_float2(v.x, v.y));
                                a3 = cadd(a3, make_float2(v.x, v.y));
                                a0 = cadd(a0, make_float2(v.z, v.w));
                                a1 = cadd(a1, make_float2(v.z, v.w));
                                a2 = cadd(a2, make_float2(v.z, v.w));
                                a3 = cadd(a3, make_float2(v.z, v.w));
                            }
                            /* tail: samples r .. Ts+2 leave the windows one by one */
                            for (int t = r; t < Ts + 3; ++t) {
                                const float2 x = s2[t];
                                if (t < Ts) a0 = cadd(a0, x);
                                if (t < Ts + 1) a1 = cadd(a1, x);
                                if (t < Ts + 2) a2 = cadd(a2, x);
                                a3 = cadd(a3, x);
                            }
                            en[0] = (a0.x * a0.x) + (a0.y * a0.y);
                            en[1] = (a1.x * a1.x) + (a1.y * a1.y);
                            en[2] = (a2.x * a2.x) + (a2.y * a2.y);
                            en[3] = (a3.x * a3.x) + (a3.y * a3.y);
                        }
                        /* sum the tone energies in tone order on the lane of tone 0 */
                        float ft[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                        for (int mm = 0; mm < M; ++mm) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) ft[k] += __shfl_sync(0xffffffffu, en[k], lane_base + mm);
                        }
                        if (m == 0 && e < hi) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int o = e + k;
                                if (o >= lo && o < hi) *prod_slot(c, o, D) = make_float2(ft[k] * phv[k].x, ft[k] * phv[k].y);
                            }
                        }
                    }
                } else {
                    for (int i = lo + slot; __any_sync(0xffffffffu, i < hi); i += nslots) {
                        float en = 0.0f;
                        float2 ph = make_float2(0.0f, 0.0f);
                        if (m == 0 && i < hi) ph = __ldg(phi_ft + i);
                        if (i < hi) {
                            const float2 a = window_sum(ftone + i * step, Ts);
                            en = (a.x * a.x) + (a.y * a.y);
                        }
                        float ft = 0.0f;
                        for (int mm = 0; mm < M; ++mm) ft += __shfl_sync(0xffffffffu, en, lane_base + mm);
                        if (m == 0 && i < hi) *prod_slot(c, i, D) = make_float2(ft * ph.x, ft * ph.y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->bar_int[ck]);
                WMARK(4);
            }
        } else if (is_front) {
            /* front team: frequency estimate of the NEXT frame from its TMA-staged raw samples */
            for (int g = 0; g < G; ++g) {
                FskCapShared *cs = cap_ptrs(smem, cfg, g).cs;
                if (cs->fon && cs->stg_valid) mbar_wait(&ctl->bar_in[g], (par_in >> g) & 1u);
            }
            const Team ft = { nfront, front_idx, lane, 2 };
            front_stage<G, false>(cfg, smem, tb, ft, slot ^ 1, warp);
        }
        par ^= nchunks >= 64 ? ~0ull : ((1ull << nchunks) - 1ull);
        PHASE_MARK(4); /* own role's time in the streaming phase */
        __syncthreads();
        PHASE_MARK(5); /* waiting for the slowest role */

        /* ---- A11: resample at the timing instant, decide, per-symbol statistics ----------
           lanes t .. t+M-1 share one (capture, symbol): one tone each */
        for (int it = tid; __any_sync(0xffffffffu, it < G * Nsym * M); it += nthreads) {
            const int m = it & (M - 1), q = it / M;
            const int g = q % G, i = q / G;
            const bool valid = it < G * Nsym * M;
            const CapPtrs c = cap_ptrs(smem, cfg, valid ? g : 0);
            FskCapShared *cs = c.cs;
            const bool on = valid && cs->active && !cs->nan_flag;
            float tm = 0.0f;
            if (on) {
                /* the two integrator outputs around the timing instant, re-summed from f_dc in
                   the same ascending order as the streaming pass */
                const float fract = cs->fract, omf = cs->one_minus_fract;
                const int st = (i + 1) * P;
                const int olo = (st + cs->low) * step, ohi = (st + cs->high) * step;
                float2 a, b;
                window_sum2(c.fdc + m * TS + olo, c.fdc + m * TS + ohi, Ts, a, b);
                float2 t = make_float2(omf * a.x, omf * a.y);
                t = cadd(t, make_float2(fract * b.x, fract * b.y));
                tm = (t.x * t.x) + (t.y * t.y);
            }
            float tmax[4];
            const unsigned lb = lane & ~(M - 1);
            for (int mm = 0; mm < M; ++mm) tmax[mm] = __shfl_sync(0xffffffffu, tm, lb + mm);
            if (on) {
                const size_t cap = cap0 + g;
                const size_t frame = (size_t)cs->frame;
                float *sig = c.mag + Ndft, *nse = sig + Nsym, *sqm = nse + Nsym, *rfl = sqm + Nsym;
                const float rf = sqrtf(tm);
                rfl[m * Nsym + i] = rf; /* tone-major, as rx_filt */
                if (args.rx_filt) args.rx_filt[(cap * args.max_frames + frame) * (size_t)(M * Nsym) + m * Nsym + i] = rf;
                if (m == 0) {
                    float mx = tmax[0];
                    int sym = 0;
                    for (int mm = 1; mm < M; ++mm)
                        if (tmax[mm] > mx) {
                            mx = tmax[mm];
                            sym = mm;
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
                    for (int mm = 0; mm < M; ++mm) sum += tmax[mm];
                    sig[i] = mx;
                    nse[i] = (sum - mx) / (float)(M - 1);
                    sqm[i] = sqrtf(mx);
                }
            }
        }
        if (args.bits) {
            /* reference returns early on a NaN timing estimate and the CLI re-emits its bit
               buffer: repeat the previous frame's bits (zeros on the first frame) */
            for (int g = 0; g < G; ++g) {
                FskCapShared *cs = cap_ptrs(smem, cfg, g).cs;
                if (!cs->active || !cs->nan_flag) continue;
                unsigned char *bo = args.bits + ((cap0 + g) * args.max_frames + (size_t)cs->frame) * (size_t)Nbits;
                for (int i = tid; i < Nbits; i += nthreads) bo[i] = cs->frame ? bo[i - Nbits] : 0;
            }
        }
        __syncthreads();
        PHASE_MARK(6);
        if (warp == 0) {
            /* four serial sums over the symbols per capture, lane (g, s) */
            const int g = lane >> 2, s = lane & 3;
            const bool on = g < G && cap_ptrs(smem, cfg, g < G ? g : 0).cs->active &&
                            !cap_ptrs(smem, cfg, g < G ? g : 0).cs->nan_flag;
            const CapPtrs c = cap_ptrs(smem, cfg, on ? g : 0);
            FskCapShared *cs = c.cs;
            if (on) {
                const float *sig = c.mag + Ndft;
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
                const float *fe = cfg.est_type ? cs->f2_est[slot] : cs->f_est[slot];
                const float fc_avg = (fe[0] + fe[1]) / 2.0f;
                const float fc_tx = (float)((cfg.f1_tx + cfg.f1_tx + cfg.tone_spacing) / 2);
                cs->foff = fc_tx - fc_avg;
            }
        }
        __syncthreads();
        PHASE_MARK(7);

        /* ---- A12: soft decisions ([U mpdecode_core.c::fsk_rx_filt_to_llrs]) ----------------- */
        if (args.llr) {
            const int bps = M == 2 ? 1 : 2;
            for (int it = tid; it < G * Nsym; it += nthreads) {
                const int g = it % G, i = it / G;
                const CapPtrs c = cap_ptrs(smem, cfg, g);
                FskCapShared *cs = c.cs;
                if (!cs->active || cs->nan_flag) continue;
                const float *rfl = c.mag + Ndft + 3 * Nsym;
                float l2[2];
                symbol_llrs(rfl + i, Nsym, M, cs->v_est, cs->SNRest, l2);
                float *lo = args.llr + ((cap0 + g) * args.max_frames + (size_t)cs->frame) * (size_t)Nbits;
                for (int k = 0; k < bps; ++k) lo[bps * i + k] = l2[k];
            }
            __syncthreads();
        }

        /* ---- per-frame record, advance, next frame's bookkeeping ---------------------------- */
        if (tid == 0) {
            for (int g = 0; g < G; ++g) {
                FskCapShared *cs = cap_ptrs(smem, cfg, g).cs;
                if (!cs->active) continue;
                if (args.stats) {
                    pirip_frame_stats st;
                    for (int m = 0; m < 4; ++m) {
                        st.f_est[m] = m < M ? cs->f_est[slot][m] : 0.0f;
                        st.f2_est[m] = m < M ? cs->f2_est[slot][m] : 0.0f;
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
                    args.stats[(cap0 + g) * args.max_frames + (size_t)cs->frame] = st;
                }
                cs->pos += (unsigned long long)cs->nin;
                cs->nin = cs->nin_next;
                cs->frame += 1;
            }
            bookkeeping();
        }
        __syncthreads();
        if (pipelined) slot ^= 1;
        PHASE_MARK(8);
    }
#ifdef FSK_PHASE_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == tim_warp || warp == osc_warp) && args.Sf_out == nullptr &&
        args.stats == nullptr)
        printf("w%d worker: wait-osc %lld mix %lld bar %lld integrate %lld\n", warp, wph[1], wph[2], wph[3], wph[4]);
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == tim_warp || warp == osc_warp) && args.Sf_out == nullptr &&
        args.stats == nullptr)
        printf("w%d load %lld fft %lld iir+peaks %lld stream-own %lld stream-wait %lld decide %lld sums %lld adv %lld book %lld\n",
               warp, tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7], tph[8], tph[0]);
#endif

    /* ---- results and state ------------------------------------------------------------ */
    for (int g = 0; g < G; ++g) {
        const size_t cap = cap0 + g;
        if (cap >= args.n_captures) break;
        const CapPtrs c = cap_ptrs(smem, cfg, g);
        /* drain a staged copy that will not be consumed before the CTA retires */
        if (c.cs->stg_valid) mbar_wait(&ctl->bar_in[g], (par_in >> g) & 1u);
        if (args.nframes && tid == 0) args.nframes[cap] = (int32_t)c.cs->frame;
        if (args.Sf_out)
            for (int i = tid; i < Ndft; i += nthreads) args.Sf_out[cap * (size_t)Ndft + i] = c.Sf[i];
        if (args.state) {
            float *state = args.state + cap * args.state_stride_floats;
            FskDevStateHdr *h = reinterpret_cast<FskDevStateHdr *>(state);
            float *sSf = state + sizeof(FskDevStateHdr) / sizeof(float);
            float2 *sfdc = reinterpret_cast<float2 *>(sSf + Ndft);
            for (int i = tid; i < Ndft; i += nthreads) sSf[i] = c.Sf[i];
            for (int m = 0; m < M; ++m)
                for (int i = tid; i < Nmem; i += nthreads) sfdc[m * Nmem + i] = c.fdc[m * TS + i];
            if (tid == 0) {
                for (int m = 0; m < 4; ++m) h->phi_c[m] = c.cs->phi_c[m];
                h->norm_rx_timing = c.cs->norm_rx_timing;
                h->ppm = c.cs->ppm;
                h->snr_est = c.cs->snr_est;
                h->EbNodB = c.cs->EbNodB;
                h->nin = c.cs->nin;
            }
        }
    }
}

template <int G>
cudaError_t prepare_g(const FskDevCfg &cfg) {
    return cudaFuncSetAttribute(fsk_demod_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
}

template <int G>
cudaError_t launch_g(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream) {
    const unsigned grid = (unsigned)((n_captures + G - 1) / G);
    fsk_demod_kernel<G><<<grid, cfg.threads, cfg.smem_bytes, stream>>>(cfg, args);
    return cudaGetLastError();
}

}  // namespace

cudaError_t fsk_kernel_prepare(const FskDevCfg &cfg) {
    switch (cfg.G) {
        case 1: return prepare_g<1>(cfg);
        case 2: return prepare_g<2>(cfg);
        case 3: return prepare_g<3>(cfg);
        case 4: return prepare_g<4>(cfg);
        case 5: return prepare_g<5>(cfg);
        case 6: return prepare_g<6>(cfg);
        case 8: return prepare_g<8>(cfg);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t fsk_launch_demod(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream) {
    if (n_captures == 0) return cudaSuccess;
    switch (cfg.G) {
        case 1: return launch_g<1>(cfg, args, n_captures, stream);
        case 2: return launch_g<2>(cfg, args, n_captures, stream);
        case 3: return launch_g<3>(cfg, args, n_captures, stream);
        case 4: return launch_g<4>(cfg, args, n_captures, stream);
        case 5: return launch_g<5>(cfg, args, n_captures, stream);
        case 6: return launch_g<6>(cfg, args, n_captures, stream);
        case 8: return launch_g<8>(cfg, args, n_captures, stream);
        default: return cudaErrorInvalidValue;
    }
}
