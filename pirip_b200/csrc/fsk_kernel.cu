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
 * in and the bits out.
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
 * at once, and (b) the frame is a streaming pipeline: warp 0 spins the oscillators and
 * publishes its progress, the worker warps mix and integrate chunk by chunk right behind
 * it, and warp 1 accumulates the timing line right behind them, so everything downstream
 * of the oscillator hides under its latency.  FFTs, IIR, peak search, decisions and LLRs
 * are ordinary data-parallel phases between CTA barriers.
 */
#include "exact_math.h"
#include "fsk_dev.h"

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    float2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_volatile(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
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
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

/* FFT scratch index swizzle: XOR bits [5:4] of the element index into bits [3:2] and
   [1:0], so the radix-4 stages with m = 1 and m = 4 (16 lanes, 8-byte elements) touch 16
   distinct bank pairs; unit-stride stages are unaffected (constant XOR per 16-block) */
__device__ __forceinline__ int fswz(int e) { return e ^ (5 * ((e >> 4) & 3)); }

/* magnitude of the fft-shifted spectrum, written as N floats to out */
template <int PER>
__device__ __forceinline__ void fft_shift_mag(const float2 *buf, float *out, int N, int lane) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = lane + (q << 5);
        const float2 X = buf[fswz((i + (N >> 1)) & (N - 1))];
        out[i] = sqrtf((X.x * X.x) + (X.y * X.y));
    }
}

/* One warp: Hann-window Ndft samples of x, FFT them (kiss_fft decimation-in-time factor
   order 4,4,..,(2)) in the warp's scratch slot, write the fft-shifted magnitudes to out. */
__device__ void warp_fft_mag(const FskDevCfg &cfg, const float2 *x, float2 *buf, float *out, int lane) {
    const int N = cfg.Ndft;
    const float2 *tw = cfg.twiddle; /* read through L1: shared by every warp on the SM */
    for (int n = lane; n < N; n += 32) {
        const float h = __ldg(cfg.hann + n);
        const float2 v = x[n];
        float2 w;
        w.x = h * v.x;
        w.y = h * v.y;
        int t = n, pos = 0;
        for (int s = 0; s < cfg.nfac; ++s) {
            const int p = cfg.fac_p[s];
            const int d = t & (p - 1);
            t = (p == 4) ? (t >> 2) : (t >> 1);
            pos += d << cfg.fac_log2m[s];
        }
        buf[fswz(pos)] = w;
    }
    __syncwarp();
    for (int s = cfg.nfac - 1; s >= 0; --s) {
        const int p = cfg.fac_p[s], m = cfg.fac_m[s], lm = cfg.fac_log2m[s], fs = cfg.fac_fstride[s];
        if (p == 4) {
            for (int b = lane; b < (N >> 2); b += 32) {
                const int blk = b >> lm, k = b & (m - 1);
                const int base = ((blk << 2) << lm) + k;
                const int i0 = fswz(base), i1 = fswz(base + m), i2 = fswz(base + 2 * m), i3 = fswz(base + 3 * m);
                float2 F0 = buf[i0];
                const float2 s0 = cmul(buf[i1], __ldg(tw + k * fs));
                const float2 s1 = cmul(buf[i2], __ldg(tw + 2 * k * fs));
                const float2 s2 = cmul(buf[i3], __ldg(tw + 3 * k * fs));
                const float2 s5 = csub(F0, s1);
                F0 = cadd(F0, s1);
                const float2 s3 = cadd(s0, s2);
                const float2 s4 = csub(s0, s2);
                buf[i2] = csub(F0, s3);
                buf[i0] = cadd(F0, s3);
                buf[i1] = make_float2(s5.x + s4.y, s5.y - s4.x);
                buf[i3] = make_float2(s5.x - s4.y, s5.y + s4.x);
            }
        } else {
            for (int b = lane; b < (N >> 1); b += 32) {
                const int blk = b >> lm, k = b & (m - 1);
                const int base = ((blk << 1) << lm) + k;
                const int i0 = fswz(base), i1 = fswz(base + m);
                const float2 t = cmul(buf[i1], __ldg(tw + k * fs));
                const float2 F0 = buf[i0];
                buf[i1] = csub(F0, t);
                buf[i0] = cadd(F0, t);
            }
        }
        __syncwarp();
    }
    switch (N >> 5) {
        case 1: fft_shift_mag<1>(buf, out, N, lane); break;
        case 2: fft_shift_mag<2>(buf, out, N, lane); break;
        case 4: fft_shift_mag<4>(buf, out, N, lane); break;
        case 8: fft_shift_mag<8>(buf, out, N, lane); break;
        case 16: fft_shift_mag<16>(buf, out, N, lane); break;
        default: fft_shift_mag<32>(buf, out, N, lane); break;
    }
    __syncwarp();
}

/* log I0(x), Numerical Recipes polynomial in the log domain ([U mpdecode_core.c::bessi0]) */
__device__ float logbessi0(float x) {
    const float ax = fabsf(x);
    double y, ans;
    if ((double)ax < 3.75) {
        y = (double)x / 3.75;
        y *= y;
        ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
        return (float)em_log(ans);
    }
    y = 3.75 / (double)ax;
    ans = 0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2)))))));
    return (float)(((double)ax - 0.5 * em_log((double)ax)) + em_log(ans));
}

/* linear-log-MAP max* ([U mpdecode_core.c::max_star0]) */
__device__ float max_star0(float delta1, float delta2) {
    const double AJIAN = -0.24904163195436, TJIAN = 2.50681740420944;
    const float diff = delta2 - delta1;
    if ((double)diff > TJIAN) return delta2;
    if ((double)diff < -TJIAN) return delta1;
    if (diff > 0) return (float)((double)delta2 + AJIAN * ((double)diff - TJIAN));
    return (float)((double)delta1 - AJIAN * ((double)diff + TJIAN));
}

/* u8 -> float exactly as ((float)b - 127.0)/128.0 */
__device__ __forceinline__ float u8f(unsigned b) { return (float)((int)b - 127) * 0.0078125f; }

/* number of integrator outputs whose whole window has been mixed once `j` new samples of
   this frame are in place (window i covers f_dc[i*step, i*step+Ts), new samples start at nold) */
__device__ __forceinline__ int outputs_ready(int nold, int j, int Ts, int step, int nint) {
    const int have = nold + j - Ts;
    if (have < 0) return 0;
    const int r = have / step + 1;
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

struct CapPtrs {
    float2 *fdc;
    float2 *xin; /* frame input; later the timing-line products */
    float *mag;  /* FFT magnitudes; later tmpS | sig | nse | sqm | rx_filt */
    float *Sf;
    FskCapShared *cs;
};

__device__ __forceinline__ CapPtrs cap_ptrs(unsigned char *smem, const FskDevCfg &cfg, int g) {
    unsigned char *cb = smem + (size_t)g * cfg.cap_stride;
    CapPtrs p;
    p.fdc = reinterpret_cast<float2 *>(cb + cfg.off_fdc);
    p.xin = reinterpret_cast<float2 *>(cb + cfg.off_xin);
    p.mag = reinterpret_cast<float *>(cb + cfg.off_mag);
    p.Sf = reinterpret_cast<float *>(cb + cfg.off_Sf);
    p.cs = reinterpret_cast<FskCapShared *>(cb + cfg.off_scal);
    return p;
}

/* Timing-line product of integrator output e.  Output e becomes ready nold - Ts + 1 samples
   ahead of the input index, so the products live in the consumed part of the input buffer
   shifted back by D = Ts + Ts/4 + 1 (>= nold_max - Ts + 1); the first D go to the idle
   magnitude region. */
__device__ __forceinline__ float2 *prod_slot(const CapPtrs &c, int e, int D) {
    return e < D ? reinterpret_cast<float2 *>(c.mag) + e : c.xin + (e - D);
}

template <int G>
__global__ void __launch_bounds__(512, 1) fsk_demod_kernel(const __grid_constant__ FskDevCfg cfg,
                                                           const __grid_constant__ FskBatchArgs args) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int M = cfg.M, Nmem = cfg.Nmem, Ndft = cfg.Ndft, nint = cfg.nint, Nsym = cfg.Nsym, P = cfg.P, Ts = cfg.Ts;
    const int Nbits = cfg.Nbits, TS = cfg.fdc_tone_stride, CH = cfg.chunk, step = Ts / P;
    const size_t cap0 = (size_t)blockIdx.x * G;
    FskCtaShared *ctl = reinterpret_cast<FskCtaShared *>(smem + (size_t)G * cfg.cap_stride + cfg.off_ctl);
    float2 *fft_slot = reinterpret_cast<float2 *>(smem + (size_t)G * cfg.cap_stride + cfg.off_fft) + (size_t)warp * Ndft;
    const float tc = cfg.tc;
    const float omtc = 1.0f - tc;
    const int D = Ts + Ts / 4 + 1;
    unsigned long long par = 0; /* phase parity of every per-chunk mbarrier, bit ck */
    if (tid == 0) {
        for (int i = 0; i < FSK_MAX_CHUNKS; ++i) {
            mbar_init(&ctl->bar_osc[i], 1);
            mbar_init(&ctl->bar_int[i], nwarps - 2);
        }
    }

    /* ---- initial state ------------------------------------------------------------ */
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
                c.cs->f_est[m] = 0.0f;
                c.cs->f2_est[m] = 0.0f;
            }
            c.cs->SNRest = c.cs->v_est = c.cs->rx_sig_pow = c.cs->rx_nse_pow = c.cs->foff = c.cs->rx_timing = 0.0f;
            c.cs->pos = 0;
            c.cs->frame = 0;
            c.cs->active = exists ? 1 : 0;
            c.cs->nan_flag = 0;
        }
    }
    __syncthreads();

    for (;;) {
        /* ---- frame bookkeeping: which captures still have a whole frame to read ------- */
        if (tid == 0) {
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
                if (act) {
                    any = 1;
                    mx = cs->nin > mx ? cs->nin : mx;
                }
            }
            ctl->any_active = any;
            ctl->max_nin = mx;
        }
        __syncthreads();
        if (!ctl->any_active) break;
        const int max_nin = ctl->max_nin;

        /* ---- A5: input conversion ([U fsk_demod.c]) + A8: slide the f_dc memory ---------- */
        for (int g = 0; g < G; ++g) {
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            if (!c.cs->active) continue;
            const int nin = c.cs->nin, nold = c.cs->nold;
            const unsigned char *iq = args.iq + (cap0 + g) * args.capture_stride_bytes;
            if (cfg.fmt == PIRIP_FMT_U8) {
                const unsigned char *src = iq + 2 * c.cs->pos;
                /* 4-byte loads (two samples) over the aligned body, single samples at the ends */
                const int mis = (int)((reinterpret_cast<size_t>(src) >> 1) & 1);
                if (mis && tid == 0) c.xin[0] = make_float2(u8f(src[0]), u8f(src[1]));
                const int npair = (nin - mis) >> 1;
                const unsigned *s4 = reinterpret_cast<const unsigned *>(src + 2 * mis);
                for (int k = tid; k < npair; k += nthreads) {
                    const unsigned w = __ldg(s4 + k);
                    float2 *dst = c.xin + mis + 2 * k;
                    dst[0] = make_float2(u8f(w & 0xff), u8f((w >> 8) & 0xff));
                    dst[1] = make_float2(u8f((w >> 16) & 0xff), u8f(w >> 24));
                }
                if (((nin - mis) & 1) && tid == 32) {
                    const int j = nin - 1;
                    c.xin[j] = make_float2(u8f(src[2 * j]), u8f(src[2 * j + 1]));
                }
            } else if (cfg.fmt == PIRIP_FMT_S16) {
                const short *src = reinterpret_cast<const short *>(iq) + 2 * c.cs->pos;
                float *xf = reinterpret_cast<float *>(c.xin);
                for (int i = tid; i < 2 * nin; i += nthreads) xf[i] = (float)src[i] / 825.0f; /* FDMDV_SCALE */
            } else {
                const float2 *src = reinterpret_cast<const float2 *>(iq) + c.cs->pos;
                for (int i = tid; i < nin; i += nthreads) c.xin[i] = src[i];
            }
            if (nold <= nin) {
                for (int m = 0; m < M; ++m)
                    for (int i = tid; i < nold; i += nthreads) c.fdc[m * TS + i] = c.fdc[m * TS + i + nin];
            } else if (tid < M) {
                for (int i = 0; i < nold; ++i) c.fdc[tid * TS + i] = c.fdc[tid * TS + i + nin];
            }
        }
        __syncthreads();

        /* ---- A7: tone frequency estimate --------------------------------------------- */
        for (int job = warp; job < G * cfg.max_numffts; job += nwarps) {
            const int g = job / cfg.max_numffts, j = job - g * cfg.max_numffts;
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            if (!c.cs->active || j >= c.cs->numffts) continue;
            warp_fft_mag(cfg, c.xin + j * (Ndft >> 1), fft_slot, c.mag + (size_t)j * Ndft, lane);
        }
        __syncthreads();
        for (int g = 0; g < G; ++g) {
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            if (!c.cs->active) continue;
            const int numffts = c.cs->numffts;
            for (int i = tid; i < Ndft; i += nthreads) {
                float s = c.Sf[i];
                for (int j = 0; j < numffts; ++j) s = (s * omtc) + (c.mag[j * Ndft + i] * tc);
                c.Sf[i] = s;
                c.mag[i] = s; /* tmpS: scratch copy the peak search blanks */
            }
        }
        __syncthreads();
        for (int g = warp; g < G; g += nwarps) {
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            if (!c.cs->active) continue;
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
                for (int m = 0; m < M; ++m) c.cs->f_est[m] = (float)fi[m] * cfg.fs_over_ndft;
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
                    for (int m = 0; m < M; ++m) c.cs->f2_est[m] = foff + (float)(m * cfg.tone_spacing);
                }
            }
            __syncwarp();
            if (lane < M) {
                const float f = cfg.est_type ? c.cs->f2_est[lane] : c.cs->f_est[lane];
                const float arg = (float)(EM_TWO_PI * (double)(f / (float)cfg.Fs));
                double sn, csn;
                em_sincos((double)arg, &sn, &csn);
                c.cs->dphi[lane] = make_float2((float)csn, (float)sn);
            }
        }
        __syncthreads();

        /* ================= streaming phase: oscillators -> mix -> integrate -> timing line ===== */
        const int nchunks = (max_nin + CH - 1) / CH;
        if (warp == 0) {
            /* A8: the oscillator recursion phi *= dphi is serial; lane (g, m) carries tone m of
               capture g, all lanes step together and publish progress once per chunk */
            const int g = lane / M, m = lane - g * M;
            const bool on = lane < G * M && cap_ptrs(smem, cfg, g < G ? g : 0).cs->active;
            const CapPtrs c = cap_ptrs(smem, cfg, on ? g : 0);
            const int my_nin = on ? c.cs->nin : 0;
            const float2 d = on ? c.cs->dphi[m] : make_float2(1.0f, 0.0f);
            float2 ph = on ? c.cs->phi_c[m] : make_float2(1.0f, 0.0f);
            float2 *dst = c.fdc + m * TS + c.cs->nold;
            /* smallest nin over the active lanes: below it no lane needs a bounds check */
            int lo = on ? my_nin : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int o = __shfl_xor_sync(0xffffffffu, lo, off);
                lo = o < lo ? o : lo;
            }
            for (int j0 = 0; j0 < max_nin; j0 += CH) {
                const int jend = j0 + CH < max_nin ? j0 + CH : max_nin;
                if (jend <= lo) {
                    if (on) {
#pragma unroll 8
                        for (int j = j0; j < jend; ++j) {
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
                if (lane == 0) mbar_arrive(&ctl->bar_osc[j0 / CH]);
            }
            if (on) {
                const float av = sqrtf((ph.x * ph.x) + (ph.y * ph.y));
                c.cs->phi_c[m] = make_float2(ph.x / av, ph.y / av);
            }
        } else if (warp == 1) {
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
                for (; done < avail; ++done) {
                    const float2 p = *prod_slot(c, done, D);
                    tre += p.x;
                    tim += p.y;
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
        } else {
            /* workers: A8 mix and A9 integrate, chunk by chunk behind the oscillator warp */
            const int wt = tid - 64, nwt = nthreads - 64;
            const int lch = 31 - __clz(CH); /* CH is a power of two */
            for (int ck = 0; ck < nchunks; ++ck) {
                const int j0 = ck * CH;
                const int jend = j0 + CH < max_nin ? j0 + CH : max_nin;
                mbar_wait(&ctl->bar_osc[ck], (unsigned)(par >> ck) & 1u);
                /* mix: f_dc = in * conj(phi), in place over the oscillator values */
                for (int it = wt; it < (G << lch); it += nwt) {
                    const int g = it >> lch, j = j0 + (it & (CH - 1));
                    const CapPtrs c = cap_ptrs(smem, cfg, g);
                    if (!c.cs->active || j >= c.cs->nin) continue;
                    const float2 x = c.xin[j];
                    float2 *f = c.fdc + c.cs->nold + j;
                    for (int m = 0; m < M; ++m) {
                        const float2 o = f[m * TS];
                        f[m * TS] = make_float2(x.x * o.x + x.y * o.y, x.y * o.x - x.x * o.y);
                    }
                }
                named_bar_sync(1, nwt);
                /* integrate every output whose window became complete with this chunk, form
                   sum_m |f_int|^2 and multiply by the timing oscillator */
                /* the first chunk also owns the outputs that lie entirely in the old samples */
                const int extra = ck == 0 ? (Ts + Ts / 4 + 1) : 0;
                if (step == 1) {
                    const int max_pairs = ((CH + extra) >> 1) + 2;
                    for (int it = wt; it < max_pairs * G; it += nwt) {
                        const int g = it % G, pr = it / G;
                        const CapPtrs c = cap_ptrs(smem, cfg, g);
                        if (!c.cs->active) continue;
                        const int nin = c.cs->nin, nold = c.cs->nold;
                        const int lo = ck == 0 ? 0 : outputs_ready(nold, j0 < nin ? j0 : nin, Ts, 1, nint);
                        const int hi = outputs_ready(nold, jend < nin ? jend : nin, Ts, 1, nint);
                        const int e = (lo & ~1) + 2 * pr;
                        if (e >= hi) continue;
                        float ft0 = 0.0f, ft1 = 0.0f;
                        for (int m = 0; m < M; ++m) {
                            const float4 *src = reinterpret_cast<const float4 *>(c.fdc + m * TS + e);
                            float4 v = src[0];
                            float2 a0 = make_float2(0.0f + v.x, 0.0f + v.y);
                            a0 = cadd(a0, make_float2(v.z, v.w));
                            float2 a1 = make_float2(0.0f + v.z, 0.0f + v.w);
                            int r = 2;
#pragma unroll 4
                            for (; r + 1 < Ts; r += 2) {
                                v = src[r >> 1];
                                a0 = cadd(a0, make_float2(v.x, v.y));
                                a1 = cadd(a1, make_float2(v.x, v.y));
                                a0 = cadd(a0, make_float2(v.z, v.w));
                                a1 = cadd(a1, make_float2(v.z, v.w));
                            }
                            v = src[r >> 1];
                            if (r == Ts) { /* Ts even */
                                a1 = cadd(a1, make_float2(v.x, v.y));
                            } else {
                                a0 = cadd(a0, make_float2(v.x, v.y));
                                a1 = cadd(a1, make_float2(v.x, v.y));
                                a1 = cadd(a1, make_float2(v.z, v.w));
                            }
                            ft0 += (a0.x * a0.x) + (a0.y * a0.y);
                            ft1 += (a1.x * a1.x) + (a1.y * a1.y);
                        }
                        if (e >= lo) {
                            const float2 ph = __ldg(cfg.phi_ft + e);
                            *prod_slot(c, e, D) = make_float2(ft0 * ph.x, ft0 * ph.y);
                        }
                        if (e + 1 < hi) {
                            const float2 ph = __ldg(cfg.phi_ft + e + 1);
                            *prod_slot(c, e + 1, D) = make_float2(ft1 * ph.x, ft1 * ph.y);
                        }
                    }
                } else {
                    const int max_out = (CH + extra) / step + 2;
                    for (int it = wt; it < max_out * G; it += nwt) {
                        const int g = it % G, k = it / G;
                        const CapPtrs c = cap_ptrs(smem, cfg, g);
                        if (!c.cs->active) continue;
                        const int nin = c.cs->nin, nold = c.cs->nold;
                        const int lo = ck == 0 ? 0 : outputs_ready(nold, j0 < nin ? j0 : nin, Ts, step, nint);
                        const int hi = outputs_ready(nold, jend < nin ? jend : nin, Ts, step, nint);
                        const int i = lo + k;
                        if (i >= hi) continue;
                        float ft = 0.0f;
                        for (int m = 0; m < M; ++m) {
                            const float2 a = window_sum(c.fdc + m * TS + i * step, Ts);
                            ft += (a.x * a.x) + (a.y * a.y);
                        }
                        const float2 ph = __ldg(cfg.phi_ft + i);
                        *prod_slot(c, i, D) = make_float2(ft * ph.x, ft * ph.y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->bar_int[ck]);
            }
        }
        par ^= nchunks >= 64 ? ~0ull : ((1ull << nchunks) - 1ull);
        __syncthreads();

        /* ---- A11: resample at the timing instant, decide, per-symbol statistics ---------- */
        for (int it = tid; it < G * Nsym; it += nthreads) {
            const int g = it % G, i = it / G;
            const CapPtrs c = cap_ptrs(smem, cfg, g);
            FskCapShared *cs = c.cs;
            if (!cs->active || cs->nan_flag) continue;
            const size_t cap = cap0 + g;
            const size_t frame = (size_t)cs->frame;
            float *sig = c.mag + Ndft, *nse = sig + Nsym, *sqm = nse + Nsym, *rfl = sqm + Nsym;
            const float fract = cs->fract, omf = cs->one_minus_fract;
            const int st = (i + 1) * P;
            const int olo = (st + cs->low) * step, ohi = (st + cs->high) * step;
            float tmax[4];
            for (int m = 0; m < M; ++m) {
                /* the two integrator outputs around the timing instant, re-summed from f_dc in
                   the same ascending order as the streaming pass */
                const float2 a = window_sum(c.fdc + m * TS + olo, Ts);
                const float2 b = (ohi == olo) ? a : window_sum(c.fdc + m * TS + ohi, Ts);
                float2 t = make_float2(omf * a.x, omf * a.y);
                t = cadd(t, make_float2(fract * b.x, fract * b.y));
                tmax[m] = (t.x * t.x) + (t.y * t.y);
            }
            float mx = tmax[0];
            int sym = 0;
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
            for (int m = 0; m < M; ++m) {
                const float rf = sqrtf(tmax[m]);
                rfl[m * Nsym + i] = rf; /* tone-major, as rx_filt */
                if (args.rx_filt) args.rx_filt[(cap * args.max_frames + frame) * (size_t)(M * Nsym) + m * Nsym + i] = rf;
                sum += tmax[m];
            }
            sig[i] = mx;
            nse[i] = (sum - mx) / (float)(M - 1);
            sqm[i] = sqrtf(mx);
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
                const float *fe = cfg.est_type ? cs->f2_est : cs->f_est;
                const float fc_avg = (fe[0] + fe[1]) / 2.0f;
                const float fc_tx = (float)((cfg.f1_tx + cfg.f1_tx + cfg.tone_spacing) / 2);
                cs->foff = fc_tx - fc_avg;
            }
        }
        __syncthreads();

        /* ---- A12: soft decisions ([U mpdecode_core.c::fsk_rx_filt_to_llrs]) ----------------- */
        if (args.llr) {
            const int bps = M == 2 ? 1 : 2;
            for (int it = tid; it < G * Nsym; it += nthreads) {
                const int g = it % G, i = it / G;
                const CapPtrs c = cap_ptrs(smem, cfg, g);
                FskCapShared *cs = c.cs;
                if (!cs->active || cs->nan_flag) continue;
                const float *rfl = c.mag + Ndft + 3 * Nsym;
                const float v_est = cs->v_est, scale = 2.0f * cs->SNRest;
                float num[2] = { -1000000.0f, -1000000.0f }, den[2] = { -1000000.0f, -1000000.0f };
                for (int m = 0; m < M; ++m) {
                    const float yr = rfl[m * Nsym + i];
                    const float env = (float)sqrt((double)(yr * yr / (v_est * v_est)));
                    const float metric = logbessi0(scale * env);
                    int mask = 1 << (bps - 1);
                    for (int k = 0; k < bps; ++k) {
                        if (mask & m)
                            num[k] = max_star0(num[k], metric);
                        else
                            den[k] = max_star0(den[k], metric);
                        mask >>= 1;
                    }
                }
                float *lo = args.llr + ((cap0 + g) * args.max_frames + (size_t)cs->frame) * (size_t)Nbits;
                for (int k = 0; k < bps; ++k) lo[bps * i + k] = -(num[k] - den[k]);
            }
            __syncthreads();
        }

        /* ---- per-frame record, advance ------------------------------------------------- */
        if (tid < G) {
            FskCapShared *cs = cap_ptrs(smem, cfg, tid).cs;
            if (cs->active) {
                if (args.stats) {
                    pirip_frame_stats s;
                    for (int m = 0; m < 4; ++m) {
                        s.f_est[m] = m < M ? cs->f_est[m] : 0.0f;
                        s.f2_est[m] = m < M ? cs->f2_est[m] : 0.0f;
                    }
                    s.norm_rx_timing = cs->norm_rx_timing;
                    s.ppm = cs->ppm;
                    s.SNRest = cs->SNRest;
                    s.v_est = cs->v_est;
                    s.rx_sig_pow = cs->rx_sig_pow;
                    s.rx_nse_pow = cs->rx_nse_pow;
                    s.EbNodB = cs->EbNodB;
                    s.snr_est = cs->snr_est;
                    s.foff = cs->foff;
                    s.rx_timing = cs->rx_timing;
                    s.nin_next = cs->nin_next;
                    s.sample_offset = (int32_t)cs->pos;
                    args.stats[(cap0 + tid) * args.max_frames + (size_t)cs->frame] = s;
                }
                cs->pos += (unsigned long long)cs->nin;
                cs->nin = cs->nin_next;
                cs->frame += 1;
            }
        }
        __syncthreads();
    }

    /* ---- results and state ------------------------------------------------------------ */
    for (int g = 0; g < G; ++g) {
        const size_t cap = cap0 + g;
        if (cap >= args.n_captures) break;
        const CapPtrs c = cap_ptrs(smem, cfg, g);
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
