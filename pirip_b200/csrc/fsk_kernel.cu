/*
 * fsk_kernel.cu — the FSK demodulator hot path on sm_100a.
 *
 * One persistent CTA per capture runs the whole `fsk_demod` main loop
 * ([U codec2 fsk_demod.c], pirip call sites [R README.md:105,109],
 * [R test/loopback_rtl_sdr.sh:16]) on the device: per frame it converts nin IQ
 * samples, estimates the tone frequencies (windowed FFT blocks + IIR spectrum +
 * peak pick / mask, [U fsk.c::fsk_demod_freq_est]), mixes every tone to DC,
 * integrates over a symbol at (Nsym+1)*P offsets, extracts the symbol-rate timing
 * line, resamples, decides bits and maps soft decisions
 * ([U fsk.c::fsk_demod_core], [U mpdecode_core.c::fsk_rx_filt_to_llrs]).
 * All demodulator state (f_dc memory, Sf, oscillator phases, nin) stays in shared
 * memory between frames; HBM traffic is the u8 IQ in and the bits out.
 *
 * NUMERICS CONTRACT.  This kernel performs the same IEEE-754 operations in the same
 * order as the reference algorithm: sequential accumulation where the reference
 * accumulates sequentially (integrators, timing line, oscillator recursion), the same
 * radix-4/2 decimation-in-time butterflies as kiss_fft, and exact_math.h for every
 * transcendental.  It must be compiled with -fmad=false -prec-div=true
 * -prec-sqrt=true -ftz=false.  The result is bit-identical output — bits AND
 * floats — to a CPU evaluation of the algorithm, which is what tests/ check.
 * Parallelism comes from (a) independent captures across CTAs and (b) the
 * independent outputs inside a frame (FFT butterflies, (tone, offset) integrator
 * outputs, symbols); only the three inherently serial recurrences (oscillator phase,
 * timing-line sum, per-frame power sums) run on single lanes.
 */
#include "exact_math.h"
#include "fsk_dev.h"

namespace {

typedef FskFrameShared FrameShared;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    float2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

/* FFT scratch is padded by one element every four so that the radix-4 stages with
   m = 1 and m = 4 do not serialise on shared-memory banks */
__device__ __forceinline__ int fpad(int i) { return i + (i >> 2); }

/* magnitude of the fft-shifted spectrum: every lane first pulls its PER bins into
   registers, then the slot is overwritten with floats */
template <int PER>
__device__ __forceinline__ void fft_shift_mag(float2 *buf, int N, int lane) {
    float mag[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = lane + (q << 5);
        const float2 X = buf[fpad((i + (N >> 1)) & (N - 1))];
        mag[q] = sqrtf((X.x * X.x) + (X.y * X.y));
    }
    __syncwarp();
    float *fm = reinterpret_cast<float *>(buf);
#pragma unroll
    for (int q = 0; q < PER; ++q) fm[lane + (q << 5)] = mag[q];
}

/* One warp: Hann-window Ndft samples of x, FFT them (kiss_fft decimation-in-time
   factor order 4,4,..,(2)), and leave the fft-shifted magnitudes as Ndft floats at the
   start of buf. */
__device__ void warp_fft_mag(const FskDevCfg &cfg, const float2 *x, float2 *buf, const float2 *tw, const float *hann,
                             int lane) {
    const int N = cfg.Ndft;
    for (int n = lane; n < N; n += 32) {
        const float h = __ldg(hann + n);
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
        buf[fpad(pos)] = w;
    }
    __syncwarp();
    for (int s = cfg.nfac - 1; s >= 0; --s) {
        const int p = cfg.fac_p[s], m = cfg.fac_m[s], lm = cfg.fac_log2m[s], fs = cfg.fac_fstride[s];
        if (p == 4) {
            for (int b = lane; b < (N >> 2); b += 32) {
                const int blk = b >> lm, k = b & (m - 1);
                const int base = ((blk << 2) << lm) + k;
                const int i0 = fpad(base), i1 = fpad(base + m), i2 = fpad(base + 2 * m), i3 = fpad(base + 3 * m);
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
                const int i0 = fpad(base), i1 = fpad(base + m);
                const float2 t = cmul(buf[i1], __ldg(tw + k * fs));
                const float2 F0 = buf[i0];
                buf[i1] = csub(F0, t);
                buf[i0] = cadd(F0, t);
            }
        }
        __syncwarp();
    }
    /* fft shift (DC -> bin N/2) and magnitude; overwrite the slot with N floats */
    switch (N >> 5) {
        case 1: fft_shift_mag<1>(buf, N, lane); break;
        case 2: fft_shift_mag<2>(buf, N, lane); break;
        case 4: fft_shift_mag<4>(buf, N, lane); break;
        case 8: fft_shift_mag<8>(buf, N, lane); break;
        case 16: fft_shift_mag<16>(buf, N, lane); break;
        default: fft_shift_mag<32>(buf, N, lane); break;
    }
}

/* log I0(x), Numerical Recipes polynomial in the log domain
   ([U mpdecode_core.c::bessi0]) */
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

template <int kThreads>
__global__ void __launch_bounds__(kThreads, FSK_MIN_CTAS) fsk_demod_kernel(const __grid_constant__ FskDevCfg cfg,
                                                             const __grid_constant__ FskBatchArgs args) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2 *f_dc = reinterpret_cast<float2 *>(smem + cfg.off_fdc);  /* [M][Nmem] */
    float2 *regA = reinterpret_cast<float2 *>(smem + cfg.off_A);    /* FFT scratch, then f_int [M][nint] */
    float2 *xin = reinterpret_cast<float2 *>(smem + cfg.off_B);     /* frame input, then timing products */
    float *Sf = reinterpret_cast<float *>(smem + cfg.off_Sf);
    float *tmpS = reinterpret_cast<float *>(smem + cfg.off_tmp);
    const float2 *tw = cfg.twiddle; /* read through L1 (__ldg): shared by every CTA on the SM */
    const float *hann = cfg.hann;
    FrameShared *fs = reinterpret_cast<FrameShared *>(smem + cfg.off_misc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const size_t cap = blockIdx.x;
    const int M = cfg.M, Nmem = cfg.Nmem, Ndft = cfg.Ndft, nint = cfg.nint, Nsym = cfg.Nsym, P = cfg.P, Ts = cfg.Ts;
    const int Nbits = cfg.Nbits;
    const unsigned char *iq = args.iq + cap * args.capture_stride_bytes;
    float *state = args.state ? args.state + cap * args.state_stride_floats : nullptr;

    /* ---- load tables and state ------------------------------------------------ */
    if (state) {
        const FskDevStateHdr *h = reinterpret_cast<const FskDevStateHdr *>(state);
        const float *sSf = state + sizeof(FskDevStateHdr) / sizeof(float);
        const float2 *sfdc = reinterpret_cast<const float2 *>(sSf + Ndft);
        for (int i = tid; i < Ndft; i += kThreads) Sf[i] = sSf[i];
        for (int i = tid; i < M * Nmem; i += kThreads) f_dc[i] = sfdc[i];
        if (tid == 0) {
            for (int m = 0; m < 4; ++m) fs->phi_c[m] = h->phi_c[m];
            fs->norm_rx_timing = h->norm_rx_timing;
            fs->ppm = h->ppm;
            fs->snr_est = h->snr_est;
            fs->EbNodB = h->EbNodB;
            fs->nin = h->nin;
        }
    } else {
        for (int i = tid; i < Ndft; i += kThreads) Sf[i] = 0.0f;
        for (int i = tid; i < M * Nmem; i += kThreads) f_dc[i] = make_float2(0.0f, 0.0f);
        if (tid == 0) {
            for (int m = 0; m < 4; ++m) fs->phi_c[m] = make_float2(1.0f, 0.0f);
            fs->norm_rx_timing = 0.0f;
            fs->ppm = 0.0f;
            fs->snr_est = 0.0f;
            fs->EbNodB = 0.0f;
            fs->nin = cfg.N;
        }
    }
    if (tid < 4) {
        fs->f_est[tid] = 0.0f;
        fs->f2_est[tid] = 0.0f;
    }
    __syncthreads();

    const float tc = cfg.tc;
    const float omtc = 1.0f - tc;
    size_t pos = 0;
    size_t frame = 0;
    unsigned char *bits_out = args.bits ? args.bits + cap * args.max_frames * (size_t)Nbits : nullptr;
    float *llr_out = args.llr ? args.llr + cap * args.max_frames * (size_t)Nbits : nullptr;
    float *filt_out = args.rx_filt ? args.rx_filt + cap * args.max_frames * (size_t)(M * Nsym) : nullptr;
    pirip_frame_stats *stats_out = args.stats ? args.stats + cap * args.max_frames : nullptr;

    while (frame < args.max_frames) {
        const int nin = fs->nin;
        if (pos + (size_t)nin > args.samples_per_capture) break; /* short read ends the loop */
        const int nold = Nmem - nin;

        /* ---- A5: input conversion ([U fsk_demod.c]) ---------------------------- */
        if (cfg.fmt == PIRIP_FMT_U8) {
            const unsigned char *src = iq + 2 * pos;
            /* 16-byte vector loads over the aligned body, bytes at the ragged ends */
            const size_t a0 = reinterpret_cast<size_t>(src);
            const int head = (int)((16 - (a0 & 15)) & 15); /* bytes before alignment */
            const int nbytes = 2 * nin;
            const int hb = head < nbytes ? head : nbytes;
            float *xf = reinterpret_cast<float *>(xin);
            for (int i = tid; i < hb; i += kThreads) xf[i] = (float)((int)src[i] - 127) * 0.0078125f;
            const int nvec = (nbytes - hb) >> 4;
            const uint4 *v = reinterpret_cast<const uint4 *>(src + hb);
            for (int c = tid; c < nvec; c += kThreads) {
                const uint4 q = __ldg(v + c);
                const unsigned w[4] = { q.x, q.y, q.z, q.w };
                float *o = xf + hb + (c << 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[4 * k + 0] = (float)((int)(w[k] & 0xff) - 127) * 0.0078125f;
                    o[4 * k + 1] = (float)((int)((w[k] >> 8) & 0xff) - 127) * 0.0078125f;
                    o[4 * k + 2] = (float)((int)((w[k] >> 16) & 0xff) - 127) * 0.0078125f;
                    o[4 * k + 3] = (float)((int)(w[k] >> 24) - 127) * 0.0078125f;
                }
            }
            for (int i = hb + (nvec << 4) + tid; i < nbytes; i += kThreads)
                xf[i] = (float)((int)src[i] - 127) * 0.0078125f;
        } else if (cfg.fmt == PIRIP_FMT_S16) {
            const short *src = reinterpret_cast<const short *>(iq) + 2 * pos;
            float *xf = reinterpret_cast<float *>(xin);
            for (int i = tid; i < 2 * nin; i += kThreads) xf[i] = (float)src[i] / 825.0f; /* FDMDV_SCALE */
        } else {
            const float2 *src = reinterpret_cast<const float2 *>(iq) + pos;
            for (int i = tid; i < nin; i += kThreads) xin[i] = src[i];
        }
        __syncthreads();

        /* ---- A7: tone frequency estimate --------------------------------------- */
        const int numffts = nin / (Ndft >> 1) - 1;
        for (int j = warp; j < numffts; j += kWarps)
            warp_fft_mag(cfg, xin + j * (Ndft >> 1), regA + (size_t)j * cfg.fft_slot_stride, tw, hann, lane);
        __syncthreads();
        for (int i = tid; i < Ndft; i += kThreads) {
            float s = Sf[i];
            for (int j = 0; j < numffts; ++j) {
                const float mg = reinterpret_cast<const float *>(regA + (size_t)j * cfg.fft_slot_stride)[i];
                s = (s * omtc) + (mg * tc);
            }
            Sf[i] = s;
            tmpS[i] = s;
        }
        __syncthreads();
        if (warp == 0) {
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
                if (lane == 0) fs->freqi[r] = ibest - (Ndft >> 1);
                __syncwarp();
            }
            if (lane == 0) {
                int fi[4];
                for (int m = 0; m < M; ++m) fi[m] = fs->freqi[m];
                for (int a = 1; a < M; ++a) { /* ascending insertion sort */
                    const int key = fi[a];
                    int b = a - 1;
                    while (b >= 0 && fi[b] > key) {
                        fi[b + 1] = fi[b];
                        --b;
                    }
                    fi[b + 1] = key;
                }
                for (int m = 0; m < M; ++m) fs->f_est[m] = (float)fi[m] * cfg.fs_over_ndft;
            }
        } else if (warp == 1) {
            /* estimator 2: slide the tone mask over Sf, first maximum wins */
            float best = 0.0f;
            int bbest = 0x7fffffff;
            for (int b = cfg.st + lane; b < cfg.en - cfg.len_mask; b += 32) {
                float corr = 0.0f;
                for (int q = 0; q < cfg.nmask; ++q) corr += Sf[b + cfg.mask_pos[q]];
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
                for (int m = 0; m < M; ++m) fs->f2_est[m] = foff + (float)(m * cfg.tone_spacing);
            }
        }
        __syncthreads();

        /* ---- A8: shift memory, spin the oscillators, mix to DC ------------------ */
        if (nold <= nin) {
            for (int m = 0; m < M; ++m)
                for (int i = tid; i < nold; i += kThreads) f_dc[m * Nmem + i] = f_dc[m * Nmem + i + nin];
        } else if (tid < M) {
            for (int i = 0; i < nold; ++i) f_dc[tid * Nmem + i] = f_dc[tid * Nmem + i + nin];
        }
        if (tid < M) {
            const float f = cfg.est_type ? fs->f2_est[tid] : fs->f_est[tid];
            const float arg = (float)(EM_TWO_PI * (double)(f / (float)cfg.Fs));
            double sn, cs;
            em_sincos((double)arg, &sn, &cs);
            fs->dphi[tid] = make_float2((float)cs, (float)sn);
        }
        __syncthreads();
        if (tid < M) {
            /* the oscillator recursion is inherently serial: one lane per tone */
            const float2 d = fs->dphi[tid];
            float2 ph = fs->phi_c[tid];
            float2 *dst = f_dc + tid * Nmem + nold;
#pragma unroll 8
            for (int j = 0; j < nin; ++j) {
                ph = cmul(ph, d);
                dst[j] = ph;
            }
            const float av = sqrtf((ph.x * ph.x) + (ph.y * ph.y));
            fs->phi_c[tid] = make_float2(ph.x / av, ph.y / av);
        }
        __syncthreads();
        for (int j = tid; j < nin; j += kThreads) {
            const float2 x = xin[j];
            for (int m = 0; m < M; ++m) {
                const float2 o = f_dc[m * Nmem + nold + j];
                f_dc[m * Nmem + nold + j] = make_float2(x.x * o.x + x.y * o.y, x.y * o.x - x.x * o.y);
            }
        }
        __syncthreads();

        /* ---- A9: integrate one symbol period at every timing offset -------------- */
        float2 *f_int = regA;
        const int step = Ts / P;
        if (step == 1 && Ts >= 4) {
            /* P == Ts: four adjacent windows per thread share their loads; every window is
               still accumulated in ascending sample order */
            const int nquad = (nint + 3) >> 2;
            for (int m = 0; m < M; ++m) {
                for (int q = tid; q < nquad; q += kThreads) {
                    const int i = q << 2;
                    const float2 *src = f_dc + m * Nmem + i;
                    float2 a0 = make_float2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0;
                    float2 v = src[0];
                    a0 = cadd(a0, v);
                    v = src[1];
                    a0 = cadd(a0, v);
                    a1 = cadd(a1, v);
                    v = src[2];
                    a0 = cadd(a0, v);
                    a1 = cadd(a1, v);
                    a2 = cadd(a2, v);
#pragma unroll 7
                    for (int j = 3; j < Ts; ++j) {
                        v = src[j];
                        a0 = cadd(a0, v);
                        a1 = cadd(a1, v);
                        a2 = cadd(a2, v);
                        a3 = cadd(a3, v);
                    }
                    f_int[m * nint + i] = a0;
                    if (i + 1 < nint) {
                        v = src[Ts];
                        a1 = cadd(a1, v);
                        f_int[m * nint + i + 1] = a1;
                        if (i + 2 < nint) {
                            const float2 w = src[Ts + 1];
                            a2 = cadd(cadd(a2, v), w);
                            f_int[m * nint + i + 2] = a2;
                            if (i + 3 < nint) {
                                a3 = cadd(cadd(cadd(a3, v), w), src[Ts + 2]);
                                f_int[m * nint + i + 3] = a3;
                            }
                        }
                    }
                }
            }
        } else {
            for (int m = 0; m < M; ++m) {
                for (int i = tid; i < nint; i += kThreads) {
                    const float2 *src = f_dc + m * Nmem + i * step;
                    float ar = 0.0f, ai = 0.0f;
#pragma unroll 4
                    for (int j = 0; j < Ts; ++j) {
                        const float2 v = src[j];
                        ar += v.x;
                        ai += v.y;
                    }
                    f_int[m * nint + i] = make_float2(ar, ai);
                }
            }
        }
        __syncthreads();

        /* ---- A10: timing line ---------------------------------------------------- */
        float2 *prod = xin;
        for (int i = tid; i < nint; i += kThreads) {
            float ft1 = 0.0f;
            for (int m = 0; m < M; ++m) {
                const float2 v = f_int[m * nint + i];
                ft1 += (v.x * v.x) + (v.y * v.y);
            }
            const float2 ph = __ldg(cfg.phi_ft + i);
            prod[i] = make_float2(ft1 * ph.x, ft1 * ph.y);
        }
        __syncthreads();
        if (tid == 0) {
            /* the timing-line sum is a serial float accumulation: real and imaginary chains
               interleaved on one lane, then the per-frame scalar estimates */
            float tre = 0.0f, tim = 0.0f;
#pragma unroll 8
            for (int i = 0; i < nint; ++i) {
                const float2 p = prod[i];
                tre += p.x;
                tim += p.y;
            }
            if (tre != tre || tim != tim) {
                fs->nan_flag = 1;
                fs->nin_next = nin;
            } else {
                fs->nan_flag = 0;
                const float norm = (float)((double)em_atan2f(tim, tre) / EM_TWO_PI);
                const float rx_timing = norm * (float)P;
                const float old = fs->norm_rx_timing;
                fs->norm_rx_timing = norm;
                const float dn = norm - old;
                if ((double)fabsf(dn) < .2) {
                    const float appm = (float)(1e6 * (double)dn / (double)(float)Nsym);
                    fs->ppm = (float)(.9 * (double)fs->ppm + .1 * (double)appm);
                }
                int nn = nin;
                if (!cfg.burst_mode && !cfg.lock_nin) {
                    if ((double)norm > 0.25)
                        nn = cfg.N + Ts / 4;
                    else if ((double)norm < -0.25)
                        nn = cfg.N - Ts / 4;
                    else
                        nn = cfg.N;
                }
                fs->nin_next = nn;
                const int low = (int)floorf(rx_timing);
                fs->low = low;
                fs->fract = rx_timing - (float)low;
                fs->one_minus_fract = 1.0f - fs->fract;
                fs->high = (int)ceilf(rx_timing);
                fs->rx_timing = rx_timing;
            }
        }
        __syncthreads();

        /* ---- A11: resample, decide, statistics ---------------------------------- */
        float *sig = reinterpret_cast<float *>(xin) + 2 * nint; /* region B tail: 3*Nsym floats */
        float *nse = sig + Nsym;
        float *sqm = nse + Nsym;
        if (!fs->nan_flag) {
            const float fract = fs->fract, omf = fs->one_minus_fract;
            const int low = fs->low, high = fs->high;
            for (int i = tid; i < Nsym; i += kThreads) {
                const int st = (i + 1) * P;
                float tmax[4];
                for (int m = 0; m < M; ++m) {
                    const float2 a = f_int[m * nint + st + low];
                    const float2 b = f_int[m * nint + st + high];
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
                if (bits_out) {
                    unsigned char *bo = bits_out + frame * (size_t)Nbits;
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
                    /* tone-major rx_filt, also staged in f_dc-free smem for the LLR map */
                    if (filt_out) filt_out[frame * (size_t)(M * Nsym) + m * Nsym + i] = rf;
                    sum += tmax[m];
                }
                sig[i] = mx;
                nse[i] = (sum - mx) / (float)(M - 1);
                sqm[i] = sqrtf(mx);
            }
        } else if (bits_out) {
            /* reference returns early on a NaN timing estimate and the CLI re-emits its bit
               buffer: repeat the previous frame's bits (zeros on the first frame) */
            for (int i = tid; i < Nbits; i += kThreads)
                bits_out[frame * (size_t)Nbits + i] = frame ? bits_out[(frame - 1) * (size_t)Nbits + i] : 0;
        }
        __syncthreads();
        if (tid < 4 && !fs->nan_flag) {
            /* four independent serial sums over the symbols, one lane each */
            const float *src = tid == 1 ? nse : (tid == 3 ? sqm : sig);
            float acc = tid == 1 ? 1E-12f : 0.0f;
            for (int i = 0; i < Nsym; ++i) acc += src[i];
            fs->sums[tid] = acc;
        }
        __syncthreads();
        if (tid == 0 && !fs->nan_flag) {
            float rx_sig_pow = fs->sums[0], rx_nse_pow = fs->sums[1], stdebno = fs->sums[2], meanebno = fs->sums[3];
            rx_sig_pow = rx_sig_pow / (float)Nsym;
            rx_nse_pow = rx_nse_pow / (float)Nsym;
            fs->rx_sig_pow = rx_sig_pow;
            fs->rx_nse_pow = rx_nse_pow;
            fs->v_est = (float)sqrt((double)(rx_sig_pow - rx_nse_pow));
            fs->SNRest = rx_sig_pow / rx_nse_pow;
            meanebno = meanebno / (float)Nsym;
            stdebno = (stdebno / (float)Nsym) - (meanebno * meanebno);
            if ((double)stdebno > 0.0)
                stdebno = (float)sqrt((double)stdebno);
            else
                stdebno = 0.0f;
            const float ratio = (float)((1e-6 + (double)meanebno) / (1e-6 + (double)stdebno));
            fs->EbNodB = -6.0f + (20.0f * em_log10f(ratio));
            fs->snr_est = (float)(.5 * (double)fs->snr_est + .5 * (double)fs->EbNodB);
            const float *fe = cfg.est_type ? fs->f2_est : fs->f_est;
            const float fc_avg = (fe[0] + fe[1]) / 2.0f;
            const float fc_tx = (float)((cfg.f1_tx + cfg.f1_tx + cfg.tone_spacing) / 2);
            fs->foff = fc_tx - fc_avg;
        }
        __syncthreads();

        /* ---- A12: soft decisions ([U mpdecode_core.c::fsk_rx_filt_to_llrs]) ------- */
        if (llr_out && !fs->nan_flag) {
            const float v_est = fs->v_est, scale = 2.0f * fs->SNRest;
            const float fract = fs->fract, omf = fs->one_minus_fract;
            const int low = fs->low, high = fs->high;
            const int bps = M == 2 ? 1 : 2;
            for (int i = tid; i < Nsym; i += kThreads) {
                const int st = (i + 1) * P;
                float num[2] = { -1000000.0f, -1000000.0f }, den[2] = { -1000000.0f, -1000000.0f };
                for (int m = 0; m < M; ++m) {
                    const float2 a = f_int[m * nint + st + low];
                    const float2 b = f_int[m * nint + st + high];
                    float2 t = make_float2(omf * a.x, omf * a.y);
                    t = cadd(t, make_float2(fract * b.x, fract * b.y));
                    const float yr = sqrtf((t.x * t.x) + (t.y * t.y));
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
                for (int k = 0; k < bps; ++k) llr_out[frame * (size_t)Nbits + bps * i + k] = -(num[k] - den[k]);
            }
        }
        if (tid == 0) {
            if (stats_out) {
                pirip_frame_stats s;
                for (int m = 0; m < 4; ++m) {
                    s.f_est[m] = m < M ? fs->f_est[m] : 0.0f;
                    s.f2_est[m] = m < M ? fs->f2_est[m] : 0.0f;
                }
                s.norm_rx_timing = fs->norm_rx_timing;
                s.ppm = fs->ppm;
                s.SNRest = fs->SNRest;
                s.v_est = fs->v_est;
                s.rx_sig_pow = fs->rx_sig_pow;
                s.rx_nse_pow = fs->rx_nse_pow;
                s.EbNodB = fs->EbNodB;
                s.snr_est = fs->snr_est;
                s.foff = fs->foff;
                s.rx_timing = fs->rx_timing;
                s.nin_next = fs->nin_next;
                s.sample_offset = (int32_t)pos;
                stats_out[frame] = s;
            }
            fs->nin = fs->nin_next;
        }
        pos += (size_t)nin;
        frame++;
        __syncthreads();
    }

    /* ---- save state ------------------------------------------------------------- */
    if (args.nframes && tid == 0) args.nframes[cap] = (int32_t)frame;
    if (args.Sf_out)
        for (int i = tid; i < Ndft; i += kThreads) args.Sf_out[cap * (size_t)Ndft + i] = Sf[i];
    if (state) {
        FskDevStateHdr *h = reinterpret_cast<FskDevStateHdr *>(state);
        float *sSf = state + sizeof(FskDevStateHdr) / sizeof(float);
        float2 *sfdc = reinterpret_cast<float2 *>(sSf + Ndft);
        for (int i = tid; i < Ndft; i += kThreads) sSf[i] = Sf[i];
        for (int i = tid; i < M * Nmem; i += kThreads) sfdc[i] = f_dc[i];
        if (tid == 0) {
            for (int m = 0; m < 4; ++m) h->phi_c[m] = fs->phi_c[m];
            h->norm_rx_timing = fs->norm_rx_timing;
            h->ppm = fs->ppm;
            h->snr_est = fs->snr_est;
            h->EbNodB = fs->EbNodB;
            h->nin = fs->nin;
        }
    }
}

}  // namespace

cudaError_t fsk_kernel_prepare(const FskDevCfg &cfg) {
    return cudaFuncSetAttribute(fsk_demod_kernel<FSK_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                cfg.smem_bytes);
}

cudaError_t fsk_launch_demod(const FskDevCfg &cfg, const FskBatchArgs &args, size_t n_captures, cudaStream_t stream) {
    if (n_captures == 0) return cudaSuccess;
    fsk_demod_kernel<FSK_THREADS><<<(unsigned)n_captures, FSK_THREADS, cfg.smem_bytes, stream>>>(cfg, args);
    return cudaGetLastError();
}
