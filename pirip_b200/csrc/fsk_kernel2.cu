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
 *      FFT warps        the frequency estimate of the NEXT frame of every capture (its position is
 *                       already known and its TMA copy was issued a frame ahead) as a pool of
 *                       (capture, block) tasks: Hann + radix-4/2 FFT, magnitudes applied to the IIR
 *                       spectrum in block order, peak pick / mask, dphi.
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
/* Wait for the phase with the given parity.  try_wait returns after a short system-dependent time when the
   phase is still pending (measured on B200: a spinning loop of try_wait was 42% of all issued instructions),
   so a pending wait backs off with nanosleep and leaves the issue slots to the warps that have work. */
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
/* release / acquire accesses to a shared-memory flag (CTA scope): the ordered hand-off of the IIR spectrum between FFT warps */
__device__ __forceinline__ void st_release_shared(int *p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_shared(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

/* hardware named barriers: a warp blocked in bar.sync costs no issue slots (unlike a polled mbarrier) */
__device__ __forceinline__ void nbar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

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

/* ---- frequency estimator ([U fsk.c::fsk_demod_freq_est]) as a pool of FFT tasks --------------------------------
   Task (g, j) = FFT block j of the frame being estimated for capture g.  The FFT warps take tasks round-robin, so
   the blocks of one capture are transformed concurrently by different warps; the single-pole magnitude IIR must run
   over the blocks in order, so each task applies its magnitudes (held in registers) to the spectrum when its turn
   comes (cs->sf_turn), and the task of the last block picks the tones. */

/* tone picking of one capture from its updated spectrum Sf (peak estimator with blanking, mask estimator, dphi) */
__device__ __noinline__ void pick_tones(const FskDevCfg &cfg, Fsk2CapShared *cs, const float *Sf, float *tmpS, int fb, int M, int lane) {
    const int Ndft = cfg.Ndft;
    for (int i = lane; i < Ndft; i += 32) tmpS[i] = Sf[i];
    __syncwarp();
    /* estimator 1: M rounds of argmax (first maximum wins) + blanking, on the scratch copy */
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
        __syncwarp(); /* every lane has read its part of tmpS before anyone blanks */
        for (int j = f_min + lane; j < f_max; j += 32) tmpS[j] = 0.0f;
        if (lane == 0) cs->freqi[r] = ibest - (Ndft >> 1);
        __syncwarp();
    }
    if (lane == 0) {
        int fi[4];
        for (int m = 0; m < M; ++m) fi[m] = cs->freqi[m];
        for (int a = 1; a < M; ++a) { /* ascending insertion sort */
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
}

/* one FFT task: block j of capture cs for the frame of parity fb; src = that frame's raw samples */
template <int LOGN>
__device__ __noinline__ void fft_task(const FskDevCfg &cfg, Fsk2CapShared *cs, float *Sf2, const unsigned char *src, float2 *buf,
                                      const float2 *tw, const float *hann, const unsigned short *perm, int j, int nb, int fb, int tag,
                                      int lane) {
    constexpr int N = 1 << LOGN, Q = N / 32;
    float mg[Q];
    warp_fft_mag<LOGN>(src, cfg.fmt, j * (N >> 1), buf, mg, tw, hann, perm, lane);
    /* ordered IIR: wait for block j - 1 of this capture */
    int *turn = &cs->sf_turn;
    {
        int tries = 0;
        while (ld_acquire_shared(turn) != j) {
            if (++tries > 64) __nanosleep(100);
        }
    }
    const float tc = cfg.tc, omtc = 1.0f - tc;
    const float *Sf_in = j == 0 ? Sf2 + (size_t)(fb ^ 1) * N : Sf2 + (size_t)fb * N;
    float *Sf_out = Sf2 + (size_t)fb * N;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int i = lane + (q << 5);
        Sf_out[i] = (Sf_in[i] * omtc) + (mg[q] * tc);
    }
    __syncwarp();
    if (lane == 0) {
        if (j + 1 < nb) {
            st_release_shared(turn, j + 1);
        } else {
            st_release_shared(turn, 0);
            st_release_shared(&cs->est_done, tag); /* the spectrum is complete: the tones can be picked (pick_tones) */
        }
    }
}

/* ---- 256-point fast path: two FFT blocks per warp, two radix-4 levels per register pass ----------------------
   Same butterflies, twiddles and operation order as kiss_fft's 4x4x4x4 decimation in time (and as warp_fft_mag),
   so the same floats; only the data movement differs.  Lanes 0-15 transform block j0, lanes 16-31 block j0 + 1.
   Pass A: lane L holds the 16 (digit-reversed) positions 16 L + i and runs levels m = 1 and m = 4 in registers.
   Exchange through shared memory (real parts, then imaginary parts: 2 KB of scratch serve both blocks).
   Pass B: lane L holds positions 16 h + L and runs levels m = 16 and m = 64; position = frequency bin.
   The twiddle of every level-1 butterfly is tw[0] = (1, -0): multiplying by it changes at most the sign of a zero,
   which the magnitudes cannot see, so level 1 skips the multiplications. */
__device__ __forceinline__ void bfly4(float2 &F0, float2 &F1, float2 &F2, float2 &F3, float2 T0, float2 T1, float2 T2) {
    const float2 s0 = cmul(F1, T0), s1 = cmul(F2, T1), s2 = cmul(F3, T2);
    const float2 s5 = csub(F0, s1);
    const float2 G0 = cadd(F0, s1);
    const float2 s3 = cadd(s0, s2), s4 = csub(s0, s2);
    F2 = csub(G0, s3);
    F0 = cadd(G0, s3);
    F1 = make_float2(s5.x + s4.y, s5.y - s4.x);
    F3 = make_float2(s5.x - s4.y, s5.y + s4.x);
}
__device__ __forceinline__ void bfly4_unit(float2 &F0, float2 &F1, float2 &F2, float2 &F3) {
    const float2 s0 = F1, s1 = F2, s2 = F3;
    const float2 s5 = csub(F0, s1);
    const float2 G0 = cadd(F0, s1);
    const float2 s3 = cadd(s0, s2), s4 = csub(s0, s2);
    F2 = csub(G0, s3);
    F0 = cadd(G0, s3);
    F1 = make_float2(s5.x + s4.y, s5.y - s4.x);
    F3 = make_float2(s5.x - s4.y, s5.y + s4.x);
}

#define FSK2_XCHG_HALF (17 * 16) /* floats per block in the exchange scratch (16 rows of 17): the blocks end up 16 banks apart */

/* magnitudes of bins 16 h + L (L = lane & 15) of block j0 (lanes 0-15) / j0 + 1 (lanes 16-31) in mg[h] */
__device__ __forceinline__ void fft256_pair(const unsigned char *src, int fmt, int first0, bool two, float *xbuf, const float2 *tw,
                                            const float *hann, float (&mg)[16], int lane) {
    const int half = lane >> 4, L = lane & 15;
    const int first = first0 + ((two && half) ? 128 : 0); /* an absent second block repeats the first (discarded) */
    float2 x[16];
    {
        const int c = (L & 3) * 4 + (L >> 2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = (i & 3) * 64 + (i >> 2) * 16 + c; /* input index whose digit-reversed position is 16 L + i */
            const float h = hann[n];
            const float2 v = raw_sample(src, fmt, first + n);
            x[i] = make_float2(h * v.x, h * v.y);
        }
    }
    /* level m = 1: positions 4 q + {0,1,2,3} */
#pragma unroll
    for (int q = 0; q < 4; ++q) bfly4_unit(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    /* level m = 4: positions k + {0,4,8,12}, twiddles tw[16 k], tw[32 k], tw[48 k] */
#pragma unroll
    for (int k = 0; k < 4; ++k) bfly4(x[k], x[k + 4], x[k + 8], x[k + 12], tw[16 * k], tw[32 * k], tw[48 * k]);
    /* exchange: position 16 L + i -> lane i; rows padded to 17 words and the two blocks 16 banks apart, so the 32
       lanes hit 32 distinct banks both ways and every access is base + immediate */
    float *xw = xbuf + half * FSK2_XCHG_HALF + 17 * L; /* writer view: position 16 L + i at word 17 L + i */
    const float *xr = xbuf + half * FSK2_XCHG_HALF + L; /* reader view: position 16 h + L at word 17 h + L */
    float2 y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) xw[i] = x[i].x;
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 16; ++h) y[h].x = xr[17 * h];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) xw[i] = x[i].y;
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 16; ++h) y[h].y = xr[17 * h];
    __syncwarp();
    /* level m = 16: positions 64 b + L + 16 r, twiddles tw[4 L], tw[8 L], tw[12 L] */
    {
        const float2 T0 = tw[4 * L], T1 = tw[8 * L], T2 = tw[12 * L];
#pragma unroll
        for (int b = 0; b < 4; ++b) bfly4(y[4 * b], y[4 * b + 1], y[4 * b + 2], y[4 * b + 3], T0, T1, T2);
    }
    /* level m = 64: positions k + 64 r with k = 16 c + L */
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = 16 * c + L;
        bfly4(y[c], y[c + 4], y[c + 8], y[c + 12], tw[k], tw[2 * k], tw[3 * k]);
    }
#pragma unroll
    for (int h = 0; h < 16; ++h) mg[h] = sqrtf((y[h].x * y[h].x) + (y[h].y * y[h].y));
}

/* FFT task pair: blocks j0 and j0 + 1 (if it exists) of capture cs for the frame of parity fb */
__device__ __noinline__ void fft256_task(const FskDevCfg &cfg, Fsk2CapShared *cs, float *Sf2, const unsigned char *src, float2 *buf,
                                         const float2 *tw, const float *hann, int j0, int nb, int fb, int tag, int lane) {
    constexpr int N = 256;
    const bool two = j0 + 1 < nb;
    float mg[16];
    fft256_pair(src, cfg.fmt, j0 * (N >> 1), two, reinterpret_cast<float *>(buf), tw, hann, mg, lane);
    /* ordered IIR: wait for the blocks before j0 of this capture (short spin, then back off) */
    int *turn = &cs->sf_turn;
    {
        int tries = 0;
        while (ld_acquire_shared(turn) != j0) {
            if (++tries > 64) __nanosleep(100);
        }
    }
    const int half = lane >> 4, L = lane & 15;
    const float tc = cfg.tc, omtc = 1.0f - tc;
    const float *Sf_in = j0 == 0 ? Sf2 + (size_t)(fb ^ 1) * N : Sf2 + (size_t)fb * N;
    float *Sf_out = Sf2 + (size_t)fb * N;
    if (half == 0) {
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            const int i = (16 * h + L + (N >> 1)) & (N - 1); /* fft shift: DC at bin N/2 */
            Sf_out[i] = (Sf_in[i] * omtc) + (mg[h] * tc);
        }
    }
    __syncwarp();
    if (half == 1 && two) {
#pragma unroll
        for (int h = 0; h < 16; ++h) {
            const int i = (16 * h + L + (N >> 1)) & (N - 1);
            Sf_out[i] = (Sf_out[i] * omtc) + (mg[h] * tc);
        }
    }
    __syncwarp();
    const int jn = j0 + (two ? 2 : 1);
    if (lane == 0) {
        if (jn < nb) {
            st_release_shared(turn, jn);
        } else {
            st_release_shared(turn, 0);
            st_release_shared(&cs->est_done, tag); /* the spectrum is complete: the tones can be picked (pick_tones) */
        }
    }
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

/* Streamed ingest: the host copies the captures to the device in time slices while the kernel runs and publishes how many
   samples of every capture have landed; a frame is fetched only once all of it is there. */
__device__ __noinline__ void wait_landed(const FskBatchArgs &args, unsigned long long need) {
    if (!args.landed) return;
    if (need > args.samples_per_capture) need = args.samples_per_capture;
    const volatile unsigned long long *p = args.landed;
    /* the copies are issued by the host right after the launch; if they never come (a tool serialised the launch), fail
       the launch after a few seconds instead of occupying the GPU for ever */
    unsigned spins = 0;
    while (*p < need) {
        __nanosleep(400);
        if (++spins > 12000000u) __trap();
    }
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
    const size_t batch = args.total_captures ? args.total_captures : args.n_captures;
    const size_t safe = (base + (batch - 1) * args.capture_stride_bytes + args.samples_per_capture * bps) & ~(size_t)15;
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

/* NWIN adjacent integrator windows of one tone: window k sums ring samples s0 + k .. s0 + k + TS - 1 in ascending order
   from zero (packed f32x2 adds: real and imaginary part together); |sum|^2 of window k in en[k].  s0 even. */
template <int TS, int NWIN>
__device__ __forceinline__ void integrate_windows(const float2 *ring_m, unsigned s0, unsigned umask, float (&en)[NWIN]) {
    u64 acc[NWIN];
#pragma unroll
    for (int k = 0; k < NWIN; ++k) acc[k] = 0ull;
    const unsigned u0 = s0 >> 1;
    constexpr int NU = (NWIN + TS) / 2, NB = 8; /* 16-byte units, loaded NB at a time ahead of the adds that use them */
#pragma unroll
    for (int r0 = 0; r0 < NU; r0 += NB) {
        ulonglong2 v[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (r0 + b < NU) {
                const unsigned u = (u0 + r0 + b) & umask;
                const unsigned pu = u ^ ((u >> 3) & 7u);
                v[b] = *reinterpret_cast<const ulonglong2 *>(ring_m + 2 * pu);
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int r = r0 + b;
            if (r < NU) {
#pragma unroll
                for (int k = 0; k < NWIN; ++k) {
                    if (2 * r >= k && 2 * r < k + TS) acc[k] = add2(acc[k], v[b].x);
                    if (2 * r + 1 >= k && 2 * r + 1 < k + TS) acc[k] = add2(acc[k], v[b].y);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NWIN; ++k) {
        const float2 a = upk(acc[k]);
        en[k] = (a.x * a.x) + (a.y * a.y);
    }
}

/* one window, any Ts / alignment */
__device__ __forceinline__ float window_energy(const float2 *ring_m, unsigned s0, unsigned rmask, int Ts) {
    u64 acc = 0ull;
    for (int j = 0; j < Ts; ++j) acc = add2(acc, *reinterpret_cast<const u64 *>(ring_m + rsw(s0 + j, rmask)));
    const float2 a = upk(acc);
    return (a.x * a.x) + (a.y * a.y);
}

#ifdef FSK2_TIMING
/* Debug trace (build with EXTRA=-DFSK2_TIMING): lane 0 of every warp of block 0 stores (code, clock64) pairs for a
   couple of steady-state frames; read back with pirip_debug_trace().  Global stores only: nothing waits on them. */
#define FSK2_TRACE_N 512
__device__ long long g_trace[32][FSK2_TRACE_N];
#define TR(code)                                                                                                   \
    do {                                                                                                           \
        if (tr_on && tr_n < FSK2_TRACE_N) { /* a barrier only blocks at the next access to protected state: touch it first */ \
            const int sink_ = *reinterpret_cast<volatile int *>(&ctl->G);                                           \
            g_trace[warp][tr_n++] = ((long long)(code) << 48) | ((clock64() + (sink_ & 0)) & 0xffffffffffffll);     \
        }                                                                                                          \
    } while (0)
#else
#define TR(code)
#endif

template <int M, int NW>
__global__ void __launch_bounds__(NW * 32, 1) fsk2_kernel(const __grid_constant__ FskDevCfg cfg, const __grid_constant__ FskBatchArgs args) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nthreads = NW * 32;
    /* streaming phase: workers [0, W), timing warp, oscillator warp.  The frequency estimate of the next frame is a
       phase of its own run by all NF = NW warps (FFT, streaming and decision code each fit the instruction cache;
       run concurrently they thrashed it: "no instruction" was the top stall reason) */
    constexpr int W = NW - 2;
    const int NF = cfg.k2.fft_warps; /* warps of the FFT phase: NW / 2 with a dedicated scratch slot, the rest borrow the
                                        idle upper part of a capture's energy ring */
    constexpr int osc_warp = NW - 1, tim_warp = NW - 2;
    constexpr int CKE = FSK2_CKPT_EVERY;
    const int Ndft = cfg.Ndft, nint = cfg.nint, Nsym = cfg.Nsym, P = cfg.P, Ts = cfg.Ts, Nmem = cfg.Nmem, Nbits = cfg.Nbits;
    const int step = Ts / P;
    const int CH = cfg.k2.chunk, R = cfg.k2.ring, ER = cfg.k2.ering, NCK = cfg.k2.nckpt, TAILN = cfg.k2.tail;
    const unsigned rmask = (unsigned)R - 1u, umask = ((unsigned)R >> 1) - 1u;
    const int RT = R + 2; /* tone stride of the ring: 16 bytes of skew between tones */
    const int fmt = cfg.fmt;
    const int nb = cfg.k2.numffts;
    unsigned char *cta = smem + (size_t)cfg.k2.gmax * cfg.k2.cap_stride;
    Fsk2CtaShared *ctl = reinterpret_cast<Fsk2CtaShared *>(cta + cfg.k2.off_ctl);
    float2 *tw = reinterpret_cast<float2 *>(cta + cfg.k2.off_tw);
    float *hann = reinterpret_cast<float *>(cta + cfg.k2.off_hann);
    unsigned short *perm = reinterpret_cast<unsigned short *>(cta + cfg.k2.off_perm);
    float2 *fft_slots = reinterpret_cast<float2 *>(cta + cfg.k2.off_fft);
    float2 *phi_ft = reinterpret_cast<float2 *>(cta + cfg.k2.off_phift); /* timing-line oscillator, [nint] */
    const unsigned char **srcp = reinterpret_cast<const unsigned char **>(ctl + 1); /* [gmax][2] raw-sample pointer per staging buffer */

    for (int i = tid; i < Ndft; i += nthreads) {
        tw[i] = cfg.twiddle[i];
        hann[i] = cfg.hann[i];
        perm[i] = cfg.perm[i];
    }
    for (int i = tid; i < nint; i += nthreads) phi_ft[i] = cfg.phi_ft[i];

    /* list mode: only the captures the fast kernel flagged (their count is known on the device only) */
    size_t ncap = args.n_captures;
    if (args.cap_count) {
        const size_t n = (size_t)*args.cap_count;
        ncap = n < ncap ? n : ncap;
    }
    const int grid = (int)gridDim.x;
    const int ncap_cta = ncap > blockIdx.x ? (int)((ncap - blockIdx.x + grid - 1) / grid) : 0;
#ifdef FSK2_TIMING
    bool tr_on = false;
    int tr_n = 0;
#endif

    for (int k0 = 0; k0 < ncap_cta; k0 += cfg.k2.gmax) {
        const int G = ncap_cta - k0 < cfg.k2.gmax ? ncap_cta - k0 : cfg.k2.gmax;
        /* worker warps with a capture; the others have no part in the streaming hand-offs and run FFT tasks of the next
           frame during the streaming phase instead */
        const int nwork = G < W ? G : W;
        const int osc_cnt = 32 * (nwork + 1), int_cnt = nwork;
        __syncthreads();
        /* ---- barriers and initial state of the group ---- */
        if (tid == 0) {
            for (int i = 0; i < FSK2_MAX_CHUNKS; ++i) {
                mbar_init(&ctl->bar_int[i], int_cnt);
                mbar_init(&ctl->bar_tim[i], 1);
            }
            for (int g = 0; g < FSK2_MAX_G; ++g) {
                mbar_init(&ctl->bar_in[g][0], 1);
                mbar_init(&ctl->bar_in[g][1], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            ctl->G = G;
            /* hand-out order of the post-decision queue: one round of FFT tasks, the per-frame sums (they need only the
               decisions), the remaining FFT tasks with the tone picking of a capture one round behind its last FFT task
               (so that it rarely waits), the remaining pickings */
            {
                const bool prs = cfg.log2_ndft == 8;
                const int tpc_ = prs ? (nb + 1) >> 1 : nb;
                const int nf = G * tpc_;
                int n = 0, picked = 0;
                const bool mixed = cfg.k2.task_queue == 2;
                for (int t = 0; t < nf; ++t) {
                    ctl->queue[n++] = (unsigned short)((FSK2_Q_FFT << 12) | t);
                    if (mixed && t == (NW < nf ? NW : nf) - 1)
                        for (int g = 0; g < G; ++g) ctl->queue[n++] = (unsigned short)((FSK2_Q_SUMS << 12) | g);
                    while (mixed && picked < G && (picked + 1) * tpc_ - 1 <= t - NW)
                        ctl->queue[n++] = (unsigned short)((FSK2_Q_PICK << 12) | picked++);
                }
                if (!mixed)
                    for (int g = 0; g < G; ++g) ctl->queue[n++] = (unsigned short)((FSK2_Q_SUMS << 12) | g);
                while (picked < G) ctl->queue[n++] = (unsigned short)((FSK2_Q_PICK << 12) | picked++);
                ctl->q_len = n;
            }
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
                cs->sf_turn = 0;
                cs->est_done = 0;
                cs->cap = (int)(blockIdx.x + (size_t)(k0 + g) * grid);
                if (args.cap_list) cs->cap = args.cap_list[cs->cap];
                cs->stg_valid[0] = cs->stg_valid[1] = 0;
                cs->stg_fills[0] = cs->stg_fills[1] = 0;
                cs->fast = 0;
                /* frame 0 exists for some nin iff its shortest form fits */
                cs->fon = args.max_frames > 0 && (unsigned long long)(cfg.N - Ts / 4) <= args.samples_per_capture;
                cs->active = args.max_frames > 0 && (unsigned long long)cfg.N <= args.samples_per_capture;
            }
        }
        __syncthreads();
        if (tid < G) {
            const Cap2 c = cap2(smem, cfg, tid);
            const unsigned char *s0 = nullptr;
            if (c.cs->fon) {
                wait_landed(args, (unsigned long long)(cfg.N + Ts / 4 + 8));
                stage_issue(cfg, args, c, 0, 0, &ctl->bar_in[tid][0], &s0);
            }
            srcp[2 * tid + 0] = s0;
        }
        __syncthreads();

        /* the FFT tasks of this warp for the frame of parity fbt (staging buffer fbt) */
        const bool pairs = cfg.log2_ndft == 8;        /* 256-point fast path: two blocks per task */
        const int tpc = pairs ? (nb + 1) >> 1 : nb;   /* tasks per capture */
        /* tasks [0, t_early) are run during the streaming phase by the worker warps without a capture (dedicated scratch
           slots only), the rest after the decisions by all FFT warps */
        const int early_pool = (W - nwork) < NW / 2 ? (W - nwork) : NW / 2;
        const int t_early = early_pool * cfg.k2.fft_early < G * tpc ? early_pool * cfg.k2.fft_early : G * tpc;
        auto fft_tasks = [&](int fbt, int t_first, int t_end, int idx, int stride, int slot, int tag) {
            float2 *buf = slot < NW / 2 ? fft_slots + (size_t)slot * FSK2_FFT_SLOT_ELEMS(Ndft)
                                        : reinterpret_cast<float2 *>(cap2(smem, cfg, slot - NW / 2).ering + cfg.k2.ering_fft_off);
            for (int t = t_first + idx; t < t_end; t += stride) {
                const int g = t / tpc, j = (t - g * tpc) * (pairs ? 2 : 1);
                const Cap2 c = cap2(smem, cfg, g);
                Fsk2CapShared *cs = c.cs;
                if (!cs->fon) continue;
                if (cs->stg_valid[fbt]) mbar_wait(&ctl->bar_in[g][fbt], (unsigned)(cs->stg_fills[fbt] - 1) & 1u);
                const unsigned char *src = srcp[2 * g + fbt];
                TR(52);
                switch (cfg.log2_ndft) {
                    case 6: fft_task<6>(cfg, cs, c.Sf, src, buf, tw, hann, perm, j, nb, fbt, tag, lane); break;
                    case 7: fft_task<7>(cfg, cs, c.Sf, src, buf, tw, hann, perm, j, nb, fbt, tag, lane); break;
                    case 8: fft256_task(cfg, cs, c.Sf, src, buf, tw, hann, j, nb, fbt, tag, lane); break;
                    case 9: fft_task<9>(cfg, cs, c.Sf, src, buf, tw, hann, perm, j, nb, fbt, tag, lane); break;
                    default: fft_task<10>(cfg, cs, c.Sf, src, buf, tw, hann, perm, j, nb, fbt, tag, lane); break;
                }
                TR(53);
            }
        };
        /* tone picking for the frame of parity fbt: one warp per capture, highest warps first (the low ones do the
           per-frame sums of the same phase) */
        auto pick_phase = [&](int fbt) {
            if (warp >= NF) return;
            float *tmpS = reinterpret_cast<float *>(warp < NW / 2 ? fft_slots + (size_t)warp * FSK2_FFT_SLOT_ELEMS(Ndft)
                                                                   : reinterpret_cast<float2 *>(cap2(smem, cfg, warp - NW / 2).ering + cfg.k2.ering_fft_off));
            for (int g = NF - 1 - warp; g < G; g += NF) {
                const Cap2 c = cap2(smem, cfg, g);
                if (c.cs->fon) pick_tones(cfg, c.cs, c.Sf + (size_t)fbt * Ndft, tmpS, fbt, M, lane);
            }
        };
        /* per-capture end-of-frame bookkeeping by lane g of warp 0: does capture g have a whole frame left; which
           integrator path it takes; fetch the frame after it.  The CTA-wide flags come from a warp reduction. */
        auto bookkeeping = [&](int f_next) {
            /* warp 0 only */
            const bool mine = lane < G;
            const Cap2 c = cap2(smem, cfg, mine ? lane : 0);
            Fsk2CapShared *cs = c.cs;
            int act = 0, nin = 0, gen = 0;
            if (mine) {
                act = cs->active;
                if (act && ((size_t)f_next >= args.max_frames || cs->pos + (unsigned long long)cs->nin > args.samples_per_capture))
                    act = 0; /* short read ends the loop */
                cs->active = act;
                cs->nold = Nmem - cs->nin;
                /* the 16-window integrator needs 16-byte aligned window starts: f_dc index 0 at an even position */
                cs->fast = cfg.k2.fast_ts && step == 1 && ((((unsigned)cs->pos - (unsigned)cs->nold) & 1u) == 0);
                nin = act ? cs->nin : 0;
                gen = act && !cs->fast;
                /* plan and fetch frame f_next + 1 into the staging buffer frame f_next - 1 used */
                cs->fon = 0;
                if (act) {
                    const unsigned long long npos = cs->pos + (unsigned long long)cs->nin;
                    if ((size_t)f_next + 1 < args.max_frames && npos + (unsigned long long)(cfg.N - Ts / 4) <= args.samples_per_capture) {
                        const int bn = (f_next + 1) & 1;
                        const unsigned char *s1 = nullptr;
                        wait_landed(args, npos + (unsigned long long)(cfg.N + Ts / 4 + 8));
                        stage_issue(cfg, args, c, bn, npos, &ctl->bar_in[lane][bn], &s1);
                        srcp[2 * lane + bn] = s1;
                        cs->fon = 1;
                    }
                }
            }
            const unsigned any = __ballot_sync(0xffffffffu, act != 0);
            const unsigned anyg = __ballot_sync(0xffffffffu, gen != 0);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int o = __shfl_xor_sync(0xffffffffu, nin, off);
                nin = o > nin ? o : nin;
            }
            if (lane == 0) {
                ctl->any_active = any != 0;
                ctl->max_nin = nin;
                ctl->any_generic = anyg != 0;
            }
        };

        /* ---- prologue: frequency estimate of frame 0, then plan frame 0 / fetch frame 1 ---- */
        if (warp < NF) fft_tasks(0, 0, G * tpc, warp, NF, warp, 0);
        __syncthreads();
        pick_phase(0);
        __syncthreads();
        if (warp == 0) bookkeeping(0);
        __syncthreads();
        TR(1);

        unsigned par = 0; /* phase parity of the per-chunk mbarriers, bit c */
        for (int f = 0;; ++f) {
            if (!ctl->any_active) break;
            const int fb = f & 1;
#ifdef FSK2_TIMING
            tr_on = blockIdx.x == 0 && lane == 0 && f >= 100 && f < 102;
            TR(3);
#endif
            const int max_nin = ctl->max_nin;
            const int nchunks = (max_nin + CH - 1) / CH;
            if (tid == 0) { /* the queue of this frame opens after the streaming phase's barrier */
                ctl->task_next = 0;
                ctl->decide_done = 0;
            }
            /* visibility of frame f's staged samples (the copy was issued a frame ago) */
            for (int g = 0; g < G; ++g) {
                Fsk2CapShared *cs = cap2(smem, cfg, g).cs;
                if (cs->active && cs->stg_valid[fb]) mbar_wait(&ctl->bar_in[g][fb], (unsigned)(cs->stg_fills[fb] - 1) & 1u);
            }
            TR(2);

            /* ================= streaming phase ================= */
            if (warp == osc_warp) {
                /* A8: oscillator recursion, lane (g, m); two steps per 16-byte store */
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
                const bool vec_ok = __all_sync(0xffffffffu, !on || (p0 & 1u) == 0);
                for (int ck = 0; ck < nchunks; ++ck) {
                    if (ck >= 3) mbar_wait(&ctl->bar_int[ck - 3], (par >> (ck - 3)) & 1u); /* ring slots free again */
                    TR(20);
                    const int j0 = ck * CH;
                    const int jend = j0 + CH < max_nin ? j0 + CH : max_nin;
                    if (jend <= lo && vec_ok) {
                        if (on) {
                            unsigned u = ((p0 + (unsigned)j0) & rmask) >> 1; /* 16-byte unit of the ring */
                            int j = j0;
                            for (; j + 16 <= jend; j += 16) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const float2 p1 = cmul(ph, d);
                                    ph = cmul(p1, d);
                                    const unsigned uu = (u + q) & umask;
                                    const unsigned pu = uu ^ ((uu >> 3) & 7u);
                                    *reinterpret_cast<float4 *>(ring_m + 2 * pu) = make_float4(p1.x, p1.y, ph.x, ph.y);
                                }
                                u = (u + 8) & umask;
                                ck_m[(j + 16) / CKE] = ph;
                            }
                            for (; j < jend; ++j) {
                                ph = cmul(ph, d);
                                ring_m[rsw(p0 + (unsigned)j, rmask)] = ph;
                                if (((j + 1) & (CKE - 1)) == 0) ck_m[(j + 1) / CKE] = ph;
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
                    TR(22);
                    /* hand the chunk to the workers: named barriers 1..4 in rotation (the oscillator is never more
                       than three chunks ahead of the integrators) */
                    nbar_arrive(1 + (ck & 3), osc_cnt);
                }
                if (on) {
                    const float av = sqrtf((ph.x * ph.x) + (ph.y * ph.y));
                    c.cs->phi_c[m] = make_float2(ph.x / av, ph.y / av);
                }
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
                    TR(30);
                    mbar_wait(&ctl->bar_int[ck], (par >> ck) & 1u);
                    TR(31);
                    const int avail = on ? windows_after_chunk(v, ck, CH, Ts, step, nint) : 0;
                    /* the workers left the timing-line products e_i * phi_ft[i]; eight per iteration while aligned */
                    const float2 *pr = reinterpret_cast<const float2 *>(er);
                    while ((done & 7) == 0 && done + 8 <= avail) {
                        const float4 *src4 = reinterpret_cast<const float4 *>(pr + (done & (ER - 1)));
                        const float4 a0 = src4[0], a1 = src4[1], a2 = src4[2], a3 = src4[3];
                        tre += a0.x, tim += a0.y;
                        tre += a0.z, tim += a0.w;
                        tre += a1.x, tim += a1.y;
                        tre += a1.z, tim += a1.w;
                        tre += a2.x, tim += a2.y;
                        tre += a2.z, tim += a2.w;
                        tre += a3.x, tim += a3.y;
                        tre += a3.z, tim += a3.w;
                        done += 8;
                    }
                    while (done < avail) {
                        const float2 v2 = pr[done & (ER - 1)];
                        tre += v2.x;
                        tim += v2.y;
                        done += 1;
                    }
                    TR(32);
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
            } else if (warp >= nwork) {
                /* A7 for frame f + 1, first part: its samples were fetched a frame ahead and its position is known */
                if (warp - nwork < early_pool) fft_tasks(fb ^ 1, 0, t_early, warp - nwork, early_pool, warp - nwork, f + 1);
            } else {
                /* A8 mix of chunk ck of capture g by one warp: f_dc = in * conj(phi) in place over the oscillator values;
                   four samples per lane, all loads first */
                auto mix_chunk = [&](int g, int ck) {
                    const Cap2 c = cap2(smem, cfg, g);
                    const FrameView v = load_view(c.cs, true);
                    const unsigned char *src = srcp[2 * g + fb];
                    const int j0 = ck * CH;
                    for (int q0 = 0; q0 < CH; q0 += 128) {
                        float2 x[4], o[4][M];
                        unsigned sl[4];
                        bool ok[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = j0 + q0 + lane + 32 * q;
                            ok[q] = j < v.nin;
                            sl[q] = rsw(v.p0 + (unsigned)j, rmask);
                            x[q] = ok[q] ? raw_sample(src, fmt, j) : make_float2(0.0f, 0.0f);
#pragma unroll
                            for (int m = 0; m < M; ++m) o[q][m] = c.ring[m * RT + sl[q]];
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (ok[q]) {
#pragma unroll
                                for (int m = 0; m < M; ++m) c.ring[m * RT + sl[q]] = mixc(x[q], o[q][m]);
                            }
                        }
                    }
                    __syncwarp();
                };
                /* A9 for the windows of capture g completed by chunk ck, by one warp: integrate, |.|^2, sum over the tones,
                   multiply by the timing oscillator */
                auto integrate_chunk = [&](int g, int ck) {
                    const Cap2 c = cap2(smem, cfg, g);
                    const FrameView v = load_view(c.cs, true);
                    const int w_lo = windows_after_chunk(v, ck - 1, CH, Ts, step, nint);
                    const int w_hi = windows_after_chunk(v, ck, CH, Ts, step, nint);
                    float2 *prod = reinterpret_cast<float2 *>(c.ering);
                    if (v.fast) {
                        /* NWIN windows of one tone per lane: lanes (m, q), so the M tones of a window group sit QPC lanes apart;
                           the tone energies are summed in tone order across those lanes and each lane turns NWIN / M of the
                           sums into timing-line products */
                        constexpr int NWIN = 4 * M, QPC = 32 / M;
                        const int m = lane / QPC, q = lane & (QPC - 1);
                        const float2 *ring_m = c.ring + m * RT;
                        for (int e0 = w_lo + NWIN * q; __any_sync(0xffffffffu, e0 < w_hi); e0 += 128) {
                            float en[NWIN];
#pragma unroll
                            for (int k = 0; k < NWIN; ++k) en[k] = 0.0f;
                            if (e0 < w_hi) {
                                const unsigned s0 = v.B + (unsigned)e0;
                                switch (Ts) {
                                    case 24: integrate_windows<24, NWIN>(ring_m, s0, umask, en); break;
                                    case 20: integrate_windows<20, NWIN>(ring_m, s0, umask, en); break;
                                    default: integrate_windows<40, NWIN>(ring_m, s0, umask, en); break;
                                }
                            }
                            constexpr int WPL = NWIN / M; /* windows whose products this lane forms */
                            float ft[WPL];
#pragma unroll
                            for (int k = 0; k < WPL; ++k) ft[k] = 0.0f;
#pragma unroll
                            for (int mm = 0; mm < M; ++mm) {
#pragma unroll
                                for (int k = 0; k < NWIN; ++k) {
                                    const float oe = __shfl_sync(0xffffffffu, en[k], q + QPC * mm);
                                    if (k / WPL == m) ft[k % WPL] += oe;
                                }
                            }
                            if (e0 < w_hi) {
#pragma unroll
                                for (int k = 0; k < WPL; k += 2) {
                                    const int e = e0 + m * WPL + k;
                                    const float4 ph = *reinterpret_cast<const float4 *>(phi_ft + e);
                                    *reinterpret_cast<float4 *>(prod + (e & (ER - 1))) =
                                        make_float4(ft[k] * ph.x, ft[k] * ph.y, ft[k + 1] * ph.z, ft[k + 1] * ph.w);
                                }
                            }
                        }
                    } else {
                        /* generic windows (any P, Ts, alignment): one window, all tones, per lane */
                        for (int i = w_lo + lane; i < w_hi; i += 32) {
                            float ft = 0.0f;
#pragma unroll
                            for (int m = 0; m < M; ++m) ft += window_energy(c.ring + m * RT, v.B + (unsigned)(i * step), rmask, Ts);
                            const float2 ph = phi_ft[i];
                            prod[i & (ER - 1)] = make_float2(ft * ph.x, ft * ph.y);
                        }
                    }
                };
                {
                    /* every worker warp w owns captures w, w + W, ...: per chunk it mixes the capture's samples and integrates
                       the windows they complete, so a capture needs no cross-warp barrier between the two (measured: splitting the
                       workers into mixing helpers and integrating owners, 13.3 ms, lost to this, 10.8 ms); the only hand-offs are
                       oscillator -> workers (named barrier) and workers -> timing warp (mbarrier) */
                    for (int ck = 0; ck < nchunks; ++ck) {
                        TR(40);
                        nbar_sync(1 + (ck & 3), osc_cnt);
                        TR(41);
                        for (int g = warp; g < G; g += W) {
                            if (!cap2(smem, cfg, g).cs->active) continue;
                            mix_chunk(g, ck);
                            TR(42);
                            if (ck >= 3) mbar_wait(&ctl->bar_tim[ck - 3], (par >> (ck - 3)) & 1u);
                            integrate_chunk(g, ck);
                        }
                        TR(44);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ctl->bar_int[ck]);
                    }
                }
            }
            par ^= nchunks >= 32 ? ~0u : ((1u << nchunks) - 1u);
            __syncthreads();
            TR(10);

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
                    int jj = cq * CKE;
                    for (; jj < jf; ++jj) { /* lead-in: phases only */
#pragma unroll
                        for (int m = 0; m < M; ++m) ph[m] = cmul(ph[m], d[m]);
                    }
                    const int jl = last - nold;
#pragma unroll 4
                    for (; jj < jl; ++jj) {
                        const float2 x = raw_sample(src, fmt, jj);
                        const int id = jj + nold;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            ph[m] = cmul(ph[m], d[m]);
                            const float2 fv = mixc(x, ph[m]);
                            const u64 v = pk(fv.x, fv.y);
                            if (id < olo + Ts) a[m] = add2(a[m], v);
                            if (id >= ohi) b[m] = add2(b[m], v);
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
            TR(14);
            /* the four serial per-frame sums of capture g and the estimates derived from them, lanes 0..3 of one warp */
            auto frame_sums = [&](int g) {
                const Cap2 c = cap2(smem, cfg, g);
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
                __syncwarp();
            };
            /* f_dc tail of this frame for the next frame's decisions (the ring keeps it until then); after every warp's
               decision lanes, which read the previous tail */
            auto copy_tail = [&]() {
                for (int t = tid; t < G * M * TAILN; t += nthreads) {
                    const int gm = t / TAILN, k = t - gm * TAILN;
                    const int g = gm / M, m = gm - g * M;
                    const Cap2 c = cap2(smem, cfg, g);
                    if (!c.cs->active) continue;
                    const unsigned pend = (unsigned)c.cs->pos + (unsigned)c.cs->nin;
                    c.tail[m * TAILN + k] = c.ring[m * RT + rsw(pend - (unsigned)TAILN + (unsigned)k, rmask)];
                }
            };
            /* measured on B200 (1 s captures, ms per launch, static / queue / queue in mixed order): 9 captures per CTA (36 FFT
               tasks on 16 warps: a ragged third round) 10.9 / 10.35 / 10.9; 8 captures per CTA (two full rounds) 9.66 / 9.89 /
               10.63.  So the queue is used when the FFT tasks do not fill whole rounds, FFT tasks first: serial tasks running
               beside FFT tasks only pay off in the ragged round (different code side by side costs instruction-cache misses) */
            if (cfg.k2.task_queue == 2 || (cfg.k2.task_queue == 1 && (G * tpc) % NF != 0)) {
                /* One queue for everything between the decisions and the next frame: the FFT tasks of frame f + 1 (its
                   samples were fetched a frame ahead and touch nothing the decisions use), then the per-frame sums of every
                   capture, then the tone picking of every capture.  A warp takes the next task whenever it is free, so the
                   serial tasks (sums: 4 lanes, picking: one warp) run beside the FFT tasks instead of after them.  Tasks are
                   handed out in order and wait only for lower-numbered ones (IIR turn, est_done) or for the decision lanes. */
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    atomicAdd(&ctl->decide_done, 1);
                }
                const int n_all = ctl->q_len;
                float2 *buf = warp < NW / 2 ? fft_slots + (size_t)warp * FSK2_FFT_SLOT_ELEMS(Ndft)
                                            : reinterpret_cast<float2 *>(cap2(smem, cfg, warp - NW / 2).ering + cfg.k2.ering_fft_off);
                for (;;) {
                    int t = 0;
                    if (lane == 0) t = atomicAdd(&ctl->task_next, 1);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t >= n_all) break;
                    const int code = ctl->queue[t], kind = code >> 12, qi = code & 0xfff;
                    if (kind == FSK2_Q_FFT) {
                        if (qi >= t_early) fft_tasks(fb ^ 1, qi, qi + 1, 0, 1, warp, f + 1); /* below t_early: done while streaming */
                    } else if (kind == FSK2_Q_SUMS) {
                        int tries = 0;
                        while (ld_acquire_shared(&ctl->decide_done) < NW) {
                            if (++tries > 16) __nanosleep(200);
                        }
                        frame_sums(qi);
                    } else {
                        const Cap2 c = cap2(smem, cfg, qi);
                        if (c.cs->fon) {
                            int tries = 0;
                            while (ld_acquire_shared(&c.cs->est_done) != f + 1) {
                                if (++tries > 16) __nanosleep(200);
                            }
                            pick_tones(cfg, c.cs, c.Sf + (size_t)(fb ^ 1) * Ndft, reinterpret_cast<float *>(buf), fb ^ 1, M, lane);
                        }
                    }
                }
                __syncthreads();
                TR(11);
                copy_tail();
            } else {
                /* A7 for frame f + 1 (all warps), from the staging buffer whose copy was issued at the end of the last
                   frame: touches nothing the decisions of frame f use */
                if (warp < NF) fft_tasks(fb ^ 1, t_early, G * tpc, warp, NF, warp, f + 1);
                __syncthreads();
                TR(11);
                copy_tail();
                if (warp < G) frame_sums(warp);
                pick_phase(fb ^ 1); /* tones of frame f + 1 from the spectrum the FFT tasks just updated */
            }
            __syncthreads();
            TR(12);

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
            TR(13);

            /* ---- per-frame record, advance, next frame's bookkeeping: lane g of warp 0 ---- */
            if (warp == 0) {
                if (lane < G) {
                    Fsk2CapShared *cs = cap2(smem, cfg, lane).cs;
                    if (cs->active) {
                        if (args.stats) {
                            pirip_frame_stats *st = &args.stats[(size_t)cs->cap * args.max_frames + (size_t)f];
                            for (int m = 0; m < 4; ++m) {
                                st->f_est[m] = m < M ? cs->f_est[fb][m] : 0.0f;
                                st->f2_est[m] = m < M ? cs->f2_est[fb][m] : 0.0f;
                            }
                            st->norm_rx_timing = cs->norm_rx_timing;
                            st->ppm = cs->ppm;
                            st->SNRest = cs->SNRest;
                            st->v_est = cs->v_est;
                            st->rx_sig_pow = cs->rx_sig_pow;
                            st->rx_nse_pow = cs->rx_nse_pow;
                            st->EbNodB = cs->EbNodB;
                            st->snr_est = cs->snr_est;
                            st->foff = cs->foff;
                            st->rx_timing = cs->rx_timing;
                            st->nin_next = cs->nin_next;
                            st->sample_offset = (int32_t)cs->pos;
                        }
                        cs->pos += (unsigned long long)cs->nin;
                        cs->nin = cs->nin_next;
                        cs->frame += 1;
                    }
                }
                __syncwarp();
                bookkeeping(f + 1);
            }
            __syncthreads();
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
    const int nw = cfg.k2.threads / 32;
    if (cfg.M == 2) {
        if (nw == 16) return launch_mn<2, 16>(cfg, args, n, stream, prep);
        if (nw == 20) return launch_mn<2, 20>(cfg, args, n, stream, prep);
        if (nw == 24) return launch_mn<2, 24>(cfg, args, n, stream, prep);
    } else {
        if (nw == 16) return launch_mn<4, 16>(cfg, args, n, stream, prep);
        if (nw == 20) return launch_mn<4, 20>(cfg, args, n, stream, prep);
        if (nw == 24) return launch_mn<4, 24>(cfg, args, n, stream, prep);
    }
    return cudaErrorInvalidValue;
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

#ifdef FSK2_TIMING
extern "C" int pirip_debug_trace(long long *out, size_t n) {
    if (n > 32u * FSK2_TRACE_N) n = 32u * FSK2_TRACE_N;
    return (int)cudaMemcpyFromSymbol(out, g_trace, n * sizeof(long long));
}
#endif
