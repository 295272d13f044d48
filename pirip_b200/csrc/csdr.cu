/*
 * csdr.cu — sample conditioning on sm_100a: u8 -> float, real-tap FIR on the complex stream,
 * decimation, optional s16 requantisation, fused in one kernel; plus the libcsdr-named host
 * entry points.  Contract: include/pirip_csdr.h.
 *
 * Algorithm ([U ha7ilm/csdr libcsdr.c::convert_u8_f / fir_decimate_cc / convert_f_s16], call
 * site [R README.md:109]):  y[k] = sum_t x[D k + t] * taps[t], one continuous causal FIR
 * anchored at stream sample 0, accumulated in tap order in float (no FMA contraction, so the
 * result is bit-identical to the scalar loop).
 *
 * Three kernels, one arithmetic:
 *  fir_decimate_u8_tma   u8 input, D = 45 / 30 with 79 taps, 16-byte aligned streams: every sample read and converted
 *                        once (rotating accumulators), lane-private conversion table, taps as constants, raw bytes
 *                        staged per warp by 3-D TMA tensor copies.  The default for pirip's pipelines.
 *  fir_decimate_u8_fast  the same compute with the raw bytes loaded straight into registers (no tensor map needed).
 *  fir_decimate_kernel   generic (any D / L / alignment, f32 input): one CTA produces TILE consecutive outputs of one
 *                        stream from a float2 tile staged in shared memory, one thread per output.
 * HBM traffic is the algorithmic 2 B per input sample (+ L - D samples of halo per run / tile) and 8/D (f32) or
 * 4/D (s16) B of output.
 */
#include <cuda.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pirip_csdr.h"
#include "host_util.h"

namespace {

constexpr int kTile = 128;

struct DecimArgs {
    const unsigned char *in; /* u8 IQ or complexf, device */
    size_t in_stride_bytes;
    size_t n_in;   /* samples per stream */
    size_t n_out;  /* outputs per stream */
    int D, L;
    const float *taps; /* device */
    const float *lut;  /* device, 256 exact convert_u8_f values */
    float2 *out_f32;
    short2 *out_s16;
    size_t out_stride;
};

template <bool U8IN>
__global__ void __launch_bounds__(kTile) fir_decimate_kernel(const __grid_constant__ DecimArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *taps = reinterpret_cast<float *>(smem_raw);
    float *lut = taps + ((a.L + 3) & ~3);
    float2 *xs = reinterpret_cast<float2 *>(lut + 256);
    const int tid = threadIdx.x;
    const size_t stream = blockIdx.y;
    const size_t k0 = (size_t)blockIdx.x * kTile;
    if (k0 >= a.n_out) return;
    const int nk = (int)((a.n_out - k0) < (size_t)kTile ? (a.n_out - k0) : (size_t)kTile);
    const int count = a.D * (nk - 1) + a.L; /* input samples this tile needs */
    const size_t s0 = k0 * (size_t)a.D;

    for (int i = tid; i < a.L; i += kTile) taps[i] = a.taps[i];
    if (U8IN) {
        for (int i = tid; i < 256; i += kTile) lut[i] = a.lut[i];
        __syncthreads();
        const unsigned char *src = a.in + stream * a.in_stride_bytes + 2 * s0;
        const int mis = (int)((reinterpret_cast<size_t>(src) >> 1) & 1);
        if (mis && tid == 0) xs[0] = make_float2(lut[src[0]], lut[src[1]]);
        const int npair = (count - mis) >> 1;
        const unsigned *s4 = reinterpret_cast<const unsigned *>(src + 2 * mis);
        for (int p = tid; p < npair; p += kTile) {
            const unsigned w = __ldg(s4 + p);
            float2 *dst = xs + mis + 2 * p;
            dst[0] = make_float2(lut[w & 0xff], lut[(w >> 8) & 0xff]);
            dst[1] = make_float2(lut[(w >> 16) & 0xff], lut[w >> 24]);
        }
        if (((count - mis) & 1) && tid == kTile - 1) {
            const int j = count - 1;
            xs[j] = make_float2(lut[src[2 * j]], lut[src[2 * j + 1]]);
        }
    } else {
        const float2 *src = reinterpret_cast<const float2 *>(a.in + stream * a.in_stride_bytes) + s0;
        for (int i = tid; i < count; i += kTile) xs[i] = src[i];
    }
    __syncthreads();
    if (tid < nk) {
        const float2 *x = xs + tid * a.D;
        float acci = 0.0f, accq = 0.0f;
#pragma unroll 4
        for (int t = 0; t < a.L; ++t) {
            const float2 v = x[t];
            const float h = taps[t];
            acci += v.x * h;
            accq += v.y * h;
        }
        const size_t o = stream * a.out_stride + k0 + tid;
        if (a.out_f32) a.out_f32[o] = make_float2(acci, accq);
        if (a.out_s16) {
            /* convert_f_s16: (short)(in * SHRT_MAX), C truncation toward zero */
            a.out_s16[o] = make_short2((short)__float2int_rz(acci * 32767.0f), (short)__float2int_rz(accq * 32767.0f));
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Fast path for u8 input and the decimator shapes of pirip's pipelines (D = 45 / 30 with the default 79 taps,
 * [R README.md:109,172], [R script/frame_repeater:36]): the stage is HBM-class work (2 B in per sample, 7 flops),
 * so the kernel is built around instruction count and shared-memory bandwidth, not around a staged float tile.
 *
 *  - One thread owns NBLK consecutive outputs of one stream and walks its D*NBLK (+ halo) input samples ONCE, in
 *    order: sample r of input block j is tap r of output j, tap D + r of output j - 1, ... so every sample is
 *    loaded and converted once and feeds ceil(L / D) rotating accumulators (the generic kernel reads every staged
 *    sample L / D = 1.76 times from shared memory).  Each output still accumulates its products in tap order from
 *    zero with separate multiply and add, so the floats are those of the scalar loop.
 *  - Raw bytes come straight from global memory into registers (4-byte loads; a thread's words are contiguous, so
 *    every 32-byte sector is fetched from HBM once and re-hit in L1); there is no staged copy of the samples at all.
 *  - convert_u8_f is a table lookup.  The table holds the 256 exact values in 32 lane-private columns (row stride
 *    256 B, table aligned to 64 KB in the shared window), so a lookup is conflict-free for any data and its address
 *    is ONE byte-permute: {table base | lane column} with the sample byte dropped into address bits 8-15.
 *  - Taps are kernel parameters: with the block loop unrolled every tap is an immediate constant-bank operand.
 *  Per input sample: 0.5 LDG + 2 PRMT + 2 LDS + 7.0 FMUL/FADD, nothing else.
 */
constexpr int kFastThreads = 512;
constexpr int kFastMaxTaps = 96;
constexpr unsigned kFastSmem = 2 * 65536 - 1024; /* room for a 64 KB table on a 64 KB boundary of the shared window */

struct DecimFastArgs {
    const unsigned char *in;
    size_t in_stride_bytes;
    size_t n_out;      /* outputs per stream */
    size_t n_streams;
    const float *lut;  /* device, 256 exact convert_u8_f values */
    float2 *out_f32;
    short2 *out_s16;
    size_t out_stride;
    unsigned dyn_smem;
    int lockstep; /* TMA kernel: CTA-wide barrier before every pipeline step */
    float taps[kFastMaxTaps];
};

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
/* table address of byte K of word w for this lane: bytes {base.0, w.K, base.2, base.3} */
template <int K>
__device__ __forceinline__ unsigned lut_addr(unsigned lane_base, unsigned w) {
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lane_base), "r"(w), "n"(0x3200 | ((4 + K) << 4)));
    return r;
}

template <bool F32, bool S16>
__device__ __forceinline__ void emit_output(const DecimFastArgs &a, size_t o, float2 acc) {
    if (F32) a.out_f32[o] = acc;
    /* convert_f_s16: (short)(in * SHRT_MAX), C truncation toward zero */
    if (S16) a.out_s16[o] = make_short2((short)__float2int_rz(acc.x * 32767.0f), (short)__float2int_rz(acc.y * 32767.0f));
}

/* accumulate input sample (xi, xq) = sample r of an input block into the outputs it is a tap of: tap aa D + r of the
   output aa blocks back, whose accumulator is (jbm - aa) mod NA with jbm = block index mod NA */
template <int D, int L, int NA>
__device__ __forceinline__ void fir_feed(const DecimFastArgs &a, float2 (&acc)[NA], int jbm, int r, float xi, float xq) {
#pragma unroll
    for (int aa = 0; aa < NA; ++aa) {
        const int t = aa * D + r;
        if (t < L) {
            const int ai = (jbm - aa + 4 * NA) % NA;
            acc[ai].x += xi * a.taps[t];
            acc[ai].y += xq * a.taps[t];
        }
    }
}

/* the 8 * SUBC samples of pipeline step `sub` of a body that starts at block j of the group */
template <int D, int L, int NA, int SUBC, bool F32, bool S16>
__device__ __forceinline__ void fir_substep(const DecimFastArgs &a, float2 (&acc)[NA], const unsigned (&cur)[4 * SUBC], const int sub,
                                            const int j, const size_t obase, const unsigned lb, const bool valid) {
#pragma unroll
    for (int i = 0; i < 8 * SUBC; ++i) {
        const int s = sub * 8 * SUBC + i; /* sample of the body */
        const int u = s / D, r = s - u * D;
        const unsigned ww = cur[i >> 1];
        const float xi = lds_f32((i & 1) ? lut_addr<2>(lb, ww) : lut_addr<0>(lb, ww));
        const float xq = lds_f32((i & 1) ? lut_addr<3>(lb, ww) : lut_addr<1>(lb, ww));
        fir_feed<D, L, NA>(a, acc, u % NA, r, xi, xq);
        if (r == D - 1) {
            /* the output whose last taps were in this block: NA - 1 blocks back */
            const int done = (u + 1) % NA;
            const int o = j + u - (NA - 1);
            if (o >= 0 && valid) emit_output<F32, S16>(a, obase + (size_t)o, acc[done]);
            acc[done] = make_float2(0.0f, 0.0f);
        }
    }
}

/* the L - D samples after a group finish its last NA - 1 outputs; loaded without touching a byte beyond the last
   sample the stream is known to hold */
template <int D, int L, int NA, int NBLK, bool F32, bool S16>
__device__ __forceinline__ void fir_halo(const DecimFastArgs &a, float2 (&acc)[NA], const uint4 *src, const size_t obase, const unsigned lb) {
    constexpr int HS = L - D;
    constexpr int HW = (HS + 1) / 2; /* words */
    unsigned w[HW];
    const unsigned *s32 = reinterpret_cast<const unsigned *>(src);
#pragma unroll
    for (int c = 0; c < HW / 4; ++c) {
        const uint4 v = __ldg(src + c);
        w[4 * c] = v.x, w[4 * c + 1] = v.y, w[4 * c + 2] = v.z, w[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int i = (HW / 4) * 4; i < HW; ++i) {
        if (i == HW - 1 && (HS & 1))
            w[i] = *reinterpret_cast<const unsigned short *>(s32 + i);
        else
            w[i] = __ldg(s32 + i);
    }
#pragma unroll
    for (int sI = 0; sI < HS; ++sI) {
        const int e = sI / D, r = sI - e * D;
        const unsigned ww = w[sI >> 1];
        const float xi = lds_f32((sI & 1) ? lut_addr<2>(lb, ww) : lut_addr<0>(lb, ww));
        const float xq = lds_f32((sI & 1) ? lut_addr<3>(lb, ww) : lut_addr<1>(lb, ww));
#pragma unroll
        for (int aa = e + 1; aa < NA; ++aa) {
            const int t = aa * D + r;
            if (t < L) {
                const int ai = ((NBLK + e) - aa + 4 * NA) % NA;
                acc[ai].x += xi * a.taps[t];
                acc[ai].y += xq * a.taps[t];
            }
        }
        if (r == D - 1 || sI == HS - 1) emit_output<F32, S16>(a, obase + (size_t)(NBLK + e - (NA - 1)), acc[((NBLK + e) + 1) % NA]);
    }
}

/* the n_out % NBLK outputs at the end of every stream: plain tap loop, one output per thread */
template <int D, int L, int NBLK, bool F32, bool S16>
__device__ __forceinline__ void fir_remainder(const DecimFastArgs &a, const unsigned lb, const size_t gps, const size_t tid_global,
                                              const size_t nthreads_global) {
    const size_t rem = a.n_out - gps * NBLK;
    for (size_t i = tid_global; i < rem * a.n_streams; i += nthreads_global) {
        const size_t stream = i / rem, k = gps * NBLK + (i - stream * rem);
        const unsigned short *x = reinterpret_cast<const unsigned short *>(a.in + stream * a.in_stride_bytes) + k * (size_t)D;
        float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 1
        for (int t = 0; t < L; ++t) {
            const unsigned ww = x[t];
            const float h = a.taps[t];
            acc.x += lds_f32(lut_addr<0>(lb, ww)) * h;
            acc.y += lds_f32(lut_addr<1>(lb, ww)) * h;
        }
        emit_output<F32, S16>(a, stream * a.out_stride + k, acc);
    }
}

/* fill the lane-private conversion table; returns this lane's column base */
__device__ __forceinline__ unsigned fir_lut_setup(const DecimFastArgs &a, const unsigned sbase, const int tid, const int nthreads) {
    const unsigned lut0 = (sbase + 65535u) & ~65535u;
    if (lut0 + 65536u > sbase + a.dyn_smem) __trap();
    for (int i = tid; i < 256 * 32; i += nthreads) {
        const float v = a.lut[i >> 5];
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(lut0 + ((unsigned)(i >> 5) << 8) + ((unsigned)(i & 31) << 2)), "f"(v) : "memory");
    }
    return lut0;
}

template <int D, int L, int NBLK, bool F32, bool S16>
__global__ void __launch_bounds__(kFastThreads, 1) fir_decimate_u8_fast(const __grid_constant__ DecimFastArgs a) {
    constexpr int NA = (L + D - 1) / D; /* outputs a sample contributes to = live accumulators */
    /* blocks per unrolled body: a whole number of 16-byte chunks (8 samples) and of accumulator rotations */
    constexpr int PB16 = (D % 8 == 0) ? 1 : (D % 4 == 0) ? 2 : (D % 2 == 0) ? 4 : 8;
    constexpr int U = (PB16 % NA == 0) ? PB16 : PB16 * NA;
    constexpr int CB = U * D / 8; /* 16-byte chunks per body */
    constexpr int SUBC = (CB % 9 == 0) ? 9 : (CB % 5 == 0) ? 5 : (CB % 3 == 0) ? 3 : 1; /* chunks per pipeline step */
    constexpr int NSUB = CB / SUBC;
    constexpr int HS = L - D; /* halo samples after the group */
    static_assert(NBLK % U == 0, "group must be whole bodies");
    static_assert(L <= kFastMaxTaps && L > D, "tap count");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned sbase = smem_addr(smem_raw);
    const unsigned lut0 = fir_lut_setup(a, sbase, tid, kFastThreads);
    __syncthreads();
    const unsigned lb = lut0 + 4u * (unsigned)lane;

    const size_t gps = a.n_out / NBLK; /* whole groups per stream */
    const size_t total = gps * a.n_streams;
    for (size_t gid = (size_t)blockIdx.x * kFastThreads + tid; gid < total; gid += (size_t)gridDim.x * kFastThreads) {
        const size_t stream = gid / gps, grp = gid - stream * gps;
        const size_t k0 = grp * NBLK;
        const uint4 *src = reinterpret_cast<const uint4 *>(a.in + stream * a.in_stride_bytes + 2 * (size_t)D * k0);
        const size_t obase = stream * a.out_stride + k0;
        float2 acc[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) acc[i] = make_float2(0.0f, 0.0f);
        unsigned cur[4 * SUBC], nxt[4 * SUBC];
#pragma unroll
        for (int c = 0; c < SUBC; ++c) {
            const uint4 v = __ldg(src + c);
            cur[4 * c] = v.x, cur[4 * c + 1] = v.y, cur[4 * c + 2] = v.z, cur[4 * c + 3] = v.w;
        }
#pragma unroll 1
        for (int j = 0; j < NBLK; j += U) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                /* the next step's chunks are in flight while this step's samples are consumed */
                if (sub + 1 < NSUB || j + U < NBLK) {
#pragma unroll
                    for (int c = 0; c < SUBC; ++c) {
                        const uint4 v = __ldg(src + (sub + 1) * SUBC + c);
                        nxt[4 * c] = v.x, nxt[4 * c + 1] = v.y, nxt[4 * c + 2] = v.z, nxt[4 * c + 3] = v.w;
                    }
                }
                fir_substep<D, L, NA, SUBC, F32, S16>(a, acc, cur, sub, j, obase, lb, true);
#pragma unroll
                for (int c = 0; c < 4 * SUBC; ++c) cur[c] = nxt[c];
            }
            src += CB;
        }
        fir_halo<D, L, NA, NBLK, F32, S16>(a, acc, src, obase, lb);
    }
    fir_remainder<D, L, NBLK, F32, S16>(a, lb, gps, (size_t)blockIdx.x * kFastThreads + tid, (size_t)gridDim.x * kFastThreads);
}

/* ---- the same arithmetic with the raw bytes staged by TMA ------------------------------------------------------
   The register-streaming kernel above is bound by L1TEX tag look-ups (ncu: l1tex throughput 89-99 %): the lanes of a warp
   own runs 1440 B apart, so every 16-byte load touches 32 different lines.  Here a warp owns 32 consecutive runs of one
   stream and fetches pipeline step s of all of them with ONE 3-D tensor copy (coordinates {byte in run, run, stream},
   box = 144 B x 32 runs; TMA wants the box to start on a 16-byte boundary of global memory, hence steps of whole
   16-byte chunks rather than whole 90-byte blocks) into a 4.5 KB buffer of its own, two buffers per warp, completion on
   an mbarrier: no L1 look-ups.  Lane l reads its 144 B with nine 16-byte shared loads (row pitch 9 chunks, odd:
   conflict-free per quarter-warp) and feeds the same fir_substep.  Runs beyond the end of a stream are zero-filled by
   the tensor map and their outputs not written.  With a.lockstep the warps of a CTA meet at a named barrier before
   every step, so that they walk the 69 KB unrolled body together and share its instruction-cache lines. */
constexpr int kStepBytes = 144;          /* bytes of one run per pipeline step (9 chunks of 16) */
constexpr int kTmaBuf = 32 * kStepBytes; /* one staging buffer: 32 runs */
constexpr int kTmaWarps = kFastThreads / 32;
/* 64 KB table on a 64 KB boundary of the shared window + 32 buffers on either side of it + mbarriers; with the window's
   first KB reserved that is 213 248 B, the rest is slack for other bases (the kernel traps if the pieces do not fit) */
constexpr unsigned kTmaSmem = 220 * 1024;

__device__ __forceinline__ void mbar_init1(unsigned bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

template <int D, int L, int NBLK, bool F32, bool S16>
__global__ void __launch_bounds__(kFastThreads, 1)
fir_decimate_u8_tma(const __grid_constant__ DecimFastArgs a, const __grid_constant__ CUtensorMap tmap) {
    constexpr int NA = (L + D - 1) / D;
    constexpr int PB16 = (D % 8 == 0) ? 1 : (D % 4 == 0) ? 2 : (D % 2 == 0) ? 4 : 8;
    constexpr int U = (PB16 % NA == 0) ? PB16 : PB16 * NA;
    constexpr int CB = U * D / 8;
    constexpr int SUBC = kStepBytes / 16;
    constexpr int NSUB = CB / SUBC;
    constexpr int NSTEPS = (NBLK / U) * NSUB; /* pipeline steps per run */
    static_assert(CB % SUBC == 0 && NBLK % U == 0 && NSTEPS >= 2, "a run must be whole pipeline steps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned sbase = smem_addr(smem_raw);
    const unsigned lut0 = fir_lut_setup(a, sbase, tid, kFastThreads);
    const unsigned lb = lut0 + 4u * (unsigned)lane;
    /* staging buffers: below the table as far as they fit, the rest above it; mbarriers in the last 256 bytes */
    const unsigned bars = sbase + a.dyn_smem - 256u;
    const unsigned lo0 = (sbase + 127u) & ~127u;
    const unsigned below = (lut0 - lo0) / kTmaBuf;
    unsigned buf0, buf1, bar0, bar1;
    {
        const unsigned i0 = 2u * (unsigned)warp, i1 = i0 + 1u;
        buf0 = i0 < below ? lo0 + i0 * kTmaBuf : lut0 + 65536u + (i0 - below) * kTmaBuf;
        buf1 = i1 < below ? lo0 + i1 * kTmaBuf : lut0 + 65536u + (i1 - below) * kTmaBuf;
        bar0 = bars + 8u * i0;
        bar1 = bars + 8u * i1;
        if (buf0 + kTmaBuf > bars || buf1 + kTmaBuf > bars) __trap();
        if (lane == 0) {
            mbar_init1(bar0);
            mbar_init1(bar1);
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned par0 = 0, par1 = 0;
    const bool lockstep = a.lockstep != 0;

    const size_t gps = a.n_out / NBLK;     /* whole runs (groups of NBLK outputs) per stream */
    const size_t tps = (gps + 31) / 32;    /* warp tasks per stream */
    const size_t ntasks = tps * a.n_streams;
    /* the task loop is uniform over the CTA (a warp without a task only keeps the barriers company) */
    for (size_t base = (size_t)blockIdx.x * kTmaWarps; base < ntasks; base += (size_t)gridDim.x * kTmaWarps) {
        const size_t task = base + warp;
        const bool live = task < ntasks;
        const size_t stream = live ? task / tps : 0, wt = live ? task - stream * tps : 0;
        const size_t grp = wt * 32 + (size_t)lane;
        const bool valid = live && grp < gps;
        const size_t k0 = grp * NBLK;
        const size_t obase = stream * a.out_stride + k0;
        const int row0 = (int)(wt * 32), st = (int)stream;
        if (live && lane == 0) {
            mbar_expect(bar0, kTmaBuf);
            tma_load_3d(buf0, &tmap, 0, row0, st, bar0);
            mbar_expect(bar1, kTmaBuf);
            tma_load_3d(buf1, &tmap, kStepBytes, row0, st, bar1);
        }
        float2 acc[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) acc[i] = make_float2(0.0f, 0.0f);
#pragma unroll 1
        for (int j = 0; j < NBLK; j += U) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                const int step = (j / U) * NSUB + sub;
                const bool odd = (step & 1) != 0;
                if (lockstep) asm volatile("bar.sync 1, %0;" ::"n"(kFastThreads) : "memory");
                if (live) {
                    const unsigned bar = odd ? bar1 : bar0, buf = odd ? buf1 : buf0;
                    if (odd) {
                        mbar_wait_parity(bar, par1);
                        par1 ^= 1u;
                    } else {
                        mbar_wait_parity(bar, par0);
                        par0 ^= 1u;
                    }
                    unsigned cur[4 * SUBC];
                    const unsigned mine = buf + (unsigned)lane * kStepBytes;
#pragma unroll
                    for (int c = 0; c < SUBC; ++c)
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(cur[4 * c]), "=r"(cur[4 * c + 1]), "=r"(cur[4 * c + 2]), "=r"(cur[4 * c + 3])
                                     : "r"(mine + 16u * c)
                                     : "memory");
                    fir_substep<D, L, NA, SUBC, F32, S16>(a, acc, cur, sub, j, obase, lb, valid);
                    /* every lane has consumed its words: the buffer can take the step after next */
                    __syncwarp();
                    if (lane == 0 && step + 2 < NSTEPS) {
                        mbar_expect(bar, kTmaBuf);
                        tma_load_3d(buf, &tmap, kStepBytes * (step + 2), row0, st, bar);
                    }
                }
            }
        }
        if (valid) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.in + stream * a.in_stride_bytes + 2 * (size_t)D * (k0 + NBLK));
            fir_halo<D, L, NA, NBLK, F32, S16>(a, acc, src, obase, lb);
        }
        __syncwarp();
    }
    fir_remainder<D, L, NBLK, F32, S16>(a, lb, gps, (size_t)blockIdx.x * kFastThreads + tid, (size_t)gridDim.x * kFastThreads);
}

template <int D, int L, int NBLK>
cudaError_t fast_launch(const DecimFastArgs &a, const CUtensorMap *tmap, bool prepare_only, int sms, cudaStream_t stream) {
    const bool f = a.out_f32 != nullptr, s = a.out_s16 != nullptr;
#define PIRIP_FAST_CASE(F, S)                                                                                              \
    do {                                                                                                                   \
        if (prepare_only) {                                                                                                \
            cudaError_t e = cudaFuncSetAttribute(fir_decimate_u8_fast<D, L, NBLK, F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem); \
            if (e != cudaSuccess) return e;                                                                                \
            e = cudaFuncSetAttribute(fir_decimate_u8_tma<D, L, NBLK, F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmem); \
            if (e != cudaSuccess) return e;                                                                                \
        } else if (f == F && s == S) {                                                                                     \
            if (tmap)                                                                                                      \
                fir_decimate_u8_tma<D, L, NBLK, F, S><<<sms, kFastThreads, a.dyn_smem, stream>>>(a, *tmap);                  \
            else                                                                                                           \
                fir_decimate_u8_fast<D, L, NBLK, F, S><<<sms, kFastThreads, a.dyn_smem, stream>>>(a);                        \
            return cudaGetLastError();                                                                                     \
        }                                                                                                                  \
    } while (0)
    PIRIP_FAST_CASE(true, true);
    PIRIP_FAST_CASE(true, false);
    PIRIP_FAST_CASE(false, true);
#undef PIRIP_FAST_CASE
    return cudaSuccess;
}

/* which (D, L) shapes have a fast instantiation; *nblk = outputs per thread (run = 2 D nblk bytes) */
cudaError_t fast_dispatch(int D, int L, const DecimFastArgs &a, const CUtensorMap *tmap, bool prepare_only, int sms, cudaStream_t stream,
                          bool *handled, int *nblk) {
    *handled = true;
    if (D == 45 && L == 79) {
        *nblk = 16;
        return fast_launch<45, 79, 16>(a, tmap, prepare_only, sms, stream);
    }
    if (D == 30 && L == 79) {
        *nblk = 12;
        return fast_launch<30, 79, 12>(a, tmap, prepare_only, sms, stream);
    }
    *handled = false;
    *nblk = 0;
    return cudaSuccess;
}

/* cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked) */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

/* tensor view of the batch for the TMA kernel: {byte in run, run, stream}, box = one pipeline step of 32 runs */
bool make_run_map(CUtensorMap *map, const void *in, size_t in_stride_bytes, size_t n_streams, size_t runs_per_stream, int run_bytes,
                  int box_bytes) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || runs_per_stream == 0) return false;
    if (runs_per_stream > 0xffffffffull || n_streams > 0xffffffffull || in_stride_bytes >= (1ull << 40)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)run_bytes, (cuuint64_t)runs_per_stream, (cuuint64_t)n_streams};
    const cuuint64_t strides[2] = {(cuuint64_t)run_bytes, (cuuint64_t)in_stride_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_bytes, 32u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t out_count(size_t n, int D, int L) { return n >= (size_t)L ? (n - (size_t)L) / (size_t)D + 1 : 0; }

}  // namespace

struct pirip_decim_plan {
    int D, L, device;
    float *d_taps = nullptr, *d_lut = nullptr;
    int smem_u8 = 0, smem_f32 = 0;
    bool fast = false; /* a fir_decimate_u8_fast / _tma instantiation exists for (D, L) */
    int fast_nblk = 0; /* outputs per thread of that instantiation */
    int sms = 0;
    std::vector<float> taps;
    /* reusable buffers of the host-buffer entry point */
    unsigned char *d_in = nullptr;
    float2 *d_out = nullptr;
    short2 *d_out16 = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_out16 = 0;
};

static cudaError_t decim_launch(pirip_decim_plan *p, bool u8in, const void *d_in, size_t n_streams, size_t n_in,
                                size_t in_stride_bytes, complexf *d_out_f32, int16_t *d_out_s16, size_t out_stride,
                                cudaStream_t stream) {
    DecimArgs a;
    a.in = (const unsigned char *)d_in;
    a.in_stride_bytes = in_stride_bytes;
    a.n_in = n_in;
    a.n_out = out_count(n_in, p->D, p->L);
    a.D = p->D;
    a.L = p->L;
    a.taps = p->d_taps;
    a.lut = p->d_lut;
    a.out_f32 = (float2 *)d_out_f32;
    a.out_s16 = (short2 *)d_out_s16;
    a.out_stride = out_stride;
    if (a.n_out == 0 || n_streams == 0) return cudaSuccess;
    /* u8 streams on 16-byte boundaries take the register-streaming kernel */
    if (u8in && p->fast && (reinterpret_cast<size_t>(d_in) & 15) == 0 && (in_stride_bytes & 15) == 0 && !getenv("PIRIP_DECIM_GENERIC")) {
        DecimFastArgs f;
        f.in = a.in;
        f.in_stride_bytes = in_stride_bytes;
        f.n_out = a.n_out;
        f.n_streams = n_streams;
        f.lut = p->d_lut;
        f.out_f32 = a.out_f32;
        f.out_s16 = a.out_s16;
        f.out_stride = out_stride;
        memset(f.taps, 0, sizeof(f.taps));
        memcpy(f.taps, p->taps.data(), sizeof(float) * p->taps.size());
        /* raw bytes staged by TMA when the tensor view can be encoded; register-streaming loads otherwise */
        CUtensorMap tmap;
        const bool use_tma = !getenv("PIRIP_DECIM_NO_TMA") &&
                             make_run_map(&tmap, d_in, in_stride_bytes, n_streams, a.n_out / (size_t)p->fast_nblk, 2 * p->D * p->fast_nblk, kStepBytes);
        f.dyn_smem = use_tma ? kTmaSmem : kFastSmem;
        f.lockstep = getenv("PIRIP_DECIM_LOCKSTEP") ? atoi(getenv("PIRIP_DECIM_LOCKSTEP")) : 1; /* measured: 0.323 vs 0.332 ms */
        bool handled = false;
        int nblk = 0;
        cudaError_t e = fast_dispatch(p->D, p->L, f, use_tma ? &tmap : nullptr, false, p->sms, stream, &handled, &nblk);
        if (handled) {
            pirip::count_launch();
            return e;
        }
    }
    dim3 grid((unsigned)((a.n_out + kTile - 1) / kTile), (unsigned)n_streams);
    if (u8in)
        fir_decimate_kernel<true><<<grid, kTile, p->smem_u8, stream>>>(a);
    else
        fir_decimate_kernel<false><<<grid, kTile, p->smem_f32, stream>>>(a);
    pirip::count_launch();
    return cudaGetLastError();
}

extern "C" pirip_decim_plan *pirip_decim_plan_create(int decimation, const float *taps, int taps_length, int device) {
    if (decimation < 1 || taps_length < 1 || !taps) {
        pirip::set_error("invalid decimator configuration");
        return nullptr;
    }
    pirip_decim_plan *p = new pirip_decim_plan();
    p->D = decimation;
    p->L = taps_length;
    p->device = device;
    const int tile_samples = decimation * (kTile - 1) + taps_length;
    p->smem_u8 = (((taps_length + 3) & ~3) + 256) * 4 + tile_samples * 8;
    p->smem_f32 = p->smem_u8;
    if (p->smem_u8 > 227 * 1024) {
        pirip::set_error("decimation * tile + taps does not fit shared memory");
        delete p;
        return nullptr;
    }
    float lut[256];
    for (int b = 0; b < 256; ++b) lut[b] = ((float)b) / (UCHAR_MAX / 2.0) - 1.0; /* convert_u8_f */
    p->taps.assign(taps, taps + taps_length);
    if (!pirip::cuda_ok(cudaSetDevice(device), "cudaSetDevice") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_taps, sizeof(float) * (size_t)taps_length), "cudaMalloc taps") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_lut, sizeof(lut)), "cudaMalloc lut") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_taps, taps, sizeof(float) * (size_t)taps_length, cudaMemcpyHostToDevice), "copy taps") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice), "copy lut") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(fir_decimate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_u8), "smem attr") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(fir_decimate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_f32), "smem attr")) {
        pirip_decim_plan_destroy(p);
        return nullptr;
    }
    {
        DecimFastArgs f;
        memset(&f, 0, sizeof(f));
        f.dyn_smem = kFastSmem;
        bool handled = false;
        cudaDeviceProp prop;
        if (!pirip::cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties") ||
            !pirip::cuda_ok(fast_dispatch(p->D, p->L, f, nullptr, true, 0, nullptr, &handled, &p->fast_nblk), "smem attr (fast decimator)")) {
            pirip_decim_plan_destroy(p);
            return nullptr;
        }
        p->fast = handled;
        p->sms = prop.multiProcessorCount;
    }
    return p;
}

extern "C" void pirip_decim_plan_destroy(pirip_decim_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_taps);
    cudaFree(p->d_lut);
    cudaFree(p->d_in);
    cudaFree(p->d_out);
    cudaFree(p->d_out16);
    delete p;
}

extern "C" size_t pirip_decim_out_count(const pirip_decim_plan *p, size_t n) { return p ? out_count(n, p->D, p->L) : 0; }

extern "C" int pirip_decim_u8_batch_dev(pirip_decim_plan *p, const uint8_t *d_iq_u8, size_t n_streams,
                                        size_t samples_per_stream, size_t in_stride_bytes, complexf *d_out_f32,
                                        int16_t *d_out_s16, size_t out_stride, void *stream) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (!d_iq_u8 && n_streams) return pirip::fail(-2, "input is NULL");
    if (in_stride_bytes < 2 * samples_per_stream) return pirip::fail(-4, "in_stride_bytes smaller than one stream");
    if (out_stride < out_count(samples_per_stream, p->D, p->L)) return pirip::fail(-4, "out_stride smaller than the output count");
    if (!pirip::cuda_ok(cudaSetDevice(p->device), "cudaSetDevice")) return -10;
    if (!pirip::cuda_ok(decim_launch(p, true, d_iq_u8, n_streams, samples_per_stream, in_stride_bytes, d_out_f32, d_out_s16,
                                     out_stride, (cudaStream_t)stream),
                        "fir_decimate_kernel launch"))
        return -11;
    return 0;
}

template <typename T>
static bool grow(T *&ptr, size_t &cap, size_t need, const char *what) {
    if (need <= cap && ptr) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    if (!pirip::cuda_ok(cudaMalloc(&ptr, need ? need : 16), what)) return false;
    cap = need;
    return true;
}

extern "C" int pirip_decim_u8_batch(pirip_decim_plan *p, const uint8_t *iq_u8, size_t n_streams, size_t samples_per_stream,
                                    complexf *out_f32, int16_t *out_s16) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (n_streams == 0) return 0;
    if (!iq_u8) return pirip::fail(-2, "input is NULL");
    if (!pirip::cuda_ok(cudaSetDevice(p->device), "cudaSetDevice")) return -10;
    const size_t nout = out_count(samples_per_stream, p->D, p->L);
    const size_t in_bytes = 2 * samples_per_stream;
    const size_t stride = (in_bytes + 15) / 16 * 16;
    if (!grow(p->d_in, p->cap_in, stride * n_streams, "cudaMalloc in")) return -12;
    if (out_f32 && !grow(p->d_out, p->cap_out, sizeof(float2) * nout * n_streams, "cudaMalloc out")) return -12;
    if (out_s16 && !grow(p->d_out16, p->cap_out16, sizeof(short2) * nout * n_streams, "cudaMalloc out16")) return -12;
    if (!pirip::cuda_ok(cudaMemcpy2D(p->d_in, stride, iq_u8, in_bytes, in_bytes, n_streams, cudaMemcpyHostToDevice), "H2D")) return -12;
    int rc = pirip_decim_u8_batch_dev(p, p->d_in, n_streams, samples_per_stream, stride, out_f32 ? (complexf *)p->d_out : nullptr,
                                      out_s16 ? (int16_t *)p->d_out16 : nullptr, nout, nullptr);
    if (rc) return rc;
    if (out_f32 && nout && !pirip::cuda_ok(cudaMemcpy(out_f32, p->d_out, sizeof(float2) * nout * n_streams, cudaMemcpyDeviceToHost), "D2H")) return -13;
    if (out_s16 && nout && !pirip::cuda_ok(cudaMemcpy(out_s16, p->d_out16, sizeof(short2) * nout * n_streams, cudaMemcpyDeviceToHost), "D2H")) return -13;
    return pirip::cuda_ok(cudaDeviceSynchronize(), "fir_decimate_kernel") ? 0 : -13;
}

/* ---------------------------------------------------------------- libcsdr-named host functions */

extern "C" void convert_u8_f(unsigned char *input, float *output, int input_size) {
    for (int i = 0; i < input_size; ++i) output[i] = ((float)input[i]) / (UCHAR_MAX / 2.0) - 1.0;
}

extern "C" void convert_f_s16(float *input, short *output, int input_size) {
    for (int i = 0; i < input_size; ++i) output[i] = (short)(input[i] * SHRT_MAX);
}

extern "C" int firdes_filter_len(float transition_bw) {
    int result = 4.0 / transition_bw;
    if (result % 2 == 0) result++; /* symmetric FIR: odd number of taps */
    return result;
}

/* windowed-sinc low-pass, Hamming window, normalised to unit DC gain; csdr's PI is a float */
extern "C" void firdes_lowpass_f(float *output, int length, float cutoff_rate, int window) {
    (void)window; /* WINDOW_HAMMING, csdr's default, is the only window on pirip's path */
    const float pi_f = (float)3.14159265358979323846;
    auto hamming = [pi_f](float rate) -> float {
        rate = 0.5 + rate / 2;
        return 0.54 - 0.46 * cos((double)(2 * pi_f * rate)); /* C semantics: cos() is the double function */
    };
    const int middle = length / 2;
    output[middle] = 2 * pi_f * cutoff_rate * hamming(0);
    for (int i = 1; i <= middle; ++i)
        output[middle - i] = output[middle + i] = (sin((double)(2 * pi_f * cutoff_rate * i)) / i) * hamming((float)i / middle);
    float sum = 0;
    for (int i = 0; i < length; ++i) sum += output[i];
    for (int i = 0; i < length; ++i) output[i] = output[i] / sum;
}

extern "C" int fir_decimate_cc(complexf *input, complexf *output, int input_size, int decimation, float *taps, int taps_length) {
    pirip_decim_plan *p = pirip_decim_plan_create(decimation, taps, taps_length, pirip::current_device());
    if (!p) {
        fprintf(stderr, "libpirip_b200: fir_decimate_cc: %s\n", pirip::last_error());
        abort();
    }
    const size_t nout = out_count((size_t)input_size, decimation, taps_length);
    int rc = 0;
    float2 *d_in = nullptr, *d_out = nullptr;
    if (nout) {
        if (!pirip::cuda_ok(cudaMalloc(&d_in, sizeof(float2) * (size_t)input_size), "cudaMalloc") ||
            !pirip::cuda_ok(cudaMalloc(&d_out, sizeof(float2) * nout), "cudaMalloc") ||
            !pirip::cuda_ok(cudaMemcpy(d_in, input, sizeof(float2) * (size_t)input_size, cudaMemcpyHostToDevice), "H2D") ||
            !pirip::cuda_ok(decim_launch(p, false, d_in, 1, (size_t)input_size, sizeof(float2) * (size_t)input_size, (complexf *)d_out,
                                         nullptr, nout, nullptr),
                            "fir_decimate_kernel launch") ||
            !pirip::cuda_ok(cudaMemcpy(output, d_out, sizeof(float2) * nout, cudaMemcpyDeviceToHost), "D2H"))
            rc = -1;
    }
    cudaFree(d_in);
    cudaFree(d_out);
    pirip_decim_plan_destroy(p);
    if (rc) {
        fprintf(stderr, "libpirip_b200: fir_decimate_cc: %s\n", pirip::last_error());
        abort();
    }
    return (int)nout;
}
