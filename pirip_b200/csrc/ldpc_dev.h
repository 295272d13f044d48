/*
 * ldpc_dev.h - device side of the LDPC decoder, shared by the batch decoder (ldpc.cu) and the receive-chain kernel
 * (rx_chain.cu).  [U codec2 mpdecode_core.c::run_ldpc_decoder / SumProduct / phi0]; SURVEY.md 8a row A14.
 */
#ifndef PIRIP_LDPC_DEV_H
#define PIRIP_LDPC_DEV_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.h"

struct LdpcGraph {
    int N, NPB, E, max_iter;
    const uint16_t *c_off;  /* [NPB+1] first edge of every check */
    const uint16_t *v_off;  /* [N+1] first entry of every variable in v_edge */
    const uint16_t *c_vpos; /* [E] check order: position of the edge in variable order */
    const unsigned *v_rec;  /* [E] variable order: check << 8 | position of the edge in its check */
    /* shared-memory layout of the decoder kernel (bytes): tables once per CTA, then one block per warp (codeword) */
    int off_voff, off_vedge, off_echk, tables_bytes;
    int w_q, w_chk, w_llr, w_scr, warp_bytes;
    int warps;  /* codewords per CTA */
    float ms_scale; /* normalisation of the min-sum check update */
};

/* [U mpdecode_core.c::phi0] CML closed form with the coarse steps above 5 */
__device__ __forceinline__ float phi0(float x) {
    if (x > 10.0f) return 0.0f;
    if ((double)x < 9.08e-5) return 10.0f;
    if (x > 9.0f) return (float)1.6881e-4;
    if (x > 8.0f) return (float)4.5887e-4;
    if (x > 7.0f) return (float)1.2473e-3;
    if (x > 6.0f) return (float)3.3906e-3;
    if (x > 5.0f) return (float)9.2168e-3;
    const float z = (float)em_exp((double)x);
    return (float)em_log((double)((z + 1.0f) / (z - 1.0f)));
}
/* the same function in single precision with the hardware exponential / logarithm (tolerance mode) */
__device__ __forceinline__ float phi0_fast(float x) {
    if (x > 10.0f) return 0.0f;
    if (x < 9.08e-5f) return 10.0f;
    if (x > 9.0f) return 1.6881e-4f;
    if (x > 8.0f) return 4.5887e-4f;
    if (x > 7.0f) return 1.2473e-3f;
    if (x > 6.0f) return 3.3906e-3f;
    if (x > 5.0f) return 9.2168e-3f;
    const float z = __expf(x);
    return __logf(__fdividef(z + 1.0f, z - 1.0f));
}

enum { LDPC_SP_EXACT = 0, LDPC_SP_FAST = 1, LDPC_MINSUM = 2 };
constexpr int kLdpcMaxWarps = 8;

/* One WARP per codeword, up to 8 codewords per CTA, the Tanner graph tables loaded once per CTA.
   A lane owns whole check nodes in the check phase and whole variable nodes in the variable phase and walks their edges
   in table order, so in the exact mode every float sum is formed in the order of the scalar algorithm
   ([U mpdecode_core.c::SumProduct]) and bits / iterations / parity-check counts are identical to it.  Messages: one float per
   edge (the variable-to-check message: phi0(|q|) with the sign of q in the sign bit, or q itself for min-sum); the
   check-to-variable messages are never stored - a check leaves its sum (or its two smallest magnitudes) and sign parity, and
   the variable phase forms r from them.  The parity-check count is a warp reduction; the loop ends as soon as it reaches
   NumberParityBits (early exit per codeword, no CTA barrier inside the iterations). */
/* check-to-variable message from what the check left behind (cs: sum / smallest magnitude with the parity in the sign bit;
   ca: min-sum only, second smallest magnitude with the position of the smallest in its 6 low mantissa bits) */
template <int MODE>
__device__ __forceinline__ float r_msg(float qe, float cs, float ca, int pos) {
    const unsigned neg = (__float_as_uint(cs) ^ __float_as_uint(qe)) & 0x80000000u;
    float mag;
    if (MODE == LDPC_MINSUM) {
        const unsigned a2 = __float_as_uint(ca);
        mag = ((int)(a2 & 63u) == pos) ? __uint_as_float(a2 & ~63u) : fabsf(cs);
    } else {
        const float d = __fsub_rn(fabsf(cs), fabsf(qe));
        mag = MODE == LDPC_SP_EXACT ? phi0(d) : phi0_fast(d);
    }
    return __uint_as_float(__float_as_uint(mag) ^ neg);
}
/* variable-to-check message as stored: phi0(|t|) (or |t| for min-sum) with the sign the scalar code gives t (t > 0 ? 0 : 1) */
template <int MODE>
__device__ __forceinline__ float q_store(float t) {
    const float a = MODE == LDPC_MINSUM ? fabsf(t) : (MODE == LDPC_SP_EXACT ? phi0(fabsf(t)) : phi0_fast(fabsf(t)));
    return t > 0.0f ? a : -a;
}

/* Per-CTA view of the graph tables in shared memory and one warp's message storage */
struct LdpcSmem {
    const uint16_t *c_off, *v_off, *c_vpos;
    const unsigned *v_rec;
    float *q;      /* [E] variable order */
    float2 *chk;   /* [NPB] what a check leaves behind (see r_msg) */
    float *llr;    /* [N] */
    unsigned *scr; /* [256] scratch words (receive-chain kernel) */
};

/* all threads of the CTA: copy the graph tables to shared memory (followed by a CTA barrier by the caller) */
__device__ __forceinline__ void ldpc_load_tables(const LdpcGraph &g, unsigned char *smem_raw) {
    uint16_t *c_off = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *v_off = reinterpret_cast<uint16_t *>(smem_raw + g.off_voff);
    uint16_t *c_vpos = reinterpret_cast<uint16_t *>(smem_raw + g.off_vedge);
    unsigned *v_rec = reinterpret_cast<unsigned *>(smem_raw + g.off_echk);
    for (int i = threadIdx.x; i <= g.NPB; i += blockDim.x) c_off[i] = g.c_off[i];
    for (int i = threadIdx.x; i <= g.N; i += blockDim.x) v_off[i] = g.v_off[i];
    for (int i = threadIdx.x; i < g.E; i += blockDim.x) {
        c_vpos[i] = g.c_vpos[i];
        v_rec[i] = g.v_rec[i];
    }
}
__device__ __forceinline__ LdpcSmem ldpc_smem_view(const LdpcGraph &g, unsigned char *smem_raw, int warp) {
    LdpcSmem m;
    m.c_off = reinterpret_cast<const uint16_t *>(smem_raw);
    m.v_off = reinterpret_cast<const uint16_t *>(smem_raw + g.off_voff);
    m.c_vpos = reinterpret_cast<const uint16_t *>(smem_raw + g.off_vedge);
    m.v_rec = reinterpret_cast<const unsigned *>(smem_raw + g.off_echk);
    unsigned char *wb = smem_raw + g.tables_bytes + (size_t)warp * g.warp_bytes;
    m.q = reinterpret_cast<float *>(wb + g.w_q);
    m.chk = reinterpret_cast<float2 *>(wb + g.w_chk);
    m.llr = reinterpret_cast<float *>(wb + g.w_llr);
    m.scr = reinterpret_cast<unsigned *>(wb + g.w_scr);
    return m;
}

/* One codeword by one WARP.
   A lane owns whole check nodes in the check phase and whole variable nodes in the variable phase and walks their edges
   in table order, so in the exact mode every float sum is formed in the order of the scalar algorithm
   ([U mpdecode_core.c::SumProduct]) and bits / iterations / parity-check counts are identical to it.  Messages: one float per
   edge, stored in VARIABLE order (the variable phase reads and writes them sequentially; the check phase gathers them through
   an index table): the variable-to-check message phi0(|q|) with the sign of q in the sign bit, or q itself for min-sum.
   The check-to-variable messages are never stored - a check leaves its sum (or its two smallest magnitudes) and sign parity,
   and the variable phase forms r from them.  The parity-check count is a warp reduction; the loop ends as soon as it reaches
   NumberParityBits.  `load(v)` delivers the channel LLR of variable v; `store(v, bit)` receives the decisions of the last
   iteration (every lane calls it for its variables).  Returns the iteration count, *pcc_out the parity checks passed. */
template <int MODE, class Load, class Store>
__device__ __forceinline__ int ldpc_decode_one(const LdpcGraph &g, const LdpcSmem &m, int lane, Load load, Store store, int *pcc_out) {
    const uint16_t *c_off = m.c_off, *v_off = m.v_off, *c_vpos = m.c_vpos;
    const unsigned *v_rec = m.v_rec;
    float *q = m.q, *llr = m.llr;
    float2 *chk = m.chk;
    {
        for (int v = lane; v < g.N; v += 32) {
            const float x = load(v);
            llr[v] = x;
            float mm;
            if (MODE == LDPC_MINSUM)
                mm = x;
            else {
                const float ph = MODE == LDPC_SP_EXACT ? phi0(fabsf(x)) : phi0_fast(fabsf(x));
                mm = x < 0.0f ? -ph : ph;
            }
            for (int j = v_off[v]; j < v_off[v + 1]; ++j) q[j] = mm;
        }
        __syncwarp();
        int result = g.max_iter, pcc = 0;
        for (int iter = 0; iter < g.max_iter; ++iter) {
            /* check-node phase */
            int local = 0;
            for (int c = lane; c < g.NPB; c += 32) {
                const int e0 = c_off[c], e1 = c_off[c + 1];
                unsigned sign = 0;
                if (MODE == LDPC_MINSUM) {
                    float m1 = 3.0e38f, m2 = 3.0e38f;
                    int pos = 0;
                    for (int e = e0; e < e1; ++e) {
                        const float v = q[c_vpos[e]];
                        const float a = fabsf(v);
                        sign ^= __float_as_uint(v) & 0x80000000u;
                        if (a < m1) {
                            m2 = m1;
                            m1 = a;
                            pos = e - e0;
                        } else if (a < m2)
                            m2 = a;
                    }
                    chk[c] = make_float2(__uint_as_float(__float_as_uint(m1 * g.ms_scale) | sign),
                                         __uint_as_float((__float_as_uint(m2 * g.ms_scale) & ~63u) | (unsigned)(pos & 63)));
                } else {
                    const float v0 = q[c_vpos[e0]];
                    float phi_sum = fabsf(v0);
                    sign = __float_as_uint(v0) & 0x80000000u;
                    for (int e = e0 + 1; e < e1; ++e) {
                        const float v = q[c_vpos[e]];
                        phi_sum = __fadd_rn(phi_sum, fabsf(v));
                        sign ^= __float_as_uint(v) & 0x80000000u;
                    }
                    chk[c] = make_float2(__uint_as_float(__float_as_uint(phi_sum) | sign), 0.0f);
                }
                if (sign == 0) local++;
            }
            pcc = __reduce_add_sync(0xffffffffu, local);
            __syncwarp();
            const bool last = pcc == g.NPB || iter == g.max_iter - 1;
            /* variable-node phase */
            for (int v = lane; v < g.N; v += 32) {
                const int j0 = v_off[v], j1 = v_off[v + 1];
                float Qi = llr[v];
                float r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (j0 + k < j1) {
                        const unsigned rec = v_rec[j0 + k];
                        const float2 cs = chk[rec >> 8];
                        r[k] = r_msg<MODE>(q[j0 + k], cs.x, cs.y, (int)(rec & 63u));
                        Qi = __fadd_rn(Qi, r[k]);
                    }
                /* variables with more than 8 edges (none of codec2's codes): the rest of the sum */
                for (int j = j0 + 8; j < j1; ++j) {
                    const unsigned rec = v_rec[j];
                    const float2 cs = chk[rec >> 8];
                    Qi = __fadd_rn(Qi, r_msg<MODE>(q[j], cs.x, cs.y, (int)(rec & 63u)));
                }
                if (last) store(v, Qi < 0.0f ? 1 : 0);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (j0 + k < j1) q[j0 + k] = q_store<MODE>(__fsub_rn(Qi, r[k]));
                for (int j = j0 + 8; j < j1; ++j) {
                    const unsigned rec = v_rec[j];
                    const float2 cs = chk[rec >> 8];
                    q[j] = q_store<MODE>(__fsub_rn(Qi, r_msg<MODE>(q[j], cs.x, cs.y, (int)(rec & 63u))));
                }
            }
            __syncwarp();
            if (pcc == g.NPB) {
                result = iter + 1;
                break;
            }
        }
        __syncwarp();
        *pcc_out = pcc;
        return result;
    }
}

/* Batch decoder: one warp per codeword, up to 8 codewords per CTA, the Tanner graph tables loaded once per CTA; a CTA loops
   over codewords (early exit per codeword, no CTA barrier inside the iterations). */
template <int MODE>
__global__ void __launch_bounds__(32 * kLdpcMaxWarps) ldpc_decode_warp_kernel(const __grid_constant__ LdpcGraph g, const float *llr_in, size_t n_cw,
                                                                              unsigned char *bits_out, int *iter_out, int *pcc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ldpc_load_tables(g, smem_raw);
    __syncthreads();
    if (warp >= g.warps) return;
    const LdpcSmem m = ldpc_smem_view(g, smem_raw, warp);
    for (size_t cw = (size_t)blockIdx.x * g.warps + warp; cw < n_cw; cw += (size_t)gridDim.x * g.warps) {
        const float *in = llr_in + cw * g.N;
        unsigned char *out = bits_out + cw * g.N;
        int pcc = 0;
        const int it = ldpc_decode_one<MODE>(
            g, m, lane, [&](int v) { return in[v]; }, [&](int v, int bit) { out[v] = (unsigned char)bit; }, &pcc);
        if (lane == 0) {
            if (iter_out) iter_out[cw] = it;
            if (pcc_out) pcc_out[cw] = pcc;
        }
        __syncwarp();
    }
}

#endif
