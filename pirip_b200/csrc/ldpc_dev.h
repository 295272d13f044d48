/*
 * ldpc_dev.h - device side of the LDPC decoder, shared by the batch decoder (ldpc.cu) and the receive-chain kernel
 * (rx_chain.cu).  [U codec2 mpdecode_core.c::run_ldpc_decoder / SumProduct / phi0]; SURVEY.md 8a row A14.
 */
#ifndef PIRIP_LDPC_DEV_H
#define PIRIP_LDPC_DEV_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "exact_math.h"

struct LdpcGraph {
    int N, NPB, E, max_iter;
    const uint16_t *c_off;  /* [NPB+1] first edge of every check */
    const uint16_t *v_off;  /* [N+1] first entry of every variable in v_edge */
    const uint16_t *c_vpos; /* [E] check order: position of the edge in variable order */
    const uint16_t *c_idx8; /* [NPB][8] the same positions, one 16-byte record per check, 0xffff where a check has fewer edges */
    /* shared-memory layout of the decoder kernel (bytes): tables once per CTA, then one block per warp (codeword) */
    int off_voff, off_vedge, off_idx8, tables_bytes;
    int w_q, w_llr, w_scr, warp_bytes;
    int warps;  /* codewords per CTA */
    int max_vdeg; /* largest number of edges of a variable */
    int v_split;  /* variables from here on have at most two edges (the parity part of a repeat-accumulate code) */
    int max_cdeg; /* largest number of edges of a check */
    int v_deg_a;  /* > 0: every variable below v_split has exactly this many edges */
    int v_tail_regular, v_tail_base; /* variables from v_split on: two edges each (the last one may have one), first entry v_tail_base */
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
constexpr int kLdpcMaxWarps = 24;

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
    const uint4 *c_idx8;
    float *q;      /* [E] variable order: q on the way to the checks, r on the way back */
    float *llr;    /* [N] */
    unsigned *scr; /* [256] scratch words (receive-chain kernel) */
};

/* all threads of the CTA: copy the graph tables to shared memory (followed by a CTA barrier by the caller) */
__device__ __forceinline__ void ldpc_load_tables(const LdpcGraph &g, unsigned char *smem_raw) {
    uint16_t *c_off = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *v_off = reinterpret_cast<uint16_t *>(smem_raw + g.off_voff);
    uint16_t *c_vpos = reinterpret_cast<uint16_t *>(smem_raw + g.off_vedge);
    for (int i = threadIdx.x; i <= g.NPB; i += blockDim.x) c_off[i] = g.c_off[i];
    for (int i = threadIdx.x; i <= g.N; i += blockDim.x) v_off[i] = g.v_off[i];
    for (int i = threadIdx.x; i < g.E; i += blockDim.x) c_vpos[i] = g.c_vpos[i];
    uint16_t *idx8 = reinterpret_cast<uint16_t *>(smem_raw + g.off_idx8);
    for (int i = threadIdx.x; i < 8 * g.NPB; i += blockDim.x) idx8[i] = g.c_idx8[i];
}
__device__ __forceinline__ LdpcSmem ldpc_smem_view(const LdpcGraph &g, unsigned char *smem_raw, int warp) {
    LdpcSmem m;
    m.c_off = reinterpret_cast<const uint16_t *>(smem_raw);
    m.v_off = reinterpret_cast<const uint16_t *>(smem_raw + g.off_voff);
    m.c_vpos = reinterpret_cast<const uint16_t *>(smem_raw + g.off_vedge);
    m.c_idx8 = reinterpret_cast<const uint4 *>(smem_raw + g.off_idx8);
    unsigned char *wb = smem_raw + g.tables_bytes + (size_t)warp * g.warp_bytes;
    m.q = reinterpret_cast<float *>(wb + g.w_q);
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
/* The messages of a codeword live in ONE array of E floats in variable order, used in place by both phases:
     check phase     reads the variable-to-check messages q of its edges (gathered through c_vpos, up to 8 in flight before they are
                     combined, and the next check's behind them), forms sign parity and sum / two smallest magnitudes, and writes the
                     check-to-variable message r of every edge back to the edge's slot;
     variable phase  reads its r's sequentially, Q = llr + sum r in edge order, decision, writes q = Q - r back.
   A pad of +0 (sum-product: x + 0 = x exactly, sign unchanged) or of the largest float (min-sum) stands in for the edges a check
   does not have, so sums and minima are those of the scalar order ([U mpdecode_core.c::SumProduct]).
   The variable phase is unrolled VA wide for the variables below g.v_split and 2 wide from there on (the degree-2 parity
   variables of a repeat-accumulate code); edges beyond the unrolled width take a loop. */
template <int MODE, int VA, int CD, class Load, class Store>
__device__ __forceinline__ int ldpc_decode_impl(const LdpcGraph &g, const LdpcSmem &m, int lane, Load load, Store store, int *pcc_out) {
    const uint16_t *c_off = m.c_off, *v_off = m.v_off, *c_vpos = m.c_vpos;
    float *q = m.q, *llr = m.llr;
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
        if (g.v_tail_regular && v >= g.v_split) {
            /* the degree-2 chain lives in two planes (first edges, second edges): see the variable phase */
            const int t = v - g.v_split;
            q[g.v_tail_base + t] = mm;
            if (g.v_tail_base + 2 * t + 2 <= g.E) q[g.v_tail_base + (g.N - g.v_split) + t] = mm;
        } else
            for (int j = v_off[v]; j < v_off[v + 1]; ++j) q[j] = mm;
    }
    __syncwarp();
    int result = g.max_iter, pcc = 0;
    for (int iter = 0; iter < g.max_iter; ++iter) {
        /* check-node phase */
        int local = 0;
#pragma unroll(MODE == LDPC_SP_EXACT ? 1 : 2)
        for (int c = lane; c < g.NPB; c += 32) {
            const float pad = MODE == LDPC_MINSUM ? 3.0e38f : 0.0f;
            int idx[CD];
            float v[CD];
            /* the positions of the check's first eight edges in one 16-byte load */
            const uint4 rec = m.c_idx8[c];
            const unsigned w[4] = { rec.x, rec.y, rec.z, rec.w };
#pragma unroll
            for (int k = 0; k < CD; ++k) {
                const unsigned u = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
                idx[k] = u == 0xffffu ? -1 : (int)u;
            }
            int e0 = 0, e1 = 0;
            if (g.max_cdeg > CD) e0 = c_off[c], e1 = c_off[c + 1]; /* only checks with more edges than the record holds */
#pragma unroll
            for (int k = 0; k < CD; ++k) v[k] = idx[k] >= 0 ? q[idx[k]] : pad;
            unsigned sign = 0;
            if (MODE == LDPC_MINSUM) {
                /* smallest and second smallest magnitude without tracking a position: the edge that holds the smallest gets the
                   second smallest, every other edge the smallest - with a tie both are equal, so the result is the same */
                float m1 = 3.0e38f, m2 = 3.0e38f;
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    const float a = fabsf(v[k]);
                    sign ^= __float_as_uint(v[k]) & 0x80000000u;
                    m2 = fminf(m2, fmaxf(m1, a));
                    m1 = fminf(m1, a);
                }
                for (int e = e0 + CD; e < e1; ++e) {
                    const float w = q[c_vpos[e]];
                    const float a = fabsf(w);
                    sign ^= __float_as_uint(w) & 0x80000000u;
                    m2 = fminf(m2, fmaxf(m1, a));
                    m1 = fminf(m1, a);
                }
                const float s1 = m1 * g.ms_scale, s2 = m2 * g.ms_scale;
#pragma unroll
                for (int k = 0; k < CD; ++k)
                    if (idx[k] >= 0) {
                        const unsigned neg = (sign ^ __float_as_uint(v[k])) & 0x80000000u;
                        q[idx[k]] = __uint_as_float(__float_as_uint(fabsf(v[k]) <= m1 ? s2 : s1) ^ neg);
                    }
                for (int e = e0 + CD; e < e1; ++e) {
                    const int j = c_vpos[e];
                    const float w = q[j];
                    const unsigned neg = (sign ^ __float_as_uint(w)) & 0x80000000u;
                    q[j] = __uint_as_float(__float_as_uint(fabsf(w) <= m1 ? s2 : s1) ^ neg);
                }
            } else {
                float phi_sum = fabsf(v[0]);
                sign = __float_as_uint(v[0]) & 0x80000000u;
#pragma unroll
                for (int k = 1; k < CD; ++k) {
                    phi_sum = __fadd_rn(phi_sum, fabsf(v[k]));
                    sign ^= __float_as_uint(v[k]) & 0x80000000u;
                }
                for (int e = e0 + CD; e < e1; ++e) {
                    const float w = q[c_vpos[e]];
                    phi_sum = __fadd_rn(phi_sum, fabsf(w));
                    sign ^= __float_as_uint(w) & 0x80000000u;
                }
                const float cs = __uint_as_float(__float_as_uint(phi_sum) | sign);
#pragma unroll
                for (int k = 0; k < CD; ++k)
                    if (idx[k] >= 0) q[idx[k]] = r_msg<MODE>(v[k], cs, 0.0f, k);
                for (int e = e0 + CD; e < e1; ++e) {
                    const int j = c_vpos[e];
                    q[j] = r_msg<MODE>(q[j], cs, 0.0f, e - e0);
                }
            }
            if (sign == 0) local++;
        }
        pcc = __reduce_add_sync(0xffffffffu, local);
        __syncwarp();
        const bool last = pcc == g.NPB || iter == g.max_iter - 1;
        /* variable-node phase */
        auto vnode = [&](int v, auto W, int j0, int j1) {
            constexpr int VW = decltype(W)::value;
            float Qi = llr[v];
            float r[VW];
#pragma unroll
            for (int k = 0; k < VW; ++k)
                if (j0 + k < j1) {
                    r[k] = q[j0 + k];
                    Qi = __fadd_rn(Qi, r[k]);
                }
            for (int j = j0 + VW; j < j1; ++j) Qi = __fadd_rn(Qi, q[j]);
            if (last) store(v, Qi < 0.0f ? 1 : 0);
#pragma unroll
            for (int k = 0; k < VW; ++k)
                if (j0 + k < j1) q[j0 + k] = q_store<MODE>(__fsub_rn(Qi, r[k]));
            for (int j = j0 + VW; j < j1; ++j) q[j] = q_store<MODE>(__fsub_rn(Qi, q[j]));
        };
        if (g.v_deg_a == 4 && VA == 4) {
            /* every variable below v_split has four edges: one 16-byte load and store per variable (32-bit accesses 16 bytes apart
               would be 4-way bank conflicts), sums in edge order */
            for (int v = lane; v < g.v_split; v += 32) {
                float4 *q4 = reinterpret_cast<float4 *>(q) + v;
                const float4 r = *q4;
                const float Qi = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(llr[v], r.x), r.y), r.z), r.w);
                if (last) store(v, Qi < 0.0f ? 1 : 0);
                *q4 = make_float4(q_store<MODE>(__fsub_rn(Qi, r.x)), q_store<MODE>(__fsub_rn(Qi, r.y)), q_store<MODE>(__fsub_rn(Qi, r.z)),
                                  q_store<MODE>(__fsub_rn(Qi, r.w)));
            }
        } else if (g.v_deg_a > 0)
            /* every variable below v_split has v_deg_a edges: offsets by arithmetic, no table reads */
            for (int v = lane; v < g.v_split; v += 32) vnode(v, std::integral_constant<int, VA>(), v * g.v_deg_a, (v + 1) * g.v_deg_a);
        else
            for (int v = lane; v < g.v_split; v += 32) vnode(v, std::integral_constant<int, VA>(), v_off[v], v_off[v + 1]);
        if (g.v_tail_regular) {
            /* the degree-2 chain of a repeat-accumulate code: two edges each but the last variable.  Stored as two planes - the
               first edges of all chain variables, then the second edges - so that a check's two chain edges (the first edge of
               variable K + c, the second of K + c - 1) and a variable's two edges are consecutive words across the lanes:
               conflict-free in both phases (side by side they were 2-way conflicts in the check phase) */
            const int T2 = g.N - g.v_split;
            for (int v = g.v_split + lane; v < g.N; v += 32) {
                const int t = v - g.v_split;
                const bool two = g.v_tail_base + 2 * t + 2 <= g.E;
                float *qa = q + g.v_tail_base + t, *qb = q + g.v_tail_base + T2 + t;
                const float r0 = *qa, r1 = two ? *qb : 0.0f;
                float Qi = __fadd_rn(llr[v], r0);
                if (two) Qi = __fadd_rn(Qi, r1);
                if (last) store(v, Qi < 0.0f ? 1 : 0);
                *qa = q_store<MODE>(__fsub_rn(Qi, r0));
                if (two) *qb = q_store<MODE>(__fsub_rn(Qi, r1));
            }
        } else
            for (int v = g.v_split + lane; v < g.N; v += 32) vnode(v, std::integral_constant<int, 2>(), v_off[v], v_off[v + 1]);
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

template <int MODE, class Load, class Store>
__device__ __forceinline__ int ldpc_decode_one(const LdpcGraph &g, const LdpcSmem &m, int lane, Load load, Store store, int *pcc_out) {
    if (g.max_vdeg <= 4 && g.max_cdeg <= 6) return ldpc_decode_impl<MODE, 4, 6>(g, m, lane, load, store, pcc_out);
    if (g.max_vdeg <= 4) return ldpc_decode_impl<MODE, 4, 8>(g, m, lane, load, store, pcc_out);
    return ldpc_decode_impl<MODE, 8, 8>(g, m, lane, load, store, pcc_out);
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
