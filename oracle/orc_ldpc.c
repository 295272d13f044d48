/*
 * orc_ldpc.c — CPU oracle: FSK LLR mapping and LDPC sum-product decoder.
 * TEST INFRASTRUCTURE (see oracle.h).
 * PARITY UNPINNED: restatement of codec2 src/mpdecode_core.c (CML-derived), an
 * un-vendored dependency of pirip (build_codec2.sh:3-5); SURVEY.md 8a rows
 * A12, A14, A15, App. A.6.  On the path through `rtl_fsk --code H_256_512_4`
 * [R README.md:184] and the framer call at [R tx/rpitx_fsk.cpp:395].
 *
 * Known restatement choices (DESIGN.md "oracle confidence"):
 *  - phi0() is the CML closed form with the coarse steps above x=5; upstream's
 *    later fixed-point/LUT quantisation is not recoverable here.
 *  - log(I0(x)) is evaluated in the log domain so it cannot overflow.
 *  - edge order inside a node (which fixes the float summation order) is:
 *    check p -> data columns in H_rows order, then parity K+p-1, then K+p;
 *    variable v -> checks in H_cols order; parity variable K+p -> check p, p+1.
 */
#include "oracle.h"
#include "orc_math.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- LLRs */

/* log of the modified Bessel function I0 — Numerical Recipes polynomial (split at 3.75)
   [U mpdecode_core.c::bessi0], taken in the log domain */
static float logbessi0(float x) {
    float ax = fabsf(x);
    double y, ans;
    if (ax < 3.75) {
        y = x / 3.75;
        y *= y;
        ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
        return (float)om_log(ans);
    }
    y = 3.75 / ax;
    ans = 0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2)))))));
    return (float)(((double)ax - 0.5 * om_log((double)ax)) + om_log(ans));
}

#define AJIAN -0.24904163195436
#define TJIAN 2.50681740420944

/* [U mpdecode_core.c::max_star0] linear-log-MAP */
static float max_star0(float delta1, float delta2) {
    float diff = delta2 - delta1;
    if (diff > TJIAN)
        return delta2;
    else if (diff < -TJIAN)
        return delta1;
    else if (diff > 0)
        return delta2 + AJIAN * (diff - TJIAN);
    else
        return delta1 - AJIAN * (diff + TJIAN);
}

/* [U mpdecode_core.c::fsk_rx_filt_to_llrs = FskDemod + Somap]; llr > 0 means bit 0 */
void orc_fsk_rx_filt_to_llrs(float llr[], const float rx_filt[], float v_est, float SNRest, int M, int nsyms) {
    int bps = M == 2 ? 1 : 2;
    float *sym_lik = (float *)malloc(sizeof(float) * (size_t)(M * nsyms));
    float scale_factor = 2 * SNRest;
    for (int i = 0; i < nsyms; ++i)
        for (int j = 0; j < M; ++j) {
            float yr = rx_filt[j * nsyms + i];
            float y_envelope = sqrt(yr * yr / (v_est * v_est));
            sym_lik[i * M + j] = logbessi0(scale_factor * y_envelope);
        }
    for (int n = 0; n < nsyms; ++n) {
        float num[2], den[2];
        for (int k = 0; k < bps; ++k) {
            num[k] = -1000000;
            den[k] = -1000000;
        }
        for (int i = 0; i < M; ++i) {
            float metric = sym_lik[n * M + i];
            int mask = 1 << (bps - 1);
            for (int k = 0; k < bps; ++k) {
                if (mask & i)
                    num[k] = max_star0(num[k], metric);
                else
                    den[k] = max_star0(den[k], metric);
                mask = mask >> 1;
            }
        }
        for (int k = 0; k < bps; ++k) llr[bps * n + k] = -(num[k] - den[k]);
    }
    free(sym_lik);
}

/* ---------------------------------------------------------------- LDPC */

/* [U mpdecode_core.c::phi0] */
float orc_phi0(float x) {
    float z;
    if (x > 10)
        return 0;
    else if (x < 9.08e-5)
        return 10;
    else if (x > 9)
        return 1.6881e-4;
    else if (x > 8)
        return 4.5887e-4;
    else if (x > 7)
        return 1.2473e-3;
    else if (x > 6)
        return 3.3906e-3;
    else if (x > 5)
        return 9.2168e-3;
    else {
        z = (float)om_exp(x);
        return (float)om_log((z + 1) / (z - 1));
    }
}

/* [U mpdecode_core.c::encode] */
void orc_ldpc_encode(const struct orc_ldpc *ldpc, const uint8_t ibits[], uint8_t pbits[]) {
    unsigned prev = 0;
    for (int p = 0; p < ldpc->NumberParityBits; ++p) {
        unsigned par = 0;
        for (int i = 0; i < ldpc->max_row_weight; ++i) {
            int ind = ldpc->H_rows[p + i * ldpc->NumberParityBits];
            if (ind) par = par + ibits[ind - 1];
        }
        unsigned tmp = (par + prev) & 1;
        prev = tmp;
        pbits[p] = (uint8_t)tmp;
    }
}

struct c_sub { uint16_t index, socket; float message; };
struct v_sub { uint16_t index, socket; float message; uint8_t sign; };
struct c_node { int degree; struct c_sub *subs; };
struct v_node { int degree; float initial_value; struct v_sub *subs; };

/* [U mpdecode_core.c::init_c_v_nodes] Tanner graph from the H_rows / H_cols tables */
static void build_graph(const struct orc_ldpc *l, const float input[], struct c_node *c_nodes, struct v_node *v_nodes) {
    const int N = l->CodeLength, NPB = l->NumberParityBits, K = N - NPB;
    const int H1 = (l->NumberRowsHcols != N);

    for (int p = 0; p < NPB; ++p) {
        int count = 0;
        for (int j = 0; j < l->max_row_weight; ++j)
            if (l->H_rows[p + j * NPB] > 0) count++;
        int deg = count;
        if (H1) deg += (p == 0) ? 1 : 2;
        c_nodes[p].degree = deg;
        c_nodes[p].subs = (struct c_sub *)calloc((size_t)deg, sizeof(struct c_sub));
        int e = 0;
        for (int j = 0; j < l->max_row_weight; ++j) {
            int ind = l->H_rows[p + j * NPB];
            if (ind > 0) c_nodes[p].subs[e++].index = (uint16_t)(ind - 1);
        }
        if (H1) {
            if (p > 0) c_nodes[p].subs[e++].index = (uint16_t)(K + p - 1);
            c_nodes[p].subs[e++].index = (uint16_t)(K + p);
        }
        assert(e == deg);
    }
    for (int v = 0; v < N; ++v) {
        int deg = 0;
        if (H1 && v >= K) {
            deg = (v != N - 1) ? 2 : 1;
        } else {
            for (int j = 0; j < l->max_col_weight; ++j)
                if (l->H_cols[v + j * l->NumberRowsHcols] > 0) deg++;
        }
        v_nodes[v].degree = deg;
        v_nodes[v].initial_value = input[v];
        v_nodes[v].subs = (struct v_sub *)calloc((size_t)deg, sizeof(struct v_sub));
        int e = 0;
        if (H1 && v >= K) {
            v_nodes[v].subs[e++].index = (uint16_t)(v - K);
            if (v != N - 1) v_nodes[v].subs[e++].index = (uint16_t)(v - K + 1);
        } else {
            for (int j = 0; j < l->max_col_weight; ++j) {
                int ind = l->H_cols[v + j * l->NumberRowsHcols];
                if (ind > 0) v_nodes[v].subs[e++].index = (uint16_t)(ind - 1);
            }
        }
        assert(e == deg);
        for (int j = 0; j < deg; ++j) {
            const struct c_node *c = &c_nodes[v_nodes[v].subs[j].index];
            int found = 0;
            for (int ci = 0; ci < c->degree; ++ci)
                if (c->subs[ci].index == v) {
                    v_nodes[v].subs[j].socket = (uint16_t)ci;
                    found = 1;
                    break;
                }
            assert(found && "H_rows and H_cols disagree");
            v_nodes[v].subs[j].message = orc_phi0(fabsf(input[v]));
            v_nodes[v].subs[j].sign = input[v] < 0 ? 1 : 0;
        }
    }
    for (int p = 0; p < NPB; ++p)
        for (int j = 0; j < c_nodes[p].degree; ++j) {
            const struct v_node *v = &v_nodes[c_nodes[p].subs[j].index];
            int found = 0;
            for (int vi = 0; vi < v->degree; ++vi)
                if (v->subs[vi].index == p) {
                    c_nodes[p].subs[j].socket = (uint16_t)vi;
                    found = 1;
                    break;
                }
            assert(found && "H_rows and H_cols disagree");
        }
}

/* [U mpdecode_core.c::SumProduct] flooding schedule, early exit on all checks satisfied */
static int sum_product(int *parityCheckCount, uint8_t DecodedBits[], struct c_node c_nodes[], struct v_node v_nodes[],
                       int CodeLength, int NumberParityBits, int max_iter) {
    int result = max_iter;
    for (int iter = 0; iter < max_iter; ++iter) {
        /* check-node update */
        int ssum = 0;
        for (int j = 0; j < NumberParityBits; ++j) {
            struct c_sub *c0 = &c_nodes[j].subs[0];
            int sign = v_nodes[c0->index].subs[c0->socket].sign;
            float phi_sum = v_nodes[c0->index].subs[c0->socket].message;
            for (int i = 1; i < c_nodes[j].degree; ++i) {
                struct c_sub *cp = &c_nodes[j].subs[i];
                struct v_sub *vp = &v_nodes[cp->index].subs[cp->socket];
                phi_sum += vp->message;
                sign ^= vp->sign;
            }
            if (sign == 0) ssum++;
            for (int i = 0; i < c_nodes[j].degree; ++i) {
                struct c_sub *cp = &c_nodes[j].subs[i];
                struct v_sub *vp = &v_nodes[cp->index].subs[cp->socket];
                if (sign ^ vp->sign)
                    cp->message = -orc_phi0(phi_sum - vp->message);
                else
                    cp->message = orc_phi0(phi_sum - vp->message);
            }
        }
        /* variable-node update */
        for (int i = 0; i < CodeLength; ++i) {
            float Qi = v_nodes[i].initial_value;
            for (int j = 0; j < v_nodes[i].degree; ++j) {
                struct v_sub *vp = &v_nodes[i].subs[j];
                Qi += c_nodes[vp->index].subs[vp->socket].message;
            }
            DecodedBits[i] = Qi < 0 ? 1 : 0;
            for (int j = 0; j < v_nodes[i].degree; ++j) {
                struct v_sub *vp = &v_nodes[i].subs[j];
                float temp_sum = Qi - c_nodes[vp->index].subs[vp->socket].message;
                vp->message = orc_phi0(fabsf(temp_sum));
                vp->sign = temp_sum > 0 ? 0 : 1;
            }
        }
        *parityCheckCount = ssum;
        if (ssum == NumberParityBits) {
            result = iter + 1;
            break;
        }
    }
    return result;
}

/* [U mpdecode_core.c::run_ldpc_decoder] */
int orc_run_ldpc_decoder(const struct orc_ldpc *ldpc, uint8_t out_char[], const float input[], int *parityCheckCount) {
    const int N = ldpc->CodeLength, NPB = ldpc->NumberParityBits;
    struct c_node *c_nodes = (struct c_node *)calloc((size_t)NPB, sizeof(struct c_node));
    struct v_node *v_nodes = (struct v_node *)calloc((size_t)N, sizeof(struct v_node));
    build_graph(ldpc, input, c_nodes, v_nodes);
    for (int i = 0; i < N; ++i) out_char[i] = 0;
    *parityCheckCount = 0;
    int iter = sum_product(parityCheckCount, out_char, c_nodes, v_nodes, N, NPB, ldpc->max_iter);
    for (int i = 0; i < NPB; ++i) free(c_nodes[i].subs);
    for (int i = 0; i < N; ++i) free(v_nodes[i].subs);
    free(c_nodes);
    free(v_nodes);
    return iter;
}
