/*
 * ldpc.cu — LDPC sum-product decoder and soft-decision map on sm_100a, plus the host side of
 * FSK_LDPC framing (code tables, RA encoder, CRC16, pack/unpack, UW framer/deframer).
 * Contract: include/pirip_ldpc.h.
 *
 * Decoder ([U codec2 mpdecode_core.c::run_ldpc_decoder / init_c_v_nodes / SumProduct / phi0],
 * on pirip's path through `rtl_fsk --code H_256_512_4` [R README.md:184]):
 * flooding schedule, check update r = +-phi0(sum phi - phi_self) with the sign parity, variable
 * update Q = llr + sum r, extrinsic q = Q - r, stop as soon as every check is satisfied.
 *
 * Mapping: one CTA per codeword; messages live in shared memory indexed by edge (edges numbered
 * check by check).  A thread owns whole check nodes in the check phase and whole variable nodes
 * in the variable phase, and walks their edges in table order, so every float sum is formed in
 * the same order as the scalar algorithm and the decoded bits, iteration count and parity-check
 * count are bit-identical to it.  phi0 uses exact_math.h (double exp/log, IEEE order).
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pirip_ldpc.h"
#include "exact_math.h"
#include "host_util.h"
#include "llr_dev.h"
#include "ldpc_dev.h"

namespace {

/* ------------------------------------------------------------------ device side */

__global__ void llr_kernel(const float *rx_filt, float *llr, float v_est, float SNRest, int M, int nsyms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsyms) return;
    float l2[2];
    symbol_llrs(rx_filt + i, nsyms, M, v_est, SNRest, l2);
    const int bps = M == 2 ? 1 : 2;
    for (int k = 0; k < bps; ++k) llr[bps * i + k] = l2[k];
}

/* ------------------------------------------------------------------ code tables */

struct builtin_code {
    const char *name;
    int max_iter, CodeLength, NumberParityBits, NumberRowsHcols, max_row_weight, max_col_weight;
    const uint16_t *H_rows, *H_cols;
};
#include "ldpc_tables.inc"

struct CodeEntry {
    std::string name;
    int max_iter, CodeLength, NumberParityBits, NumberRowsHcols, max_row_weight, max_col_weight;
    std::vector<uint16_t> H_rows, H_cols;
    bool standin = false; /* a generated table of the same shape, NOT codec2's matrix */
    bool warned = false;
};

/* codec2's generated table files (src/H_256_512_4.h + .c and friends) as text: `#define <NAME>_<FIELD> n` and
   `<NAME>_H_rows[] = { ... }` / `<NAME>_H_cols[] = { ... }`.  Every name with both arrays and its dimensions becomes an entry. */
static void parse_codec2_tables(const std::string &txt, std::vector<CodeEntry> &out) {
    auto is_id = [](char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_'; };
    auto define_of = [&](const std::string &name, const char *field, int dflt) {
        const std::string key = name + "_" + field;
        size_t pos = 0;
        while ((pos = txt.find(key, pos)) != std::string::npos) {
            const size_t end = pos + key.size();
            if ((pos == 0 || !is_id(txt[pos - 1])) && end < txt.size() && !is_id(txt[end])) return atoi(txt.c_str() + end);
            pos = end;
        }
        return dflt;
    };
    auto array_of = [&](const std::string &name, const char *suffix, std::vector<uint16_t> &v) {
        const std::string key = name + suffix;
        size_t pos = 0;
        while ((pos = txt.find(key, pos)) != std::string::npos) {
            size_t q = pos + key.size();
            pos = q;
            while (q < txt.size() && (txt[q] == ' ' || txt[q] == '[' || txt[q] == ']')) ++q;
            if (q >= txt.size() || txt[q] != '=') continue; /* a declaration (extern ...;), not the definition */
            const size_t open = txt.find('{', q), close = txt.find('}', q);
            if (open == std::string::npos || close == std::string::npos || close < open) return false;
            v.clear();
            for (size_t i = open + 1; i < close;) {
                if (txt[i] >= '0' && txt[i] <= '9') {
                    v.push_back((uint16_t)strtoul(txt.c_str() + i, nullptr, 10));
                    while (i < close && txt[i] >= '0' && txt[i] <= '9') ++i;
                } else
                    ++i;
            }
            return !v.empty();
        }
        return false;
    };
    size_t pos = 0;
    const std::string tag = "_H_rows";
    while ((pos = txt.find(tag, pos)) != std::string::npos) {
        size_t b = pos;
        while (b > 0 && is_id(txt[b - 1])) --b;
        const std::string name = txt.substr(b, pos - b);
        pos += tag.size();
        bool have = false;
        for (const CodeEntry &e : out) have = have || e.name == name;
        if (name.empty() || have) continue;
        CodeEntry e;
        e.name = name;
        if (!array_of(name, "_H_rows", e.H_rows) || !array_of(name, "_H_cols", e.H_cols)) continue;
        e.CodeLength = define_of(name, "CODELENGTH", 0);
        e.NumberParityBits = define_of(name, "NUMBERPARITYBITS", 0);
        e.NumberRowsHcols = define_of(name, "NUMBERROWSHCOLS", 0);
        e.max_row_weight = define_of(name, "MAX_ROW_WEIGHT", 0);
        e.max_col_weight = define_of(name, "MAX_COL_WEIGHT", 0);
        e.max_iter = define_of(name, "MAX_ITER", 15);
        if (e.CodeLength <= 0 || e.NumberParityBits <= 0 || e.NumberRowsHcols <= 0 || e.max_row_weight <= 0 || e.max_col_weight <= 0 ||
            e.H_rows.size() != (size_t)e.NumberParityBits * e.max_row_weight || e.H_cols.size() != (size_t)e.NumberRowsHcols * e.max_col_weight) {
            fprintf(stderr, "pirip_b200: PIRIP_LDPC_TABLES: table %s is incomplete (dimensions / array sizes), ignored\n", name.c_str());
            continue;
        }
        out.push_back(std::move(e));
    }
}

std::mutex g_codes_mu;
std::vector<CodeEntry> &codes() {
    static std::vector<CodeEntry> v = [] {
        std::vector<CodeEntry> w;
        for (const builtin_code &b : builtin_codes) {
            CodeEntry e;
            e.name = b.name;
            e.max_iter = b.max_iter;
            e.CodeLength = b.CodeLength;
            e.NumberParityBits = b.NumberParityBits;
            e.NumberRowsHcols = b.NumberRowsHcols;
            e.max_row_weight = b.max_row_weight;
            e.max_col_weight = b.max_col_weight;
            e.H_rows.assign(b.H_rows, b.H_rows + (size_t)b.NumberParityBits * b.max_row_weight);
            e.H_cols.assign(b.H_cols, b.H_cols + (size_t)b.NumberRowsHcols * b.max_col_weight);
            e.standin = true;
            w.push_back(std::move(e));
        }
        /* PIRIP_LDPC_TABLES=<codec2>/src/H_256_512_4.h:<codec2>/src/H_256_512_4.c[:...]: codec2's own tables take the genuine
           names; the stand-in of the same name stays reachable as <name>_STANDIN */
        if (const char *ev = getenv("PIRIP_LDPC_TABLES")) {
            std::string txt, list(ev);
            size_t a = 0;
            while (a <= list.size()) {
                size_t b = list.find(':', a);
                if (b == std::string::npos) b = list.size();
                const std::string path = list.substr(a, b - a);
                if (!path.empty()) {
                    FILE *fp = fopen(path.c_str(), "rb");
                    if (!fp)
                        fprintf(stderr, "pirip_b200: PIRIP_LDPC_TABLES: cannot read %s\n", path.c_str());
                    else {
                        char buf[1 << 14];
                        size_t n;
                        while ((n = fread(buf, 1, sizeof buf, fp)) > 0) txt.append(buf, n);
                        fclose(fp);
                        txt.push_back('\n');
                    }
                }
                a = b + 1;
            }
            std::vector<CodeEntry> real;
            parse_codec2_tables(txt, real);
            for (CodeEntry &r : real) {
                for (CodeEntry &e : w)
                    if (e.name == r.name && e.standin) e.name += "_STANDIN";
                w.push_back(std::move(r));
            }
        }
        w.reserve(64); /* keep table pointers handed out by ldpc_codes_setup stable */
        return w;
    }();
    return v;
}

/* Tanner graph in edge-list form; node-local edge order = table order (see file header) */
bool build_graph(const struct LDPC *l, std::vector<uint16_t> &c_off, std::vector<uint16_t> &v_off, std::vector<uint16_t> &v_edge,
                 std::string &why) {
    const int N = l->CodeLength, NPB = l->NumberParityBits, K = N - NPB;
    const bool H1 = l->NumberRowsHcols != N;
    if (N <= 0 || NPB <= 0 || K <= 0 || N > 60000) {
        why = "bad code dimensions";
        return false;
    }
    std::vector<std::vector<int>> cvars((size_t)NPB);
    for (int p = 0; p < NPB; ++p) {
        for (int j = 0; j < l->max_row_weight; ++j) {
            const int ind = l->H_rows[p + j * NPB];
            if (ind > 0) cvars[(size_t)p].push_back(ind - 1);
        }
        if (H1) {
            if (p > 0) cvars[(size_t)p].push_back(K + p - 1);
            cvars[(size_t)p].push_back(K + p);
        }
    }
    c_off.assign((size_t)NPB + 1, 0);
    size_t E = 0;
    for (int p = 0; p < NPB; ++p) {
        c_off[(size_t)p] = (uint16_t)E;
        E += cvars[(size_t)p].size();
    }
    if (E >= 65535) { /* 0xffff marks a missing edge in the per-check index records */
        why = "too many edges";
        return false;
    }
    c_off[(size_t)NPB] = (uint16_t)E;
    v_off.assign((size_t)N + 1, 0);
    v_edge.clear();
    for (int v = 0; v < N; ++v) {
        v_off[(size_t)v] = (uint16_t)v_edge.size();
        std::vector<int> checks;
        if (H1 && v >= K) {
            checks.push_back(v - K);
            if (v != N - 1) checks.push_back(v - K + 1);
        } else {
            for (int j = 0; j < l->max_col_weight; ++j) {
                const int ind = l->H_cols[v + j * l->NumberRowsHcols];
                if (ind > 0) checks.push_back(ind - 1);
            }
        }
        for (int c : checks) {
            if (c < 0 || c >= NPB) {
                why = "H_cols refers to a check that does not exist";
                return false;
            }
            int pos = -1;
            for (size_t i = 0; i < cvars[(size_t)c].size(); ++i)
                if (cvars[(size_t)c][i] == v) {
                    pos = (int)i;
                    break;
                }
            if (pos < 0) {
                why = "H_rows and H_cols disagree";
                return false;
            }
            v_edge.push_back((uint16_t)(c_off[(size_t)c] + pos));
        }
    }
    v_off[(size_t)N] = (uint16_t)v_edge.size();
    if (v_edge.size() != E) {
        why = "H_rows and H_cols disagree (edge counts)";
        return false;
    }
    return true;
}

}  // namespace

/* ------------------------------------------------------------------ batched decoder plan */

struct pirip_ldpc_plan {
    LdpcGraph g;
    int device, smem;
    int mode = LDPC_SP_EXACT;
    int num_sms = 148;
    uint16_t *d_c_off = nullptr, *d_v_off = nullptr, *d_c_vpos = nullptr, *d_c_idx8 = nullptr;
    float *d_llr = nullptr;
    unsigned char *d_bits = nullptr;
    int *d_iter = nullptr, *d_pcc = nullptr;
    size_t cap = 0;
};

extern "C" pirip_ldpc_plan *pirip_ldpc_plan_create(const struct LDPC *l, int device) {
    if (!l || !l->H_rows || !l->H_cols) {
        pirip::set_error("LDPC tables missing");
        return nullptr;
    }
    std::vector<uint16_t> c_off, v_off, v_edge;
    std::string why;
    if (!build_graph(l, c_off, v_off, v_edge, why)) {
        pirip::set_error("LDPC code rejected: " + why);
        return nullptr;
    }
    pirip_ldpc_plan *p = new pirip_ldpc_plan();
    p->device = device;
    p->g.N = l->CodeLength;
    p->g.NPB = l->NumberParityBits;
    p->g.E = (int)v_edge.size();
    p->g.max_iter = l->max_iter > 0 ? l->max_iter : 15;
    /* v_edge[j] = edge id (check order) of the j-th entry in variable order */
    std::vector<uint16_t> c_vpos((size_t)p->g.E);
    {
        for (int j = 0; j < p->g.E; ++j) c_vpos[(size_t)v_edge[(size_t)j]] = (uint16_t)j;
    }
    {
        auto al = [](int v) { return (v + 15) / 16 * 16; };
        LdpcGraph &g = p->g;
        int o = al((g.NPB + 1) * 2);
        g.off_voff = o;
        o += al((g.N + 1) * 2);
        g.off_vedge = o;
        o += al(g.E * 2);
        g.off_idx8 = o;
        o += al(g.NPB * 16);
        g.tables_bytes = o;
        int w = 0;
        g.w_q = w;
        w += al(g.E * 4);
        g.w_llr = w;
        w += al(g.N * 4);
        g.w_scr = w; /* 256 words of scratch for the receive-chain kernel: hard-bit words of two frames, decoded bits */
        w += 1024;
        g.warp_bytes = w;
        int warps = (227 * 1024 - g.tables_bytes) / w;
        /* one CTA per SM with as many codewords (warps) as shared memory holds: the tables are loaded once for all of them */
        if (warps > kLdpcMaxWarps) warps = kLdpcMaxWarps;
        g.warps = warps;
        g.max_vdeg = 0;
        for (int v = 0; v < g.N; ++v) g.max_vdeg = std::max(g.max_vdeg, (int)v_off[(size_t)v + 1] - (int)v_off[(size_t)v]);
        g.v_split = g.N;
        while (g.v_split > 0 && (int)v_off[(size_t)g.v_split] - (int)v_off[(size_t)g.v_split - 1] <= 2) --g.v_split;
        g.v_split = (g.v_split + 31) / 32 * 32; /* whole rounds of lanes on either side */
        if (g.v_split > g.N) g.v_split = g.N;
        g.max_cdeg = 0;
        for (int c = 0; c < g.NPB; ++c) g.max_cdeg = std::max(g.max_cdeg, (int)c_off[(size_t)c + 1] - (int)c_off[(size_t)c]);
        g.v_deg_a = g.v_split > 0 ? (int)v_off[1] - (int)v_off[0] : 0;
        for (int v = 0; v < g.v_split; ++v)
            if ((int)v_off[(size_t)v + 1] - (int)v_off[(size_t)v] != g.v_deg_a) g.v_deg_a = 0;
        g.v_tail_base = g.v_split < g.N ? (int)v_off[(size_t)g.v_split] : 0;
        g.v_tail_regular = g.v_split < g.N ? 1 : 0;
        for (int v = g.v_split; v < g.N; ++v) {
            const int d = (int)v_off[(size_t)v + 1] - (int)v_off[(size_t)v];
            if (!(d == 2 || (v == g.N - 1 && d >= 1 && d <= 2)) || (int)v_off[(size_t)v] != g.v_tail_base + 2 * (v - g.v_split)) g.v_tail_regular = 0;
        }
        if (getenv("PIRIP_LDPC_NO_REGULAR")) g.v_deg_a = 0, g.v_tail_regular = 0;
        g.ms_scale = 0.75f;
        p->smem = g.tables_bytes + (warps > 0 ? warps : 0) * w;
    }
    if (p->g.warps < 1) {
        pirip::set_error("code too large for one CTA's shared memory");
        delete p;
        return nullptr;
    }
    if (p->g.v_tail_regular) {
        /* the degree-2 chain is stored in two planes (ldpc_dev.h, variable phase): variable-order index -> storage position */
        const int tb = p->g.v_tail_base, T2 = p->g.N - p->g.v_split;
        for (uint16_t &j : c_vpos)
            if ((int)j >= tb) {
                const int t = ((int)j - tb) >> 1, sidx = ((int)j - tb) & 1;
                j = (uint16_t)(tb + sidx * T2 + t);
            }
    }
    for (int c = 0; c < p->g.NPB; ++c)
        if (c_off[(size_t)c + 1] - c_off[(size_t)c] > 64) {
            pirip::set_error("check nodes with more than 64 edges are not supported");
            delete p;
            return nullptr;
        }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) p->num_sms = sms;
    }
    if (const char *e = getenv("PIRIP_LDPC_MODE")) {
        if (!strcmp(e, "minsum")) p->mode = LDPC_MINSUM;
        else if (!strcmp(e, "fast")) p->mode = LDPC_SP_FAST;
    }
    std::vector<uint16_t> c_idx8((size_t)p->g.NPB * 8, (uint16_t)0xffff);
    for (int c = 0; c < p->g.NPB; ++c)
        for (int e = c_off[(size_t)c], k = 0; e < c_off[(size_t)c + 1] && k < 8; ++e, ++k) c_idx8[(size_t)c * 8 + (size_t)k] = c_vpos[(size_t)e];
    if (!pirip::cuda_ok(cudaSetDevice(device), "cudaSetDevice") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_c_idx8, 2 * c_idx8.size()), "cudaMalloc") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_c_idx8, c_idx8.data(), 2 * c_idx8.size(), cudaMemcpyHostToDevice), "copy") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_c_off, 2 * c_off.size()), "cudaMalloc") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_v_off, 2 * v_off.size()), "cudaMalloc") ||
        !pirip::cuda_ok(cudaMalloc(&p->d_c_vpos, 2 * c_vpos.size()), "cudaMalloc") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_c_off, c_off.data(), 2 * c_off.size(), cudaMemcpyHostToDevice), "copy") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_v_off, v_off.data(), 2 * v_off.size(), cudaMemcpyHostToDevice), "copy") ||
        !pirip::cuda_ok(cudaMemcpy(p->d_c_vpos, c_vpos.data(), 2 * c_vpos.size(), cudaMemcpyHostToDevice), "copy") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(ldpc_decode_warp_kernel<LDPC_SP_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem), "smem attr") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(ldpc_decode_warp_kernel<LDPC_SP_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem), "smem attr") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(ldpc_decode_warp_kernel<LDPC_MINSUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem), "smem attr")) {
        pirip_ldpc_plan_destroy(p);
        return nullptr;
    }
    p->g.c_off = p->d_c_off;
    p->g.v_off = p->d_v_off;
    p->g.c_vpos = p->d_c_vpos;
    p->g.c_idx8 = p->d_c_idx8;
    return p;
}

extern "C" int pirip_ldpc_plan_set_mode(pirip_ldpc_plan *p, int mode) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (mode != LDPC_SP_EXACT && mode != LDPC_SP_FAST && mode != LDPC_MINSUM) return pirip::fail(-2, "unknown decoder mode");
    p->mode = mode;
    return 0;
}

extern "C" int pirip_ldpc_plan_mode(const pirip_ldpc_plan *p) { return p ? p->mode : -1; }

extern "C" void pirip_ldpc_plan_destroy(pirip_ldpc_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_c_off);
    cudaFree(p->d_v_off);
    cudaFree(p->d_c_vpos);
    cudaFree(p->d_c_idx8);
    cudaFree(p->d_llr);
    cudaFree(p->d_bits);
    cudaFree(p->d_iter);
    cudaFree(p->d_pcc);
    delete p;
}

extern "C" int pirip_ldpc_decode_batch_dev(pirip_ldpc_plan *p, const float *d_llr, size_t n, uint8_t *d_bits, int32_t *d_iter,
                                           int32_t *d_pcc, void *stream) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (n == 0) return 0;
    if (!d_llr || !d_bits) return pirip::fail(-2, "NULL buffer");
    if (!pirip::cuda_ok(cudaSetDevice(p->device), "cudaSetDevice")) return -10;
    const size_t ctas_needed = (n + p->g.warps - 1) / p->g.warps;
    const size_t resident = (size_t)p->num_sms * ((p->smem + 1024) * 2 <= 228 * 1024 ? 2 : 1);
    const unsigned grid = (unsigned)(ctas_needed < resident ? ctas_needed : resident); /* persistent CTAs loop over the codewords */
    const unsigned block = 32u * (unsigned)p->g.warps;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->mode == LDPC_MINSUM)
        ldpc_decode_warp_kernel<LDPC_MINSUM><<<grid, block, p->smem, st>>>(p->g, d_llr, n, d_bits, d_iter, d_pcc);
    else if (p->mode == LDPC_SP_FAST)
        ldpc_decode_warp_kernel<LDPC_SP_FAST><<<grid, block, p->smem, st>>>(p->g, d_llr, n, d_bits, d_iter, d_pcc);
    else
        ldpc_decode_warp_kernel<LDPC_SP_EXACT><<<grid, block, p->smem, st>>>(p->g, d_llr, n, d_bits, d_iter, d_pcc);
    pirip::count_launch();
    return pirip::cuda_ok(cudaGetLastError(), "ldpc_decode_warp_kernel launch") ? 0 : -11;
}

extern "C" int pirip_ldpc_decode_batch(pirip_ldpc_plan *p, const float *llr, size_t n, uint8_t *bits, int32_t *iter, int32_t *pcc) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (n == 0) return 0;
    if (!pirip::cuda_ok(cudaSetDevice(p->device), "cudaSetDevice")) return -10;
    if (n > p->cap) {
        cudaFree(p->d_llr);
        cudaFree(p->d_bits);
        cudaFree(p->d_iter);
        cudaFree(p->d_pcc);
        p->d_llr = nullptr;
        p->d_bits = nullptr;
        p->d_iter = p->d_pcc = nullptr;
        p->cap = 0;
        if (!pirip::cuda_ok(cudaMalloc(&p->d_llr, sizeof(float) * n * p->g.N), "cudaMalloc") ||
            !pirip::cuda_ok(cudaMalloc(&p->d_bits, n * p->g.N), "cudaMalloc") ||
            !pirip::cuda_ok(cudaMalloc(&p->d_iter, sizeof(int) * n), "cudaMalloc") ||
            !pirip::cuda_ok(cudaMalloc(&p->d_pcc, sizeof(int) * n), "cudaMalloc"))
            return -12;
        p->cap = n;
    }
    if (!pirip::cuda_ok(cudaMemcpy(p->d_llr, llr, sizeof(float) * n * p->g.N, cudaMemcpyHostToDevice), "H2D")) return -12;
    const int rc = pirip_ldpc_decode_batch_dev(p, p->d_llr, n, p->d_bits, p->d_iter, p->d_pcc, nullptr);
    if (rc) return rc;
    if (!pirip::cuda_ok(cudaMemcpy(bits, p->d_bits, n * p->g.N, cudaMemcpyDeviceToHost), "ldpc_decode_warp_kernel")) return -13;
    if (iter && !pirip::cuda_ok(cudaMemcpy(iter, p->d_iter, sizeof(int) * n, cudaMemcpyDeviceToHost), "D2H")) return -13;
    if (pcc && !pirip::cuda_ok(cudaMemcpy(pcc, p->d_pcc, sizeof(int) * n, cudaMemcpyDeviceToHost), "D2H")) return -13;
    return 0;
}

/* ------------------------------------------------------------------ codec2-named entry points */

extern "C" int ldpc_codes_num(void) {
    std::lock_guard<std::mutex> lk(g_codes_mu);
    return (int)codes().size();
}

extern "C" void ldpc_codes_list(void) {
    std::lock_guard<std::mutex> lk(g_codes_mu);
    fprintf(stderr, "\n");
    for (const CodeEntry &e : codes()) {
        const int K = e.CodeLength - e.NumberParityBits;
        fprintf(stderr, "%-20s rate %3.2f (%d,%d) \n", e.name.c_str(), (float)K / e.CodeLength, e.CodeLength, K);
    }
    fprintf(stderr, "\n");
}

/* A stand-in answers to the genuine name (pirip's command lines run unchanged, [R README.md:184]) but says so, loudly, once:
   frames coded with it do not interoperate with codec2 / real pirip radios.  `<name>_STANDIN` selects it silently. */
extern "C" int ldpc_codes_find(char name[]) {
    if (!name) return -1;
    std::lock_guard<std::mutex> lk(g_codes_mu);
    const std::string want(name);
    for (size_t i = 0; i < codes().size(); ++i) {
        CodeEntry &e = codes()[i];
        if (e.name == want) {
            if (e.standin && !e.warned && !getenv("PIRIP_LDPC_STANDIN_OK")) {
                e.warned = true;
                fprintf(stderr,
                        "pirip_b200: WARNING: LDPC code %s is a STAND-IN table (same N, K and layout, NOT codec2's parity-check matrix):\n"
                        "pirip_b200: WARNING: frames coded with it do not decode on codec2 / real pirip radios and theirs do not decode here.\n"
                        "pirip_b200: WARNING: load codec2's table with PIRIP_LDPC_TABLES=<codec2>/src/%s.h:<codec2>/src/%s.c or ldpc_codes_add().\n",
                        name, name, name);
            }
            return (int)i;
        }
        if (e.standin && e.name + "_STANDIN" == want) return (int)i;
    }
    return -1;
}

extern "C" void ldpc_codes_setup(struct LDPC *ldpc, char name[]) {
    const int idx = ldpc_codes_find(name);
    assert(idx != -1); /* as upstream: an unknown code name is a programming error */
    std::lock_guard<std::mutex> lk(g_codes_mu);
    CodeEntry &e = codes()[(size_t)idx];
    memset(ldpc, 0, sizeof *ldpc);
    strncpy(ldpc->name, e.name.c_str(), sizeof(ldpc->name) - 1);
    ldpc->max_iter = e.max_iter;
    ldpc->dec_type = 0;
    ldpc->q_scale_factor = 1;
    ldpc->r_scale_factor = 1;
    ldpc->CodeLength = e.CodeLength;
    ldpc->NumberParityBits = e.NumberParityBits;
    ldpc->NumberRowsHcols = e.NumberRowsHcols;
    ldpc->max_row_weight = e.max_row_weight;
    ldpc->max_col_weight = e.max_col_weight;
    ldpc->H_rows = e.H_rows.data();
    ldpc->H_cols = e.H_cols.data();
    ldpc->ldpc_data_bits_per_frame = e.CodeLength - e.NumberParityBits;
    ldpc->ldpc_coded_bits_per_frame = e.CodeLength;
    ldpc->protection_mode = 0;
    ldpc->data_bits_per_frame = ldpc->ldpc_data_bits_per_frame;
    ldpc->coded_bits_per_frame = ldpc->ldpc_coded_bits_per_frame;
}

extern "C" int ldpc_codes_add(const char *name, int max_iter, int CodeLength, int NumberParityBits, int NumberRowsHcols,
                              int max_row_weight, int max_col_weight, const uint16_t *H_rows, const uint16_t *H_cols) {
    if (!name || !H_rows || !H_cols) return -1;
    CodeEntry e;
    e.name = name;
    e.max_iter = max_iter;
    e.CodeLength = CodeLength;
    e.NumberParityBits = NumberParityBits;
    e.NumberRowsHcols = NumberRowsHcols;
    e.max_row_weight = max_row_weight;
    e.max_col_weight = max_col_weight;
    e.H_rows.assign(H_rows, H_rows + (size_t)NumberParityBits * max_row_weight);
    e.H_cols.assign(H_cols, H_cols + (size_t)NumberRowsHcols * max_col_weight);
    std::lock_guard<std::mutex> lk(g_codes_mu);
    codes().reserve(64); /* keep table pointers handed out by ldpc_codes_setup stable */
    for (CodeEntry &o : codes())
        if (o.name == e.name && o.standin) o.name += "_STANDIN"; /* a real table takes the genuine name over */
    codes().push_back(std::move(e));
    return (int)codes().size() - 1;
}

extern "C" void ldpc_encode_frame(struct LDPC *ldpc, unsigned char ibits[], unsigned char pbits[]) {
    unsigned prev = 0;
    const int NPB = ldpc->NumberParityBits;
    for (int p = 0; p < NPB; ++p) {
        unsigned par = 0;
        for (int i = 0; i < ldpc->max_row_weight; ++i) {
            const int ind = ldpc->H_rows[p + i * NPB];
            if (ind) par += ibits[ind - 1];
        }
        prev = (par + prev) & 1u; /* accumulate along the staircase */
        pbits[p] = (unsigned char)prev;
    }
}

extern "C" int run_ldpc_decoder(struct LDPC *ldpc, uint8_t out_char[], float input[], int *parityCheckCount) {
    /* one plan per call site would be wasteful: cache the plan for the last table seen */
    static thread_local pirip_ldpc_plan *cached = nullptr;
    static thread_local const uint16_t *cached_rows = nullptr;
    static thread_local int cached_iter = 0;
    if (!cached || cached_rows != ldpc->H_rows || cached_iter != ldpc->max_iter) {
        pirip_ldpc_plan_destroy(cached);
        cached = pirip_ldpc_plan_create(ldpc, pirip::current_device());
        cached_rows = ldpc->H_rows;
        cached_iter = ldpc->max_iter;
        if (!cached) {
            fprintf(stderr, "libpirip_b200: run_ldpc_decoder: %s\n", pirip::last_error());
            abort();
        }
    }
    int32_t iter = 0, pcc = 0;
    if (pirip_ldpc_decode_batch(cached, input, 1, out_char, &iter, &pcc) != 0) {
        fprintf(stderr, "libpirip_b200: run_ldpc_decoder: %s\n", pirip::last_error());
        abort();
    }
    if (parityCheckCount) *parityCheckCount = pcc;
    return iter;
}

extern "C" void fsk_rx_filt_to_llrs(float llr[], float rx_filt[], float v_est, float SNRest, int M, int nsyms) {
    const int bps = M == 2 ? 1 : 2;
    float *d_f = nullptr, *d_l = nullptr;
    bool ok = pirip::cuda_ok(cudaMalloc(&d_f, sizeof(float) * (size_t)(M * nsyms)), "cudaMalloc") &&
              pirip::cuda_ok(cudaMalloc(&d_l, sizeof(float) * (size_t)(bps * nsyms)), "cudaMalloc") &&
              pirip::cuda_ok(cudaMemcpy(d_f, rx_filt, sizeof(float) * (size_t)(M * nsyms), cudaMemcpyHostToDevice), "H2D");
    if (ok) {
        llr_kernel<<<(nsyms + 127) / 128, 128>>>(d_f, d_l, v_est, SNRest, M, nsyms);
        pirip::count_launch();
        ok = pirip::cuda_ok(cudaGetLastError(), "llr_kernel launch") &&
             pirip::cuda_ok(cudaMemcpy(llr, d_l, sizeof(float) * (size_t)(bps * nsyms), cudaMemcpyDeviceToHost), "llr_kernel");
    }
    cudaFree(d_f);
    cudaFree(d_l);
    if (!ok) {
        fprintf(stderr, "libpirip_b200: fsk_rx_filt_to_llrs: %s\n", pirip::last_error());
        abort();
    }
}

/* ------------------------------------------------------------------ framing (host) */

/* CRC-16/CCITT-FALSE, byte-wise form */
extern "C" unsigned short freedv_gen_crc16(unsigned char *data_p, int length) {
    unsigned short crc = 0xFFFF;
    while (length-- > 0) {
        unsigned char x = (unsigned char)((crc >> 8) ^ *data_p++);
        x ^= x >> 4;
        crc = (unsigned short)((crc << 8) ^ ((unsigned short)x << 12) ^ ((unsigned short)x << 5) ^ (unsigned short)x);
    }
    return crc;
}

extern "C" void freedv_pack(unsigned char *bytes, unsigned char *bits, int nbits) {
    memset(bytes, 0, (size_t)((nbits + 7) / 8));
    int bit = 7, byte = 0;
    for (int i = 0; i < nbits; ++i) {
        bytes[byte] |= (unsigned char)(bits[i] << bit);
        if (--bit < 0) {
            bit = 7;
            byte++;
        }
    }
}

extern "C" void freedv_unpack(unsigned char *bits, unsigned char *bytes, int nbits) {
    int bit = 7, byte = 0;
    for (int i = 0; i < nbits; ++i) {
        bits[i] = (bytes[byte] >> bit) & 0x1;
        if (--bit < 0) {
            bit = 7;
            byte++;
        }
    }
}

extern "C" int freedv_check_crc16_unpacked(unsigned char *bits, int nbits) {
    assert((nbits % 8) == 0);
    const int nbytes = nbits / 8;
    std::vector<unsigned char> bytes((size_t)nbytes);
    freedv_pack(bytes.data(), bits, nbits);
    const unsigned short tx = (unsigned short)((bytes[(size_t)nbytes - 2] << 8) | bytes[(size_t)nbytes - 1]);
    return tx == freedv_gen_crc16(bytes.data(), nbytes - 2);
}

extern "C" void ofdm_generate_payload_data_bits(uint8_t payload_data_bits[], int n) {
    /* [U ofdm.c::ofdm_rand] LCG from seed 1, bit = r > 16384 */
    uint64_t seed = 1;
    for (int i = 0; i < n; ++i) {
        seed = (1103515245ull * seed + 12345) % 32768;
        payload_data_bits[i] = (uint16_t)seed > 16384;
    }
}

static const uint8_t fsk_ldpc_uw[32] = { 0, 1, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0,
                                         1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 0, 1, 0, 1 };

struct freedv {
    struct LDPC ldpc;
    int M, Rs, Fs;
    int bits_per_modem_frame; /* payload bits: K */
    int bits_per_frame;       /* UW + codeword */
    /* receive side */
    std::vector<float> twoframes_llr, frame_llr;
    std::vector<uint8_t> twoframes_hard;
    int frame_llr_nbits = 0;
    int state = 0, best_location = 0, baduw = 0;
    int thresh1 = 5, thresh2 = 6, baduw_thresh = 1;
    int last_uw_err = 0, last_iter = 0, last_pcc = 0;
};

extern "C" struct freedv *freedv_open_advanced(int mode, struct freedv_advanced *adv) {
    if (mode != FREEDV_MODE_FSK_LDPC || !adv || !adv->codename) return NULL;
    if (ldpc_codes_find(adv->codename) < 0) return NULL;
    struct freedv *f = new freedv();
    ldpc_codes_setup(&f->ldpc, adv->codename);
    f->M = adv->M;
    f->Rs = adv->Rs;
    f->Fs = adv->Fs;
    f->bits_per_modem_frame = f->ldpc.data_bits_per_frame;
    f->bits_per_frame = 32 + f->ldpc.coded_bits_per_frame;
    f->twoframes_llr.assign((size_t)(2 * f->bits_per_frame), 0.0f);
    f->twoframes_hard.assign((size_t)(2 * f->bits_per_frame), 0);
    return f;
}

extern "C" void freedv_close(struct freedv *f) { delete f; }
extern "C" int freedv_get_bits_per_modem_frame(struct freedv *f) { return f->bits_per_modem_frame; }
extern "C" int freedv_tx_fsk_ldpc_bits_per_frame(struct freedv *f) { return f->bits_per_frame; }

extern "C" void freedv_tx_fsk_ldpc_framer(struct freedv *f, uint8_t frame[], uint8_t payload_data[]) {
    memcpy(frame, fsk_ldpc_uw, sizeof fsk_ldpc_uw);
    memcpy(frame + sizeof fsk_ldpc_uw, payload_data, (size_t)f->bits_per_modem_frame);
    ldpc_encode_frame(&f->ldpc, frame + sizeof fsk_ldpc_uw, frame + sizeof fsk_ldpc_uw + f->bits_per_modem_frame);
}

extern "C" int freedv_fsk_ldpc_rx_llrs(struct freedv *f, const float llr[], int nbits, uint8_t payload_bits[]) {
    const int bpf = f->bits_per_frame;
    int rx_status = 0;
    /* queue the modem frame; run the sync state machine once per whole frame of new LLRs */
    f->frame_llr.insert(f->frame_llr.end(), llr, llr + nbits);
    f->frame_llr_nbits += nbits;
    if (f->frame_llr_nbits >= bpf) {
        memmove(f->twoframes_llr.data(), f->twoframes_llr.data() + bpf, sizeof(float) * (size_t)bpf);
        memcpy(f->twoframes_llr.data() + bpf, f->frame_llr.data(), sizeof(float) * (size_t)bpf);
        f->frame_llr.erase(f->frame_llr.begin(), f->frame_llr.begin() + bpf);
        f->frame_llr_nbits -= bpf;
        for (int i = 0; i < 2 * bpf; ++i) f->twoframes_hard[(size_t)i] = f->twoframes_llr[(size_t)i] < 0;

        int next_state = f->state, errors = 0;
        if (f->state == 0) {
            int best_errors = 32;
            f->best_location = 0;
            for (int i = 0; i < bpf; ++i) {
                errors = 0;
                for (int u = 0; u < 32; ++u) errors += f->twoframes_hard[(size_t)(i + u)] ^ fsk_ldpc_uw[u];
                if (errors < best_errors) {
                    best_errors = errors;
                    f->best_location = i;
                }
            }
            if (best_errors <= f->thresh1) {
                errors = best_errors;
                next_state = 1;
                f->baduw = 0;
            }
        } else {
            for (int u = 0; u < 32; ++u) errors += f->twoframes_hard[(size_t)(f->best_location + u)] ^ fsk_ldpc_uw[u];
            if (errors > f->thresh2) {
                f->baduw++;
                if (f->baduw >= f->baduw_thresh) next_state = 0;
            } else {
                f->baduw = 0;
            }
        }
        f->last_uw_err = errors;
        f->last_iter = f->last_pcc = 0;
        if (next_state == 1) {
            std::vector<uint8_t> decoded((size_t)f->ldpc.CodeLength);
            int pcc = 0;
            const int iter = run_ldpc_decoder(&f->ldpc, decoded.data(), &f->twoframes_llr[(size_t)(f->best_location + 32)], &pcc);
            f->last_iter = iter;
            f->last_pcc = pcc;
            if (pcc != f->ldpc.NumberParityBits) rx_status |= FREEDV_RX_BIT_ERRORS;
            if (freedv_check_crc16_unpacked(decoded.data(), f->bits_per_modem_frame)) {
                rx_status |= FREEDV_RX_BITS;
                memcpy(payload_bits, decoded.data(), (size_t)f->bits_per_modem_frame);
            } else if (f->state == 0) {
                next_state = 0; /* CRC failed on the first frame after a UW hit: dud UW */
            }
        }
        f->state = next_state;
    }
    if (f->state == 1) rx_status |= FREEDV_RX_SYNC;
    return rx_status;
}

extern "C" void freedv_fsk_ldpc_rx_info(struct freedv *f, int info[6]) {
    info[0] = f->state;
    info[1] = f->best_location;
    info[2] = f->last_uw_err;
    info[3] = f->last_iter;
    info[4] = f->last_pcc;
    info[5] = f->frame_llr_nbits;
}

/* ------------------------------------------------------------------ batched coded receive chain (config 4) */

namespace {

struct RxArgs {
    const float *llr;      /* [ncap][llr_stride] the demodulator's LLR stream of every capture */
    size_t llr_stride;
    const int32_t *nframes; /* [ncap] modem frames demodulated */
    size_t ncap, max_frames;
    int Nbits, bpf, K, thresh1, thresh2, baduw_thresh;
    unsigned uw_word;      /* bit u = UW bit u */
    uint8_t *status;       /* [ncap][max_frames] rx_status of every modem frame */
    uint8_t *payload;      /* [ncap][max_frames][K/8] packed payload where FREEDV_RX_BITS is set (zero elsewhere) */
    int32_t *info;         /* optional [ncap][max_frames][4]: uw_loc, uw_err, iter, pcc of the frame step taken at that modem frame */
};

/* FSK_LDPC receive side of every capture of a batch in ONE launch ([U freedv_fsk.c::freedv_rx_fsk_ldpc_data], the state
   machine of freedv_fsk_ldpc_rx_llrs below; `rtl_fsk --code H_256_512_4 -b` [R README.md:184,292], record layout
   [R tx/frame_repeater.c:55-62]).  One WARP per capture: it walks the capture's modem frames, runs the UW search over the
   two-frame window (hard bits packed by ballot, 32 start positions per step), the sync state machine, the LDPC decoder
   (ldpc_decode_one, the same routine as the batch decoder) on the frame after the UW, and the CRC-16 check; it writes one
   status byte and - when the CRC passes - the packed payload per modem frame.  The two-frame window and the LLR queue of
   the scalar code are views of the capture's LLR stream, so nothing is buffered. */
template <int MODE>
__global__ void __launch_bounds__(32 * kLdpcMaxWarps) fsk_ldpc_rx_kernel(const __grid_constant__ LdpcGraph g, const __grid_constant__ RxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ldpc_load_tables(g, smem_raw);
    __syncthreads();
    if (warp >= g.warps) return;
    const LdpcSmem m = ldpc_smem_view(g, smem_raw, warp);
    unsigned *hw = m.scr;        /* hard-bit words of the two-frame window */
    unsigned *dec = m.scr + 160; /* decoded bits, bit v of word v / 32 */
    const int bpf = a.bpf, nhw = (2 * bpf + 31) / 32;
    for (size_t cap = (size_t)blockIdx.x * g.warps + warp; cap < a.ncap; cap += (size_t)gridDim.x * g.warps) {
        const float *S = a.llr + cap * a.llr_stride;
        const int nf = a.nframes[cap];
        int state = 0, best_loc = 0, baduw = 0, t = 0;
        long long nq = 0;
        for (int f = 0; f < nf; ++f) {
            nq += a.Nbits;
            int rx_status = 0;
            int uw_err = 0, iters = 0, pccv = 0, stepped = 0;
            if (nq >= bpf) {
                stepped = 1;
                const long long base = (long long)(t - 1) * bpf; /* stream index of the window's first LLR (zeros before the stream) */
                for (int k = 0; k < nhw; ++k) {
                    const int i = 32 * k + lane;
                    const long long idx = base + i;
                    const float x = (idx >= 0 && i < 2 * bpf) ? S[idx] : 0.0f;
                    const unsigned w = __ballot_sync(0xffffffffu, x < 0.0f);
                    if (lane == 0) hw[k] = w;
                }
                if (lane == 0) hw[nhw] = 0;
                __syncwarp();
                auto uw_errors = [&](int i) {
                    const unsigned lo = hw[i >> 5], hi = hw[(i >> 5) + 1];
                    const unsigned win = __funnelshift_r(lo, hi, i & 31);
                    return __popc(win ^ a.uw_word);
                };
                int next_state = state, errors = 0;
                if (state == 0) {
                    int be = 32, bl = 0;
                    for (int i = lane; i < bpf; i += 32) {
                        const int e = uw_errors(i);
                        if (e < be) {
                            be = e;
                            bl = i;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const int oe = __shfl_xor_sync(0xffffffffu, be, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
                        if (oe < be || (oe == be && ol < bl)) {
                            be = oe;
                            bl = ol;
                        }
                    }
                    best_loc = be < 32 ? bl : 0; /* the scalar search starts from best_errors = 32, location 0 */
                    if (be <= a.thresh1) {
                        errors = be;
                        next_state = 1;
                        baduw = 0;
                    }
                } else {
                    errors = uw_errors(best_loc);
                    if (errors > a.thresh2) {
                        baduw++;
                        if (baduw >= a.baduw_thresh) next_state = 0;
                    } else
                        baduw = 0;
                }
                uw_err = errors;
                if (next_state == 1) {
                    for (int k = lane; k < (g.N + 31) / 32; k += 32) dec[k] = 0;
                    __syncwarp();
                    const long long cw0 = base + best_loc + 32;
                    iters = ldpc_decode_one<MODE>(
                        g, m, lane, [&](int v) { return (cw0 + v >= 0) ? S[cw0 + v] : 0.0f; },
                        [&](int v, int bit) {
                            if (bit) atomicOr(&dec[v >> 5], 1u << (v & 31));
                        },
                        &pccv);
                    __syncwarp();
                    if (pccv != g.NPB) rx_status |= FREEDV_RX_BIT_ERRORS;
                    /* CRC-16/CCITT-FALSE of the payload but its last two bytes ([U freedv_api.c::freedv_gen_crc16]); bytes MSB first */
                    const int nbytes = a.K / 8;
                    auto byte_at = [&](int b) { return (unsigned)(__brev((dec[b >> 2] >> (8 * (b & 3))) & 0xffu) >> 24); };
                    int crc_ok = 0;
                    if (lane == 0) {
                        unsigned crc = 0xFFFF;
                        for (int b = 0; b < nbytes - 2; ++b) {
                            unsigned x = ((crc >> 8) ^ byte_at(b)) & 0xff;
                            x ^= x >> 4;
                            crc = ((crc << 8) ^ (x << 12) ^ (x << 5) ^ x) & 0xffff;
                        }
                        crc_ok = crc == ((byte_at(nbytes - 2) << 8) | byte_at(nbytes - 1));
                    }
                    crc_ok = __shfl_sync(0xffffffffu, crc_ok, 0);
                    if (crc_ok) {
                        rx_status |= FREEDV_RX_BITS;
                        uint8_t *po = a.payload + (cap * a.max_frames + (size_t)f) * (size_t)nbytes;
                        for (int b = lane; b < nbytes; b += 32) po[b] = (uint8_t)byte_at(b);
                    } else if (state == 0)
                        next_state = 0; /* CRC failed on the first frame after a UW hit: dud UW */
                }
                state = next_state;
                t++;
                nq -= bpf;
            }
            if (state == 1) rx_status |= FREEDV_RX_SYNC;
            if (lane == 0) {
                a.status[cap * a.max_frames + (size_t)f] = (uint8_t)rx_status;
                if (a.info) {
                    int32_t *io = a.info + (cap * a.max_frames + (size_t)f) * 4;
                    io[0] = stepped ? best_loc : -1;
                    io[1] = uw_err;
                    io[2] = iters;
                    io[3] = pccv;
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace

struct pirip_rx_plan {
    pirip_fsk_plan *fsk = nullptr;
    pirip_ldpc_plan *ldpc = nullptr;
    struct LDPC code;
    pirip_fsk_cfg cfg;
    int device = 0, Nbits = 0, bpf = 0, K = 0;
    /* device buffers of the host-buffer entry point */
    unsigned char *d_iq = nullptr;
    float *d_llr = nullptr;
    uint8_t *d_status = nullptr, *d_payload = nullptr;
    int32_t *d_nf = nullptr, *d_info = nullptr;
    size_t cap_iq = 0, cap_llr = 0, cap_out = 0, cap_n = 0;
    cudaStream_t stream = nullptr;
};

extern "C" void pirip_fsk_ldpc_rx_plan_destroy(pirip_rx_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    pirip_fsk_plan_destroy(p->fsk);
    pirip_ldpc_plan_destroy(p->ldpc);
    cudaFree(p->d_iq);
    cudaFree(p->d_llr);
    cudaFree(p->d_status);
    cudaFree(p->d_payload);
    cudaFree(p->d_nf);
    cudaFree(p->d_info);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" pirip_rx_plan *pirip_fsk_ldpc_rx_plan_create(const struct pirip_fsk_cfg *cfg, const char *codename, int device) {
    if (!cfg || !codename) {
        pirip::set_error("cfg / codename is NULL");
        return nullptr;
    }
    std::string name(codename);
    if (ldpc_codes_find(&name[0]) < 0) {
        pirip::set_error("unknown LDPC code " + name);
        return nullptr;
    }
    pirip_rx_plan *p = new pirip_rx_plan();
    p->device = device;
    p->cfg = *cfg;
    p->cfg.soft = 1;
    ldpc_codes_setup(&p->code, &name[0]);
    p->fsk = pirip_fsk_plan_create(&p->cfg, device);
    p->ldpc = p->fsk ? pirip_ldpc_plan_create(&p->code, device) : nullptr;
    if (!p->fsk || !p->ldpc) {
        pirip_fsk_ldpc_rx_plan_destroy(p);
        return nullptr;
    }
    p->Nbits = pirip_fsk_plan_nbits(p->fsk);
    p->K = p->code.CodeLength - p->code.NumberParityBits;
    p->bpf = 32 + p->code.CodeLength;
    if (2 * p->bpf / 32 + 2 > 160 || (p->code.CodeLength + 31) / 32 > 96 || (p->K % 8) != 0) {
        pirip::set_error("code too long for the receive-chain kernel's scratch");
        pirip_fsk_ldpc_rx_plan_destroy(p);
        return nullptr;
    }
    if (!pirip::cuda_ok(cudaFuncSetAttribute(fsk_ldpc_rx_kernel<LDPC_SP_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->ldpc->smem), "smem attr") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(fsk_ldpc_rx_kernel<LDPC_SP_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->ldpc->smem), "smem attr") ||
        !pirip::cuda_ok(cudaFuncSetAttribute(fsk_ldpc_rx_kernel<LDPC_MINSUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->ldpc->smem), "smem attr")) {
        pirip_fsk_ldpc_rx_plan_destroy(p);
        return nullptr;
    }
    return p;
}

extern "C" pirip_fsk_plan *pirip_fsk_ldpc_rx_plan_fsk(pirip_rx_plan *p) { return p ? p->fsk : nullptr; }
extern "C" pirip_ldpc_plan *pirip_fsk_ldpc_rx_plan_ldpc(pirip_rx_plan *p) { return p ? p->ldpc : nullptr; }
extern "C" int pirip_fsk_ldpc_rx_plan_payload_bytes(const pirip_rx_plan *p) { return p ? p->K / 8 : -1; }

extern "C" int pirip_fsk_ldpc_rx_batch_dev(pirip_rx_plan *p, const void *d_iq, size_t n_captures, size_t samples_per_capture,
                                           size_t capture_stride_bytes, float *d_llr, int32_t *d_nframes, uint8_t *d_status, uint8_t *d_payload,
                                           int32_t *d_info, size_t max_frames, void *stream) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (n_captures == 0) return 0;
    if (!d_iq || !d_llr || !d_nframes || !d_status || !d_payload) return pirip::fail(-2, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    /* demodulate: LLRs of every modem frame, [n][max_frames * Nbits] */
    int rc = pirip_fsk_demod_batch_dev(p->fsk, d_iq, n_captures, samples_per_capture, capture_stride_bytes, nullptr, d_llr, nullptr, nullptr,
                                       d_nframes, max_frames, st);
    if (rc) return rc;
    const size_t nbytes = (size_t)p->K / 8;
    if (!pirip::cuda_ok(cudaMemsetAsync(d_payload, 0, n_captures * max_frames * nbytes, st), "memset payload") ||
        !pirip::cuda_ok(cudaMemsetAsync(d_status, 0, n_captures * max_frames, st), "memset status"))
        return -12;
    RxArgs a;
    memset(&a, 0, sizeof a);
    a.llr = d_llr;
    a.llr_stride = max_frames * (size_t)p->Nbits;
    a.nframes = d_nframes;
    a.ncap = n_captures;
    a.max_frames = max_frames;
    a.Nbits = p->Nbits;
    a.bpf = p->bpf;
    a.K = p->K;
    a.thresh1 = 5;
    a.thresh2 = 6;
    a.baduw_thresh = 1;
    a.uw_word = 0;
    for (int u = 0; u < 32; ++u) a.uw_word |= (unsigned)fsk_ldpc_uw[u] << u;
    a.status = d_status;
    a.payload = d_payload;
    a.info = d_info;
    const LdpcGraph &g = p->ldpc->g;
    const size_t ctas_needed = (n_captures + g.warps - 1) / g.warps;
    const size_t resident = (size_t)p->ldpc->num_sms * ((p->ldpc->smem + 1024) * 2 <= 228 * 1024 ? 2 : 1);
    const unsigned grid = (unsigned)(ctas_needed < resident ? ctas_needed : resident);
    const unsigned block = 32u * (unsigned)g.warps;
    if (p->ldpc->mode == LDPC_MINSUM)
        fsk_ldpc_rx_kernel<LDPC_MINSUM><<<grid, block, p->ldpc->smem, st>>>(g, a);
    else if (p->ldpc->mode == LDPC_SP_FAST)
        fsk_ldpc_rx_kernel<LDPC_SP_FAST><<<grid, block, p->ldpc->smem, st>>>(g, a);
    else
        fsk_ldpc_rx_kernel<LDPC_SP_EXACT><<<grid, block, p->ldpc->smem, st>>>(g, a);
    pirip::count_launch();
    return pirip::cuda_ok(cudaGetLastError(), "fsk_ldpc_rx_kernel launch") ? 0 : -11;
}

extern "C" int pirip_fsk_ldpc_rx_batch(pirip_rx_plan *p, const void *iq, size_t n, size_t samples_per_capture, uint8_t *status, uint8_t *payload,
                                       int32_t *nframes, int32_t *info, size_t max_frames) {
    if (!p) return pirip::fail(-1, "plan is NULL");
    if (n == 0) return 0;
    if (!iq || !status || !payload) return pirip::fail(-2, "NULL buffer");
    if (!pirip::cuda_ok(cudaSetDevice(p->device), "cudaSetDevice")) return -10;
    const size_t bps = p->cfg.fmt == PIRIP_FMT_U8 ? 2 : (p->cfg.fmt == PIRIP_FMT_S16 ? 4 : 8);
    const size_t cap_bytes = samples_per_capture * bps, stride = (cap_bytes + 15) / 16 * 16;
    const size_t nbytes = (size_t)p->K / 8;
    auto ensure = [&](void **ptr, size_t &have, size_t need) {
        if (need <= have && *ptr) return true;
        cudaFree(*ptr);
        *ptr = nullptr;
        have = 0;
        if (!pirip::cuda_ok(cudaMalloc(ptr, need), "cudaMalloc")) return false;
        have = need;
        return true;
    };
    if (!p->stream && !pirip::cuda_ok(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return -12;
    if (!ensure((void **)&p->d_iq, p->cap_iq, stride * n) || !ensure((void **)&p->d_llr, p->cap_llr, sizeof(float) * n * max_frames * p->Nbits))
        return -12;
    if (n * max_frames > p->cap_out) {
        size_t c1 = 0, c2 = 0, c3 = 0;
        if (!ensure((void **)&p->d_status, c1, n * max_frames) || !ensure((void **)&p->d_payload, c2, n * max_frames * nbytes) ||
            !ensure((void **)&p->d_info, c3, sizeof(int32_t) * 4 * n * max_frames))
            return -12;
        p->cap_out = n * max_frames;
    }
    if (n > p->cap_n) {
        size_t c4 = 0;
        if (!ensure((void **)&p->d_nf, c4, sizeof(int32_t) * n)) return -12;
        p->cap_n = n;
    }
    cudaStream_t st = p->stream;
    if (!pirip::cuda_ok(cudaMemcpy2DAsync(p->d_iq, stride, iq, cap_bytes, cap_bytes, n, cudaMemcpyHostToDevice, st), "H2D iq")) return -12;
    if (!pirip::cuda_ok(cudaMemsetAsync(p->d_llr, 0, sizeof(float) * n * max_frames * p->Nbits, st), "memset llr")) return -12;
    const int rc = pirip_fsk_ldpc_rx_batch_dev(p, p->d_iq, n, samples_per_capture, stride, p->d_llr, p->d_nf, p->d_status, p->d_payload,
                                               info ? p->d_info : nullptr, max_frames, st);
    if (rc) return rc;
    if (!pirip::cuda_ok(cudaMemcpyAsync(status, p->d_status, n * max_frames, cudaMemcpyDeviceToHost, st), "D2H status") ||
        !pirip::cuda_ok(cudaMemcpyAsync(payload, p->d_payload, n * max_frames * nbytes, cudaMemcpyDeviceToHost, st), "D2H payload") ||
        (nframes && !pirip::cuda_ok(cudaMemcpyAsync(nframes, p->d_nf, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st), "D2H nframes")) ||
        (info && !pirip::cuda_ok(cudaMemcpyAsync(info, p->d_info, sizeof(int32_t) * 4 * n * max_frames, cudaMemcpyDeviceToHost, st), "D2H info")))
        return -12;
    return pirip::cuda_ok(cudaStreamSynchronize(st), "receive chain") ? 0 : -13;
}

