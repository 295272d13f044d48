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
 * Mapping: one CTA produces TILE consecutive outputs of one stream.  Its D*(TILE-1)+L input
 * samples are read once from HBM (coalesced 4-byte loads), converted through a 256-entry
 * table of the exact convert_u8_f values and staged as float2 in shared memory; every thread
 * then runs the L-tap loop for one output out of shared memory.  HBM traffic is therefore the
 * algorithmic 2 B per input sample (+ L-D samples of halo per tile) and 8/D B of output.
 */
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

size_t out_count(size_t n, int D, int L) { return n >= (size_t)L ? (n - (size_t)L) / (size_t)D + 1 : 0; }

}  // namespace

struct pirip_decim_plan {
    int D, L, device;
    float *d_taps = nullptr, *d_lut = nullptr;
    int smem_u8 = 0, smem_f32 = 0;
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
