// Microbenchmark: scalar vs packed (f32x2) FP32 add/mul issue rate and dependent-chain latency on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32x2 fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE, int CH>
__global__ void k(float *out, long long *cyc, int iters, float seed) {
    float a[CH]; u64 p[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    const float b = seed * 1.0001f; const u64 pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) a[i] = fadd(a[i], b);
                if (MODE == 1) p[i] = add2(p[i], pb);
                if (MODE == 2) a[i] = fmul(a[i], b);
                if (MODE == 3) p[i] = mul2(p[i], pb);
                if (MODE == 4) { float x = a[i]; float y = __uint_as_float((unsigned)p[i]);  /* scalar complex mul recursion */
                    float nx = fadd(fmul(x, b), -fmul(y, seed)); float ny = fadd(fmul(x, seed), fmul(y, b)); a[i] = nx; p[i] = __float_as_uint(ny); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0; 
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int CH>
void run(const char *name, int threads) {
    float *out; long long *cyc; cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<MODE, CH><<<148, threads>>>(out, cyc, iters, 1.0f);
    k<MODE, CH><<<148, threads>>>(out, cyc, iters, 1.0f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)iters * 8 * CH;            // per thread
    double warps = threads / 32.0;
    printf("%-10s chains=%d threads=%4d: %.2f cycles per op per thread-chain-step, %.2f warp-instr/clk/SM (%s)\n", name, CH, threads,
           (double)c / (iters * 8.0), ops * warps / (double)c * (MODE == 4 ? 6 : 1), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("FADD", 32); run<1, 1>("FADD2", 32); run<2, 1>("FMUL", 32); run<3, 1>("FMUL2", 32); run<4, 1>("cmul-chain", 32);
    run<0, 8>("FADD", 1024); run<1, 8>("FADD2", 1024); run<2, 8>("FMUL", 1024); run<3, 8>("FMUL2", 1024);
    run<0, 8>("FADD", 512); run<1, 8>("FADD2", 512);
    run<0, 4>("FADD", 256); run<1, 4>("FADD2", 256);
    run<4, 1>("cmul-chain", 128); run<4, 2>("cmul-chain", 128);
    return 0;
}
