// Are add.rn.f32x2 / mul.rn.f32x2 bit-identical to add.rn.f32 / mul.rn.f32 on sm_100a?  (no ftz, no fma)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ unsigned hashu(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k(unsigned long long *bad, int mode) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nb = 0;
    for (int it = 0; it < 4096; ++it) {
        unsigned a0 = hashu(id * 4099u + it * 7919u + 1u), a1 = hashu(a0 + 0x9e3779b9u), b0 = hashu(a1 + 17u), b1 = hashu(b0 + 33u);
        if (mode == 1) { /* similar exponents: cancellation and subnormal results */
            b0 = (a0 & 0xff800000u) ^ 0x80000000u | (b0 & 0x007fffffu);
            b1 = (a1 & 0xff800000u) | (b1 & 0x007fffffu);
        }
        if (mode == 2) { a0 &= 0x807fffffu; a1 = (a1 & 0x80ffffffu); b0 = (b0 & 0x80ffffffu) | 0x00800000u; } /* subnormal inputs */
        const float fa0 = __uint_as_float(a0), fa1 = __uint_as_float(a1), fb0 = __uint_as_float(b0), fb1 = __uint_as_float(b1);
        const u64 pa = ((u64)a1 << 32) | a0, pb = ((u64)b1 << 32) | b0;
        const u64 s = add2(pa, pb), p = mul2(pa, pb);
        const unsigned s0 = __float_as_uint(__fadd_rn(fa0, fb0)), s1 = __float_as_uint(__fadd_rn(fa1, fb1));
        const unsigned p0 = __float_as_uint(__fmul_rn(fa0, fb0)), p1 = __float_as_uint(__fmul_rn(fa1, fb1));
        const bool nan_s0 = (s0 & 0x7fffffffu) > 0x7f800000u, nan_s1 = (s1 & 0x7fffffffu) > 0x7f800000u;
        const bool nan_p0 = (p0 & 0x7fffffffu) > 0x7f800000u, nan_p1 = (p1 & 0x7fffffffu) > 0x7f800000u;
        if (!nan_s0 && (unsigned)s != s0) nb++;
        if (!nan_s1 && (unsigned)(s >> 32) != s1) nb++;
        if (!nan_p0 && (unsigned)p != p0) nb++;
        if (!nan_p1 && (unsigned)(p >> 32) != p1) nb++;
    }
    if (nb) atomicAdd(bad, nb);
}
int main() {
    unsigned long long *bad, h;
    cudaMalloc(&bad, 8);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(bad, 0, 8);
        k<<<1024, 256>>>(bad, mode);
        cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %llu mismatches of %llu packed results (%s)\n", mode, h, 1024ull * 256 * 4096 * 4, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
