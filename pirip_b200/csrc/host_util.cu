/* host_util.cu — error reporting, launch accounting, library identification */
#include "host_util.h"

#include <atomic>

#include "../../include/pirip_fsk.h"

namespace {
thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{ 0 };
}  // namespace

namespace pirip {

void set_error(const std::string &msg) { g_last_error = msg; }

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    return dev;
}

const char *last_error() { return g_last_error.c_str(); }

}  // namespace pirip

extern "C" uint64_t pirip_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char *pirip_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *pirip_version(void) { return "pirip_b200 0.2 (sm_100a; fast kernel with exact in-kernel replay + exact-order kernel)"; }

/* ---- bit packing for the gather of decoded bit streams (config 5): 8 one-bit-per-byte values -> one byte, MSB first, the
   order of codec2's freedv_pack ([R tx/rpitx_fsk.cpp:38,75-83]) ---- */
namespace {
__global__ void pack_bits_kernel(const uint8_t *__restrict__ bits, size_t nbytes_out, uint8_t *__restrict__ out) {
    /* one thread: 16 input bytes -> 2 output bytes */
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t npair = nbytes_out >> 1;
    if (i < npair) {
        const uint4 v = reinterpret_cast<const uint4 *>(bits)[i];
        auto pk8 = [](unsigned lo, unsigned hi) {
            /* bytes b0..b3 in lo (b0 least significant byte = first bit), b4..b7 in hi; each 0/1 */
            unsigned r = 0;
            r |= ((lo >> 0) & 1u) << 7;
            r |= ((lo >> 8) & 1u) << 6;
            r |= ((lo >> 16) & 1u) << 5;
            r |= ((lo >> 24) & 1u) << 4;
            r |= ((hi >> 0) & 1u) << 3;
            r |= ((hi >> 8) & 1u) << 2;
            r |= ((hi >> 16) & 1u) << 1;
            r |= ((hi >> 24) & 1u) << 0;
            return r;
        };
        reinterpret_cast<unsigned short *>(out)[i] = (unsigned short)(pk8(v.x, v.y) | (pk8(v.z, v.w) << 8));
    } else if (i == npair && (nbytes_out & 1)) {
        const uint2 v = reinterpret_cast<const uint2 *>(bits)[2 * npair];
        unsigned r = 0;
        for (int k = 0; k < 4; ++k) r |= ((v.x >> (8 * k)) & 1u) << (7 - k), r |= ((v.y >> (8 * k)) & 1u) << (3 - k);
        out[nbytes_out - 1] = (uint8_t)r;
    }
}
}  // namespace

extern "C" int pirip_pack_bits_dev(const uint8_t *d_bits, size_t nbits, uint8_t *d_packed, void *stream) {
    if (nbits == 0) return 0;
    if (!d_bits || !d_packed) return pirip::fail(-2, "pirip_pack_bits_dev: NULL buffer");
    if ((nbits & 7) != 0 || (reinterpret_cast<size_t>(d_bits) & 15) != 0 || (reinterpret_cast<size_t>(d_packed) & 1) != 0)
        return pirip::fail(-4, "pirip_pack_bits_dev: nbits must be a multiple of 8, d_bits 16-byte and d_packed 2-byte aligned");
    const size_t nbytes = nbits / 8, nthreads = nbytes / 2 + 1;
    pack_bits_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_bits, nbytes, d_packed);
    if (!pirip::cuda_ok(cudaGetLastError(), "pack_bits_kernel launch")) return -11;
    pirip::count_launch();
    return 0;
}
