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
extern "C" const char *pirip_version(void) { return "pirip_b200 0.1 (sm_100a; exact-order numerics)"; }
