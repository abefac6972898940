// capi.cu -- library-level entry points and error plumbing of libxrs.so.
#include <atomic>

#include "common.cuh"

namespace xrs {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const std::string &msg) { g_last_error = msg; }

int fail(const std::string &msg) {
    g_last_error = msg;
    return 1;
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return 2;
}

}  // namespace xrs

extern "C" {

int xrs_version(void) { return XRS_VERSION; }

int xrs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *xrs_last_error(void) { return xrs::g_last_error.c_str(); }

uint64_t xrs_launch_count(void) { return xrs::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
