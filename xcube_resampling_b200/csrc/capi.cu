// capi.cu -- library-level entry points and error plumbing of libxrs.so.
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

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

// ---- optional per-kernel timing ---------------------------------------------------------
struct ProfileRecord {
    const char *name;
    cudaEvent_t e0, e1;
};
static std::atomic<int> g_profile_on{0};
static std::mutex g_profile_mutex;
static std::vector<ProfileRecord> g_profile_records;
static thread_local int g_profile_open = -1;  // index of the record begun by this thread

void profile_begin(const char *name, cudaStream_t st) {
    g_profile_open = -1;
    if (!g_profile_on.load(std::memory_order_relaxed)) return;
    ProfileRecord r;
    r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    g_profile_records.push_back(r);
    g_profile_open = static_cast<int>(g_profile_records.size()) - 1;
}

void profile_end(cudaStream_t st) {
    if (g_profile_open < 0) return;
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    if (g_profile_open < static_cast<int>(g_profile_records.size()))
        cudaEventRecord(g_profile_records[g_profile_open].e1, st);
    g_profile_open = -1;
}

}  // namespace xrs

extern "C" {

int xrs_profile_enable(int32_t on) {
    xrs::g_profile_on.store(on ? 1 : 0);
    return 0;
}

int32_t xrs_profile_collect(char *names, int64_t names_len, double *total_ms, int64_t *launches, int32_t max_entries) {
    std::vector<xrs::ProfileRecord> recs;
    {
        std::lock_guard<std::mutex> lock(xrs::g_profile_mutex);
        recs.swap(xrs::g_profile_records);
    }
    std::vector<const char *> order;
    std::vector<double> ms;
    std::vector<int64_t> cnt;
    for (auto &r : recs) {
        float t = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
            size_t k = 0;
            while (k < order.size() && std::strcmp(order[k], r.name) != 0) ++k;
            if (k == order.size()) {
                order.push_back(r.name);
                ms.push_back(0.0);
                cnt.push_back(0);
            }
            ms[k] += t;
            cnt[k] += 1;
        } else {
            cudaGetLastError();
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    int32_t n = 0;
    int64_t pos = 0;
    if (names && names_len > 0) names[0] = 0;
    for (size_t k = 0; k < order.size() && n < max_entries; ++k) {
        const int64_t len = static_cast<int64_t>(std::strlen(order[k]));
        if (!names || pos + len + 2 > names_len) break;
        std::memcpy(names + pos, order[k], len);
        names[pos + len] = '\n';
        names[pos + len + 1] = 0;
        pos += len + 1;
        if (total_ms) total_ms[n] = ms[k];
        if (launches) launches[n] = cnt[k];
        ++n;
    }
    return n;
}

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
