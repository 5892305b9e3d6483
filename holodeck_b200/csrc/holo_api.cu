// holo_api.cu -- version / error reporting of libholo_b200 (see include/holo_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "holo_api.cuh"

namespace holo {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static bool g_profiling = false;
static double g_profile[8];
static int g_nprofile = 0;
bool profiling_on() { return g_profiling; }
void store_profile(const double* ms, int n) {
    g_nprofile = n < 8 ? n : 8;
    for (int i = 0; i < g_nprofile; ++i) g_profile[i] = ms[i];
}
}  // namespace holo

extern "C" {

int64_t holo_launch_count(void) { return (int64_t)holo::g_launches.load(); }

void holo_set_profiling(int on) { holo::g_profiling = (on != 0); }

int holo_get_profile(double* ms, int n) {
    int m = holo::g_nprofile < n ? holo::g_nprofile : n;
    for (int i = 0; i < m; ++i) ms[i] = holo::g_profile[i];
    return m;
}

int holo_abi_version(void) { return HOLO_ABI_VERSION; }

const char* holo_last_error(void) { return holo::g_err; }

int holo_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        holo::set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return 0;
    }
    return n;
}

int holo_check_launch(const char* who) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        holo::set_error("%s: kernel launch failed: %s", who, cudaGetErrorString(e));
        return HOLO_ERR_CUDA;
    }
    return HOLO_OK;
}

}  // extern "C"
