// holo_api.cuh -- error plumbing shared by the translation units of libholo_b200.
#pragma once
#include <cuda_runtime.h>

#include "../../include/holo_b200.h"

namespace holo {
void set_error(const char* fmt, ...);
void count_launches(int n);
bool profiling_on();
void store_profile(const double* ms, int n);

// Records up to 8 events on a stream; `finish()` synchronises and stores the gaps as the profile.
struct StageTimer {
    cudaEvent_t ev[8];
    int n = 0;
    bool on;
    cudaStream_t st;
    explicit StageTimer(cudaStream_t s) : on(profiling_on()), st(s) {}
    void mark() {
        if (!on || n >= 8) return;
        cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        ++n;
    }
    void finish() {
        if (!on || n < 2) return;
        cudaEventSynchronize(ev[n - 1]);
        double ms[8];
        for (int i = 0; i + 1 < n; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
            ms[i] = t;
        }
        store_profile(ms, n - 1);
        for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
        n = 0;
    }
};
}

extern "C" int holo_check_launch(const char* who);

namespace holo {
// holo_realize.cu: realised GWB of a column slab of a wider grid (used by the eccentric harmonic sum)
int realize_gwb_columns(const double* number, const double* h2fdf, int64_t ncell, int F, int R, int64_t r0,
                        uint64_t seed, double normal_threshold, const double* counts, int key_col0, int key_cols,
                        double* gwb, void* workspace, int64_t workspace_bytes, void* stream);
}

#define HOLO_REQUIRE(cond, msg)                      \
    do {                                             \
        if (!(cond)) {                               \
            holo::set_error("%s", (msg));            \
            return HOLO_ERR_ARG;                     \
        }                                            \
    } while (0)

#define HOLO_CUDA(call)                                                                    \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            holo::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                     \
            return HOLO_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)
