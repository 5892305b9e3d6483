// holo_api.cuh -- error plumbing shared by the translation units of libholo_b200.
#pragma once
#include <cuda_runtime.h>

#include "../../include/holo_b200.h"

namespace holo {
void set_error(const char* fmt, ...);
void count_launches(int n);
}

extern "C" int holo_check_launch(const char* who);

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
