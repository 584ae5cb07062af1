// common.cuh -- shared helpers of libsnb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/snb200.h"

namespace snb {

void set_error(const char *fmt, ...);

#define SNB_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            ::snb::set_error(__VA_ARGS__);    \
            return (code);                    \
        }                                     \
    } while (0)

#define SNB_LAUNCH_CHECK(name)                                                      \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ::snb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return SNB_ERR_LAUNCH;                                                  \
        }                                                                           \
    } while (0)

static inline cudaStream_t S(snb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int kNumSMs = 148;  // B200

}  // namespace snb
