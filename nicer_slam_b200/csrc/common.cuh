// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nicer_b200.h"

namespace nicer {

void set_error(const char *fmt, ...);
int num_sms();
int stream_fork(cudaStream_t main_stream, cudaStream_t side);

#define NICER_FAIL(code, ...)          \
    do {                               \
        nicer::set_error(__VA_ARGS__); \
        return (code);                 \
    } while (0)

#define NICER_CHECK_LAUNCH(name)                                                        \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) NICER_FAIL(-2, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

#define NICER_CUDA(call, name)                                                          \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) NICER_FAIL(-2, "%s: %s", name, cudaGetErrorString(e__)); \
    } while (0)

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace nicer
