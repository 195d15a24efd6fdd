// Error reporting and device queries of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace nicer {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Makes `side` wait for everything enqueued on `main` so far (fork of a stream-level dependency; also legal during
// CUDA-graph capture of `main`).  Events come from a small round-robin pool created on first use.
int stream_fork(cudaStream_t main_stream, cudaStream_t side) {
    constexpr int POOL = 64;
    static cudaEvent_t pool[POOL];
    static int next = -1;
    if (next < 0) {
        for (int i = 0; i < POOL; ++i)
            if (cudaEventCreateWithFlags(&pool[i], cudaEventDisableTiming) != cudaSuccess) NICER_FAIL(-2, "stream_fork: cudaEventCreate failed");
        next = 0;
    }
    cudaEvent_t ev = pool[next];
    next = (next + 1) % POOL;
    NICER_CUDA(cudaEventRecord(ev, main_stream), "stream_fork");
    NICER_CUDA(cudaStreamWaitEvent(side, ev, 0), "stream_fork");
    return 0;
}

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;  // B200
    }
    return cached;
}

}  // namespace nicer

extern "C" const char *nicer_last_error(void) { return nicer::g_err; }
extern "C" int nicer_version(void) { return 1; }
