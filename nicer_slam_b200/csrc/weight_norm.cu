// weight_norm of every Linear of a network in ONE launch per direction: W[r,:] = g[r] * v[r,:] / ||v[r,:]||
// (torch._weight_norm with dim = 0, which nn.utils.weight_norm's pre-forward hook calls once per layer and per forward:
// model/base_networks.py:149-153, 378-381).  One warp per row; the backward is torch's _weight_norm_interface_backward:
// dg[r] = <dW[r], v[r]> / n_r ,  dv[r,:] = (g[r] / n_r) * (dW[r,:] - v[r,:] * <dW[r], v[r]> / n_r^2).
#include "common.cuh"

namespace nicer {

constexpr int WN_WARPS = 8;
struct WnJobs {
    nicer_wn_job_t j[NICER_WN_MAX_JOBS];
    uint32_t row0[NICER_WN_MAX_JOBS + 1];
    uint32_t n;
};

__device__ __forceinline__ float wn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(WN_WARPS * 32) weight_norm_kernel(const WnJobs jobs) {
    const int lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * WN_WARPS + (threadIdx.x >> 5);
    if (row >= jobs.row0[jobs.n]) return;
    uint32_t k = 0;
    while (row >= jobs.row0[k + 1]) ++k;
    const nicer_wn_job_t &J = jobs.j[k];
    const uint32_t r = row - jobs.row0[k], C = J.cols;
    const float *v = J.v + (size_t)r * C;
    if (!BACKWARD) {
        float ss = 0.f;
        for (uint32_t c = lane; c < C; c += 32) ss += v[c] * v[c];
        const float nrm = sqrtf(wn_warp_sum(ss));
        const float s = J.g[r] / nrm;
        float *w = J.w + (size_t)r * C;
        for (uint32_t c = lane; c < C; c += 32) w[c] = v[c] * s;
        if (lane == 0 && J.norm) J.norm[r] = nrm;
    } else {
        const float *dw = J.dw + (size_t)r * C;
        float dot = 0.f;
        for (uint32_t c = lane; c < C; c += 32) dot += dw[c] * v[c];
        dot = wn_warp_sum(dot);
        const float nrm = J.norm[r], g = J.g[r];
        const float a = g / nrm, b = dot / (nrm * nrm);
        float *dv = J.dv + (size_t)r * C;
        for (uint32_t c = lane; c < C; c += 32) dv[c] = a * (dw[c] - v[c] * b);
        if (lane == 0) J.dg[r] = dot / nrm;
    }
}

template <bool BACKWARD>
static int launch(const nicer_wn_job_t *jobs, uint32_t n, cudaStream_t st, const char *name) {
    if (n == 0) return 0;
    if (!jobs) NICER_FAIL(-1, "%s: NULL jobs", name);
    if (n > NICER_WN_MAX_JOBS) NICER_FAIL(-1, "%s: at most %d jobs per call", name, NICER_WN_MAX_JOBS);
    WnJobs J;
    J.n = n;
    J.row0[0] = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const nicer_wn_job_t &j = jobs[k];
        if (!j.v || !j.g || (!BACKWARD && !j.w) || (BACKWARD && (!j.dw || !j.dv || !j.dg || !j.norm)))
            NICER_FAIL(-1, "%s: NULL pointer in job %u", name, k);
        J.j[k] = j;
        J.row0[k + 1] = J.row0[k] + j.rows;
    }
    if (J.row0[n] == 0) return 0;
    weight_norm_kernel<BACKWARD><<<div_up(J.row0[n], WN_WARPS), WN_WARPS * 32, 0, st>>>(J);
    NICER_CHECK_LAUNCH(name);
    return 0;
}

}  // namespace nicer

extern "C" int nicer_weight_norm(const nicer_wn_job_t *jobs, uint32_t n, void *stream) {
    return nicer::launch<false>(jobs, n, (cudaStream_t)stream, "nicer_weight_norm");
}
extern "C" int nicer_weight_norm_backward(const nicer_wn_job_t *jobs, uint32_t n, void *stream) {
    return nicer::launch<true>(jobs, n, (cudaStream_t)stream, "nicer_weight_norm_backward");
}
