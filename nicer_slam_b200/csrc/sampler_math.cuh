// Per-ray pieces of the hierarchical sampler shared by sampler.cu and the host emulation (tests/host_emul).
// Follows /root/reference/code/model/ray_sampler.py:21-61 (UniformSampler) and :90-166 (ImportantSampler).
#pragma once
#include "nicer_math.cuh"

namespace nicer {

// torch.linspace(start, end, steps)[i] in fp32: start + step*i in the lower half, end - step*(steps-1-i) in the upper
// half (aten/src/ATen/native/cuda/RangeFactories.cu), which is what the reference's t_vals / u are.
NHD float linspace_at(float start, float end, uint32_t steps, uint32_t i) {
    if (steps <= 1) return start;
    const float step = (end - start) / (float)(steps - 1);
    const uint32_t half = steps / 2;
    // plain expressions: on the device nvcc contracts them into FMAs exactly as it does in torch's own linspace kernel
    return (i < half) ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

// far end of a ray inside the cube [-bound, bound]^3 (near_far_from_cube, ray_sampler.py:21-34); 1e9 when the ray
// misses, then clamped to far_cap.
NHD float cube_far(const float o[3], const float d[3], float bound, float far_cap) {
    float near = -3.402823466e38f, far = 3.402823466e38f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float den = d[a] + 1e-15f;
        const float tmin = (-bound - o[a]) / den, tmax = (bound - o[a]) / den;
        const float lo = (tmin < tmax) ? tmin : tmax, hi = (tmin > tmax) ? tmin : tmax;
        near = fmaxf(near, lo);
        far = fminf(far, hi);
    }
    if (far < near) far = 1e9f;
    return fminf(far, far_cap);
}

// near*(1-t) + far*t with every product rounded on its own (torch runs mul, mul, add as separate kernels)
NHD float lerp_nf(float near, float far, float t) { return fmul_exact(near, 1.0f - t) + fmul_exact(far, t); }

// z of coarse sample i of N between near and far; with a stratified draw u in [0,1) when jitter (ray_sampler.py:47-58)
NHD float uniform_z(float near, float far, uint32_t N, uint32_t i, bool jitter, float u) {
    const float z = lerp_nf(near, far, linspace_at(0.f, 1.f, N, i));
    if (!jitter) return z;
    float lower = z, upper = z;
    if (i > 0) lower = 0.5f * (z + lerp_nf(near, far, linspace_at(0.f, 1.f, N, i - 1)));
    if (i + 1 < N) upper = 0.5f * (lerp_nf(near, far, linspace_at(0.f, 1.f, N, i + 1)) + z);
    return lower + fmul_exact(upper - lower, u);
}

// searchsorted(cdf, u, right=True): number of entries <= u (cdf is non-decreasing, n entries)
NHD uint32_t upper_bound(const float *cdf, uint32_t n, float u) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one inverse-CDF sample (ray_sampler.py:127-140)
NHD float invert_cdf(const float *cdf, const float *z, uint32_t n, float u) {
    const uint32_t ind = upper_bound(cdf, n, u);
    const uint32_t below = ind > 0 ? ind - 1 : 0, above = ind < n - 1 ? ind : n - 1;
    const float c0 = cdf[below], c1 = cdf[above];
    float denom = c1 - c0;
    if (denom < 1e-5f) denom = 1.0f;
    const float t = (u - c0) / denom;
    return z[below] + fmul_exact(t, z[above] - z[below]);
}

}  // namespace nicer
