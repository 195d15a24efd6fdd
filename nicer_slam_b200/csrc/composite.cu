// SDF -> density (voxel-counter beta, Laplace CDF) -> alpha compositing, one warp per ray with a
// warp-segmented scan for the transmittance, and its backward (reverse scan).
// Replaces GridPredefineDensity.forward (/root/reference/code/model/density.py:33-67),
// SLAMNetwork.volume_rendering (model/network.py:349-370) and the weighted sums
// (network.py:147-148, 338-341) plus ImportantSampler's weight computation (ray_sampler.py:105-112)
// and SLAMNetwork.update_voxels (network.py:62-76).
#include "common.cuh"
#include "composite_math.cuh"

namespace nicer {

constexpr int CMP_WARPS = 4;

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ float warp_incl_scan_rev(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += n;
    }
    return v;
}
// Exclusive prefix / suffix sums computed from the SHIFTED values (never as inclusive - self: the last sample of a
// ray has a free energy of ~1e10*sigma and subtracting it back would cancel the whole prefix in fp32).
__device__ __forceinline__ float warp_excl_scan(float v, int lane) {
    float prev = __shfl_up_sync(0xffffffffu, v, 1);
    return warp_incl_scan(lane == 0 ? 0.f : prev, lane);
}
__device__ __forceinline__ float warp_excl_scan_rev(float v, int lane) {
    float next = __shfl_down_sync(0xffffffffu, v, 1);
    return warp_incl_scan_rev(lane == 31 ? 0.f : next, lane);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool FULL>
__global__ void __launch_bounds__(CMP_WARPS * 32)
composite_forward_kernel(const float *__restrict__ sdf, const float *__restrict__ X, const float *__restrict__ Z,
                         const float *__restrict__ rgb, const float *__restrict__ grad,
                         const float *__restrict__ voxels, int res, uint32_t R, uint32_t S, float *weights,
                         float *rgb_out, float *depth_out, float *normal_out, float *wsum_out) {
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
    if (r >= R) return;
    float carry = 0.f;
    float a_rgb[3] = {0.f, 0.f, 0.f}, a_n[3] = {0.f, 0.f, 0.f}, a_w = 0.f, a_wz = 0.f;
    for (uint32_t base = 0; base < S; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < S;
        const size_t p = (size_t)r * S + (valid ? i : S - 1);
        const float s = sdf[p];
        const float beta = beta_lookup(voxels, res, X[3 * p], X[3 * p + 1], X[3 * p + 2]);
        const float sigma = laplace_density(s, beta);
        const float zi = Z[p];
        const float delta = (i + 1 < S) ? (Z[p + 1] - zi) : 1e10f;
        const float E = valid ? delta * sigma : 0.f;
        const float pre = warp_excl_scan(E, lane);
        const float excl = pre + carry;
        const float T = expf(-excl);
        const float alpha = 1.0f - expf(-E);
        const float w = valid ? alpha * T : 0.f;
        if (valid) weights[p] = w;
        if (FULL && valid) {
            a_w += w;
            a_wz += w * zi;
#pragma unroll
            for (int c = 0; c < 3; ++c) a_rgb[c] += w * rgb[3 * p + c];
            const float gx = grad[3 * p], gy = grad[3 * p + 1], gz = grad[3 * p + 2];
            const float den = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-6f;
            a_n[0] += w * (gx / den); a_n[1] += w * (gy / den); a_n[2] += w * (gz / den);
        }
        carry += __shfl_sync(0xffffffffu, pre + E, 31);
    }
    if (FULL) {
        a_w = warp_sum(a_w); a_wz = warp_sum(a_wz);
#pragma unroll
        for (int c = 0; c < 3; ++c) { a_rgb[c] = warp_sum(a_rgb[c]); a_n[c] = warp_sum(a_n[c]); }
        if (lane == 0) {
            wsum_out[r] = a_w;
            depth_out[r] = a_wz / (a_w + 1e-8f);
#pragma unroll
            for (int c = 0; c < 3; ++c) { rgb_out[3 * (size_t)r + c] = a_rgb[c]; normal_out[3 * (size_t)r + c] = a_n[c]; }
        }
    }
}

__global__ void __launch_bounds__(CMP_WARPS * 32)
composite_backward_kernel(const float *__restrict__ sdf, const float *__restrict__ X, const float *__restrict__ Z,
                          const float *__restrict__ rgb, const float *__restrict__ grad,
                          const float *__restrict__ voxels, int res, uint32_t R, uint32_t S,
                          const float *__restrict__ weights, const float *__restrict__ depth_out,
                          const float *__restrict__ wsum, const float *__restrict__ g_rgb_out,
                          const float *__restrict__ g_depth_out, const float *__restrict__ g_normal_out,
                          const float *__restrict__ g_weights, float *g_sdf, float *g_rgb, float *g_grad) {
    __shared__ float carries[CMP_WARPS][33];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * CMP_WARPS + wrp;
    if (r >= R) return;
    const uint32_t n_chunks = (S + 31) / 32;   // <= 32 (S <= 1024)
    // pass 1: free-energy prefix at the start of every chunk
    {
        float carry = 0.f;
        for (uint32_t ch = 0; ch < n_chunks; ++ch) {
            const uint32_t i = ch * 32 + lane;
            const bool valid = i < S;
            const size_t p = (size_t)r * S + (valid ? i : S - 1);
            const float beta = beta_lookup(voxels, res, X[3 * p], X[3 * p + 1], X[3 * p + 2]);
            const float sigma = laplace_density(sdf[p], beta);
            const float delta = (i + 1 < S) ? (Z[p + 1] - Z[p]) : 1e10f;
            const float E = valid ? delta * sigma : 0.f;
            if (lane == 0) carries[wrp][ch] = carry;
            carry += warp_sum(E);
        }
    }
    __syncwarp();
    float go_rgb[3] = {0.f, 0.f, 0.f}, go_n[3] = {0.f, 0.f, 0.f};
    if (g_rgb_out) { go_rgb[0] = g_rgb_out[3 * (size_t)r]; go_rgb[1] = g_rgb_out[3 * (size_t)r + 1]; go_rgb[2] = g_rgb_out[3 * (size_t)r + 2]; }
    if (g_normal_out) { go_n[0] = g_normal_out[3 * (size_t)r]; go_n[1] = g_normal_out[3 * (size_t)r + 1]; go_n[2] = g_normal_out[3 * (size_t)r + 2]; }
    const float go_d = g_depth_out ? g_depth_out[r] : 0.f;
    const float dep = depth_out[r];
    const float inv_ws = 1.0f / (wsum[r] + 1e-8f);
    float rcarry = 0.f;
    for (int ch = (int)n_chunks - 1; ch >= 0; --ch) {
        const uint32_t i = (uint32_t)ch * 32 + lane;
        const bool valid = i < S;
        const size_t p = (size_t)r * S + (valid ? i : S - 1);
        const float s = sdf[p];
        const float beta = beta_lookup(voxels, res, X[3 * p], X[3 * p + 1], X[3 * p + 2]);
        const float sigma = laplace_density(s, beta);
        const float zi = Z[p];
        const float delta = (i + 1 < S) ? (Z[p + 1] - zi) : 1e10f;
        const float E = valid ? delta * sigma : 0.f;
        const float excl = warp_excl_scan(E, lane) + carries[wrp][ch];
        const float T = expf(-excl);
        const float eE = expf(-E);
        const float w = valid ? weights[p] : 0.f;
        const float gvec[3] = {grad[3 * p], grad[3 * p + 1], grad[3 * p + 2]};
        const float cvec[3] = {rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]};
        float wbar = valid ? composite_wbar(go_rgb, go_n, go_d, cvec, gvec, zi, dep, inv_ws, g_weights ? g_weights[p] : 0.f) : 0.f;
        const float cbar = -wbar * w;                       // dL/d(prefix_i) = -wbar_i * alpha_i * T_i
        const float spre = warp_excl_scan_rev(cbar, lane);
        const float sexcl = spre + rcarry;                   // sum over samples after i
        if (valid) {
            float gs, gc[3], gg[3];
            composite_sample_backward(sexcl, wbar, T, eE, delta, s, beta, w, go_rgb, go_n, gvec, &gs, gc, gg);
            g_sdf[p] = gs;
#pragma unroll
            for (int c = 0; c < 3; ++c) { g_rgb[3 * p + c] = gc[c]; g_grad[3 * p + c] = gg[c]; }
        }
        rcarry += __shfl_sync(0xffffffffu, spre + cbar, 0);
    }
}

__global__ void voxel_count_kernel(const float *__restrict__ X, uint32_t P, float *voxels, int res) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float x = X[3 * (size_t)p], y = X[3 * (size_t)p + 1], z = X[3 * (size_t)p + 2];
    if (fabsf(x) > 0.99f || fabsf(y) > 0.99f || fabsf(z) > 0.99f) return;
    const int ix = (int)((x + 1.0f) / 2.0f * (float)res);
    const int iy = (int)((y + 1.0f) / 2.0f * (float)res);
    const int iz = (int)((z + 1.0f) / 2.0f * (float)res);
    atomicAdd(&voxels[((size_t)ix * res + iy) * res + iz], 1.0f);
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_composite_forward(const float *sdf, const float *x, const float *z, const float *rgb,
                                       const float *grad, const float *voxels, uint32_t voxel_res, uint32_t R,
                                       uint32_t S, float *weights, float *rgb_out, float *depth_out,
                                       float *normal_out, float *wsum, void *stream) {
    if (R == 0 || S == 0) return 0;
    if (S > 1024) NICER_FAIL(-1, "nicer_composite_forward: S must be <= 1024 (got %u)", S);
    if (!sdf || !x || !z || !rgb || !grad || !voxels || !weights || !rgb_out || !depth_out || !normal_out || !wsum)
        NICER_FAIL(-1, "nicer_composite_forward: NULL pointer");
    composite_forward_kernel<true><<<div_up(R, CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        sdf, x, z, rgb, grad, voxels, (int)voxel_res, R, S, weights, rgb_out, depth_out, normal_out, wsum);
    NICER_CHECK_LAUNCH("nicer_composite_forward");
    return 0;
}

extern "C" int nicer_sampler_weights(const float *sdf, const float *x, const float *z, const float *voxels,
                                     uint32_t voxel_res, uint32_t R, uint32_t S, float *weights, void *stream) {
    if (R == 0 || S == 0) return 0;
    if (S > 1024) NICER_FAIL(-1, "nicer_sampler_weights: S must be <= 1024 (got %u)", S);
    if (!sdf || !x || !z || !voxels || !weights) NICER_FAIL(-1, "nicer_sampler_weights: NULL pointer");
    composite_forward_kernel<false><<<div_up(R, CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        sdf, x, z, nullptr, nullptr, voxels, (int)voxel_res, R, S, weights, nullptr, nullptr, nullptr, nullptr);
    NICER_CHECK_LAUNCH("nicer_sampler_weights");
    return 0;
}

extern "C" int nicer_composite_backward(const float *sdf, const float *x, const float *z, const float *rgb,
                                        const float *grad, const float *voxels, uint32_t voxel_res, uint32_t R,
                                        uint32_t S, const float *weights, const float *depth_out, const float *wsum,
                                        const float *g_rgb_out, const float *g_depth_out, const float *g_normal_out,
                                        const float *g_weights, float *g_sdf, float *g_rgb, float *g_grad,
                                        void *stream) {
    if (R == 0 || S == 0) return 0;
    if (S > 1024) NICER_FAIL(-1, "nicer_composite_backward: S must be <= 1024 (got %u)", S);
    if (!sdf || !x || !z || !rgb || !grad || !voxels || !weights || !depth_out || !wsum || !g_sdf || !g_rgb || !g_grad)
        NICER_FAIL(-1, "nicer_composite_backward: NULL pointer");
    composite_backward_kernel<<<div_up(R, CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        sdf, x, z, rgb, grad, voxels, (int)voxel_res, R, S, weights, depth_out, wsum, g_rgb_out, g_depth_out,
        g_normal_out, g_weights, g_sdf, g_rgb, g_grad);
    NICER_CHECK_LAUNCH("nicer_composite_backward");
    return 0;
}

extern "C" int nicer_voxel_count(const float *x, uint32_t P, float *voxels, uint32_t voxel_res, void *stream) {
    if (P == 0) return 0;
    if (!x || !voxels) NICER_FAIL(-1, "nicer_voxel_count: NULL pointer");
    voxel_count_kernel<<<div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(x, P, voxels, (int)voxel_res);
    NICER_CHECK_LAUNCH("nicer_voxel_count");
    return 0;
}
