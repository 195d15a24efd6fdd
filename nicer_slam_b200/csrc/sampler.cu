// Hierarchical ray sampler (model/ray_sampler.py:21-61, 90-166) in two kernels around the no-grad SDF pass:
//   nicer_sampler_uniform   far end of every ray in the cube, N_eval coarse depths (stratified when training) and their
//                           3-D points, one thread per (ray, sample);
//   nicer_sampler_resample  one WARP per ray: density -> transmittance weights (warp scan, as compositing) -> pdf -> cdf
//                           (warp scan) -> inverse-CDF samples (binary search in shared memory) -> merge with near / far /
//                           the randomly picked coarse depths -> sort (bitonic, shared memory) -> eikonal depth pick.
// The reference runs this tail as ~36 elementwise / cumsum / searchsorted / gather / sort kernels over [R, 640] tensors.
#include "common.cuh"
#include "composite_math.cuh"
#include "sampler_math.cuh"

namespace nicer {

constexpr int SMP_WARPS = 4;
constexpr int SMP_MAX_U = 1024;      // coarse samples per ray
constexpr int SMP_SORT = 256;        // slots of the per-warp sort (N + 2 + n_extra <= 256)

__global__ void __launch_bounds__(256)
sampler_uniform_kernel(const float *__restrict__ cam_loc, const float *__restrict__ ray_dirs, float near, float far_cap, float bound,
                       int use_cube, const float *__restrict__ rnd, uint32_t R, uint32_t N, float *z, float *far_out, float *points) {
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= (size_t)R * N) return;
    const uint32_t r = (uint32_t)(e / N), i = (uint32_t)(e - (size_t)r * N);
    const float o[3] = {cam_loc[3 * (size_t)r], cam_loc[3 * (size_t)r + 1], cam_loc[3 * (size_t)r + 2]};
    const float d[3] = {ray_dirs[3 * (size_t)r], ray_dirs[3 * (size_t)r + 1], ray_dirs[3 * (size_t)r + 2]};
    const float far = use_cube ? cube_far(o, d, bound, far_cap) : far_cap;
    const float zi = uniform_z(near, far, N, i, rnd != nullptr, rnd ? rnd[e] : 0.f);
    z[e] = zi;
    if (i == 0 && far_out) far_out[r] = far;
    if (points) {
#pragma unroll
        for (int a = 0; a < 3; ++a) points[3 * e + a] = o[a] + fmul_exact(zi, d[a]);
    }
}

__device__ __forceinline__ float w_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ float w_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(SMP_WARPS * 32)
sampler_resample_kernel(const float *__restrict__ sdf, const float *__restrict__ X, const float *__restrict__ Z,
                        const float *__restrict__ voxels, int res, uint32_t R, uint32_t U, uint32_t N,
                        const int64_t *__restrict__ sel, uint32_t n_extra, float near, const float *__restrict__ far,
                        const int64_t *__restrict__ eik_idx, float *z_out, float *z_eik, float *weights_out) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    float *cdf = sm + (size_t)wrp * (2 * SMP_MAX_U + SMP_SORT);      // [U] (first the weights, then the cdf)
    float *zs = cdf + SMP_MAX_U;                                    // [U] coarse depths of the ray
    float *srt = zs + SMP_MAX_U;                                    // [SMP_SORT]
    const uint32_t r = blockIdx.x * SMP_WARPS + wrp;
    if (r >= R) return;
    // ---- weights (ray_sampler.py:105-112), exactly as the compositing kernel computes them
    float carry = 0.f;
    for (uint32_t base = 0; base < U; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < U;
        const size_t p = (size_t)r * U + (valid ? i : U - 1);
        const float beta = beta_lookup(voxels, res, X[3 * p], X[3 * p + 1], X[3 * p + 2]);
        const float sigma = laplace_density(sdf[p], beta);
        const float zi = Z[p];
        const float delta = (i + 1 < U) ? (Z[p + 1] - zi) : 1e10f;
        const float E = valid ? delta * sigma : 0.f;
        const float prev = __shfl_up_sync(0xffffffffu, E, 1);
        const float pre = w_incl_scan(lane == 0 ? 0.f : prev, lane);
        const float T = expf(-(pre + carry));
        const float w = (1.0f - expf(-E)) * T;
        if (valid) { cdf[i] = w; zs[i] = zi; if (weights_out) weights_out[p] = w; }
        carry += __shfl_sync(0xffffffffu, pre + E, 31);
    }
    __syncwarp();
    // ---- pdf = (w[:-1] + 1e-5) / sum ; cdf = [0, cumsum(pdf)]
    const uint32_t M = U - 1;
    float tot = 0.f;
    for (uint32_t i = lane; i < M; i += 32) tot += cdf[i] + 1e-5f;
    tot = w_sum(tot);
    float run = 0.f;
    for (uint32_t base = 0; base < M; base += 32) {
        const uint32_t i = base + lane;
        const float pdf = (i < M) ? (cdf[i] + 1e-5f) / tot : 0.f;
        const float inc = w_incl_scan(pdf, lane) + run;
        __syncwarp();
        if (i < M) cdf[i] = inc;                 // temporarily: inclusive sum at i (shifted below)
        run = __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
    // shift right by one: process from the top so that nothing is overwritten before it is read
    for (int base = (int)((M + 31) / 32) * 32 - 32; base >= 0; base -= 32) {
        const uint32_t i = (uint32_t)base + lane;
        const float v = (i < M) ? cdf[i] : 0.f;
        __syncwarp();
        if (i < M) cdf[i + 1] = v;
        __syncwarp();
    }
    // The pdf sums to one, so the last cdf entry is 1 up to the rounding of the scan.  searchsorted(right=True) at u = 1
    // lands on the last bin only if that entry is <= 1; when rounding pushes it above, the reference falls into its
    // denom < 1e-5 branch and returns the previous coarse depth instead (a coin flip that depends on torch's reduction order).
    // Clamp, i.e. take the outcome of exact arithmetic.
    if (lane == 0) { cdf[0] = 0.f; cdf[M] = fminf(cdf[M], 1.0f); }
    __syncwarp();
    // ---- inverse CDF at u = linspace(0, 1, N), extras, sort
    const uint32_t S = N + 2 + n_extra;
    const uint32_t SN = (S <= 64) ? 64u : ((S <= 128) ? 128u : (uint32_t)SMP_SORT);    // sort width (power of two >= S)
    for (uint32_t k = lane; k < SN; k += 32) {
        float v = 3.402823466e38f;
        if (k < N) v = invert_cdf(cdf, zs, U, linspace_at(0.f, 1.f, N, k));
        else if (k == N) v = near;
        else if (k == N + 1) v = far[r];
        else if (k < S) v = zs[(uint32_t)sel[k - N - 2]];
        srt[k] = v;
    }
    __syncwarp();
    for (uint32_t size = 2; size <= SN; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < SN / 2; t += 32) {
                const uint32_t lo = 2 * t - (t & (stride - 1));      // index with bit `stride` cleared
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const float a = srt[lo], b = srt[hi];
                if ((a > b) == up) { srt[lo] = b; srt[hi] = a; }
            }
            __syncwarp();
        }
    }
    for (uint32_t k = lane; k < S; k += 32) z_out[(size_t)r * S + k] = srt[k];
    if (lane == 0 && z_eik) z_eik[r] = srt[(uint32_t)eik_idx[r]];
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_sampler_uniform(const float *cam_loc, const float *ray_dirs, float near, float far_cap, float bound, int use_cube,
                                     const float *rnd, uint32_t R, uint32_t N, float *z, float *far, float *points, void *stream) {
    if (R == 0 || N == 0) return 0;
    if (!cam_loc || !ray_dirs || !z) NICER_FAIL(-1, "nicer_sampler_uniform: NULL pointer");
    const size_t total = (size_t)R * N;
    sampler_uniform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cam_loc, ray_dirs, near, far_cap, bound,
                                                                                          use_cube, rnd, R, N, z, far, points);
    NICER_CHECK_LAUNCH("nicer_sampler_uniform");
    return 0;
}

extern "C" int nicer_sampler_resample(const float *sdf, const float *x, const float *z, const float *voxels, uint32_t voxel_res,
                                      uint32_t R, uint32_t U, uint32_t N, const int64_t *sel, uint32_t n_extra, float near,
                                      const float *far, const int64_t *eik_idx, float *z_out, float *z_eik, float *weights,
                                      void *stream) {
    if (R == 0) return 0;
    if (!sdf || !x || !z || !voxels || !far || !z_out || (n_extra && !sel) || (z_eik && !eik_idx))
        NICER_FAIL(-1, "nicer_sampler_resample: NULL pointer");
    if (U < 2 || U > SMP_MAX_U) NICER_FAIL(-1, "nicer_sampler_resample: 2 <= N_samples_eval <= %d (got %u)", SMP_MAX_U, U);
    if (N + 2 + n_extra > SMP_SORT) NICER_FAIL(-1, "nicer_sampler_resample: N_samples + 2 + N_samples_extra <= %d", SMP_SORT);
    const size_t smem = (size_t)SMP_WARPS * (2 * SMP_MAX_U + SMP_SORT) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        NICER_CUDA(cudaFuncSetAttribute(sampler_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "nicer_sampler_resample");
        attr = true;
    }
    sampler_resample_kernel<<<div_up(R, SMP_WARPS), SMP_WARPS * 32, smem, (cudaStream_t)stream>>>(
        sdf, x, z, voxels, (int)voxel_res, R, U, N, sel, n_extra, near, far, eik_idx, z_out, z_eik, weights);
    NICER_CHECK_LAUNCH("nicer_sampler_resample");
    return 0;
}
