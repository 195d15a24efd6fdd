// Camera / ray helpers of the hot path as single kernels (math in geometry_math.cuh):
//   nicer_pose_from_cam7(_backward)   <- get_camera_from_tensor / quad2rotation   utils/general.py:52-100
//   nicer_camera_rays(_backward)      <- get_camera_params / lift                 utils/rend_util.py:68-93,107-129
//   nicer_inv4x4(_backward)           <- torch.inverse(pose) of the flow / warp blocks   model/network.py:157,171
//   nicer_ray_points(_backward)       <- points = cam_loc + z * dir and the per-sample view directions
//                                        model/network.py:112-117 (and their sum-over-samples backward)
// All tensors are tiny next to the network kernels; the point is the launch count (the reference issues ~200 elementwise
// kernels for these per iteration, forward + backward).
#include "common.cuh"
#include "geometry_math.cuh"

namespace nicer {

__global__ void pose_from_cam7_kernel(const float *__restrict__ cam7, uint32_t B, float *pose) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) pose_from_cam7(cam7 + 7 * (size_t)b, pose + 16 * (size_t)b);
}
__global__ void pose_from_cam7_backward_kernel(const float *__restrict__ cam7, const float *__restrict__ g_pose, uint32_t B, float *g_cam7) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) pose_from_cam7_backward(cam7 + 7 * (size_t)b, g_pose + 16 * (size_t)b, g_cam7 + 7 * (size_t)b);
}

__global__ void inv4x4_kernel(const float *__restrict__ A, uint32_t B, float *Ai) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) inv4x4(A + 16 * (size_t)b, Ai + 16 * (size_t)b);
}
__global__ void inv4x4_backward_kernel(const float *__restrict__ Ai, const float *__restrict__ G, uint32_t B, float *GA) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) inv4x4_backward(Ai + 16 * (size_t)b, G + 16 * (size_t)b, GA + 16 * (size_t)b);
}

__global__ void camera_rays_kernel(const float *__restrict__ uv, const float *__restrict__ pose, const float *__restrict__ K, uint32_t B,
                                   uint32_t N, float *dirs, float *cam_loc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const uint32_t b = i / N;
    const float *P = pose + 16 * (size_t)b;
    float d[3], v[3], p[3];
    camera_ray(P, K + 16 * (size_t)b, uv[2 * (size_t)i], uv[2 * (size_t)i + 1], d, v, p);
    dirs[3 * (size_t)i] = d[0]; dirs[3 * (size_t)i + 1] = d[1]; dirs[3 * (size_t)i + 2] = d[2];
    if (i - b * N == 0) { cam_loc[3 * b] = P[3]; cam_loc[3 * b + 1] = P[7]; cam_loc[3 * b + 2] = P[11]; }
}

// one block per frame: g_R = sum_n g_v (x) p_n ; the translation cancels in v = (R p + t) - t, so g_t = g_loc
constexpr int CR_BLOCK = 256;
__global__ void __launch_bounds__(CR_BLOCK)
camera_rays_backward_kernel(const float *__restrict__ uv, const float *__restrict__ pose, const float *__restrict__ K, uint32_t N,
                            const float *__restrict__ g_dirs, const float *__restrict__ g_loc, float *g_pose) {
    const uint32_t b = blockIdx.x;
    const float *P = pose + 16 * (size_t)b, *Kb = K + 16 * (size_t)b;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    for (uint32_t n = threadIdx.x; n < N; n += CR_BLOCK) {
        const size_t i = (size_t)b * N + n;
        float d[3], v[3], p[3], gv[3];
        camera_ray(P, Kb, uv[2 * i], uv[2 * i + 1], d, v, p);
        const float gd[3] = {g_dirs[3 * i], g_dirs[3 * i + 1], g_dirs[3 * i + 2]};
        camera_ray_backward(v, gd, gv);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[3 * r + c] += gv[r] * p[c];
    }
    __shared__ float red[CR_BLOCK / 32][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        float a = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = a;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
        float out = 0.f;
        if (r < 3 && c < 3) {
            for (int w = 0; w < CR_BLOCK / 32; ++w) out += red[w][3 * r + c];
        } else if (r < 3 && g_loc) {
            out = g_loc[3 * b + r];
        }
        g_pose[16 * (size_t)b + threadIdx.x] = out;
    }
}

__global__ void ray_points_kernel(const float *__restrict__ loc, const float *__restrict__ dirs, const float *__restrict__ z, uint32_t R,
                                  uint32_t S, float *points, float *dirs_flat) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * S) return;
    const uint32_t r = i / S;
    const float zz = z[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d = dirs[3 * (size_t)r + k];
        points[3 * (size_t)i + k] = loc[3 * (size_t)r + k] + fmul_exact(zz, d);     // o + (z * d): two roundings, as torch does
        if (dirs_flat) dirs_flat[3 * (size_t)i + k] = d;
    }
}

// warp per ray: g_loc = sum_s g_p, g_dir = sum_s (z g_p + g_dirs_flat)
__global__ void ray_points_backward_kernel(const float *__restrict__ z, uint32_t R, uint32_t S, const float *__restrict__ g_points,
                                           const float *__restrict__ g_dirs_flat, float *g_loc, float *g_dirs) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint32_t s = lane; s < S; s += 32) {
        const size_t i = (size_t)r * S + s;
        const float zz = z[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float gp = g_points ? g_points[3 * i + k] : 0.f;
            a[k] += gp;
            a[3 + k] += zz * gp + (g_dirs_flat ? g_dirs_flat[3 * i + k] : 0.f);
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { g_loc[3 * (size_t)r + k] = a[k]; g_dirs[3 * (size_t)r + k] = a[3 + k]; }
    }
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_pose_from_cam7(const float *cam7, uint32_t B, float *pose, void *stream) {
    if (B == 0) return 0;
    if (!cam7 || !pose) NICER_FAIL(-1, "nicer_pose_from_cam7: NULL pointer");
    pose_from_cam7_kernel<<<div_up(B, 64), 64, 0, (cudaStream_t)stream>>>(cam7, B, pose);
    NICER_CHECK_LAUNCH("nicer_pose_from_cam7");
    return 0;
}

extern "C" int nicer_pose_from_cam7_backward(const float *cam7, const float *g_pose, uint32_t B, float *g_cam7, void *stream) {
    if (B == 0) return 0;
    if (!cam7 || !g_pose || !g_cam7) NICER_FAIL(-1, "nicer_pose_from_cam7_backward: NULL pointer");
    pose_from_cam7_backward_kernel<<<div_up(B, 64), 64, 0, (cudaStream_t)stream>>>(cam7, g_pose, B, g_cam7);
    NICER_CHECK_LAUNCH("nicer_pose_from_cam7_backward");
    return 0;
}

extern "C" int nicer_inv4x4(const float *A, uint32_t B, float *Ai, void *stream) {
    if (B == 0) return 0;
    if (!A || !Ai) NICER_FAIL(-1, "nicer_inv4x4: NULL pointer");
    inv4x4_kernel<<<div_up(B, 32), 32, 0, (cudaStream_t)stream>>>(A, B, Ai);
    NICER_CHECK_LAUNCH("nicer_inv4x4");
    return 0;
}

extern "C" int nicer_inv4x4_backward(const float *Ai, const float *G, uint32_t B, float *GA, void *stream) {
    if (B == 0) return 0;
    if (!Ai || !G || !GA) NICER_FAIL(-1, "nicer_inv4x4_backward: NULL pointer");
    inv4x4_backward_kernel<<<div_up(B, 32), 32, 0, (cudaStream_t)stream>>>(Ai, G, B, GA);
    NICER_CHECK_LAUNCH("nicer_inv4x4_backward");
    return 0;
}

extern "C" int nicer_camera_rays(const float *uv, const float *pose, const float *K, uint32_t B, uint32_t N, float *dirs,
                                 float *cam_loc, void *stream) {
    if (B == 0 || N == 0) return 0;
    if (!uv || !pose || !K || !dirs || !cam_loc) NICER_FAIL(-1, "nicer_camera_rays: NULL pointer");
    camera_rays_kernel<<<div_up(B * N, 256), 256, 0, (cudaStream_t)stream>>>(uv, pose, K, B, N, dirs, cam_loc);
    NICER_CHECK_LAUNCH("nicer_camera_rays");
    return 0;
}

extern "C" int nicer_camera_rays_backward(const float *uv, const float *pose, const float *K, uint32_t B, uint32_t N,
                                          const float *g_dirs, const float *g_loc, float *g_pose, void *stream) {
    if (B == 0) return 0;
    if (!uv || !pose || !K || !g_dirs || !g_pose) NICER_FAIL(-1, "nicer_camera_rays_backward: NULL pointer");
    camera_rays_backward_kernel<<<B, CR_BLOCK, 0, (cudaStream_t)stream>>>(uv, pose, K, N, g_dirs, g_loc, g_pose);
    NICER_CHECK_LAUNCH("nicer_camera_rays_backward");
    return 0;
}

extern "C" int nicer_ray_points(const float *cam_loc, const float *dirs, const float *z, uint32_t R, uint32_t S, float *points,
                                float *dirs_flat, void *stream) {
    if (R == 0 || S == 0) return 0;
    if (!cam_loc || !dirs || !z || !points) NICER_FAIL(-1, "nicer_ray_points: NULL pointer");
    ray_points_kernel<<<div_up(R * S, 256), 256, 0, (cudaStream_t)stream>>>(cam_loc, dirs, z, R, S, points, dirs_flat);
    NICER_CHECK_LAUNCH("nicer_ray_points");
    return 0;
}

extern "C" int nicer_ray_points_backward(const float *z, uint32_t R, uint32_t S, const float *g_points, const float *g_dirs_flat,
                                         float *g_loc, float *g_dirs, void *stream) {
    if (R == 0) return 0;
    if (!z || !g_loc || !g_dirs) NICER_FAIL(-1, "nicer_ray_points_backward: NULL pointer");
    ray_points_backward_kernel<<<div_up(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(z, R, S, g_points, g_dirs_flat, g_loc, g_dirs);
    NICER_CHECK_LAUNCH("nicer_ray_points_backward");
    return 0;
}
