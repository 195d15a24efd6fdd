// Optical-flow projection (model/network.py:153-165): the surface point of pixel p of frame i (rendered depth along its
// ray) is projected into frame j with K_j * (w2c_j * X) and the pixel position in i is subtracted.  One kernel per direction
// instead of ~12 matmul / permute / divide kernels forward and ~25 backward; gradients to the rendered depth, the rays, the
// camera centres and the world-to-camera matrices (hand-derived), K and uv are constants.
//   depth [B*n], dirs [B*n,3], loc [B,3], w2c [E,4,4] (of frame j of each edge), K [E,4,4], uv [B,n,2], idii [E] (int64)
//   flow [E,n,2]
#include "common.cuh"
#include "warp_math.cuh"

namespace nicer {

__global__ void __launch_bounds__(256)
flow_project_kernel(const float *__restrict__ depth, const float *__restrict__ dirs, const float *__restrict__ loc,
                    const float *__restrict__ w2c, const float *__restrict__ K, const float *__restrict__ uv,
                    const int64_t *__restrict__ idii, uint32_t E, uint32_t n, float *flow) {
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= E * n) return;
    const uint32_t e = t / n, p = t - e * n;
    const uint32_t i = (uint32_t)idii[e], ray = i * n + p;
    float X[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) X[a] = loc[3 * i + a] + depth[ray] * dirs[3 * (size_t)ray + a];
    const WarpProj pr = warp_project(X, w2c + 16 * e, K + 16 * e, 2.0f, 2.0f);
    flow[2 * (size_t)t] = pr.proj[0] / pr.zden - uv[2 * (size_t)ray];
    flow[2 * (size_t)t + 1] = pr.proj[1] / pr.zden - uv[2 * (size_t)ray + 1];
}

__global__ void __launch_bounds__(256)
flow_project_backward_kernel(const float *__restrict__ depth, const float *__restrict__ dirs, const float *__restrict__ loc,
                             const float *__restrict__ w2c, const float *__restrict__ K, const int64_t *__restrict__ idii, uint32_t E,
                             uint32_t n, const float *__restrict__ g_flow, float *g_depth, float *g_dirs, float *g_loc, float *g_w2c) {
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= E * n) return;
    const uint32_t e = t / n, p = t - e * n;
    const uint32_t i = (uint32_t)idii[e], ray = i * n + p;
    const float d = depth[ray];
    float dir[3], X[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { dir[a] = dirs[3 * (size_t)ray + a]; X[a] = loc[3 * i + a] + d * dir[a]; }
    const float *Wt = w2c + 16 * e, *Kt = K + 16 * e;
    const WarpProj pr = warp_project(X, Wt, Kt, 2.0f, 2.0f);
    const float gu = g_flow[2 * (size_t)t], gv = g_flow[2 * (size_t)t + 1];
    const float gproj[3] = {gu / pr.zden, gv / pr.zden, -(gu * pr.proj[0] + gv * pr.proj[1]) / (pr.zden * pr.zden)};
    float gcam[3], gX[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gcam[c] = Kt[c] * gproj[0] + Kt[4 + c] * gproj[1] + Kt[8 + c] * gproj[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) gX[c] = Wt[c] * gcam[0] + Wt[4 + c] * gcam[1] + Wt[8 + c] * gcam[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        atomicAdd(&g_w2c[16 * e + 4 * a + 0], gcam[a] * X[0]);
        atomicAdd(&g_w2c[16 * e + 4 * a + 1], gcam[a] * X[1]);
        atomicAdd(&g_w2c[16 * e + 4 * a + 2], gcam[a] * X[2]);
        atomicAdd(&g_w2c[16 * e + 4 * a + 3], gcam[a]);
        atomicAdd(&g_dirs[3 * (size_t)ray + a], d * gX[a]);
        atomicAdd(&g_loc[3 * i + a], gX[a]);
    }
    atomicAdd(&g_depth[ray], gX[0] * dir[0] + gX[1] * dir[1] + gX[2] * dir[2]);
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_flow_project(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K, const float *uv,
                                  const int64_t *idii, uint32_t E, uint32_t n, float *flow, void *stream) {
    if (E == 0 || n == 0) return 0;
    if (!depth || !dirs || !loc || !w2c || !K || !uv || !idii || !flow) NICER_FAIL(-1, "nicer_flow_project: NULL pointer");
    flow_project_kernel<<<div_up(E * n, 256), 256, 0, (cudaStream_t)stream>>>(depth, dirs, loc, w2c, K, uv, idii, E, n, flow);
    NICER_CHECK_LAUNCH("nicer_flow_project");
    return 0;
}

// g_depth [B*n], g_dirs [B*n,3], g_loc [B,3], g_w2c [E,4,4] must arrive zeroed (accumulated with atomics: a frame can be the
// source of several edges)
extern "C" int nicer_flow_project_backward(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                           const int64_t *idii, uint32_t E, uint32_t n, const float *g_flow, float *g_depth, float *g_dirs,
                                           float *g_loc, float *g_w2c, void *stream) {
    if (E == 0 || n == 0) return 0;
    if (!depth || !dirs || !loc || !w2c || !K || !idii || !g_flow || !g_depth || !g_dirs || !g_loc || !g_w2c)
        NICER_FAIL(-1, "nicer_flow_project_backward: NULL pointer");
    flow_project_backward_kernel<<<div_up(E * n, 256), 256, 0, (cudaStream_t)stream>>>(depth, dirs, loc, w2c, K, idii, E, n, g_flow, g_depth,
                                                                                     g_dirs, g_loc, g_w2c);
    NICER_CHECK_LAUNCH("nicer_flow_project_backward");
    return 0;
}
