// Photometric warp sampling (model/network.py:167-279) as one kernel per direction: every pixel of every frame i, lifted
// with its rendered depth, is projected into every frame t of the batch and the colour image of t is sampled bilinearly.
// The reference does this with ~35 elementwise / matmul / grid_sample kernels forward and ~50 backward per patch size.
//   sampled [B(t)][E][3], mask [B(t)][E] (uint8) with E = B * N * pp elements (frame i major)
// backward: d/d(depth [B*N]), d/d(dirs [E,3]), d/d(loc [B,3]), d/d(w2c [B,4,4])  (K and the images get no gradient)
#include "common.cuh"
#include "warp_math.cuh"

namespace nicer {

constexpr int WP_BLOCK = 256;

__global__ void __launch_bounds__(WP_BLOCK)
warp_sample_kernel(const float *__restrict__ depth, const float *__restrict__ dirs, const float *__restrict__ loc,
                   const float *__restrict__ w2c, const float *__restrict__ K, const float *__restrict__ img, uint32_t B, uint32_t N,
                   uint32_t pp, uint32_t H, uint32_t W, float *sampled, uint8_t *mask) {
    const uint32_t E = B * N * pp;
    const uint32_t e = blockIdx.x * WP_BLOCK + threadIdx.x;
    if (e >= E) return;
    const uint32_t i = e / (N * pp), ray = e / pp;
    const float d = depth[ray];
    float pts[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) pts[a] = loc[3 * i + a] + d * dirs[3 * (size_t)e + a];
    for (uint32_t t = 0; t < B; ++t) {
        const WarpProj pr = warp_project(pts, w2c + 16 * t, K + 16 * t, (float)W, (float)H);
        float out[3];
        bilinear3(img + (size_t)t * H * W * 3, (int)H, (int)W, pr.nu, pr.nv, out, nullptr, nullptr, nullptr);
        const size_t o = (size_t)t * E + e;
        sampled[3 * o] = out[0]; sampled[3 * o + 1] = out[1]; sampled[3 * o + 2] = out[2];
        mask[o] = (pr.nu > -1.0f && pr.nu < 1.0f && pr.nv > -1.0f && pr.nv < 1.0f && pr.proj[2] > 0.f) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(WP_BLOCK)
warp_sample_backward_kernel(const float *__restrict__ depth, const float *__restrict__ dirs, const float *__restrict__ loc,
                            const float *__restrict__ w2c, const float *__restrict__ K, const float *__restrict__ img, uint32_t B,
                            uint32_t N, uint32_t pp, uint32_t H, uint32_t W, const float *__restrict__ g_sampled, float *g_depth,
                            float *g_dirs, float *g_loc, float *g_w2c) {
    extern __shared__ float sh[];          // [B][12] w2c gradients, then [B][3] loc gradients
    float *sw = sh, *sl = sh + 12 * B;
    for (uint32_t k = threadIdx.x; k < 15 * B; k += WP_BLOCK) sh[k] = 0.f;
    __syncthreads();
    const uint32_t E = B * N * pp;
    const uint32_t e = blockIdx.x * WP_BLOCK + threadIdx.x;
    const bool live = e < E;
    const uint32_t ec = live ? e : E - 1;
    const uint32_t i = ec / (N * pp), ray = ec / pp;
    const int lane = threadIdx.x & 31;
    const float d = depth[ray];
    float dir[3], pts[3], gp[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 3; ++a) { dir[a] = dirs[3 * (size_t)ec + a]; pts[a] = loc[3 * i + a] + d * dir[a]; }
    // every lane of the warp walks all target frames: the 12 w2c-gradient terms of a frame are summed over the warp with
    // shuffles and added to shared memory once per warp (all 256 threads of a block adding to the same 12 words was the whole
    // cost of this kernel: 0.22 ms for 65 k elements)
    for (uint32_t t = 0; t < B; ++t) {
        const float *Wt = w2c + 16 * t, *Kt = K + 16 * t;
        const WarpProj pr = warp_project(pts, Wt, Kt, (float)W, (float)H);
        const size_t o = (size_t)t * E + ec;
        const float g[3] = {g_sampled[3 * o], g_sampled[3 * o + 1], g_sampled[3 * o + 2]};
        float gcam[3] = {0.f, 0.f, 0.f};
        const bool on = live && !(g[0] == 0.f && g[1] == 0.f && g[2] == 0.f);
        if (on) {
            float out[3], dnu, dnv;
            bilinear3(img + (size_t)t * H * W * 3, (int)H, (int)W, pr.nu, pr.nv, out, g, &dnu, &dnv);
            const float ax = dnu * 2.0f / (float)W, ay = dnv * 2.0f / (float)H;
            const float gproj[3] = {ax / pr.zden, ay / pr.zden, -(ax * pr.proj[0] + ay * pr.proj[1]) / (pr.zden * pr.zden)};
#pragma unroll
            for (int c = 0; c < 3; ++c) gcam[c] = Kt[c] * gproj[0] + Kt[4 + c] * gproj[1] + Kt[8 + c] * gproj[2];
#pragma unroll
            for (int c = 0; c < 3; ++c) gp[c] += Wt[c] * gcam[0] + Wt[4 + c] * gcam[1] + Wt[8 + c] * gcam[2];
        }
        if (__ballot_sync(0xffffffffu, on) == 0u) continue;          // warp-uniform
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = gcam[a] * (c < 3 ? pts[c] : 1.0f);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0 && v != 0.f) atomicAdd(&sw[12 * t + 4 * a + c], v);
            }
        }
    }
    // camera-centre gradient: one shared-memory add per warp when the whole warp belongs to one source frame
    const uint32_t i0 = __shfl_sync(0xffffffffu, i, 0);
    const bool same = __all_sync(0xffffffffu, i == i0);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = live ? gp[a] : 0.f;
        if (same) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) atomicAdd(&sl[3 * i0 + a], v);
        } else if (live) {
            atomicAdd(&sl[3 * i + a], v);
        }
    }
    if (live) {
#pragma unroll
        for (int a = 0; a < 3; ++a) g_dirs[3 * (size_t)e + a] = d * gp[a];
        const float gd = gp[0] * dir[0] + gp[1] * dir[1] + gp[2] * dir[2];
        if (pp == 1) g_depth[ray] = gd; else atomicAdd(&g_depth[ray], gd);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < 12 * B; k += WP_BLOCK)
        if (sw[k] != 0.f) atomicAdd(&g_w2c[16 * (k / 12) + (k % 12)], sw[k]);
    for (uint32_t k = threadIdx.x; k < 3 * B; k += WP_BLOCK)
        if (sl[k] != 0.f) atomicAdd(&g_loc[k], sl[k]);
}

// Ground-truth colour / depth of the patch pixels in their own frame (network.py:226-246): in-image test, integer pixel
// (truncated, clamped), 1 where the pixel is outside the image.  uvp [B*M,2], img [B,H,W,3], dep [B,H,W]
__global__ void warp_gt_kernel(const float *__restrict__ uvp, const float *__restrict__ img, const float *__restrict__ dep, uint32_t B,
                               uint32_t M, uint32_t H, uint32_t W, float *gt_rgb, float *gt_depth, uint8_t *inside) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * M) return;
    const uint32_t b = e / M;
    const float u = uvp[2 * (size_t)e], v = uvp[2 * (size_t)e + 1];
    const bool in = (0.f <= u) && (0.f <= v) && (u < (float)W) && (v < (float)H);
    float r[3] = {1.0f, 1.0f, 1.0f}, d = 1.0f;
    if (in) {
        const uint32_t ui = min((uint32_t)u, W - 1), vi = min((uint32_t)v, H - 1);
        const size_t px = ((size_t)b * H + vi) * W + ui;
        r[0] = img[3 * px]; r[1] = img[3 * px + 1]; r[2] = img[3 * px + 2];
        d = dep[px];
    }
    gt_rgb[3 * (size_t)e] = r[0]; gt_rgb[3 * (size_t)e + 1] = r[1]; gt_rgb[3 * (size_t)e + 2] = r[2];
    gt_depth[e] = d;
    inside[e] = in ? 1 : 0;
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_warp_sample(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                 const float *img, uint32_t B, uint32_t N, uint32_t pp, uint32_t H, uint32_t W, float *sampled,
                                 uint8_t *mask, void *stream) {
    if (B == 0 || N == 0 || pp == 0) return 0;
    if (!depth || !dirs || !loc || !w2c || !K || !img || !sampled || !mask) NICER_FAIL(-1, "nicer_warp_sample: NULL pointer");
    warp_sample_kernel<<<div_up(B * N * pp, WP_BLOCK), WP_BLOCK, 0, (cudaStream_t)stream>>>(depth, dirs, loc, w2c, K, img, B, N, pp, H, W,
                                                                                            sampled, mask);
    NICER_CHECK_LAUNCH("nicer_warp_sample");
    return 0;
}

// g_depth [B*N], g_loc [B,3], g_w2c [B,4,4] must arrive zeroed (accumulated with atomics); g_dirs [E,3] is written
extern "C" int nicer_warp_sample_backward(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                          const float *img, uint32_t B, uint32_t N, uint32_t pp, uint32_t H, uint32_t W,
                                          const float *g_sampled, float *g_depth, float *g_dirs, float *g_loc, float *g_w2c,
                                          void *stream) {
    if (B == 0 || N == 0 || pp == 0) return 0;
    if (!depth || !dirs || !loc || !w2c || !K || !img || !g_sampled || !g_depth || !g_dirs || !g_loc || !g_w2c)
        NICER_FAIL(-1, "nicer_warp_sample_backward: NULL pointer");
    if (B > 256) NICER_FAIL(-1, "nicer_warp_sample_backward: at most 256 frames");
    warp_sample_backward_kernel<<<div_up(B * N * pp, WP_BLOCK), WP_BLOCK, 15 * B * sizeof(float), (cudaStream_t)stream>>>(
        depth, dirs, loc, w2c, K, img, B, N, pp, H, W, g_sampled, g_depth, g_dirs, g_loc, g_w2c);
    NICER_CHECK_LAUNCH("nicer_warp_sample_backward");
    return 0;
}

extern "C" int nicer_warp_gt(const float *uvp, const float *img, const float *dep, uint32_t B, uint32_t M, uint32_t H, uint32_t W,
                             float *gt_rgb, float *gt_depth, uint8_t *inside, void *stream) {
    if (B == 0 || M == 0) return 0;
    if (!uvp || !img || !dep || !gt_rgb || !gt_depth || !inside) NICER_FAIL(-1, "nicer_warp_gt: NULL pointer");
    warp_gt_kernel<<<div_up(B * M, 256), 256, 0, (cudaStream_t)stream>>>(uvp, img, dep, B, M, H, W, gt_rgb, gt_depth, inside);
    NICER_CHECK_LAUNCH("nicer_warp_gt");
    return 0;
}
