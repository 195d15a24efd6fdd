// Per-element math of the photometric warp block (/root/reference/code/model/network.py:167-279): a pixel of frame i,
// lifted with its rendered depth, is projected into frame t and the full-resolution colour image of t is sampled
// bilinearly there (F.grid_sample, align_corners=True, zeros padding).  Forward and hand-derived backward.
#pragma once
#include "nicer_math.cuh"

namespace nicer {

struct WarpProj {
    float cam[3];      // point in the camera frame of t
    float proj[3];     // K3 * cam
    float nu, nv;      // normalised sampling coordinates in [-1, 1]
    float zden;        // proj.z + 1e-8
};

// pts: world point; W2C: 4x4 row-major world-to-camera of the target frame; K: 4x4 intrinsics of the target frame
NHD WarpProj warp_project(const float pts[3], const float *W2C, const float *K, float Wimg, float Himg) {
    WarpProj r;
#pragma unroll
    for (int a = 0; a < 3; ++a) r.cam[a] = W2C[4 * a] * pts[0] + W2C[4 * a + 1] * pts[1] + W2C[4 * a + 2] * pts[2] + W2C[4 * a + 3];
#pragma unroll
    for (int a = 0; a < 3; ++a) r.proj[a] = K[4 * a] * r.cam[0] + K[4 * a + 1] * r.cam[1] + K[4 * a + 2] * r.cam[2];
    r.zden = r.proj[2] + 1e-8f;
    r.nu = r.proj[0] / r.zden / Wimg * 2.0f - 1.0f;
    r.nv = r.proj[1] / r.zden / Himg * 2.0f - 1.0f;
    return r;
}

// bilinear sample of img [H][W][3] (NHWC slice of one frame) at normalised (nu, nv), align_corners=True, zeros padding;
// optionally the derivative of sum_c g[c] * out[c] w.r.t. (nu, nv)
NHD void bilinear3(const float *img, int H, int W, float nu, float nv, float out[3], const float *g, float *dnu, float *dnv) {
    const float x = (nu + 1.0f) * 0.5f * (float)(W - 1), y = (nv + 1.0f) * 0.5f * (float)(H - 1);
    const float x0f = floorf(x), y0f = floorf(y);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = x - x0f, wx0 = 1.0f - wx1, wy1 = y - y0f, wy0 = 1.0f - wy1;
    out[0] = out[1] = out[2] = 0.f;
    float gx = 0.f, gy = 0.f;
    const int xs[2] = {x0, x1}, ys[2] = {y0, y1};
    const float wxs[2] = {wx0, wx1}, wys[2] = {wy0, wy1};
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (xs[i] < 0 || xs[i] >= W || ys[j] < 0 || ys[j] >= H) continue;
            const float *px = img + ((size_t)ys[j] * W + xs[i]) * 3;
            const float w = wxs[i] * wys[j];
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { out[c] += w * px[c]; if (g) dot += g[c] * px[c]; }
            if (g) {
                gx += (i ? 1.0f : -1.0f) * wys[j] * dot;
                gy += (j ? 1.0f : -1.0f) * wxs[i] * dot;
            }
        }
    if (g) { *dnu = gx * 0.5f * (float)(W - 1); *dnv = gy * 0.5f * (float)(H - 1); }
}

}  // namespace nicer
