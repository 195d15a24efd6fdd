// Per-sample density / compositing math shared by composite.cu and the host emulation (tests/host_emul).
// Follows /root/reference/code/model/density.py:37-60 and model/network.py:349-370.
#pragma once
#include "nicer_math.cuh"

namespace nicer {

NHD float beta_lookup(const float *voxels, int res, float x, float y, float z) {
    float count = 0.f;
    if (!(fabsf(x) > 0.99f || fabsf(y) > 0.99f || fabsf(z) > 0.99f)) {
        const int ix = (int)((x + 1.0f) / 2.0f * (float)res);
        const int iy = (int)((y + 1.0f) / 2.0f * (float)res);
        const int iz = (int)((z + 1.0f) / 2.0f * (float)res);
        count = voxels[((size_t)ix * res + iy) * res + iz];
    }
    // a*exp(-b*1e-4*count*d)+c with python-double constants rounded to fp32 per op, as torch does
    const float k1 = (float)(-0.0116544676 * 0.0001);
    float t = k1 * count;
    t = t * 5.37538f;
    return 0.01207724805f * expf(t) + 0.0023639156f;
}

NHD float sgn(float s) { return (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f); }

NHD float laplace_density(float s, float beta) {
    const float alpha = 1.0f / beta;
    return alpha * (0.5f + 0.5f * sgn(s) * expm1f(-fabsf(s) / beta));
}


// dL/dw_i from the upstream gradients of the per-ray outputs (rgb, normal map, depth, weights)
NHD float composite_wbar(const float go_rgb[3], const float go_n[3], float go_d, const float c[3], const float g[3],
                         float zi, float depth, float inv_ws, float gw_ext) {
    const float den = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) + 1e-6f;
    float wbar = go_rgb[0] * c[0] + go_rgb[1] * c[1] + go_rgb[2] * c[2];
    wbar += go_n[0] * (g[0] / den) + go_n[1] * (g[1] / den) + go_n[2] * (g[2] / den);
    wbar += go_d * (zi - depth) * inv_ws;
    return wbar + gw_ext;
}

// Given the suffix sum of cbar over later samples, finish one sample's backward.
NHD void composite_sample_backward(float suffix, float wbar, float T, float eE, float delta, float s, float beta,
                                   float w, const float go_rgb[3], const float go_n[3], const float g[3],
                                   float *g_sdf, float g_rgb[3], float g_grad[3]) {
    const float Ebar = suffix + wbar * T * eE;   // + dL/dalpha_i * dalpha/dE
    const float sigbar = Ebar * delta;
    // d sigma / d s = -(0.5/beta^2) * (expm1(-|s|/beta) + 1): torch differentiates expm1 as result+1, which is
    // exactly 0 once expm1 saturates to -1; 0 at s == 0, where sign()/abs() have zero gradient.
    const float dsig = (s == 0.f) ? 0.f : -((1.0f / beta) * 0.5f) * (expm1f(-fabsf(s) / beta) + 1.0f) / beta;
    *g_sdf = sigbar * dsig;
    const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float den = nrm + 1e-6f;
    const float nb[3] = {w * go_n[0], w * go_n[1], w * go_n[2]};
    const float dotg = nb[0] * g[0] + nb[1] * g[1] + nb[2] * g[2];
    const float k = (nrm > 0.f) ? dotg / (nrm * den * den) : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g_rgb[c] = w * go_rgb[c];
        g_grad[c] = nb[c] / den - g[c] * k;
    }
}

}  // namespace nicer
