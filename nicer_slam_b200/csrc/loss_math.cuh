// Per-element math of SLAMLoss (/root/reference/code/model/loss.py:113-233) and the scale-and-shift-invariant depth loss
// (/root/reference/code/utils/MiDaS.py:6-140): values and hand-derived gradients, shared by the CUDA kernels (loss.cu)
// and the host emulation used by the CPU tests.
#pragma once
#include "nicer_math.cuh"

namespace nicer {

NHD float sgnf(float v) { return v > 0.f ? 1.0f : (v < 0.f ? -1.0f : 0.f); }   // d|v|/dv as torch defines it (0 at 0)

// F.normalize(v, p=2, dim=-1): v / max(|v|, 1e-12)
NHD float normalize3(const float v[3], float out[3]) {
    const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const float d = n > 1e-12f ? n : 1e-12f;
    out[0] = v[0] / d; out[1] = v[1] / d; out[2] = v[2] / d;
    return d;
}

// Mono-normal terms of one ray (loss.py:94-100): pred/gt are multiplied by the mask m (0/1) BEFORE normalisation.
// l1 = sum_c |n_c - g_c|, cs = 1 - n.g;  grad = d(w_l1 * l1 + w_cos * cs)/d(pred)  (times m; 0 for masked rays)
NHD void normal_terms(const float pred[3], const float gt[3], float m, float w_l1, float w_cos, float *l1, float *cs, float grad[3]) {
    const float v[3] = {pred[0] * m, pred[1] * m, pred[2] * m}, g0[3] = {gt[0] * m, gt[1] * m, gt[2] * m};
    float n[3], g[3];
    const float dn = normalize3(v, n);
    normalize3(g0, g);
    *l1 = fabsf(n[0] - g[0]) + fabsf(n[1] - g[1]) + fabsf(n[2] - g[2]);
    *cs = 1.0f - (n[0] * g[0] + n[1] * g[1] + n[2] * g[2]);
    if (m == 0.f) { grad[0] = grad[1] = grad[2] = 0.f; return; }
    float u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) u[c] = w_l1 * sgnf(n[c] - g[c]) - w_cos * g[c];
    const float nu = n[0] * u[0] + n[1] * u[1] + n[2] * u[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) grad[c] = (u[c] - n[c] * nu) / dn;
}

// eikonal term of one point: (|g| - 1)^2 and its gradient (0 at g = 0, as torch's norm backward)
NHD float eikonal_term(const float g[3], float grad[3]) {
    const float a = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float k = a > 0.f ? 2.0f * (a - 1.0f) / a : 0.f;
    grad[0] = k * g[0]; grad[1] = k * g[1]; grad[2] = k * g[2];
    return (a - 1.0f) * (a - 1.0f);
}

// smoothness term of one point pair (loss.py:84-88): | g1/(|g1|+1e-5) - g2/(|g2|+1e-5) |
NHD float smooth_term(const float g1[3], const float g2[3], float grad1[3], float grad2[3]) {
    const float a1 = sqrtf(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]), a2 = sqrtf(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
    const float e1 = a1 + 1e-5f, e2 = a2 + 1e-5f;
    float d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = g1[c] / e1 - g2[c] / e2;
    const float s = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    float q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) q[c] = s > 0.f ? d[c] / s : 0.f;          // dL/dn1 = q, dL/dn2 = -q
    const float g1q = g1[0] * q[0] + g1[1] * q[1] + g1[2] * q[2], g2q = g2[0] * q[0] + g2[1] * q[1] + g2[2] * q[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        grad1[c] = q[c] / e1 - (a1 > 0.f ? g1[c] * g1q / (a1 * e1 * e1) : 0.f);
        grad2[c] = -(q[c] / e2 - (a2 > 0.f ? g2[c] * g2q / (a2 * e2 * e2) : 0.f));
    }
    return s;
}

// per-image least squares for (scale, shift) of the SSI depth loss, in fp32 like the reference (MiDaS.py:6-26)
NHD void scale_shift(float a00, float a01, float a11, float b0, float b1, float *scale, float *shift) {
    const float det = a00 * a11 - a01 * a01;
    if (det != 0.f) {
        *scale = (a11 * b0 - a01 * b1) / det;
        *shift = (-a01 * b0 + a00 * b1) / det;
    } else {
        *scale = 0.f; *shift = 0.f;
    }
}

}  // namespace nicer
