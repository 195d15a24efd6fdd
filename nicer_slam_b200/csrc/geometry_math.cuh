// Per-element math of the camera / ray helpers on the hot path (host + device), forward and hand-derived backward:
//   quad2rotation / get_camera_from_tensor   /root/reference/code/utils/general.py:52-100
//   get_camera_params / lift                 /root/reference/code/utils/rend_util.py:68-93,107-129
// The reference runs these as ~100 tiny elementwise kernels per call (and as many again in backward); here each is
// one kernel.  Expression order follows the reference so that the results agree to the last bits that matter.
#pragma once
#include "nicer_math.cuh"

namespace nicer {

// [7] (quat w,x,y,z un-normalised; translation) -> c2w 4x4 row-major
NHD void pose_from_cam7(const float *c, float *P) {
    const float qr = c[0], qi = c[1], qj = c[2], qk = c[3];
    const float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
    P[0] = 1.0f - two_s * (qj * qj + qk * qk); P[1] = two_s * (qi * qj - qk * qr);        P[2] = two_s * (qi * qk + qj * qr);
    P[4] = two_s * (qi * qj + qk * qr);        P[5] = 1.0f - two_s * (qi * qi + qk * qk); P[6] = two_s * (qj * qk - qi * qr);
    P[8] = two_s * (qi * qk - qj * qr);        P[9] = two_s * (qj * qk + qi * qr);        P[10] = 1.0f - two_s * (qi * qi + qj * qj);
    P[3] = c[4]; P[7] = c[5]; P[11] = c[6];
    P[12] = 0.f; P[13] = 0.f; P[14] = 0.f; P[15] = 1.0f;
}

// g [16] = dL/dpose -> gc [7] = dL/dcam7.  R = I + s * M(q) with s = 2/|q|^2:  dR = s dM + M ds,  ds/dq = -s^2 q.
NHD void pose_from_cam7_backward(const float *c, const float *g, float *gc) {
    const float qr = c[0], qi = c[1], qj = c[2], qk = c[3];
    const float s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
    // M entries (R = I + s M on the diagonal the "1 -" sign is folded in)
    const float m00 = -(qj * qj + qk * qk), m01 = qi * qj - qk * qr, m02 = qi * qk + qj * qr;
    const float m10 = qi * qj + qk * qr, m11 = -(qi * qi + qk * qk), m12 = qj * qk - qi * qr;
    const float m20 = qi * qk - qj * qr, m21 = qj * qk + qi * qr, m22 = -(qi * qi + qj * qj);
    const float g00 = g[0], g01 = g[1], g02 = g[2], g10 = g[4], g11 = g[5], g12 = g[6], g20 = g[8], g21 = g[9], g22 = g[10];
    const float gs = g00 * m00 + g01 * m01 + g02 * m02 + g10 * m10 + g11 * m11 + g12 * m12 + g20 * m20 + g21 * m21 + g22 * m22;
    // dL/dq through M (times s)
    const float dr = -qk * g01 + qj * g02 + qk * g10 - qi * g12 - qj * g20 + qi * g21;
    const float di = qj * g01 + qk * g02 + qj * g10 - 2.0f * qi * g11 - qr * g12 + qk * g20 + qr * g21 - 2.0f * qi * g22;
    const float dj = -2.0f * qj * g00 + qi * g01 + qr * g02 + qi * g10 + qk * g12 - qr * g20 + qk * g21 - 2.0f * qj * g22;
    const float dk = -2.0f * qk * g00 - qr * g01 + qi * g02 + qr * g10 - 2.0f * qk * g11 + qj * g12 + qi * g20 + qj * g21;
    const float ds = -s * s * gs;
    gc[0] = s * dr + ds * qr;
    gc[1] = s * di + ds * qi;
    gc[2] = s * dj + ds * qj;
    gc[3] = s * dk + ds * qk;
    gc[4] = g[3]; gc[5] = g[7]; gc[6] = g[11];
}

// pixel -> camera-space point at depth 1 (lift, with skew)
NHD void lift_pixel(const float *K, float x, float y, float p[3]) {
    const float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
    p[0] = (x - cx + cy * sk / fy - sk * y / fy) / fx;
    p[1] = (y - cy) / fy;
    p[2] = 1.0f;
}

// ray direction of a pixel: v = (R p + t) - t, d = v / (v . v)   (rend_util.py:88-92: divided by the SQUARED norm)
NHD void camera_ray(const float *P, const float *K, float x, float y, float d[3], float v[3], float p[3]) {
    lift_pixel(K, x, y, p);
    // world = bmm(pose, [p; 1]) is a K = 4 GEMM: a sequential FMA chain over k (what the BLAS kernels do), then "- t";
    // the squared norm is an elementwise product followed by a sum (separately rounded products, no FMA)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float w = fmaf(P[4 * i + 2], p[2], fmaf(P[4 * i + 1], p[1], fmul_exact(P[4 * i], p[0]))) + P[4 * i + 3];
        v[i] = w - P[4 * i + 3];
    }
    const float n = (fmul_exact(v[0], v[0]) + fmul_exact(v[1], v[1])) + fmul_exact(v[2], v[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = v[i] / n;
}

// dL/dv from dL/dd:  d = v / n, n = v.v  ->  g_v = g_d / n - 2 v (v . g_d) / n^2
NHD void camera_ray_backward(const float v[3], const float gd[3], float gv[3]) {
    const float n = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float vg = v[0] * gd[0] + v[1] * gd[1] + v[2] * gd[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) gv[i] = gd[i] / n - 2.0f * v[i] * vg / (n * n);
}

// 4x4 inverse by Gauss-Jordan elimination with partial pivoting (what torch.inverse / linalg.inv_ex do through LU with
// partial pivoting; network.py:157,171 invert the c2w poses).  A, Ai row-major; a singular matrix gives inf / nan, as LU does.
NHD void inv4x4(const float *A, float *Ai) {
    float m[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { m[r][c] = A[4 * r + c]; m[r][4 + c] = (r == c) ? 1.0f : 0.0f; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        float best = fabsf(m[k][k]);
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            const float a = fabsf(m[r][k]);
            if (a > best) { best = a; piv = r; }
        }
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            if (r == piv) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { const float t = m[k][c]; m[k][c] = m[r][c]; m[r][c] = t; }
            }
        }
        const float inv = 1.0f / m[k][k];
#pragma unroll
        for (int c = 0; c < 8; ++c) m[k][c] *= inv;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const float f = m[r][k];
#pragma unroll
            for (int c = 0; c < 8; ++c) m[r][c] -= f * m[k][c];
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) Ai[4 * r + c] = m[r][4 + c];
}

// dL/dA = -Ai^T G Ai^T   (G = dL/dAi)
NHD void inv4x4_backward(const float *Ai, const float *G, float *GA) {
    float T[16];     // T = Ai^T G
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += Ai[4 * k + r] * G[4 * k + c];
            T[4 * r + c] = a;
        }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += T[4 * r + k] * Ai[4 * c + k];
            GA[4 * r + c] = -a;
        }
}

}  // namespace nicer
