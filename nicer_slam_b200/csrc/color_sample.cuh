// One-thread-per-point evaluation of the rendering (color) network, mode "idr":
//   rgb = sigmoid(MLP_relu([x, PE_mv(view), normals, feat, enc_color(x)]))
// (/root/reference/code/model/base_networks.py:333-392) and its backward.
#pragma once
#include "sdf_sample.cuh"

namespace nicer {

struct ColorNetView {
    const float *W0t;     // [d_in][64]
    const float *Wt[3];   // hidden l = 1..n-1 transposed
    const float *WL;      // [3][64] row-major (output layer)
    const float *b0;
    const float *b[3];
    float bl[3];
    const LevelInfo *lv;
    const float *table;   // NULL: no color grid
    int L, n_hidden, multires_view, d_view, feature, d_in, off_normal, off_feat, off_grid;
    float df;
    bool detached;
};

NHD float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

template <int C>
NHD void color_forward_sample(const ColorNetView &nv, const float *X, const float *V, const float *N,
                              const float *feat_fm, uint32_t p, uint32_t P, float *col, int cs, float *rgb,
                              float *A_fm, float *DYDX, float *H0) {
    const size_t Ps = P;
    const int n = nv.n_hidden;
    float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
    float v[3] = {V[3 * (size_t)p], V[3 * (size_t)p + 1], V[3 * (size_t)p + 2]};
    float acc[NICER_W];
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) acc[j] = nv.b0[j];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        axpy64(acc, nv.W0t + d * NICER_W, x[d]);
        axpy64(acc, nv.W0t + (3 + d) * NICER_W, v[d]);
        if (H0) { H0[(size_t)d * Ps + p] = x[d]; H0[(size_t)(3 + d) * Ps + p] = v[d]; }
    }
    {
        float fr = 1.0f;
        for (int f = 0; f < nv.multires_view; ++f) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(v[d] * fr, &s, &c);
                const int ks = 6 + 6 * f + d, kc = ks + 3;
                axpy64(acc, nv.W0t + ks * NICER_W, s);
                axpy64(acc, nv.W0t + kc * NICER_W, c);
                if (H0) { H0[(size_t)ks * Ps + p] = s; H0[(size_t)kc * Ps + p] = c; }
            }
            fr *= 2.0f;
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float nd = N[3 * (size_t)p + d];
        axpy64(acc, nv.W0t + (nv.off_normal + d) * NICER_W, nd);
        if (H0) H0[(size_t)(nv.off_normal + d) * Ps + p] = nd;
    }
    for (int j = 0; j < nv.feature; ++j) axpy64(acc, nv.W0t + (nv.off_feat + j) * NICER_W, feat_fm[(size_t)j * Ps + p]);
    if (nv.table) {
        float u[3];
        to_unit(x, nv.df, u);
        for (int l = 0; l < nv.L; ++l) {
            float feat[C], dfeat[3][C];
            if (DYDX) encode_level<C, true>(nv.table, nv.lv[l], u, feat, dfeat);
            else      encode_level<C, false>(nv.table, nv.lv[l], u, feat, dfeat);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                axpy64(acc, nv.W0t + (nv.off_grid + l * C + c) * NICER_W, feat[c]);
                if (H0) H0[(size_t)(nv.off_grid + l * C + c) * Ps + p] = feat[c];
            }
            if (DYDX) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < C; ++c) DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p] = dfeat[d][c];
            }
        }
    }
    for (int l = 1; l <= n; ++l) {
#pragma unroll
        for (int j = 0; j < NICER_W; ++j) {
            const float a = fmaxf(acc[j], 0.f);
            A_fm[((size_t)(l - 1) * NICER_W + j) * Ps + p] = a;
            col[j * cs] = a;
        }
        if (l < n) {
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) acc[j] = nv.b[l - 1][j];
            mv_acc64(acc, nv.Wt[l - 1], col, cs, NICER_W);
        }
    }
    float o[3] = {nv.bl[0], nv.bl[1], nv.bl[2]};
    for (int k = 0; k < NICER_W; ++k) {
        const float a = col[k * cs];
        o[0] += nv.WL[k] * a;
        o[1] += nv.WL[NICER_W + k] * a;
        o[2] += nv.WL[2 * NICER_W + k] * a;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[3 * (size_t)p + c] = sigmoidf_(o[c]);
}

template <int C>
NHD void color_backward_sample(const ColorNetView &nv, const float *X, const float *V, const float *N,
                               const float *feat_fm, uint32_t p, uint32_t P, const float *rgb, const float *A_fm,
                               const float *DYDX, const float *g_rgb, float *grad_x, float *grad_view,
                               float *grad_normals, float *grad_feat_fm, float *grad_table, float *ZB, float *OB,
                               float *col, int cs) {
    const size_t Ps = P;
    const int n = nv.n_hidden;
    float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
    float v[3] = {V[3 * (size_t)p], V[3 * (size_t)p + 1], V[3 * (size_t)p + 2]};
    float ob[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float y = rgb[3 * (size_t)p + c];
        ob[c] = g_rgb[3 * (size_t)p + c] * (1.0f - y) * y;
        OB[(size_t)c * Ps + p] = ob[c];
    }
    for (int k = 0; k < NICER_W; ++k) {
        const size_t o = ((size_t)(n - 1) * NICER_W + k) * Ps + p;
        const float abar = nv.WL[k] * ob[0] + nv.WL[NICER_W + k] * ob[1] + nv.WL[2 * NICER_W + k] * ob[2];
        const float zb = A_fm[o] > 0.f ? abar : 0.f;
        ZB[o] = zb;
        col[k * cs] = zb;
    }
    float q[NICER_W];
    for (int l = n - 1; l >= 1; --l) {
#pragma unroll
        for (int k = 0; k < NICER_W; ++k) q[k] = col[k * cs];
        const float *Wt = nv.Wt[l - 1];
        for (int k = 0; k < NICER_W; ++k) {
            const size_t o = ((size_t)(l - 1) * NICER_W + k) * Ps + p;
            const float abar = dot64(Wt + k * NICER_W, q);
            const float zb = A_fm[o] > 0.f ? abar : 0.f;
            ZB[o] = zb;
            col[k * cs] = zb;
        }
    }
#pragma unroll
    for (int k = 0; k < NICER_W; ++k) q[k] = col[k * cs];

    float xb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        xb[d] = dot64(nv.W0t + d * NICER_W, q);
    }
    float vb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        vb[d] = dot64(nv.W0t + (3 + d) * NICER_W, q);
    }
    {
        float fr = 1.0f;
        for (int f = 0; f < nv.multires_view; ++f) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(v[d] * fr, &s, &c);
                const int ks = 6 + 6 * f + d, kc = ks + 3;
                const float hs = dot64(nv.W0t + ks * NICER_W, q), hc = dot64(nv.W0t + kc * NICER_W, q);
                vb[d] += fr * (c * hs - s * hc);
            }
            fr *= 2.0f;
        }
    }
    if (grad_view) {
#pragma unroll
        for (int d = 0; d < 3; ++d) grad_view[3 * (size_t)p + d] = vb[d];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int k = nv.off_normal + d;
        grad_normals[3 * (size_t)p + d] = dot64(nv.W0t + k * NICER_W, q);
    }
    for (int j = 0; j < nv.feature; ++j) {
        const int k = nv.off_feat + j;
        grad_feat_fm[(size_t)j * Ps + p] = dot64(nv.W0t + k * NICER_W, q);
    }
    float xu[3] = {0.f, 0.f, 0.f};
    if (nv.table && !nv.detached) {
        float u[3];
        to_unit(x, nv.df, u);
        for (int l = 0; l < nv.L; ++l) {
            float gy[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int k = nv.off_grid + l * C + c;
                gy[c] = dot64(nv.W0t + k * NICER_W, q);
                if (DYDX) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) xu[d] += gy[c] * DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p];
                }
            }
            const LevelInfo li = nv.lv[l];
            Cell3 cell = locate3(li, u);
            if (cell.inside) {
                uint32_t idx[8];
                corner_indices(li, cell, idx);
                float wt[8];
                corner_weights(cell, wt);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float vv[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) vv[c] = wt[k] * gy[c];
                    scatter_entry<C>(grad_table, li, idx[k], vv);
                }
            }
        }
    }
    if (grad_x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) grad_x[3 * (size_t)p + d] += xb[d] + xu[d] / 2.0f / nv.df;
    }
}

}  // namespace nicer
