// One-thread-per-point evaluation of an SDF network (hash grid + NeRF PE -> weight-normed Softplus MLP)
// with the analytic d sdf/dx, and its full backward including the second-order terms that the
// reference obtains through autograd double backward
// (/root/reference/code/model/base_networks.py:155-221; hashencoder/hashgrid.py:54-134).
//
// Notation (n hidden layers of width 64):
//   h0 = [x, sin(2^f x), cos(2^f x) (f<m), enc(u)],  u = (x/df + 1)/2
//   z_1 = W_0 h0 + b_0, a_l = softplus(z_l), z_{l+1} = W_l a_l + b_l, out = W_n a_n + b_n = [sdf, feat]
//   gradient pass:   r_n = W_n[0,:],  q_l = r_l * sp'(z_l),  r_{l-1} = W_{l-1}^T q_l,  g = J^T r_0
//   backward of it:  tangent pass t_0 = J g_bar, u_1 = W_0 t_0, tan_l = u_l * sp'(z_l), u_{l+1} = W_l tan_l
//                    dL/dz_l = abar_l * sp'(z_l) + u_l * r_l * sp''(z_l)
//                    dL/dW_l = zbar_{l+1} a_l^T + q_{l+1} tan_l^T     (accumulated by nicer_outer_accum)
//   Second derivatives of the hash encoding w.r.t. x are dropped, as in the reference
//   (hashgrid.py:134 returns None for the inputs of the second backward); PE second derivatives are kept.
#pragma once
#include "nicer_math.cuh"

#ifndef NICER_ATOMIC_ADD
#ifdef __CUDA_ARCH__
#define NICER_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define NICER_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#endif
#endif

namespace nicer {

constexpr uint32_t F_SDF_ONLY = 1u, F_ACCUMULATE = 2u, F_NO_FEAT = 4u;
constexpr int COL_ROWS = 72;  // private scratch column height (>= max d_in of an SDF net = 39 + 32)

struct SdfNetView {
    const float *W0t;     // [d_in][64]
    const float *Wt[3];   // W_l^T for l = 1..n-1: Wt[l-1][k*64+j] = W_l[j][k]
    const float *WLt;     // [64][64]: WLt[k*64+j] = W_n[1+j][k]
    const float *wl_sdf;  // [64] = W_n[0,:]
    const float *b0;      // [64]
    const float *b[3];    // b_l for l = 1..n-1
    const float *bl_feat; // [64] = b_n[1:]
    float bl_sdf;
    const LevelInfo *lv;
    const float *table;
    int L, n_hidden, multires, d_pe, d_in;
    float df;
};

template <int C>
NHD void scatter_entry(float *grad_table, const LevelInfo &li, uint32_t idx, const float v[C]) {
    if (!grad_table) return;        // table gradient not wanted (pose-only tracking)
    float *p = grad_table + ((size_t)li.offset + idx) * C;
#ifdef __CUDA_ARCH__
    if constexpr (C == 8) {
        atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
        atomicAdd(reinterpret_cast<float4 *>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
    } else if constexpr (C == 4) {
        atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else if constexpr (C == 2) {
        atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(p + c, v[c]);
    }
#else
    for (int c = 0; c < C; ++c) p[c] += v[c];
#endif
}

NHD void to_unit(const float x[3], float df, float u[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = (x[d] / df + 1.0f) / 2.0f;
}

template <int C>
NHD void sdf_forward_sample(const SdfNetView &nv, const float *X, uint32_t p, uint32_t P, uint32_t flags,
                            float *col, int cs, float *sdf, float *feat_fm, float *grad, float *Z, float *R,
                            float *DYDX, float *H0, uint32_t Pf = 0) {
    // Pf: only the first Pf points have a feature row block (feat_fm is [64][Pf]); 0 = all P points
    const size_t Ps = P, Pfs = Pf ? Pf : P;
    const int n = nv.n_hidden;
    const bool sdf_only = (flags & F_SDF_ONLY) != 0;
    const bool accumulate = (flags & F_ACCUMULATE) != 0;
    float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
    float u[3];
    to_unit(x, nv.df, u);

    float acc[NICER_W];
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) acc[j] = nv.b0[j];
    // ---- layer 0, input generated on the fly
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        axpy64(acc, nv.W0t + d * NICER_W, x[d]);
        if (H0) H0[(size_t)d * Ps + p] = x[d];
    }
    {
        float fr = 1.0f;
        for (int f = 0; f < nv.multires; ++f) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(x[d] * fr, &s, &c);
                const int ks = 3 + 6 * f + d, kc = ks + 3;
                axpy64(acc, nv.W0t + ks * NICER_W, s);
                axpy64(acc, nv.W0t + kc * NICER_W, c);
                if (H0) { H0[(size_t)ks * Ps + p] = s; H0[(size_t)kc * Ps + p] = c; }
            }
            fr *= 2.0f;
        }
    }
    for (int l = 0; l < nv.L; ++l) {
        float feat[C], dfeat[3][C];
        if (sdf_only) encode_level<C, false>(nv.table, nv.lv[l], u, feat, dfeat);
        else          encode_level<C, true>(nv.table, nv.lv[l], u, feat, dfeat);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            axpy64(acc, nv.W0t + (nv.d_pe + l * C + c) * NICER_W, feat[c]);
            if (H0) H0[(size_t)(nv.d_pe + l * C + c) * Ps + p] = feat[c];
        }
        if (!sdf_only) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int c = 0; c < C; ++c) DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p] = dfeat[d][c];
        }
    }
    // ---- hidden layers
    for (int l = 1; l <= n; ++l) {
        if (!sdf_only) {
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) Z[((size_t)(l - 1) * NICER_W + j) * Ps + p] = acc[j];
        }
#pragma unroll
        for (int j = 0; j < NICER_W; ++j) col[j * cs] = softplus100(acc[j]);
        if (l < n) {
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) acc[j] = nv.b[l - 1][j];
            mv_acc64(acc, nv.Wt[l - 1], col, cs, NICER_W);
        }
    }
    // ---- output layer
    {
        float s = nv.bl_sdf;
        for (int k = 0; k < NICER_W; ++k) s += nv.wl_sdf[k] * col[k * cs];
        if (accumulate) sdf[p] += s; else sdf[p] = s;
    }
    if (sdf_only) return;
    if (!(flags & F_NO_FEAT) && p < Pfs) {
#pragma unroll
        for (int j = 0; j < NICER_W; ++j) acc[j] = nv.bl_feat[j];
        mv_acc64(acc, nv.WLt, col, cs, NICER_W);
        if (accumulate) {
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) feat_fm[(size_t)j * Pfs + p] += acc[j];
        } else {
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) feat_fm[(size_t)j * Pfs + p] = acc[j];
        }
    }
    // ---- gradient pass: q holds q_{l+1}
    float *q = acc;
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) q[j] = nv.wl_sdf[j] * dsoftplus100(Z[((size_t)(n - 1) * NICER_W + j) * Ps + p]);
    for (int l = n - 1; l >= 1; --l) {
        const float *Wt = nv.Wt[l - 1];
        for (int k = 0; k < NICER_W; ++k) {
            float r = dot64(Wt + k * NICER_W, q);
            R[((size_t)(l - 1) * NICER_W + k) * Ps + p] = r;
            col[k * cs] = r * dsoftplus100(Z[((size_t)(l - 1) * NICER_W + k) * Ps + p]);
        }
#pragma unroll
        for (int k = 0; k < NICER_W; ++k) q[k] = col[k * cs];
    }
    float g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] = dot64(nv.W0t + d * NICER_W, q);
    {
        float fr = 1.0f;
        for (int f = 0; f < nv.multires; ++f) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(x[d] * fr, &s, &c);
                float rs = dot64(nv.W0t + (3 + 6 * f + d) * NICER_W, q);
                float rc = dot64(nv.W0t + (3 + 6 * f + 3 + d) * NICER_W, q);
                g[d] += fr * (c * rs - s * rc);
            }
            fr *= 2.0f;
        }
    }
    float gu[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < nv.L; ++l) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float r0 = dot64(nv.W0t + (nv.d_pe + l * C + c) * NICER_W, q);
#pragma unroll
            for (int d = 0; d < 3; ++d) gu[d] += r0 * DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        g[d] += gu[d] / 2.0f / nv.df;
        if (accumulate) grad[3 * (size_t)p + d] += g[d]; else grad[3 * (size_t)p + d] = g[d];
    }
}

template <int C>
NHD void sdf_backward_sample(const SdfNetView &nv, const float *X, uint32_t p, uint32_t P, const float *Z,
                             const float *R, const float *DYDX, const float *g_sdf, const float *g_feat_fm,
                             const float *g_grad, float *grad_x, float *grad_table, float *ZB, float *QB,
                             float *AB, float *TAN, float *T0, float *col, int cs, uint32_t Pf = 0) {
    // Pf: g_sdf [Pf] and g_feat_fm [64][Pf] cover the first Pf points only (zero upstream gradient beyond); 0 = all P points
    const size_t Ps = P, Pfs = Pf ? Pf : P;
    const int n = nv.n_hidden;
    float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
    float u[3];
    to_unit(x, nv.df, u);
    const float gs = (g_sdf && p < Pfs) ? g_sdf[p] : 0.f;
    float gg[3] = {0.f, 0.f, 0.f}, ggu[3];
    if (g_grad) { gg[0] = g_grad[3 * (size_t)p]; gg[1] = g_grad[3 * (size_t)p + 1]; gg[2] = g_grad[3 * (size_t)p + 2]; }
#pragma unroll
    for (int d = 0; d < 3; ++d) ggu[d] = gg[d] / 2.0f / nv.df;

    // ---------------- tangent pass (forward-mode in direction gg)
    float acc[NICER_W];
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) acc[j] = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        axpy64(acc, nv.W0t + d * NICER_W, gg[d]);
        T0[(size_t)d * Ps + p] = gg[d];
    }
    {
        float fr = 1.0f;
        for (int f = 0; f < nv.multires; ++f) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(x[d] * fr, &s, &c);
                const int ks = 3 + 6 * f + d, kc = ks + 3;
                const float ts = fr * c * gg[d], tc = -fr * s * gg[d];
                axpy64(acc, nv.W0t + ks * NICER_W, ts);
                axpy64(acc, nv.W0t + kc * NICER_W, tc);
                T0[(size_t)ks * Ps + p] = ts;
                T0[(size_t)kc * Ps + p] = tc;
            }
            fr *= 2.0f;
        }
    }
    for (int l = 0; l < nv.L; ++l) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float t = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) t += ggu[d] * DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p];
            const int k = nv.d_pe + l * C + c;
            T0[(size_t)k * Ps + p] = t;
            axpy64(acc, nv.W0t + k * NICER_W, t);
        }
    }
    for (int l = 1; l <= n; ++l) {
#pragma unroll
        for (int j = 0; j < NICER_W; ++j) {
            const size_t o = ((size_t)(l - 1) * NICER_W + j) * Ps + p;
            const float z = Z[o];
            const float r = (l < n) ? R[o] : nv.wl_sdf[j];
            const SpEval sp = sp_eval(z);
            const float s1 = sp.s1, s2 = sp.s2;
            const float tan = acc[j] * s1;
            TAN[o] = tan;
            QB[o] = r * s1;
            AB[o] = sp.a;
            ZB[o] = acc[j] * r * s2;
            col[j * cs] = tan;
        }
        if (l < n) {
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) acc[j] = 0.f;
            mv_acc64(acc, nv.Wt[l - 1], col, cs, NICER_W);
        }
    }
    // ---------------- reverse pass
    float *q = acc;
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) q[j] = (g_feat_fm && p < Pfs) ? g_feat_fm[(size_t)j * Pfs + p] : 0.f;
    for (int k = 0; k < NICER_W; ++k) {
        const size_t o = ((size_t)(n - 1) * NICER_W + k) * Ps + p;
        const float abar = nv.wl_sdf[k] * gs + dot64(nv.WLt + k * NICER_W, q);
        const float zb = abar * dsoftplus100(Z[o]) + ZB[o];
        ZB[o] = zb;
        col[k * cs] = zb;
    }
    for (int l = n - 1; l >= 1; --l) {
#pragma unroll
        for (int k = 0; k < NICER_W; ++k) q[k] = col[k * cs];
        const float *Wt = nv.Wt[l - 1];
        for (int k = 0; k < NICER_W; ++k) {
            const size_t o = ((size_t)(l - 1) * NICER_W + k) * Ps + p;
            const float abar = dot64(Wt + k * NICER_W, q);
            const float zb = abar * dsoftplus100(Z[o]) + ZB[o];
            ZB[o] = zb;
            col[k * cs] = zb;
        }
    }
#pragma unroll
    for (int k = 0; k < NICER_W; ++k) q[k] = col[k * cs];
    // hbar0[k] = (W_0^T zbar_1)[k] -> col
    for (int k = 0; k < nv.d_in; ++k) col[k * cs] = dot64(nv.W0t + k * NICER_W, q);
    // r_0 = W_0^T q_1 (second-order terms)
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) q[j] = QB[(size_t)j * Ps + p];

    float xb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) xb[d] = col[d * cs];
    {
        float fr = 1.0f;
        for (int f = 0; f < nv.multires; ++f) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(x[d] * fr, &s, &c);
                const int ks = 3 + 6 * f + d, kc = ks + 3;
                xb[d] += fr * (c * col[ks * cs] - s * col[kc * cs]);
                const float rs = dot64(nv.W0t + ks * NICER_W, q), rc = dot64(nv.W0t + kc * NICER_W, q);
                xb[d] += gg[d] * (fr * fr) * (-s * rs - c * rc);
            }
            fr *= 2.0f;
        }
    }
    float xu[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < nv.L; ++l) {
        float gy1[C], gy2[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int k = nv.d_pe + l * C + c;
            gy1[c] = col[k * cs];
            gy2[c] = dot64(nv.W0t + k * NICER_W, q);
#pragma unroll
            for (int d = 0; d < 3; ++d) xu[d] += gy1[c] * DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p];
        }
        const LevelInfo li = nv.lv[l];
        Cell3 cell = locate3(li, u);
        if (cell.inside) {
            uint32_t idx[8];
            corner_indices(li, cell, idx);
            float wt[8], dw0[8], dw1[8], dw2[8];
            corner_weights(cell, wt);
            corner_dweights(cell, 0, dw0);
            corner_dweights(cell, 1, dw1);
            corner_dweights(cell, 2, dw2);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float v[C];
                const float w2 = dw0[k] * ggu[0] + dw1[k] * ggu[1] + dw2[k] * ggu[2];
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = wt[k] * gy1[c] + w2 * gy2[c];
                scatter_entry<C>(grad_table, li, idx[k], v);
            }
        }
    }
    if (grad_x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) grad_x[3 * (size_t)p + d] += xb[d] + xu[d] / 2.0f / nv.df;
    }
}

}  // namespace nicer
