// TEST INFRASTRUCTURE ONLY.  Host (g++) build of the per-point device functions in
// nicer_slam_b200/csrc/*_sample.cuh behind the same C ABI as libnicer_b200.so, so that on a box
// without a GPU the CPU test-suite can (1) check the kernels' math against the oracle and (2) run
// the Python host layer (autograd wiring, buffer layouts) end to end on CPU tensors.
// The product never loads this library: nicer_slam_b200/_lib.py only ever opens libnicer_b200.so
// and raises if it is missing; tests/conftest.py swaps the handle explicitly for the CPU tests.
#include <stdarg.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

#include "../../include/nicer_b200.h"
#include "../../nicer_slam_b200/csrc/color_sample.cuh"
#include "../../nicer_slam_b200/csrc/geometry_math.cuh"
#include "../../nicer_slam_b200/csrc/loss_math.cuh"
#include "../../nicer_slam_b200/csrc/warp_math.cuh"
#include "../../nicer_slam_b200/csrc/sampler_math.cuh"
#include "../../nicer_slam_b200/csrc/composite_math.cuh"

using namespace nicer;

static thread_local char g_err[512] = "";
static int fail(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}
extern "C" const char *nicer_last_error(void) { return g_err; }
extern "C" int nicer_version(void) { return -1; }   // negative: emulation
extern "C" int nicer_set_tensor_cores(int) { return 0; }

namespace {
struct HostSdf {
    std::vector<float> W0t, Wt[3], WLt, wl_sdf, b0, b[3], bl_feat;
    std::vector<LevelInfo> lv;
    SdfNetView nv;
    explicit HostSdf(const nicer_sdf_net_t &net) {
        const int n = net.n_hidden, L = net.grid.L, C = net.grid.C;
        const int d_pe = 3 + 6 * net.multires, d_in = d_pe + L * C;
        W0t.assign(COL_ROWS * NICER_W, 0.f);
        for (int j = 0; j < NICER_W; ++j) for (int k = 0; k < d_in; ++k) W0t[k * NICER_W + j] = net.W[0][j * d_in + k];
        for (int l = 1; l < n; ++l) {
            Wt[l - 1].assign(NICER_W * NICER_W, 0.f); b[l - 1].assign(NICER_W, 0.f);
            for (int j = 0; j < NICER_W; ++j) { b[l - 1][j] = net.b[l][j]; for (int k = 0; k < NICER_W; ++k) Wt[l - 1][k * NICER_W + j] = net.W[l][j * NICER_W + k]; }
        }
        const int nfeat = net.d_out - 1;
        WLt.assign(NICER_W * NICER_W, 0.f); wl_sdf.assign(NICER_W, 0.f); b0.assign(NICER_W, 0.f); bl_feat.assign(NICER_W, 0.f);
        for (int j = 0; j < nfeat; ++j) for (int k = 0; k < NICER_W; ++k) WLt[k * NICER_W + j] = net.W[n][(1 + j) * NICER_W + k];
        for (int i = 0; i < NICER_W; ++i) { wl_sdf[i] = net.W[n][i]; b0[i] = net.b[0][i]; if (i < nfeat) bl_feat[i] = net.b[n][1 + i]; }
        for (int l = 0; l < L; ++l) lv.push_back(make_level(net.grid.offsets, l, host_level_scales(net.grid.L, net.grid.S, net.grid.H).s[l]));
        nv.W0t = W0t.data();
        for (int i = 0; i < 3; ++i) { nv.Wt[i] = Wt[i].data(); nv.b[i] = b[i].data(); }
        nv.WLt = WLt.data(); nv.wl_sdf = wl_sdf.data(); nv.b0 = b0.data(); nv.bl_feat = bl_feat.data();
        nv.bl_sdf = net.b[n][0]; nv.lv = lv.data(); nv.table = net.grid.table;
        nv.L = L; nv.n_hidden = n; nv.multires = net.multires; nv.d_pe = d_pe; nv.d_in = d_in; nv.df = net.grid.divide_factor;
    }
};

struct HostColor {
    std::vector<float> W0t, Wt[3], WL, b0, b[3];
    std::vector<LevelInfo> lv;
    ColorNetView nv;
    explicit HostColor(const nicer_color_net_t &net) {
        const int n = net.n_hidden;
        const bool hg = net.grid.table != nullptr;
        const int L = hg ? net.grid.L : 0, C = hg ? net.grid.C : 0;
        const int d_view = 3 + 6 * net.multires_view, F = net.feature, d_in = 3 + d_view + 3 + F + L * C;
        W0t.assign(160 * NICER_W, 0.f);
        for (int j = 0; j < NICER_W; ++j) for (int k = 0; k < d_in; ++k) W0t[k * NICER_W + j] = net.W[0][j * d_in + k];
        for (int l = 1; l < n; ++l) {
            Wt[l - 1].assign(NICER_W * NICER_W, 0.f); b[l - 1].assign(NICER_W, 0.f);
            for (int j = 0; j < NICER_W; ++j) { b[l - 1][j] = net.b[l][j]; for (int k = 0; k < NICER_W; ++k) Wt[l - 1][k * NICER_W + j] = net.W[l][j * NICER_W + k]; }
        }
        WL.assign(net.W[n], net.W[n] + 3 * NICER_W); b0.assign(net.b[0], net.b[0] + NICER_W);
        for (int l = 0; l < L; ++l) lv.push_back(make_level(net.grid.offsets, l, host_level_scales(net.grid.L, net.grid.S, net.grid.H).s[l]));
        nv.W0t = W0t.data();
        for (int i = 0; i < 3; ++i) { nv.Wt[i] = Wt[i].data(); nv.b[i] = b[i].data(); }
        nv.WL = WL.data(); nv.b0 = b0.data();
        nv.bl[0] = net.b[n][0]; nv.bl[1] = net.b[n][1]; nv.bl[2] = net.b[n][2];
        nv.lv = lv.data(); nv.table = net.grid.table; nv.L = L; nv.n_hidden = n; nv.multires_view = net.multires_view;
        nv.d_view = d_view; nv.feature = F; nv.d_in = d_in; nv.off_normal = 3 + d_view; nv.off_feat = 3 + d_view + 3;
        nv.off_grid = 3 + d_view + 3 + F; nv.df = hg ? net.grid.divide_factor : 1.0f; nv.detached = net.grid_detached != 0;
    }
};
}  // namespace

#define DISPATCH_C(C, CALL)            \
    switch (C) {                       \
        case 2: { constexpr int CC = 2; CALL; } break; \
        case 4: { constexpr int CC = 4; CALL; } break; \
        case 8: { constexpr int CC = 8; CALL; } break; \
        default: return fail("bad C"); \
    }

extern "C" int nicer_sdf_forward(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t P_feat, uint32_t flags, float *sdf,
                                 float *feat_fm, float *grad, float *Z, float *R, float *DYDX, float *H0, void *) {
    HostSdf h(*net);
    float col[COL_ROWS];
    for (uint32_t p = 0; p < P; ++p)
        DISPATCH_C(net->grid.C, (sdf_forward_sample<CC>(h.nv, x, p, P, flags, col, 1, sdf, feat_fm, grad, Z, R, DYDX, H0, P_feat)));
    return 0;
}

extern "C" int nicer_sdf_backward(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t P_feat, const float *Z,
                                  const float *R, const float *DYDX, const float * /*H0*/, const float *g_sdf, const float *g_feat_fm,
                                  const float *g_grad, float *grad_x, float *grad_table, float *ZB, float *QB,
                                  float *AB, float *TAN, float *T0, float *tan_sum, float * /*GY*/, void *, void *) {
    HostSdf h(*net);
    float col[COL_ROWS];
    for (uint32_t p = 0; p < P; ++p)
        DISPATCH_C(net->grid.C, (sdf_backward_sample<CC>(h.nv, x, p, P, Z, R, DYDX, g_sdf, g_feat_fm, g_grad, grad_x,
                                                         grad_table, ZB, QB, AB, TAN, T0, col, 1, P_feat)));
    if (tan_sum) {
        const size_t row0 = (size_t)(net->n_hidden - 1) * 64;
        for (int j = 0; j < 64; ++j) {
            float a = 0.f;
            for (uint32_t p = 0; p < P; ++p) a += TAN[(row0 + j) * P + p];
            tan_sum[j] += a;
        }
    }
    return 0;
}

extern "C" int nicer_color_forward(const nicer_color_net_t *net, const float *x, const float *view,
                                   const float *normals, const float *feat_fm, uint32_t P, float *rgb, float *A_fm,
                                   float *DYDX, float *H0, void *) {
    HostColor h(*net);
    float col[NICER_W];
    const uint32_t C = net->grid.table ? net->grid.C : 2;
    for (uint32_t p = 0; p < P; ++p)
        DISPATCH_C(C, (color_forward_sample<CC>(h.nv, x, view, normals, feat_fm, p, P, col, 1, rgb, A_fm, DYDX, H0)));
    return 0;
}

extern "C" int nicer_color_backward(const nicer_color_net_t *net, const float *x, const float *view,
                                    const float *normals, const float *feat_fm, uint32_t P, const float *rgb,
                                    const float *A_fm, const float *DYDX, const float *g_rgb, float *grad_x,
                                    float *grad_view, float *grad_normals, float *grad_feat_fm, float *grad_table,
                                    float *ZB, float *OB, float * /*GY*/, void *, void *) {
    HostColor h(*net);
    float col[NICER_W];
    const uint32_t C = net->grid.table ? net->grid.C : 2;
    for (uint32_t p = 0; p < P; ++p)
        DISPATCH_C(C, (color_backward_sample<CC>(h.nv, x, view, normals, feat_fm, p, P, rgb, A_fm, DYDX, g_rgb, grad_x,
                                                grad_view, grad_normals, grad_feat_fm, grad_table, ZB, OB, col, 1)));
    return 0;
}

extern "C" int nicer_outer_accum(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N,
                                 uint32_t P, float *C, uint32_t ldc, float *bias, void *) {
    for (uint32_t m = 0; m < M; ++m) {
        for (uint32_t n = 0; n < N; ++n) {
            double s = 0;
            for (uint32_t p = 0; p < P; ++p) s += (double)A[(size_t)m * lda + p] * B[(size_t)n * ldb + p];
            C[(size_t)m * ldc + n] += (float)s;
        }
        if (bias) { double s = 0; for (uint32_t p = 0; p < P; ++p) s += A[(size_t)m * lda + p]; bias[m] += (float)s; }
    }
    return 0;
}

extern "C" int nicer_outer_accum_batch(const nicer_oa_job_t *jobs, uint32_t n_jobs, uint32_t P, void *st) {
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const nicer_oa_job_t &q = jobs[j];
        if (q.M == 0 || q.N == 0) continue;
        nicer_outer_accum(q.A, q.lda, q.M, q.B, q.ldb, q.N, P, q.C, q.ldc, q.bias, st);
    }
    return 0;
}

static void ray_forward(const float *sdf, const float *X, const float *Z, const float *voxels, int res, uint32_t r,
                        uint32_t S, std::vector<float> &E, std::vector<float> &T, std::vector<float> &beta) {
    float carry = 0.f;
    for (uint32_t i = 0; i < S; ++i) {
        const size_t p = (size_t)r * S + i;
        beta[i] = beta_lookup(voxels, res, X[3 * p], X[3 * p + 1], X[3 * p + 2]);
        const float sigma = laplace_density(sdf[p], beta[i]);
        const float delta = (i + 1 < S) ? (Z[p + 1] - Z[p]) : 1e10f;
        E[i] = delta * sigma;
        T[i] = expf(-carry);
        carry += E[i];
    }
}

extern "C" int nicer_composite_forward(const float *sdf, const float *x, const float *z, const float *rgb,
                                       const float *grad, const float *voxels, uint32_t res, uint32_t R, uint32_t S,
                                       float *weights, float *rgb_out, float *depth_out, float *normal_out,
                                       float *wsum, void *) {
    std::vector<float> E(S), T(S), beta(S);
    for (uint32_t r = 0; r < R; ++r) {
        ray_forward(sdf, x, z, voxels, (int)res, r, S, E, T, beta);
        float aw = 0, awz = 0, ac[3] = {0, 0, 0}, an[3] = {0, 0, 0};
        for (uint32_t i = 0; i < S; ++i) {
            const size_t p = (size_t)r * S + i;
            const float w = (1.0f - expf(-E[i])) * T[i];
            weights[p] = w;
            if (!rgb_out) continue;
            aw += w; awz += w * z[p];
            const float *g = grad + 3 * p;
            const float den = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) + 1e-6f;
            for (int c = 0; c < 3; ++c) { ac[c] += w * rgb[3 * p + c]; an[c] += w * (g[c] / den); }
        }
        if (rgb_out) {
            wsum[r] = aw; depth_out[r] = awz / (aw + 1e-8f);
            for (int c = 0; c < 3; ++c) { rgb_out[3 * r + c] = ac[c]; normal_out[3 * r + c] = an[c]; }
        }
    }
    return 0;
}

extern "C" int nicer_sampler_weights(const float *sdf, const float *x, const float *z, const float *voxels,
                                     uint32_t res, uint32_t R, uint32_t S, float *weights, void *st) {
    return nicer_composite_forward(sdf, x, z, nullptr, nullptr, voxels, res, R, S, weights, nullptr, nullptr, nullptr,
                                   nullptr, st);
}

// ---------------------------------------------------------------------------------------------- flow projection
extern "C" int nicer_flow_project(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K, const float *uv,
                                  const int64_t *idii, uint32_t E, uint32_t n, float *flow, void *) {
    for (uint32_t t = 0; t < E * n; ++t) {
        const uint32_t e = t / n, p = t - e * n, i = (uint32_t)idii[e], ray = i * n + p;
        float X[3];
        for (int a = 0; a < 3; ++a) X[a] = loc[3 * i + a] + depth[ray] * dirs[3 * (size_t)ray + a];
        const WarpProj pr = warp_project(X, w2c + 16 * e, K + 16 * e, 2.0f, 2.0f);
        flow[2 * (size_t)t] = pr.proj[0] / pr.zden - uv[2 * (size_t)ray];
        flow[2 * (size_t)t + 1] = pr.proj[1] / pr.zden - uv[2 * (size_t)ray + 1];
    }
    return 0;
}
extern "C" int nicer_flow_project_backward(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                           const int64_t *idii, uint32_t E, uint32_t n, const float *g_flow, float *g_depth, float *g_dirs,
                                           float *g_loc, float *g_w2c, void *) {
    for (uint32_t t = 0; t < E * n; ++t) {
        const uint32_t e = t / n, p = t - e * n, i = (uint32_t)idii[e], ray = i * n + p;
        const float d = depth[ray];
        float dir[3], X[3];
        for (int a = 0; a < 3; ++a) { dir[a] = dirs[3 * (size_t)ray + a]; X[a] = loc[3 * i + a] + d * dir[a]; }
        const float *Wt = w2c + 16 * e, *Kt = K + 16 * e;
        const WarpProj pr = warp_project(X, Wt, Kt, 2.0f, 2.0f);
        const float gu = g_flow[2 * (size_t)t], gv = g_flow[2 * (size_t)t + 1];
        const float gproj[3] = {gu / pr.zden, gv / pr.zden, -(gu * pr.proj[0] + gv * pr.proj[1]) / (pr.zden * pr.zden)};
        float gcam[3], gX[3];
        for (int c = 0; c < 3; ++c) gcam[c] = Kt[c] * gproj[0] + Kt[4 + c] * gproj[1] + Kt[8 + c] * gproj[2];
        for (int c = 0; c < 3; ++c) gX[c] = Wt[c] * gcam[0] + Wt[4 + c] * gcam[1] + Wt[8 + c] * gcam[2];
        for (int a = 0; a < 3; ++a) {
            for (int c = 0; c < 3; ++c) g_w2c[16 * e + 4 * a + c] += gcam[a] * X[c];
            g_w2c[16 * e + 4 * a + 3] += gcam[a];
            g_dirs[3 * (size_t)ray + a] += d * gX[a];
            g_loc[3 * i + a] += gX[a];
        }
        g_depth[ray] += gX[0] * dir[0] + gX[1] * dir[1] + gX[2] * dir[2];
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- Adam
extern "C" int nicer_adam_step(float *p, float *g, float *m, float *v, uint64_t n, double lr, double beta1, double beta2, double eps,
                               uint64_t step, int zero_grad, void *) {
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2), b2 = (float)beta2;
    const float bc2_sqrt = (float)sqrt(bc2), neg_step = (float)(-(lr / bc1)), e = (float)eps;
    // rounded operation by operation like torch's CPU kernels (no FMA contraction on this build)
    for (uint64_t i = 0; i < n; ++i) {
        const float gk = g[i];
        volatile float d = gk - m[i], wd = w1 * d;
        m[i] = m[i] + wd;
        volatile float t = v[i] * b2, a = w2 * gk, ag = a * gk;
        v[i] = t + ag;
        volatile float q = sqrtf(v[i]) / bc2_sqrt, denom = q + e, r = m[i] / denom, sr = neg_step * r;
        p[i] = p[i] + sr;
        if (zero_grad) g[i] = 0.f;
    }
    return 0;
}
extern "C" int nicer_set_adam_variant(int) { return 0; }

// ---------------------------------------------------------------------------------------------- weight norm
extern "C" int nicer_weight_norm(const nicer_wn_job_t *jobs, uint32_t n, void *) {
    for (uint32_t k = 0; k < n; ++k) {
        const nicer_wn_job_t &J = jobs[k];
        for (uint32_t r = 0; r < J.rows; ++r) {
            const float *v = J.v + (size_t)r * J.cols;
            float ss = 0.f;
            for (uint32_t c = 0; c < J.cols; ++c) ss += v[c] * v[c];
            const float nrm = sqrtf(ss), s = J.g[r] / nrm;
            for (uint32_t c = 0; c < J.cols; ++c) J.w[(size_t)r * J.cols + c] = v[c] * s;
            if (J.norm) J.norm[r] = nrm;
        }
    }
    return 0;
}
extern "C" int nicer_weight_norm_backward(const nicer_wn_job_t *jobs, uint32_t n, void *) {
    for (uint32_t k = 0; k < n; ++k) {
        const nicer_wn_job_t &J = jobs[k];
        for (uint32_t r = 0; r < J.rows; ++r) {
            const float *v = J.v + (size_t)r * J.cols, *dw = J.dw + (size_t)r * J.cols;
            float dot = 0.f;
            for (uint32_t c = 0; c < J.cols; ++c) dot += dw[c] * v[c];
            const float nrm = J.norm[r], a = J.g[r] / nrm, b = dot / (nrm * nrm);
            for (uint32_t c = 0; c < J.cols; ++c) J.dv[(size_t)r * J.cols + c] = a * (dw[c] - v[c] * b);
            J.dg[r] = dot / nrm;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- hierarchical sampler
extern "C" int nicer_sampler_uniform(const float *cam_loc, const float *ray_dirs, float near, float far_cap, float bound, int use_cube,
                                     const float *rnd, uint32_t R, uint32_t N, float *z, float *far_out, float *points, void *) {
    for (uint32_t r = 0; r < R; ++r) {
        const float far = use_cube ? cube_far(cam_loc + 3 * r, ray_dirs + 3 * r, bound, far_cap) : far_cap;
        if (far_out) far_out[r] = far;
        for (uint32_t i = 0; i < N; ++i) {
            const size_t e = (size_t)r * N + i;
            z[e] = uniform_z(near, far, N, i, rnd != nullptr, rnd ? rnd[e] : 0.f);
            if (points)
                for (int a = 0; a < 3; ++a) points[3 * e + a] = cam_loc[3 * r + a] + fmul_exact(z[e], ray_dirs[3 * r + a]);
        }
    }
    return 0;
}

extern "C" int nicer_sampler_resample(const float *sdf, const float *x, const float *z, const float *voxels, uint32_t res,
                                      uint32_t R, uint32_t U, uint32_t N, const int64_t *sel, uint32_t n_extra, float near,
                                      const float *far, const int64_t *eik_idx, float *z_out, float *z_eik, float *weights,
                                      void *st) {
    std::vector<float> w((size_t)R * U);
    nicer_composite_forward(sdf, x, z, nullptr, nullptr, voxels, res, R, U, w.data(), nullptr, nullptr, nullptr, nullptr, st);
    const uint32_t S = N + 2 + n_extra;
    std::vector<float> cdf(U), srt(S);
    for (uint32_t r = 0; r < R; ++r) {
        const float *wr = w.data() + (size_t)r * U, *zr = z + (size_t)r * U;
        if (weights) memcpy(weights + (size_t)r * U, wr, sizeof(float) * U);
        float tot = 0.f;
        for (uint32_t i = 0; i + 1 < U; ++i) tot += wr[i] + 1e-5f;
        cdf[0] = 0.f;
        float run = 0.f;
        for (uint32_t i = 0; i + 1 < U; ++i) { run += (wr[i] + 1e-5f) / tot; cdf[i + 1] = run; }
        cdf[U - 1] = std::min(cdf[U - 1], 1.0f);      // see sampler.cu: the outcome of exact arithmetic at u = 1
        for (uint32_t k = 0; k < N; ++k) srt[k] = invert_cdf(cdf.data(), zr, U, linspace_at(0.f, 1.f, N, k));
        srt[N] = near;
        srt[N + 1] = far[r];
        for (uint32_t j = 0; j < n_extra; ++j) srt[N + 2 + j] = zr[(uint32_t)sel[j]];
        std::sort(srt.begin(), srt.end());
        memcpy(z_out + (size_t)r * S, srt.data(), sizeof(float) * S);
        if (z_eik) z_eik[r] = srt[(uint32_t)eik_idx[r]];
    }
    return 0;
}

extern "C" int nicer_composite_backward(const float *sdf, const float *x, const float *z, const float *rgb,
                                        const float *grad, const float *voxels, uint32_t res, uint32_t R, uint32_t S,
                                        const float *weights, const float *depth_out, const float *wsum,
                                        const float *g_rgb_out, const float *g_depth_out, const float *g_normal_out,
                                        const float *g_weights, float *g_sdf, float *g_rgb, float *g_grad, void *) {
    std::vector<float> E(S), T(S), beta(S);
    for (uint32_t r = 0; r < R; ++r) {
        ray_forward(sdf, x, z, voxels, (int)res, r, S, E, T, beta);
        float go_rgb[3] = {0, 0, 0}, go_n[3] = {0, 0, 0};
        for (int c = 0; c < 3; ++c) { if (g_rgb_out) go_rgb[c] = g_rgb_out[3 * r + c]; if (g_normal_out) go_n[c] = g_normal_out[3 * r + c]; }
        const float go_d = g_depth_out ? g_depth_out[r] : 0.f, dep = depth_out[r], inv_ws = 1.0f / (wsum[r] + 1e-8f);
        float suffix = 0.f;
        for (int i = (int)S - 1; i >= 0; --i) {
            const size_t p = (size_t)r * S + i;
            const float delta = ((uint32_t)i + 1 < S) ? (z[p + 1] - z[p]) : 1e10f;
            const float w = weights[p];
            const float wbar = composite_wbar(go_rgb, go_n, go_d, rgb + 3 * p, grad + 3 * p, z[p], dep, inv_ws,
                                              g_weights ? g_weights[p] : 0.f);
            composite_sample_backward(suffix, wbar, T[i], expf(-E[i]), delta, sdf[p], beta[i], w, go_rgb, go_n,
                                      grad + 3 * p, g_sdf + p, g_rgb + 3 * p, g_grad + 3 * p);
            suffix += -wbar * w;
        }
    }
    return 0;
}

extern "C" int nicer_voxel_count(const float *X, uint32_t P, float *voxels, uint32_t res, void *) {
    for (uint32_t p = 0; p < P; ++p) {
        const float x = X[3 * (size_t)p], y = X[3 * (size_t)p + 1], z = X[3 * (size_t)p + 2];
        if (fabsf(x) > 0.99f || fabsf(y) > 0.99f || fabsf(z) > 0.99f) continue;
        const int ix = (int)((x + 1.0f) / 2.0f * (float)res), iy = (int)((y + 1.0f) / 2.0f * (float)res),
                  iz = (int)((z + 1.0f) / 2.0f * (float)res);
        voxels[((size_t)ix * res + iy) * res + iz] += 1.0f;
    }
    return 0;
}

// ---- drop-in hash op (D == 3 only in the emulation), via the shared nicer_math.cuh helpers
template <int C>
static void hash_fwd(const float *in, const float *emb, const int32_t *off, float *out, uint32_t B, uint32_t L, float S,
                     uint32_t H, int dx, float *dy_dx) {
    for (uint32_t l = 0; l < L; ++l) {
        const LevelInfo li = make_level(off, l, host_level_scales(L, S, H).s[l]);
        for (uint32_t b = 0; b < B; ++b) {
            float feat[C], df[3][C];
            if (dx) encode_level<C, true>(emb, li, in + 3 * (size_t)b, feat, df);
            else encode_level<C, false>(emb, li, in + 3 * (size_t)b, feat, df);
            for (int c = 0; c < C; ++c) out[((size_t)l * B + b) * C + c] = feat[c];
            if (dx) for (int d = 0; d < 3; ++d) for (int c = 0; c < C; ++c) dy_dx[(((size_t)b * L + l) * 3 + d) * C + c] = df[d][c];
        }
    }
}

template <int C>
static void hash_bwd(const float *grad, const float *in, const int32_t *off, float *gg, uint32_t B, uint32_t L, float S,
                     uint32_t H, const float *ggx /* NULL: first order */) {
    for (uint32_t l = 0; l < L; ++l) {
        const LevelInfo li = make_level(off, l, host_level_scales(L, S, H).s[l]);
        for (uint32_t b = 0; b < B; ++b) {
            Cell3 cell = locate3(li, in + 3 * (size_t)b);
            if (!cell.inside) continue;
            uint32_t idx[8]; corner_indices(li, cell, idx);
            float wt[8];
            if (!ggx) corner_weights(cell, wt);
            else {
                float d0[8], d1[8], d2[8];
                corner_dweights(cell, 0, d0); corner_dweights(cell, 1, d1); corner_dweights(cell, 2, d2);
                for (int k = 0; k < 8; ++k) wt[k] = d0[k] * ggx[3 * (size_t)b] + d1[k] * ggx[3 * (size_t)b + 1] + d2[k] * ggx[3 * (size_t)b + 2];
            }
            for (int k = 0; k < 8; ++k) {
                float v[C];
                for (int c = 0; c < C; ++c) v[c] = wt[k] * grad[((size_t)l * B + b) * C + c];
                scatter_entry<C>(gg, li, idx[k], v);
            }
        }
    }
}

extern "C" int nicer_hash_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                                         float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                         uint32_t H, int calc_grad_inputs, float *dy_dx, void *) {
    if (D != 3) return fail("emulation: D must be 3");
    DISPATCH_C(C, (hash_fwd<CC>(inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs, dy_dx)));
    return 0;
}

extern "C" int nicer_hash_encode_backward(const float *grad, const float *inputs, const float *, const int32_t *offsets,
                                          float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                          uint32_t H, int calc_grad_inputs, const float *dy_dx, float *grad_inputs, void *) {
    if (D != 3) return fail("emulation: D must be 3");
    DISPATCH_C(C, (hash_bwd<CC>(grad, inputs, offsets, grad_embeddings, B, L, S, H, nullptr)));
    if (calc_grad_inputs)
        for (uint32_t b = 0; b < B; ++b) for (int d = 0; d < 3; ++d) {
            float r = 0;
            for (uint32_t l = 0; l < L; ++l) for (uint32_t c = 0; c < C; ++c)
                r += grad[((size_t)l * B + b) * C + c] * dy_dx[(((size_t)b * L + l) * 3 + d) * C + c];
            grad_inputs[3 * (size_t)b + d] = r;
        }
    return 0;
}

extern "C" int nicer_hash_encode_second_backward(const float *grad, const float *inputs, const float *,
                                                 const int32_t *offsets, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                                 float S, uint32_t H, int, const float *dy_dx, const float *ggx,
                                                 float *grad_grad, float *grad2_embeddings, void *) {
    if (D != 3) return fail("emulation: D must be 3");
    for (uint32_t l = 0; l < L; ++l) for (uint32_t b = 0; b < B; ++b) for (uint32_t c = 0; c < C; ++c) {
        float r = 0;
        for (int d = 0; d < 3; ++d) r += ggx[3 * (size_t)b + d] * dy_dx[(((size_t)b * L + l) * 3 + d) * C + c];
        grad_grad[((size_t)l * B + b) * C + c] = r;
    }
    DISPATCH_C(C, (hash_bwd<CC>(grad, inputs, offsets, grad2_embeddings, B, L, S, H, ggx)));
    return 0;
}

// ---------------------------------------------------------------------------------------------- camera / ray helpers
extern "C" int nicer_inv4x4(const float *A, uint32_t B, float *Ai, void *) {
    for (uint32_t b = 0; b < B; ++b) inv4x4(A + 16 * (size_t)b, Ai + 16 * (size_t)b);
    return 0;
}
extern "C" int nicer_inv4x4_backward(const float *Ai, const float *G, uint32_t B, float *GA, void *) {
    for (uint32_t b = 0; b < B; ++b) inv4x4_backward(Ai + 16 * (size_t)b, G + 16 * (size_t)b, GA + 16 * (size_t)b);
    return 0;
}
extern "C" int nicer_pose_from_cam7(const float *cam7, uint32_t B, float *pose, void *) {
    for (uint32_t b = 0; b < B; ++b) pose_from_cam7(cam7 + 7 * (size_t)b, pose + 16 * (size_t)b);
    return 0;
}
extern "C" int nicer_pose_from_cam7_backward(const float *cam7, const float *g_pose, uint32_t B, float *g_cam7, void *) {
    for (uint32_t b = 0; b < B; ++b) pose_from_cam7_backward(cam7 + 7 * (size_t)b, g_pose + 16 * (size_t)b, g_cam7 + 7 * (size_t)b);
    return 0;
}
extern "C" int nicer_camera_rays(const float *uv, const float *pose, const float *K, uint32_t B, uint32_t N, float *dirs,
                                 float *cam_loc, void *) {
    for (uint32_t b = 0; b < B; ++b) {
        const float *P = pose + 16 * (size_t)b;
        for (uint32_t n = 0; n < N; ++n) {
            const size_t i = (size_t)b * N + n;
            float v[3], p[3];
            camera_ray(P, K + 16 * (size_t)b, uv[2 * i], uv[2 * i + 1], dirs + 3 * i, v, p);
        }
        cam_loc[3 * b] = P[3]; cam_loc[3 * b + 1] = P[7]; cam_loc[3 * b + 2] = P[11];
    }
    return 0;
}
extern "C" int nicer_camera_rays_backward(const float *uv, const float *pose, const float *K, uint32_t B, uint32_t N,
                                          const float *g_dirs, const float *g_loc, float *g_pose, void *) {
    for (uint32_t b = 0; b < B; ++b) {
        const float *P = pose + 16 * (size_t)b;
        double acc[9] = {0};
        for (uint32_t n = 0; n < N; ++n) {
            const size_t i = (size_t)b * N + n;
            float d[3], v[3], p[3], gv[3];
            camera_ray(P, K + 16 * (size_t)b, uv[2 * i], uv[2 * i + 1], d, v, p);
            camera_ray_backward(v, g_dirs + 3 * i, gv);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) acc[3 * r + c] += (double)gv[r] * p[c];
        }
        float *g = g_pose + 16 * (size_t)b;
        for (int k = 0; k < 16; ++k) g[k] = 0.f;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) g[4 * r + c] = (float)acc[3 * r + c];
            if (g_loc) g[4 * r + 3] = g_loc[3 * b + r];
        }
    }
    return 0;
}
extern "C" int nicer_ray_points(const float *cam_loc, const float *dirs, const float *z, uint32_t R, uint32_t S, float *points,
                                float *dirs_flat, void *) {
    for (uint32_t r = 0; r < R; ++r)
        for (uint32_t s = 0; s < S; ++s) {
            const size_t i = (size_t)r * S + s;
            for (int k = 0; k < 3; ++k) {
                points[3 * i + k] = cam_loc[3 * (size_t)r + k] + fmul_exact(z[i], dirs[3 * (size_t)r + k]);
                if (dirs_flat) dirs_flat[3 * i + k] = dirs[3 * (size_t)r + k];
            }
        }
    return 0;
}
extern "C" int nicer_ray_points_backward(const float *z, uint32_t R, uint32_t S, const float *g_points, const float *g_dirs_flat,
                                         float *g_loc, float *g_dirs, void *) {
    for (uint32_t r = 0; r < R; ++r) {
        double a[6] = {0};
        for (uint32_t s = 0; s < S; ++s) {
            const size_t i = (size_t)r * S + s;
            for (int k = 0; k < 3; ++k) {
                const float gp = g_points ? g_points[3 * i + k] : 0.f;
                a[k] += gp;
                a[3 + k] += (double)z[i] * gp + (g_dirs_flat ? g_dirs_flat[3 * i + k] : 0.f);
            }
        }
        for (int k = 0; k < 3; ++k) { g_loc[3 * (size_t)r + k] = (float)a[k]; g_dirs[3 * (size_t)r + k] = (float)a[3 + k]; }
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------- SLAMLoss terms
extern "C" int nicer_slam_loss(const nicer_loss_t *args, double *, float *maskf, float *terms, void *) {
    const nicer_loss_t &a = *args;
    double rgb = 0, nl1 = 0, ncos = 0, gtd = 0, gtd_cnt = 0, eik = 0, sm = 0;
    std::vector<double> fr(5 * (size_t)a.B, 0.0);
    for (uint32_t r = 0; r < a.R; ++r) {
        bool pos = false, neg = false;
        for (uint32_t s = 0; s < a.S; ++s) { const float v = a.sdf[(size_t)r * a.S + s]; pos |= v > 0.f; neg |= v < 0.f; }
        const float m = (a.mask_gt && a.mask_gt[r] > 0.5f && pos && neg) ? 1.0f : 0.f;
        maskf[r] = m;
        if (a.rgb_pred)
            for (int c = 0; c < 3; ++c) {
                const float d = a.rgb_pred[3 * (size_t)r + c] - a.rgb_gt[3 * (size_t)r + c];
                rgb += fabsf(d);
                if (a.g_rgb) a.g_rgb[3 * (size_t)r + c] = a.w_rgb * sgnf(d) / (3.0f * (float)a.R);
            }
        if (a.normal_pred) {
            float l1, cs, g[3];
            normal_terms(a.normal_pred + 3 * (size_t)r, a.normal_gt + 3 * (size_t)r, m, a.w_normal_l1 / (float)a.R,
                         a.w_normal_cos / (float)a.R, &l1, &cs, g);
            nl1 += l1; ncos += cs;
            if (a.g_normal) for (int c = 0; c < 3; ++c) a.g_normal[3 * (size_t)r + c] = g[c];
        }
        if (a.depth_pred && a.depth_gt) {
            const float md = a.depth_mask_all ? 1.0f : m;
            if (md != 0.f) {
                const double p = a.depth_pred[r], t = a.depth_gt[r] * 50.0f + 0.5f;
                double *f = fr.data() + 5 * (r / a.N);
                f[0] += p * p; f[1] += p; f[2] += 1.0; f[3] += p * t; f[4] += t;
            }
        }
        if (a.depth_pred && a.gt_depth && a.gt_depth_valid[r] > 0.f) { gtd += fabsf(a.depth_pred[r] - a.gt_depth[r]); gtd_cnt += 1.0; }
    }
    if (a.grad_theta)
        for (uint32_t i = 0; i < a.G; ++i) {
            const float *g1 = a.grad_theta + 3 * (size_t)i;
            float ge[3] = {0, 0, 0}, gs1[3] = {0, 0, 0}, gs2[3] = {0, 0, 0};
            if (a.w_eik > 0.f) eik += eikonal_term(g1, ge);
            if (a.grad_theta_nei && a.w_smooth > 0.f) sm += smooth_term(g1, a.grad_theta_nei + 3 * (size_t)i, gs1, gs2);
            const float ke = a.w_eik / (float)a.G, ks = a.w_smooth / (float)a.G;
            for (int c = 0; c < 3; ++c) {
                if (a.g_theta) a.g_theta[3 * (size_t)i + c] = ke * ge[c] + ks * gs1[c];
                if (a.g_theta_nei) a.g_theta_nei[3 * (size_t)i + c] = ks * gs2[c];
            }
        }
    const bool depth_on = a.depth_pred && a.depth_gt;
    std::vector<float> scale(a.B, 0.f), shift(a.B, 0.f);
    double mt = 0;
    for (uint32_t b = 0; b < a.B; ++b) {
        const double *f = fr.data() + 5 * b;
        if (depth_on) scale_shift((float)f[0], (float)f[1], (float)f[2], (float)f[3], (float)f[4], &scale[b], &shift[b]);
        mt += f[2];
    }
    const float m_total = (float)mt;
    double mse = 0, reg = 0;
    if (a.depth_pred)
        for (uint32_t r = 0; r < a.R; ++r) {
            float g = 0.f;
            if (depth_on && m_total > 0.f) {
                const uint32_t b = r / a.N, n = r - b * a.N;
                const float s = scale[b], sh = shift[b];
                auto msk = [&](uint32_t q) { return a.depth_mask_all ? 1.0f : maskf[q]; };
                auto dif = [&](uint32_t q) { return msk(q) * ((s * a.depth_pred[q] + sh) - (a.depth_gt[q] * 50.0f + 0.5f)); };
                const float m = msk(r);
                const float res = (s * a.depth_pred[r] + sh) - (a.depth_gt[r] * 50.0f + 0.5f);
                mse += (double)(m * res * res);
                g = m * res * s / m_total;
                const float d0 = m * res;
                float sg = 0.f;
                if (n > 0) sg += msk(r - 1) * m * sgnf(d0 - dif(r - 1));
                if (n + 1 < a.N) {
                    const float mm = m * msk(r + 1), dd = dif(r + 1) - d0;
                    reg += (double)(mm * fabsf(dd));
                    sg -= mm * sgnf(dd);
                }
                g += 0.5f * s * m * sg / m_total;
                g *= a.w_depth;
            }
            if (a.gt_depth && gtd_cnt > 0 && a.gt_depth_valid[r] > 0.f) g += a.w_gt_depth * sgnf(a.depth_pred[r] - a.gt_depth[r]) / (float)gtd_cnt;
            if (a.g_depth) a.g_depth[r] = g;
        }
    float depth = 0.f;
    if (depth_on && m_total > 0.f) depth = (float)(mse / (2.0 * (double)m_total)) + 0.5f * (float)(reg / (double)m_total);
    terms[NICER_LOSS_RGB] = a.rgb_pred ? (float)(rgb / (3.0 * (double)a.R)) : 0.f;
    terms[NICER_LOSS_DEPTH] = depth;
    terms[NICER_LOSS_GT_DEPTH] = a.gt_depth ? (float)(gtd / gtd_cnt) : 0.f;
    terms[NICER_LOSS_NORMAL_L1] = a.normal_pred ? (float)(nl1 / (double)a.R) : 0.f;
    terms[NICER_LOSS_NORMAL_COS] = a.normal_pred ? (float)(ncos / (double)a.R) : 0.f;
    terms[NICER_LOSS_EIKONAL] = (a.grad_theta && a.w_eik > 0.f) ? (float)(eik / (double)a.G) : 0.f;
    terms[NICER_LOSS_SMOOTH] = (a.grad_theta_nei && a.w_smooth > 0.f) ? (float)(sm / (double)a.G) : 0.f;
    terms[NICER_LOSS_SUM] = a.w_rgb * terms[NICER_LOSS_RGB] + a.w_depth * terms[NICER_LOSS_DEPTH] +
                            (a.gt_depth ? a.w_gt_depth * terms[NICER_LOSS_GT_DEPTH] : 0.f) +
                            a.w_normal_l1 * terms[NICER_LOSS_NORMAL_L1] + a.w_normal_cos * terms[NICER_LOSS_NORMAL_COS] +
                            a.w_eik * terms[NICER_LOSS_EIKONAL] + a.w_smooth * terms[NICER_LOSS_SMOOTH];
    return 0;
}


// ---------------------------------------------------------------------------------------------- warp sampling
extern "C" int nicer_warp_sample(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                 const float *img, uint32_t B, uint32_t N, uint32_t pp, uint32_t H, uint32_t W, float *sampled,
                                 uint8_t *mask, void *) {
    const uint32_t E = B * N * pp;
    for (uint32_t e = 0; e < E; ++e) {
        const uint32_t i = e / (N * pp), ray = e / pp;
        float pts[3];
        for (int a = 0; a < 3; ++a) pts[a] = loc[3 * i + a] + depth[ray] * dirs[3 * (size_t)e + a];
        for (uint32_t t = 0; t < B; ++t) {
            const WarpProj pr = warp_project(pts, w2c + 16 * t, K + 16 * t, (float)W, (float)H);
            const size_t o = (size_t)t * E + e;
            bilinear3(img + (size_t)t * H * W * 3, (int)H, (int)W, pr.nu, pr.nv, sampled + 3 * o, nullptr, nullptr, nullptr);
            mask[o] = (pr.nu > -1.0f && pr.nu < 1.0f && pr.nv > -1.0f && pr.nv < 1.0f && pr.proj[2] > 0.f) ? 1 : 0;
        }
    }
    return 0;
}
extern "C" int nicer_warp_sample_backward(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                          const float *img, uint32_t B, uint32_t N, uint32_t pp, uint32_t H, uint32_t W,
                                          const float *g_sampled, float *g_depth, float *g_dirs, float *g_loc, float *g_w2c, void *) {
    const uint32_t E = B * N * pp;
    for (uint32_t e = 0; e < E; ++e) {
        const uint32_t i = e / (N * pp), ray = e / pp;
        const float d = depth[ray];
        float dir[3], pts[3], gp[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) { dir[a] = dirs[3 * (size_t)e + a]; pts[a] = loc[3 * i + a] + d * dir[a]; }
        for (uint32_t t = 0; t < B; ++t) {
            const float *Wt = w2c + 16 * t, *Kt = K + 16 * t;
            const WarpProj pr = warp_project(pts, Wt, Kt, (float)W, (float)H);
            const float *g = g_sampled + 3 * ((size_t)t * E + e);
            float out[3], dnu, dnv;
            bilinear3(img + (size_t)t * H * W * 3, (int)H, (int)W, pr.nu, pr.nv, out, g, &dnu, &dnv);
            const float ax = dnu * 2.0f / (float)W, ay = dnv * 2.0f / (float)H;
            const float gproj[3] = {ax / pr.zden, ay / pr.zden, -(ax * pr.proj[0] + ay * pr.proj[1]) / (pr.zden * pr.zden)};
            float gcam[3];
            for (int c = 0; c < 3; ++c) gcam[c] = Kt[c] * gproj[0] + Kt[4 + c] * gproj[1] + Kt[8 + c] * gproj[2];
            for (int c = 0; c < 3; ++c) gp[c] += Wt[c] * gcam[0] + Wt[4 + c] * gcam[1] + Wt[8 + c] * gcam[2];
            for (int a = 0; a < 3; ++a) {
                for (int c = 0; c < 3; ++c) g_w2c[16 * t + 4 * a + c] += gcam[a] * pts[c];
                g_w2c[16 * t + 4 * a + 3] += gcam[a];
            }
        }
        for (int a = 0; a < 3; ++a) { g_dirs[3 * (size_t)e + a] = d * gp[a]; g_loc[3 * i + a] += gp[a]; }
        g_depth[ray] += gp[0] * dir[0] + gp[1] * dir[1] + gp[2] * dir[2];
    }
    return 0;
}

extern "C" int nicer_warp_gt(const float *uvp, const float *img, const float *dep, uint32_t B, uint32_t M, uint32_t H, uint32_t W,
                             float *gt_rgb, float *gt_depth, uint8_t *inside, void *) {
    for (uint32_t e = 0; e < B * M; ++e) {
        const uint32_t b = e / M;
        const float u = uvp[2 * (size_t)e], v = uvp[2 * (size_t)e + 1];
        const bool in = (0.f <= u) && (0.f <= v) && (u < (float)W) && (v < (float)H);
        float r[3] = {1.0f, 1.0f, 1.0f}, d = 1.0f;
        if (in) {
            const uint32_t ui = std::min((uint32_t)u, W - 1), vi = std::min((uint32_t)v, H - 1);
            const size_t px = ((size_t)b * H + vi) * W + ui;
            r[0] = img[3 * px]; r[1] = img[3 * px + 1]; r[2] = img[3 * px + 2];
            d = dep[px];
        }
        for (int c = 0; c < 3; ++c) gt_rgb[3 * (size_t)e + c] = r[c];
        gt_depth[e] = d;
        inside[e] = in ? 1 : 0;
    }
    return 0;
}

extern "C" size_t nicer_masked_l1_mean_workspace(void) { return 16; }
extern "C" int nicer_masked_l1_mean(const float *a, const float *b, const unsigned char *mask, uint32_t n_mask, uint32_t inner,
                                    uint32_t b_len, void *, float *out, void *) {
    double sum = 0.0, cnt = 0.0;
    for (uint32_t m = 0; m < n_mask; ++m) {
        if (!mask[m]) continue;
        cnt += inner;
        for (uint32_t c = 0; c < inner; ++c) {
            const size_t i = (size_t)m * inner + c;
            sum += fabsf(a[i] - b[i % b_len]);
        }
    }
    out[0] = (float)sum / (float)cnt;
    out[1] = (float)cnt;
    return 0;
}
extern "C" int nicer_masked_l1_mean_backward(const float *a, const float *b, const unsigned char *mask, uint32_t n_mask,
                                             uint32_t inner, uint32_t b_len, const float *out, const float *g, float *ga, void *) {
    for (size_t i = 0; i < (size_t)n_mask * inner; ++i)
        ga[i] = mask[i / inner] ? g[0] * sgnf(a[i] - b[i % b_len]) / out[1] : 0.f;
    return 0;
}
