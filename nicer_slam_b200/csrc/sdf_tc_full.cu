// Tensor-core (tcgen05, sm_100a) SDF network, main differentiable pass: forward + analytic d sdf/dx.
// Same math and same saved tensors (Z, R, DYDX) as the fp32 SIMT kernels of sdf_net.cu / sdf_sample.cuh; the
// 64-wide mat-vecs of every layer, of the feature head and of the transposed (gradient) chain run as
// tcgen05.mma kind::tf32 with the 3xTF32 split:
//   * one thread = one point = one TMEM lane; activations / adjoints are written with tcgen05.st as the A operand
//     (hi and lo column ranges), accumulator rows are read back with tcgen05.ld.32x32b by the same thread;
//   * weights (hi, lo) live in shared memory once per CTA in the UMMA K-major no-swizzle layout [k/4][rows][4]
//     (LBO = rows*16 B between the two 16-byte K-chunks of a step, SBO = 128 B between 8-row groups).  tcgen05 has no
//     interleaved MN-major form for 32-bit operands, so the transposed products of the gradient chain
//     r_{l-1} = W_{l-1}^T q_l use separately staged K-major copies of W^T; to fit shared memory the pass is split in
//     two kernels: A (layers + feature head, saves Z) and B (gradient chain from Z, saves R, emits d sdf/dx);
//   * a CTA holds two independent 128-point tiles (256 threads) that share the weights and overlap each other's
//     MMA and epilogue phases; tiles synchronise on named barriers and one mbarrier each.
#include "sdf_tc_plan.cuh"

namespace nicer {

template <int C>
__global__ void __launch_bounds__(TCF_THREADS, 1)
sdf_forward_tc_a_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                      uint32_t P, uint32_t flags, float *sdf, float *feat_fm, float *Z, float *DYDX, float *H0) {
    // H0 != NULL: its grid rows (and DYDX) were already written by grid_encode_kernel; this kernel adds the x / PE rows
    extern __shared__ __align__(16) float smem[];
    __shared__ TcfShared sh;
    LevelInfo *lv;
    Tile t = tcf_setup(net, ls, pl, smem, sh, lv);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const bool accumulate = (flags & NICER_SDF_ACCUMULATE) != 0;
    const bool want_feat = (flags & NICER_SDF_NO_FEAT) == 0;
    const float *wl = smem + pl.wl_sdf;
    const float bl_sdf = net.b[n][0];

    const uint32_t tiles = (P + 127u) / 128u;
    // tile index space: CTA b handles tiles 2*b + {0,1}, 2*(b+grid) + {0,1}, ...
    for (uint32_t tt = blockIdx.x * 2 + (threadIdx.x >> 7); tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (threadIdx.x & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;
        const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
        float u[3];
        to_unit(x, df, u);
        // ---------------- network input -> A   (columns: [32 grid | 39 PE | pad])
        float gf[32];
        if (H0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) gf[k] = (k < L * C) ? __ldg(H0 + (size_t)(39 + k) * Ps + p) : 0.f;
        }
        {
            float pe[48];
            pe[0] = x[0]; pe[1] = x[1]; pe[2] = x[2];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float sc[12];
                pe_sincos<6>(x[d], sc);
#pragma unroll
                for (int f = 0; f < 6; ++f) { pe[3 + 6 * f + d] = sc[2 * f]; pe[3 + 6 * f + 3 + d] = sc[2 * f + 1]; }
            }
#pragma unroll
            for (int k = 39; k < 48; ++k) pe[k] = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 6; ++c8) st_a8(t, 4 + c8, &pe[c8 * 8]);      // columns 32..79
            if (H0 && valid) {
#pragma unroll
                for (int k = 0; k < 39; ++k) H0[(size_t)k * Ps + p] = pe[k];
            }
        }
        if (H0) {
            // grid features were gathered by grid_encode_kernel into the grid rows of H0 (and DYDX): coalesced reads
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) st_a8(t, c8, &gf[c8 * 8]);
        } else {
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                float feat[C], dfeat[3][C];
                if (l < L) {
                    encode_level<C, true>(net.grid.table, lv[l], u, feat, dfeat);
                    if (valid) {
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int c = 0; c < C; ++c) DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p] = dfeat[d][c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) feat[c] = 0.f;
                }
                st_a_small<C>(t, l * C, feat);
            }
        }
        // ---------------- hidden layers
        float s_out = bl_sdf;
        for (int l = 0; l < n; ++l) {
            mat_gemm(t, pl, l, smem);
            const float *bias = smem + pl.bias[l];
            const bool last = (l == n - 1);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                ld_d8(t, c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = c8 * 8 + i;
                    const float z = v[i] + bias[j];
                    if (valid) Z[((size_t)l * NICER_W + j) * Ps + p] = z;
                    const SpEval sp = sp_eval(z);
                    v[i] = sp.a;
                    if (last) s_out += wl[j] * sp.a;
                }
                st_a8(t, c8, v);
            }
        }
        if (valid) { if (accumulate) sdf[p] += s_out; else sdf[p] = s_out; }
        // ---------------- feature head
        if (want_feat) {
            mat_gemm(t, pl, n, smem);
            const float *bias = smem + pl.bias[n];
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                ld_d8(t, c8, v);
                tc::wait_ld();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = c8 * 8 + i;
                        float *dst = feat_fm + (size_t)j * Ps + p;
                        const float f = v[i] + bias[j];
                        if (accumulate) *dst += f; else *dst = f;
                    }
                }
            }
        }
    }
    tile_teardown(sh);
}

// Kernel B: gradient chain  q_n = W_n[0,:] * sp'(z_n),  r_{l-1} = W_{l-1}^T q_l,  q_l = r_l * sp'(z_l),  g = J^T r_0
// from the saved pre-activations Z; saves R (adjoints r_l, l < n) and emits d sdf/dx.
template <int C>
__global__ void __launch_bounds__(TCF_THREADS, 1)
sdf_forward_tc_b_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                        uint32_t P, uint32_t flags, float *grad, const float *Z, float *R, const float *DYDX) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcfShared sh;
    LevelInfo *lv;
    Tile t = tcf_setup(net, ls, pl, smem, sh, lv);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const bool accumulate = (flags & NICER_SDF_ACCUMULATE) != 0;
    const float *wl = smem + pl.wl_sdf;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + (threadIdx.x >> 7); tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (threadIdx.x & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;
        // q_n -> A
        {
            float zv[NICER_W];
            load64(Z, (size_t)(n - 1) * NICER_W, Ps, p, zv);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = wl[c8 * 8 + i] * dsoftplus100(zv[c8 * 8 + i]);
                st_a8(t, c8, v);
            }
        }
        for (int l = n - 1; l >= 1; --l) {
            mat_issue(t, pl, l, smem);          // r_l = W_l^T q_{l+1}
            float zv[NICER_W];
            load64(Z, (size_t)(l - 1) * NICER_W, Ps, p, zv);
            gemm_wait(t);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                ld_d8(t, c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const size_t o = ((size_t)(l - 1) * NICER_W + c8 * 8 + i) * Ps + p;
                    if (valid) R[o] = v[i];
                    v[i] *= dsoftplus100(zv[c8 * 8 + i]);
                }
                st_a8(t, c8, v);
            }
        }
        mat_issue(t, pl, 0, smem);              // r_0 = W_0^T q_1   (80 columns: [32 grid | 39 PE | pad])
        const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
        float dyv[96];
#pragma unroll
        for (int k = 0; k < 96; ++k) dyv[k] = (k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
        gemm_wait(t);
        float g[3];
        {
            float rp[40];   // PE part: columns 32..71
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &rp[c8 * 8]);
            tc::wait_ld();
            g[0] = rp[0]; g[1] = rp[1]; g[2] = rp[2];
            float fr = 1.0f;
#pragma unroll
            for (int f = 0; f < 6; ++f) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float s, c;
                    sincosf(x[d] * fr, &s, &c);
                    g[d] += fr * (c * rp[3 + 6 * f + d] - s * rp[3 + 6 * f + 3 + d]);
                }
                fr *= 2.0f;
            }
            float gu[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float rh[8];    // grid part: columns 0..31
                ld_d8(t, c8, rh);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = c8 * 8 + i;
                    if (k < L * C) {
                        const int l = k / C, c = k % C;
#pragma unroll
                        for (int d = 0; d < 3; ++d) gu[d] += rh[i] * dyv[(l * 3 + d) * C + c];
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] += gu[d] / 2.0f / df;
        }
        if (valid) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (accumulate) grad[3 * (size_t)p + d] += g[d]; else grad[3 * (size_t)p + d] = g[d];
            }
        }
    }
    tile_teardown(sh);
}

int launch_grid_encode(const nicer_grid_t *g, const float *x, uint32_t P, float *F, float *DYDX, cudaStream_t st);

// two-threads-per-point variants (sdf_tc_split.cu), selected per kernel by tc_split_mask() (1 = A, 2 = B, 4 = T, 8 = R)
unsigned tc_split_mask();
int launch_tcs_a(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, uint32_t flags, float *sdf, float *feat_fm, float *Z,
                 float *DYDX, float *H0, cudaStream_t st);
int launch_tcs_b(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t flags, float *grad, const float *Z, float *R,
                 const float *DYDX, const float *H0, cudaStream_t st);
int launch_tcs_t(const nicer_sdf_net_t *net, const float *x, uint32_t P, const float *Z, const float *R, const float *DYDX,
                 const float *H0, const float *g_grad, float *ZB, float *QB, float *AB, float *TAN, float *T0, float *tan_sum,
                 cudaStream_t st);
int launch_row_sum_accum(const float *rows, uint32_t n_rows, uint32_t P, float *out, cudaStream_t st);
int launch_tcs_r(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, const float *Z, const float *DYDX, const float *H0,
                 const float *g_sdf, const float *g_feat_fm, const float *g_grad, float *grad_x, float *ZB, const float *QB, float *GY,
                 cudaStream_t st);

int launch_sdf_forward_tc(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, uint32_t flags, float *sdf, float *feat_fm,
                          float *grad, float *Z, float *R, float *DYDX, float *H0, cudaStream_t st) {
    if (Pf != P && !(tc_split_mask() & 1u)) NICER_FAIL(-1, "nicer_sdf_forward: P_feat < P needs the two-threads-per-point kernels (NICER_TC_SPLIT)");
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const uint32_t pairs = div_up(div_up(P, 128), 2);
    const uint32_t grid = pairs < (uint32_t)num_sms() ? pairs : (uint32_t)num_sms();
    const TcfPlan pa = plan_a(net), pb = plan_b(net);
    const size_t smem_a = (size_t)pa.total_floats * sizeof(float), smem_b = (size_t)pb.total_floats * sizeof(float);
    if (H0) {       // gathers at full occupancy, into the grid rows of the saved network input
        if (int e = launch_grid_encode(&net->grid, x, P, H0 + (size_t)39 * P, DYDX, st)) return e;
    }
    const unsigned split = tc_split_mask();
    // kernel B reads the positional encoding back from the saved input (NICER_PE_RELOAD=0: recomputes it)
    static const bool pe_reload = [] { const char *e = getenv("NICER_PE_RELOAD"); return !(e && e[0] == '0'); }();
    const float *pe_h0 = pe_reload ? H0 : nullptr;
#define LAUNCH(CC)                                                                                                      \
    do {                                                                                                                \
        NICER_CUDA(cudaFuncSetAttribute(sdf_forward_tc_a_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a), \
                   "nicer_sdf_forward(tc A)");                                                                          \
        NICER_CUDA(cudaFuncSetAttribute(sdf_forward_tc_b_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b), \
                   "nicer_sdf_forward(tc B)");                                                                          \
        if (split & 1u) { if (int e = launch_tcs_a(net, x, P, Pf, flags, sdf, feat_fm, Z, DYDX, H0, st)) return e; }      \
        else sdf_forward_tc_a_kernel<CC><<<grid, TCF_THREADS, smem_a, st>>>(*net, ls, pa, x, P, flags, sdf, feat_fm, Z, DYDX, H0); \
        if (split & 2u) { if (int e = launch_tcs_b(net, x, P, flags, grad, Z, R, DYDX, pe_h0, st)) return e; }                     \
        else sdf_forward_tc_b_kernel<CC><<<grid, TCF_THREADS, smem_b, st>>>(*net, ls, pb, x, P, flags, grad, Z, R, DYDX);   \
    } while (0)
    switch (net->grid.C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_sdf_forward(tc)");
    return 0;
}


// ================================================================================================ backward
// Kernel T: tangent pass (forward-mode derivative of the network in direction g_bar = dL/d(d sdf/dx)):
//   t_0 = J g_bar, u_1 = W_0 t_0, tan_l = u_l * sp'(z_l), u_{l+1} = W_l tan_l
// and the per-layer buffers the weight-gradient GEMMs and kernel R need: TAN, QB = r_l sp'(z_l), AB = a_l,
// ZB <- u_l r_l sp''(z_l) (the second-order part of dL/dz_l) and T0.
template <int C>
__global__ void __launch_bounds__(TCF_THREADS, 1)
sdf_backward_tc_t_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                         uint32_t P, const float *Z, const float *R, const float *DYDX, const float *g_grad, float *ZB,
                         float *QB, float *AB, float *TAN, float *T0) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcfShared sh;
    LevelInfo *lv;
    Tile t = tcf_setup(net, ls, pl, smem, sh, lv);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const float *wl = smem + pl.wl_sdf;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + (threadIdx.x >> 7); tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (threadIdx.x & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;
        const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
        float gg[3] = {0.f, 0.f, 0.f}, ggu[3];
        if (g_grad) { gg[0] = g_grad[3 * (size_t)p]; gg[1] = g_grad[3 * (size_t)p + 1]; gg[2] = g_grad[3 * (size_t)p + 2]; }
#pragma unroll
        for (int d = 0; d < 3; ++d) ggu[d] = gg[d] / 2.0f / df;
        // ---- t_0: PE part (columns 32..70), rows 0..38 of T0 / H0
        {
            float tp[48];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                tp[d] = gg[d];
                if (valid) T0[(size_t)d * Ps + p] = gg[d];
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float sc[12];
                pe_sincos<6, true>(x[d], sc);
                float fr = 1.0f;
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    const int ks = 3 + 6 * f + d, kc = ks + 3;
                    const float ts = fr * sc[2 * f + 1] * gg[d], tcv = -fr * sc[2 * f] * gg[d];
                    tp[ks] = ts; tp[kc] = tcv;
                    if (valid) { T0[(size_t)ks * Ps + p] = ts; T0[(size_t)kc * Ps + p] = tcv; }
                    fr *= 2.0f;
                }
            }
#pragma unroll
            for (int k = 39; k < 48; ++k) tp[k] = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 6; ++c8) st_a8(t, 4 + c8, &tp[c8 * 8]);
        }
        // ---- t_0: grid part (columns 0..31), rows 39.. of T0
        {
            float dyv[96];
#pragma unroll
            for (int k = 0; k < 96; ++k) dyv[k] = (k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
            float tv[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int l = k / C, c = k % C;
                tv[k] = (k < L * C) ? ggu[0] * dyv[(l * 3 + 0) * C + c] + ggu[1] * dyv[(l * 3 + 1) * C + c] + ggu[2] * dyv[(l * 3 + 2) * C + c] : 0.f;
                if (valid && k < L * C) T0[(size_t)(39 + k) * Ps + p] = tv[k];
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) st_a8(t, c8, &tv[c8 * 8]);
        }
        mat_issue(t, pl, 0, smem);   // u_1 = W_0 t_0
        for (int l = 1; l <= n; ++l) {
            float zv[NICER_W], rv[NICER_W];
            load64(Z, (size_t)(l - 1) * NICER_W, Ps, p, zv);
            if (l < n) {
                load64(R, (size_t)(l - 1) * NICER_W, Ps, p, rv);
            } else {
#pragma unroll
                for (int j = 0; j < NICER_W; ++j) rv[j] = wl[j];
            }
            gemm_wait(t);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                ld_d8(t, c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = c8 * 8 + i;
                    const size_t o = ((size_t)(l - 1) * NICER_W + j) * Ps + p;
                    const SpEval sp = sp_eval(zv[j]);
                    const float u = v[i];
                    const float tan = u * sp.s1;
                    if (valid) {
                        TAN[o] = tan;
                        QB[o] = rv[j] * sp.s1;
                        AB[o] = sp.a;
                        ZB[o] = u * rv[j] * sp.s2;
                    }
                    v[i] = tan;
                }
                if (l < n) st_a8(t, c8, v);
            }
            if (l < n) mat_issue(t, pl, l, smem);   // u_{l+1} = W_l tan_l
        }
    }
    tile_teardown(sh);
}

// Kernel R: reverse pass.  abar_n = W_n^T [g_sdf, g_feat],  zbar_l = abar_l sp'(z_l) + ZB_l (second-order part from
// kernel T),  abar_{l-1} = W_{l-1}^T zbar_l,  hbar_0 = W_0^T zbar_1,  r_0 = W_0^T q_1 (for the second-order terms);
// writes zbar_l into ZB, scatters first- and second-order grid gradients, accumulates dL/dx.
template <int C>
__global__ void __launch_bounds__(TCF_THREADS, 1)
sdf_backward_tc_r_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                         uint32_t P, const float *Z, const float *DYDX, const float *g_sdf, const float *g_feat_fm,
                         const float *g_grad, float *grad_x, float *ZB, const float *QB, float *GY) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcfShared sh;
    LevelInfo *lv;
    Tile t = tcf_setup(net, ls, pl, smem, sh, lv);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const float *wl = smem + pl.wl_sdf;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + (threadIdx.x >> 7); tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (threadIdx.x & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;
        const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
        float u[3];
        to_unit(x, df, u);
        const float gs = g_sdf ? g_sdf[p] : 0.f;
        float gg[3] = {0.f, 0.f, 0.f}, ggu[3];
        if (g_grad) { gg[0] = g_grad[3 * (size_t)p]; gg[1] = g_grad[3 * (size_t)p + 1]; gg[2] = g_grad[3 * (size_t)p + 2]; }
#pragma unroll
        for (int d = 0; d < 3; ++d) ggu[d] = gg[d] / 2.0f / df;
        // ---- A = g_feat -> abar_n' = (W_n[1:])^T g_feat
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = g_feat_fm ? g_feat_fm[(size_t)(c8 * 8 + i) * Ps + p] : 0.f;
            st_a8(t, c8, v);
        }
        mat_issue(t, pl, n, smem);
        for (int l = n; l >= 1; --l) {
            float zv[NICER_W], cv[NICER_W];
            load64(Z, (size_t)(l - 1) * NICER_W, Ps, p, zv);
            load64_rw(ZB, (size_t)(l - 1) * NICER_W, Ps, p, cv);
            gemm_wait(t);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                ld_d8(t, c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = c8 * 8 + i;
                    const size_t o = ((size_t)(l - 1) * NICER_W + k) * Ps + p;
                    const float abar = (l == n) ? v[i] + wl[k] * gs : v[i];
                    const float zb = abar * dsoftplus100(zv[k]) + cv[k];
                    if (valid) ZB[o] = zb;
                    v[i] = zb;
                }
                st_a8(t, c8, v);
            }
            mat_issue(t, pl, l - 1, smem);   // abar_{l-1} = W_{l-1}^T zbar_l   (l == 1: hbar_0, 80 columns)
        }
        float q1[NICER_W], dyv[96];
        load64(QB, 0, Ps, p, q1);
#pragma unroll
        for (int k = 0; k < 96; ++k) dyv[k] = (k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
        gemm_wait(t);
        // ---- hbar_0: PE part -> dL/dx, grid part kept for the scatter
        float xb[3];
        float gy1[32];
        float pesc[36];     // sin/cos of the PE, reused for the second-order term below
        {
            float hp[40];
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &hp[c8 * 8]);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) ld_d8(t, c8, &gy1[c8 * 8]);
            tc::wait_ld();
            xb[0] = hp[0]; xb[1] = hp[1]; xb[2] = hp[2];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pe_sincos<6, true>(x[d], &pesc[12 * d]);
                float fr = 1.0f;
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    xb[d] += fr * (pesc[12 * d + 2 * f + 1] * hp[3 + 6 * f + d] - pesc[12 * d + 2 * f] * hp[3 + 6 * f + 3 + d]);
                    fr *= 2.0f;
                }
            }
        }
        // ---- r_0 = W_0^T q_1 (second-order terms)
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) st_a8(t, c8, &q1[c8 * 8]);
        mat_gemm(t, pl, 0, smem);
        {
            float rp[40];
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &rp[c8 * 8]);
            tc::wait_ld();
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float fr = 1.0f;
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    xb[d] += gg[d] * (fr * fr) * (-pesc[12 * d + 2 * f] * rp[3 + 6 * f + d] - pesc[12 * d + 2 * f + 1] * rp[3 + 6 * f + 3 + d]);
                    fr *= 2.0f;
                }
            }
        }
        float xu[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int l = 0; l < 32 / C; ++l) {
            if (l < L) {
                float gy2[C];
                ld_d_small<C>(t, l * C, gy2);
                tc::wait_ld();
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int d = 0; d < 3; ++d) xu[d] += gy1[l * C + c] * dyv[(l * 3 + d) * C + c];
                // grid gradients go to the scatter kernel (grid_scatter.cu): first-order rows, then second-order rows
                if (valid) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        GY[(size_t)(l * C + c) * Ps + p] = gy1[l * C + c];
                        GY[(size_t)((L + l) * C + c) * Ps + p] = gy2[c];
                    }
                }
            }
        }
        if (grad_x && valid) {
#pragma unroll
            for (int d = 0; d < 3; ++d) grad_x[3 * (size_t)p + d] += xb[d] + xu[d] / 2.0f / df;
        }
    }
    tile_teardown(sh);
}

int launch_grid_scatter(const nicer_grid_t *g, const float *x, uint32_t P, const float *GY1, const float *GY2,
                        const float *g_grad, float *grad_table, cudaStream_t st);

int launch_sdf_backward_tc(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, const float *Z, const float *R,
                           const float *DYDX, const float *H0, const float *g_sdf, const float *g_feat_fm, const float *g_grad, float *grad_x,
                           float *grad_table, float *ZB, float *QB, float *AB, float *TAN, float *T0, float *tan_sum, float *GY,
                           cudaStream_t st, cudaStream_t scatter_st) {
    const unsigned split = tc_split_mask();
    if (Pf != P && !(split & 8u)) NICER_FAIL(-1, "nicer_sdf_backward: P_feat < P needs the two-threads-per-point kernels (NICER_TC_SPLIT)");
    static const bool pe_reload = [] { const char *e = getenv("NICER_PE_RELOAD"); return !(e && e[0] == '0'); }();
    const float *pe_h0 = pe_reload ? H0 : nullptr;
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const uint32_t pairs = div_up(div_up(P, 128), 2);
    const uint32_t grid = pairs < (uint32_t)num_sms() ? pairs : (uint32_t)num_sms();
    const TcfPlan pt = plan_t(net), pr = plan_r(net);
    const size_t smem_t = (size_t)pt.total_floats * sizeof(float), smem_r = (size_t)pr.total_floats * sizeof(float);
#define LAUNCH(CC)                                                                                                       \
    do {                                                                                                                 \
        NICER_CUDA(cudaFuncSetAttribute(sdf_backward_tc_t_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t), \
                   "nicer_sdf_backward(tc T)");                                                                          \
        NICER_CUDA(cudaFuncSetAttribute(sdf_backward_tc_r_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r), \
                   "nicer_sdf_backward(tc R)");                                                                          \
        if (split & 4u) { if (int e = launch_tcs_t(net, x, P, Z, R, DYDX, pe_h0, g_grad, ZB, QB, AB, TAN, T0, tan_sum, st)) return e; } \
        else sdf_backward_tc_t_kernel<CC><<<grid, TCF_THREADS, smem_t, st>>>(*net, ls, pt, x, P, Z, R, DYDX, g_grad, ZB, QB, AB, TAN, T0); \
        if (split & 8u) { if (int e = launch_tcs_r(net, x, P, Pf, Z, DYDX, pe_h0, g_sdf, g_feat_fm, g_grad, grad_x, ZB, QB, GY, st)) return e; } \
        else sdf_backward_tc_r_kernel<CC><<<grid, TCF_THREADS, smem_r, st>>>(*net, ls, pr, x, P, Z, DYDX, g_sdf, g_feat_fm, g_grad, grad_x, \
                                                                        ZB, QB, GY);                                     \
    } while (0)
    switch (net->grid.C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_sdf_backward(tc)");
    if (tan_sum && !(split & 4u)) {      // the one-thread-per-point tangent kernel does not sum tan_n itself
        if (int e = launch_row_sum_accum(TAN + (size_t)(net->n_hidden - 1) * NICER_W * P, NICER_W, P, tan_sum, st)) return e;
    }
    if (scatter_st && scatter_st != st) {
        if (int e = stream_fork(st, scatter_st)) return e;
    } else {
        scatter_st = st;
    }
    return launch_grid_scatter(&net->grid, x, P, GY, g_grad ? GY + (size_t)net->grid.L * net->grid.C * P : nullptr, g_grad,
                               grad_table, scatter_st);
}

}  // namespace nicer
