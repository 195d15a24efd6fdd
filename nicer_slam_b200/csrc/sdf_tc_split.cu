// Tensor-core (tcgen05, sm_100a) SDF network main pass, TWO THREADS PER POINT.
//
// Same four kernels, same math, same saved tensors and the same staged operands as sdf_tc_full.cu (A: layers + feature
// head, B: gradient chain, T: tangent pass, R: reverse pass; plans in sdf_tc_plan.cuh).  What changes is the mapping of
// the per-point epilogue work: a 128-point tile is served by 256 threads, thread pair (h = 0, 1) of a point sharing its TMEM
// lane -- warp w and warp w + 4 of a tile address the same lane quarter -- and splitting the 64 accumulator columns in
// halves [32 h, 32 h + 32).  The 80-column layer-0 operand [32 grid | 39 PE | pad] splits the same way: h = 0 owns the grid
// part (features, d feat/dx rows, grid-gradient hand-over), h = 1 the x / positional-encoding part.
//
// Why: with one thread per point the kernels held 64-wide register arrays (254-255 registers, 8 warps per SM) and ncu
// showed them waiting on their own loads (long scoreboard 48 %, tensor pipe 6-9 %, issue slots 20-33 % busy;
// profiles/r01_ncu_sdf_backward_summary.txt).  With half-width arrays a CTA runs 16 warps at 128 registers and the per-tile
// dependent chain between two MMA groups is half as long.  What matters at 128 registers is that NO loaded value is spilled
// (a spill store waits for its load, which serialises the prefetch): the epilogues therefore work in 8-column chunks, the saved
// rows of a layer are prefetched under its MMAs (kernel T: two chunks ahead of their use), and wide per-point arrays (the 96
// d feat/dx rows) are consumed level by level.  The few per-point scalars that need both halves (sdf = w_n . a_n, d sdf/dx,
// dL/dx) cross through a 2 KB shared-memory hand-over, ordered by the tile's named barrier.  Measured steps: DESIGN.md 4.
//
// Batched second point set: the points behind the first Pf (eikonal samples riding behind the main-pass points) skip the
// feature head (kernel A) and have no upstream sdf / feature gradient (kernel R); everything else treats them like any point.
#include "sdf_tc_plan.cuh"
#include "tc_tile2.cuh"

namespace nicer {

__device__ __forceinline__ void mat_issue2(Tile &t, const TcfPlan &pl, int i, float *smem) {
    gemm_issue2(t, tc::smem_u32(smem + pl.m[i].hi), tc::smem_u32(smem + pl.m[i].lo), pl.m[i].K, pl.m[i].rows);
}
__device__ __forceinline__ void mat_gemm2(Tile &t, const TcfPlan &pl, int i, float *smem) {
    mat_issue2(t, pl, i, smem);
    gemm_wait(t);
}

// ------------------------------------------------------------------------------------------------ kernel A
template <int C>
__global__ void __launch_bounds__(TCS_THREADS, 1)
sdf_forward_tcs_a_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                         uint32_t P, uint32_t Pf, uint32_t flags, float *sdf, float *feat_fm, float *Z, float *DYDX, float *H0) {
    // H0 != NULL: its grid rows (and DYDX) were already written by grid_encode_kernel; this kernel adds the x / PE rows
    // Pf <= P: only the first Pf points get features (feat_fm is [64][Pf]); tiles beyond skip the feature head (eikonal points
    // batched behind the main-pass points)
    extern __shared__ __align__(16) float smem[];
    __shared__ TcsShared sh;
    LevelInfo *lv;
    tcf_stage_all(net, ls, pl, smem, lv);
    Tile t = tile_setup2(sh, TCF_ALO, TCF_D);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const int h = (threadIdx.x >> 7) & 1, c0 = 4 * h, tile = threadIdx.x >> 8, lane = threadIdx.x & 127;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const bool accumulate = (flags & NICER_SDF_ACCUMULATE) != 0;
    const bool want_feat = (flags & NICER_SDF_NO_FEAT) == 0;
    const float *wl = smem + pl.wl_sdf;
    const float bl_sdf = net.b[n][0];

    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;
        // ---------------- network input -> A   (columns: [32 grid | 39 PE | pad])
        if (h == 1) {
            const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
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
        } else if (H0) {
            // grid features were gathered by grid_encode_kernel into the grid rows of H0 (and DYDX): coalesced reads
            float gf[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) gf[k] = (k < L * C) ? __ldg(H0 + (size_t)(39 + k) * Ps + p) : 0.f;
            st_half(t, 0, gf);
        } else {
            const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
            float u[3];
            to_unit(x, df, u);
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
        float s_part = (h == 0) ? bl_sdf : 0.f;
        for (int l = 0; l < n; ++l) {
            mat_gemm2(t, pl, l, smem);
            const float *bias = smem + pl.bias[l] + c0 * 8;
            const bool last = (l == n - 1);
            float v[32];
            ld_half(t, c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float z = v[i] + bias[i];
                if (valid) Z[((size_t)l * NICER_W + c0 * 8 + i) * Ps + p] = z;
                const SpEval sp = sp_eval(z);
                v[i] = sp.a;
                if (last) s_part += wl[c0 * 8 + i] * sp.a;
            }
            st_half(t, c0, v);
        }
        // ---------------- sdf = b + w_n . a_n: upper half of the dot product handed over through shared memory
        const bool tile_feat = want_feat && tt * 128u < Pf;                   // tile-uniform
        if (h == 1) sh.xch[tile][lane][0] = s_part;
        if (tile_feat) mat_issue2(t, pl, n, smem); else tile_sync2(t);        // both contain the tile barrier
        if (h == 0 && valid) {
            const float s_out = s_part + sh.xch[tile][lane][0];
            if (accumulate) sdf[p] += s_out; else sdf[p] = s_out;
        }
        // ---------------- feature head
        if (tile_feat) {
            gemm_wait(t);
            const float *bias = smem + pl.bias[n] + c0 * 8;
            float v[32];
            ld_half(t, c0, v);
            if (valid && p < Pf) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float *dst = feat_fm + (size_t)(c0 * 8 + i) * Pf + p;
                    const float f = v[i] + bias[i];
                    if (accumulate) *dst += f; else *dst = f;
                }
            }
        }
    }
    tile_teardown2(sh);
}

// ------------------------------------------------------------------------------------------------ kernel B
// gradient chain  q_n = W_n[0,:] * sp'(z_n),  r_{l-1} = W_{l-1}^T q_l,  q_l = r_l * sp'(z_l),  g = J^T r_0
template <int C>
__global__ void __launch_bounds__(TCS_THREADS, 1)
sdf_forward_tcs_b_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                         uint32_t P, uint32_t flags, float *grad, const float *Z, float *R, const float *DYDX,
                         const float *__restrict__ H0) {
    // H0 != NULL: rows 3..38 hold sin / cos (2^f x_d) as kernel A wrote them; reading them back replaces 18 sincosf per point
    extern __shared__ __align__(16) float smem[];
    __shared__ TcsShared sh;
    LevelInfo *lv;
    tcf_stage_all(net, ls, pl, smem, lv);
    Tile t = tile_setup2(sh, TCF_ALO, TCF_D);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const int h = (threadIdx.x >> 7) & 1, c0 = 4 * h, tile = threadIdx.x >> 8, lane = threadIdx.x & 127;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const bool accumulate = (flags & NICER_SDF_ACCUMULATE) != 0;
    const float *wl = smem + pl.wl_sdf + c0 * 8;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;
        {   // q_n -> A
            float v[32];
            load32(Z, (size_t)(n - 1) * NICER_W + c0 * 8, Ps, p, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = wl[i] * dsoftplus100(v[i]);
            st_half(t, c0, v);
        }
        for (int l = n - 1; l >= 1; --l) {
            mat_issue2(t, pl, l, smem);          // r_l = W_l^T q_{l+1}
            float zv[32];
            load32(Z, (size_t)(l - 1) * NICER_W + c0 * 8, Ps, p, zv);
            gemm_wait(t);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {        // 8 columns at a time: the prefetched rows + one chunk fit 128 registers
                float v[8];
                ld_d8(t, c0 + c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (valid) R[((size_t)(l - 1) * NICER_W + (c0 + c8) * 8 + i) * Ps + p] = v[i];
                    v[i] *= dsoftplus100(zv[c8 * 8 + i]);
                }
                st_a8(t, c0 + c8, v);
            }
        }
        mat_issue2(t, pl, 0, smem);              // r_0 = W_0^T q_1   (80 columns: [32 grid | 39 PE | pad])
        if (h == 1) {
            float sc[36];   // sin / cos (2^f x_d) in the order of the input vector: [6 f + d] sin, [6 f + 3 + d] cos
            if (H0) {
#pragma unroll
                for (int k = 0; k < 36; ++k) sc[k] = __ldg(H0 + (size_t)(3 + k) * Ps + p);
            } else {
                const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
                float fr = 1.0f;
#pragma unroll
                for (int f = 0; f < 6; ++f) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) sincosf(x[d] * fr, &sc[6 * f + d], &sc[6 * f + 3 + d]);
                    fr *= 2.0f;
                }
            }
            gemm_wait(t);
            float rp[40];   // PE part: columns 32..71
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &rp[c8 * 8]);
            tc::wait_ld();
            float g[3] = {rp[0], rp[1], rp[2]};
            float fr = 1.0f;
#pragma unroll
            for (int f = 0; f < 6; ++f) {
#pragma unroll
                for (int d = 0; d < 3; ++d) g[d] += fr * (sc[6 * f + 3 + d] * rp[3 + 6 * f + d] - sc[6 * f + d] * rp[3 + 6 * f + 3 + d]);
                fr *= 2.0f;
            }
            tile_sync2(t);                       // grid part of d sdf/dx from the other half
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float gd = g[d] + sh.xch[tile][lane][d];
                    if (accumulate) grad[3 * (size_t)p + d] += gd; else grad[3 * (size_t)p + d] = gd;
                }
            }
        } else {
            float dyv[96];
#pragma unroll
            for (int k = 0; k < 96; ++k) dyv[k] = (k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
            gemm_wait(t);
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
            for (int d = 0; d < 3; ++d) sh.xch[tile][lane][d] = gu[d] / 2.0f / df;
            tile_sync2(t);
        }
    }
    tile_teardown2(sh);
}

// Sums of 8 per-lane values over the 32 lanes of a warp with 9 shuffles (halving exchange, then two butterfly steps): every
// lane returns the warp total of column `col` (lanes that differ only in their two low bits hold the same column).
__device__ __forceinline__ float warp_colsum8(const float v[8], int lane, int &col) {
    const bool u4 = (lane & 16) != 0, u2 = (lane & 8) != 0, u1 = (lane & 4) != 0;
    float a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (u4 ? v[4 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, u4 ? v[i] : v[4 + i], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) b[i] = (u2 ? a[2 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, u2 ? a[i] : a[2 + i], 8);
    float c = (u1 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, u1 ? b[0] : b[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    col = (u4 ? 4 : 0) + (u2 ? 2 : 0) + (u1 ? 1 : 0);
    return c;
}

// ------------------------------------------------------------------------------------------------ kernel T
// tangent pass: t_0 = J g_bar, u_1 = W_0 t_0, tan_l = u_l * sp'(z_l), u_{l+1} = W_l tan_l; writes TAN, QB, AB, ZB, T0
template <int C>
__global__ void __launch_bounds__(TCS_THREADS, 1)
sdf_backward_tcs_t_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                          uint32_t P, const float *Z, const float *R, const float *DYDX, const float *__restrict__ H0,
                          const float *g_grad, float *ZB, float *QB, float *AB, float *TAN, float *T0, float *tan_sum) {
    // H0 != NULL: sin / cos of the positional encoding are read back from the saved input (rows 3..38) instead of recomputed
    // tan_sum != NULL: [64] += sum over the points of tan_n (the second-order part of dL/dW_n[0,:], i.e. of the sdf row of the
    //                  last layer): column sums per warp by shuffles, per CTA in shared memory, one atomic per column at the end
    extern __shared__ __align__(16) float smem[];
    __shared__ TcsShared sh;
    __shared__ float colsum_s[NICER_W];
    if (threadIdx.x < NICER_W) colsum_s[threadIdx.x] = 0.f;
    LevelInfo *lv;
    tcf_stage_all(net, ls, pl, smem, lv);
    Tile t = tile_setup2(sh, TCF_ALO, TCF_D);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const int h = (threadIdx.x >> 7) & 1, c0 = 4 * h, tile = threadIdx.x >> 8, lane = threadIdx.x & 127;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const float *wl = smem + pl.wl_sdf + c0 * 8;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;
        float gg[3] = {0.f, 0.f, 0.f};
        if (g_grad) { gg[0] = g_grad[3 * (size_t)p]; gg[1] = g_grad[3 * (size_t)p + 1]; gg[2] = g_grad[3 * (size_t)p + 2]; }
        if (h == 1) {
            // ---- t_0: PE part (columns 32..70), rows 0..38 of T0
            float tp[48];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                tp[d] = gg[d];
                if (valid) T0[(size_t)d * Ps + p] = gg[d];
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float sc[12];
                if (H0) {
#pragma unroll
                    for (int f = 0; f < 6; ++f) {
                        sc[2 * f] = __ldg(H0 + (size_t)(3 + 6 * f + d) * Ps + p);
                        sc[2 * f + 1] = __ldg(H0 + (size_t)(3 + 6 * f + 3 + d) * Ps + p);
                    }
                } else {
                    pe_sincos<6, true>(X[3 * (size_t)p + d], sc);
                }
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
        } else {
            // ---- t_0: grid part (columns 0..31), rows 39.. of T0
            float ggu[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) ggu[d] = gg[d] / 2.0f / df;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float dy[3][8], tv[8];      // the d feat/dx rows of these 8 features (24 loads in flight per chunk)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = c8 * 8 + i, l = k / C, c = k % C;
#pragma unroll
                    for (int d = 0; d < 3; ++d) dy[d][i] = (k < L * C) ? __ldg(DYDX + (size_t)((l * 3 + d) * C + c) * Ps + p) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = c8 * 8 + i;
                    tv[i] = (k < L * C) ? ggu[0] * dy[0][i] + ggu[1] * dy[1][i] + ggu[2] * dy[2][i] : 0.f;
                    if (valid && k < L * C) T0[(size_t)(39 + k) * Ps + p] = tv[i];
                }
                st_a8(t, c8, tv);
            }
        }
        mat_issue2(t, pl, 0, smem);   // u_1 = W_0 t_0
        for (int l = 1; l <= n; ++l) {
            // z_l and r_l rows in 8-column chunks, two chunks ahead of their use (the first two are issued under the MMAs): with
            // both sets of 32 prefetched plus the four output streams ptxas spilled the loaded values, i.e. waited for every load
            float zq[4][8], rq[4][8];
            const size_t row0 = (size_t)(l - 1) * NICER_W + c0 * 8;
#define T_LOAD(CH)                                                                                                     \
    do {                                                                                                               \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) {                                                                \
            zq[CH][i] = __ldg(Z + (row0 + (CH) * 8 + i) * Ps + p);                                                     \
            rq[CH][i] = (l < n) ? __ldg(R + (row0 + (CH) * 8 + i) * Ps + p) : wl[(CH) * 8 + i];                        \
        }                                                                                                              \
    } while (0)
            T_LOAD(0);
            T_LOAD(1);
            gemm_wait(t);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float v[8];
                if (c8 == 0) T_LOAD(2);
                if (c8 == 1) T_LOAD(3);
                ld_d8(t, c0 + c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const size_t o = (row0 + c8 * 8 + i) * Ps + p;
                    const SpEval sp = sp_eval(zq[c8][i]);
                    const float u = v[i];
                    const float tan = u * sp.s1;
                    if (valid) {
                        TAN[o] = tan;
                        QB[o] = rq[c8][i] * sp.s1;
                        AB[o] = sp.a;
                        ZB[o] = u * rq[c8][i] * sp.s2;
                    }
                    v[i] = tan;
                }
                if (l < n) {
                    st_a8(t, c0 + c8, v);
                } else if (tan_sum) {
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    }
                    int col;
                    const float cs = warp_colsum8(v, threadIdx.x & 31, col);
                    if ((threadIdx.x & 3) == 0) atomicAdd(&colsum_s[(c0 + c8) * 8 + col], cs);
                }
            }
#undef T_LOAD
            if (l < n) mat_issue2(t, pl, l, smem);   // u_{l+1} = W_l tan_l
        }
    }
    tile_teardown2(sh);
    if (tan_sum && threadIdx.x < NICER_W) atomicAdd(tan_sum + threadIdx.x, colsum_s[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ kernel R
// reverse pass.  abar_n = W_n^T [g_sdf, g_feat],  zbar_l = abar_l sp'(z_l) + ZB_l,  abar_{l-1} = W_{l-1}^T zbar_l,
// hbar_0 = W_0^T zbar_1,  r_0 = W_0^T q_1 (second-order terms); writes zbar_l into ZB, the grid gradients into GY, dL/dx.
template <int C>
__global__ void __launch_bounds__(TCS_THREADS, 1)
sdf_backward_tcs_r_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcfPlan pl, const float *__restrict__ X,
                          uint32_t P, uint32_t Pf, const float *Z, const float *DYDX, const float *__restrict__ H0, const float *g_sdf,
                          const float *g_feat_fm, const float *g_grad, float *grad_x, float *ZB, const float *QB, float *GY) {
    // Pf <= P: g_sdf [Pf] and g_feat_fm [64][Pf] cover the first Pf points; the upstream gradient of the others is zero
    extern __shared__ __align__(16) float smem[];
    __shared__ TcsShared sh;
    LevelInfo *lv;
    tcf_stage_all(net, ls, pl, smem, lv);
    Tile t = tile_setup2(sh, TCF_ALO, TCF_D);
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    const int h = (threadIdx.x >> 7) & 1, c0 = 4 * h, tile = threadIdx.x >> 8, lane = threadIdx.x & 127;
    const size_t Ps = P;
    const float df = net.grid.divide_factor;
    const float *wl = smem + pl.wl_sdf + c0 * 8;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;
        const bool has_up = p < Pf;
        const float gs = (g_sdf && has_up) ? g_sdf[p] : 0.f;
        // ---- A = g_feat -> abar_n' = (W_n[1:])^T g_feat
        {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (g_feat_fm && has_up) ? g_feat_fm[(size_t)(c0 * 8 + i) * Pf + p] : 0.f;
            st_half(t, c0, v);
        }
        mat_issue2(t, pl, n, smem);
        for (int l = n; l >= 1; --l) {
            float zv[32], cv[32];
            const size_t row0 = (size_t)(l - 1) * NICER_W + c0 * 8;
            load32(Z, row0, Ps, p, zv);
            load32_rw(ZB, row0, Ps, p, cv);
            gemm_wait(t);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {        // 8 columns at a time: the prefetched rows + one chunk fit 128 registers
                float v[8];
                ld_d8(t, c0 + c8, v);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = c8 * 8 + i;
                    const float abar = (l == n) ? v[i] + wl[j] * gs : v[i];
                    const float zb = abar * dsoftplus100(zv[j]) + cv[j];
                    if (valid) ZB[(row0 + j) * Ps + p] = zb;
                    v[i] = zb;
                }
                st_a8(t, c0 + c8, v);
            }
            mat_issue2(t, pl, l - 1, smem);   // abar_{l-1} = W_{l-1}^T zbar_l   (l == 1: hbar_0, 80 columns)
        }
        if (h == 1) {
            float q1[32];
            load32(QB, (size_t)c0 * 8, Ps, p, q1);
            // ---- PE part of hbar_0 and of r_0 -> dL/dx
            float gg[3] = {0.f, 0.f, 0.f};
            if (g_grad) { gg[0] = g_grad[3 * (size_t)p]; gg[1] = g_grad[3 * (size_t)p + 1]; gg[2] = g_grad[3 * (size_t)p + 2]; }
            float pesc[36];     // sin/cos of the PE ([12 d + 2 f] sin, [12 d + 2 f + 1] cos), used for both terms below
            if (H0) {           // read back what kernel A saved (in flight under the MMAs) instead of recomputing
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int f = 0; f < 6; ++f) {
                        pesc[12 * d + 2 * f] = __ldg(H0 + (size_t)(3 + 6 * f + d) * Ps + p);
                        pesc[12 * d + 2 * f + 1] = __ldg(H0 + (size_t)(3 + 6 * f + 3 + d) * Ps + p);
                    }
            }
            gemm_wait(t);
            float xb[3];
            {
                float hp[40];
#pragma unroll
                for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &hp[c8 * 8]);
                tc::wait_ld();
                xb[0] = hp[0]; xb[1] = hp[1]; xb[2] = hp[2];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (!H0) pe_sincos<6, true>(X[3 * (size_t)p + d], &pesc[12 * d]);
                    float fr = 1.0f;
#pragma unroll
                    for (int f = 0; f < 6; ++f) {
                        xb[d] += fr * (pesc[12 * d + 2 * f + 1] * hp[3 + 6 * f + d] - pesc[12 * d + 2 * f] * hp[3 + 6 * f + 3 + d]);
                        fr *= 2.0f;
                    }
                }
            }
            st_half(t, c0, q1);
            mat_gemm2(t, pl, 0, smem);          // r_0 = W_0^T q_1
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
            tile_sync2(t);                      // grid part of dL/dx from the other half
            if (grad_x && valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) grad_x[3 * (size_t)p + d] += xb[d] + sh.xch[tile][lane][d];
            }
        } else {
            // ---- grid part of hbar_0 (first-order grid gradient, dL/dx through d feat/dx) and of r_0 (second-order)
            float dyv[96];
#pragma unroll
            for (int k = 0; k < 96; ++k) dyv[k] = (k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
            gemm_wait(t);
            float xu[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                if (l < L) {
                    float gy1[C];
                    ld_d_small<C>(t, l * C, gy1);
                    tc::wait_ld();
#pragma unroll
                    for (int c = 0; c < C; ++c) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) xu[d] += gy1[c] * dyv[(l * 3 + d) * C + c];
                        if (valid) GY[(size_t)(l * C + c) * Ps + p] = gy1[c];
                    }
                }
            }
            {
                float q1[32];
                load32(QB, (size_t)c0 * 8, Ps, p, q1);
                st_half(t, c0, q1);
            }
            mat_gemm2(t, pl, 0, smem);          // r_0 = W_0^T q_1
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                if (l < L) {
                    float gy2[C];
                    ld_d_small<C>(t, l * C, gy2);
                    tc::wait_ld();
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < C; ++c) GY[(size_t)((L + l) * C + c) * Ps + p] = gy2[c];
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) sh.xch[tile][lane][d] = xu[d] / 2.0f / df;
            tile_sync2(t);
        }
    }
    tile_teardown2(sh);
}

// NICER_TC_SPLIT: bit mask of the kernels that run in this two-threads-per-point form (1 = A, 2 = B, 4 = T, 8 = R); the others
// run as the one-thread-per-point kernels of sdf_tc_full.cu.  Default: see tc_split_mask().
unsigned tc_split_mask() {
    static const unsigned m = [] {
        const char *e = getenv("NICER_TC_SPLIT");
        return e ? (unsigned)atoi(e) : 15u;
    }();
    return m;
}

#define TCS_DISPATCH(KERNEL, SMEM, WHAT, ...)                                                                              \
    do {                                                                                                                   \
        switch (net->grid.C) {                                                                                             \
            case 2:                                                                                                        \
                NICER_CUDA(cudaFuncSetAttribute(KERNEL<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)), WHAT); \
                KERNEL<2><<<grid, TCS_THREADS, SMEM, st>>>(__VA_ARGS__);                                                   \
                break;                                                                                                     \
            case 4:                                                                                                        \
                NICER_CUDA(cudaFuncSetAttribute(KERNEL<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)), WHAT); \
                KERNEL<4><<<grid, TCS_THREADS, SMEM, st>>>(__VA_ARGS__);                                                   \
                break;                                                                                                     \
            default:                                                                                                       \
                NICER_CUDA(cudaFuncSetAttribute(KERNEL<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)), WHAT); \
                KERNEL<8><<<grid, TCS_THREADS, SMEM, st>>>(__VA_ARGS__);                                                   \
                break;                                                                                                     \
        }                                                                                                                  \
        NICER_CHECK_LAUNCH(WHAT);                                                                                          \
    } while (0)

static uint32_t tcs_grid(uint32_t P) {
    const uint32_t pairs = div_up(div_up(P, 128), 2);
    return pairs < (uint32_t)num_sms() ? pairs : (uint32_t)num_sms();
}

int launch_tcs_a(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, uint32_t flags, float *sdf, float *feat_fm, float *Z,
                 float *DYDX, float *H0, cudaStream_t st) {
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const uint32_t grid = tcs_grid(P);
    const TcfPlan pl = plan_a(net);
    const size_t smem = (size_t)pl.total_floats * sizeof(float);
    TCS_DISPATCH(sdf_forward_tcs_a_kernel, smem, "nicer_sdf_forward(tcs A)", *net, ls, pl, x, P, Pf, flags, sdf, feat_fm, Z, DYDX, H0);
    return 0;
}

int launch_tcs_b(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t flags, float *grad, const float *Z, float *R,
                 const float *DYDX, const float *H0, cudaStream_t st) {
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const uint32_t grid = tcs_grid(P);
    const TcfPlan pl = plan_b(net);
    const size_t smem = (size_t)pl.total_floats * sizeof(float);
    TCS_DISPATCH(sdf_forward_tcs_b_kernel, smem, "nicer_sdf_forward(tcs B)", *net, ls, pl, x, P, flags, grad, Z, R, DYDX, H0);
    return 0;
}

int launch_tcs_t(const nicer_sdf_net_t *net, const float *x, uint32_t P, const float *Z, const float *R, const float *DYDX,
                 const float *H0, const float *g_grad, float *ZB, float *QB, float *AB, float *TAN, float *T0, float *tan_sum,
                 cudaStream_t st) {
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const uint32_t grid = tcs_grid(P);
    const TcfPlan pl = plan_t(net);
    const size_t smem = (size_t)pl.total_floats * sizeof(float);
    TCS_DISPATCH(sdf_backward_tcs_t_kernel, smem, "nicer_sdf_backward(tcs T)", *net, ls, pl, x, P, Z, R, DYDX, H0, g_grad, ZB, QB, AB, TAN,
                 T0, tan_sum);
    return 0;
}

int launch_tcs_r(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, const float *Z, const float *DYDX, const float *H0,
                 const float *g_sdf, const float *g_feat_fm, const float *g_grad, float *grad_x, float *ZB, const float *QB, float *GY,
                 cudaStream_t st) {
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const uint32_t grid = tcs_grid(P);
    const TcfPlan pl = plan_r(net);
    const size_t smem = (size_t)pl.total_floats * sizeof(float);
    TCS_DISPATCH(sdf_backward_tcs_r_kernel, smem, "nicer_sdf_backward(tcs R)", *net, ls, pl, x, P, Pf, Z, DYDX, H0, g_sdf, g_feat_fm, g_grad,
                 grad_x, ZB, QB, GY);
    return 0;
}

}  // namespace nicer
