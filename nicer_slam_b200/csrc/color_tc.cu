// Tensor-core (tcgen05, 3xTF32) rendering network, forward and backward (mode "idr",
// /root/reference/code/model/base_networks.py:333-392).  Same scheme as sdf_tc_full.cu (tc_tile.cuh).
//
// The 129-wide input [x 3 | PE(view) 27 | normals 3 | feat 64 | grid 32] is fed to layer 0 as two K blocks that reuse
// the same TMEM operand columns and accumulate into one accumulator:
//   block a (K = 72): [grid 32 | x, PE(view), normals 33 | pad 7]      block b (K = 64): the SDF feature vector
// so a tile needs 72+72 operand columns + 64 accumulator columns and two tiles fit the 512 TMEM columns of an SM.
// The 3-wide output layer and its transposed product are fp32 dot products in registers.
#include "color_sample.cuh"
#include "tc_tile.cuh"
#include "tc_tile2.cuh"

namespace nicer {

constexpr int CT_KA = 72;                 // block a, zero padded
constexpr int CT_ALO = 72, CT_D = 144;    // forward tile columns: A hi [0,72), A lo [72,144), D [144,208)
constexpr int CB_ALO = 64, CB_D = 128;    // backward tile columns: A hi [0,64), A lo [64,128), D [128,208) (80 wide)
constexpr int CT_NA = 80;                 // block a as an N (transposed product), padded to a multiple of 16

struct ColorPlan {
    int w_hi[3], w_lo[3];                 // fwd: W0a, W0b, W1      bwd: W1^T, W0a^T, W0b^T
    int b0, b1, wl, lv, total_floats;
};

static ColorPlan color_plan_fwd() {
    ColorPlan p; int o = 0;
    const int sz[3] = {NICER_W * CT_KA, NICER_W * NICER_W, NICER_W * NICER_W};
    for (int i = 0; i < 3; ++i) { p.w_hi[i] = o; o += sz[i]; p.w_lo[i] = o; o += sz[i]; }
    p.b0 = o; o += NICER_W; p.b1 = o; o += NICER_W; p.wl = o; o += 4 * NICER_W;
    p.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS; p.total_floats = o;
    return p;
}
static ColorPlan color_plan_bwd() {
    ColorPlan p; int o = 0;
    const int sz[3] = {NICER_W * NICER_W, CT_NA * NICER_W, NICER_W * NICER_W};
    for (int i = 0; i < 3; ++i) { p.w_hi[i] = o; o += sz[i]; p.w_lo[i] = o; o += sz[i]; }
    p.b0 = o; o += NICER_W; p.b1 = o; o += NICER_W; p.wl = o; o += 4 * NICER_W;
    p.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS; p.total_floats = o;
    return p;
}

// block-a column k -> column of the reference input vector (-1: padding). d_view = 27, feature = 64.
__device__ __forceinline__ int ct_col_a(int k, int n_grid) {
    if (k < 32) return (k < n_grid) ? 97 + k : -1;
    if (k < 65) return k - 32;
    return -1;
}

// generic staging into hi/lo [k/4][rows][4]: B[n][k] = src(n, k) given by the functor
template <class F>
__device__ void ct_stage(int rows, int K, float *hi, float *lo, F src) {
    for (int i = threadIdx.x; i < rows * K; i += blockDim.x) {
        const int n = i / K, k = i - n * K;
        const float w = src(n, k);
        const float h = tc::tf32_hi(w);
        const int dst = ((k >> 2) * rows + n) * 4 + (k & 3);
        hi[dst] = h;
        lo[dst] = w - h;
    }
}

__device__ __forceinline__ void ct_issue(Tile &t, const ColorPlan &pl, int i, int K, int N, float *smem, bool acc_first = false) {
    gemm_issue(t, tc::smem_u32(smem + pl.w_hi[i]), tc::smem_u32(smem + pl.w_lo[i]), K, N, acc_first);
}

// x (3), PE_4(view) (27), normals (3) -> 40 values (33 + zero padding), in the reference order
template <bool DOUBLING>
__device__ __forceinline__ void ct_xvn(const float x[3], const float v[3], const float nrm[3], float out[40]) {
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2];
    out[3] = v[0]; out[4] = v[1]; out[5] = v[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float sc[8];
        pe_sincos<4, DOUBLING>(v[d], sc);
#pragma unroll
        for (int f = 0; f < 4; ++f) { out[6 + 6 * f + d] = sc[2 * f]; out[6 + 6 * f + 3 + d] = sc[2 * f + 1]; }
    }
    out[30] = nrm[0]; out[31] = nrm[1]; out[32] = nrm[2];
#pragma unroll
    for (int k = 33; k < 40; ++k) out[k] = 0.f;
}

template <int C>
__global__ void __launch_bounds__(TCF_THREADS, 1)
color_forward_tc_kernel(const nicer_color_net_t net, const LevelScales ls, const ColorPlan pl, const float *__restrict__ X,
                        const float *__restrict__ V, const float *__restrict__ Nrm, const float *__restrict__ feat_fm, uint32_t P,
                        float *rgb, float *A_fm, float *DYDX, float *H0) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcfShared sh;
    const int tid = threadIdx.x;
    const bool has_grid = net.grid.table != nullptr;
    const int L = has_grid ? (int)net.grid.L : 0;
    const int n_grid = L * C;
    const int d_in = 33 + 64 + n_grid;
    const float *W0 = net.W[0], *W1 = net.W[1];
    ct_stage(NICER_W, CT_KA, smem + pl.w_hi[0], smem + pl.w_lo[0], [&](int n, int k) {
        const int c = ct_col_a(k, n_grid);
        return c >= 0 ? W0[(size_t)n * d_in + c] : 0.f; });
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[1], smem + pl.w_lo[1], [&](int n, int k) { return W0[(size_t)n * d_in + 33 + k]; });
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[2], smem + pl.w_lo[2], [&](int n, int k) { return W1[(size_t)n * NICER_W + k]; });
    for (int i = tid; i < NICER_W; i += TCF_THREADS) { smem[pl.b0 + i] = net.b[0][i]; smem[pl.b1 + i] = net.b[1][i]; }
    for (int i = tid; i < 3 * NICER_W; i += TCF_THREADS) smem[pl.wl + i] = net.W[2][i];
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + pl.lv);
    for (int l = tid; l < L; l += TCF_THREADS) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    Tile t = tile_setup(sh, CT_ALO, CT_D);
    const size_t Ps = P;
    const float df = has_grid ? net.grid.divide_factor : 1.0f;
    const float bl[3] = {net.b[2][0], net.b[2][1], net.b[2][2]};
    const float *wl = smem + pl.wl;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + (tid >> 7); tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (tid & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;
        const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
        {
            const float v[3] = {__ldg(V + 3 * (size_t)p), __ldg(V + 3 * (size_t)p + 1), __ldg(V + 3 * (size_t)p + 2)};
            const float nr[3] = {__ldg(Nrm + 3 * (size_t)p), __ldg(Nrm + 3 * (size_t)p + 1), __ldg(Nrm + 3 * (size_t)p + 2)};
            float xvn[40];
            ct_xvn<false>(x, v, nr, xvn);
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) st_a8(t, 4 + c8, &xvn[c8 * 8]);     // columns 32..71
            if (H0 && valid) {
#pragma unroll
                for (int k = 0; k < 33; ++k) H0[(size_t)k * Ps + p] = xvn[k];  // rows 0..32 of the input (for dW0)
            }
        }
        if (H0) {
            // grid features (and DYDX) were gathered by grid_encode_kernel into the grid rows of H0
            float gf[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) gf[k] = (k < n_grid) ? __ldg(H0 + (size_t)(97 + k) * Ps + p) : 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) st_a8(t, c8, &gf[c8 * 8]);
        } else {
            float u[3];
            to_unit(x, df, u);
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                float feat[C], dfeat[3][C];
                if (l < L) {
                    if (DYDX) {
                        encode_level<C, true>(net.grid.table, lv[l], u, feat, dfeat);
                        if (valid) {
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int c = 0; c < C; ++c) DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p] = dfeat[d][c];
                        }
                    } else {
                        encode_level<C, false>(net.grid.table, lv[l], u, feat, dfeat);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) feat[c] = 0.f;
                }
                st_a_small<C>(t, l * C, feat);
            }
        }
        ct_issue(t, pl, 0, CT_KA, NICER_W, smem);                 // block a
        float fv[NICER_W];
        load64(feat_fm, 0, Ps, p, fv);
        gemm_wait(t);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) st_a8(t, c8, &fv[c8 * 8]);
        ct_issue(t, pl, 1, NICER_W, NICER_W, smem, true);         // block b accumulates
        gemm_wait(t);
        // hidden layer 1 (ReLU), layer 2, output
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            ld_d8(t, c8, v);
            tc::wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = fmaxf(v[i] + smem[pl.b0 + c8 * 8 + i], 0.f);
                if (valid) A_fm[((size_t)(c8 * 8 + i)) * Ps + p] = v[i];
            }
            st_a8(t, c8, v);
        }
        ct_issue(t, pl, 2, NICER_W, NICER_W, smem);
        gemm_wait(t);
        float o[3] = {bl[0], bl[1], bl[2]};
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            ld_d8(t, c8, v);
            tc::wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = c8 * 8 + i;
                const float a = fmaxf(v[i] + smem[pl.b1 + k], 0.f);
                if (valid) A_fm[((size_t)(NICER_W + k)) * Ps + p] = a;
                o[0] += wl[k] * a; o[1] += wl[NICER_W + k] * a; o[2] += wl[2 * NICER_W + k] * a;
            }
        }
        if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[3 * (size_t)p + c] = sigmoidf_(o[c]);
        }
    }
    tile_teardown(sh);
}

template <int C>
__global__ void __launch_bounds__(TCF_THREADS, 1)
color_backward_tc_kernel(const nicer_color_net_t net, const LevelScales ls, const ColorPlan pl, const float *__restrict__ X,
                         const float *__restrict__ V, const float *__restrict__ Nrm, uint32_t P, const float *__restrict__ rgb,
                         const float *__restrict__ A_fm, const float *__restrict__ DYDX, const float *__restrict__ g_rgb,
                         float *grad_x, float *grad_view, float *grad_normals, float *grad_feat_fm, float *ZB,
                         float *OB, float *GY) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcfShared sh;
    const int tid = threadIdx.x;
    const bool has_grid = net.grid.table != nullptr;
    const int L = has_grid ? (int)net.grid.L : 0;
    const int n_grid = L * C;
    const int d_in = 33 + 64 + n_grid;
    const bool detached = net.grid_detached != 0;
    const float *W0 = net.W[0], *W1 = net.W[1];
    // W1^T: B[n = k_in][k = j_out] = W1[j][k_in]
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[0], smem + pl.w_lo[0], [&](int n, int k) { return W1[(size_t)k * NICER_W + n]; });
    // W0a^T: rows = block-a columns (80, padded), contraction over W0's rows
    ct_stage(CT_NA, NICER_W, smem + pl.w_hi[1], smem + pl.w_lo[1], [&](int n, int k) {
        const int c = (n < CT_KA) ? ct_col_a(n, n_grid) : -1;
        return c >= 0 ? W0[(size_t)k * d_in + c] : 0.f; });
    // W0b^T: rows = feature columns
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[2], smem + pl.w_lo[2], [&](int n, int k) { return W0[(size_t)k * d_in + 33 + n]; });
    for (int i = tid; i < 3 * NICER_W; i += TCF_THREADS) smem[pl.wl + i] = net.W[2][i];
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + pl.lv);
    for (int l = tid; l < L; l += TCF_THREADS) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    Tile t = tile_setup(sh, CB_ALO, CB_D);
    const size_t Ps = P;
    const float df = has_grid ? net.grid.divide_factor : 1.0f;
    const float *wl = smem + pl.wl;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + (tid >> 7); tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (tid & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;
        // ---- output layer backward in registers, zbar_2 -> A
        float ob[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float y = __ldg(rgb + 3 * (size_t)p + c);
            ob[c] = __ldg(g_rgb + 3 * (size_t)p + c) * (1.0f - y) * y;
            if (valid) OB[(size_t)c * Ps + p] = ob[c];
        }
        {
            float a2[NICER_W];
            load64(A_fm, NICER_W, Ps, p, a2);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = c8 * 8 + i;
                    const float abar = wl[k] * ob[0] + wl[NICER_W + k] * ob[1] + wl[2 * NICER_W + k] * ob[2];
                    v[i] = a2[k] > 0.f ? abar : 0.f;
                    if (valid) ZB[((size_t)(NICER_W + k)) * Ps + p] = v[i];
                }
                st_a8(t, c8, v);
            }
        }
        ct_issue(t, pl, 0, NICER_W, NICER_W, smem);       // abar_1 = W1^T zbar_2
        float a1[NICER_W];
        load64(A_fm, 0, Ps, p, a1);
        gemm_wait(t);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            ld_d8(t, c8, v);
            tc::wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = c8 * 8 + i;
                v[i] = a1[k] > 0.f ? v[i] : 0.f;
                if (valid) ZB[(size_t)k * Ps + p] = v[i];
            }
            st_a8(t, c8, v);
        }
        ct_issue(t, pl, 1, NICER_W, CT_NA, smem);          // hbar block a (80 columns) = W0a^T zbar_1
        const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
        const float v3[3] = {__ldg(V + 3 * (size_t)p), __ldg(V + 3 * (size_t)p + 1), __ldg(V + 3 * (size_t)p + 2)};
        const float nr[3] = {__ldg(Nrm + 3 * (size_t)p), __ldg(Nrm + 3 * (size_t)p + 1), __ldg(Nrm + 3 * (size_t)p + 2)};
        float dyv[96];
        const bool use_dx = (DYDX != nullptr) && !detached;
#pragma unroll
        for (int k = 0; k < 96; ++k) dyv[k] = (use_dx && k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
        gemm_wait(t);
        float xb[3], gy[32];
        {
            float hp[40];
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &hp[c8 * 8]);     // columns 32..71: x, view PE, normals
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) ld_d8(t, c8, &gy[c8 * 8]);         // columns 0..31: grid
            tc::wait_ld();
            xb[0] = hp[0]; xb[1] = hp[1]; xb[2] = hp[2];
            float vb[3] = {hp[3], hp[4], hp[5]};
            float xvn[40];
            ct_xvn<true>(x, v3, nr, xvn);
            float fr = 1.0f;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float s = xvn[6 + 6 * f + d], c = xvn[6 + 6 * f + 3 + d];
                    vb[d] += fr * (c * hp[6 + 6 * f + d] - s * hp[6 + 6 * f + 3 + d]);
                }
                fr *= 2.0f;
            }
            if (valid) {
                if (grad_view) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) grad_view[3 * (size_t)p + d] = vb[d];
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) grad_normals[3 * (size_t)p + d] = hp[30 + d];
            }
        }
        ct_issue(t, pl, 2, NICER_W, NICER_W, smem);        // hbar block b (features) = W0b^T zbar_1; A still holds zbar_1
        // grid: dL/dx through the grid; dL/d(enc) goes to the scatter kernel (grid_scatter.cu)
        float xu[3] = {0.f, 0.f, 0.f};
        if (has_grid && !detached) {
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                if (l < L) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) xu[d] += gy[l * C + c] * dyv[(l * 3 + d) * C + c];
                        if (valid) GY[(size_t)(l * C + c) * Ps + p] = gy[l * C + c];
                    }
                }
            }
        }
        if (grad_x && valid) {
#pragma unroll
            for (int d = 0; d < 3; ++d) grad_x[3 * (size_t)p + d] += xb[d] + xu[d] / 2.0f / df;
        }
        gemm_wait(t);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            ld_d8(t, c8, v);
            tc::wait_ld();
            if (valid) {
#pragma unroll
                for (int i = 0; i < 8; ++i) grad_feat_fm[(size_t)(c8 * 8 + i) * Ps + p] = v[i];
            }
        }
    }
    tile_teardown(sh);
}


// =====================================================================================================================
// Two threads per point (the default; the scheme of sdf_tc_split.cu): a 128-point tile is served by 256 threads, the thread pair
// of a point splitting the 64 accumulator columns in halves.  Of the 72-column block a [32 grid | 33 x, PE(view), normals | pad],
// h = 0 owns the grid part (features, d feat/dx rows, grid-gradient hand-over) and h = 1 the rest; the 3-wide output layer is a
// partial dot product per half, handed over through shared memory.  Same plans, operands and saved tensors as the kernels above.
__device__ __forceinline__ void ct_issue2(Tile &t, const ColorPlan &pl, int i, int K, int N, float *smem, bool acc_first = false) {
    gemm_issue2(t, tc::smem_u32(smem + pl.w_hi[i]), tc::smem_u32(smem + pl.w_lo[i]), K, N, acc_first);
}

template <int C>
__global__ void __launch_bounds__(TCS_THREADS, 1)
color_forward_tcs_kernel(const nicer_color_net_t net, const LevelScales ls, const ColorPlan pl, const float *__restrict__ X,
                         const float *__restrict__ V, const float *__restrict__ Nrm, const float *__restrict__ feat_fm, uint32_t P,
                         float *rgb, float *A_fm, float *DYDX, float *H0) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcsShared sh;
    const int tid = threadIdx.x;
    const bool has_grid = net.grid.table != nullptr;
    const int L = has_grid ? (int)net.grid.L : 0;
    const int n_grid = L * C;
    const int d_in = 33 + 64 + n_grid;
    const float *W0 = net.W[0], *W1 = net.W[1];
    ct_stage(NICER_W, CT_KA, smem + pl.w_hi[0], smem + pl.w_lo[0], [&](int n, int k) {
        const int c = ct_col_a(k, n_grid);
        return c >= 0 ? W0[(size_t)n * d_in + c] : 0.f; });
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[1], smem + pl.w_lo[1], [&](int n, int k) { return W0[(size_t)n * d_in + 33 + k]; });
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[2], smem + pl.w_lo[2], [&](int n, int k) { return W1[(size_t)n * NICER_W + k]; });
    for (int i = tid; i < NICER_W; i += TCS_THREADS) { smem[pl.b0 + i] = net.b[0][i]; smem[pl.b1 + i] = net.b[1][i]; }
    for (int i = tid; i < 3 * NICER_W; i += TCS_THREADS) smem[pl.wl + i] = net.W[2][i];
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + pl.lv);
    for (int l = tid; l < L; l += TCS_THREADS) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    Tile t = tile_setup2(sh, CT_ALO, CT_D);
    const int h = (tid >> 7) & 1, c0 = 4 * h, tile = tid >> 8, lane = tid & 127;
    const size_t Ps = P;
    const float df = has_grid ? net.grid.divide_factor : 1.0f;
    const float bl[3] = {net.b[2][0], net.b[2][1], net.b[2][2]};
    const float *wl = smem + pl.wl + c0 * 8;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;
        if (h == 1) {
            const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
            const float v[3] = {__ldg(V + 3 * (size_t)p), __ldg(V + 3 * (size_t)p + 1), __ldg(V + 3 * (size_t)p + 2)};
            const float nr[3] = {__ldg(Nrm + 3 * (size_t)p), __ldg(Nrm + 3 * (size_t)p + 1), __ldg(Nrm + 3 * (size_t)p + 2)};
            float xvn[40];
            ct_xvn<false>(x, v, nr, xvn);
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) st_a8(t, 4 + c8, &xvn[c8 * 8]);     // columns 32..71
            if (H0 && valid) {
#pragma unroll
                for (int k = 0; k < 33; ++k) H0[(size_t)k * Ps + p] = xvn[k];  // rows 0..32 of the input (for dW0)
            }
        } else if (H0) {
            // grid features (and DYDX) were gathered by grid_encode_kernel into the grid rows of H0
            float gf[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) gf[k] = (k < n_grid) ? __ldg(H0 + (size_t)(97 + k) * Ps + p) : 0.f;
            st_half(t, 0, gf);
        } else {
            const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
            float u[3];
            to_unit(x, df, u);
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                float feat[C], dfeat[3][C];
                if (l < L) {
                    if (DYDX) {
                        encode_level<C, true>(net.grid.table, lv[l], u, feat, dfeat);
                        if (valid) {
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int c = 0; c < C; ++c) DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p] = dfeat[d][c];
                        }
                    } else {
                        encode_level<C, false>(net.grid.table, lv[l], u, feat, dfeat);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) feat[c] = 0.f;
                }
                st_a_small<C>(t, l * C, feat);
            }
        }
        ct_issue2(t, pl, 0, CT_KA, NICER_W, smem);                // block a
        float fv[32];
        load32(feat_fm, (size_t)c0 * 8, Ps, p, fv);
        gemm_wait(t);
        st_half(t, c0, fv);
        ct_issue2(t, pl, 1, NICER_W, NICER_W, smem, true);        // block b accumulates
        gemm_wait(t);
        // hidden layer 1 (ReLU), layer 2, output
        {
            float v[32];
            ld_half(t, c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                v[i] = fmaxf(v[i] + smem[pl.b0 + c0 * 8 + i], 0.f);
                if (valid) A_fm[((size_t)(c0 * 8 + i)) * Ps + p] = v[i];
            }
            st_half(t, c0, v);
        }
        ct_issue2(t, pl, 2, NICER_W, NICER_W, smem);
        gemm_wait(t);
        float o[3] = {0.f, 0.f, 0.f};
        {
            float v[32];
            ld_half(t, c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float a = fmaxf(v[i] + smem[pl.b1 + c0 * 8 + i], 0.f);
                if (valid) A_fm[((size_t)(NICER_W + c0 * 8 + i)) * Ps + p] = a;
                o[0] += wl[i] * a; o[1] += wl[NICER_W + i] * a; o[2] += wl[2 * NICER_W + i] * a;
            }
        }
        if (h == 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) sh.xch[tile][lane][c] = o[c];
        }
        tile_sync2(t);
        if (h == 0 && valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[3 * (size_t)p + c] = sigmoidf_((bl[c] + o[c]) + sh.xch[tile][lane][c]);
        }
        // (the next write to xch comes after the tile barriers of the next tile's MMA groups)
    }
    tile_teardown2(sh);
}

template <int C>
__global__ void __launch_bounds__(TCS_THREADS, 1)
color_backward_tcs_kernel(const nicer_color_net_t net, const LevelScales ls, const ColorPlan pl, const float *__restrict__ X,
                          const float *__restrict__ V, const float *__restrict__ Nrm, uint32_t P, const float *__restrict__ rgb,
                          const float *__restrict__ A_fm, const float *__restrict__ DYDX, const float *__restrict__ g_rgb,
                          float *grad_x, float *grad_view, float *grad_normals, float *grad_feat_fm, float *ZB,
                          float *OB, float *GY) {
    extern __shared__ __align__(16) float smem[];
    __shared__ TcsShared sh;
    const int tid = threadIdx.x;
    const bool has_grid = net.grid.table != nullptr;
    const int L = has_grid ? (int)net.grid.L : 0;
    const int n_grid = L * C;
    const int d_in = 33 + 64 + n_grid;
    const bool detached = net.grid_detached != 0;
    const float *W0 = net.W[0], *W1 = net.W[1];
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[0], smem + pl.w_lo[0], [&](int n, int k) { return W1[(size_t)k * NICER_W + n]; });
    ct_stage(CT_NA, NICER_W, smem + pl.w_hi[1], smem + pl.w_lo[1], [&](int n, int k) {
        const int c = (n < CT_KA) ? ct_col_a(n, n_grid) : -1;
        return c >= 0 ? W0[(size_t)k * d_in + c] : 0.f; });
    ct_stage(NICER_W, NICER_W, smem + pl.w_hi[2], smem + pl.w_lo[2], [&](int n, int k) { return W0[(size_t)k * d_in + 33 + n]; });
    for (int i = tid; i < 3 * NICER_W; i += TCS_THREADS) smem[pl.wl + i] = net.W[2][i];
    Tile t = tile_setup2(sh, CB_ALO, CB_D);
    const int h = (tid >> 7) & 1, c0 = 4 * h, tile = tid >> 8, lane = tid & 127;
    const size_t Ps = P;
    const float df = has_grid ? net.grid.divide_factor : 1.0f;
    const float *wl = smem + pl.wl + c0 * 8;
    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;
        // ---- output layer backward in registers, zbar_2 -> A (this thread's 32 columns)
        float ob[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float y = __ldg(rgb + 3 * (size_t)p + c);
            ob[c] = __ldg(g_rgb + 3 * (size_t)p + c) * (1.0f - y) * y;
            if (valid && h == 0) OB[(size_t)c * Ps + p] = ob[c];
        }
        {
            float v[32];
            load32(A_fm, (size_t)NICER_W + c0 * 8, Ps, p, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float abar = wl[i] * ob[0] + wl[NICER_W + i] * ob[1] + wl[2 * NICER_W + i] * ob[2];
                v[i] = v[i] > 0.f ? abar : 0.f;
                if (valid) ZB[((size_t)(NICER_W + c0 * 8 + i)) * Ps + p] = v[i];
            }
            st_half(t, c0, v);
        }
        ct_issue2(t, pl, 0, NICER_W, NICER_W, smem);       // abar_1 = W1^T zbar_2
        {
            float a1[32], v[32];
            load32(A_fm, (size_t)c0 * 8, Ps, p, a1);
            gemm_wait(t);
            ld_half(t, c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                v[i] = a1[i] > 0.f ? v[i] : 0.f;
                if (valid) ZB[(size_t)(c0 * 8 + i) * Ps + p] = v[i];
            }
            st_half(t, c0, v);
        }
        ct_issue2(t, pl, 1, NICER_W, CT_NA, smem);         // hbar block a (80 columns) = W0a^T zbar_1
        if (h == 1) {
            const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
            const float v3[3] = {__ldg(V + 3 * (size_t)p), __ldg(V + 3 * (size_t)p + 1), __ldg(V + 3 * (size_t)p + 2)};
            const float nr[3] = {__ldg(Nrm + 3 * (size_t)p), __ldg(Nrm + 3 * (size_t)p + 1), __ldg(Nrm + 3 * (size_t)p + 2)};
            float xvn[40];
            ct_xvn<true>(x, v3, nr, xvn);
            gemm_wait(t);
            float hp[40];
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) ld_d8(t, 4 + c8, &hp[c8 * 8]);     // columns 32..71: x, view PE, normals
            tc::wait_ld();
            ct_issue2(t, pl, 2, NICER_W, NICER_W, smem);   // hbar block b (features) = W0b^T zbar_1; A still holds zbar_1
            float vb[3] = {hp[3], hp[4], hp[5]};
            float fr = 1.0f;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float s = xvn[6 + 6 * f + d], c = xvn[6 + 6 * f + 3 + d];
                    vb[d] += fr * (c * hp[6 + 6 * f + d] - s * hp[6 + 6 * f + 3 + d]);
                }
                fr *= 2.0f;
            }
            if (valid) {
                if (grad_view) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) grad_view[3 * (size_t)p + d] = vb[d];
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) grad_normals[3 * (size_t)p + d] = hp[30 + d];
            }
            tile_sync2(t);                                  // grid part of dL/dx from the other half
            if (grad_x && valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) grad_x[3 * (size_t)p + d] += hp[d] + sh.xch[tile][lane][d];
            }
        } else {
            float dyv[96];
            const bool use_dx = (DYDX != nullptr) && !detached;
#pragma unroll
            for (int k = 0; k < 96; ++k) dyv[k] = (use_dx && k < L * 3 * C) ? __ldg(DYDX + (size_t)k * Ps + p) : 0.f;
            gemm_wait(t);
            // grid (columns 0..31): dL/dx through the grid; dL/d(enc) goes to the scatter kernel (grid_scatter.cu).  Consumed level by
            // level before the next MMA group overwrites the accumulator (dyv[96] + all 32 columns at once would not fit 128 registers)
            float xu[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int l = 0; l < 32 / C; ++l) {
                if (l < L) {
                    float gy[C];
                    ld_d_small<C>(t, l * C, gy);
                    tc::wait_ld();
                    if (has_grid && !detached) {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
#pragma unroll
                            for (int d = 0; d < 3; ++d) xu[d] += gy[c] * dyv[(l * 3 + d) * C + c];
                            if (valid) GY[(size_t)(l * C + c) * Ps + p] = gy[c];
                        }
                    }
                }
            }
            ct_issue2(t, pl, 2, NICER_W, NICER_W, smem);   // (both halves take part in the tile barrier of this issue)
#pragma unroll
            for (int d = 0; d < 3; ++d) sh.xch[tile][lane][d] = xu[d] / 2.0f / df;
            tile_sync2(t);
        }
        gemm_wait(t);
        {
            float v[32];
            ld_half(t, c0, v);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) grad_feat_fm[(size_t)(c0 * 8 + i) * Ps + p] = v[i];
            }
        }
    }
    tile_teardown2(sh);
}

bool tc_enabled();

// NICER_COLOR_SPLIT=0 selects the one-thread-per-point kernels
static bool color_split() {
    static const bool on = [] { const char *e = getenv("NICER_COLOR_SPLIT"); return !(e && e[0] == '0'); }();
    return on;
}

// 1: launched, 0: configuration not covered (caller falls back to the SIMT kernel), < 0: error
int launch_grid_encode(const nicer_grid_t *g, const float *x, uint32_t P, float *F, float *DYDX, cudaStream_t st);

int launch_color_forward_tc(const nicer_color_net_t *net, const float *x, const float *view, const float *normals, const float *feat_fm,
                            uint32_t P, float *rgb, float *A_fm, float *DYDX, float *H0, cudaStream_t st) {
    if (!tc_enabled() || net->n_hidden != 2 || net->multires_view != 4 || net->feature != 64) return 0;
    const bool has_grid = net->grid.table != nullptr;
    const LevelScales ls = host_level_scales(has_grid ? net->grid.L : 0, net->grid.S, net->grid.H);
    const ColorPlan pl = color_plan_fwd();
    const size_t smem = (size_t)pl.total_floats * sizeof(float);
    const uint32_t pairs = div_up(div_up(P, 128), 2);
    const uint32_t grid = pairs < (uint32_t)num_sms() ? pairs : (uint32_t)num_sms();
    if (H0 && has_grid) {      // gathers at full occupancy, into the grid rows of the saved network input
        if (int e = launch_grid_encode(&net->grid, x, P, H0 + (size_t)97 * P, DYDX, st)) return e;
    }
#define LAUNCH(CC)                                                                                                        \
    do {                                                                                                                  \
        NICER_CUDA(cudaFuncSetAttribute(color_forward_tc_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_color_forward(tc)");                                                                            \
        NICER_CUDA(cudaFuncSetAttribute(color_forward_tcs_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_color_forward(tcs)");                                                                           \
        if (color_split())                                                                                                \
            color_forward_tcs_kernel<CC><<<grid, TCS_THREADS, smem, st>>>(*net, ls, pl, x, view, normals, feat_fm, P, rgb, A_fm, DYDX, H0); \
        else                                                                                                              \
            color_forward_tc_kernel<CC><<<grid, TCF_THREADS, smem, st>>>(*net, ls, pl, x, view, normals, feat_fm, P, rgb, A_fm, DYDX, H0); \
    } while (0)
    switch (has_grid ? net->grid.C : 2) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_color_forward(tc)");
    return 1;
}

int launch_grid_scatter(const nicer_grid_t *g, const float *x, uint32_t P, const float *GY1, const float *GY2,
                        const float *g_grad, float *grad_table, cudaStream_t st);

int launch_color_backward_tc(const nicer_color_net_t *net, const float *x, const float *view, const float *normals, uint32_t P,
                             const float *rgb, const float *A_fm, const float *DYDX, const float *g_rgb, float *grad_x,
                             float *grad_view, float *grad_normals, float *grad_feat_fm, float *grad_table, float *ZB, float *OB,
                             float *GY, cudaStream_t st, cudaStream_t scatter_st) {
    if (!tc_enabled() || net->n_hidden != 2 || net->multires_view != 4 || net->feature != 64) return 0;
    const bool has_grid = net->grid.table != nullptr;
    const LevelScales ls = host_level_scales(has_grid ? net->grid.L : 0, net->grid.S, net->grid.H);
    const ColorPlan pl = color_plan_bwd();
    const size_t smem = (size_t)pl.total_floats * sizeof(float);
    const uint32_t pairs = div_up(div_up(P, 128), 2);
    const uint32_t grid = pairs < (uint32_t)num_sms() ? pairs : (uint32_t)num_sms();
#define LAUNCH(CC)                                                                                                         \
    do {                                                                                                                   \
        NICER_CUDA(cudaFuncSetAttribute(color_backward_tc_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_color_backward(tc)");                                                                            \
        NICER_CUDA(cudaFuncSetAttribute(color_backward_tcs_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_color_backward(tcs)");                                                                           \
        if (color_split())                                                                                                 \
            color_backward_tcs_kernel<CC><<<grid, TCS_THREADS, smem, st>>>(*net, ls, pl, x, view, normals, P, rgb, A_fm, DYDX, g_rgb, \
                                                                           grad_x, grad_view, grad_normals, grad_feat_fm, ZB, OB, GY); \
        else                                                                                                               \
            color_backward_tc_kernel<CC><<<grid, TCF_THREADS, smem, st>>>(*net, ls, pl, x, view, normals, P, rgb, A_fm, DYDX, g_rgb, \
                                                                          grad_x, grad_view, grad_normals, grad_feat_fm, ZB, OB, GY); \
    } while (0)
    switch (has_grid ? net->grid.C : 2) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_color_backward(tc)");
    if (has_grid && !net->grid_detached) {
        if (scatter_st && scatter_st != st) {
            if (int e = stream_fork(st, scatter_st)) return e;
        } else {
            scatter_st = st;
        }
        if (int e = launch_grid_scatter(&net->grid, x, P, GY, nullptr, nullptr, grad_table, scatter_st)) return e;
    }
    return 1;
}

}  // namespace nicer
