// Shared by the tcgen05 SDF main-pass kernels (sdf_tc_full.cu: one thread per point; sdf_tc_split.cu: two threads per
// point): TMEM column map, the staged-operand plans of the four kernels (A forward, B gradient chain, T tangent pass,
// R reverse pass), operand staging into the UMMA K-major no-swizzle layout.
#pragma once
#include "sdf_sample.cuh"
#include "tc_tile.cuh"

namespace nicer {
constexpr int TCF_K0 = 80;          // layer-0 input columns, zero padded (d_in <= 71); multiple of 16 so it can be an N
constexpr int TCF_ALO = 80, TCF_D = 160;   // TMEM columns of a tile: A hi [0,80), A lo [80,160), accumulator [160,240)

struct TcfPlan {
    MatSpec m[5];
    int n_mats;
    int bias[5];     // float offsets of per-layer biases (kernel A), -1 unused
    int wl_sdf, lv, total_floats;
};

// kernel A: W_0 .. W_{n-1} (forward) + feature head W_n[1:, :]
inline TcfPlan plan_a(const nicer_sdf_net_t *net) {
    TcfPlan pl;
    const int n = (int)net->n_hidden, d_in = 3 + 6 * (int)net->multires + (int)(net->grid.L * net->grid.C);
    int o = 0;
    pl.n_mats = n + 1;
    for (int l = 0; l <= n; ++l) {
        MatSpec &m = pl.m[l];
        m.layer = l; m.transposed = 0; m.colmap = (l == 0) ? 1 : 0; m.col0 = 0;
        m.row0 = (l == n) ? 1 : 0;
        m.w_rows = (l == n) ? (int)net->d_out - 1 : NICER_W;
        m.w_cols = (l == 0) ? d_in : NICER_W;
        m.rows = NICER_W; m.K = (l == 0) ? TCF_K0 : NICER_W;
        m.hi = o; o += m.rows * m.K;
        m.lo = o; o += m.rows * m.K;
        pl.bias[l] = o; o += NICER_W;
    }
    for (int l = n + 1; l < 5; ++l) pl.bias[l] = -1;
    pl.wl_sdf = o; o += NICER_W;
    pl.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    pl.total_floats = o;
    return pl;
}

// kernel B: W_{n-1}^T .. W_1^T (64 x 64) and W_0^T (80 rows x 64)
inline TcfPlan plan_b(const nicer_sdf_net_t *net) {
    TcfPlan pl;
    const int n = (int)net->n_hidden, d_in = 3 + 6 * (int)net->multires + (int)(net->grid.L * net->grid.C);
    int o = 0;
    pl.n_mats = n;
    for (int l = 0; l < n; ++l) {       // m[l] = W_l^T
        MatSpec &m = pl.m[l];
        m.layer = l; m.transposed = 1; m.colmap = (l == 0) ? 1 : 0; m.col0 = 0;
        m.row0 = 0; m.w_rows = NICER_W; m.w_cols = (l == 0) ? d_in : NICER_W;
        m.rows = (l == 0) ? TCF_K0 : NICER_W; m.K = NICER_W;
        m.hi = o; o += m.rows * m.K;
        m.lo = o; o += m.rows * m.K;
    }
    for (int l = 0; l < 5; ++l) pl.bias[l] = -1;
    pl.wl_sdf = o; o += NICER_W;
    pl.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    pl.total_floats = o;
    return pl;
}

// backward kernel T (tangent pass): W_0 .. W_{n-1} (forward orientation)
inline TcfPlan plan_t(const nicer_sdf_net_t *net) {
    TcfPlan pl = plan_a(net);
    // same operands as kernel A minus the feature head and the biases: rebuild compactly
    const int n = (int)net->n_hidden;
    int o = 0;
    pl.n_mats = n;
    for (int l = 0; l < n; ++l) {
        MatSpec &m = pl.m[l];
        m.hi = o; o += m.rows * m.K;
        m.lo = o; o += m.rows * m.K;
    }
    for (int l = 0; l < 5; ++l) pl.bias[l] = -1;
    pl.wl_sdf = o; o += NICER_W;
    pl.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    pl.total_floats = o;
    return pl;
}

// backward kernel R (reverse pass): m[0..n-1] = W_l^T as in kernel B, m[n] = (W_n[1:, :])^T (feature head transposed)
inline TcfPlan plan_r(const nicer_sdf_net_t *net) {
    TcfPlan pl = plan_b(net);
    const int n = (int)net->n_hidden;
    int o = 0;
    for (int l = 0; l < n; ++l) {
        MatSpec &m = pl.m[l];
        m.hi = o; o += m.rows * m.K;
        m.lo = o; o += m.rows * m.K;
    }
    MatSpec &f = pl.m[n];
    f.layer = n; f.transposed = 1; f.colmap = 0; f.col0 = 0;
    f.row0 = 1; f.w_rows = (int)net->d_out - 1; f.w_cols = NICER_W;
    f.rows = NICER_W; f.K = NICER_W;
    f.hi = o; o += f.rows * f.K;
    f.lo = o; o += f.rows * f.K;
    pl.n_mats = n + 1;
    pl.wl_sdf = o; o += NICER_W;
    pl.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    pl.total_floats = o;
    return pl;
}

// Column order of the layer-0 operand inside the kernel: [32 grid features | 39 PE values | zero padding] (grid
// features first so that every level's C features start at a column that is a multiple of C).  tcf_col_src maps an
// operand column to the column of the reference's input vector [PE 39 | grid L*C]  (-1: padding).
__device__ __forceinline__ int tcf_col_src(int k, int d_in) {
    if (k < 32) return (39 + k < d_in) ? 39 + k : -1;
    if (k < 71) return k - 32;
    return -1;
}

// Generic staging of a B operand with `rows` rows and K_pad contraction columns into hi/lo [k/4][rows][4].
//   transposed == false: B[n][k] = W[row0 + n][col(k)]   (W row-major [*, K_src]; layer0: col() permutes, see tcf_col_src)
//   transposed == true : B[n][k] = W[row0 + k][col(n)]   (B = W^T: rows index W's columns, contraction over W's rows)
static __device__ void tcf_stage(const float *__restrict__ W, int row0, int w_rows, int w_cols, int rows, int K_pad, bool transposed,
                          bool layer0, float *hi, float *lo) {
    for (int i = threadIdx.x; i < K_pad * rows; i += blockDim.x) {
        const int n = i / K_pad, k = i - n * K_pad;
        const int wr = transposed ? k : n;
        const int wc_raw = transposed ? n : k;
        const int wc = layer0 ? tcf_col_src(wc_raw, w_cols) : (wc_raw < w_cols ? wc_raw : -1);
        const float w = (wc >= 0 && wr < w_rows) ? W[(size_t)(row0 + wr) * w_cols + wc] : 0.f;
        const float h = tc::tf32_hi(w);
        const int dst = ((k >> 2) * rows + n) * 4 + (k & 3);
        hi[dst] = h;
        lo[dst] = w - h;
    }
}

// common per-CTA staging: operands, biases, last-layer sdf row, level table (any block size)
__device__ __forceinline__ void tcf_stage_all(const nicer_sdf_net_t &net, const LevelScales &ls, const TcfPlan &pl, float *smem,
                                              LevelInfo *&lv) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = (int)net.n_hidden, L = (int)net.grid.L;
    for (int i = 0; i < pl.n_mats; ++i) {
        const MatSpec &m = pl.m[i];
        tcf_stage(net.W[m.layer], m.row0, m.w_rows, m.w_cols, m.rows, m.K, m.transposed != 0, m.colmap != 0, smem + m.hi, smem + m.lo);
    }
    const int nfeat = (int)net.d_out - 1;
    for (int l = 0; l <= n; ++l) {
        if (pl.bias[l] < 0) continue;
        for (int i = tid; i < NICER_W; i += nthr)
            smem[pl.bias[l] + i] = (l < n) ? net.b[l][i] : ((i < nfeat) ? net.b[n][1 + i] : 0.f);
    }
    for (int i = tid; i < NICER_W; i += nthr) smem[pl.wl_sdf + i] = net.W[n][i];
    lv = reinterpret_cast<LevelInfo *>(smem + pl.lv);
    for (int l = tid; l < L; l += nthr) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
}

// staging + barriers + TMEM of a 256-thread, two-tile CTA. Returns the tile of the calling thread.
__device__ __forceinline__ Tile tcf_setup(const nicer_sdf_net_t &net, const LevelScales &ls, const TcfPlan &pl, float *smem,
                                          TcfShared &sh, LevelInfo *&lv) {
    tcf_stage_all(net, ls, pl, smem, lv);
    return tile_setup(sh, TCF_ALO, TCF_D);
}

__device__ __forceinline__ void mat_issue(Tile &t, const TcfPlan &pl, int i, float *smem) {
    gemm_issue(t, tc::smem_u32(smem + pl.m[i].hi), tc::smem_u32(smem + pl.m[i].lo), pl.m[i].K, pl.m[i].rows);
}
__device__ __forceinline__ void mat_gemm(Tile &t, const TcfPlan &pl, int i, float *smem) {
    mat_issue(t, pl, i, smem);
    gemm_wait(t);
}

}  // namespace nicer
