// Grid-gradient scatter of the fused network backward passes, as its own high-occupancy kernel.
//
// The tcgen05 backward kernels run 8 warps per SM (TMEM- and register-bound); issuing the 8 * L red.global.add per point
// from there stalls them, because every instruction that reuses an operand register of an in-flight reduction waits on
// the load/store queue (ncu: long-scoreboard stalls behind REDG on the instructions following the scatter loop).  They
// therefore only write the per-feature gradients dL/d(enc) -- GY1 [L*C][P] (first order, K2 of the reference,
// hashencoder.cu:286-331) and optionally GY2 [L*C][P] (the part of the second-order term that multiplies
// d(enc)/dx, K5, hashencoder.cu:430-508) -- feature-major, and this kernel does index arithmetic + atomics with one
// thread per (run of consecutive points, level) at full occupancy:
//     grad_table[idx_k] += w_k * GY1 + (sum_d dw_{d,k} * g_grad_d / (2 df)) * GY2        for the 8 corners k.
#include "common.cuh"
#include "sdf_sample.cuh"

namespace nicer {

constexpr int GS_BLOCK = 256;
constexpr int GS_RUN = 8;      // consecutive points (samples of one ray, sorted by depth) handled by one thread

// One (cell, level) worth of accumulated corner gradients -> atomics.  For C == 2 the two x-neighbours of a corner
// pair are adjacent 8-byte entries on dense levels and, on hashed levels, whenever the lower x coordinate is even; when
// the pair also sits in one aligned 16 bytes it takes one 16-byte reduction instead of two.
template <int C>
__device__ __forceinline__ void flush_cell(float *grad_table, const LevelInfo &li, const Cell3 &cell, const float (&acc)[8][C]) {
    uint32_t idx[8];
    corner_indices(li, cell, idx);
    if constexpr (C == 2) {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const bool lo_first = idx[k] < idx[k + 1];
            const uint32_t base = lo_first ? idx[k] : idx[k + 1], top = lo_first ? idx[k + 1] : idx[k];
            if (top == base + 1u && ((li.offset + base) & 1u) == 0u) {     // adjacent entries in one aligned 16 bytes
                float *p = grad_table + ((size_t)li.offset + base) * 2;
                const float4 v = lo_first ? make_float4(acc[k][0], acc[k][1], acc[k + 1][0], acc[k + 1][1])
                                          : make_float4(acc[k + 1][0], acc[k + 1][1], acc[k][0], acc[k][1]);
                atomicAdd(reinterpret_cast<float4 *>(p), v);
            } else {
                scatter_entry<C>(grad_table, li, idx[k], acc[k]);
                scatter_entry<C>(grad_table, li, idx[k + 1], acc[k + 1]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) scatter_entry<C>(grad_table, li, idx[k], acc[k]);
    }
}

// thread = (run of GS_RUN consecutive points, level): contributions to the same cell are summed in registers and
// flushed when the cell changes -- on the coarse dense levels consecutive samples of a ray share cells, which removes
// most of the same-address reductions.
template <int C, bool SECOND>
__global__ void __launch_bounds__(GS_BLOCK)
grid_scatter_kernel(const nicer_grid_t g, const LevelScales ls, const float *__restrict__ X, uint32_t P,
                    const float *__restrict__ GY1, const float *__restrict__ GY2, const float *__restrict__ g_grad,
                    float *grad_table) {
    const uint32_t run = blockIdx.x * GS_BLOCK + threadIdx.x;
    const uint32_t l = blockIdx.y;
    const uint32_t p0 = run * GS_RUN;
    if (p0 >= P) return;
    const size_t Ps = P;
    const LevelInfo li = make_level(g.offsets, l, level_scale(ls, (uint32_t)l));
    float acc[8][C];
    Cell3 cur;
    bool have = false;
#pragma unroll 1
    for (uint32_t i = 0; i < GS_RUN; ++i) {
        const uint32_t p = p0 + i;
        if (p >= P) break;
        const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
        float u[3];
        to_unit(x, g.divide_factor, u);
        const Cell3 cell = locate3(li, u);
        if (!cell.inside) continue;
        if (have && (cell.pg[0] != cur.pg[0] || cell.pg[1] != cur.pg[1] || cell.pg[2] != cur.pg[2])) {
            flush_cell<C>(grad_table, li, cur, acc);
            have = false;
        }
        if (!have) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[k][c] = 0.f;
            cur = cell;
            have = true;
        }
        float gy1[C], gy2[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            gy1[c] = __ldg(GY1 + (size_t)(l * C + c) * Ps + p);
            gy2[c] = SECOND ? __ldg(GY2 + (size_t)(l * C + c) * Ps + p) : 0.f;
        }
        float wt[8];
        corner_weights(cell, wt);
        if (SECOND) {
            float w2[8], dw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w2[k] = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float ggu = __ldg(g_grad + 3 * (size_t)p + d) / 2.0f / g.divide_factor;
                corner_dweights(cell, d, dw);
#pragma unroll
                for (int k = 0; k < 8; ++k) w2[k] += dw[k] * ggu;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[k][c] += wt[k] * gy1[c] + w2[k] * gy2[c];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[k][c] += wt[k] * gy1[c];
        }
    }
    if (have) flush_cell<C>(grad_table, li, cur, acc);
}

// GY2 == NULL: first-order term only
int launch_grid_scatter(const nicer_grid_t *g, const float *x, uint32_t P, const float *GY1, const float *GY2,
                        const float *g_grad, float *grad_table, cudaStream_t st) {
    if (!grad_table) return 0;      // table gradient not wanted (pose-only tracking): no scatter at all
    if (P == 0) return 0;
    const LevelScales ls = host_level_scales(g->L, g->S, g->H);
    const dim3 grid(div_up(div_up(P, GS_RUN), GS_BLOCK), g->L);
    const bool second = GY2 != nullptr && g_grad != nullptr;
#define LAUNCH(CC)                                                                                                   \
    do {                                                                                                             \
        if (second) grid_scatter_kernel<CC, true><<<grid, GS_BLOCK, 0, st>>>(*g, ls, x, P, GY1, GY2, g_grad, grad_table); \
        else grid_scatter_kernel<CC, false><<<grid, GS_BLOCK, 0, st>>>(*g, ls, x, P, GY1, nullptr, nullptr, grad_table);  \
    } while (0)
    switch (g->C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("grid_scatter");
    return 0;
}

}  // namespace nicer
