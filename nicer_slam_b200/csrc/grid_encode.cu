// Grid-feature gather of the fused network forward passes, as its own high-occupancy kernel (the mirror image of
// grid_scatter.cu).  The tcgen05 forward kernels run 8-16 warps per SM; issuing the 8 * L dependent gathers per point
// from there exposes their latency (ncu: ~30% of the stall samples of the forward kernels sat on the gather results).
// This kernel does the index arithmetic + gathers with one thread per (point, level) at full occupancy and writes
//     F    [L*C][P]      enc(x), feature-major (rows l*C + c)      -- the grid rows of the saved network input H0
//     DYDX [L*3*C][P]    d enc / d u (rows (l*3 + d)*C + c)        -- only when requested
// which the tensor-core kernels then read with coalesced loads.  Same arithmetic as K1 of the reference
// (hashencoder.cu:131-283) through encode_level (nicer_math.cuh).
#include "common.cuh"
#include "sdf_sample.cuh"

namespace nicer {

constexpr int GE_BLOCK = 256;

template <int C, bool WITH_DX>
__global__ void __launch_bounds__(GE_BLOCK)
grid_encode_kernel(const nicer_grid_t g, const LevelScales ls, const float *__restrict__ X, uint32_t P, float *__restrict__ F,
                   float *__restrict__ DYDX) {
    const uint32_t p = blockIdx.x * GE_BLOCK + threadIdx.x;
    const uint32_t l = blockIdx.y;
    if (p >= P) return;
    const size_t Ps = P;
    const float x[3] = {__ldg(X + 3 * (size_t)p), __ldg(X + 3 * (size_t)p + 1), __ldg(X + 3 * (size_t)p + 2)};
    float u[3];
    to_unit(x, g.divide_factor, u);
    const LevelInfo li = make_level(g.offsets, l, level_scale(ls, (uint32_t)l));
    float feat[C], dfeat[3][C];
    encode_level<C, WITH_DX>(g.table, li, u, feat, dfeat);
#pragma unroll
    for (int c = 0; c < C; ++c) F[(size_t)(l * C + c) * Ps + p] = feat[c];
    if (WITH_DX) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int c = 0; c < C; ++c) DYDX[((size_t)(l * 3 + d) * C + c) * Ps + p] = dfeat[d][c];
    }
}

int launch_grid_encode(const nicer_grid_t *g, const float *x, uint32_t P, float *F, float *DYDX, cudaStream_t st) {
    if (P == 0) return 0;
    const LevelScales ls = host_level_scales(g->L, g->S, g->H);
    const dim3 grid(div_up(P, GE_BLOCK), g->L);
#define LAUNCH(CC)                                                                                   \
    do {                                                                                             \
        if (DYDX) grid_encode_kernel<CC, true><<<grid, GE_BLOCK, 0, st>>>(*g, ls, x, P, F, DYDX);     \
        else grid_encode_kernel<CC, false><<<grid, GE_BLOCK, 0, st>>>(*g, ls, x, P, F, nullptr);      \
    } while (0)
    switch (g->C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("grid_encode");
    return 0;
}

}  // namespace nicer
