// Tensor-core (tcgen05, sm_100a) SDF-only forward: the sampler's 640-samples-per-ray pass and get_sdf_vals
// (/root/reference/code/model/ray_sampler.py:100-102, model/base_networks.py:27-32,223-228).
//
// Two kernels: sdf_only_tc4_kernel (default: four 128-point tiles per SM, K fed in 32-column chunks, grid features read from
// a scratch filled by grid_encode_kernel) and the earlier sdf_only_tc_kernel (NICER_TC_TILES=2: two tiles per SM, whole-K
// operands, gathers inline), kept for comparison.  In both, one thread = one point = one TMEM lane: the thread builds the
// network input (x, NeRF PE, hash/dense grid features) in registers and writes it with tcgen05.st into TMEM as the
// A operand (hi and lo halves of the 3xTF32 split); thread 0 issues tcgen05.mma kind::tf32 (M=128, N=64, K=8 per
// instruction; 3 MMAs per K-step for fp32 fidelity) against the weights held in shared memory in the UMMA K-major
// no-swizzle layout; every thread then reads its own accumulator row with tcgen05.ld.32x32b, applies bias +
// Softplus(100) and writes the next layer's A operand.  The last layer (one output: the sdf) is a 64-term fp32 dot.
#include "common.cuh"
#include "sdf_sample.cuh"
#include "tc_common.cuh"

namespace nicer {

constexpr int TC_BLOCK = 128;
constexpr int TC_K0 = 72;                 // layer-0 K padded to a multiple of 8 (d_in <= 71)
constexpr int TC_COL_AHI = 0;             // TMEM columns: A hi [0,72), A lo [72,144), D [144,208)
constexpr int TC_COL_ALO = TC_K0;
constexpr int TC_COL_D = 2 * TC_K0;
constexpr int TC_TMEM_COLS = 256;

struct TcSmemLayout {
    int w_hi[4], w_lo[4];   // per MMA layer: [K/4][64][4] floats (K-major core-matrix layout)
    int bias[4];            // per MMA layer bias [64]
    int wl_sdf, lv, total_floats;
};

static TcSmemLayout tc_layout(int n_hidden) {
    TcSmemLayout s;
    int o = 0;
    for (int l = 0; l < 4; ++l) {
        const int K = (l == 0) ? TC_K0 : NICER_W;
        s.w_hi[l] = o; if (l < n_hidden) o += K * NICER_W;
        s.w_lo[l] = o; if (l < n_hidden) o += K * NICER_W;
        s.bias[l] = o; if (l < n_hidden) o += NICER_W;
    }
    s.wl_sdf = o; o += NICER_W;
    s.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    s.total_floats = o;
    return s;
}

// stage W [64 out][K_src in] (row-major, global) as hi/lo in the UMMA K-major no-swizzle layout:
// element (n, k) -> chunk c = k/4 : base[(c*64 + n)*4 + k%4]; rows k >= K_src are zero padding
__device__ void tc_stage_weight(const float *__restrict__ W, int K_src, int K_pad, float *hi, float *lo) {
    for (int i = threadIdx.x; i < K_pad * NICER_W; i += blockDim.x) {
        const int n = i / K_pad, k = i - n * K_pad;
        const float w = (k < K_src) ? W[(size_t)n * K_src + k] : 0.f;
        const float h = tc::tf32_hi(w);
        const int dst = ((k >> 2) * NICER_W + n) * 4 + (k & 3);
        hi[dst] = h;
        lo[dst] = w - h;
    }
}

template <int C>
__global__ void __launch_bounds__(TC_BLOCK, 2)
sdf_only_tc_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcSmemLayout lay, const float *__restrict__ X,
                   uint32_t P, uint32_t accumulate, float *__restrict__ sdf) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n = (int)net.n_hidden;
    const int L = (int)net.grid.L;
    const int d_pe = 3 + 6 * (int)net.multires;
    const int d_in = d_pe + L * C;

    // ---- one-time setup: weights (hi/lo), biases, level table, mbarrier, TMEM
    for (int l = 0; l < n; ++l) {
        tc_stage_weight(net.W[l], l == 0 ? d_in : NICER_W, l == 0 ? TC_K0 : NICER_W, smem + lay.w_hi[l], smem + lay.w_lo[l]);
        for (int i = tid; i < NICER_W; i += TC_BLOCK) smem[lay.bias[l] + i] = net.b[l][i];
    }
    for (int i = tid; i < NICER_W; i += TC_BLOCK) smem[lay.wl_sdf + i] = net.W[n][i];
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + lay.lv);
    for (int l = tid; l < L; l += TC_BLOCK) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    if (tid == 0) { tc::mbar_init(&mma_bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_slot, TC_TMEM_COLS);
    // make the generic-proxy weight writes visible to the tensor core (async proxy) before any MMA
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    const float bl_sdf = net.b[n][0];
    const float df = net.grid.divide_factor;
    constexpr uint32_t IDESC = tc::idesc_tf32(128, NICER_W);
    uint32_t parity = 0;

    const uint32_t tiles = (P + TC_BLOCK - 1) / TC_BLOCK;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint32_t p = t * TC_BLOCK + tid;
        const bool valid = p < P;
        if (!valid) p = P - 1;   // keep every warp converged for the .sync.aligned TMEM ops
        // ---------------- network input -> TMEM (A operand of layer 0)
        {
            const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
            float h0[TC_K0];
            h0[0] = x[0]; h0[1] = x[1]; h0[2] = x[2];
            {
                // sin/cos(2^f x): one precise sincos per coordinate, then angle doubling (error grows ~2x per octave,
                // <~ 2e-6 at 2^5: fine for the no-grad sampler pass, whose output only places samples)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float s, c;
                    sincosf(x[d], &s, &c);
#pragma unroll
                    for (int f = 0; f < 6; ++f) {
                        h0[3 + 6 * f + d] = s;
                        h0[3 + 6 * f + 3 + d] = c;
                        const float s2 = 2.0f * s * c, c2 = fmaf(-2.0f * s, s, 1.0f);
                        s = s2; c = c2;
                    }
                }
            }
            float u[3];
            to_unit(x, df, u);
#pragma unroll
            for (int k = 39; k < TC_K0; ++k) h0[k] = 0.f;
            if (d_pe == 39) {
#pragma unroll
                for (int l = 0; l < 32 / C; ++l) {
                    if (l < L) {
                        float feat[C], dummy[3][C];
                        encode_level<C, false>(net.grid.table, lv[l], u, feat, dummy);
#pragma unroll
                        for (int c = 0; c < C; ++c) h0[39 + l * C + c] = feat[c];
                    }
                }
            }
#pragma unroll
            for (int c8 = 0; c8 < TC_K0 / 8; ++c8)
                tc::tmem_st8_split(lane_base + TC_COL_AHI + c8 * 8, lane_base + TC_COL_ALO + c8 * 8, &h0[c8 * 8]);
        }
        float a[NICER_W];
        for (int l = 0; l < n; ++l) {
            const int K = (l == 0) ? TC_K0 : NICER_W;
            tc::wait_st();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t whi = tc::smem_u32(smem + lay.w_hi[l]), wlo = tc::smem_u32(smem + lay.w_lo[l]);
                for (int ks = 0; ks < K / 8; ++ks) {
                    // one K-step = 8 tf32 = 2 chunks of 16 B; chunk stride (LBO) = 64 rows * 16 B, 8-row group stride (SBO) = 128 B
                    const uint64_t bhi = tc::smem_desc(whi + (uint32_t)ks * 2u * 1024u, 1024u, 128u);
                    const uint64_t blo = tc::smem_desc(wlo + (uint32_t)ks * 2u * 1024u, 1024u, 128u);
                    const uint32_t ahi = tmem_base + TC_COL_AHI + ks * 8, alo = tmem_base + TC_COL_ALO + ks * 8;
                    tc::mma_tf32_ts(tmem_base + TC_COL_D, ahi, bhi, IDESC, ks > 0 ? 1u : 0u);
                    tc::mma_tf32_ts(tmem_base + TC_COL_D, alo, bhi, IDESC, 1u);
                    tc::mma_tf32_ts(tmem_base + TC_COL_D, ahi, blo, IDESC, 1u);
                }
                tc::mma_commit(&mma_bar);
            }
            tc::mbar_wait(&mma_bar, parity);
            parity ^= 1u;
            __syncwarp();
            tc::fence_after_sync();
#pragma unroll
            for (int c8 = 0; c8 < NICER_W / 8; ++c8) tc::tmem_ld8(lane_base + TC_COL_D + c8 * 8, &a[c8 * 8]);
            tc::wait_ld();
            const float *bias = smem + lay.bias[l];
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) a[j] = softplus100_fast(a[j] + bias[j]);
            if (l + 1 < n) {
#pragma unroll
                for (int c8 = 0; c8 < NICER_W / 8; ++c8)
                    tc::tmem_st8_split(lane_base + TC_COL_AHI + c8 * 8, lane_base + TC_COL_ALO + c8 * 8, &a[c8 * 8]);
            }
        }
        // ---------------- output layer: sdf only
        float s = bl_sdf;
        const float *wl = smem + lay.wl_sdf;
#pragma unroll
        for (int k = 0; k < NICER_W; ++k) s += wl[k] * a[k];
        if (valid) {
            if (accumulate) sdf[p] += s; else sdf[p] = s;
        }
        // the next tile's tcgen05.st may not overtake this tile's tcgen05.ld of D (ordered by wait_ld above)
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}


// ---------------------------------------------------------------------------------------------------------------
// Four tiles per SM.  The kernel above is TMEM-bound at two 128-point tiles per SM (A hi 72 + A lo 72 + D 64 columns
// -> 256 of the 512 columns each), i.e. 8 warps per SM, and spends most of its time waiting on instruction latency.
// This variant feeds the A operand in K-chunks of 32 columns (hi 32 + lo 32 + D 64 = 128 columns per tile): a layer
// is 2 (hidden, K = 64) or 3 (layer 0, K = 72 = 32 + 32 + 8) accumulating MMA groups, each waited for before the
// operand columns are overwritten.  A CTA holds two tiles (256 threads, one staged copy of the weights), two CTAs
// fit an SM: 16 warps per SM.  Layer 0's first chunk (x and most of the PE) is issued before the grid gathers start.
constexpr int T4_THREADS = 256;
constexpr int T4_ALO = 32, T4_D = 64, T4_TILE_COLS = 128, T4_CTA_COLS = 256;

struct Tile4 {
    uint32_t tmem, lane_base;
    uint64_t *bar;
    uint32_t parity;
    int id;
    bool leader;
};

// MMAs for K-steps [ks0, ks0 + nks) of a staged [K/4][64][4] operand against the chunk columns [0, 8*nks) of the tile
__device__ __forceinline__ void chunk_issue(Tile4 &t, uint32_t whi, uint32_t wlo, int ks0, int nks, bool acc) {
    tc::wait_st();
    tc::fence_before_sync();
    asm volatile("bar.sync %0, 128;" ::"r"(t.id) : "memory");
    if (t.leader) {
        tc::fence_after_sync();
        constexpr uint32_t IDESC = tc::idesc_tf32(128, NICER_W);
        for (int ks = 0; ks < nks; ++ks) {
            const uint64_t bhi = tc::smem_desc(whi + (uint32_t)(ks0 + ks) * 2u * 1024u, 1024u, 128u);
            const uint64_t blo = tc::smem_desc(wlo + (uint32_t)(ks0 + ks) * 2u * 1024u, 1024u, 128u);
            const uint32_t ahi = t.tmem + ks * 8, alo = t.tmem + T4_ALO + ks * 8;
            tc::mma_tf32_ts(t.tmem + T4_D, ahi, bhi, IDESC, (acc || ks > 0) ? 1u : 0u);
            tc::mma_tf32_ts(t.tmem + T4_D, alo, bhi, IDESC, 1u);
            tc::mma_tf32_ts(t.tmem + T4_D, ahi, blo, IDESC, 1u);
        }
        tc::mma_commit(t.bar);
    }
}
__device__ __forceinline__ void chunk_wait(Tile4 &t) {
    tc::mbar_wait(t.bar, t.parity);
    t.parity ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
}
// 8*N8 values -> chunk columns [0, 8*N8) (hi) and [32, 32 + 8*N8) (lo)
template <int N8>
__device__ __forceinline__ void chunk_store(const Tile4 &t, const float *v) {
#pragma unroll
    for (int c8 = 0; c8 < N8; ++c8) tc::tmem_st8_split(t.lane_base + c8 * 8, t.lane_base + T4_ALO + c8 * 8, v + c8 * 8);
}

template <int C>
__global__ void __launch_bounds__(T4_THREADS, 2)
sdf_only_tc4_kernel(const nicer_sdf_net_t net, const LevelScales ls, const TcSmemLayout lay, const float *__restrict__ X,
                    uint32_t P, uint32_t accumulate, float *__restrict__ sdf, const float *__restrict__ F) {
    // F != NULL: grid features [L*C][P] gathered beforehand by grid_encode_kernel (coalesced reads here)
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, tile = tid >> 7;
    const int n = (int)net.n_hidden;
    const int L = (int)net.grid.L;
    const int d_in = 39 + L * C;

    for (int l = 0; l < n; ++l) {
        tc_stage_weight(net.W[l], l == 0 ? d_in : NICER_W, l == 0 ? TC_K0 : NICER_W, smem + lay.w_hi[l], smem + lay.w_lo[l]);
        for (int i = tid; i < NICER_W; i += T4_THREADS) smem[lay.bias[l] + i] = net.b[l][i];
    }
    for (int i = tid; i < NICER_W; i += T4_THREADS) smem[lay.wl_sdf + i] = net.W[n][i];
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + lay.lv);
    for (int l = tid; l < L; l += T4_THREADS) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_slot, T4_CTA_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    Tile4 t;
    t.tmem = tmem_base_slot + (uint32_t)tile * T4_TILE_COLS;
    t.lane_base = t.tmem + ((uint32_t)((warp & 3) * 32) << 16);
    t.bar = &bars[tile];
    t.parity = 0;
    t.id = 1 + tile;
    t.leader = (tid & 127) == 0;
    const float bl_sdf = net.b[n][0];
    const float df = net.grid.divide_factor;
    const float *wl = smem + lay.wl_sdf;

    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < tiles; tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + (tid & 127);
        const bool valid = p < P;
        if (!valid) p = P - 1;   // keep every warp converged for the .sync.aligned TMEM ops
        // ---------------- layer 0, input in the natural order [x 3 | PE 36 | grid 32 | pad 1], chunks 32 + 32 + 8
        {
            const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
            float h0[TC_K0];
            h0[0] = x[0]; h0[1] = x[1]; h0[2] = x[2];
            // sin/cos(2^f x): one precise sincos per coordinate, then angle doubling (error grows ~2x per octave,
            // <~ 2e-6 at 2^5: fine for the no-grad sampler pass, whose output only places samples)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(x[d], &s, &c);
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    h0[3 + 6 * f + d] = s;
                    h0[3 + 6 * f + 3 + d] = c;
                    const float s2 = 2.0f * s * c, c2 = fmaf(-2.0f * s, s, 1.0f);
                    s = s2; c = c2;
                }
            }
            const uint32_t w0h = tc::smem_u32(smem + lay.w_hi[0]), w0l = tc::smem_u32(smem + lay.w_lo[0]);
            chunk_store<4>(t, h0);
            chunk_issue(t, w0h, w0l, 0, 4, false);           // runs while the grid levels are gathered
            h0[71] = 0.f;
            if (F) {
#pragma unroll
                for (int k = 0; k < 32; ++k) h0[39 + k] = (k < L * C) ? __ldg(F + (size_t)k * P + p) : 0.f;
            } else {
                float u[3];
                to_unit(x, df, u);
#pragma unroll
                for (int l = 0; l < 32 / C; ++l) {
                    float feat[C], dummy[3][C];
                    if (l < L) {
                        encode_level<C, false>(net.grid.table, lv[l], u, feat, dummy);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) feat[c] = 0.f;
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) h0[39 + l * C + c] = feat[c];
                }
            }
            chunk_wait(t);
            chunk_store<4>(t, h0 + 32);
            chunk_issue(t, w0h, w0l, 4, 4, true);
            chunk_wait(t);
            chunk_store<1>(t, h0 + 64);
            chunk_issue(t, w0h, w0l, 8, 1, true);
            chunk_wait(t);
        }
        float a[NICER_W];
        for (int l = 0; l < n; ++l) {
#pragma unroll
            for (int c8 = 0; c8 < NICER_W / 8; ++c8) tc::tmem_ld8(t.lane_base + T4_D + c8 * 8, &a[c8 * 8]);
            tc::wait_ld();
            const float *bias = smem + lay.bias[l];
#pragma unroll
            for (int j = 0; j < NICER_W; ++j) a[j] = softplus100_fast(a[j] + bias[j]);
            if (l + 1 < n) {
                const uint32_t wh = tc::smem_u32(smem + lay.w_hi[l + 1]), wlo_ = tc::smem_u32(smem + lay.w_lo[l + 1]);
                chunk_store<4>(t, a);
                chunk_issue(t, wh, wlo_, 0, 4, false);
                chunk_wait(t);
                chunk_store<4>(t, a + 32);
                chunk_issue(t, wh, wlo_, 4, 4, true);
                chunk_wait(t);
            }
        }
        float s = bl_sdf;
#pragma unroll
        for (int k = 0; k < NICER_W; ++k) s += wl[k] * a[k];
        if (valid) {
            if (accumulate) sdf[p] += s; else sdf[p] = s;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_slot, T4_CTA_COLS);
}


// ---------------------------------------------------------------------------------------------------------------
// Stacked operands, two threads per point (NICER_TC_SAMPLER=s; NOT the default: measured on B200, fine net, 2.6 M points:
// 0.663 ms against 0.602 ms of the tc4 kernel above -- it issues 50 instead of 75 MMAs per tile, but its 256 TMEM columns per
// tile allow only two tiles per SM, and two tiles cannot keep the tensor pipe busy across the epilogues; tc4 runs four).
// The tc4 kernel is paced by tcgen05.mma issue: one instruction
// costs ~102 cycles whatever N <= 128 is (scripts/mma_bench.cu), and 3xTF32 at N = 64 needs three of them per K-step.  Here the
// B operand of a layer is the 128-row stack [W_hi ; W_lo]:
//     D[:, 0:128) (+)= a_hi [W_hi ; W_lo]^T      one N = 128 instruction:  hi*hi | hi*lo
//     D[:, 0:64)   += a_lo  W_hi^T               one N = 64 instruction on the first 64 rows of the same operand
// i.e. two instructions per K-step instead of three; the epilogue adds the two 64-column halves.  A tile is A hi [0,64),
// A lo [64,128), D [128,256): two tiles per SM, served by 512 threads -- thread pair (h = 0, 1) of a point shares its TMEM lane
// and splits the 64 outputs in halves, as in sdf_tc_split.cu -- so 16 warps hide the epilogue behind the other tile's MMAs.
// Layer 0 (K = 72) goes in two accumulating groups: [x | PE | 0] (40 columns, written by h = 1) is issued first and runs while
// h = 0 fetches the 32 grid features, which then reuse columns [0,32).
constexpr int S2_THREADS = 512;
constexpr int S2_ALO = 64, S2_D = 128, S2_TILE_COLS = 256, S2_TMEM = 512;
constexpr int S2_K0 = 72;                 // [x 3 | PE 36 | zero 1 | grid 32]

struct S2Layout {
    int w[4];               // per MMA layer: stacked [K/4][128][4] floats (rows 0..63 = hi, 64..127 = lo)
    int bias[4];
    int wl_sdf, total_floats;
};
static S2Layout s2_layout(int n_hidden) {
    S2Layout s;
    int o = 0;
    for (int l = 0; l < 4; ++l) {
        const int K = (l == 0) ? S2_K0 : NICER_W;
        s.w[l] = o; if (l < n_hidden) o += K * 2 * NICER_W;
        s.bias[l] = o; if (l < n_hidden) o += NICER_W;
    }
    s.wl_sdf = o; o += NICER_W;
    s.total_floats = o;
    return s;
}

// W [64 out][K_src in] -> stacked hi/lo operand; layer 0: operand column k -> source column (k < 39: k; 39: zero; 40..71: 39 + k - 40)
__device__ void s2_stage_weight(const float *__restrict__ W, int K_src, int K_pad, bool layer0, float *dst) {
    for (int i = threadIdx.x; i < K_pad * NICER_W; i += blockDim.x) {
        const int n = i / K_pad, k = i - n * K_pad;
        int ks = k;
        if (layer0) ks = (k < 39) ? k : ((k == 39) ? -1 : k - 1);
        const float w = (ks >= 0 && ks < K_src) ? W[(size_t)n * K_src + ks] : 0.f;
        const float h = tc::tf32_hi(w);
        const int base = ((k >> 2) * 2 * NICER_W) * 4 + (k & 3);
        dst[base + n * 4] = h;
        dst[base + (NICER_W + n) * 4] = w - h;
    }
}

struct Tile2 {
    uint32_t tmem, lane_base;
    uint64_t *bar;
    uint32_t parity;
    int id;
    bool leader;
};

// MMAs for K-steps [ks0, ks0 + nks) of a stacked operand against A columns [0, 8 nks)
__device__ __forceinline__ void s2_issue(Tile2 &t, uint32_t w, int ks0, int nks, bool acc) {
    tc::wait_st();
    tc::fence_before_sync();
    asm volatile("bar.sync %0, 256;" ::"r"(t.id) : "memory");
    if (t.leader) {
        tc::fence_after_sync();
        constexpr uint32_t I128 = tc::idesc_tf32(128, 2 * NICER_W), I64 = tc::idesc_tf32(128, NICER_W);
        constexpr uint32_t CHUNK = 2 * NICER_W * 16;       // one 4-float K-chunk of all 128 rows
        for (int ks = 0; ks < nks; ++ks) {
            const uint64_t b = tc::smem_desc(w + (uint32_t)(ks0 + ks) * 2u * CHUNK, CHUNK, 128u);
            tc::mma_tf32_ts(t.tmem + S2_D, t.tmem + ks * 8, b, I128, (acc || ks > 0) ? 1u : 0u);
            tc::mma_tf32_ts(t.tmem + S2_D, t.tmem + S2_ALO + ks * 8, b, I64, 1u);
        }
        tc::mma_commit(t.bar);
    }
}
__device__ __forceinline__ void s2_wait(Tile2 &t) {
    tc::mbar_wait(t.bar, t.parity);
    t.parity ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
}

template <int C>
__global__ void __launch_bounds__(S2_THREADS, 1)
sdf_only_tcs_kernel(const nicer_sdf_net_t net, const S2Layout lay, const float *__restrict__ X, uint32_t P, uint32_t accumulate,
                    float *__restrict__ sdf, const float *__restrict__ F) {
    // F: grid features [L*C][P] gathered beforehand by grid_encode_kernel (coalesced reads here)
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float xch[2][128];
    const int tid = threadIdx.x, warp = tid >> 5, tile = tid >> 8, h = (tid >> 7) & 1, lane = tid & 127, c0 = 32 * h;
    const int n = (int)net.n_hidden;
    const int L = (int)net.grid.L;
    const int d_in = 39 + L * C;

    for (int l = 0; l < n; ++l) {
        s2_stage_weight(net.W[l], l == 0 ? d_in : NICER_W, l == 0 ? S2_K0 : NICER_W, l == 0, smem + lay.w[l]);
        for (int i = tid; i < NICER_W; i += S2_THREADS) smem[lay.bias[l] + i] = net.b[l][i];
    }
    for (int i = tid; i < NICER_W; i += S2_THREADS) smem[lay.wl_sdf + i] = net.W[n][i];
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_slot, S2_TMEM);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    Tile2 t;
    t.tmem = tmem_base_slot + (uint32_t)tile * S2_TILE_COLS;
    t.lane_base = t.tmem + ((uint32_t)((warp & 3) * 32) << 16);
    t.bar = &bars[tile];
    t.parity = 0;
    t.id = 1 + tile;
    t.leader = (tid & 255) == 0;
    const float bl_sdf = net.b[n][0];
    const float *wl = smem + lay.wl_sdf + c0;
    const uint32_t w0 = tc::smem_u32(smem + lay.w[0]);

    const uint32_t tiles = (P + 127u) / 128u;
    for (uint32_t tt = blockIdx.x * 2 + tile; tt < ((tiles + 1u) & ~1u); tt += gridDim.x * 2) {
        uint32_t p = tt * 128u + lane;
        const bool valid = p < P;
        if (!valid) p = P - 1;   // keep every warp converged for the .sync.aligned TMEM ops
        // ---------------- layer 0, group 1: [x 3 | PE 36 | 0] from h = 1; h = 0 puts its feature loads in flight
        float gf[32];
        if (h == 1) {
            const float x[3] = {X[3 * (size_t)p], X[3 * (size_t)p + 1], X[3 * (size_t)p + 2]};
            float h0[40];
            h0[0] = x[0]; h0[1] = x[1]; h0[2] = x[2];
            // sin/cos(2^f x): one precise sincos per coordinate, then angle doubling (error grows ~2x per octave,
            // <~ 2e-6 at 2^5: fine for the no-grad sampler pass, whose output only places samples)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(x[d], &s, &c);
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    h0[3 + 6 * f + d] = s;
                    h0[3 + 6 * f + 3 + d] = c;
                    const float s2 = 2.0f * s * c, c2 = fmaf(-2.0f * s, s, 1.0f);
                    s = s2; c = c2;
                }
            }
            h0[39] = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < 5; ++c8) tc::tmem_st8_split(t.lane_base + c8 * 8, t.lane_base + S2_ALO + c8 * 8, &h0[c8 * 8]);
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) gf[k] = (k < L * C) ? __ldg(F + (size_t)k * P + p) : 0.f;
        }
        s2_issue(t, w0, 0, 5, false);
        s2_wait(t);
        // ---------------- layer 0, group 2: grid features reuse columns [0,32)
        if (h == 0) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) tc::tmem_st8_split(t.lane_base + c8 * 8, t.lane_base + S2_ALO + c8 * 8, &gf[c8 * 8]);
        }
        s2_issue(t, w0, 5, 4, true);
        s2_wait(t);
        // ---------------- epilogues (this thread's 32 outputs) and hidden layers
        float s_part = (h == 0) ? bl_sdf : 0.f;
        for (int l = 0; l < n; ++l) {
            const float *bias = smem + lay.bias[l] + c0;
            const bool last = (l + 1 == n);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float a[8], b[8];
                tc::tmem_ld8(t.lane_base + S2_D + c0 + c8 * 8, a);
                tc::tmem_ld8(t.lane_base + S2_D + NICER_W + c0 + c8 * 8, b);
                tc::wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a[i] = softplus100_fast((a[i] + b[i]) + bias[c8 * 8 + i]);
                    if (last) s_part += wl[c8 * 8 + i] * a[i];
                }
                if (!last) tc::tmem_st8_split(t.lane_base + c0 + c8 * 8, t.lane_base + S2_ALO + c0 + c8 * 8, a);
            }
            if (!last) {
                s2_issue(t, tc::smem_u32(smem + lay.w[l + 1]), 0, 8, false);
                s2_wait(t);
            }
        }
        // ---------------- sdf = b + w_n . a_n: the upper half of the dot product crosses through shared memory
        if (h == 1) xch[tile][lane] = s_part;
        asm volatile("bar.sync %0, 256;" ::"r"(t.id) : "memory");
        if (h == 0 && valid) {
            const float s = s_part + xch[tile][lane];
            if (accumulate) sdf[p] += s; else sdf[p] = s;
        }
        // (the next tile's first write to xch comes after two more tile barriers)
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_slot, S2_TMEM);
}

static int g_tc_enabled = -1;
bool tc_enabled() {
    if (g_tc_enabled < 0) {
        const char *e = getenv("NICER_DISABLE_TC");
        g_tc_enabled = (e && e[0] == '1') ? 0 : 1;
    }
    return g_tc_enabled == 1;
}
void set_tc_enabled(int on) { g_tc_enabled = on ? 1 : 0; }

int launch_grid_encode(const nicer_grid_t *g, const float *x, uint32_t P, float *F, float *DYDX, cudaStream_t st);

// F: optional [L*C][P] workspace; when given the gathers run in grid_encode_kernel at full occupancy first
int launch_sdf_only_tc(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t flags, float *sdf, float *F, cudaStream_t st) {
    TcSmemLayout lay = tc_layout((int)net->n_hidden);
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const size_t smem = (size_t)lay.total_floats * sizeof(float);
    const uint32_t tiles = div_up(P, TC_BLOCK);
    const uint32_t grid = tiles < (uint32_t)(2 * num_sms()) ? tiles : (uint32_t)(2 * num_sms());
    const uint32_t acc = (flags & NICER_SDF_ACCUMULATE) ? 1u : 0u;
    static const int tiles_per_sm = [] { const char *e = getenv("NICER_TC_TILES"); return (e && e[0] == '2') ? 2 : 4; }();
    // NICER_TC_SAMPLER=s selects the stacked-operand kernel (measured 0.66 ms against 0.60 ms of tc4 on the fine net: see its header)
    static const bool stacked = [] { const char *e = getenv("NICER_TC_SAMPLER"); return e && e[0] == 's'; }();
    if (tiles_per_sm == 4 && net->multires == 6) {
        if (F && !(flags & NICER_SDF_FEATURES_READY)) {
            if (int e = launch_grid_encode(&net->grid, x, P, F, nullptr, st)) return e;
        }
        if (stacked && F) {
            const S2Layout l2 = s2_layout((int)net->n_hidden);
            const size_t smem2 = (size_t)l2.total_floats * sizeof(float);
            const uint32_t pairs2 = div_up(tiles, 2);
            const uint32_t grid2 = pairs2 < (uint32_t)num_sms() ? pairs2 : (uint32_t)num_sms();
#define LAUNCHS(CC)                                                                                                  \
    do {                                                                                                             \
        NICER_CUDA(cudaFuncSetAttribute(sdf_only_tcs_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2), \
                   "nicer_sdf_forward(tcs)");                                                                        \
        sdf_only_tcs_kernel<CC><<<grid2, S2_THREADS, smem2, st>>>(*net, l2, x, P, acc, sdf, F);                      \
    } while (0)
            switch (net->grid.C) {
                case 2: LAUNCHS(2); break;
                case 4: LAUNCHS(4); break;
                default: LAUNCHS(8); break;
            }
#undef LAUNCHS
            NICER_CHECK_LAUNCH("nicer_sdf_forward(tcs)");
            return 0;
        }
        const uint32_t pairs = div_up(tiles, 2);
        const uint32_t grid4 = pairs < (uint32_t)(2 * num_sms()) ? pairs : (uint32_t)(2 * num_sms());
#define LAUNCH4(CC)                                                                                                  \
    do {                                                                                                             \
        NICER_CUDA(cudaFuncSetAttribute(sdf_only_tc4_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_sdf_forward(tc4)");                                                                        \
        sdf_only_tc4_kernel<CC><<<grid4, T4_THREADS, smem, st>>>(*net, ls, lay, x, P, acc, sdf, F);                  \
    } while (0)
        switch (net->grid.C) {
            case 2: LAUNCH4(2); break;
            case 4: LAUNCH4(4); break;
            default: LAUNCH4(8); break;
        }
#undef LAUNCH4
        NICER_CHECK_LAUNCH("nicer_sdf_forward(tc4)");
        return 0;
    }
#define LAUNCH(CC)                                                                                                  \
    do {                                                                                                            \
        NICER_CUDA(cudaFuncSetAttribute(sdf_only_tc_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_sdf_forward(tc)");                                                                        \
        sdf_only_tc_kernel<CC><<<grid, TC_BLOCK, smem, st>>>(*net, ls, lay, x, P, acc, sdf);                        \
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

}  // namespace nicer

extern "C" int nicer_set_tensor_cores(int enabled) {
    nicer::set_tc_enabled(enabled);
    return 0;
}
