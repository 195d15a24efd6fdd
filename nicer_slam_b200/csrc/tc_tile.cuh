// Shared machinery of the tcgen05 kernels: a 128-point tile (one TMEM lane per point, A operand hi/lo column
// ranges + accumulator columns), tile-level GEMM issue / wait, TMEM <-> register helpers, operand staging.
#pragma once
#include "common.cuh"
#include "nicer_math.cuh"
#include "tc_common.cuh"

namespace nicer {

constexpr int TCF_THREADS = 256;    // two 128-point tiles per CTA share one copy of the staged operands
constexpr int TCF_TILE_COLS = 256, TCF_TMEM = 512;

// One staged B operand (hi/lo) of a kernel
struct MatSpec {
    int layer;       // index into net.W
    int row0, w_rows, w_cols;   // source window: rows [row0, row0 + w_rows) of W (row-major, w_cols columns)
    int rows, K;     // operand shape: `rows` rows (N of the MMA), K contraction columns (multiple of 8)
    int transposed, colmap;     // colmap: 0 identity, otherwise a kernel-specific column permutation id
    int col0;        // first (mapped) column of the window handed to this operand
    int hi, lo;      // float offsets in dynamic shared memory
};

struct Tile {
    uint32_t tmem;        // TMEM address of the tile's column 0, lane 0
    uint32_t lane_base;   // + this warp's lane quarter
    uint64_t *bar;
    uint32_t parity;
    int id;               // named barrier id (1 or 2)
    bool leader;
    uint32_t alo, dcol;   // column offsets of the A-lo range and of the accumulator inside the tile
};

__device__ __forceinline__ void tile_sync(const Tile &t) { asm volatile("bar.sync %0, 128;" ::"r"(t.id) : "memory"); }

// D[128 x N] = A[128 x K] * B^T:  A = (hi, lo) column ranges of the tile (TMEM), B = N rows x K, K-major in shared memory.
// Split in issue / wait so that the caller can put the global loads its epilogue needs in flight while the MMAs run
// (the tcgen05 asm statements are compiler barriers for memory operations: loads are not hoisted across them).
__device__ __forceinline__ void gemm_issue(Tile &t, uint32_t whi, uint32_t wlo, int K, int N, bool acc_first = false) {
    tc::wait_st();
    tc::fence_before_sync();
    tile_sync(t);
    if (t.leader) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_tf32(128, (uint32_t)N);
        const uint32_t chunk = (uint32_t)N * 16u;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t bhi = tc::smem_desc(whi + (uint32_t)ks * 2u * chunk, chunk, 128u);
            const uint64_t blo = tc::smem_desc(wlo + (uint32_t)ks * 2u * chunk, chunk, 128u);
            const uint32_t ahi = t.tmem + ks * 8, alo = t.tmem + t.alo + ks * 8;
            tc::mma_tf32_ts(t.tmem + t.dcol, ahi, bhi, idesc, (ks > 0 || acc_first) ? 1u : 0u);
            tc::mma_tf32_ts(t.tmem + t.dcol, alo, bhi, idesc, 1u);
            tc::mma_tf32_ts(t.tmem + t.dcol, ahi, blo, idesc, 1u);
        }
        tc::mma_commit(t.bar);
    }
}
__device__ __forceinline__ void gemm_wait(Tile &t) {
    tc::mbar_wait(t.bar, t.parity);
    t.parity ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
}
__device__ __forceinline__ void tile_gemm(Tile &t, uint32_t whi, uint32_t wlo, int K, int N) {
    gemm_issue(t, whi, wlo, K, N);
    gemm_wait(t);
}

__device__ __forceinline__ void ld_d8(const Tile &t, int c8, float v[8]) { tc::tmem_ld8(t.lane_base + t.dcol + c8 * 8, v); }
__device__ __forceinline__ void st_a8(const Tile &t, int c8, const float v[8]) {
    tc::tmem_st8_split(t.lane_base + c8 * 8, t.lane_base + t.alo + c8 * 8, v);
}

// C (2, 4 or 8) consecutive A columns starting at a multiple of C, hi/lo split
template <int C>
__device__ __forceinline__ void st_a_small(const Tile &t, int col, const float v[C]) {
    float h[C], l[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { h[i] = tc::tf32_hi(v[i]); l[i] = v[i] - h[i]; }
    const uint32_t ahi = t.lane_base + col, alo = t.lane_base + t.alo + col;
    if constexpr (C == 8) {
        tc::tmem_st8(ahi, h);
        tc::tmem_st8(alo, l);
    } else if constexpr (C == 4) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ahi), "r"(__float_as_uint(h[0])),
                     "r"(__float_as_uint(h[1])), "r"(__float_as_uint(h[2])), "r"(__float_as_uint(h[3])) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(alo), "r"(__float_as_uint(l[0])),
                     "r"(__float_as_uint(l[1])), "r"(__float_as_uint(l[2])), "r"(__float_as_uint(l[3])) : "memory");
    } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ahi), "r"(__float_as_uint(h[0])),
                     "r"(__float_as_uint(h[1])) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(alo), "r"(__float_as_uint(l[0])),
                     "r"(__float_as_uint(l[1])) : "memory");
    }
}

struct TcfShared {
    uint64_t bars[2];
    uint32_t tmem_slot;
};

// small-width accumulator reads (C = 2, 4 or 8 consecutive columns)
template <int C>
__device__ __forceinline__ void ld_d_small(const Tile &t, int col, float v[C]) {
    const uint32_t a = t.lane_base + t.dcol + col;
    if constexpr (C == 8) {
        tc::tmem_ld8(a, v);
    } else if constexpr (C == 4) {
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
        v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    } else {
        uint32_t r0, r1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a) : "memory");
        v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
    }
}

// barriers + TMEM of a two-tile CTA; returns the calling thread's tile. Call after the operands are staged.
__device__ __forceinline__ Tile tile_setup(TcfShared &sh, uint32_t alo, uint32_t dcol) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { tc::mbar_init(&sh.bars[0], 1); tc::mbar_init(&sh.bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&sh.tmem_slot, TCF_TMEM);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    Tile t;
    const int tile = tid >> 7;
    t.tmem = sh.tmem_slot + (uint32_t)tile * TCF_TILE_COLS;
    t.lane_base = t.tmem + ((uint32_t)((warp & 3) * 32) << 16);
    t.bar = &sh.bars[tile];
    t.parity = 0;
    t.id = 1 + tile;
    t.leader = (tid & 127) == 0;
    t.alo = alo;
    t.dcol = dcol;
    return t;
}

__device__ __forceinline__ void tile_teardown(TcfShared &sh) {
    tc::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(sh.tmem_slot, TCF_TMEM);
}

// all 64 values of one saved layer row-block for this point (issued together: 64 loads in flight)
__device__ __forceinline__ void load64(const float *__restrict__ base, size_t row0, size_t Ps, uint32_t p, float v[NICER_W]) {
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) v[j] = __ldg(base + (row0 + j) * Ps + p);
}
// same through the coherent path (for a buffer the kernel also writes)
__device__ __forceinline__ void load64_rw(const float *base, size_t row0, size_t Ps, uint32_t p, float v[NICER_W]) {
#pragma unroll
    for (int j = 0; j < NICER_W; ++j) v[j] = base[(row0 + j) * Ps + p];
}

}  // namespace nicer
