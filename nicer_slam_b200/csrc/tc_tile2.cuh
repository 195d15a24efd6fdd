// Two-threads-per-point tile machinery shared by the tcgen05 MLP kernels (sdf_tc_split.cu, color_tc.cu): a 128-point tile is
// served by 256 threads -- warps w and w + 4 of a tile address the same TMEM lane quarter and split the accumulator columns in
// halves -- two tiles per 512-thread CTA.  See the header of sdf_tc_split.cu.
#pragma once
#include "tc_tile.cuh"

namespace nicer {

constexpr int TCS_THREADS = 512;

struct TcsShared {
    uint64_t bars[2];
    uint32_t tmem_slot;
    float xch[2][128][4];     // [tile][point][..]: partial results handed from one column half to the other
};

__device__ __forceinline__ void tile_sync2(const Tile &t) { asm volatile("bar.sync %0, 256;" ::"r"(t.id) : "memory"); }

// as gemm_issue (tc_tile.cuh) for a 256-thread tile
__device__ __forceinline__ void gemm_issue2(Tile &t, uint32_t whi, uint32_t wlo, int K, int N, bool acc_first = false) {
    tc::wait_st();
    tc::fence_before_sync();
    tile_sync2(t);
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
// barriers + TMEM of a two-tile, 512-thread CTA; returns the calling thread's tile. Call after the operands are staged.
__device__ __forceinline__ Tile tile_setup2(TcsShared &sh, uint32_t alo, uint32_t dcol) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { tc::mbar_init(&sh.bars[0], 1); tc::mbar_init(&sh.bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&sh.tmem_slot, TCF_TMEM);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    Tile t;
    const int tile = tid >> 8;
    t.tmem = sh.tmem_slot + (uint32_t)tile * TCF_TILE_COLS;
    t.lane_base = t.tmem + ((uint32_t)((warp & 3) * 32) << 16);     // warps w and w + 4 of a tile: same lane quarter
    t.bar = &sh.bars[tile];
    t.parity = 0;
    t.id = 1 + tile;
    t.leader = (tid & 255) == 0;
    t.alo = alo;
    t.dcol = dcol;
    return t;
}

__device__ __forceinline__ void tile_teardown2(TcsShared &sh) {
    tc::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(sh.tmem_slot, TCF_TMEM);
}

// 32 values of one saved layer row-block (this thread's column half) for this point, issued together
__device__ __forceinline__ void load32(const float *__restrict__ base, size_t row0, size_t Ps, uint32_t p, float v[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __ldg(base + (row0 + j) * Ps + p);
}
__device__ __forceinline__ void load32_rw(const float *base, size_t row0, size_t Ps, uint32_t p, float v[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = base[(row0 + j) * Ps + p];
}
// this thread's four accumulator chunks (32 columns starting at 32 h)
__device__ __forceinline__ void ld_half(const Tile &t, int c0, float v[32]) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) ld_d8(t, c0 + c8, &v[c8 * 8]);
    tc::wait_ld();
}
__device__ __forceinline__ void st_half(const Tile &t, int c0, const float v[32]) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) st_a8(t, c0 + c8, &v[c8 * 8]);
}

}  // namespace nicer
