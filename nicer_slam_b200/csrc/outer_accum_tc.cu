// Tensor-core (tcgen05, 3xTF32) weight/bias gradient contraction:  C[M,N] += A[M][P] * B[N][P]^T,  bias[M] += rowsum(A).
// Both operands are the feature-major buffers of the fused backward kernels, i.e. already K-major with K = samples.
// The product is computed transposed, D^T[N (128 TMEM lanes)][M (64 columns)] = Bop * Aop^T: the B rows are the M = 128
// operand (lane n = row n; rows > N are zero, row N is all ones so that D^T[N][m] = rowsum(A)[m] comes for free) and the A rows
// are the N operand.
//
// The kernel is HBM-bound by arithmetic intensity (16.8 FLOP/B at N = 64), so it is a streaming pipeline, one persistent
// CTA per SM (split-K over the CTAs of a job), in half-stages of 32 samples, warp-specialised:
//   warps 0-3   "row warps": thread n owns row n of B.  It loads its 32 samples (8 x LDG.128, two half-stages ahead, in
//               registers), splits them into (hi, lo) tf32 and writes them with tcgen05.st into the A-operand columns of
//               TMEM (two buffers of 32 + 32 columns): the M = 128 operand never touches shared memory.
//   warps 4-14  "A warps": cp.async (LDGSTS, 16 B) fills a ring of raw fp32 half-stages of the A rows 3-5 half-stages ahead;
//               the warps split one raw half-stage into one of TWO shared-memory operand tiles [A_hi ; A_lo] (128 rows per
//               16-byte K chunk, UMMA K-major no-swizzle; raw rows padded to 144 B: reads and writes bank-conflict free).
//   warp 15     one thread issues the 8 tcgen05.mma of a half-stage when both halves of the operand pair are ready
//               (mbarrier with 15 warp arrivals): B_hi x [A_hi ; A_lo] as ONE N = 128 MMA + B_lo x A_hi (N = 64), i.e. 2 instead
//               of 3 MMAs per K-step (a tcgen05.mma costs the same ~102 cycles for N = 64 and N = 128), and commits them to the
//               buffer's "free" mbarrier.  It takes no part in the loads, so the back-pressure of the MMA queue never stalls them.
// History (profiles/r02_wgrad_notes.txt): round 1 (register prefetch of one 64-sample stage, one operand buffer, 2 CTAs/SM)
// 2.0 TB/s, warps waiting on the next stage's loads; cp.async ring + two operand buffers but the MMA-issuing thread inside
// the producer barrier: 1.8-2.2 TB/s (issue back-pressure serialised with the split); dedicated MMA warp, both operands through
// shared memory: 2.4 TB/s, bound by shared-memory bandwidth (~120 KB moved per 16 KB half-stage: raw ring in + out, hi + lo
// tiles, MMA operand fetches).  This version moves ~56 KB of shared memory per half-stage.
#include <cstddef>

#include "common.cuh"
#include "tc_common.cuh"

namespace nicer {

constexpr int OT_THREADS = 512;
constexpr int OT_HS = 64;                       // samples per stage
constexpr int OT_CH = OT_HS / 4;                // 16-byte chunks per row per stage (16)
constexpr int OT_BROWS = 128, OT_AROWS = 64;
constexpr int OT_RAW_STRIDE = OT_HS * 4 + 16;   // bytes per raw row (padded: 272 = 68 words, 68 mod 32 = 4)
constexpr int OT_SLOT = OT_AROWS * OT_RAW_STRIDE;   // raw ring slot (A rows only): 17408 B
constexpr int OT_SLOTS = 4;
constexpr int OT_ROW_WARPS = 8, OT_A_WARPS = 7;     // + 1 MMA warp = 16.  Row warp w: TMEM lane quarter w % 4, sample half w / 4
constexpr int OT_A_THREADS = OT_A_WARPS * 32;       // 224
constexpr int OT_A_ITEMS = 5;                       // 64 rows x 16 chunks / 224 threads
constexpr int OT_NBUF = 3;                          // operand buffers (TMEM rows + smem tile): write / MMA overlap needs > 2
constexpr int OT_TMEM_COLS = 512;                   // D: [0,128)   A operand buffer b: hi [128 + 128 b, +64), lo [192 + 128 b, +64)

struct OtSmem {
    float acomb[OT_NBUF][OT_CH * 2 * OT_AROWS * 4]; // per buffer: chunk-major, rows [0,64) = A_hi, rows [64,128) = A_lo   (3 x 32 KB)
    unsigned char ring[OT_SLOTS * OT_SLOT];         // raw fp32 A rows                                                   (54 KB)
};

// up to OT_MAX_JOBS contractions over the same sample range in one launch (the weight gradients of one network backward):
// CTAs [j * ctas_per_job, (j+1) * ctas_per_job) split the half-stages of job j
constexpr int OT_MAX_JOBS = 8;
struct OtJobs {
    const float *A[OT_MAX_JOBS], *B[OT_MAX_JOBS];
    float *C[OT_MAX_JOBS], *bias[OT_MAX_JOBS];
    uint32_t lda[OT_MAX_JOBS], ldb[OT_MAX_JOBS], ldc[OT_MAX_JOBS], M[OT_MAX_JOBS], N[OT_MAX_JOBS];
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// half a stage (32 samples) of one B row: 8 float4 from global (zeros beyond P)
constexpr int OT_RC = OT_CH / 2;
struct RowRegs { float4 v[OT_RC]; };
__device__ __forceinline__ void row_load(RowRegs &r, const float *__restrict__ src, uint32_t p0, uint32_t P, bool active, float fill) {
#pragma unroll
    for (int c = 0; c < OT_RC; ++c) {
        r.v[c] = make_float4(fill, fill, fill, fill);
        if (active) {
            r.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p0 + c * 4 < P) r.v[c] = __ldg(reinterpret_cast<const float4 *>(src + p0 + c * 4));
        }
    }
}
// split and write to the A-operand columns of this thread's TMEM lane
__device__ __forceinline__ void row_store(const RowRegs &r, uint32_t t_hi, uint32_t t_lo) {
#pragma unroll
    for (int c2 = 0; c2 < OT_RC / 2; ++c2) {
        const float v[8] = {r.v[2 * c2].x, r.v[2 * c2].y, r.v[2 * c2].z, r.v[2 * c2].w,
                            r.v[2 * c2 + 1].x, r.v[2 * c2 + 1].y, r.v[2 * c2 + 1].z, r.v[2 * c2 + 1].w};
        tc::tmem_st8_split(t_hi + c2 * 8, t_lo + c2 * 8, v);
    }
}

__global__ void __launch_bounds__(OT_THREADS, 1)
outer_accum_tc_kernel(const OtJobs js, uint32_t P, uint32_t halves_per_cta, uint32_t ctas_per_job) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    OtSmem &sm = *reinterpret_cast<OtSmem *>(smem_raw);
    __shared__ __align__(8) uint64_t bar_free[OT_NBUF], bar_full[OT_NBUF], bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t job = blockIdx.x / ctas_per_job, cta = blockIdx.x - job * ctas_per_job;
    const float *__restrict__ A = js.A[job], *__restrict__ B = js.B[job];
    float *C = js.C[job], *bias = js.bias[job];
    const uint32_t lda = js.lda[job], ldb = js.ldb[job], ldc = js.ldc[job], M = js.M[job], N = js.N[job];
    const uint32_t n_half = (P + OT_HS - 1) / OT_HS;
    const uint32_t h0 = cta * halves_per_cta;
    const uint32_t h1 = (h0 + halves_per_cta < n_half) ? h0 + halves_per_cta : n_half;
    const uint32_t nh = h1 > h0 ? h1 - h0 : 0;

    // one-time: zero both A tiles (rows >= M are never written), barriers, TMEM
    for (int i = tid; i < (int)(sizeof(sm.acomb) / 4); i += OT_THREADS) reinterpret_cast<float *>(sm.acomb)[i] = 0.f;
    if (tid == 0) {
        for (int b = 0; b < OT_NBUF; ++b) { tc::mbar_init(&bar_free[b], 1); tc::mbar_init(&bar_full[b], OT_ROW_WARPS + OT_A_WARPS); }
        tc::mbar_init(&bar_done, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, OT_TMEM_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp < OT_ROW_WARPS) {
        // ---------------- row warps: B row n -> TMEM lane n; warps w and w + 4 share a lane quarter and take 16 samples each
        const uint32_t n = (uint32_t)((warp & 3) * 32 + lane), half = (uint32_t)(warp >> 2);
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * (OT_HS / 2);
        const bool active = n < N;
        const float *src = B + (size_t)(active ? n : 0) * ldb + half * (OT_HS / 2);
        // lanes without a B row carry a constant: the ones row (1.0 splits into hi = 1, lo = 0) when a bias is wanted, else zeros.
        // tcgen05.st is warp-collective (.sync.aligned), so every lane stores every half-stage, active or not.
        const float fill = (bias && n == N) ? 1.0f : 0.0f;
        const uint32_t Pq = P > half * (OT_HS / 2) ? P - half * (OT_HS / 2) : 0;      // bound for this warp's shifted sample index
        RowRegs r0, r1;
        row_load(r0, src, h0 * OT_HS, Pq, active, fill);
        row_load(r1, src, (h0 + 1) * OT_HS, Pq, active, fill);
        auto step = [&](uint32_t h, RowRegs &r) {        // loads run two stages ahead of the store
            const uint32_t b = h % OT_NBUF;
            if (h >= OT_NBUF) tc::mbar_wait(&bar_free[b], ((h / OT_NBUF) - 1u) & 1u);
            row_store(r, lane_base + 128 + 2 * OT_HS * b, lane_base + 128 + OT_HS + 2 * OT_HS * b);
            tc::wait_st();
            row_load(r, src, (h0 + h + 2) * OT_HS, Pq, active && h + 2 < nh, fill);
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[b]);
        };
        for (uint32_t h = 0; h < nh; h += 2) {
            step(h, r0);
            if (h + 1 < nh) step(h + 1, r1);
        }
    } else if (warp < OT_ROW_WARPS + OT_A_WARPS) {
        // ---------------- A warps: raw ring -> [A_hi ; A_lo] tiles.  Per-thread work is fixed for the whole kernel:
        //   load items  : 16-byte piece (row, chunk) -> global pointer (advances 32 samples per half-stage), ring offset
        //   split items : (8-row group, chunk, row in group); lanes 0-7 = 8 consecutive rows of one chunk -> raw / tile offsets
        const uint32_t at = (uint32_t)tid - OT_ROW_WARPS * 32;
        const uint32_t ring_u32 = tc::smem_u32(sm.ring);
        const float *ld_src[OT_A_ITEMS];
        uint32_t ld_dst[OT_A_ITEMS], ld_p[OT_A_ITEMS], sp_raw[OT_A_ITEMS], sp_hi[OT_A_ITEMS];
#pragma unroll
        for (int k = 0; k < OT_A_ITEMS; ++k) {
            const uint32_t i = at + k * OT_A_THREADS;
            ld_src[k] = nullptr; ld_dst[k] = 0; ld_p[k] = 0xffffffffu; sp_raw[k] = 0xffffffffu; sp_hi[k] = 0;
            if (i < M * OT_CH) {
                const uint32_t r = i / OT_CH, c = i % OT_CH;
                ld_p[k] = h0 * OT_HS + c * 4;
                ld_src[k] = A + (size_t)r * lda + ld_p[k];
                ld_dst[k] = r * OT_RAW_STRIDE + c * 16;
            }
            if (i < OT_AROWS * OT_CH) {
                const uint32_t g = i / (8 * OT_CH), c = (i >> 3) % OT_CH, r = g * 8 + (i & 7);
                if (r < M) {
                    sp_raw[k] = r * OT_RAW_STRIDE + c * 16;
                    sp_hi[k] = (c * 2 * OT_AROWS + r) * 16;
                }
            }
        }
        uint32_t ld_slot = 0, ld_h = 0;
        auto issue_load = [&]() {
            if (ld_h < nh) {
                const uint32_t dst0 = ring_u32 + ld_slot * OT_SLOT;
#pragma unroll
                for (int k = 0; k < OT_A_ITEMS; ++k) {
                    if (ld_src[k]) {
                        const bool ok = ld_p[k] < P;                    // P % 4 == 0: a piece is either whole or absent
                        cp_async16(dst0 + ld_dst[k], ok ? (const void *)ld_src[k] : (const void *)A, ok ? 16u : 0u);
                        ld_src[k] += OT_HS;
                        ld_p[k] += OT_HS;
                    }
                }
            }
            cp_async_commit();                   // every thread commits one group per half-stage, loads or not
            ++ld_h;
            ld_slot = (ld_slot + 1 == OT_SLOTS) ? 0 : ld_slot + 1;
        };
#pragma unroll
        for (int h = 0; h < OT_SLOTS - 1; ++h) issue_load();
        uint32_t use_slot = 0;
        for (uint32_t h = 0; h < nh; ++h) {
            asm volatile("cp.async.wait_group %0;" ::"n"(OT_SLOTS - 2) : "memory");     // half-stage h has landed (this thread's pieces)
            asm volatile("bar.sync 1, %0;" ::"n"(OT_A_THREADS) : "memory");            // ... and every other A thread's
            issue_load();                        // into the slot everybody finished reading before that barrier (half-stage h-1)
            const uint32_t b = h % OT_NBUF;
            if (h >= OT_NBUF) tc::mbar_wait(&bar_free[b], ((h / OT_NBUF) - 1u) & 1u);   // MMAs of stage h-3 are done with tile b
            const unsigned char *raw = sm.ring + (size_t)use_slot * OT_SLOT;
            unsigned char *tile = reinterpret_cast<unsigned char *>(sm.acomb[b]);
            use_slot = (use_slot + 1 == OT_SLOTS) ? 0 : use_slot + 1;
#pragma unroll
            for (int k = 0; k < OT_A_ITEMS; ++k) {
                if (sp_raw[k] != 0xffffffffu) {
                    const float4 v = *reinterpret_cast<const float4 *>(raw + sp_raw[k]);
                    float4 hh, ll;
                    hh.x = tc::tf32_hi(v.x); hh.y = tc::tf32_hi(v.y); hh.z = tc::tf32_hi(v.z); hh.w = tc::tf32_hi(v.w);
                    ll.x = v.x - hh.x; ll.y = v.y - hh.y; ll.z = v.z - hh.z; ll.w = v.w - hh.w;
                    *reinterpret_cast<float4 *>(tile + sp_hi[k]) = hh;
                    *reinterpret_cast<float4 *>(tile + sp_hi[k] + OT_AROWS * 16) = ll;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[b]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (lane == 0) {
        // ---------------- MMA warp: one thread turns every filled (TMEM rows, smem tile) pair into 8 tcgen05.mma
        constexpr uint32_t IDESC128 = tc::idesc_tf32(128, 2 * OT_AROWS), IDESC64 = tc::idesc_tf32(128, OT_AROWS);
        for (uint32_t h = 0; h < nh; ++h) {
            const uint32_t b = h % OT_NBUF;
            tc::mbar_wait(&bar_full[b], (h / OT_NBUF) & 1u);
            tc::fence_after_sync();
            const uint32_t ac = tc::smem_u32(sm.acomb[b]);
            const uint32_t a_hi = tmem + 128 + 2 * OT_HS * b, a_lo = a_hi + OT_HS;
#pragma unroll
            for (int ks = 0; ks < OT_HS / 8; ++ks) {
                // the combined operand (128 rows per chunk); its first 64 rows alone are A_hi (same chunk stride)
                const uint64_t dac = tc::smem_desc(ac + ks * 2 * (2 * OT_AROWS) * 16, 2 * OT_AROWS * 16, 128);
                // columns [0,64) += B_hi A_hi^T, columns [64,128) += B_hi A_lo^T ; then columns [0,64) += B_lo A_hi^T
                tc::mma_tf32_ts(tmem, a_hi + ks * 8, dac, IDESC128, (h == 0 && ks == 0) ? 0u : 1u);
                tc::mma_tf32_ts(tmem, a_lo + ks * 8, dac, IDESC64, 1u);
            }
            tc::mma_commit(&bar_free[b]);
        }
        if (nh > 0) tc::mma_commit(&bar_done);
    }
    if (nh > 0) {
        tc::mbar_wait(&bar_done, 0);
        __syncwarp();
        tc::fence_after_sync();
    }
    // epilogue: lane n of the accumulator holds D^T[n][0..63]
    if (nh > 0 && warp < 4) {
        const uint32_t n = warp * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8], w[8];
            tc::tmem_ld8(lane_base + c8 * 8, v);
            tc::tmem_ld8(lane_base + OT_AROWS + c8 * 8, w);
            tc::wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] += w[i];
                const uint32_t m = c8 * 8 + i;
                if (m < M) {
                    if (n < N) atomicAdd(&C[(size_t)m * ldc + n], v[i]);
                    else if (n == N && bias) atomicAdd(&bias[m], v[i]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, OT_TMEM_COLS);
}

bool tc_enabled();

static bool oa_tc_ok(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N, uint32_t P) {
    if (M > OT_AROWS || N >= OT_BROWS || P < 4096 || (P & 3u)) return false;
    if ((lda & 3u) || (ldb & 3u) || (reinterpret_cast<uintptr_t>(A) & 15u) || (reinterpret_cast<uintptr_t>(B) & 15u)) return false;
    return true;
}

static int oa_launch(const OtJobs &js, uint32_t n_jobs, uint32_t P, cudaStream_t st) {
    const uint32_t n_half = div_up(P, OT_HS);
    uint32_t per_job = (uint32_t)num_sms() / n_jobs;            // one persistent CTA per SM, split over the jobs
    if (per_job == 0) per_job = 1;
    if (per_job > n_half) per_job = n_half;
    const uint32_t hpc = div_up(n_half, per_job);
    per_job = div_up(n_half, hpc);
    const size_t smem = sizeof(OtSmem);
    NICER_CUDA(cudaFuncSetAttribute(outer_accum_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "nicer_outer_accum(tc)");
    outer_accum_tc_kernel<<<per_job * n_jobs, OT_THREADS, smem, st>>>(js, P, hpc, per_job);
    NICER_CHECK_LAUNCH("nicer_outer_accum(tc)");
    return 0;
}

// returns 1 when the tensor-core kernel was launched, 0 when the shapes do not qualify (caller falls back), < 0 on error
int launch_outer_accum_tc(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N, uint32_t P, float *C,
                          uint32_t ldc, float *bias, cudaStream_t st) {
    if (!tc_enabled() || !oa_tc_ok(A, lda, M, B, ldb, N, P)) return 0;
    OtJobs js{};
    js.A[0] = A; js.B[0] = B; js.C[0] = C; js.bias[0] = bias;
    js.lda[0] = lda; js.ldb[0] = ldb; js.ldc[0] = ldc; js.M[0] = M; js.N[0] = N;
    if (int e = oa_launch(js, 1, P, st)) return e;
    return 1;
}

// all jobs of a batch in as few launches as possible; jobs the tensor-core kernel does not cover go one by one through
// nicer_outer_accum's fallback
int launch_outer_accum_batch_tc(const nicer_oa_job_t *jobs, uint32_t n_jobs, uint32_t P, uint8_t *done, cudaStream_t st) {
    if (!tc_enabled()) return 0;
    OtJobs js{};
    uint32_t k = 0;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const nicer_oa_job_t &q = jobs[j];
        done[j] = 0;
        if (!oa_tc_ok(q.A, q.lda, q.M, q.B, q.ldb, q.N, P)) continue;
        js.A[k] = q.A; js.B[k] = q.B; js.C[k] = q.C; js.bias[k] = q.bias;
        js.lda[k] = q.lda; js.ldb[k] = q.ldb; js.ldc[k] = q.ldc; js.M[k] = q.M; js.N[k] = q.N;
        done[j] = 1;
        if (++k == OT_MAX_JOBS) {
            if (int e = oa_launch(js, k, P, st)) return e;
            k = 0;
        }
    }
    if (k) {
        if (int e = oa_launch(js, k, P, st)) return e;
    }
    return 0;
}

}  // namespace nicer
