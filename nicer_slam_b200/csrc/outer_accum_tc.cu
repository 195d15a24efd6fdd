// Tensor-core (tcgen05, 3xTF32) weight/bias gradient contraction:  C[M,N] += A[M][P] * B[N][P]^T,  bias[M] += rowsum(A).
// Both operands are the feature-major buffers of the fused backward kernels, i.e. already K-major with K = samples.
// The product is computed transposed, D^T[N (128 TMEM lanes)][M (64 columns)] = Bop * Aop^T with the B rows as the
// M=128 operand (rows >= N are zero, row N is all ones so that D^T[N][m] = rowsum(A)[m] comes for free) and the A rows
// as the N operand; both operands go through shared memory in the UMMA K-major no-swizzle layout as (hi, lo) pairs.
//
// The kernel is HBM-bound by arithmetic intensity (16.8 FLOP/B at N = 64), so it is built as a streaming pipeline, one
// persistent CTA per SM, split-K over the CTAs of a job:
//   cp.async (LDGSTS, 16 B) fills a ring of raw fp32 half-stages (32 samples x (M + N) rows) 3-5 half-stages ahead of use;
//   the 256 threads turn one raw half-stage into the (hi, lo) tf32 operand tiles of one of TWO operand buffers
//   (raw rows are padded to 144 B so that both the 8-row x 16-B reads and the core-matrix writes are bank-conflict free);
//   one thread issues the 8 tcgen05.mma of that half-stage (B_hi x [A_hi ; A_lo] as ONE N = 128 MMA + B_lo x A_hi, i.e.
//   2 instead of 3 MMAs per K-step: a tcgen05.mma costs the same ~102 cycles for N = 64 and N = 128) and commits them to
//   the buffer's mbarrier, so the MMAs of half-stage h run under the split of h + 1 and the loads of h + 2 ... h + 5.
// Round 1's version (register prefetch of ONE 64-sample stage, single operand buffer) ran at 2.0 TB/s: ncu showed the
// warps waiting on the next stage's loads (long scoreboard) with nothing else in flight.
#include "common.cuh"
#include "tc_common.cuh"

namespace nicer {

constexpr int OT_THREADS = 256;
constexpr int OT_HS = 32;                       // samples per half-stage
constexpr int OT_CH = OT_HS / 4;                // 16-byte chunks per row per half-stage (8)
constexpr int OT_BROWS = 128, OT_AROWS = 64;
constexpr int OT_RAW_STRIDE = OT_HS * 4 + 16;   // bytes per raw row (padded: 144)
constexpr int OT_MAX_SLOTS = 6;
constexpr int OT_SMEM_LIMIT = 227 * 1024;

struct OtOperands {                             // one operand buffer (48 KB)
    float bhi[OT_CH * OT_BROWS * 4], blo[OT_CH * OT_BROWS * 4];
    // A rows as ONE 128-row operand per chunk: rows [0,64) hold the hi halves, rows [64,128) the lo halves
    float acomb[OT_CH * 2 * OT_AROWS * 4];
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
__device__ __forceinline__ void cp_async_wait(int pending) {       // at most `pending` groups still in flight
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    }
}

__global__ void __launch_bounds__(OT_THREADS, 1)
outer_accum_tc_kernel(const OtJobs js, uint32_t P, uint32_t halves_per_cta, uint32_t ctas_per_job, uint32_t n_slots, uint32_t rows_b_pad) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    OtOperands *ops = reinterpret_cast<OtOperands *>(smem_raw);                  // [2]
    unsigned char *ring = smem_raw + 2 * sizeof(OtOperands);                     // n_slots x slot_bytes
    __shared__ __align__(8) uint64_t bar_free[2], bar_done;
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
    const uint32_t slot_bytes = (OT_AROWS + rows_b_pad) * OT_RAW_STRIDE;
    const uint32_t a_items = M * OT_CH, items = (M + N) * OT_CH;                 // 16-byte pieces of one half-stage
    const uint32_t a_groups = (M + 7) / 8, groups = a_groups + (N + 7) / 8;      // 8-row groups (split work)

    // one-time: zero both operand buffers (rows that are never written must stay 0), ones row at index N, barriers, TMEM
    for (int i = tid; i < (int)(2 * sizeof(OtOperands) / 4); i += OT_THREADS) reinterpret_cast<float *>(ops)[i] = 0.f;
    __syncthreads();
    if (bias && N < OT_BROWS) {
        for (int i = tid; i < 2 * OT_CH * 4; i += OT_THREADS) {
            const int b = i / (OT_CH * 4), j = i - b * (OT_CH * 4);
            ops[b].bhi[((j >> 2) * OT_BROWS + N) * 4 + (j & 3)] = 1.0f;          // exact in tf32
        }
    }
    if (tid == 0) { tc::mbar_init(&bar_free[0], 1); tc::mbar_init(&bar_free[1], 1); tc::mbar_init(&bar_done, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t ring_u32 = tc::smem_u32(ring);

    auto issue_load = [&](uint32_t h) {          // h: half-stage index relative to h0
        if (h < nh) {
            const uint32_t p0 = (h0 + h) * OT_HS;
            const uint32_t dst0 = ring_u32 + (h % n_slots) * slot_bytes;
            for (uint32_t i = tid; i < items; i += OT_THREADS) {
                const bool isA = i < a_items;
                const uint32_t k = isA ? i : i - a_items;
                const uint32_t r = k >> 3, c = k & 7;
                const uint32_t p = p0 + c * 4;
                const float *src = isA ? A + (size_t)r * lda + p : B + (size_t)r * ldb + p;
                const uint32_t row = isA ? r : OT_AROWS + r;
                const bool ok = p < P;                                  // P % 4 == 0: a piece is either whole or absent
                cp_async16(dst0 + row * OT_RAW_STRIDE + c * 16, ok ? (const void *)src : (const void *)A, ok ? 16u : 0u);
            }
        }
        cp_async_commit();                       // every thread commits one group per half-stage, loads or not
    };

    const uint32_t ahead = n_slots - 1;
    for (uint32_t h = 0; h < ahead; ++h) issue_load(h);
    for (uint32_t h = 0; h < nh; ++h) {
        cp_async_wait((int)ahead - 1);           // the group of half-stage h has landed (this thread's pieces)
        __syncthreads();                         // ... and everybody else's
        const uint32_t b = h & 1u;
        if (h >= 2) tc::mbar_wait(&bar_free[b], ((h >> 1) - 1u) & 1u);           // MMAs of half-stage h-2 are done with buffer b
        OtOperands &op = ops[b];
        const unsigned char *raw = ring + (size_t)(h % n_slots) * slot_bytes;
        // split: item = (8-row group, chunk, row in group); lanes 0-7 = 8 consecutive rows of one chunk
        for (uint32_t i = tid; i < groups * 64; i += OT_THREADS) {
            const uint32_t g = i >> 6, c = (i >> 3) & 7, r8 = i & 7;
            const bool isA = g < a_groups;
            const uint32_t r = (isA ? g : g - a_groups) * 8 + r8;
            if (r >= (isA ? M : N)) continue;
            const float4 v = *reinterpret_cast<const float4 *>(raw + (size_t)((isA ? 0 : OT_AROWS) + r) * OT_RAW_STRIDE + c * 16);
            float4 hh, ll;
            hh.x = tc::tf32_hi(v.x); hh.y = tc::tf32_hi(v.y); hh.z = tc::tf32_hi(v.z); hh.w = tc::tf32_hi(v.w);
            ll.x = v.x - hh.x; ll.y = v.y - hh.y; ll.z = v.z - hh.z; ll.w = v.w - hh.w;
            float *dh = isA ? op.acomb + ((size_t)c * 2 * OT_AROWS + r) * 4 : op.bhi + ((size_t)c * OT_BROWS + r) * 4;
            float *dl = isA ? op.acomb + ((size_t)c * 2 * OT_AROWS + OT_AROWS + r) * 4 : op.blo + ((size_t)c * OT_BROWS + r) * 4;
            *reinterpret_cast<float4 *>(dh) = hh;
            *reinterpret_cast<float4 *>(dl) = ll;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            constexpr uint32_t IDESC128 = tc::idesc_tf32(128, 2 * OT_AROWS), IDESC64 = tc::idesc_tf32(128, OT_AROWS);
            const uint32_t bh = tc::smem_u32(op.bhi), bl = tc::smem_u32(op.blo), ac = tc::smem_u32(op.acomb);
#pragma unroll
            for (int ks = 0; ks < OT_HS / 8; ++ks) {
                const uint64_t dbh = tc::smem_desc(bh + ks * 2 * OT_BROWS * 16, OT_BROWS * 16, 128);
                const uint64_t dbl = tc::smem_desc(bl + ks * 2 * OT_BROWS * 16, OT_BROWS * 16, 128);
                // the combined operand (128 rows per chunk); its first 64 rows alone are A_hi (same chunk stride)
                const uint64_t dac = tc::smem_desc(ac + ks * 2 * (2 * OT_AROWS) * 16, 2 * OT_AROWS * 16, 128);
                // columns [0,64) += B_hi A_hi^T, columns [64,128) += B_hi A_lo^T ; then columns [0,64) += B_lo A_hi^T
                tc::mma_tf32_ss(tmem, dbh, dac, IDESC128, (h == 0 && ks == 0) ? 0u : 1u);
                tc::mma_tf32_ss(tmem, dbl, dac, IDESC64, 1u);
            }
            tc::mma_commit(&bar_free[b]);
        }
        issue_load(h + ahead);                   // its slot was read (by everyone) before the barrier above, one iteration ago
    }
    cp_async_wait(0);
    if (tid == 0 && nh > 0) tc::mma_commit(&bar_done);
    // epilogue: lane n of the accumulator holds D^T[n][0..63]
    if (nh > 0) {
        tc::mbar_wait(&bar_done, 0);
        __syncwarp();
        tc::fence_after_sync();
    }
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
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

bool tc_enabled();

static bool oa_tc_ok(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N, uint32_t P) {
    if (M > OT_AROWS || N >= OT_BROWS || P < 4096 || (P & 3u)) return false;
    if ((lda & 3u) || (ldb & 3u) || (reinterpret_cast<uintptr_t>(A) & 15u) || (reinterpret_cast<uintptr_t>(B) & 15u)) return false;
    return true;
}

static int oa_launch(const OtJobs &js, uint32_t n_jobs, uint32_t P, cudaStream_t st) {
    const uint32_t n_half = div_up(P, OT_HS);
    uint32_t n_max = 0;
    for (uint32_t j = 0; j < n_jobs; ++j) n_max = js.N[j] > n_max ? js.N[j] : n_max;
    const uint32_t rows_b_pad = (n_max + 7u) & ~7u;
    const uint32_t slot = (OT_AROWS + rows_b_pad) * OT_RAW_STRIDE;
    uint32_t n_slots = (uint32_t)((OT_SMEM_LIMIT - 2 * sizeof(OtOperands) - 256) / slot);
    if (n_slots > OT_MAX_SLOTS) n_slots = OT_MAX_SLOTS;
    if (n_slots < 3) NICER_FAIL(-1, "nicer_outer_accum(tc): operand too wide for the shared-memory ring");
    uint32_t per_job = (uint32_t)num_sms() / n_jobs;            // one persistent CTA per SM, split over the jobs
    if (per_job == 0) per_job = 1;
    if (per_job > n_half) per_job = n_half;
    const uint32_t hpc = div_up(n_half, per_job);
    per_job = div_up(n_half, hpc);
    const size_t smem = 2 * sizeof(OtOperands) + (size_t)n_slots * slot;
    NICER_CUDA(cudaFuncSetAttribute(outer_accum_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "nicer_outer_accum(tc)");
    outer_accum_tc_kernel<<<per_job * n_jobs, OT_THREADS, smem, st>>>(js, P, hpc, per_job, n_slots, rows_b_pad);
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
