// Tensor-core (tcgen05, 3xTF32) weight/bias gradient contraction:  C[M,N] += A[M][P] * B[N][P]^T,  bias[M] += rowsum(A).
// Both operands are the feature-major buffers of the fused backward kernels, i.e. already K-major with K = samples.
// The product is computed transposed, D^T[N (128 TMEM lanes)][M (64 columns)] = Bop * Aop^T with the B rows as the
// M=128 operand (rows >= N are zero, row N is all ones so that D^T[N][m] = rowsum(A)[m] comes for free) and the A rows
// as the N operand; both operands go through shared memory in the UMMA K-major no-swizzle layout as (hi, lo) pairs.
// Split-K over CTAs (2 CTAs/SM, 64 samples per stage, next stage's global loads prefetched into registers while the
// MMAs of the current one run), one red.global.add per output element per CTA at the end.
#include "common.cuh"
#include "tc_common.cuh"

namespace nicer {

constexpr int OT_THREADS = 256;
constexpr int OT_KC = 64;                       // samples per stage
constexpr int OT_CH = OT_KC / 4;                // 16-byte chunks per row per stage
constexpr int OT_BROWS = 128, OT_AROWS = 64;
constexpr int OT_UNITS_A = (OT_AROWS / 8) * 4;  // a unit = 8 rows x 4 chunks, one float4 per lane
constexpr int OT_MAX_UNITS = (OT_AROWS / 8) * 4 + (OT_BROWS / 8) * 4;
constexpr int OT_PER_WARP = OT_MAX_UNITS / (OT_THREADS / 32);   // 12

struct OtSmem {
    float bhi[OT_CH * OT_BROWS * 4], blo[OT_CH * OT_BROWS * 4];   // 32 KB each
    // A rows as ONE 128-row operand per chunk: rows [0,64) hold the hi halves, rows [64,128) the lo halves, so that
    // B_hi x [A_hi ; A_lo] is a single N = 128 MMA (a tcgen05.mma costs the same ~102 cycles for N = 64 and N = 128,
    // scripts/mma_bench.cu): 2 MMAs per K-step instead of 3
    float acomb[OT_CH * 2 * OT_AROWS * 4];                        // 32 KB
};

// up to OT_MAX_JOBS contractions over the same sample range in one launch (the weight gradients of one network backward):
// CTAs [j * ctas_per_job, (j+1) * ctas_per_job) split the stages of job j
constexpr int OT_MAX_JOBS = 8;
struct OtJobs {
    const float *A[OT_MAX_JOBS], *B[OT_MAX_JOBS];
    float *C[OT_MAX_JOBS], *bias[OT_MAX_JOBS];
    uint32_t lda[OT_MAX_JOBS], ldb[OT_MAX_JOBS], ldc[OT_MAX_JOBS], M[OT_MAX_JOBS], N[OT_MAX_JOBS];
};

__global__ void __launch_bounds__(OT_THREADS, 2)
outer_accum_tc_kernel(const OtJobs js, uint32_t P, uint32_t stages_per_cta, uint32_t ctas_per_job) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    OtSmem &sm = *reinterpret_cast<OtSmem *>(smem_raw);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t job = blockIdx.x / ctas_per_job, cta = blockIdx.x - job * ctas_per_job;
    const float *__restrict__ A = js.A[job], *__restrict__ B = js.B[job];
    float *C = js.C[job], *bias = js.bias[job];
    const uint32_t lda = js.lda[job], ldb = js.ldb[job], ldc = js.ldc[job], M = js.M[job], N = js.N[job];
    const uint32_t n_stages = (P + OT_KC - 1) / OT_KC;
    const uint32_t s0 = cta * stages_per_cta;
    const uint32_t s1 = (s0 + stages_per_cta < n_stages) ? s0 + stages_per_cta : n_stages;
    const uint32_t n_units_b = ((N + 7) / 8) * 4;                 // units that touch real B rows
    const uint32_t n_units = OT_UNITS_A + n_units_b;

    // one-time: zero both operands (rows that are never loaded must stay 0), ones row at index N, barrier, TMEM
    for (int i = tid; i < (int)(sizeof(OtSmem) / 4); i += OT_THREADS) reinterpret_cast<float *>(&sm)[i] = 0.f;
    __syncthreads();
    if (bias && N < OT_BROWS) {
        for (int i = tid; i < OT_CH * 4; i += OT_THREADS) sm.bhi[((i >> 2) * OT_BROWS + N) * 4 + (i & 3)] = 1.0f;   // exact in tf32
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    uint32_t parity = 0;

    // unit u of this warp: u = warp + 8*i ; rows rg*8 + lane%8, chunk cg*4 + lane/8
    float4 pre[OT_PER_WARP];
    auto load_stage = [&](uint32_t st) {
        const uint32_t p0 = st * OT_KC;
#pragma unroll
        for (int i = 0; i < OT_PER_WARP; ++i) {
            const uint32_t u = warp + 8 * i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < n_units) {
                const bool isA = u < OT_UNITS_A;
                const uint32_t uu = isA ? u : u - OT_UNITS_A;
                const uint32_t r = (uu >> 2) * 8 + (lane & 7), c = (uu & 3) * 4 + (lane >> 3);
                const uint32_t p = p0 + c * 4;
                const uint32_t rows = isA ? M : N;
                if (r < rows && p < P) {
                    const float *src = isA ? A + (size_t)r * lda + p : B + (size_t)r * ldb + p;
                    v = __ldg(reinterpret_cast<const float4 *>(src));
                }
            }
            pre[i] = v;
        }
    };
    auto store_stage = [&]() {
#pragma unroll
        for (int i = 0; i < OT_PER_WARP; ++i) {
            const uint32_t u = warp + 8 * i;
            if (u < n_units) {
                const bool isA = u < OT_UNITS_A;
                const uint32_t uu = isA ? u : u - OT_UNITS_A;
                const uint32_t r = (uu >> 2) * 8 + (lane & 7), c = (uu & 3) * 4 + (lane >> 3);
                if (r >= (isA ? M : N)) continue;     // rows beyond the operand stay as initialised (zeros / the ones row)
                const float4 v = pre[i];
                float4 h, l;
                h.x = tc::tf32_hi(v.x); h.y = tc::tf32_hi(v.y); h.z = tc::tf32_hi(v.z); h.w = tc::tf32_hi(v.w);
                l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                float *dh = isA ? sm.acomb + ((size_t)c * 2 * OT_AROWS + r) * 4 : sm.bhi + ((size_t)c * OT_BROWS + r) * 4;
                float *dl = isA ? sm.acomb + ((size_t)c * 2 * OT_AROWS + OT_AROWS + r) * 4 : sm.blo + ((size_t)c * OT_BROWS + r) * 4;
                *reinterpret_cast<float4 *>(dh) = h;
                *reinterpret_cast<float4 *>(dl) = l;
            }
        }
    };

    if (s0 < s1) load_stage(s0);
    bool first = true;
    for (uint32_t st = s0; st < s1; ++st) {
        store_stage();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            constexpr uint32_t IDESC128 = tc::idesc_tf32(128, 2 * OT_AROWS), IDESC64 = tc::idesc_tf32(128, OT_AROWS);
            const uint32_t bh = tc::smem_u32(sm.bhi), bl = tc::smem_u32(sm.blo), ac = tc::smem_u32(sm.acomb);
#pragma unroll
            for (int ks = 0; ks < OT_KC / 8; ++ks) {
                const uint64_t dbh = tc::smem_desc(bh + ks * 2 * OT_BROWS * 16, OT_BROWS * 16, 128);
                const uint64_t dbl = tc::smem_desc(bl + ks * 2 * OT_BROWS * 16, OT_BROWS * 16, 128);
                // the combined operand (128 rows per chunk); its first 64 rows alone are A_hi (same chunk stride)
                const uint64_t dac = tc::smem_desc(ac + ks * 2 * (2 * OT_AROWS) * 16, 2 * OT_AROWS * 16, 128);
                // columns [0,64) += B_hi A_hi^T, columns [64,128) += B_hi A_lo^T ; then columns [0,64) += B_lo A_hi^T
                tc::mma_tf32_ss(tmem, dbh, dac, IDESC128, (first && ks == 0) ? 0u : 1u);
                tc::mma_tf32_ss(tmem, dbl, dac, IDESC64, 1u);
            }
            tc::mma_commit(&bar);
        }
        first = false;
        if (st + 1 < s1) load_stage(st + 1);      // global loads of the next stage fly while the MMAs run
        tc::mbar_wait(&bar, parity);              // shared-memory operands may be overwritten after this
        parity ^= 1u;
    }
    // epilogue: lane n of the accumulator holds D^T[n][0..63]
    __syncwarp();
    tc::fence_after_sync();
    if (s0 < s1 && warp < 4) {
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
    if (M > OT_AROWS || N >= OT_BROWS || P < 4096) return false;
    if ((lda & 3u) || (ldb & 3u) || (reinterpret_cast<uintptr_t>(A) & 15u) || (reinterpret_cast<uintptr_t>(B) & 15u)) return false;
    return true;
}

static int oa_launch(const OtJobs &js, uint32_t n_jobs, uint32_t P, cudaStream_t st) {
    const uint32_t n_stages = div_up(P, OT_KC);
    uint32_t per_job = div_up((uint32_t)(2 * num_sms()), n_jobs);
    if (per_job > n_stages) per_job = n_stages;
    const uint32_t spc = div_up(n_stages, per_job);
    per_job = div_up(n_stages, spc);
    const size_t smem = sizeof(OtSmem);
    NICER_CUDA(cudaFuncSetAttribute(outer_accum_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "nicer_outer_accum(tc)");
    outer_accum_tc_kernel<<<per_job * n_jobs, OT_THREADS, smem, st>>>(js, P, spc, per_job);
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
