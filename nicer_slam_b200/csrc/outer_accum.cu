// C[M,N] += A[M][P] * B[N][P]^T, bias[M] += rowsum(A): the weight / bias gradient contraction over the
// sample dimension (what autograd's nn.Linear backward does with a cuBLAS sgemm in the reference,
// /root/reference/code/model/base_networks.py:176-179,378-381).  Operands are the feature-major
// buffers written by the fused backward kernels, so both are K(=P)-contiguous: split-K over CTAs,
// 32-sample tiles staged transposed in shared memory, 4xRN register blocks, one red.global.add per
// output element per CTA at the end.
#include "common.cuh"

namespace nicer {

constexpr int OA_TP = 32;       // samples per tile
constexpr int OA_THREADS = 256; // 16 x 16
constexpr int OA_RM = 4;        // rows of C per thread -> M <= 64
constexpr int OA_MPAD = 68;

template <int RN>
__global__ void __launch_bounds__(OA_THREADS)
outer_accum_kernel(const float *__restrict__ A, uint32_t lda, uint32_t M, const float *__restrict__ B, uint32_t ldb,
                   uint32_t N, uint32_t P, uint32_t tiles_per_cta, float *C, uint32_t ldc, float *bias) {
    constexpr int NPAD = 16 * RN + 4;
    __shared__ __align__(16) float As[OA_TP][OA_MPAD];
    __shared__ __align__(16) float Bs[OA_TP][NPAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lane = tid & 31, wid = tid >> 5;
    float acc[OA_RM][RN];
    float bsum[OA_RM];
#pragma unroll
    for (int r = 0; r < OA_RM; ++r) {
        bsum[r] = 0.f;
#pragma unroll
        for (int c = 0; c < RN; ++c) acc[r][c] = 0.f;
    }
    const uint32_t n_tiles = (P + OA_TP - 1) / OA_TP;
    const uint32_t t0 = blockIdx.x * tiles_per_cta;
    uint32_t t1 = t0 + tiles_per_cta;
    if (t1 > n_tiles) t1 = n_tiles;
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t p = t * OA_TP + lane;
        // warp `wid` loads rows wid, wid+8, ...; lane = sample within the tile (coalesced 128 B)
        for (int m = wid; m < 16 * OA_RM; m += OA_THREADS / 32)
            As[lane][m] = (m < (int)M && p < P) ? A[(size_t)m * lda + p] : 0.f;
        for (int n = wid; n < 16 * RN; n += OA_THREADS / 32)
            Bs[lane][n] = (n < (int)N && p < P) ? B[(size_t)n * ldb + p] : 0.f;
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < OA_TP; ++i) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[i][ty * OA_RM]);
            const float a[OA_RM] = {a4.x, a4.y, a4.z, a4.w};
            float b[RN];
#pragma unroll
            for (int c = 0; c < RN; ++c) b[c] = Bs[i][tx * RN + c];
#pragma unroll
            for (int r = 0; r < OA_RM; ++r) {
                bsum[r] += a[r];
#pragma unroll
                for (int c = 0; c < RN; ++c) acc[r][c] += a[r] * b[c];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < OA_RM; ++r) {
        const int m = ty * OA_RM + r;
        if (m >= (int)M) continue;
#pragma unroll
        for (int c = 0; c < RN; ++c) {
            const int n = tx * RN + c;
            if (n < (int)N) atomicAdd(&C[(size_t)m * ldc + n], acc[r][c]);
        }
        if (bias && tx == 0) atomicAdd(&bias[m], bsum[r]);
    }
}

int launch_outer_accum_tc(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N, uint32_t P, float *C,
                          uint32_t ldc, float *bias, cudaStream_t st);

int launch_outer_accum_batch_tc(const nicer_oa_job_t *jobs, uint32_t n_jobs, uint32_t P, uint8_t *done, cudaStream_t st);

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_outer_accum(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N,
                                 uint32_t P, float *C, uint32_t ldc, float *bias, void *stream) {
    if (P == 0 || M == 0 || N == 0) return 0;
    if (!A || !B || !C) NICER_FAIL(-1, "nicer_outer_accum: NULL pointer");
    if (M > 64 || N > 144) NICER_FAIL(-1, "nicer_outer_accum: M <= 64 and N <= 144 required (got %u, %u)", M, N);
    if (lda < P || ldb < P || ldc < N) NICER_FAIL(-1, "nicer_outer_accum: bad leading dimension");
    {
        const int r = launch_outer_accum_tc(A, lda, M, B, ldb, N, P, C, ldc, bias, (cudaStream_t)stream);
        if (r != 0) return r < 0 ? r : 0;
    }
    const uint32_t n_tiles = div_up(P, OA_TP);
    uint32_t grid = (uint32_t)(2 * num_sms());
    if (grid > div_up(n_tiles, 4)) grid = div_up(n_tiles, 4);
    if (grid == 0) grid = 1;
    const uint32_t tpc = div_up(n_tiles, grid);
    grid = div_up(n_tiles, tpc);
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 64) outer_accum_kernel<4><<<grid, OA_THREADS, 0, st>>>(A, lda, M, B, ldb, N, P, tpc, C, ldc, bias);
    else if (N <= 80) outer_accum_kernel<5><<<grid, OA_THREADS, 0, st>>>(A, lda, M, B, ldb, N, P, tpc, C, ldc, bias);
    else outer_accum_kernel<9><<<grid, OA_THREADS, 0, st>>>(A, lda, M, B, ldb, N, P, tpc, C, ldc, bias);
    NICER_CHECK_LAUNCH("nicer_outer_accum");
    return 0;
}

extern "C" int nicer_outer_accum_batch(const nicer_oa_job_t *jobs, uint32_t n_jobs, uint32_t P, void *stream) {
    if (n_jobs == 0 || P == 0) return 0;
    if (!jobs) NICER_FAIL(-1, "nicer_outer_accum_batch: jobs is NULL");
    if (n_jobs > 64) NICER_FAIL(-1, "nicer_outer_accum_batch: at most 64 jobs per call (got %u)", n_jobs);
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const nicer_oa_job_t &q = jobs[j];
        if (!q.A || !q.B || !q.C) NICER_FAIL(-1, "nicer_outer_accum_batch: NULL pointer in job %u", j);
        if (q.M > 64 || q.N > 144) NICER_FAIL(-1, "nicer_outer_accum_batch: M <= 64 and N <= 144 required (job %u: %u, %u)", j, q.M, q.N);
        if (q.lda < P || q.ldb < P || q.ldc < q.N) NICER_FAIL(-1, "nicer_outer_accum_batch: bad leading dimension in job %u", j);
    }
    uint8_t done[64] = {0};
    if (int e = launch_outer_accum_batch_tc(jobs, n_jobs, P, done, (cudaStream_t)stream)) return e;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        if (done[j] || jobs[j].M == 0 || jobs[j].N == 0) continue;
        const nicer_oa_job_t &q = jobs[j];
        if (int e = nicer_outer_accum(q.A, q.lda, q.M, q.B, q.ldb, q.N, P, q.C, q.ldc, q.bias, stream)) return e;
    }
    return 0;
}
