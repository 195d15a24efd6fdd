// Probe: does tcgen05.mma kind::tf32 accept an MN-major B operand in the no-swizzle ("interleaved") layout?
// The weights of the MLP kernels sit in shared memory as [k/4][64 rows][4] (K-major core matrices).  Read MN-major
// (LBO = 128 B, SBO = 1024 B) the very same bytes would be W^T, which would let one staged copy serve both the
// forward products a W^T... and the transposed (gradient-chain / reverse-pass) products.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mn_major_probe scripts/mn_major_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../nicer_slam_b200/csrc/tc_common.cuh"
using namespace nicer;

// variant 0: K-major (reference behaviour)      D[p][n] = sum_k A[p][k] W[n][k]
// variant 1: b_major = MN, LBO = 128, SBO = 1024, K-step advance 128 B     expect D[p][n] = sum_k A[p][k] W[k][n]
// variant 2: b_major = MN, LBO = 1024, SBO = 128, K-step advance 128 B
// variant 3: b_major = MN, LBO = 128, SBO = 1024, K-step advance 2048 B
// variant 4: b_major = MN, LBO = 1024, SBO = 128, K-step advance 2048 B
__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *W, int variant, float *D) {
    __shared__ __align__(128) float w[64 * 64];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        w[((k >> 2) * 64 + n) * 4 + (k & 3)] = tc::tf32_hi(W[n * 64 + k]);
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&slot, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = slot, lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c8 = 0; c8 < 8; ++c8) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = tc::tf32_hi(A[tid * 64 + c8 * 8 + i]);
        tc::tmem_st8(lane_base + c8 * 8, v);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t base = tc::smem_u32(w);
        for (int ks = 0; ks < 8; ++ks) {
            uint64_t desc;
            uint32_t idesc;
            if (variant == 0) {
                desc = tc::smem_desc(base + ks * 2048u, 1024u, 128u);
                idesc = tc::idesc_tf32(128, 64);
            } else {
                const uint32_t lbo = (variant & 1) ? 128u : 1024u, sbo = (variant & 1) ? 1024u : 128u;
                const uint32_t adv = (variant <= 2) ? 128u : 2048u;
                desc = tc::smem_desc(base + ks * adv, lbo, sbo);
                idesc = tc::idesc_tf32(128, 64, 0u, 1u);
            }
            tc::mma_tf32_ts(tmem + 64, tmem + ks * 8, desc, idesc, ks > 0 ? 1u : 0u);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    __syncwarp();
    tc::fence_after_sync();
    for (int c8 = 0; c8 < 8; ++c8) {
        float v[8];
        tc::tmem_ld8(lane_base + 64 + c8 * 8, v);
        tc::wait_ld();
        for (int i = 0; i < 8; ++i) D[tid * 64 + c8 * 8 + i] = v[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

static float hi(float a) {
    uint32_t u;
    memcpy(&u, &a, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(&a, &u, 4);
    return a;
}

int main() {
    float hA[128 * 64], hW[64 * 64], hD[128 * 64];
    srand(1);
    for (float &v : hA) v = (float)rand() / RAND_MAX - 0.5f;
    for (float &v : hW) v = (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dW, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dW, sizeof(hW)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW, sizeof(hW), cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 5; ++variant) {
        cudaMemset(dD, 0, sizeof(hD));
        probe<<<1, 128>>>(dA, dW, variant, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double eK = 0, eT = 0, nrm = 0;
        for (int p = 0; p < 128; ++p)
            for (int n = 0; n < 64; ++n) {
                double rK = 0, rT = 0;
                for (int k = 0; k < 64; ++k) {
                    rK += (double)hi(hA[p * 64 + k]) * hi(hW[n * 64 + k]);
                    rT += (double)hi(hA[p * 64 + k]) * hi(hW[k * 64 + n]);
                }
                const double d = hD[p * 64 + n];
                eK += (d - rK) * (d - rK); eT += (d - rT) * (d - rT); nrm += rK * rK;
            }
        printf("variant %d: rel err vs A*W^T (K-major meaning) %.3e   vs A*W (transposed meaning) %.3e\n", variant,
               sqrt(eK / nrm), sqrt(eT / nrm));
    }
    return 0;
}
