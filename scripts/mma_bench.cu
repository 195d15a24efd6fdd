// Micro-benchmark: cycles per tcgen05.mma kind::tf32 instruction for the shapes the kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench scripts/mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nicer_slam_b200/csrc/tc_common.cuh"
using namespace nicer;

// mode 0: SS, same accumulator; 1: SS, 4 rotating accumulators; 2: TS (A from TMEM), same accumulator; 3: TS rotating
__global__ void __launch_bounds__(128, 1) k(int M, int N, int mode, int reps, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_tf32((uint32_t)M, (uint32_t)N);
        const uint32_t a = tc::smem_u32(smem), b = tc::smem_u32(smem + 64 * 1024);
        const uint64_t da = tc::smem_desc(a, (uint32_t)M * 16, 128), db = tc::smem_desc(b, (uint32_t)N * 16, 128);
        const int nacc = (mode & 1) ? (N <= 64 ? 4 : (N <= 128 ? 2 : 1)) : 1;
        const uint32_t dbase = tmem + 128;       // A (TS) lives in columns [0,64)
        long long t0 = clock64();
        if (mode >= 4) {
            // the weight-gradient kernel's sequence: pairs (N = 128 into columns [0,128), then N = 64 into columns [0,64))
            //   mode 4: SS, one accumulator set    5: SS, two accumulator sets alternating per pair
            //   mode 6: TS, one accumulator set    7: TS, two accumulator sets alternating per pair
            const uint32_t i128 = tc::idesc_tf32(128, 128), i64 = tc::idesc_tf32(128, 64);
            const uint64_t db128 = tc::smem_desc(b, 128 * 16, 128);
            for (int r = 0; r < reps / 2; ++r) {
                const uint32_t d = dbase + (((mode & 1) && (r & 1)) ? 128u : 0u);
                if (mode >= 6) {
                    tc::mma_tf32_ts(d, tmem + (r & 3) * 8, db128, i128, 1u);
                    tc::mma_tf32_ts(d, tmem + 32 + (r & 3) * 8, db128, i64, 1u);
                } else {
                    tc::mma_tf32_ss(d, da, db128, i128, 1u);
                    tc::mma_tf32_ss(d, da, db128, i64, 1u);
                }
            }
        } else
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = dbase + (uint32_t)(r % nacc) * (uint32_t)N;
            if (mode >= 2) tc::mma_tf32_ts(d, tmem + (r & 7) * 8, db, idesc, 1u);
            else tc::mma_tf32_ss(d, da, db, idesc, 1u);
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long *d;
    cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int reps = 2000;
    const int Ms[2] = {128, 64};
    const int Ns[4] = {64, 80, 128, 256};
    for (int mi = 0; mi < 2; ++mi)
        for (int ni = 0; ni < 4; ++ni)
            for (int mode = 0; mode < 8; ++mode) {
                if (Ms[mi] == 64 && mode >= 2) continue;
                if (mode >= 4 && ni != 2) continue;
                k<<<148, 128, 160 * 1024>>>(Ms[mi], Ns[ni], mode, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[148];
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                printf("M=%3d N=%3d mode=%d (%s, %s%s): %7.1f cycles/MMA  %s\n", Ms[mi], Ns[ni], mode, (mode & 2) ? "TS" : "SS",
                       (mode & 1) ? "rotating acc" : "one acc", mode >= 4 ? ", N=128/N=64 pairs" : "", (double)h[0] / reps,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
// Measured on B200 (sm_100a, 1965 MHz), cycles per instruction, one issuing thread, K = 8 (tf32):
//   M=128 or 64, N = 64 / 80 / 128 : 101.9 (SS or TS, same or rotating accumulators)
//   M=128, N = 256                 : 160.9 (SS), 136.1 (TS);   M=64, N=256: 148.9 (SS)
// i.e. ~102 cycles per tcgen05.mma whatever N <= 128 is: a 64-wide layer uses a third of the tensor pipe's peak, and
// the instruction count (K-steps x 3 for 3xTF32) is what bounds the 64-wide MLP kernels, not the FLOPs.
