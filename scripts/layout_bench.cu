// Micro-benchmark: HBM throughput of the MLP kernels' access pattern -- every warp reads (and writes) one 128-byte line from
// each of 64 rows per step -- for the feature-major layout [rows][P] (row stride P floats, what the kernels use) against a
// tile-major layout [P/128][rows][128] (a tile's rows are one contiguous block).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/layout_bench scripts/layout_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// mode 0: read only; 1: read + write another buffer
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float *__restrict__ src, float *dst, uint32_t P, uint32_t blocks, size_t row_stride,
                                            size_t tile_stride, float *sink) {
    const uint32_t tiles = P / 128;
    float acc = 0.f;
    for (uint32_t tt = blockIdx.x * 2 + (threadIdx.x >> 7); tt < tiles; tt += gridDim.x * 2) {
        for (uint32_t b = 0; b < blocks; ++b) {        // `blocks` groups of 64 rows (like the layers of one kernel)
            const size_t base = (size_t)tt * tile_stride + (size_t)b * 64 * row_stride + (threadIdx.x & 127);
            float v[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __ldg(src + base + j * row_stride);
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                acc += v[j];
                if (MODE == 1) dst[base + j * row_stride] = v[j] * 1.0001f;
            }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const uint32_t P = 401408, blocks = 6;      // 384 rows: 616 MB per buffer
    const size_t rows = 64 * blocks, n = (size_t)rows * P;
    float *a, *b, *sink;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int layout = 0; layout < 2; ++layout) {
        const size_t row_stride = layout == 0 ? P : 128, tile_stride = layout == 0 ? 128 : rows * 128;
        for (int mode = 0; mode < 2; ++mode) {
            for (int threads_cfg = 0; threads_cfg < 2; ++threads_cfg) {
                const int grid = threads_cfg == 0 ? 148 : 296;      // 1 or 2 CTAs per SM
                float best = 1e9f;
                for (int rep = 0; rep < 5; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) k<0><<<grid, 256>>>(a, b, P, blocks, row_stride, tile_stride, sink);
                    else k<1><<<grid, 256>>>(a, b, P, blocks, row_stride, tile_stride, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                const double bytes = (double)n * 4 * (mode == 0 ? 1 : 2);
                printf("%s  %s  grid %d: %.3f ms  %.2f TB/s  (%s)\n", layout == 0 ? "feature-major [rows][P]      " : "tile-major [P/128][rows][128]",
                       mode == 0 ? "read      " : "read+write", grid, best, bytes / best / 1e9, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
