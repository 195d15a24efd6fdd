// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the tensor-core kernels.
// One thread = one point = one TMEM lane: activations are written to TMEM with tcgen05.st and consumed as the
// A operand of tcgen05.mma (kind::tf32) directly from TMEM; weights sit in shared memory in the UMMA K-major,
// no-swizzle ("interleaved") core-matrix layout; the accumulator row of a point is read back by the same thread
// with tcgen05.ld.32x32b.  fp32 fidelity: 3xTF32 (a = hi + lo, a*b ~= hi*bhi + lo*bhi + hi*blo, fp32 accumulate).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nicer {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp, all lanes converged)
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
// tcgen05.commit: the mbarrier is arrived on when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors
// Instruction descriptor, kind::tf32, D = fp32, A and B K-major, dense (cute::UMMA::InstrDescriptor bit layout).
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
    return (1u << 4)                 // c_format  = F32
           | (2u << 7)               // a_format  = TF32
           | (2u << 10)              // b_format  = TF32
           | (a_mn_major << 15) | (b_mn_major << 16)
           | ((N >> 3) << 17)        // n_dim
           | ((M >> 4) << 24);       // m_dim
}

// Shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start address, leading-dimension byte
// offset and stride-dimension byte offset in 16-byte units, version = 1 (Blackwell).
// K-major operand [rows][K]: core matrix = 8 rows x 16 B stored as 128 contiguous bytes;
//   LBO = distance between the two 16-byte K-chunks of one K=32B step, SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr_bytes >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;   // version
    return d;
}

// D[tmem] (+)= A[tmem] * B[smem]^T   (A from TMEM: "TS" form), single thread issues
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T  ("SS" form)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- TMEM <-> registers, 32 lanes x 32 bit, 8 consecutive columns per call (lane = thread within its warp quarter)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float v[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// 3xTF32 split: hi keeps the top 19 bits (sign, exponent, 10 mantissa bits), lo = a - hi exactly
// (round-to-nearest tf32: |lo| <= 2^-12 |a|, so the hardware's truncation of lo to tf32 costs <= 2^-23 |a|).
// Rounding is done with two integer ALU ops (add half an ulp of tf32 to the magnitude, clear the 13 dropped bits =
// round-to-nearest, ties away from zero, what cvt.rna.tf32.f32 does): cvt.rna runs on the XU pipe at ~4 results per
// clock per SM (ncu: sm__inst_executed_pipe_xu at 75% in the weight-gradient kernel, which was bound by it).
__device__ __forceinline__ float tf32_hi(float a) {
    return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}

// write 8 values as (hi, lo) into the two A-operand column ranges
__device__ __forceinline__ void tmem_st8_split(uint32_t taddr_hi, uint32_t taddr_lo, const float v[8]) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = tf32_hi(v[i]); l[i] = v[i] - h[i]; }
    tmem_st8(taddr_hi, h);
    tmem_st8(taddr_lo, l);
}

}  // namespace tc
}  // namespace nicer
