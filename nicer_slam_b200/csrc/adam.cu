// Fused dense Adam step for the hash-grid tables: one pass over (param, grad, exp_avg, exp_avg_sq) that
// also zeroes the gradient, replacing torch.optim.Adam's ~8 elementwise kernels per tensor plus optimizer.zero_grad()
// (/root/reference/code/training/volsdf_train.py:174 `torch.optim.Adam(para_list, betas=(0.9, 0.99), eps=1e-15)`, :547, :576).
// The reference steps EVERY entry of the 133 M-row color table each iteration (momentum drifts on untouched rows), so the
// update is dense: 5 x 4 B read/written per entry.  Arithmetic follows torch's _single_tensor_adam / _multi_tensor_adam
// (amsgrad = False, weight_decay = 0, maximize = False) operation by operation so that results are bit-identical:
//     exp_avg    = lerp(exp_avg, grad, 1 - beta1)              = fma(1 - beta1, grad - exp_avg, exp_avg)
//     exp_avg_sq = exp_avg_sq * beta2 ; += (1 - beta2) * grad * grad   (mul rounded, then addcmul: fma((1-b2)*g, g, .))
//     denom      = sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps
//     param      = param + (-lr / (1 - beta1^t)) * (exp_avg / denom)
#include "common.cuh"

namespace nicer {

// How torch's CUDA kernels round (set once from a probe against torch.optim.Adam on the GPU, scripts/gpu_adam_probe.py):
//   bit 0: exp_avg_sq += (1-b2) * (g*g)  [foreach addcmul: scalar * (t1*t2)]  instead of ((1-b2)*g) * g  [single-tensor addcmul]
//   bit 1: sqrt(v) * (1/bc2_sqrt) instead of sqrt(v) / bc2_sqrt
//   bit 2: lerp without FMA
//   bit 3: param += (step_size * m) / denom instead of fma(step_size, m / denom, param)
static int g_adam_variant = 1;

template <int VAR>
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float w1, float beta2, float w2, float bc2_sqrt, float inv_bc2_sqrt,
                                         float eps, float neg_step) {
    m = (VAR & 4) ? __fadd_rn(m, __fmul_rn(w1, __fsub_rn(g, m))) : __fmaf_rn(w1, __fsub_rn(g, m), m);
    v = (VAR & 1) ? __fmaf_rn(w2, __fmul_rn(g, g), __fmul_rn(v, beta2)) : __fmaf_rn(__fmul_rn(w2, g), g, __fmul_rn(v, beta2));
    const float sq = __fsqrt_rn(v);
    const float denom = __fadd_rn((VAR & 2) ? __fmul_rn(sq, inv_bc2_sqrt) : __fdiv_rn(sq, bc2_sqrt), eps);
    p = (VAR & 8) ? __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_step, m), denom)) : __fmaf_rn(neg_step, __fdiv_rn(m, denom), p);
}

template <int VAR>
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, size_t n, float w1, float beta2,
            float w2, float bc2_sqrt, float inv_bc2_sqrt, float eps, float neg_step, int zero_grad) {
    const size_t i0 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= n) return;
    if (i0 + 4 <= n) {
        float4 P4 = *reinterpret_cast<float4 *>(p + i0), G4 = *reinterpret_cast<float4 *>(g + i0);
        float4 M4 = *reinterpret_cast<float4 *>(m + i0), V4 = *reinterpret_cast<float4 *>(v + i0);
        float *pp = &P4.x, *gg = &G4.x, *mm = &M4.x, *vv = &V4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) adam_one<VAR>(pp[k], gg[k], mm[k], vv[k], w1, beta2, w2, bc2_sqrt, inv_bc2_sqrt, eps, neg_step);
        *reinterpret_cast<float4 *>(p + i0) = P4;
        *reinterpret_cast<float4 *>(m + i0) = M4;
        *reinterpret_cast<float4 *>(v + i0) = V4;
        if (zero_grad) *reinterpret_cast<float4 *>(g + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (size_t i = i0; i < n; ++i) {
            adam_one<VAR>(p[i], g[i], m[i], v[i], w1, beta2, w2, bc2_sqrt, inv_bc2_sqrt, eps, neg_step);
            if (zero_grad) g[i] = 0.f;
        }
    }
}

}  // namespace nicer

// step: 1-based count of this update.  The scalar factors are formed on the host in double and rounded to fp32 exactly as
// torch does for its Python-scalar arguments (value=1-beta2, step_size=lr/bias_correction1, sqrt(bias_correction2)).
extern "C" int nicer_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, double lr, double beta1,
                               double beta2, double eps, uint64_t step, int zero_grad, void *stream) {
    if (n == 0) return 0;
    if (!param || !grad || !exp_avg || !exp_avg_sq) NICER_FAIL(-1, "nicer_adam_step: NULL pointer");
    if (step == 0) NICER_FAIL(-1, "nicer_adam_step: step counts from 1");
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u)
        NICER_FAIL(-1, "nicer_adam_step: pointers must be 16-byte aligned");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2), b2 = (float)beta2;
    const float bc2_sqrt = (float)sqrt(bc2), neg_step = (float)(-(lr / bc1));
    const float inv_bc2_sqrt = 1.0f / bc2_sqrt;
    const size_t threads = (n + 3) / 4;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(V) nicer::adam_kernel<V><<<grid, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, (size_t)n, w1, b2, w2, bc2_sqrt, inv_bc2_sqrt, \
                                                              (float)eps, neg_step, zero_grad)
    switch (nicer::g_adam_variant & 15) {
        case 0: LAUNCH(0); break; case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break; case 5: LAUNCH(5); break; case 6: LAUNCH(6); break; case 7: LAUNCH(7); break;
        case 8: LAUNCH(8); break; case 9: LAUNCH(9); break; case 10: LAUNCH(10); break; case 11: LAUNCH(11); break;
        case 12: LAUNCH(12); break; case 13: LAUNCH(13); break; case 14: LAUNCH(14); break; default: LAUNCH(15); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_adam_step");
    return 0;
}

/* test hook: choose how the update is rounded (see g_adam_variant above) */
extern "C" int nicer_set_adam_variant(int variant) {
    nicer::g_adam_variant = variant;
    return 0;
}
