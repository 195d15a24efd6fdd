// SLAMLoss terms over the rendered rays and the eikonal points as three kernels (+ one memset) instead of the ~300
// elementwise / reduction kernels autograd runs for the reference formulation (loss.py:113-233, MiDaS.py:6-140):
//   loss_rays_kernel      warp per ray: foreground flag from the ray's SDF sign change, RGB L1, mono-normal L1 + cosine,
//                         per-image least-squares sums of the scale-and-shift-invariant depth loss, sensor-depth L1 sums;
//                         writes the gradients that need no global denominator
//   loss_points_kernel    thread per eikonal point: eikonal + smoothness terms and their gradients
//   loss_finalize_kernel  one block: per-image (scale, shift), SSI depth loss (MSE + 0.5 * list-axis gradient term),
//                         depth gradients, the term values
// Sums are accumulated in double.  Every gradient is d(sum_i w_i term_i)/d(input); the caller scales by dL/d(that sum).
#include "common.cuh"
#include "loss_math.cuh"

namespace nicer {

enum { ACC_RGB = 0, ACC_NL1 = 1, ACC_NCOS = 2, ACC_GTD = 3, ACC_GTD_CNT = 4, ACC_EIK = 5, ACC_SMOOTH = 6, ACC_FRAMES = 8 };

constexpr int LR_WARPS = 8;

__global__ void __launch_bounds__(LR_WARPS * 32)
loss_rays_kernel(const nicer_loss_t a, double *acc, float *maskf) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * LR_WARPS + warp;
    __shared__ double red[LR_WARPS][5];
    double part[5] = {0, 0, 0, 0, 0};
    if (r < a.R) {
        // foreground: the SDF changes sign along the ray (loss.py:165-167)
        bool pos = false, neg = false;
        for (uint32_t s = lane; s < a.S; s += 32) {
            const float v = a.sdf[(size_t)r * a.S + s];
            pos |= v > 0.f;
            neg |= v < 0.f;
        }
        const bool fg = __any_sync(0xffffffffu, pos) && __any_sync(0xffffffffu, neg);
        if (lane == 0) {
            const float m = (a.mask_gt && a.mask_gt[r] > 0.5f && fg) ? 1.0f : 0.f;
            maskf[r] = m;
            if (a.rgb_pred) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float d = a.rgb_pred[3 * (size_t)r + c] - a.rgb_gt[3 * (size_t)r + c];
                    s += fabsf(d);
                    if (a.g_rgb) a.g_rgb[3 * (size_t)r + c] = a.w_rgb * sgnf(d) / (3.0f * (float)a.R);
                }
                part[0] = s;
            }
            if (a.normal_pred) {
                float l1, cs, g[3];
                normal_terms(a.normal_pred + 3 * (size_t)r, a.normal_gt + 3 * (size_t)r, m, a.w_normal_l1 / (float)a.R,
                             a.w_normal_cos / (float)a.R, &l1, &cs, g);
                part[1] = l1; part[2] = cs;
                if (a.g_normal) { a.g_normal[3 * (size_t)r] = g[0]; a.g_normal[3 * (size_t)r + 1] = g[1]; a.g_normal[3 * (size_t)r + 2] = g[2]; }
            }
            if (a.depth_pred && a.depth_gt) {
                const float md = a.depth_mask_all ? 1.0f : m;
                if (md != 0.f) {
                    const double p = a.depth_pred[r], t = a.depth_gt[r] * 50.0f + 0.5f;
                    double *f = acc + ACC_FRAMES + 5 * (r / a.N);
                    atomicAdd(f + 0, p * p); atomicAdd(f + 1, p); atomicAdd(f + 2, 1.0); atomicAdd(f + 3, p * t); atomicAdd(f + 4, t);
                }
            }
            if (a.depth_pred && a.gt_depth && a.gt_depth_valid[r] > 0.f) {
                part[3] = fabsf(a.depth_pred[r] - a.gt_depth[r]);
                part[4] = 1.0;
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) red[warp][k] = part[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int w = 0; w < LR_WARPS; ++w) s += red[w][threadIdx.x];
        if (s != 0) atomicAdd(acc + threadIdx.x, s);          // ACC_RGB .. ACC_GTD_CNT are 0..4
    }
}

constexpr int LP_BLOCK = 256;
__global__ void __launch_bounds__(LP_BLOCK)
loss_points_kernel(const nicer_loss_t a, double *acc) {
    const uint32_t i = blockIdx.x * LP_BLOCK + threadIdx.x;
    float eik = 0.f, sm = 0.f;
    if (i < a.G) {
        const float g1[3] = {a.grad_theta[3 * (size_t)i], a.grad_theta[3 * (size_t)i + 1], a.grad_theta[3 * (size_t)i + 2]};
        float ge[3] = {0.f, 0.f, 0.f}, gs1[3] = {0.f, 0.f, 0.f}, gs2[3] = {0.f, 0.f, 0.f};
        if (a.w_eik > 0.f) eik = eikonal_term(g1, ge);
        if (a.grad_theta_nei && a.w_smooth > 0.f) {
            const float g2[3] = {a.grad_theta_nei[3 * (size_t)i], a.grad_theta_nei[3 * (size_t)i + 1], a.grad_theta_nei[3 * (size_t)i + 2]};
            sm = smooth_term(g1, g2, gs1, gs2);
        }
        const float ke = a.w_eik / (float)a.G, ks = a.w_smooth / (float)a.G;
        if (a.g_theta) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a.g_theta[3 * (size_t)i + c] = ke * ge[c] + ks * gs1[c];
        }
        if (a.g_theta_nei) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a.g_theta_nei[3 * (size_t)i + c] = ks * gs2[c];
        }
    }
    __shared__ double red[LP_BLOCK / 32][2];
    double e = eik, s = sm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = e; red[threadIdx.x >> 5][1] = s; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0;
        for (int w = 0; w < LP_BLOCK / 32; ++w) t += red[w][threadIdx.x];
        atomicAdd(acc + ACC_EIK + threadIdx.x, t);
    }
}

constexpr int LF_BLOCK = 1024;
__global__ void __launch_bounds__(LF_BLOCK)
loss_finalize_kernel(const nicer_loss_t a, const double *acc, const float *maskf, float *terms) {
    __shared__ float scale[NICER_LOSS_MAX_FRAMES], shift[NICER_LOSS_MAX_FRAMES];
    __shared__ double red[LF_BLOCK / 32][2];
    __shared__ float m_total_s;
    const bool depth_on = a.depth_pred && a.depth_gt;
    if (threadIdx.x < a.B && depth_on) {
        const double *f = acc + ACC_FRAMES + 5 * threadIdx.x;
        scale_shift((float)f[0], (float)f[1], (float)f[2], (float)f[3], (float)f[4], &scale[threadIdx.x], &shift[threadIdx.x]);
    }
    if (threadIdx.x == 0) {
        double mt = 0;
        for (uint32_t b = 0; b < a.B; ++b) mt += acc[ACC_FRAMES + 5 * b + 2];
        m_total_s = (float)mt;
    }
    __syncthreads();
    const float m_total = m_total_s;
    const float gtd_cnt = (float)acc[ACC_GTD_CNT];
    double mse = 0, reg = 0;
    if (a.depth_pred) {
        for (uint32_t r = threadIdx.x; r < a.R; r += LF_BLOCK) {
            float g = 0.f;
            if (depth_on && m_total > 0.f) {
                const uint32_t b = r / a.N, n = r - b * a.N;
                const float s = scale[b], sh = shift[b];
                auto msk = [&](uint32_t q) { return a.depth_mask_all ? 1.0f : maskf[q]; };
                auto dif = [&](uint32_t q) { return msk(q) * ((s * a.depth_pred[q] + sh) - (a.depth_gt[q] * 50.0f + 0.5f)); };
                const float m = msk(r);
                const float res = (s * a.depth_pred[r] + sh) - (a.depth_gt[r] * 50.0f + 0.5f);
                mse += (double)(m * res * res);
                g = m * res * s / m_total;
                const float d0 = m * res;
                float sg = 0.f;
                if (n > 0) sg += msk(r - 1) * m * sgnf(d0 - dif(r - 1));
                if (n + 1 < a.N) {
                    const float mm = m * msk(r + 1), dd = dif(r + 1) - d0;
                    reg += (double)(mm * fabsf(dd));
                    sg -= mm * sgnf(dd);
                }
                g += 0.5f * s * m * sg / m_total;
                g *= a.w_depth;
            }
            if (a.gt_depth && gtd_cnt > 0.f && a.gt_depth_valid[r] > 0.f) g += a.w_gt_depth * sgnf(a.depth_pred[r] - a.gt_depth[r]) / gtd_cnt;
            if (a.g_depth) a.g_depth[r] = g;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mse += __shfl_xor_sync(0xffffffffu, mse, o); reg += __shfl_xor_sync(0xffffffffu, reg, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = mse; red[threadIdx.x >> 5][1] = reg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ms = 0, rg = 0;
        for (int w = 0; w < LF_BLOCK / 32; ++w) { ms += red[w][0]; rg += red[w][1]; }
        float depth = 0.f;
        if (depth_on && m_total > 0.f) depth = (float)(ms / (2.0 * (double)m_total)) + 0.5f * (float)(rg / (double)m_total);
        terms[NICER_LOSS_RGB] = a.rgb_pred ? (float)(acc[ACC_RGB] / (3.0 * (double)a.R)) : 0.f;
        terms[NICER_LOSS_DEPTH] = depth;
        // mean over an empty selection is NaN in the reference (0/0); keep that
        terms[NICER_LOSS_GT_DEPTH] = a.gt_depth ? (float)(acc[ACC_GTD] / acc[ACC_GTD_CNT]) : 0.f;
        terms[NICER_LOSS_NORMAL_L1] = a.normal_pred ? (float)(acc[ACC_NL1] / (double)a.R) : 0.f;
        terms[NICER_LOSS_NORMAL_COS] = a.normal_pred ? (float)(acc[ACC_NCOS] / (double)a.R) : 0.f;
        terms[NICER_LOSS_EIKONAL] = (a.grad_theta && a.w_eik > 0.f) ? (float)(acc[ACC_EIK] / (double)a.G) : 0.f;
        terms[NICER_LOSS_SMOOTH] = (a.grad_theta_nei && a.w_smooth > 0.f) ? (float)(acc[ACC_SMOOTH] / (double)a.G) : 0.f;
        terms[NICER_LOSS_SUM] = a.w_rgb * terms[NICER_LOSS_RGB] + a.w_depth * terms[NICER_LOSS_DEPTH] +
                                (a.gt_depth ? a.w_gt_depth * terms[NICER_LOSS_GT_DEPTH] : 0.f) +
                                a.w_normal_l1 * terms[NICER_LOSS_NORMAL_L1] + a.w_normal_cos * terms[NICER_LOSS_NORMAL_COS] +
                                a.w_eik * terms[NICER_LOSS_EIKONAL] + a.w_smooth * terms[NICER_LOSS_SMOOTH];
    }
}


// ---- masked L1 mean: mean |a - b| over the entries whose mask is set (the photometric-warp and optical-flow terms,
// model/loss.py:93-104,145-152 of the reference: torch.abs(x[mask] - y[mask]).mean()).  a, b: [n_mask * inner], mask: one byte
// per `inner` consecutive values; b may be shorter than a (b_len values, repeated: the warp target is the same for every
// target frame).  One block: the tensors are a few 100 k values, the reduction order is fixed (deterministic).
constexpr int ML_BLOCK = 256, ML_MAX_BLOCKS = 128;
// ws: [ML_MAX_BLOCKS] double partial sums, [ML_MAX_BLOCKS] uint32 partial counts, one uint32 ticket (zeroed by the launcher).
// The last block to finish adds the partials in block order: the result does not depend on the block schedule.
__global__ void __launch_bounds__(ML_BLOCK)
masked_l1_mean_kernel(const float *__restrict__ a, const float *__restrict__ b, const unsigned char *__restrict__ mask,
                      uint32_t n_mask, uint32_t inner, uint32_t b_len, double *ws, float *out) {
    double sum = 0.0;
    uint32_t cnt = 0;
    for (uint32_t m = blockIdx.x * ML_BLOCK + threadIdx.x; m < n_mask; m += gridDim.x * ML_BLOCK) {
        if (!mask[m]) continue;             // masked-out entries are never read into the sum (NaN / Inf there are dropped)
        ++cnt;
        for (uint32_t c = 0; c < inner; ++c) {
            const size_t i = (size_t)m * inner + c;
            sum += (double)fabsf(a[i] - b[i % b_len]);
        }
    }
    __shared__ double ssum[ML_BLOCK / 32];
    __shared__ uint32_t scnt[ML_BLOCK / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; scnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    uint32_t *pcnt = reinterpret_cast<uint32_t *>(ws + ML_MAX_BLOCKS), *ticket = pcnt + ML_MAX_BLOCKS;
    if (threadIdx.x == 0) {
        double t = 0.0;
        uint32_t n = 0;
        for (int w = 0; w < ML_BLOCK / 32; ++w) { t += ssum[w]; n += scnt[w]; }
        ws[blockIdx.x] = t;
        pcnt[blockIdx.x] = n;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        uint32_t n = 0;
        for (uint32_t k = 0; k < gridDim.x; ++k) { t += ws[k]; n += pcnt[k]; }
        const float count = (float)n * (float)inner;
        out[0] = (float)t / count;          // 0 / 0 = NaN for an empty selection, like the mean of an empty tensor
        out[1] = count;
    }
}
// ga = g * sign(a - b) / count on the selected entries, 0 elsewhere
__global__ void masked_l1_mean_backward_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                               const unsigned char *__restrict__ mask, uint32_t n_mask, uint32_t inner,
                                               uint32_t b_len, const float *__restrict__ out, const float *__restrict__ g,
                                               float *ga) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_mask * inner) return;
    float r = 0.f;
    if (mask[i / inner]) {
        const float d = a[i] - b[i % b_len];
        r = g[0] * sgnf(d) / out[1];
    }
    ga[i] = r;
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_slam_loss(const nicer_loss_t *args, double *acc, float *maskf, float *terms, void *stream) {
    if (!args || !acc || !maskf || !terms) NICER_FAIL(-1, "nicer_slam_loss: NULL pointer");
    const nicer_loss_t &a = *args;
    if (a.R == 0 || a.B == 0 || a.N == 0 || a.R != a.B * a.N) NICER_FAIL(-1, "nicer_slam_loss: R must equal B * N (got %u, %u, %u)", a.R, a.B, a.N);
    if (a.B > NICER_LOSS_MAX_FRAMES) NICER_FAIL(-1, "nicer_slam_loss: at most %d frames (got %u)", NICER_LOSS_MAX_FRAMES, a.B);
    if (!a.sdf || a.S == 0) NICER_FAIL(-1, "nicer_slam_loss: sdf is required (foreground mask)");
    if ((a.rgb_pred && !a.rgb_gt) || (a.normal_pred && !a.normal_gt) || (a.gt_depth && !a.gt_depth_valid))
        NICER_FAIL(-1, "nicer_slam_loss: a prediction is given without its target");
    cudaStream_t st = (cudaStream_t)stream;
    NICER_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * (ACC_FRAMES + 5 * (size_t)a.B), st), "nicer_slam_loss");
    loss_rays_kernel<<<div_up(a.R, LR_WARPS), LR_WARPS * 32, 0, st>>>(a, acc, maskf);
    if (a.grad_theta && a.G > 0) loss_points_kernel<<<div_up(a.G, LP_BLOCK), LP_BLOCK, 0, st>>>(a, acc);
    loss_finalize_kernel<<<1, LF_BLOCK, 0, st>>>(a, acc, maskf, terms);
    NICER_CHECK_LAUNCH("nicer_slam_loss");
    return 0;
}

extern "C" size_t nicer_masked_l1_mean_workspace(void) { return ML_MAX_BLOCKS * (sizeof(double) + sizeof(uint32_t)) + 16; }

extern "C" int nicer_masked_l1_mean(const float *a, const float *b, const unsigned char *mask, uint32_t n_mask, uint32_t inner,
                                    uint32_t b_len, void *workspace, float *out, void *stream) {
    if (!a || !b || !mask || !out || !workspace) NICER_FAIL(-1, "nicer_masked_l1_mean: NULL pointer");
    if (inner == 0 || b_len == 0) NICER_FAIL(-1, "nicer_masked_l1_mean: inner and b_len must be > 0");
    cudaStream_t st = (cudaStream_t)stream;
    double *ws = static_cast<double *>(workspace);
    NICER_CUDA(cudaMemsetAsync(reinterpret_cast<uint32_t *>(ws + ML_MAX_BLOCKS) + ML_MAX_BLOCKS, 0, sizeof(uint32_t), st), "nicer_masked_l1_mean");
    uint32_t blocks = div_up(n_mask, ML_BLOCK * 4);
    blocks = blocks < 1 ? 1 : (blocks > ML_MAX_BLOCKS ? ML_MAX_BLOCKS : blocks);
    masked_l1_mean_kernel<<<blocks, ML_BLOCK, 0, st>>>(a, b, mask, n_mask, inner, b_len, ws, out);
    NICER_CHECK_LAUNCH("nicer_masked_l1_mean");
    return 0;
}

extern "C" int nicer_masked_l1_mean_backward(const float *a, const float *b, const unsigned char *mask, uint32_t n_mask,
                                             uint32_t inner, uint32_t b_len, const float *out, const float *g, float *ga,
                                             void *stream) {
    if (n_mask == 0) return 0;
    if (!a || !b || !mask || !out || !g || !ga) NICER_FAIL(-1, "nicer_masked_l1_mean_backward: NULL pointer");
    if (inner == 0 || b_len == 0) NICER_FAIL(-1, "nicer_masked_l1_mean_backward: inner and b_len must be > 0");
    const size_t n = (size_t)n_mask * inner;
    masked_l1_mean_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, mask, n_mask, inner, b_len, out, g, ga);
    NICER_CHECK_LAUNCH("nicer_masked_l1_mean_backward");
    return 0;
}
