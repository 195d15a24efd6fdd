// Drop-in replacement of the reference's native op (hashencoder/src/hashencoder.cu, K1-K5) with the
// reference's own memory layouts, so hashencoder/hashgrid.py's autograd wiring can call it unchanged:
//   outputs / grad / grad_grad [L,B,C];  dy_dx [B,L,D,C];  accumulators arrive zeroed.
// Differences in *how* (not what): one thread handles a whole (point, level) with 128-bit table
// loads (C=4,8) / 64-bit (C=2), vector red.global.add for the scatters (one per corner instead of one
// per channel pair), first- and second-order work fused per call, launch on the caller's stream.
#include "common.cuh"
#include "nicer_math.cuh"

namespace nicer {

template <uint32_t D>
__device__ __forceinline__ uint32_t vertex_index(const LevelInfo &li, const uint32_t p[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= li.hashmap_size) { index += p[d] * stride; stride *= li.resolution; }
    }
    if (stride > li.hashmap_size) {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) index ^= p[d] * primes[d];
    }
    return index % li.hashmap_size;
}

template <uint32_t D>
struct Cell {
    uint32_t pg[D];
    float w[D], dw[D];
    bool inside;
};

template <uint32_t D>
__device__ __forceinline__ Cell<D> locate(const LevelInfo &li, const float *x) {
    Cell<D> c;
    c.inside = true;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        float v = x[d];
        if (v < 0.f || v > 1.f) c.inside = false;
        float p = __fmul_rn(v, li.scale);   // not contracted with the subtraction below
        c.pg[d] = (uint32_t)floorf(p);
        p -= (float)c.pg[d];
        c.dw[d] = sstep_d(p);
        c.w[d] = sstep(p);
    }
    return c;
}

template <uint32_t C>
__device__ __forceinline__ void vec_store(float *dst, const float v[C]) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (uint32_t i = 0; i < C; i += 4) *reinterpret_cast<float4 *>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (C == 2) {
        *reinterpret_cast<float2 *>(dst) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (uint32_t i = 0; i < C; ++i) dst[i] = v[i];
    }
}

template <uint32_t C>
__device__ __forceinline__ void vec_load(const float *src, float v[C]) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (uint32_t i = 0; i < C; i += 4) {
            float4 a = *reinterpret_cast<const float4 *>(src + i);
            v[i] = a.x; v[i + 1] = a.y; v[i + 2] = a.z; v[i + 3] = a.w;
        }
    } else if constexpr (C == 2) {
        float2 a = *reinterpret_cast<const float2 *>(src);
        v[0] = a.x; v[1] = a.y;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < C; ++i) v[i] = src[i];
    }
}

template <uint32_t C>
__device__ __forceinline__ void vec_red(float *dst, const float v[C]) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (uint32_t i = 0; i < C; i += 4) atomicAdd(reinterpret_cast<float4 *>(dst + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    } else if constexpr (C == 2) {
        atomicAdd(reinterpret_cast<float2 *>(dst), make_float2(v[0], v[1]));
    } else {
#pragma unroll
        for (uint32_t i = 0; i < C; ++i) atomicAdd(dst + i, v[i]);
    }
}

// ---- K1: forward (+ dy_dx)
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
hash_forward_kernel(const float *__restrict__ inputs, const float *__restrict__ grid, const int *__restrict__ offsets,
                    float *__restrict__ outputs, uint32_t B, uint32_t L, const LevelScales ls, bool want_dx,
                    float *__restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelInfo li = make_level(offsets, level, level_scale(ls, (uint32_t)level));
    const float *tab = grid + (size_t)li.offset * C;
    const Cell<D> c = locate<D>(li, inputs + (size_t)b * D);
    float *out = outputs + ((size_t)level * B + b) * C;
    float *dd = want_dx ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;
    float res[C];
#pragma unroll
    for (uint32_t ch = 0; ch < C; ++ch) res[ch] = 0.f;
    if (!c.inside) {
        vec_store<C>(out, res);
        if (dd) {
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) vec_store<C>(dd + d * C, res);
        }
        return;
    }
    float val[1u << D][C];
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); ++k) {
        uint32_t pl[D];
        float w = 1.f;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            const bool up = (k >> d) & 1u;
            pl[d] = c.pg[d] + (up ? 1u : 0u);
            w *= up ? c.w[d] : 1.f - c.w[d];
        }
        vec_load<C>(tab + (size_t)vertex_index<D>(li, pl) * C, val[k]);
#pragma unroll
        for (uint32_t ch = 0; ch < C; ++ch) res[ch] += w * val[k][ch];
    }
    vec_store<C>(out, res);
    if (dd) {
#pragma unroll
        for (uint32_t gd = 0; gd < D; ++gd) {
            float rg[C];
#pragma unroll
            for (uint32_t ch = 0; ch < C; ++ch) rg[ch] = 0.f;
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); ++k) {
                if ((k >> gd) & 1u) continue;
                float w = li.scale;
#pragma unroll
                for (uint32_t d = 0; d < D; ++d)
                    if (d != gd) w *= ((k >> d) & 1u) ? c.w[d] : 1.f - c.w[d];
                const uint32_t kr = k | (1u << gd);
#pragma unroll
                for (uint32_t ch = 0; ch < C; ++ch) rg[ch] += w * (val[kr][ch] - val[k][ch]) * c.dw[gd];
            }
            vec_store<C>(dd + gd * C, rg);
        }
    }
}

// ---- K2 (+K3): scatter into the grid and, per point, the input gradient.
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
hash_backward_kernel(const float *__restrict__ grad, const float *__restrict__ inputs, const int *__restrict__ offsets,
                     float *__restrict__ grad_grid, uint32_t B, uint32_t L, const LevelScales ls) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelInfo li = make_level(offsets, level, level_scale(ls, (uint32_t)level));
    const Cell<D> c = locate<D>(li, inputs + (size_t)b * D);
    if (!c.inside) return;
    float g[C];
    vec_load<C>(grad + ((size_t)level * B + b) * C, g);
    float *gt = grad_grid + (size_t)li.offset * C;
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); ++k) {
        uint32_t pl[D];
        float w = 1.f;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            const bool up = (k >> d) & 1u;
            pl[d] = c.pg[d] + (up ? 1u : 0u);
            w *= up ? c.w[d] : 1.f - c.w[d];
        }
        float v[C];
#pragma unroll
        for (uint32_t ch = 0; ch < C; ++ch) v[ch] = w * g[ch];
        vec_red<C>(gt + (size_t)vertex_index<D>(li, pl) * C, v);
    }
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
hash_input_backward_kernel(const float *__restrict__ grad, const float *__restrict__ dy_dx,
                           float *__restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float r[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) r[d] = 0.f;
    const float *dd = dy_dx + (size_t)b * L * D * C;
    for (uint32_t l = 0; l < L; ++l) {
        float g[C];
        vec_load<C>(grad + ((size_t)l * B + b) * C, g);
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            float v[C];
            vec_load<C>(dd + ((size_t)l * D + d) * C, v);
#pragma unroll
            for (uint32_t ch = 0; ch < C; ++ch) r[d] += g[ch] * v[ch];
        }
    }
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) grad_inputs[(size_t)b * D + d] = r[d];
}

// ---- K4 + K5 fused: grad_grad = ggx . dy_dx ; second-order scatter.
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
hash_second_backward_kernel(const float *__restrict__ grad, const float *__restrict__ inputs,
                            const int *__restrict__ offsets, const float *__restrict__ dy_dx,
                            const float *__restrict__ ggx, float *__restrict__ grad_grad,
                            float *__restrict__ grad2_grid, uint32_t B, uint32_t L, const LevelScales ls) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float gx[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) gx[d] = ggx[(size_t)b * D + d];
    {   // K4 (no range test of its own: dy_dx is zero for out-of-range points, hashencoder.cu:405-458)
        const float *dd = dy_dx + ((size_t)b * L + level) * D * C;
        float r[C];
#pragma unroll
        for (uint32_t ch = 0; ch < C; ++ch) r[ch] = 0.f;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            float v[C];
            vec_load<C>(dd + d * C, v);
#pragma unroll
            for (uint32_t ch = 0; ch < C; ++ch) r[ch] += gx[d] * v[ch];
        }
        vec_store<C>(grad_grad + ((size_t)level * B + b) * C, r);
    }
    const LevelInfo li = make_level(offsets, level, level_scale(ls, (uint32_t)level));
    const Cell<D> c = locate<D>(li, inputs + (size_t)b * D);
    if (!c.inside) return;
    float g[C];
    vec_load<C>(grad + ((size_t)level * B + b) * C, g);
    float *gt = grad2_grid + (size_t)li.offset * C;
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); ++k) {
        uint32_t pl[D];
        float coef = 0.f;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) pl[d] = c.pg[d] + ((k >> d) & 1u);
#pragma unroll
        for (uint32_t gd = 0; gd < D; ++gd) {
            float w = li.scale;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d)
                if (d != gd) w *= ((k >> d) & 1u) ? c.w[d] : 1.f - c.w[d];
            w *= gx[gd] * c.dw[gd];
            coef += ((k >> gd) & 1u) ? w : -w;
        }
        float v[C];
#pragma unroll
        for (uint32_t ch = 0; ch < C; ++ch) v[ch] = coef * g[ch];
        vec_red<C>(gt + (size_t)vertex_index<D>(li, pl) * C, v);
    }
}

template <uint32_t D>
static int launch_forward(uint32_t C, dim3 g, cudaStream_t st, const float *in, const float *emb, const int *off,
                          float *out, uint32_t B, uint32_t L, float S, uint32_t H, bool dx, float *dy_dx) {
    const LevelScales ls = host_level_scales(L, S, H);
    switch (C) {
        case 1: hash_forward_kernel<D, 1><<<g, 256, 0, st>>>(in, emb, off, out, B, L, ls, dx, dy_dx); break;
        case 2: hash_forward_kernel<D, 2><<<g, 256, 0, st>>>(in, emb, off, out, B, L, ls, dx, dy_dx); break;
        case 4: hash_forward_kernel<D, 4><<<g, 256, 0, st>>>(in, emb, off, out, B, L, ls, dx, dy_dx); break;
        case 8: hash_forward_kernel<D, 8><<<g, 256, 0, st>>>(in, emb, off, out, B, L, ls, dx, dy_dx); break;
        default: NICER_FAIL(-1, "GridEncoding: C must be 1, 2, 4, or 8.");
    }
    return 0;
}

template <uint32_t D>
static int launch_backward(uint32_t C, dim3 g, cudaStream_t st, const float *grad, const float *in, const int *off,
                           float *gg, uint32_t B, uint32_t L, float S, uint32_t H, bool dx, const float *dy_dx,
                           float *gin) {
    const uint32_t gb = div_up(B, 256);
    const LevelScales ls = host_level_scales(L, S, H);
    switch (C) {
#define CASE(CC)                                                                                       \
    case CC:                                                                                           \
        hash_backward_kernel<D, CC><<<g, 256, 0, st>>>(grad, in, off, gg, B, L, ls);                 \
        if (dx) hash_input_backward_kernel<D, CC><<<gb, 256, 0, st>>>(grad, dy_dx, gin, B, L);         \
        break;
        CASE(1) CASE(2) CASE(4) CASE(8)
#undef CASE
        default: NICER_FAIL(-1, "GridEncoding: C must be 1, 2, 4, or 8.");
    }
    return 0;
}

template <uint32_t D>
static int launch_second(uint32_t C, dim3 g, cudaStream_t st, const float *grad, const float *in, const int *off,
                         const float *dy_dx, const float *ggx, float *gg, float *g2, uint32_t B, uint32_t L, float S,
                         uint32_t H) {
    const LevelScales ls = host_level_scales(L, S, H);
    switch (C) {
        case 2: hash_second_backward_kernel<D, 2><<<g, 256, 0, st>>>(grad, in, off, dy_dx, ggx, gg, g2, B, L, ls); break;
        case 4: hash_second_backward_kernel<D, 4><<<g, 256, 0, st>>>(grad, in, off, dy_dx, ggx, gg, g2, B, L, ls); break;
        case 8: hash_second_backward_kernel<D, 8><<<g, 256, 0, st>>>(grad, in, off, dy_dx, ggx, gg, g2, B, L, ls); break;
        default: NICER_FAIL(-1, "GridEncoding: C must be 1, 2, 4, or 8.");  // C=1 unsupported as in hashencoder.cu:708-714
    }
    return 0;
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_hash_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                                         float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                         uint32_t H, int calc_grad_inputs, float *dy_dx, void *stream) {
    if (B == 0) return 0;
    if (!inputs || !embeddings || !offsets || !outputs || (calc_grad_inputs && !dy_dx))
        NICER_FAIL(-1, "nicer_hash_encode_forward: NULL pointer");
    dim3 g(div_up(B, 256), L, 1);
    int e;
    if (D == 3) e = launch_forward<3>(C, g, (cudaStream_t)stream, inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs != 0, dy_dx);
    else if (D == 2) e = launch_forward<2>(C, g, (cudaStream_t)stream, inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs != 0, dy_dx);
    else NICER_FAIL(-1, "GridEncoding: D must be 2 or 3.");
    if (e) return e;
    NICER_CHECK_LAUNCH("nicer_hash_encode_forward");
    return 0;
}

extern "C" int nicer_hash_encode_backward(const float *grad, const float *inputs, const float *embeddings,
                                          const int32_t *offsets, float *grad_embeddings, uint32_t B, uint32_t D,
                                          uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                          const float *dy_dx, float *grad_inputs, void *stream) {
    (void)embeddings;
    if (B == 0) return 0;
    if (!grad || !inputs || !offsets || !grad_embeddings || (calc_grad_inputs && (!dy_dx || !grad_inputs)))
        NICER_FAIL(-1, "nicer_hash_encode_backward: NULL pointer");
    dim3 g(div_up(B, 256), L, 1);
    int e;
    if (D == 3) e = launch_backward<3>(C, g, (cudaStream_t)stream, grad, inputs, offsets, grad_embeddings, B, L, S, H, calc_grad_inputs != 0, dy_dx, grad_inputs);
    else if (D == 2) e = launch_backward<2>(C, g, (cudaStream_t)stream, grad, inputs, offsets, grad_embeddings, B, L, S, H, calc_grad_inputs != 0, dy_dx, grad_inputs);
    else NICER_FAIL(-1, "GridEncoding: D must be 2 or 3.");
    if (e) return e;
    NICER_CHECK_LAUNCH("nicer_hash_encode_backward");
    return 0;
}

extern "C" int nicer_hash_encode_second_backward(const float *grad, const float *inputs, const float *embeddings,
                                                 const int32_t *offsets, uint32_t B, uint32_t D, uint32_t C,
                                                 uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                                 const float *dy_dx, const float *grad_grad_inputs, float *grad_grad,
                                                 float *grad2_embeddings, void *stream) {
    (void)embeddings; (void)calc_grad_inputs;
    if (B == 0) return 0;
    if (!grad || !inputs || !offsets || !dy_dx || !grad_grad_inputs || !grad_grad || !grad2_embeddings)
        NICER_FAIL(-1, "nicer_hash_encode_second_backward: NULL pointer");
    dim3 g(div_up(B, 256), L, 1);
    int e;
    if (D == 3) e = launch_second<3>(C, g, (cudaStream_t)stream, grad, inputs, offsets, dy_dx, grad_grad_inputs, grad_grad, grad2_embeddings, B, L, S, H);
    else if (D == 2) e = launch_second<2>(C, g, (cudaStream_t)stream, grad, inputs, offsets, dy_dx, grad_grad_inputs, grad_grad, grad2_embeddings, B, L, S, H);
    else NICER_FAIL(-1, "GridEncoding: D must be 2 or 3.");
    if (e) return e;
    NICER_CHECK_LAUNCH("nicer_hash_encode_second_backward");
    return 0;
}
