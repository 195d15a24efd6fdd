// Per-point math shared by every kernel (and compiled for the host by tests/host_emul to check the
// formulas on a CPU-only box).  No reference code: the algorithms follow
// /root/reference/code/hashencoder/src/hashencoder.cu (index/hash :35-73, smoothstep :115-121,
// K1 :131-283, K2 :286-373, K5 :461-625) and torch's softplus / weight-normed Linear semantics
// (model/base_networks.py:149-179), restated for a one-thread-per-point formulation.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define NHD __host__ __device__ __forceinline__
#define NDEV __device__ __forceinline__
#else
#define NHD inline
#define NDEV inline
#include <string.h>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 v = {a, b, c, d}; return v; }
static inline float2 make_float2(float a, float b) { float2 v = {a, b}; return v; }
#endif

#define NICER_W 64  // hidden width

namespace nicer {

// ------------------------------------------------------------------ activations
// torch.nn.Softplus(beta=100, threshold=20) and its first / second derivative
// (ATen softplus, softplus_backward, softplus_double_backward).
constexpr float SP_BETA = 100.0f;
constexpr float SP_THRESH = 20.0f;

// Device fast path: one exp shared by softplus / sigmoid / sigmoid', built from ex2.approx + lg2.approx + rcp.approx
// with the argument-reduction error of exp and the rounding error of 1+e compensated, so that the result is fp32-faithful
// (|error| <~ 2e-7 relative on e, <~ 3e-9 absolute on softplus) at ~25 instructions instead of ~75 for libm's
// expf + log1pf + IEEE division.  The host build (tests/host_emul) keeps libm.
struct SpEval { float a, s1, s2; };   // softplus(z), softplus'(z), softplus''(z)

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exp(t) for t <= 20.  PRECISE: the rounding error of t*log2(e) is compensated (|rel err| ~ 2 ulp); otherwise the error
// is |t| * 2^-24 (<= 1.2e-6 for t <= 20), enough for the no-grad sampler pass.
template <bool PRECISE>
__device__ __forceinline__ float exp_fast(float t) {
    const float L2E = 1.4426950408889634f, L2E_LO = 1.925963033500803e-8f, LN2 = 0.6931471805599453f;
    const float y = t * L2E;
    const float e = ex2_approx(y);
    if (!PRECISE) return e;
    const float r = fmaf(t, L2E, -y) + t * L2E_LO;      // t*log2(e) - y, to ~2^-48
    return fmaf(e, r * LN2, e);
}
// softplus / its derivatives from one exponential.  log(1+e) = lg2.approx(1+e)*ln2 has an absolute error of ~2e-7
// (the approximation's 2^-22 in [1,2] plus the rounding of 1+e), i.e. ~2e-9 on softplus(z) = log(1+e)/100: the
// activations feed linear layers, where only the absolute error matters, so no extra correction is spent here.
template <bool PRECISE = true>
__device__ __forceinline__ SpEval sp_eval(float z) {
    SpEval r;
    const float t = z * SP_BETA;
    if (t > SP_THRESH) { r.a = z; r.s1 = 1.0f; r.s2 = 0.0f; return r; }
    const float e = exp_fast<PRECISE>(t);
    const float u = 1.0f + e;
    r.a = lg2_approx(u) * (0.6931471805599453f * 0.01f);
    r.s1 = e * rcp_approx(u);
    r.s2 = (1.0f - r.s1) * r.s1 * SP_BETA;
    return r;
}
__device__ __forceinline__ float softplus100(float z) { return sp_eval<true>(z).a; }
__device__ __forceinline__ float softplus100_fast(float z) { return sp_eval<false>(z).a; }
__device__ __forceinline__ float dsoftplus100(float z) { return sp_eval<true>(z).s1; }
__device__ __forceinline__ float d2softplus100(float z) { return sp_eval<true>(z).s2; }
#else
inline float softplus100(float z) {
    float t = z * SP_BETA;
    return t > SP_THRESH ? z : log1pf(expf(t)) / SP_BETA;
}
inline float dsoftplus100(float z) {
    float t = z * SP_BETA;
    if (t > SP_THRESH) return 1.0f;
    float e = expf(t);
    return e / (e + 1.0f);
}
inline float d2softplus100(float z) {
    float t = z * SP_BETA;
    if (!(t < SP_THRESH)) return 0.0f;
    float s = 1.0f / (1.0f + expf(-t));
    return (1.0f - s) * s * SP_BETA;
}
template <bool PRECISE = true>
inline SpEval sp_eval(float z) { SpEval r = {softplus100(z), dsoftplus100(z), d2softplus100(z)}; return r; }
inline float softplus100_fast(float z) { return softplus100(z); }
#endif

// x*y rounded to fp32 and never contracted into a following subtraction (the fractional part of x*scale must
// be taken from the ROUNDED product, as the un-fused evaluation does)
NHD float fmul_exact(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}

NHD float sstep(float v) { return v * v * (3.0f - 2.0f * v); }
NHD float sstep_d(float v) { return 6.0f * v * (1.0f - v); }

// NeRF positional encoding of one coordinate: sc[2f] = sin(2^f x), sc[2f+1] = cos(2^f x), f < F (model/embedder.py).
// DOUBLING (device, backward kernels only): sincosf at every STEP-th frequency, the ones between by angle doubling
// (<= 2 doublings, abs error < 5e-7).  The forward kernels keep one sincosf per frequency: rendered depths feed the
// discontinuous validity masks of the warp loss, where a last-bit difference can move a pixel across a mask edge.
template <int F, bool DOUBLING = false>
NHD void pe_sincos(float x, float *sc) {
#if defined(__CUDA_ARCH__)
    constexpr int STEP = DOUBLING ? ((F % 3 == 0) ? 3 : 2) : 1;
    float s = 0.f, c = 1.f, fr = 1.0f;
#pragma unroll
    for (int f = 0; f < F; ++f) {
        if (f % STEP == 0) {
            sincosf(x * fr, &s, &c);
        } else {
            const float s2 = 2.0f * s * c, c2 = fmaf(-2.0f * s, s, 1.0f);
            s = s2; c = c2;
        }
        sc[2 * f] = s; sc[2 * f + 1] = c;
        fr *= 2.0f;
    }
#else
    float fr = 1.0f;
    for (int f = 0; f < F; ++f) {
        sincosf(x * fr, &sc[2 * f], &sc[2 * f + 1]);
        fr *= 2.0f;
    }
#endif
}

// ------------------------------------------------------------------ grid geometry
struct LevelInfo {
    uint32_t offset;        // first entry of the level
    uint32_t hashmap_size;  // entries in the level
    uint32_t resolution;    // ceil(scale) + 1
    float scale;            // exp2f(level*S)*H - 1
    uint32_t dense;         // 1: p0 + res*(p1 + res*p2), 0: XOR hash   (decided per level, see level_is_dense)
    uint32_t mode;          // how index % hashmap_size is taken: LEVEL_MOD_MASK / _WRAP / _GENERAL (make_level)
    uint32_t mask;          // LEVEL_MOD_MASK: index % hashmap_size == index & mask for every index the level can produce
};
constexpr uint32_t LEVEL_MOD_GENERAL = 0u, LEVEL_MOD_MASK = 1u, LEVEL_MOD_WRAP = 2u;
constexpr int LEVEL_INFO_WORDS = sizeof(LevelInfo) / 4;

// Per-level scale exp2f(level*S)*H - 1 (hashencoder.cu:180).  The kernels evaluate it on the DEVICE with the reference's own
// expression, so that nvcc emits what it emits for the reference (MUFU.EX2 + one FFMA: see the SASS of kernel_grid in
// oracle/_ref) and sample positions are bit-identical to the reference's CUDA build on the same GPU: one ulp of `scale` at
// resolution 2048 already moves a sample by 1e-4 of a cell (tests/test_gpu_ref_cuda.py).  The host-compiled emulation of
// the per-point functions (tests/host_emul, CPU tests only) and the CPU oracle use the C library's exp2f; against those the
// GPU differs by that last bit of scale (~1e-5 relative on interpolated features at the finest levels).
struct LevelScales { float s[16]; float S; uint32_t H; };

inline LevelScales host_level_scales(uint32_t L, float S, uint32_t H) {
    LevelScales ls;
    for (uint32_t l = 0; l < 16; ++l) ls.s[l] = (l < L) ? exp2f((float)l * S) * (float)H - 1.0f : 0.f;
    ls.S = S;
    ls.H = H;
    return ls;
}

NHD float level_scale(const LevelScales &ls, uint32_t level) {
#ifdef __CUDA_ARCH__
    return exp2f(level * ls.S) * ls.H - 1.0f;
#else
    return ls.s[level];
#endif
}

NHD bool level_is_dense(const LevelInfo &li);

NHD LevelInfo make_level(const int32_t *offsets, uint32_t level, float scale) {
    LevelInfo li;
    li.offset = (uint32_t)offsets[level];
    li.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    li.scale = scale;
    li.resolution = (uint32_t)ceilf(li.scale) + 1u;
    li.dense = level_is_dense(li) ? 1u : 0u;
    // The reference reduces every index modulo hashmap_size (hashencoder.cu:72).  Power-of-two levels (all hashed levels
    // of the reference's configurations) take a mask.  For a dense level the largest vertex coordinate is
    // floor(scale) + 1 (u <= 1), so the largest index is known: below hashmap_size the modulo is the identity (mask of
    // all ones), below 2*hashmap_size (integer scales, where u == 1 reaches coordinate `resolution`) it is one
    // conditional subtraction; anything else keeps the general modulo.
    const uint32_t hs = li.hashmap_size;
    const unsigned long long r = li.resolution, pmax = (unsigned long long)floorf(li.scale) + 1ull;
    const unsigned long long max_index = pmax * (1ull + r + r * r);
    li.mask = 0u;
    if (hs != 0u && (hs & (hs - 1u)) == 0u) { li.mode = LEVEL_MOD_MASK; li.mask = hs - 1u; }
    else if (li.dense && max_index < (unsigned long long)hs) { li.mode = LEVEL_MOD_MASK; li.mask = 0xffffffffu; }
    else if (li.dense && max_index < 2ull * hs) li.mode = LEVEL_MOD_WRAP;
    else li.mode = LEVEL_MOD_GENERAL;
    return li;
}

// The reference's index rule (hashencoder.cu:54-73) accumulates p[d]*stride with stride *= resolution while
// stride <= hashmap_size (uint32 arithmetic, wraps) and switches to the XOR hash when the final stride exceeds the level
// size; which of the two happens depends on the level only, so it is decided once per level.
NHD bool level_is_dense(const LevelInfo &li) {
    uint32_t stride = 1;
    const uint32_t hs = li.hashmap_size, res = li.resolution;
    if (stride <= hs) stride *= res;
    if (stride <= hs) stride *= res;
    if (stride <= hs) stride *= res;
    return !(stride > hs);
}

// entry index (not multiplied by C) of grid vertex p within a level
NHD uint32_t vertex_index3(const LevelInfo &li, bool dense, uint32_t px, uint32_t py, uint32_t pz) {
    const uint32_t res = li.resolution;
    return dense ? px + res * (py + res * pz) : ((px * 1u) ^ (py * 2654435761u) ^ (pz * 805459861u));
}

// Interpolation cell of a point u in [0,1]^3 at one level.
struct Cell3 {
    uint32_t pg[3];
    float w[3];   // smoothstep(frac)
    float dw[3];  // smoothstep'(frac)
    float scale;
    bool inside;
};

NHD Cell3 locate3(const LevelInfo &li, const float u[3]) {
    Cell3 c;
    c.inside = !(u[0] < 0.f || u[0] > 1.f || u[1] < 0.f || u[1] > 1.f || u[2] < 0.f || u[2] > 1.f);
    c.scale = li.scale;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float p = fmul_exact(u[d], li.scale);
        float f = floorf(p);
        c.pg[d] = (uint32_t)f;
        p -= (float)c.pg[d];
        c.dw[d] = sstep_d(p);
        c.w[d] = sstep(p);
    }
    return c;
}

NHD void corner_indices(const LevelInfo &li, const Cell3 &c, uint32_t idx[8]) {
    const bool dense = li.dense != 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        idx[k] = vertex_index3(li, dense, c.pg[0] + (k & 1), c.pg[1] + ((k >> 1) & 1), c.pg[2] + ((k >> 2) & 1));
    const uint32_t hs = li.hashmap_size;
    if (li.mode == LEVEL_MOD_MASK) {       // level-uniform branches
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] &= li.mask;
    } else if (li.mode == LEVEL_MOD_WRAP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = idx[k] >= hs ? idx[k] - hs : idx[k];
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = idx[k] >= hs ? idx[k] % hs : idx[k];
    }
}

// trilinear weights in the reference's corner order (bit d of k selects the upper vertex in dim d)
NHD void corner_weights(const Cell3 &c, float wt[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float w = 1.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) w *= ((k >> d) & 1) ? c.w[d] : 1.0f - c.w[d];
        wt[k] = w;
    }
}

// d(feature)/d(u_gd) weights per corner: scale * s'(f_gd) * prod_{d != gd} w_d, signed (+ upper, - lower)
NHD void corner_dweights(const Cell3 &c, int gd, float wt[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float w = c.scale;
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (d != gd) w *= ((k >> d) & 1) ? c.w[d] : 1.0f - c.w[d];
        w *= c.dw[gd];
        wt[k] = ((k >> gd) & 1) ? w : -w;
    }
}

template <int C>
NHD void load_entry(const float *table, const LevelInfo &li, uint32_t idx, float out[C]) {
    const float *p = table + ((size_t)li.offset + idx) * C;
    if constexpr (C == 8) {
        float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
        out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else if constexpr (C == 4) {
        float4 a = *reinterpret_cast<const float4 *>(p);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    } else if constexpr (C == 2) {
        float2 a = *reinterpret_cast<const float2 *>(p);
        out[0] = a.x; out[1] = a.y;
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c] = p[c];
    }
}

// Level encode of a point: features feat[C] and (optionally) d feat / d u  dfeat[3][C].
// Out-of-range points give zeros (hashencoder.cu:152-177).
template <int C, bool WITH_DX>
NHD void encode_level(const float *table, const LevelInfo &li, const float u[3], float feat[C], float dfeat[3][C]) {
    Cell3 c = locate3(li, u);
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        feat[ch] = 0.f;
        if (WITH_DX) { dfeat[0][ch] = 0.f; dfeat[1][ch] = 0.f; dfeat[2][ch] = 0.f; }
    }
    if (!c.inside) return;
    uint32_t idx[8];
    corner_indices(li, c, idx);
    float val[8][C];
#pragma unroll
    for (int k = 0; k < 8; ++k) load_entry<C>(table, li, idx[k], val[k]);
    float wt[8];
    corner_weights(c, wt);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) feat[ch] += wt[k] * val[k][ch];
    if (WITH_DX) {
#pragma unroll
        for (int gd = 0; gd < 3; ++gd) {
            // the reference sums (right-left) pairs: w * (v_r - v_l) * s'(f)  (hashencoder.cu:245-276)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((k >> gd) & 1) continue;
                float w = c.scale;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    if (d != gd) w *= ((k >> d) & 1) ? c.w[d] : 1.0f - c.w[d];
                const int kr = k | (1 << gd);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) dfeat[gd][ch] += w * (val[kr][ch] - val[k][ch]) * c.dw[gd];
            }
        }
    }
}

// ------------------------------------------------------------------ small dense helpers
// acc[0..63] += a * row[0..63]   (row 16-byte aligned)
NHD void axpy64(float acc[NICER_W], const float *row, float a) {
    const float4 *w = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int j = 0; j < NICER_W / 4; ++j) {
        float4 v = w[j];
        acc[4 * j + 0] += v.x * a;
        acc[4 * j + 1] += v.y * a;
        acc[4 * j + 2] += v.z * a;
        acc[4 * j + 3] += v.w * a;
    }
}

// sum_j row[j] * q[j]
NHD float dot64(const float *row, const float q[NICER_W]) {
    const float4 *w = reinterpret_cast<const float4 *>(row);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < NICER_W / 4; ++j) {
        float4 v = w[j];
        s0 += v.x * q[4 * j + 0];
        s1 += v.y * q[4 * j + 1];
        s2 += v.z * q[4 * j + 2];
        s3 += v.w * q[4 * j + 3];
    }
    return (s0 + s1) + (s2 + s3);
}

// acc += Wt[K][64]^T-applied: acc[j] += sum_k Wt[k][j] * col[k*cs]
NHD void mv_acc64(float acc[NICER_W], const float *Wt, const float *col, int cs, int K) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) axpy64(acc, Wt + k * NICER_W, col[k * cs]);
}

}  // namespace nicer
