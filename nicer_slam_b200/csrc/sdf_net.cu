// Fused SDF network kernels (sm_100a): hash/dense grid gather + NeRF PE + weight-normed Softplus MLP +
// analytic d sdf/dx in one kernel; full first+second order backward in another.
// Replaces ImplicitNetworkGrid.forward / get_outputs / gradient / get_sdf_vals and the autograd double
// backward behind them (/root/reference/code/model/base_networks.py:155-238).
//
// Mapping: one thread per point, a CTA of 128 points; the network weights are staged once per
// (persistent) CTA in shared memory, transposed so that every weight read is a warp-broadcast LDS.128;
// each thread owns a private shared-memory column for the activations of the layer in flight.
// Saved tensors are feature-major ([row][P]) so every global access of a warp is a coalesced 128 B line.
#include "common.cuh"
#include "sdf_sample.cuh"

namespace nicer {

constexpr int SDF_BLOCK = 128;
constexpr int SDF_CS = SDF_BLOCK;  // column stride: thread t owns col[k*CS + t] (conflict-free)

struct SdfSmemLayout {
    int W0t, Wt[3], WLt, wl_sdf, b0, b[3], bl_feat, lv, col, total_floats;
};

static SdfSmemLayout sdf_layout(int n_hidden) {
    SdfSmemLayout s;
    int o = 0;
    s.W0t = o; o += COL_ROWS * NICER_W;
    for (int i = 0; i < 3; ++i) { s.Wt[i] = o; if (i < n_hidden - 1) o += NICER_W * NICER_W; }
    s.WLt = o; o += NICER_W * NICER_W;
    s.wl_sdf = o; o += NICER_W;
    s.b0 = o; o += NICER_W;
    for (int i = 0; i < 3; ++i) { s.b[i] = o; if (i < n_hidden - 1) o += NICER_W; }
    s.bl_feat = o; o += NICER_W;
    s.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    s.col = o; o += COL_ROWS * SDF_CS;
    s.total_floats = o;
    return s;
}

// Stage the network into shared memory (transposed) and build the device view.
__device__ void stage_sdf_net(const nicer_sdf_net_t &net, const LevelScales &ls, const SdfSmemLayout &lay, float *smem, SdfNetView &nv) {
    const int n = (int)net.n_hidden;
    const int L = (int)net.grid.L, C = (int)net.grid.C;
    const int d_pe = 3 + 6 * (int)net.multires;
    const int d_in = d_pe + L * C;
    const int tid = threadIdx.x, nt = blockDim.x;
    float *W0t = smem + lay.W0t;
    for (int i = tid; i < NICER_W * d_in; i += nt) {
        int j = i / d_in, k = i - j * d_in;
        W0t[k * NICER_W + j] = net.W[0][i];
    }
    for (int l = 1; l < n; ++l) {
        float *Wt = smem + lay.Wt[l - 1];
        for (int i = tid; i < NICER_W * NICER_W; i += nt) {
            int j = i / NICER_W, k = i - j * NICER_W;
            Wt[k * NICER_W + j] = net.W[l][i];
        }
        for (int i = tid; i < NICER_W; i += nt) smem[lay.b[l - 1] + i] = net.b[l][i];
    }
    {
        float *WLt = smem + lay.WLt;
        const int nfeat = (int)net.d_out - 1;  // <= 64
        for (int i = tid; i < NICER_W * NICER_W; i += nt) {
            int j = i / NICER_W, k = i - j * NICER_W;   // output feature j, input k
            WLt[k * NICER_W + j] = (j < nfeat) ? net.W[n][(size_t)(1 + j) * NICER_W + k] : 0.f;
        }
        for (int i = tid; i < NICER_W; i += nt) {
            smem[lay.wl_sdf + i] = net.W[n][i];
            smem[lay.b0 + i] = net.b[0][i];
            smem[lay.bl_feat + i] = (i < nfeat) ? net.b[n][1 + i] : 0.f;
        }
    }
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + lay.lv);
    for (int l = tid; l < L; l += nt) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    nv.W0t = W0t;
    for (int i = 0; i < 3; ++i) { nv.Wt[i] = smem + lay.Wt[i]; nv.b[i] = smem + lay.b[i]; }
    nv.WLt = smem + lay.WLt;
    nv.wl_sdf = smem + lay.wl_sdf;
    nv.b0 = smem + lay.b0;
    nv.bl_feat = smem + lay.bl_feat;
    nv.bl_sdf = net.b[n][0];
    nv.lv = lv;
    nv.table = net.grid.table;
    nv.L = L; nv.n_hidden = n; nv.multires = (int)net.multires; nv.d_pe = d_pe; nv.d_in = d_in;
    nv.df = net.grid.divide_factor;
}

template <int C>
__global__ void __launch_bounds__(SDF_BLOCK, 2)
sdf_forward_kernel(const nicer_sdf_net_t net, const LevelScales ls, const SdfSmemLayout lay, const float *__restrict__ X, uint32_t P,
                   uint32_t flags, float *sdf, float *feat_fm, float *grad, float *Z, float *R, float *DYDX, float *H0, uint32_t Pf) {
    extern __shared__ __align__(16) float smem[];
    SdfNetView nv;
    stage_sdf_net(net, ls, lay, smem, nv);
    __syncthreads();
    float *col = smem + lay.col + threadIdx.x;
    const uint32_t tiles = (P + SDF_BLOCK - 1) / SDF_BLOCK;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t p = t * SDF_BLOCK + threadIdx.x;
        if (p < P) sdf_forward_sample<C>(nv, X, p, P, flags, col, SDF_CS, sdf, feat_fm, grad, Z, R, DYDX, H0, Pf);
    }
}

template <int C>
__global__ void __launch_bounds__(SDF_BLOCK, 2)
sdf_backward_kernel(const nicer_sdf_net_t net, const LevelScales ls, const SdfSmemLayout lay, const float *__restrict__ X, uint32_t P,
                    const float *Z, const float *R, const float *DYDX, const float *g_sdf, const float *g_feat_fm,
                    const float *g_grad, float *grad_x, float *grad_table, float *ZB, float *QB, float *AB,
                    float *TAN, float *T0, uint32_t Pf) {
    extern __shared__ __align__(16) float smem[];
    SdfNetView nv;
    stage_sdf_net(net, ls, lay, smem, nv);
    __syncthreads();
    float *col = smem + lay.col + threadIdx.x;
    const uint32_t tiles = (P + SDF_BLOCK - 1) / SDF_BLOCK;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t p = t * SDF_BLOCK + threadIdx.x;
        if (p < P)
            sdf_backward_sample<C>(nv, X, p, P, Z, R, DYDX, g_sdf, g_feat_fm, g_grad, grad_x, grad_table, ZB, QB, AB,
                                   TAN, T0, col, SDF_CS, Pf);
    }
}

static int check_sdf_net(const nicer_sdf_net_t *net, const char *who) {
    if (!net) NICER_FAIL(-1, "%s: net is NULL", who);
    const uint32_t C = net->grid.C, L = net->grid.L;
    if (!(C == 2 || C == 4 || C == 8)) NICER_FAIL(-1, "%s: level_dim C must be 2, 4 or 8 (got %u)", who, C);
    if (L < 1 || L > NICER_MAX_LEVELS) NICER_FAIL(-1, "%s: num_levels must be in [1,%d] (got %u)", who, NICER_MAX_LEVELS, L);
    if (L * C > 32) NICER_FAIL(-1, "%s: L*C must be <= 32 (got %u)", who, L * C);
    if (net->multires > 6) NICER_FAIL(-1, "%s: multires must be <= 6 (got %u)", who, net->multires);
    if (net->n_hidden < 1 || net->n_hidden > 4) NICER_FAIL(-1, "%s: n_hidden must be in [1,4] (got %u)", who, net->n_hidden);
    if (net->d_out < 1 || net->d_out > 65) NICER_FAIL(-1, "%s: d_out must be in [1,65] (got %u)", who, net->d_out);
    if (!net->grid.table || !net->grid.offsets) NICER_FAIL(-1, "%s: grid pointers are NULL", who);
    for (uint32_t l = 0; l <= net->n_hidden; ++l)
        if (!net->W[l] || !net->b[l]) NICER_FAIL(-1, "%s: weight/bias %u is NULL", who, l);
    if (!(net->grid.divide_factor > 0.f)) NICER_FAIL(-1, "%s: divide_factor must be > 0", who);
    return 0;
}

bool tc_enabled();
int launch_sdf_only_tc(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t flags, float *sdf, float *F, cudaStream_t st);
int launch_sdf_forward_tc(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, uint32_t flags, float *sdf, float *feat_fm,
                          float *grad, float *Z, float *R, float *DYDX, float *H0, cudaStream_t st);

int launch_sdf_backward_tc(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t Pf, const float *Z, const float *R,
                           const float *DYDX, const float *H0, const float *g_sdf, const float *g_feat_fm, const float *g_grad, float *grad_x,
                           float *grad_table, float *ZB, float *QB, float *AB, float *TAN, float *T0, float *tan_sum, float *GY,
                           cudaStream_t st, cudaStream_t scatter_st);

// out[r] += sum_p rows[r][p]   (one block per row)
__global__ void __launch_bounds__(256) row_sum_accum_kernel(const float *__restrict__ rows, uint32_t P, float *out) {
    const float *row = rows + (size_t)blockIdx.x * P;
    float a = 0.f;
    for (uint32_t p = threadIdx.x; p < P; p += 256) a += row[p];
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out + blockIdx.x, t);
    }
}
int launch_row_sum_accum(const float *rows, uint32_t n_rows, uint32_t P, float *out, cudaStream_t st) {
    row_sum_accum_kernel<<<n_rows, 256, 0, st>>>(rows, P, out);
    NICER_CHECK_LAUNCH("nicer_sdf_backward(tan sum)");
    return 0;
}

template <typename K>
static int prep_kernel(K kernel, size_t smem_bytes, const char *who) {
    NICER_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes), who);
    return 0;
}

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_sdf_forward(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t P_feat, uint32_t flags, float *sdf,
                                 float *feat_fm, float *grad, float *Z, float *R, float *DYDX, float *H0, void *stream) {
    if (int e = check_sdf_net(net, "nicer_sdf_forward")) return e;
    if (P == 0) return 0;
    if (P_feat > P) NICER_FAIL(-1, "nicer_sdf_forward: P_feat (%u) > P (%u)", P_feat, P);
    const uint32_t Pf = P_feat ? P_feat : P;
    if (!x || !sdf) NICER_FAIL(-1, "nicer_sdf_forward: x/sdf is NULL");
    const bool sdf_only = flags & NICER_SDF_ONLY;
    if (!sdf_only) {
        if (!grad || !Z || !DYDX) NICER_FAIL(-1, "nicer_sdf_forward: grad/Z/DYDX required unless NICER_SDF_ONLY");
        if (net->n_hidden > 1 && !R) NICER_FAIL(-1, "nicer_sdf_forward: R required for n_hidden > 1");
        if (!(flags & NICER_SDF_NO_FEAT) && !feat_fm) NICER_FAIL(-1, "nicer_sdf_forward: feat_fm is NULL");
    }
    if (net->n_hidden > 3) NICER_FAIL(-1, "nicer_sdf_forward: n_hidden > 3 not built");
    // tensor-core (tcgen05, 3xTF32) path for the sdf-only pass; multires 6 + 64-wide layers is what it is built for
    if (sdf_only && tc_enabled() && net->multires == 6)
        return launch_sdf_only_tc(net, x, P, flags, sdf, H0, (cudaStream_t)stream);     // H0: optional [L*C][P] feature workspace
    if (!sdf_only && tc_enabled() && net->multires == 6)
        return launch_sdf_forward_tc(net, x, P, Pf, flags, sdf, feat_fm, grad, Z, R, DYDX, H0, (cudaStream_t)stream);
    SdfSmemLayout lay = sdf_layout((int)net->n_hidden);
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const size_t smem = (size_t)lay.total_floats * sizeof(float);
    const uint32_t tiles = div_up(P, SDF_BLOCK);
    const uint32_t grid = tiles < (uint32_t)(2 * num_sms()) ? tiles : (uint32_t)(2 * num_sms());
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(CC)                                                                                            \
    do {                                                                                                      \
        if (int e = prep_kernel(sdf_forward_kernel<CC>, smem, "nicer_sdf_forward")) return e;                 \
        sdf_forward_kernel<CC><<<grid, SDF_BLOCK, smem, st>>>(*net, ls, lay, x, P, flags, sdf, feat_fm, grad, Z, R, DYDX, H0, Pf); \
    } while (0)
    switch (net->grid.C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_sdf_forward");
    return 0;
}

extern "C" int nicer_sdf_backward(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t P_feat, const float *Z,
                                  const float *R, const float *DYDX, const float *H0, const float *g_sdf, const float *g_feat_fm,
                                  const float *g_grad, float *grad_x, float *grad_table, float *ZB, float *QB,
                                  float *AB, float *TAN, float *T0, float *tan_sum, float *GY, void *stream,
                                  void *scatter_stream) {
    if (int e = check_sdf_net(net, "nicer_sdf_backward")) return e;
    if (P == 0) return 0;
    if (P_feat > P) NICER_FAIL(-1, "nicer_sdf_backward: P_feat (%u) > P (%u)", P_feat, P);
    const uint32_t Pf = P_feat ? P_feat : P;
    if (!x || !Z || !DYDX || !ZB || !QB || !AB || !TAN || !T0)      /* grad_table may be NULL: no table gradient wanted */
        NICER_FAIL(-1, "nicer_sdf_backward: a required pointer is NULL");
    if (net->n_hidden > 1 && !R) NICER_FAIL(-1, "nicer_sdf_backward: R required for n_hidden > 1");
    if (net->n_hidden > 3) NICER_FAIL(-1, "nicer_sdf_backward: n_hidden > 3 not built");
    if (tc_enabled() && net->multires == 6 && !GY) NICER_FAIL(-1, "nicer_sdf_backward: GY workspace is NULL");
    if (tc_enabled() && net->multires == 6)
        return launch_sdf_backward_tc(net, x, P, Pf, Z, R, DYDX, H0, g_sdf, g_feat_fm, g_grad, grad_x, grad_table, ZB, QB, AB, TAN, T0,
                                      tan_sum, GY, (cudaStream_t)stream, (cudaStream_t)scatter_stream);
    SdfSmemLayout lay = sdf_layout((int)net->n_hidden);
    const LevelScales ls = host_level_scales(net->grid.L, net->grid.S, net->grid.H);
    const size_t smem = (size_t)lay.total_floats * sizeof(float);
    const uint32_t tiles = div_up(P, SDF_BLOCK);
    const uint32_t grid = tiles < (uint32_t)(2 * num_sms()) ? tiles : (uint32_t)(2 * num_sms());
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(CC)                                                                                              \
    do {                                                                                                        \
        if (int e = prep_kernel(sdf_backward_kernel<CC>, smem, "nicer_sdf_backward")) return e;                 \
        sdf_backward_kernel<CC><<<grid, SDF_BLOCK, smem, st>>>(*net, ls, lay, x, P, Z, R, DYDX, g_sdf, g_feat_fm, g_grad, \
                                                               grad_x, grad_table, ZB, QB, AB, TAN, T0, Pf);    \
    } while (0)
    switch (net->grid.C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_sdf_backward");
    if (tan_sum) return launch_row_sum_accum(TAN + (size_t)(net->n_hidden - 1) * NICER_W * P, NICER_W, P, tan_sum, st);
    return 0;
}
