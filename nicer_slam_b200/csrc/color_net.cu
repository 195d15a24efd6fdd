// Fused rendering-network kernels (sm_100a): color hash-grid gather + view PE + ReLU MLP + sigmoid.
// Replaces RenderingNetwork.forward (mode "idr") and its autograd backward
// (/root/reference/code/model/base_networks.py:333-392).  Same one-thread-per-point scheme as sdf_net.cu.
#include "common.cuh"
#include "color_sample.cuh"

namespace nicer {

constexpr int COL_BLOCK = 128;
constexpr int COLOR_MAX_DIN = 144;

struct ColorSmemLayout {
    int W0t, Wt[3], WL, b0, b[3], lv, col, total_floats;
};

static ColorSmemLayout color_layout(int n_hidden) {
    ColorSmemLayout s;
    int o = 0;
    s.W0t = o; o += COLOR_MAX_DIN * NICER_W;
    for (int i = 0; i < 3; ++i) { s.Wt[i] = o; if (i < n_hidden - 1) o += NICER_W * NICER_W; }
    s.WL = o; o += 4 * NICER_W;
    s.b0 = o; o += NICER_W;
    for (int i = 0; i < 3; ++i) { s.b[i] = o; if (i < n_hidden - 1) o += NICER_W; }
    s.lv = o; o += NICER_MAX_LEVELS * LEVEL_INFO_WORDS;
    s.col = o; o += NICER_W * COL_BLOCK;
    s.total_floats = o;
    return s;
}

__device__ void stage_color_net(const nicer_color_net_t &net, const LevelScales &ls, const ColorSmemLayout &lay, float *smem, ColorNetView &nv) {
    const int n = (int)net.n_hidden;
    const bool has_grid = net.grid.table != nullptr;
    const int L = has_grid ? (int)net.grid.L : 0, C = has_grid ? (int)net.grid.C : 0;
    const int d_view = 3 + 6 * (int)net.multires_view;
    const int F = (int)net.feature;
    const int d_in = 3 + d_view + 3 + F + L * C;
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
    for (int i = tid; i < 3 * NICER_W; i += nt) smem[lay.WL + i] = net.W[n][i];
    for (int i = tid; i < NICER_W; i += nt) smem[lay.b0 + i] = net.b[0][i];
    LevelInfo *lv = reinterpret_cast<LevelInfo *>(smem + lay.lv);
    for (int l = tid; l < L; l += nt) lv[l] = make_level(net.grid.offsets, (uint32_t)l, level_scale(ls, (uint32_t)l));
    nv.W0t = W0t;
    for (int i = 0; i < 3; ++i) { nv.Wt[i] = smem + lay.Wt[i]; nv.b[i] = smem + lay.b[i]; }
    nv.WL = smem + lay.WL;
    nv.b0 = smem + lay.b0;
    nv.bl[0] = net.b[n][0]; nv.bl[1] = net.b[n][1]; nv.bl[2] = net.b[n][2];
    nv.lv = lv;
    nv.table = net.grid.table;
    nv.L = L; nv.n_hidden = n; nv.multires_view = (int)net.multires_view; nv.d_view = d_view; nv.feature = F;
    nv.d_in = d_in; nv.off_normal = 3 + d_view; nv.off_feat = 3 + d_view + 3; nv.off_grid = 3 + d_view + 3 + F;
    nv.df = has_grid ? net.grid.divide_factor : 1.0f;
    nv.detached = net.grid_detached != 0;
}

template <int C>
__global__ void __launch_bounds__(COL_BLOCK, 2)
color_forward_kernel(const nicer_color_net_t net, const LevelScales ls, const ColorSmemLayout lay, const float *__restrict__ X,
                     const float *__restrict__ V, const float *__restrict__ N, const float *__restrict__ feat_fm,
                     uint32_t P, float *rgb, float *A_fm, float *DYDX, float *H0) {
    extern __shared__ __align__(16) float smem[];
    ColorNetView nv;
    stage_color_net(net, ls, lay, smem, nv);
    __syncthreads();
    float *col = smem + lay.col + threadIdx.x;
    const uint32_t tiles = (P + COL_BLOCK - 1) / COL_BLOCK;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t p = t * COL_BLOCK + threadIdx.x;
        if (p < P) color_forward_sample<C>(nv, X, V, N, feat_fm, p, P, col, COL_BLOCK, rgb, A_fm, DYDX, H0);
    }
}

template <int C>
__global__ void __launch_bounds__(COL_BLOCK, 2)
color_backward_kernel(const nicer_color_net_t net, const LevelScales ls, const ColorSmemLayout lay, const float *__restrict__ X,
                      const float *__restrict__ V, const float *__restrict__ N, const float *__restrict__ feat_fm,
                      uint32_t P, const float *rgb, const float *A_fm, const float *DYDX, const float *g_rgb,
                      float *grad_x, float *grad_view, float *grad_normals, float *grad_feat_fm, float *grad_table,
                      float *ZB, float *OB) {
    extern __shared__ __align__(16) float smem[];
    ColorNetView nv;
    stage_color_net(net, ls, lay, smem, nv);
    __syncthreads();
    float *col = smem + lay.col + threadIdx.x;
    const uint32_t tiles = (P + COL_BLOCK - 1) / COL_BLOCK;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t p = t * COL_BLOCK + threadIdx.x;
        if (p < P)
            color_backward_sample<C>(nv, X, V, N, feat_fm, p, P, rgb, A_fm, DYDX, g_rgb, grad_x, grad_view, grad_normals,
                                     grad_feat_fm, grad_table, ZB, OB, col, COL_BLOCK);
    }
}

static int check_color_net(const nicer_color_net_t *net, const char *who) {
    if (!net) NICER_FAIL(-1, "%s: net is NULL", who);
    const bool has_grid = net->grid.table != nullptr;
    if (has_grid) {
        const uint32_t C = net->grid.C, L = net->grid.L;
        if (!(C == 2 || C == 4 || C == 8)) NICER_FAIL(-1, "%s: level_dim C must be 2, 4 or 8 (got %u)", who, C);
        if (L < 1 || L > NICER_MAX_LEVELS || L * C > 32) NICER_FAIL(-1, "%s: bad grid shape L=%u C=%u", who, L, C);
        if (!net->grid.offsets) NICER_FAIL(-1, "%s: grid offsets NULL", who);
        if (!(net->grid.divide_factor > 0.f)) NICER_FAIL(-1, "%s: divide_factor must be > 0", who);
    }
    if (net->multires_view > 4) NICER_FAIL(-1, "%s: multires_view must be <= 4", who);
    if (net->feature > 64) NICER_FAIL(-1, "%s: feature must be <= 64", who);
    if (net->n_hidden < 1 || net->n_hidden > 3) NICER_FAIL(-1, "%s: n_hidden must be in [1,3]", who);
    for (uint32_t l = 0; l <= net->n_hidden; ++l)
        if (!net->W[l] || !net->b[l]) NICER_FAIL(-1, "%s: weight/bias %u is NULL", who, l);
    return 0;
}

int launch_color_forward_tc(const nicer_color_net_t *net, const float *x, const float *view, const float *normals, const float *feat_fm,
                            uint32_t P, float *rgb, float *A_fm, float *DYDX, float *H0, cudaStream_t st);
int launch_color_backward_tc(const nicer_color_net_t *net, const float *x, const float *view, const float *normals, uint32_t P,
                             const float *rgb, const float *A_fm, const float *DYDX, const float *g_rgb, float *grad_x,
                             float *grad_view, float *grad_normals, float *grad_feat_fm, float *grad_table, float *ZB, float *OB,
                             float *GY, cudaStream_t st, cudaStream_t scatter_st);

}  // namespace nicer

using namespace nicer;

extern "C" int nicer_color_forward(const nicer_color_net_t *net, const float *x, const float *view,
                                   const float *normals, const float *feat_fm, uint32_t P, float *rgb, float *A_fm,
                                   float *DYDX, float *H0, void *stream) {
    if (int e = check_color_net(net, "nicer_color_forward")) return e;
    if (P == 0) return 0;
    if (!x || !view || !normals || !feat_fm || !rgb || !A_fm) NICER_FAIL(-1, "nicer_color_forward: NULL pointer");
    {
        const int r = launch_color_forward_tc(net, x, view, normals, feat_fm, P, rgb, A_fm, DYDX, H0, (cudaStream_t)stream);
        if (r != 0) return r < 0 ? r : 0;
    }
    ColorSmemLayout lay = color_layout((int)net->n_hidden);
    const LevelScales ls = host_level_scales(net->grid.table ? net->grid.L : 0, net->grid.S, net->grid.H);
    const size_t smem = (size_t)lay.total_floats * sizeof(float);
    const uint32_t tiles = div_up(P, COL_BLOCK);
    const uint32_t grid = tiles < (uint32_t)(2 * num_sms()) ? tiles : (uint32_t)(2 * num_sms());
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t C = net->grid.table ? net->grid.C : 2;
#define LAUNCH(CC)                                                                                                   \
    do {                                                                                                             \
        NICER_CUDA(cudaFuncSetAttribute(color_forward_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_color_forward");                                                                           \
        color_forward_kernel<CC><<<grid, COL_BLOCK, smem, st>>>(*net, ls, lay, x, view, normals, feat_fm, P, rgb, A_fm, DYDX, H0); \
    } while (0)
    switch (C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_color_forward");
    return 0;
}

extern "C" int nicer_color_backward(const nicer_color_net_t *net, const float *x, const float *view,
                                    const float *normals, const float *feat_fm, uint32_t P, const float *rgb,
                                    const float *A_fm, const float *DYDX, const float *g_rgb, float *grad_x,
                                    float *grad_view, float *grad_normals, float *grad_feat_fm, float *grad_table,
                                    float *ZB, float *OB, float *GY, void *stream, void *scatter_stream) {
    if (int e = check_color_net(net, "nicer_color_backward")) return e;
    if (P == 0) return 0;
    if (!x || !view || !normals || !feat_fm || !rgb || !A_fm || !g_rgb || !grad_normals || !grad_feat_fm || !ZB || !OB)
        NICER_FAIL(-1, "nicer_color_backward: NULL pointer");
    if (net->grid.table && !net->grid_detached && grad_table && !GY)
        NICER_FAIL(-1, "nicer_color_backward: GY workspace required for the table gradient");
    {
        const int r = launch_color_backward_tc(net, x, view, normals, P, rgb, A_fm, DYDX, g_rgb, grad_x, grad_view, grad_normals,
                                               grad_feat_fm, grad_table, ZB, OB, GY, (cudaStream_t)stream,
                                               (cudaStream_t)scatter_stream);
        if (r != 0) return r < 0 ? r : 0;
    }
    ColorSmemLayout lay = color_layout((int)net->n_hidden);
    const LevelScales ls = host_level_scales(net->grid.table ? net->grid.L : 0, net->grid.S, net->grid.H);
    const size_t smem = (size_t)lay.total_floats * sizeof(float);
    const uint32_t tiles = div_up(P, COL_BLOCK);
    const uint32_t grid = tiles < (uint32_t)(2 * num_sms()) ? tiles : (uint32_t)(2 * num_sms());
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t C = net->grid.table ? net->grid.C : 2;
#define LAUNCH(CC)                                                                                                    \
    do {                                                                                                              \
        NICER_CUDA(cudaFuncSetAttribute(color_backward_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                   "nicer_color_backward");                                                                           \
        color_backward_kernel<CC><<<grid, COL_BLOCK, smem, st>>>(*net, ls, lay, x, view, normals, feat_fm, P, rgb, A_fm, DYDX, \
                                                                  g_rgb, grad_x, grad_view, grad_normals, grad_feat_fm, \
                                                                  grad_table, ZB, OB);                                \
    } while (0)
    switch (C) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        default: LAUNCH(8); break;
    }
#undef LAUNCH
    NICER_CHECK_LAUNCH("nicer_color_backward");
    return 0;
}
