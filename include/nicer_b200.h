/*
 * nicer_b200.h — C ABI of libnicer_b200.so (hand-written sm_100a CUDA for the NICER-SLAM
 * volume-rendering hot path).  Plain pointers and sizes only: every pointer is a DEVICE pointer
 * unless named host_*, `stream` is a cudaStream_t passed as void*, and no function allocates:
 * the caller owns every buffer (same ownership rule as the reference's native op).
 *
 * Every function returns 0 on success and a negative code on error; nicer_last_error() returns
 * a thread-local message.  Launches are asynchronous on `stream`.
 *
 * Reference interfaces replaced (paths relative to /root/reference/code):
 *   nicer_hash_encode_forward          <- hash_encode_forward          hashencoder/src/hashencoder.h:13, hashencoder.cu:758-781
 *   nicer_hash_encode_backward         <- hash_encode_backward         hashencoder.h:14,  hashencoder.cu:783-813
 *   nicer_hash_encode_second_backward  <- hash_encode_second_backward  hashencoder.h:15,  hashencoder.cu:816-854
 *   nicer_sdf_forward / _backward      <- ImplicitNetworkGrid.forward/get_outputs/gradient + autograd double backward
 *                                         model/base_networks.py:155-221 (hash encode + NeRF PE + weight-normed
 *                                         Softplus(100) MLP + d sdf/dx computed in-kernel)
 *   nicer_color_forward / _backward    <- RenderingNetwork.forward (mode "idr")  model/base_networks.py:333-392
 *   nicer_outer_accum                  <- the nn.Linear weight/bias gradient GEMMs autograd runs for the above
 *   nicer_composite_forward/_backward  <- GridPredefineDensity (model/density.py:33-67) + SLAMNetwork.volume_rendering
 *                                         and the weighted sums (model/network.py:137-151,338-345,349-370)
 *   nicer_voxel_count                  <- SLAMNetwork.update_voxels  model/network.py:62-76
 */
#ifndef NICER_B200_H
#define NICER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NICER_MAX_LEVELS 16
#define NICER_HIDDEN 64          /* hidden width of all three MLPs in every shipped conf */
#define NICER_MAX_HIDDEN_LAYERS 4

const char *nicer_last_error(void);
int nicer_version(void);
/* 1 (default): the sdf-only pass runs on the tcgen05 (3xTF32) kernel; 0: fp32 SIMT kernel (A/B testing).
 * Also controlled by the environment variable NICER_DISABLE_TC=1. */
int nicer_set_tensor_cores(int enabled);

/* ---------------------------------------------------------------------------------------------
 * Drop-in native op of hashencoder/ (same argument meaning and layouts as hashencoder.h:13-15).
 *   inputs  [B,D] fp32 in [0,1] (points outside produce zeros / are skipped)
 *   embeddings [N,C] fp32; offsets [L+1] int32 (device)
 *   outputs / grad / grad_grad [L,B,C];  dy_dx [B,L,D,C]
 *   accumulators (grad_embeddings, grad_inputs, grad2_embeddings) must arrive zeroed.
 *   D in {2,3}; C in {1,2,4,8} (second_backward: C in {2,4,8}, as the reference).
 * ------------------------------------------------------------------------------------------- */
int nicer_hash_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                              float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                              uint32_t H, int calc_grad_inputs, float *dy_dx, void *stream);

int nicer_hash_encode_backward(const float *grad, const float *inputs, const float *embeddings,
                               const int32_t *offsets, float *grad_embeddings, uint32_t B, uint32_t D,
                               uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                               const float *dy_dx, float *grad_inputs, void *stream);

int nicer_hash_encode_second_backward(const float *grad, const float *inputs, const float *embeddings,
                                      const int32_t *offsets, uint32_t B, uint32_t D, uint32_t C,
                                      uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                      const float *dy_dx, const float *grad_grad_inputs, float *grad_grad,
                                      float *grad2_embeddings, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused networks.  Conventions:
 *   P            number of points (ray-samples)
 *   x            [P,3] row-major world points in [-1,1]^3
 *   "fm" buffers feature-major: [rows][P] (row stride P), so a warp of consecutive points is coalesced
 *   weights      effective (weight-normed) matrices, row-major [out,in], fp32
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float *table;        /* [n_entries, C] */
    const int32_t *offsets;    /* device [L+1] */
    uint32_t L, C, H;          /* levels, features per level (2|4|8), base resolution */
    float S;                   /* log2(per_level_scale) (hashgrid.py:31) */
    float divide_factor;       /* ImplicitNetworkGrid.divide_factor (base_networks.py:158) */
} nicer_grid_t;

typedef struct {
    nicer_grid_t grid;
    uint32_t multires;         /* NeRF PE frequencies for x (6 in all confs) */
    uint32_t n_hidden;         /* hidden layers of width 64: coarse 1, fine 3 */
    uint32_t d_out;            /* 1 + feature_vector_size (65) */
    const float *W[NICER_MAX_HIDDEN_LAYERS + 1];   /* W[0] [64,d_in], W[1..n-1] [64,64], W[n] [d_out,64] */
    const float *b[NICER_MAX_HIDDEN_LAYERS + 1];
} nicer_sdf_net_t;

/* Workspace saved by nicer_sdf_forward for nicer_sdf_backward (all fm):
 *   Z    [n_hidden*64][P]   pre-activations z_l
 *   R    [(n_hidden-1)*64][P] adjoints r_l = d sdf / d a_l for l<n (r_n is W[n][0,:])   (may be NULL if n_hidden==1)
 *   DYDX [L*3*C][P]         d enc / d u   (K1's dy_dx, feature-major)
 *   H0   [d_in][P]          the network input [x, PE(x), grid(x)] (operand of the layer-0 weight gradient, so that
 *                           the backward pass never re-reads the grid)
 */
#define NICER_SDF_ONLY 1u        /* sdf only: no feat, no gradient, nothing saved (sampler pass, get_sdf_vals); H0, when
                                    not NULL, is a [L*C][P] scratch for the grid features (gathered by a separate
                                    high-occupancy kernel instead of inside the tensor-core kernel) */
#define NICER_SDF_FEATURES_READY 8u /* with NICER_SDF_ONLY: H0 already holds the grid features of x (skip the gather kernel) */
#define NICER_SDF_ACCUMULATE 2u  /* add into sdf/feat/grad instead of overwriting (coarse+fine sum, base_networks.py:40) */
#define NICER_SDF_NO_FEAT 4u     /* gradient() path: sdf + gradient only (base_networks.py:195-206) */

/* P_feat (both calls): only the first P_feat points carry features and upstream sdf / feature gradients -- feat_fm, g_feat_fm are
 * [64][P_feat] and g_sdf is [P_feat]; the points behind them (the eikonal samples of network.py:313-336 batched behind the
 * main-pass points, which only need d sdf/dx) skip the feature head.  0 means P (every point). */
int nicer_sdf_forward(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t P_feat, uint32_t flags,
                      float *sdf /*[P]*/, float *feat_fm /*[64][P_feat]*/, float *grad /*[P,3]*/,
                      float *Z, float *R, float *DYDX, float *H0, void *stream);

/* Backward of (sdf, feat, grad) w.r.t. x, the grid and (through the workspace below) the weights.
 *   H0 [d_in][P] | NULL       the saved network input (its x / positional-encoding rows are read back instead of recomputed)
 *   g_sdf [P_feat] | NULL, g_feat_fm [64][P_feat] | NULL, g_grad [P,3] | NULL   upstream gradients
 *   grad_x [P,3] | NULL       accumulated (+=)
 *   grad_table [n_entries,C]  accumulated with atomics (first- and second-order terms both land here)
 *   Outputs for nicer_outer_accum (all fm, written):
 *     ZB  [n_hidden*64][P]  dL/dz_l          QB  [n_hidden*64][P]  q_l = r_l * softplus'(z_l)
 *     AB  [n_hidden*64][P]  a_l              TAN [n_hidden*64][P]  tangent of a_l in direction g_grad
 *     T0  [d_in][P]         tangent of the network input (H0, the input itself, is saved by the forward)
 *     tan_sum [64] | NULL   accumulated (+=): sum over the points of tan_n, the second-order part of dL/dW_n[0,:] (the sdf row
 *                           of the last layer), so that the caller needs no separate reduction over TAN
 *   Workspace: GY [2*L*C][P]  dL/d(enc) (first-order rows, then the second-order rows) handed from the MLP backward
 *                             kernel to the grid-scatter kernel
 */
int nicer_sdf_backward(const nicer_sdf_net_t *net, const float *x, uint32_t P, uint32_t P_feat,
                       const float *Z, const float *R, const float *DYDX, const float *H0,
                       const float *g_sdf, const float *g_feat_fm, const float *g_grad,
                       float *grad_x, float *grad_table,
                       float *ZB, float *QB, float *AB, float *TAN, float *T0, float *tan_sum, float *GY, void *stream,
                       void *scatter_stream);
/* scatter_stream (both backward calls): stream the grid-scatter kernel runs on, after everything enqueued on `stream`
 * by the call (NULL or == stream: same stream).  grad_table is complete when scatter_stream has drained; the caller
 * joins the streams.  Lets the atomics-bound scatter overlap with the weight-gradient GEMMs that follow. */

typedef struct {
    nicer_grid_t grid;         /* grid.table == NULL: no color grid (use_grid_feature=false) */
    uint32_t multires_view;    /* 4 */
    uint32_t feature;          /* 64 */
    uint32_t n_hidden;         /* 2 */
    uint32_t grid_detached;    /* color_stage == "base" (base_networks.py:337-339): no grid / x gradient through it */
    const float *W[NICER_MAX_HIDDEN_LAYERS + 1];   /* W[0] [64,d_in], ..., W[n] [3,64] */
    const float *b[NICER_MAX_HIDDEN_LAYERS + 1];
} nicer_color_net_t;

/* rgb [P,3] = sigmoid(MLP([x, PE(view), normals, feat, grid(x)])).  A_fm [n_hidden*64][P] saved (post-ReLU).
 * DYDX [L*3*C][P] saved only when want_dx (x requires grad and grid not detached), else NULL.
 * H0 [d_in][P] | NULL: the network input saved for the layer-0 weight gradient; only rows [0,33) (x, PE(view),
 * normals) and [33+feature, d_in) (grid) are written -- rows [33,33+feature) are feat_fm itself. */
int nicer_color_forward(const nicer_color_net_t *net, const float *x, const float *view /*[P,3]*/,
                        const float *normals /*[P,3]*/, const float *feat_fm /*[64][P]*/, uint32_t P,
                        float *rgb, float *A_fm, float *DYDX, float *H0, void *stream);

/* g_rgb [P,3] upstream.  Outputs: grad_x (+=, NULL ok), grad_view [P,3] (written, NULL ok),
 * grad_normals [P,3] (written), grad_feat_fm [64][P] (written), grad_table (atomics, NULL when detached),
 * ZB [n_hidden*64][P] dL/dz_l, OB [3][P] dL/d(pre-sigmoid) (for nicer_outer_accum).
 * Workspace: GY [L*C][P] dL/d(enc) for the grid-scatter kernel (NULL ok when there is no grid or it is detached). */
int nicer_color_backward(const nicer_color_net_t *net, const float *x, const float *view,
                         const float *normals, const float *feat_fm, uint32_t P, const float *rgb,
                         const float *A_fm, const float *DYDX, const float *g_rgb,
                         float *grad_x, float *grad_view, float *grad_normals, float *grad_feat_fm,
                         float *grad_table, float *ZB, float *OB, float *GY, void *stream, void *scatter_stream);

/* C[M,N] (row stride ldc) += A[M][P] * B[N][P]^T ;  bias[M] += rowsum(A) when bias != NULL.
 * A, B feature-major with row strides lda, ldb (>= P).  M <= 64, N <= 144. */
int nicer_outer_accum(const float *A, uint32_t lda, uint32_t M, const float *B, uint32_t ldb, uint32_t N,
                      uint32_t P, float *C, uint32_t ldc, float *bias, void *stream);

/* Several contractions over the SAME P samples in one call (the weight gradients of one network backward): one kernel
 * launch per 8 jobs instead of one per job.  `jobs` is a HOST array; each job has the meaning of nicer_outer_accum. */
typedef struct {
    const float *A; uint32_t lda, M;
    const float *B; uint32_t ldb, N;
    float *C; uint32_t ldc;
    float *bias;
} nicer_oa_job_t;
int nicer_outer_accum_batch(const nicer_oa_job_t *jobs, uint32_t n_jobs, uint32_t P, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Density + alpha compositing, one warp per ray.  R rays x S samples (S <= 1024), P = R*S.
 *   sdf [P], x [P,3] (for the beta lookup), z [R,S], rgb [P,3], grad [P,3] (SDF gradients), voxels [res^3]
 *   outputs: weights [R,S], rgb_out [R,3], depth_out [R] (= sum w z / (sum w + 1e-8)), normal_out [R,3]
 *            (= sum w g/(|g|+1e-6), before the R^T rotation), wsum [R]
 * ------------------------------------------------------------------------------------------- */
int nicer_composite_forward(const float *sdf, const float *x, const float *z, const float *rgb,
                            const float *grad, const float *voxels, uint32_t voxel_res, uint32_t R,
                            uint32_t S, float *weights, float *rgb_out, float *depth_out,
                            float *normal_out, float *wsum, void *stream);

/* upstream: g_rgb_out [R,3] | NULL, g_depth_out [R] | NULL, g_normal_out [R,3] | NULL, g_weights [R,S] | NULL.
 * outputs (written): g_sdf [P], g_rgb [P,3], g_grad [P,3]. */
int nicer_composite_backward(const float *sdf, const float *x, const float *z, const float *rgb,
                             const float *grad, const float *voxels, uint32_t voxel_res, uint32_t R,
                             uint32_t S, const float *weights, const float *depth_out, const float *wsum,
                             const float *g_rgb_out, const float *g_depth_out, const float *g_normal_out,
                             const float *g_weights, float *g_sdf, float *g_rgb, float *g_grad, void *stream);

/* Sampler weights: density + transmittance for the no-grad pass (ray_sampler.py:105-112). weights [R,S]. */
int nicer_sampler_weights(const float *sdf, const float *x, const float *z, const float *voxels,
                          uint32_t voxel_res, uint32_t R, uint32_t S, float *weights, void *stream);

/* ---- optical-flow projection i -> j (model/network.py:153-165): flow[e,p] = pix(K_e * (w2c_e * (loc[i] + depth[i,p] * dirs[i,p]))) - uv[i,p]
 * with i = idii[e]; w2c / K are those of the edge's target frame.  Backward accumulates into zeroed g_depth, g_dirs, g_loc, g_w2c. */
int nicer_flow_project(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K, const float *uv,
                       const int64_t *idii, uint32_t E, uint32_t n, float *flow, void *stream);
int nicer_flow_project_backward(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                                const int64_t *idii, uint32_t E, uint32_t n, const float *g_flow, float *g_depth, float *g_dirs,
                                float *g_loc, float *g_w2c, void *stream);

/* ---- fused dense Adam step (+ gradient zeroing) for the hash-grid tables; bit-identical to
 * torch.optim.Adam(betas=(beta1,beta2), eps=eps) without amsgrad / weight decay (volsdf_train.py:174, 547, 576).
 * All four arrays hold n fp32 values, 16-byte aligned; step is the 1-based update count. */
int nicer_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, double lr, double beta1,
                    double beta2, double eps, uint64_t step, int zero_grad, void *stream);
int nicer_set_adam_variant(int variant);   /* test hook: rounding variant of the update (csrc/adam.cu) */

/* ---- weight_norm of all Linear layers of a network in one launch (torch._weight_norm, dim 0; model/base_networks.py:149-153):
 * forward: w[r,:] = g[r] * v[r,:] / ||v[r,:]||, norm[r] = ||v[r,:]|| (norm optional); backward (needs norm): dv, dg from dw. */
#define NICER_WN_MAX_JOBS 8
typedef struct {
    const float *v;      /* [rows, cols] */
    const float *g;      /* [rows] */
    float *w;            /* forward out [rows, cols] */
    float *norm;         /* forward out / backward in [rows] */
    const float *dw;     /* backward in [rows, cols] */
    float *dv;           /* backward out [rows, cols] */
    float *dg;           /* backward out [rows] */
    uint32_t rows, cols;
} nicer_wn_job_t;
int nicer_weight_norm(const nicer_wn_job_t *jobs, uint32_t n, void *stream);
int nicer_weight_norm_backward(const nicer_wn_job_t *jobs, uint32_t n, void *stream);

/* Coarse depths of the hierarchical sampler (UniformSampler.get_z_vals + near_far_from_cube, model/ray_sampler.py:21-61):
 * far[r] = exit of ray r from the cube [-bound, bound]^3 clamped to far_cap (use_cube = take_sphere_intersection; else
 * far_cap itself), z[r, i] = near*(1-t_i) + far*t_i with t = linspace(0, 1, N), stratified inside the half-way intervals
 * with the caller's uniform draws rnd [R,N] (NULL: no jitter, model.eval()); points [R*N,3] = cam_loc + z*ray_dir
 * (optional).  cam_loc, ray_dirs [R,3]. */
int nicer_sampler_uniform(const float *cam_loc, const float *ray_dirs, float near, float far_cap, float bound, int use_cube,
                          const float *rnd, uint32_t R, uint32_t N, float *z, float *far, float *points, void *stream);

/* The rest of ImportantSampler.get_z_vals (model/ray_sampler.py:105-159), one warp per ray: weights of the U coarse
 * samples from their SDF (density, transmittance), pdf = (w[:-1]+1e-5)/sum, cdf, inverse CDF at linspace(0,1,N)
 * (searchsorted right=True, the reference's denom < 1e-5 rule), merged with near, far[r] and the n_extra coarse depths
 * z[r, sel[j]], sorted -> z_out [R, N+2+n_extra]; z_eik[r] = z_out[r, eik_idx[r]] (optional).  weights [R,U] optional
 * output.  sel, eik_idx: int64 (torch.randperm / torch.randint results).  U <= 1024, N+2+n_extra <= 256. */
int nicer_sampler_resample(const float *sdf, const float *x, const float *z, const float *voxels, uint32_t voxel_res,
                           uint32_t R, uint32_t U, uint32_t N, const int64_t *sel, uint32_t n_extra, float near,
                           const float *far, const int64_t *eik_idx, float *z_out, float *z_eik, float *weights,
                           void *stream);

/* voxels[idx(x)] += 1 for every point with all |x_i| <= 0.99 (network.py:62-76). */
int nicer_voxel_count(const float *x, uint32_t P, float *voxels, uint32_t voxel_res, void *stream);

/* ---- SLAMLoss terms over rays and eikonal points (model/loss.py:113-233, utils/MiDaS.py:6-140) in three kernels.
 * A NULL prediction switches its term off.  terms[] receives the UNWEIGHTED term values and, in NICER_LOSS_SUM,
 * sum_i w_i * term_i; every g_* output (NULL ok) receives d(that sum)/d(input).
 * Workspace: acc double[8 + 5*B], maskf float[R] (the foreground & gt mask per ray). */
#define NICER_LOSS_MAX_FRAMES 64
enum { NICER_LOSS_RGB = 0, NICER_LOSS_DEPTH, NICER_LOSS_GT_DEPTH, NICER_LOSS_NORMAL_L1, NICER_LOSS_NORMAL_COS,
       NICER_LOSS_EIKONAL, NICER_LOSS_SMOOTH, NICER_LOSS_SUM, NICER_LOSS_TERMS };
typedef struct {
    uint32_t R, S, B, N, G;             /* rays = B frames x N pixels, S samples per ray, G eikonal points */
    uint32_t depth_mask_all;            /* depth term uses an all-ones mask (loss.py:171-173, Replica scan 4) */
    const float *sdf;                   /* [R,S]: foreground = sign change along the ray (loss.py:165-167) */
    const float *mask_gt;               /* [R]   : ground_truth["mask"] (> 0.5) */
    const float *rgb_pred, *rgb_gt;     /* [R,3] : L1 */
    const float *depth_pred;            /* [R]   : rendered depth (depth_values) */
    const float *depth_gt;              /* [R]   : mono depth; target = depth_gt * 50 + 0.5 (loss.py:91); NULL: SSI term off */
    const float *gt_depth;              /* [R]   : sensor-depth target, NULL: term off */
    const float *gt_depth_valid;        /* [R]   : ground_truth["gt_depth"]; its > 0 entries are averaged */
    const float *normal_pred, *normal_gt;          /* [R,3] */
    const float *grad_theta, *grad_theta_nei;      /* [G,3] (nei NULL: no smoothness term) */
    float w_rgb, w_depth, w_gt_depth, w_normal_l1, w_normal_cos, w_eik, w_smooth;
    float *g_rgb, *g_depth, *g_normal, *g_theta, *g_theta_nei;
} nicer_loss_t;
int nicer_slam_loss(const nicer_loss_t *args, double *acc, float *maskf, float *terms, void *stream);

/* ---- photometric warp sampling (model/network.py:167-279): every pixel of every frame i, lifted with its rendered depth,
 * is projected into every frame t of the batch and t's colour image is sampled bilinearly (grid_sample, align_corners).
 *   depth [B*N], dirs [B*N*pp,3] and loc [B,3] (rays of the patch pixels), w2c [B,4,4], K [B,4,4], img [B,H,W,3]
 *   -> sampled [B(t)][B*N*pp][3], mask [B(t)][B*N*pp] (uint8: inside the image and in front of the camera)
 * backward: g_sampled -> g_dirs (written), g_depth / g_loc / g_w2c (accumulated with atomics: must arrive zeroed). */
int nicer_warp_sample(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                      const float *img, uint32_t B, uint32_t N, uint32_t pp, uint32_t H, uint32_t W, float *sampled,
                      uint8_t *mask, void *stream);
int nicer_warp_sample_backward(const float *depth, const float *dirs, const float *loc, const float *w2c, const float *K,
                               const float *img, uint32_t B, uint32_t N, uint32_t pp, uint32_t H, uint32_t W,
                               const float *g_sampled, float *g_depth, float *g_dirs, float *g_loc, float *g_w2c, void *stream);

/* nicer_warp_gt: ground-truth colour / depth of the patch pixels in their own frame (model/network.py:226-246)
 *   uvp [B,M,2] pixel coordinates, img [B,H,W,3], dep [B,H,W] -> gt_rgb [B,M,3], gt_depth [B,M] (1 outside the image),
 *   inside [B,M] (uint8) */
int nicer_warp_gt(const float *uvp, const float *img, const float *dep, uint32_t B, uint32_t M, uint32_t H, uint32_t W,
                  float *gt_rgb, float *gt_depth, uint8_t *inside, void *stream);

/* nicer_masked_l1_mean: mean |a - b| over the selected entries = torch.abs(a[mask] - b[mask]).mean(), the photometric-warp
 * and optical-flow terms (model/loss.py:93-104,145-152).  a [n_mask*inner], mask [n_mask] (uint8, one per `inner` values),
 * b [b_len] read as b[i % b_len] (b_len = n_mask*inner when not broadcast) -> out[0] = mean, out[1] = number of selected values.
 * workspace: nicer_masked_l1_mean_workspace() bytes of device memory, 8-byte aligned (per-block partials; no initialisation needed).
 * backward: g [1] = dL/dmean -> ga [n_mask*inner] (written; 0 on unselected entries). */
size_t nicer_masked_l1_mean_workspace(void);
int nicer_masked_l1_mean(const float *a, const float *b, const unsigned char *mask, uint32_t n_mask, uint32_t inner,
                         uint32_t b_len, void *workspace, float *out, void *stream);
int nicer_masked_l1_mean_backward(const float *a, const float *b, const unsigned char *mask, uint32_t n_mask, uint32_t inner,
                                  uint32_t b_len, const float *out, const float *g, float *ga, void *stream);

/* ---- camera / ray helpers (one kernel each; the reference runs them as ~200 elementwise kernels per iteration)
 * nicer_pose_from_cam7            <- get_camera_from_tensor / quad2rotation   utils/general.py:52-100
 *   cam7 [B,7] (quaternion w,x,y,z un-normalised, translation) -> pose [B,4,4] row-major c2w
 * nicer_camera_rays               <- get_camera_params / lift                 utils/rend_util.py:68-93,107-129
 *   uv [B,N,2], pose [B,4,4], K [B,4,4] -> dirs [B,N,3] (divided by the SQUARED norm, rend_util.py:92), cam_loc [B,3]
 *   backward: g_dirs [B,N,3], g_loc [B,3] | NULL -> g_pose [B,4,4] (written; uv and K get no gradient)
 * nicer_ray_points                <- points = cam_loc + z * dir, dirs repeated per sample   model/network.py:112-117
 *   cam_loc [R,3], dirs [R,3], z [R,S] -> points [R*S,3], dirs_flat [R*S,3] | NULL
 *   backward: g_points | NULL, g_dirs_flat | NULL ([R*S,3]) -> g_loc [R,3], g_dirs [R,3] (written; z gets no gradient)
 * nicer_inv4x4                    <- torch.inverse(pose) (w2c of the flow and warp blocks)   model/network.py:157,171
 *   A [B,4,4] -> Ai [B,4,4] (Gauss-Jordan, partial pivoting); backward: Ai, G = dL/dAi -> GA = -Ai^T G Ai^T (written) */
int nicer_pose_from_cam7(const float *cam7, uint32_t B, float *pose, void *stream);
int nicer_pose_from_cam7_backward(const float *cam7, const float *g_pose, uint32_t B, float *g_cam7, void *stream);
int nicer_inv4x4(const float *A, uint32_t B, float *Ai, void *stream);
int nicer_inv4x4_backward(const float *Ai, const float *G, uint32_t B, float *GA, void *stream);
int nicer_camera_rays(const float *uv, const float *pose, const float *K, uint32_t B, uint32_t N, float *dirs,
                      float *cam_loc, void *stream);
int nicer_camera_rays_backward(const float *uv, const float *pose, const float *K, uint32_t B, uint32_t N,
                               const float *g_dirs, const float *g_loc, float *g_pose, void *stream);
int nicer_ray_points(const float *cam_loc, const float *dirs, const float *z, uint32_t R, uint32_t S, float *points,
                     float *dirs_flat, void *stream);
int nicer_ray_points_backward(const float *z, uint32_t R, uint32_t S, const float *g_points, const float *g_dirs_flat,
                              float *g_loc, float *g_dirs, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NICER_B200_H */
