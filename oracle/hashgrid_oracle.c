/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path
 * (nicer_slam_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Plain-C, single-threaded (optionally OpenMP over points for the read-only
 * passes) restatement of the reference's multi-resolution hash/dense grid
 * encoder, following /root/reference/code/hashencoder/src/hashencoder.cu:
 *
 *   grid_index()        <- get_grid_index<D,C>   hashencoder.cu:54-73
 *   hash_coords()       <- fast_hash<D>          hashencoder.cu:35-51
 *   sstep()/sstep_d()   <- smoothstep(_derivative) hashencoder.cu:115-121
 *   oracle_hash_forward <- kernel_grid           hashencoder.cu:131-283  (K1)
 *   oracle_hash_backward<- kernel_grid_backward  hashencoder.cu:286-373  (K2)
 *                          kernel_input_backward hashencoder.cu:376-402  (K3)
 *   oracle_hash_second_backward
 *                       <- kernel_grid_second_backward_grad      :405-458 (K4)
 *                          kernel_grid_second_backward_embedding :461-625 (K5)
 *
 * Layouts are the reference's: inputs [B,D] in [0,1]; grid [N,C]; offsets
 * int32 [L+1]; outputs / grad / grad_grad [L,B,C]; dy_dx [B,L,D,C];
 * accumulators (grad_grid, grad_inputs, grad2_grid) must arrive zeroed.
 *
 * Parity status: the reference ships no golden vectors for this path
 * ("parity unpinned by the reference"); this file is pinned
 * instead against the reference's own Python autograd wiring run on top of it
 * (oracle/gen_golden.py) and against an independent float64 numpy evaluation
 * of the same formulas (tests/test_oracle_hash.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXD 3
#define MAXC 8

static inline float sstep(float v) { return v * v * (3.0f - 2.0f * v); }
static inline float sstep_d(float v) { return 6 * v * (1.0f - v); }

static inline uint32_t hash_coords(uint32_t D, const uint32_t *p) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                       2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

/* hashencoder.cu:54-73: dense index while stride <= hashmap_size, stride *= resolution
 * (NOT resolution+1), hashed otherwise; always wrapped by % hashmap_size. */
static inline uint32_t grid_index(uint32_t D, uint32_t C, uint32_t ch, uint32_t hashmap_size,
                                  uint32_t resolution, const uint32_t *p) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += p[d] * stride;
        stride *= resolution;
    }
    if (stride > hashmap_size) index = hash_coords(D, p);
    return (index % hashmap_size) * C + ch;
}

typedef struct {
    uint32_t hashmap_size, resolution;
    float scale;
    float pos[MAXD], dpos[MAXD];
    uint32_t pg[MAXD];
} cell_t;

/* returns 0 when the point is out of [0,1]^D (reference: zeros / early return) */
static inline int locate(cell_t *c, const float *x, const int32_t *offsets, uint32_t level,
                         uint32_t D, float S, uint32_t H) {
    for (uint32_t d = 0; d < D; d++)
        if (x[d] < 0 || x[d] > 1) return 0;
    c->hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    c->scale = exp2f((float)level * S) * (float)H - 1.0f;
    c->resolution = (uint32_t)ceilf(c->scale) + 1;
    for (uint32_t d = 0; d < D; d++) {
        float p = x[d] * c->scale;
        c->pg[d] = (uint32_t)floorf(p);
        p -= (float)c->pg[d];
        c->dpos[d] = sstep_d(p);
        c->pos[d] = sstep(p);
    }
    return 1;
}

void oracle_hash_forward(const float *inputs, const float *grid, const int32_t *offsets,
                         float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                         uint32_t H, int calc_grad_inputs, float *dy_dx) {
#pragma omp parallel for schedule(static)
    for (int64_t bb = 0; bb < (int64_t)B; bb++) {
        uint32_t b = (uint32_t)bb;
        for (uint32_t level = 0; level < L; level++) {
            const float *g = grid + (size_t)(uint32_t)offsets[level] * C;
            const float *x = inputs + (size_t)b * D;
            float *out = outputs + ((size_t)level * B + b) * C;
            float *dd = calc_grad_inputs ? dy_dx + ((size_t)b * L + level) * D * C : 0;
            cell_t c;
            if (!locate(&c, x, offsets, level, D, S, H)) {
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
                if (dd) memset(dd, 0, sizeof(float) * D * C);
                continue;
            }
            float res[MAXC] = {0};
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pl[MAXD];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - c.pos[d]; pl[d] = c.pg[d]; }
                    else                         { w *= c.pos[d];     pl[d] = c.pg[d] + 1; }
                }
                uint32_t index = grid_index(D, C, 0, c.hashmap_size, c.resolution, pl);
                for (uint32_t ch = 0; ch < C; ch++) res[ch] += w * g[index + ch];
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = res[ch];
            if (!dd) continue;
            for (uint32_t gd = 0; gd < D; gd++) {
                float rg[MAXC] = {0};
                for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                    float w = c.scale;
                    uint32_t pl[MAXD];
                    for (uint32_t nd = 0; nd < D - 1; nd++) {
                        uint32_t d = (nd >= gd) ? nd + 1 : nd;
                        if ((idx & (1u << nd)) == 0) { w *= 1 - c.pos[d]; pl[d] = c.pg[d]; }
                        else                          { w *= c.pos[d];     pl[d] = c.pg[d] + 1; }
                    }
                    pl[gd] = c.pg[gd];
                    uint32_t il = grid_index(D, C, 0, c.hashmap_size, c.resolution, pl);
                    pl[gd] = c.pg[gd] + 1;
                    uint32_t ir = grid_index(D, C, 0, c.hashmap_size, c.resolution, pl);
                    for (uint32_t ch = 0; ch < C; ch++)
                        rg[ch] += w * (g[ir + ch] - g[il + ch]) * c.dpos[gd];
                }
                for (uint32_t ch = 0; ch < C; ch++) dd[gd * C + ch] = rg[ch];
            }
        }
    }
}

/* K2 (scatter) + K3 (input gradient). Serial over points: deterministic accumulation order. */
void oracle_hash_backward(const float *grad, const float *inputs, const float *grid,
                          const int32_t *offsets, float *grad_grid, uint32_t B, uint32_t D,
                          uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                          const float *dy_dx, float *grad_inputs) {
    (void)grid;
    for (uint32_t level = 0; level < L; level++) {
        float *gg = grad_grid + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            const float *gr = grad + ((size_t)level * B + b) * C;
            cell_t c;
            if (!locate(&c, x, offsets, level, D, S, H)) continue;
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pl[MAXD];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - c.pos[d]; pl[d] = c.pg[d]; }
                    else                         { w *= c.pos[d];     pl[d] = c.pg[d] + 1; }
                }
                uint32_t index = grid_index(D, C, 0, c.hashmap_size, c.resolution, pl);
                for (uint32_t ch = 0; ch < C; ch++) gg[index + ch] += w * gr[ch];
            }
        }
    }
    if (calc_grad_inputs) {
        for (uint32_t b = 0; b < B; b++) {
            const float *dd = dy_dx + (size_t)b * L * D * C;
            for (uint32_t d = 0; d < D; d++) {
                float r = 0;
                for (uint32_t l = 0; l < L; l++)
                    for (uint32_t ch = 0; ch < C; ch++)
                        r += grad[((size_t)l * B + b) * C + ch] * dd[(l * D + d) * C + ch];
                grad_inputs[(size_t)b * D + d] = r;
            }
        }
    }
}

/* K4 (grad_grad = ggx . dy_dx) + K5 (second-order scatter into the grid).
 * NOTE (hashencoder.cu:405-458): K4 has no out-of-range test of its own; it relies on dy_dx
 * being zero for OOB points.  K5 early-returns for OOB points.  There is no d/dx output
 * (hashgrid.py:134 returns None for inputs): second derivatives w.r.t. x are dropped. */
void oracle_hash_second_backward(const float *grad, const float *inputs, const float *grid,
                                 const int32_t *offsets, uint32_t B, uint32_t D, uint32_t C,
                                 uint32_t L, float S, uint32_t H, const float *dy_dx,
                                 const float *ggx, float *grad_grad, float *grad2_grid) {
    (void)grid;
    for (uint32_t level = 0; level < L; level++) {
        for (uint32_t b = 0; b < B; b++) {
            const float *dd = dy_dx + ((size_t)b * L + level) * D * C;
            float *o = grad_grad + ((size_t)level * B + b) * C;
            for (uint32_t ch = 0; ch < C; ch++) {
                float r = 0;
                for (uint32_t d = 0; d < D; d++) r += ggx[(size_t)b * D + d] * dd[d * C + ch];
                o[ch] = r;
            }
        }
    }
    for (uint32_t level = 0; level < L; level++) {
        float *g2 = grad2_grid + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            const float *gr = grad + ((size_t)level * B + b) * C;
            const float *gx = ggx + (size_t)b * D;
            cell_t c;
            if (!locate(&c, x, offsets, level, D, S, H)) continue;
            float cache[(1 << MAXD) * MAXC];
            memset(cache, 0, sizeof(cache));
            for (uint32_t gd = 0; gd < D; gd++) {
                for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                    float w = c.scale;
                    uint32_t loc[MAXD];
                    for (uint32_t nd = 0; nd < D - 1; nd++) {
                        uint32_t d = (nd >= gd) ? nd + 1 : nd;
                        if ((idx & (1u << nd)) == 0) { w *= 1 - c.pos[d]; loc[d] = 0; }
                        else                          { w *= c.pos[d];     loc[d] = 1; }
                    }
                    uint32_t il = 0, ir = 0;
                    loc[gd] = 0;
                    for (uint32_t d = 0; d < D; d++) il += loc[d] << d;
                    loc[gd] = 1;
                    for (uint32_t d = 0; d < D; d++) ir += loc[d] << d;
                    for (uint32_t ch = 0; ch < C; ch++) {
                        float v = w * gr[ch] * gx[gd] * c.dpos[gd];
                        cache[ir * C + ch] += v;
                        cache[il * C + ch] -= v;
                    }
                }
            }
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                uint32_t pl[MAXD];
                for (uint32_t d = 0; d < D; d++) pl[d] = c.pg[d] + ((idx >> d) & 1u);
                uint32_t index = grid_index(D, C, 0, c.hashmap_size, c.resolution, pl);
                for (uint32_t ch = 0; ch < C; ch++) g2[index + ch] += cache[idx * C + ch];
            }
        }
    }
}
