/*
 * oracle.c -- CPU restatement of the reference kernels on SuperNormal's patch-based
 * NeuS hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product (supernormal_b200/)
 * never links or calls this file.
 *
 * Every function cites the reference source it restates (paths relative to
 * /root/reference; NA = third_parties/nerfacc-0.3.5/nerfacc-0.3.5/nerfacc,
 * CS = NA/cuda/csrc).  Floating-point contraction follows what nvcc 12.9 emits for the
 * reference sources at -O3 for sm_100a (re-checked from the SASS of
 * oracle/_ref): each fused multiply-add below is an explicit fmaf(), everything else is a
 * separately rounded fp32 operation, so this file must be compiled with
 * -ffp-contract=off.
 *
 * Parity pinning:
 *   - ray marching, patch weights fwd/bwd, serial transmittance: pinned by the reference's
 *     golden vectors (NT/test_rendering.py, NT/test_pack.py, docstrings) in
 *     tests/test_oracle_golden.py and, on the GPU box, bit-for-bit against the reference
 *     kernels themselves (oracle/_ref, tests/test_ref_differential.py).
 *   - hash grid: tiny-cuda-nn @2ec562e (create_env.sh:11) is NOT vendored in the reference
 *     tree and cannot be installed here (no network): PARITY UNPINNED.  The functions
 *     below restate its published algorithm (include/tiny-cuda-nn/encodings/grid.h,
 *     common_device.h) as summarised in SURVEY.md Appendix A.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * Ray marching: CS/ray_marching.cu:9-192 (+ CS/include/helpers_contraction.h:16-21,
 * CS/include/helpers_math.h:177-180,1167-1170,1217-1220,1343-1346,1471-1475)
 * ---------------------------------------------------------------------------------- */

static inline float clampf_(float f, float a, float b) { return fmaxf(a, fminf(f, b)); } /* helpers_math.h:1167 */
static inline float signf_(float x) { return copysignf(1.0f, x); }                    /* helpers_math.h:1471 */
static inline int f2i_trunc(float x) { /* F2I.TRUNC: NaN -> 0, saturating */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (int)0x80000000;
    return (int)x;
}
static inline int clampi_(int v, int a, int b) { return v < a ? a : (v > b ? b : v); }

/* CS/ray_marching.cu:9-14 */
static inline float calc_dt(float t, float cone, float dt_min, float dt_max) {
    return clampf_(t * cone, dt_min, dt_max);
}

/* CS/ray_marching.cu:16-45 (AABB contraction only) */
static inline int grid_occupied_at(const float xyz[3], const float roi[6], const int res[3],
                                   const uint8_t *grid) {
    if (xyz[0] < roi[0] || xyz[0] > roi[3] || xyz[1] < roi[1] || xyz[1] > roi[4] ||
        xyz[2] < roi[2] || xyz[2] > roi[5])
        return 0;
    int ix[3];
    for (int a = 0; a < 3; ++a) {
        float u = (xyz[a] - roi[a]) / (roi[3 + a] - roi[a]); /* roi_to_unit */
        ix[a] = clampi_(f2i_trunc(u * (float)res[a]), 0, res[a] - 1);
    }
    int idx = (ix[0] * res[1] + ix[1]) * res[2] + ix[2];
    return grid[idx] != 0;
}

/* CS/ray_marching.cu:48-57 */
static inline float distance_to_next_voxel(const float xyz[3], const float dir[3],
                                           const float inv_dir[3], const float roi[6],
                                           const int res[3]) {
    float t3[3];
    for (int a = 0; a < 3; ++a) {
        float r = (float)res[a];
        float ext = roi[3 + a] - roi[a];
        float u = (xyz[a] - roi[a]) / ext;
        float fl = floorf(fmaf(signf_(dir[a]), 0.5f, fmaf(r, u, 0.5f)));
        float diff = fmaf(r, -u, fl);
        t3[a] = ((diff * inv_dir[a]) / r) * ext;
    }
    float t = fminf(fminf(t3[0], t3[1]), t3[2]);
    return fmaxf(t, 0.0f);
}

/* CS/ray_marching.cu:59-75 */
static inline float advance_to_next_voxel(float t, float dt_min, const float xyz[3],
                                          const float dir[3], const float inv_dir[3],
                                          const float roi[6], const int res[3], float far) {
    float t_target = t + distance_to_next_voxel(xyz, dir, inv_dir, roi, res);
    t_target = fminf(t_target, far);
    float _t = t;
    do {
        _t += dt_min;
    } while (_t < t_target);
    return _t;
}

/* One ray of CS/ray_marching.cu:81-192.  Returns the number of samples; writes them when
 * t0s/t1s are non-null (the reference's second round). */
static int march_one(const float *o, const float *d, float near, float far, const float roi[6],
                     const int res[3], const uint8_t *grid, float step, float cone, float *t0s,
                     float *t1s) {
    float inv_dir[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    const float dt_min = step, dt_max = 1e10f;
    int j = 0;
    float t0 = near;
    float dt = calc_dt(t0, cone, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    while (t_mid < far) {
        float xyz[3] = {fmaf(t_mid, d[0], o[0]), fmaf(t_mid, d[1], o[1]), fmaf(t_mid, d[2], o[2])};
        if (grid_occupied_at(xyz, roi, res, grid)) {
            if (t0s) {
                t0s[j] = t0;
                t1s[j] = t1;
            }
            ++j;
            t0 = t1;
            t1 = t0 + calc_dt(t0, cone, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        } else {
            t_mid = advance_to_next_voxel(t_mid, dt_min, xyz, d, inv_dir, roi, res, far);
            dt = calc_dt(t_mid, cone, dt_min, dt_max);
            t0 = t_mid - dt * 0.5f;
            t1 = t_mid + dt * 0.5f;
        }
    }
    return j;
}

/* Host wrapper of CS/ray_marching.cu:194-289: count pass, exclusive cumsum, write pass.
 * packed_info int32 [n,2]; returns total samples.  Call with t0s==NULL to size outputs. */
int64_t oracle_ray_marching(int n_rays, const float *rays_o, const float *rays_d,
                            const float *t_min, const float *t_max, const float *roi,
                            const int *res, const uint8_t *grid, float step, float cone,
                            int32_t *packed_info, int64_t *ray_indices, float *t0s, float *t1s) {
    int64_t total = 0;
    for (int i = 0; i < n_rays; ++i) {
        int base = (int)total;
        int c = march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], roi, res, grid, step,
                          cone, t0s ? t0s + base : NULL, t1s ? t1s + base : NULL);
        packed_info[2 * i] = base;
        packed_info[2 * i + 1] = c;
        if (ray_indices)
            for (int j = 0; j < c; ++j) ray_indices[base + j] = i;
        total += c;
    }
    return total;
}

/* ------------------------------------------------------------------------------------
 * Patch-based rendering weights: CS/render_weight.cu:87-118 (fwd), :298-340 (bwd).
 * patch_size==1 gives the per-ray kernels :154-221.  Contraction per SURVEY Appendix C.
 * ---------------------------------------------------------------------------------- */
void oracle_weight_from_alpha_patch_fwd(int n_patches, int P, const int32_t *packed_info,
                                        const float *alphas, float *weights) {
    for (int i = 0; i < n_patches; ++i) {
        int base = packed_info[2 * i], steps = packed_info[2 * i + 1];
        for (int k = 0; k < P; ++k) {
            float T = 1.f;
            for (int j = 0; j < steps; ++j) {
                size_t id = (size_t)(base + j) * P + k;
                float a = alphas[id];
                weights[id] = a * T;
                T = T * (1.f - a);
            }
        }
    }
}

void oracle_weight_from_alpha_patch_bwd(int n_patches, int P, const int32_t *packed_info,
                                        const float *alphas, const float *weights,
                                        const float *grad_weights, float *grad_alphas) {
    for (int i = 0; i < n_patches; ++i) {
        int base = packed_info[2 * i], steps = packed_info[2 * i + 1];
        for (int k = 0; k < P; ++k) {
            float accum = 0.f;
            for (int j = 0; j < steps; ++j) {
                size_t id = (size_t)(base + j) * P + k;
                accum = fmaf(grad_weights[id], weights[id], accum);
            }
            float T = 1.f;
            for (int j = 0; j < steps; ++j) {
                size_t id = (size_t)(base + j) * P + k;
                float a = alphas[id];
                float num = fmaf(grad_weights[id], T, -accum);
                grad_alphas[id] = num / fmaxf(1.f - a, 1e-10f);
                accum = fmaf(-grad_weights[id], weights[id], accum);
                T = T * (1.f - a);
            }
        }
    }
}

/* Serial exclusive transmittance per ray: CS/render_transmittance.cu:85-112 (the naive
 * kernel; the CUB route CS/render_transmittance_cub.cu:111-134 multiplies in a tree order
 * and is only comparable through the visibility mask). */
void oracle_transmittance_from_alpha(int n_rays, const int32_t *packed_info, const float *alphas,
                                     float *trans) {
    for (int i = 0; i < n_rays; ++i) {
        int base = packed_info[2 * i], steps = packed_info[2 * i + 1];
        float T = 1.f;
        for (int j = 0; j < steps; ++j) {
            trans[base + j] = T;
            T = T * (1.f - alphas[base + j]);
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Multiresolution hash grid (tiny-cuda-nn, NOT in the reference tree -- parity unpinned).
 * Call sites: models/fields.py:26 (construct), :78 (forward).
 * ---------------------------------------------------------------------------------- */

/* fp16 emulation (round-to-nearest-even, subnormals kept) -- numerically what
 * __float2half_rn / __hfma do.  No reliance on compiler _Float16 support. */
static uint16_t f64_to_h(double x) {
    if (x != x) return 0x7e00;
    uint16_t sign = signbit(x) ? 0x8000 : 0;
    double a = fabs(x);
    if (a >= 65520.0) return sign | 0x7c00; /* rounds to inf */
    if (a < 6.103515625e-05) {              /* subnormal: multiples of 2^-24 */
        double q = a * 16777216.0;          /* exact scaling */
        double r = nearbyint(q);            /* RNE under default rounding mode */
        return sign | (uint16_t)r;          /* r==1024 correctly yields the min normal */
    }
    int e;
    double m = frexp(a, &e); /* a = m*2^e, m in [0.5,1) */
    double q = m * 2048.0;   /* 11 significant bits: [1024, 2048) */
    double r = nearbyint(q);
    if (r >= 2048.0) {
        r = 1024.0;
        e += 1;
    }
    int be = e - 1 + 15;
    if (be >= 31) return sign | 0x7c00;
    return sign | (uint16_t)((be << 10) | ((int)r - 1024));
}
static double h_to_f64(uint16_t h) {
    int s = h >> 15, e = (h >> 10) & 31, m = h & 1023;
    double v;
    if (e == 0)
        v = ldexp((double)m, -24);
    else if (e == 31)
        v = m ? NAN : INFINITY;
    else
        v = ldexp((double)(m + 1024), e - 25);
    return s ? -v : v;
}
uint16_t oracle_f32_to_h(float x) { return f64_to_h((double)x); }
float oracle_h_to_f32(uint16_t h) { return (float)h_to_f64(h); }
/* __hfma: single rounding of a*b+c.  a*b is exact in double (22 significant bits); the sum
 * is exact in double unless the exponents are >30 binades apart, where the smaller term is
 * below half an fp16 ulp of the larger and sticky-correct anyway except on exact ties. */
static uint16_t hfma_(uint16_t a, uint16_t b, uint16_t c) {
    return f64_to_h(h_to_f64(a) * h_to_f64(b) + h_to_f64(c));
}

/* grid_scale / grid_resolution (tcnn common_device.h) */
float oracle_grid_scale(uint32_t level, float log2_per_level_scale, uint32_t base_resolution) {
    return exp2f((float)level * log2_per_level_scale) * (float)base_resolution - 1.0f;
}
uint32_t oracle_grid_resolution(float scale) { return (uint32_t)ceilf(scale) + 1u; }

/* Offset table: GridEncodingTemplated ctor (tcnn encodings/grid.h).  offsets has L+1 entries
 * (in table entries, not floats). Returns total entries. */
uint32_t oracle_hashgrid_offsets(uint32_t n_levels, uint32_t log2_hashmap_size,
                                 uint32_t base_resolution, float per_level_scale,
                                 uint32_t *offsets, float *scales, uint32_t *resolutions) {
    uint32_t offset = 0;
    float l2 = log2f(per_level_scale);
    for (uint32_t i = 0; i < n_levels; ++i) {
        float scale = oracle_grid_scale(i, l2, base_resolution);
        uint32_t res = oracle_grid_resolution(scale);
        uint32_t max_params = UINT32_MAX / 2;
        uint32_t n = powf((float)res, 3.0f) > (float)max_params ? max_params : res * res * res;
        n = (n + 7u) / 8u * 8u;
        uint32_t cap = 1u << log2_hashmap_size;
        if (n > cap) n = cap;
        offsets[i] = offset;
        if (scales) scales[i] = scale;
        if (resolutions) resolutions[i] = res;
        offset += n;
    }
    offsets[n_levels] = offset;
    return offset;
}

static inline uint32_t grid_index(uint32_t hashmap_size, uint32_t res, const uint32_t p[3]) {
    uint32_t stride = 1, index = 0;
    for (int dim = 0; dim < 3 && stride <= hashmap_size; ++dim) {
        index += p[dim] * stride;
        stride *= res;
    }
    if (hashmap_size < stride) index = p[0] ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
    return index % hashmap_size;
}

/* kernel_grid forward for F=2 (tcnn encodings/grid.h).  table_h: fp16 bits [entries,2];
 * out_h: fp16 bits, row-major [N, L*2] (the layout models/fields.py:78 receives).
 * Levels >= n_active are written as zero (== the mask of models/fields.py:81-83).
 * dy_dx (optional, fp32 [N, L*2, 3]) follows the same kernel's gradient branch. */
void oracle_hashgrid_fwd(int64_t n, const float *x, const uint16_t *table_h, uint32_t n_levels,
                         const uint32_t *offsets, const float *scales, const uint32_t *resolutions,
                         uint32_t n_active, uint16_t *out_h, float *dy_dx) {
    const uint32_t C = n_levels * 2;
    for (int64_t i = 0; i < n; ++i) {
        for (uint32_t l = 0; l < n_levels; ++l) {
            uint16_t *o = out_h + i * C + 2 * l;
            float *dd = dy_dx ? dy_dx + (i * C + 2 * l) * 3 : NULL;
            if (l >= n_active) {
                o[0] = o[1] = 0;
                if (dd) memset(dd, 0, 6 * sizeof(float));
                continue;
            }
            const uint16_t *g = table_h + (size_t)offsets[l] * 2;
            uint32_t size = offsets[l + 1] - offsets[l];
            float scale = scales[l];
            uint32_t res = resolutions[l];
            float pos[3];
            uint32_t pg[3];
            for (int d = 0; d < 3; ++d) { /* pos_fract */
                float p = fmaf(scale, x[3 * i + d], 0.5f);
                float fl = floorf(p);
                pg[d] = (uint32_t)(int)fl;
                pos[d] = p - fl;
            }
            uint16_t r0 = 0, r1 = 0;
            for (uint32_t c = 0; c < 8; ++c) {
                float w = 1.f;
                uint32_t pl[3];
                for (int d = 0; d < 3; ++d) {
                    if ((c & (1u << d)) == 0) {
                        w *= 1.f - pos[d];
                        pl[d] = pg[d];
                    } else {
                        w *= pos[d];
                        pl[d] = pg[d] + 1;
                    }
                }
                uint32_t idx = grid_index(size, res, pl) * 2;
                uint16_t wh = f64_to_h((double)w);
                r0 = hfma_(wh, g[idx], r0);
                r1 = hfma_(wh, g[idx + 1], r1);
            }
            o[0] = r0;
            o[1] = r1;
            if (dd) {
                float grads[2][3] = {{0, 0, 0}, {0, 0, 0}};
                for (int gd = 0; gd < 3; ++gd) {
                    for (uint32_t c = 0; c < 4; ++c) {
                        float w = scale;
                        uint32_t pl[3];
                        for (int nd = 0; nd < 2; ++nd) {
                            int d = nd >= gd ? nd + 1 : nd;
                            if ((c & (1u << nd)) == 0) {
                                w *= 1.f - pos[d];
                                pl[d] = pg[d];
                            } else {
                                w *= pos[d];
                                pl[d] = pg[d] + 1;
                            }
                        }
                        pl[gd] = pg[gd];
                        uint32_t il = grid_index(size, res, pl) * 2;
                        pl[gd] = pg[gd] + 1;
                        uint32_t ir = grid_index(size, res, pl) * 2;
                        for (int f = 0; f < 2; ++f) {
                            float vr = (float)h_to_f64(g[ir + f]), vl = (float)h_to_f64(g[il + f]);
                            /* weight * (right - left) * pos_derivative(=1), separately rounded */
                            grads[f][gd] += w * (vr - vl) * 1.0f;
                        }
                    }
                }
                for (int f = 0; f < 2; ++f)
                    for (int gd = 0; gd < 3; ++gd) dd[f * 3 + gd] = grads[f][gd];
            }
        }
    }
}

/* kernel_grid_backward (table gradient), accumulated in double so that it can serve as the
 * "exact" value the fp32-atomic CUDA kernel and tcnn's fp16-atomic kernel both approximate.
 * dL_dy: fp32 row-major [N, L*2]; grad: double [entries,2] (caller zero-fills). */
void oracle_hashgrid_bwd_table(int64_t n, const float *x, const float *dL_dy, uint32_t n_levels,
                               const uint32_t *offsets, const float *scales,
                               const uint32_t *resolutions, uint32_t n_active, double *grad) {
    const uint32_t C = n_levels * 2;
    for (int64_t i = 0; i < n; ++i) {
        for (uint32_t l = 0; l < n_levels && l < n_active; ++l) {
            double *g = grad + (size_t)offsets[l] * 2;
            uint32_t size = offsets[l + 1] - offsets[l];
            float scale = scales[l];
            uint32_t res = resolutions[l];
            float pos[3];
            uint32_t pg[3];
            for (int d = 0; d < 3; ++d) {
                float p = fmaf(scale, x[3 * i + d], 0.5f);
                float fl = floorf(p);
                pg[d] = (uint32_t)(int)fl;
                pos[d] = p - fl;
            }
            for (uint32_t c = 0; c < 8; ++c) {
                float w = 1.f;
                uint32_t pl[3];
                for (int d = 0; d < 3; ++d) {
                    if ((c & (1u << d)) == 0) {
                        w *= 1.f - pos[d];
                        pl[d] = pg[d];
                    } else {
                        w *= pos[d];
                        pl[d] = pg[d] + 1;
                    }
                }
                uint32_t idx = grid_index(size, res, pl) * 2;
                g[idx] += (double)w * (double)dL_dy[i * C + 2 * l];
                g[idx + 1] += (double)w * (double)dL_dy[i * C + 2 * l + 1];
            }
        }
    }
}

/* Indices + trilinear weights of the 8 corners (for building differentiable PyTorch
 * oracles without re-deriving the hash in Python).  idx: int64 [N, L, 8] ABSOLUTE entry
 * index (offset included); w: fp32 [N, L, 8]; frac/cell optional. */
void oracle_hashgrid_corners(int64_t n, const float *x, uint32_t n_levels, const uint32_t *offsets,
                             const float *scales, const uint32_t *resolutions, int64_t *idx,
                             float *w) {
    for (int64_t i = 0; i < n; ++i)
        for (uint32_t l = 0; l < n_levels; ++l) {
            uint32_t size = offsets[l + 1] - offsets[l];
            float pos[3];
            uint32_t pg[3];
            for (int d = 0; d < 3; ++d) {
                float p = fmaf(scales[l], x[3 * i + d], 0.5f);
                float fl = floorf(p);
                pg[d] = (uint32_t)(int)fl;
                pos[d] = p - fl;
            }
            for (uint32_t c = 0; c < 8; ++c) {
                float ww = 1.f;
                uint32_t pl[3];
                for (int d = 0; d < 3; ++d) {
                    if ((c & (1u << d)) == 0) {
                        ww *= 1.f - pos[d];
                        pl[d] = pg[d];
                    } else {
                        ww *= pos[d];
                        pl[d] = pg[d] + 1;
                    }
                }
                idx[(i * n_levels + l) * 8 + c] = (int64_t)offsets[l] + grid_index(size, resolutions[l], pl);
                w[(i * n_levels + l) * 8 + c] = ww;
            }
        }
}
