/*
 * snb200.h -- C ABI of libsnb200.so: the B200 (sm_100a) kernels behind SuperNormal's
 * patch-based NeuS training hot path.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; no torch / ATen type crosses this ABI;
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_`;
 *   - every entry point is asynchronous on `stream`, never allocates, never synchronises;
 *   - data-dependent sizes are written to caller-provided device counters;
 *   - return 0 on success, a negative snb_status on argument errors; the message is
 *     available (thread-local) from snb_last_error().
 *
 * Each group cites the reference interface it replaces (paths relative to the reference
 * checkout; NA = third_parties/nerfacc-0.3.5/nerfacc-0.3.5/nerfacc, CS = NA/cuda/csrc).
 */
#ifndef SNB200_H_
#define SNB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *snb_stream_t;

enum snb_status {
    SNB_OK = 0,
    SNB_ERR_NULL = -1,      /* required pointer is null */
    SNB_ERR_ARG = -2,       /* size / enum out of range */
    SNB_ERR_ALIGN = -3,     /* pointer not aligned for vector access */
    SNB_ERR_LAUNCH = -4,    /* cudaGetLastError() after launch */
    SNB_ERR_CAPACITY = -5   /* caller-provided capacity too small (host-checkable cases) */
};

#define SNB_MAX_LEVELS 16
#define SNB_HIDDEN 64 /* d_hidden of models/fields.py (config/diligent.conf:64) */

const char *snb_last_error(void);
int32_t snb_version(void);

/* ---------------------------------------------------------------------------------------
 * Ray marching.  Replaces nerfacc.cuda._C.ray_marching (CS/ray_marching.cu:194-289, bound at
 * CS/pybind.cu:162-206), called from NA/ray_marching.py:177-190.  AABB contraction only.
 * The reference's count pass / cumsum / .item() / write pass becomes three sync-free calls.
 * grid_binary: uint8 (torch.bool) [res_x,res_y,res_z] C-order.  roi: f32[6].
 * ------------------------------------------------------------------------------------- */
int32_t snb_march_count(int32_t n_rays, const float *rays_o, const float *rays_d, const float *t_min,
                        const float *t_max, const float *roi, int32_t res_x, int32_t res_y, int32_t res_z,
                        const uint8_t *grid_binary, float step_size, float cone_angle,
                        int32_t *num_steps, snb_stream_t stream);
/* exclusive scan of num_steps -> packed_info i32[n,2] = (offset,count); total -> *total (device i32) */
int32_t snb_packed_info_from_counts(int32_t n_rays, const int32_t *num_steps, int32_t *packed_info,
                                    int32_t *total, snb_stream_t stream);
/* second round: writes t_starts/t_ends f32[S] and ray ids (i64 and/or i32; either may be null).
 * Samples whose packed offset is >= capacity are dropped (the caller sized the outputs). */
int32_t snb_march_emit(int32_t n_rays, const float *rays_o, const float *rays_d, const float *t_min,
                       const float *t_max, const float *roi, int32_t res_x, int32_t res_y, int32_t res_z,
                       const uint8_t *grid_binary, float step_size, float cone_angle,
                       const int32_t *packed_info, int64_t capacity, int64_t *ray_indices_i64,
                       int32_t *ray_indices_i32, float *t_starts, float *t_ends, snb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Packing / transmittance / rendering weights / accumulation.
 * Replaces NA/pack.py:47-77 (pack_info), _C.transmittance_from_alpha_forward_{cub,naive}
 * (CS/render_transmittance_cub.cu:111-134, CS/render_transmittance.cu:85-112),
 * _C.weight_from_alpha_patch_based_{forward,backward}_naive (CS/render_weight.cu:502-543,
 * :583-626; P==1 gives weight_from_alpha_{forward,backward}_naive :154-221) and the
 * scatter_add_ of NA/vol_rendering.py:311-335 / :176-198.
 * ------------------------------------------------------------------------------------- */
int32_t snb_count_by_ray(int64_t n_samples, const int64_t *ray_indices, int32_t n_rays,
                         int32_t *num_steps /* zero-filled by callee */, snb_stream_t stream);
int32_t snb_max_i64(int64_t n, const int64_t *v, int64_t *out /* device, 1 elem; -1 if n==0 */, snb_stream_t stream);
int32_t snb_transmittance_from_alpha(int32_t n_rays, const int32_t *packed_info, const float *alphas,
                                     float *transmittance, snb_stream_t stream);
int32_t snb_weight_from_alpha_patch_fwd(int32_t n_patches, int32_t patch_size, const int32_t *packed_info,
                                        const float *alphas, float *weights, snb_stream_t stream);
int32_t snb_weight_from_alpha_patch_bwd(int32_t n_patches, int32_t patch_size, const int32_t *packed_info,
                                        const float *alphas, const float *weights, const float *grad_weights,
                                        float *grad_alphas, snb_stream_t stream);
/* out[idx[s], k, :] += w[s,k] * (values ? values[s,k,:] : 1).  out f32[n_out,P,D] zero-filled by callee. */
int32_t snb_accumulate_fwd(int64_t n_samples, int32_t P, int32_t D, const float *weights, const float *values,
                           const int64_t *ray_indices, int32_t n_out, float *out, snb_stream_t stream);
/* grad_w[s,k] = sum_d go[idx,k,d]*(values?values[s,k,d]:1);  grad_v[s,k,d] = go[idx,k,d]*w[s,k] (nullable) */
int32_t snb_accumulate_bwd(int64_t n_samples, int32_t P, int32_t D, const float *weights, const float *values,
                           const int64_t *ray_indices, const float *grad_out, float *grad_weights,
                           float *grad_values, snb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Occupancy grid.  Replaces NA/grid.py:197-239 (_update) + _C.contract_inv
 * (CS/contraction.cu:36-61,89-113; AABB: x*(max-min)+min).
 * ------------------------------------------------------------------------------------- */
/* x[i] = ((coord(idx_i) + rand[i]) / res) * (roi_max-roi_min) + roi_min.  indices null = all cells
 * in order.  rand f32[n,3] (torch.rand_like in the reference, injected by the caller). */
int32_t snb_occgrid_points(int64_t n, const int64_t *indices, const float *rand, int32_t res_x, int32_t res_y,
                           int32_t res_z, const float *roi, float *x, snb_stream_t stream);
/* occs[idx] = max(occs[idx]*decay, occ[i]) (NA/grid.py:232) */
int32_t snb_occgrid_ema(int64_t n, const int64_t *indices, const float *occ, float decay, float *occs,
                        snb_stream_t stream);
/* binary = occs > min(mean(occs), thre) (NA/grid.py:237-239). workspace: >= 8 bytes, device. */
int32_t snb_occgrid_binarize(int64_t num_cells, const float *occs, float occ_thre, uint8_t *binary,
                             void *workspace, snb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Multiresolution hash grid.  Replaces tinycudann.Encoding(HashGrid) as used at
 * models/fields.py:26,78 (tiny-cuda-nn @2ec562e, bindings/torch; kernels kernel_grid,
 * kernel_grid_backward, kernel_grid_backward_input, kernel_grid_backward_input_backward_*).
 * F = 2 features per level.  Tables are level-major, entry-major, feature-minor.
 * ------------------------------------------------------------------------------------- */
typedef struct snb_hashgrid_meta {
    uint32_t n_levels;
    uint32_t offsets[SNB_MAX_LEVELS + 1]; /* in entries */
    float scales[SNB_MAX_LEVELS];
    uint32_t resolutions[SNB_MAX_LEVELS];
} snb_hashgrid_meta;

/* host-only: fills *h_meta; returns total entries (0 on error). */
uint32_t snb_hashgrid_make_meta(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution,
                                float per_level_scale, snb_hashgrid_meta *h_meta);
/* fp32 master params -> fp16 table (what the bindings do with params.to(half) every call) */
int32_t snb_cast_f32_to_f16(int64_t n, const float *src, void *dst_f16, snb_stream_t stream);
/* out: [N, L*2] row-major, fp16 (out_is_f32=0) or fp32 holding the same fp16-rounded values.
 * Levels >= n_active are written as zeros (models/fields.py:81-83 made exact by skipping). */
int32_t snb_hashgrid_fwd(int64_t n, const float *x, const void *table_f16, const snb_hashgrid_meta *h_meta,
                         uint32_t n_active, void *out, int32_t out_is_f32, snb_stream_t stream);
/* table_grad f32[n_entries*2] += sum w * scale * dL_dy   (fp32 atomics; caller zero-fills) */
int32_t snb_hashgrid_bwd_table(int64_t n, const float *x, const void *dL_dy, int32_t dy_is_f32, float dy_scale,
                               const snb_hashgrid_meta *h_meta, uint32_t n_active, float *table_grad,
                               snb_stream_t stream);
/* dL_dx f32[N,3] = sum_levels dL_dy * dy/dx */
int32_t snb_hashgrid_bwd_input(int64_t n, const float *x, const void *dL_dy, int32_t dy_is_f32,
                               const void *table_f16, const snb_hashgrid_meta *h_meta, uint32_t n_active,
                               float *dL_dx, snb_stream_t stream);
/* double backward of bwd_input: given g2 = dL/d(dL_dx) f32[N,3]:
 *   table_grad += d/dtable,  d_dL_dy f32[N,L*2] (nullable) = d/d(dL_dy),  dx2 f32[N,3] (nullable) = d/dx */
int32_t snb_hashgrid_bwd_bwd_input(int64_t n, const float *x, const float *g2, const void *dL_dy,
                                   int32_t dy_is_f32, const void *table_f16, const snb_hashgrid_meta *h_meta,
                                   uint32_t n_active, float *table_grad, float *d_dL_dy, float *dx2,
                                   snb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB200_H_ */
