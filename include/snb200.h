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
/* cudaMemcpyAsync on a caller-chosen stream (kind: 1 host->device, 2 device->host, 3 device->device): the host-fed loop's batch upload
 * and loss read-back (what Dataset.gen_random_patches(...).cuda() and loss.item() are in exp_runner.py:156-165, 216) without a
 * framework in between. */
int32_t snb_copy_async(void *dst, const void *src, int64_t bytes, int32_t kind, snb_stream_t stream);

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


/* =======================================================================================
 * Fused training path: the same operators as above, fused so that one
 * training step is ~10 launches with no host synchronisation.  Replaces, per step, the op
 * sequence of models/renderer.py:63-276 + models/fields.py:76-99 + exp_runner.py:191-207.
 * All struct members are DEVICE pointers (or plain sizes).
 * ===================================================================================== */
#define SNB_NET_FLOATS 2432 /* folded MLP: W0T[35][64] | b0[64] | W1[64] | b1 | inv_s | pad */
#define SNB_PATCH 9         /* rays per patch (3x3, config/diligent.conf:29) */

typedef struct snb_net {
    const void *table_f16;  /* [n_entries*2] fp16 hash table */
    const float *net;       /* [SNB_NET_FLOATS] folded weights written by snb_prep_net */
    snb_hashgrid_meta meta;
    uint32_t n_active;      /* SDFNetwork.bindwidth (models/fields.py:73-83) */
} snb_net;

typedef struct snb_patch_batch { /* outputs of Dataset.gen_random_patches, models/dataset_loader.py:223-277 */
    int32_t n_patches;
    const float *rays_o;    /* [N,3]   camera centre of the patch's view */
    const float *rays_d;    /* [N,9,3] */
    const float *plane_n;   /* [N,3]   marching-plane normal */
    const float *near_;     /* [N] */
    const float *far_;      /* [N]   (NaN for rays missing the unit sphere) */
    const float *v_inv;     /* [N,9,3,3] */
    const float *normal_gt; /* [N,9,3] */
    const float *mask;      /* [N,9] */
} snb_patch_batch;

typedef struct snb_samples { /* visible samples on the patch-centre rays, packed by patch */
    int64_t capacity;       /* max samples S */
    int64_t end_capacity;   /* max non-contiguous interval ends */
    int32_t scratch_stride; /* per-ray capacity of the marching scratch */
    int32_t *counts;        /* [N]   samples per ray */
    int32_t *end_counts;    /* [N]   interval ends per ray that need their own SDF query */
    int32_t *packed_info;   /* [N,2] (offset,count) */
    int32_t *end_packed;    /* [N,2] */
    int32_t *totals;        /* [4]   S, S_end, overflow flag, 0 */
    float *t0;              /* [capacity] */
    float *t1;              /* [capacity] */
    int32_t *patch_idx;     /* [capacity] */
    int32_t *end_slot;      /* [capacity] slot of the sample's own end query, or -1 (use next sample's start) */
    int32_t *slot_sample;   /* [end_capacity] sample owning each end slot */
    float *scratch_t0;      /* [N, scratch_stride] */
    float *scratch_t1;      /* [N, scratch_stride] */
    int32_t *launch_order;  /* [N] or NULL: patch indices, most sample chunks first; written by snb_compact_samples(_stats), read by
                             * snb_render_fused(_ad), whose CTA i then takes patch launch_order[i] (the kernel ends with its longest patch) */
} snb_samples;

/* Fold weight_norm (models/fields.py:66-67) and the variance network (models/fields.py:133-139,
 * models/renderer.py:171) into the `net` buffer; also reduces mask_sum (exp_runner.py:169-174).
 * small layout: v0[64*d_in] | g0[64] | b0[64] | v1[64] | g1 | b1 | variance,  d_in = 3 + 2*n_levels.
 * stats (device f32[8], zeroed here): [0]=mask_sum(+1e-5) [1]=normal sq-err sum [2]=bce sum
 * [3]=eikonal sum [4]=d(inv_s) */
int32_t snb_prep_net(int32_t n_levels, const float *small, float *net, int32_t n_mask, const float *mask,
                     float *stats, snb_stream_t stream);
/* Ray marching + NeuS visibility cut fused (NA/ray_marching.py:157-220 with
 * models/renderer.py:80-122 as alpha_fn): warp per centre ray, stops at T < early_stop_eps.
 * jitter: device f32[N] in [0,1) or null (stratified=False). */
int32_t snb_march_visible(const snb_patch_batch *h_batch, const snb_net *h_net, const float *roi, int32_t res_x,
                          int32_t res_y, int32_t res_z, const uint8_t *grid_binary, float step_size,
                          const float *jitter, float early_stop_eps, const snb_samples *h_samples,
                          snb_stream_t stream);
/* scan counts -> packed_info/totals, then copy scratch -> packed arrays, assign end slots */
int32_t snb_compact_samples(int32_t n_patches, const snb_samples *h_samples, snb_stream_t stream);
/* Same, and the scan CTA also resets the loss accumulators of the iteration: stats[0] = #(mask > 0.5) + 1e-5 (the
 * normal-loss normaliser, exp_runner.py:184-186), stats[1..7] = 0 -- what snb_prep_net does when given a mask. */
int32_t snb_compact_samples_stats(int32_t n_patches, const snb_samples *h_samples, int32_t n_mask, const float *mask,
                                  float *stats, snb_stream_t stream);
/* SDF at arbitrary points, no grad.  mode 0: sdf, 1: sigmoid(-80*sdf) (models/renderer.py:56-60), 2: -sdf */
int32_t snb_sdf_eval(int64_t n, const float *x, const snb_net *h_net, int32_t mode, float *out, snb_stream_t stream);
/* self-test of the tcgen05 / TMEM plumbing the fused kernels build on: D[128,N] = A[128,K] * B[N,K]^T, kind::tf32 (operands are
 * rounded to TF32 by the tensor core), fp32 accumulate.  K % 8 == 0, K <= 128, N in {32, 64}.  err (device i32) is set to 1 if the
 * MMA never signalled completion. */
int32_t snb_umma_selftest(const float *A, const float *B, float *D, int32_t K, int32_t N, int32_t *err, snb_stream_t stream);
/* same, with the A tile in the padded K-major layout (leading-dimension byte offset lbo_a = 128 or 192, csrc/umma.cuh) */
int32_t snb_umma_selftest_lbo(const float *A, const float *B, float *D, int32_t K, int32_t N, int32_t lbo_a, int32_t *err,
                              snb_stream_t stream);
/* SDF and its analytic gradient d sdf / d x in ONE pass (forward-mode through encode + MLP on the tensor cores):
 * what SDFNetwork.gradient (models/fields.py:107-119) returns for `ad` normals (models/renderer.py:225-226, :345),
 * without the autograd double pass.  sdf: f32[n] or null; grad: f32[n,3]. */
int32_t snb_sdf_eval_grad(int64_t n, const float *x, const snb_net *h_net, float *sdf, float *grad, snb_stream_t stream);
/* SDF at the 9 plane-projected rays of every sample start (+ own ends), keeping the encoded features.
 * sdf: [9*(capacity+end_capacity)]: starts at s*9+k, ends at 9*S + slot*9+k.  feats: half2 [points, 16] capacity; a row holds the live
 * levels, row stride 1/2/4/8/16 half2 = n_active rounded up to a power of two (kept for snb_sdf_bwd_patch of the SAME n_active) */
int32_t snb_sdf_fwd_patch(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                          float *sdf, void *feats, snb_stream_t stream);
/* NeuS alpha -> patch transmittance scan -> dfd normals -> accumulation (models/renderer.py:164-267).
 * comp [N,9,3], wsum [N,9]; optional per-sample outputs gradients [S,9,3], weights [S,9]. stats[3] += eikonal sum */
int32_t snb_render_fwd(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                       const float *sdf, float *comp, float *wsum, float *gradients, float *weights, float *stats,
                       snb_stream_t stream);
/* losses of exp_runner.py:191-203 (l2) and their seeds: dcomp [N,9,3], dwsum [N,9]; stats[1], stats[2] */
int32_t snb_patch_loss(const snb_patch_batch *h_batch, const float *comp, const float *wsum, float normal_weight,
                       float mask_weight, float *stats, float *dcomp, float *dwsum, snb_stream_t stream);
/* backward of snb_render_fwd: d_sdf0/d_sdf1 [S,9] (w.r.t. start / end SDF of each interval), stats[4] += d inv_s.
 * dgrad [S,9,3] optional external seed on the per-sample gradients; eikonal_weight adds the fused
 * eikonal term eikonal_weight * mean((|g|-1)^2). */
int32_t snb_render_bwd(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                       const float *sdf, const float *comp, const float *wsum, const float *dcomp, const float *dwsum,
                       const float *dgrad, float eikonal_weight, float *d_sdf0, float *d_sdf1, float *stats,
                       snb_stream_t stream);
/* snb_render_fwd + snb_patch_loss + snb_render_bwd in ONE launch (lane = sample, warp scans instead of serial chains).
 * Same results up to fp32 summation order.  d_sdf0/d_sdf1 both null: forward + losses only (stats[1..3], comp, wsum). */
int32_t snb_render_fused(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                         const float *sdf, float normal_weight, float mask_weight, float eikonal_weight, float *comp,
                         float *wsum, float *d_sdf0, float *d_sdf1, float *stats, snb_stream_t stream);
/* MLP backward + hash-table scatter for all points of snb_sdf_fwd_patch.
 * table_grad f32[n_entries*2] += ...;  net_grad f32[SNB_NET_FLOATS] += gradients w.r.t. the FOLDED weights */
int32_t snb_sdf_bwd_patch(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                          const void *feats, const float *d_sdf0, const float *d_sdf1, float *table_grad,
                          float *net_grad, snb_stream_t stream);
/* The same backward with a caller-provided device workspace.  With more than 4 active levels and a workspace of at least
 * snb_sdf_bwd_workspace_bytes(...) the work is split into two launches: (1) the tcgen05 MLP backward writes d loss / d features
 * ([level][ray][sample] float2) and the point positions into the workspace, (2) a high-occupancy scatter kernel adds them to
 * table_grad with the trilinear weights, merging contributions to the same grid cell in shared memory first and issuing the
 * reductions corner-major so that x-neighbour corners share one L2 sector request (profiles/r02_red_rate_microbench.txt: the L2
 * retires ~97 reduction SECTORS per clock whatever their width).  workspace == NULL (or too small, or <= 4 active levels): the
 * single fused kernel of snb_sdf_bwd_patch.  Same results up to fp32 summation order. */
int64_t snb_sdf_bwd_workspace_bytes(int32_t n_levels, int64_t capacity, int64_t end_capacity);
int32_t snb_sdf_bwd_patch_ws(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                             const void *feats, const float *d_sdf0, const float *d_sdf1, float *table_grad,
                             float *net_grad, void *workspace, int64_t workspace_bytes, snb_stream_t stream);
/* gradient_method = 'ad' (models/renderer.py:225-226; SDFNetwork.gradient with create_graph=True, models/fields.py:107-119) inside the fused
 * step, without autograd:
 *   snb_sdf_grad_patch      analytic d sdf / d x at every sample start of every in-patch ray (point p = 9 s + k), f32 [9 * capacity, 3]
 *   snb_render_fused_ad     snb_render_fused with those gradients as the per-sample normals; d_grad = d loss / d gradient (same layout),
 *                           d_sdf0 / d_sdf1 = the alpha path only
 *   snb_sdf_grad_bwd_patch  parameter gradients of d_grad . grad_x sdf -- the double backward through MLP and hash grid (table: value-weight
 *                           term through sigmoid', derivative-weight term through the tangent) -- accumulated like snb_sdf_bwd_patch
 *   snb_train_fwd_bwd_lean_ad  the lean iteration with these three in place of the dfd render / backward */
int32_t snb_sdf_grad_patch(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples, float *grad,
                           snb_stream_t stream);
int32_t snb_render_fused_ad(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples, const float *sdf,
                            const float *grad_in, float normal_weight, float mask_weight, float eikonal_weight, float *comp,
                            float *wsum, float *d_sdf0, float *d_sdf1, float *d_grad, float *stats, snb_stream_t stream);
int32_t snb_sdf_grad_bwd_patch(const snb_patch_batch *h_batch, const snb_net *h_net, const snb_samples *h_samples,
                               const void *feats, const float *d_grad, float *table_grad, float *net_grad, snb_stream_t stream);
/* Un-fold net_grad into gradients of (v,g,b,variance) (weight_norm / exp backward), in `small` layout. */
int32_t snb_unfold_grads(int32_t n_levels, const float *small, const float *net_grad, const float *stats,
                         float *small_grad, snb_stream_t stream);
/* Adam (torch.optim.Adam defaults, exp_runner.py:97,207) over n params; zeroes grad; optionally refreshes the
 * fp16 copy.  step_count >= 1.  grad_scale multiplies the gradient first (1/world_size after an allreduce). */
int32_t snb_adam_step(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *param_f16,
                      float lr, float beta1, float beta2, float eps, int32_t step_count, float grad_scale,
                      snb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Device-side patch sampler and single-call training step: removes the
 * per-iteration host work of Dataset.gen_random_patches / near_far_from_sphere
 * (models/dataset_loader.py:223-297: ~30 ATen launches + a host np.random.choice + H2D) and of
 * the per-kernel Python dispatch.
 * ------------------------------------------------------------------------------------- */
typedef struct snb_dataset { /* tensors Dataset.__init__ leaves on the device, models/dataset_loader.py:99-150 */
    int32_t n_images, H, W, n_train;
    const float *normals;        /* [n_images,H,W,3] world-space normals */
    const float *masks;          /* [n_images,H,W] */
    const float *intrinsics_inv; /* [n_images,4,4] */
    const float *pose;           /* [n_images,4,4] camera-to-world */
    const float *v_inverse;      /* [n_images,H,W,3,3] (Dataset.V_inverse_all), or NULL: computed in closed form from pose and
                                  * the pixel's camera-frame direction (no 36 B/pixel table; models/dataset_loader.py:114-137) */
    const int32_t *train_ids;    /* [n_train] view ids to sample from (exclude_views removed) */
} snb_dataset;

typedef struct snb_batch_out { /* writable twin of snb_patch_batch + stratified jitter */
    float *rays_o, *rays_d, *plane_n, *near_, *far_, *v_inv, *normal_gt, *mask, *jitter;
} snb_batch_out;

/* Fills a batch of n_patches random 3x3 patches: centre pixel uniform in [1,W-3]x[1,H-3]
 * (torch.randint(1, W-2), dataset_loader.py:244-245), view uniform over train_ids (:252), rays through the
 * pixel centres (:262-267), near/far from the unit sphere (:279-297, NaN when missed), jitter ~ U[0,1).
 * Counter-based RNG (Philox4x32-10) keyed by (seed, step): reproducible, no state. */
int32_t snb_sample_patches(const snb_dataset *h_ds, int32_t n_patches, uint64_t seed, uint64_t step,
                           const snb_batch_out *h_out, snb_stream_t stream);

/* Fused occupancy update (NA/grid.py:197-239 with models/renderer.py:56-60 as occ_eval_fn), sync-free:
 * warmup!=0: every cell; else every occupied cell (thinned to ~num_cells/4 if more are occupied) + num_cells/4
 * uniform random cells.  occs_prev: scratch f32[num_cells]; workspace: >= 16 bytes. */
int32_t snb_occgrid_update_fused(const snb_net *h_net, int32_t res_x, int32_t res_y, int32_t res_z, const float *roi,
                                 int32_t warmup, float ema_decay, float occ_thre, uint64_t seed, uint64_t step,
                                 float *occs, float *occs_prev, uint8_t *binary, void *workspace, snb_stream_t stream);

typedef struct snb_train_ctx { /* everything one training iteration touches; all device pointers */
    snb_patch_batch batch;
    snb_samples samples;
    snb_net net;              /* net.net = folded weights buffer, net.table_f16 = fp16 table */
    int32_t n_levels;
    int64_t small_pad;        /* floats in front of the hash table inside the flat buffers */
    float *flat_param;        /* [small_pad + n_entries*2]  small | pad | table */
    float *flat_grad, *exp_avg, *exp_avg_sq;
    float *net_grad;          /* [SNB_NET_FLOATS] */
    float *sdf;               /* [9*(capacity+end_capacity)] */
    void *feats;              /* half2 [9*(capacity+end_capacity), 16]: kept features, rows of 1/2/4/8/16 half2 (the live levels rounded up) */
    float *d_sdf0, *d_sdf1;   /* [9*capacity] */
    float *comp, *wsum, *dcomp, *dwsum; /* [N,9,3] [N,9] */
    float *stats;             /* [8] */
    const float *jitter;      /* [N] or null */
    const float *roi;         /* [6] */
    const uint8_t *grid_binary;
    int32_t res_x, res_y, res_z;
    void *bwd_workspace;      /* snb_sdf_bwd_patch_ws workspace (may be null: single-kernel backward) */
    int64_t bwd_workspace_bytes;
} snb_train_ctx;

/* prep_net -> march_visible -> compact -> sdf_fwd_patch -> render_fwd -> patch_loss -> render_bwd ->
 * sdf_bwd_patch -> unfold_grads, all enqueued on `stream` by one host call (exp_runner.py:177-206). */
int32_t snb_train_fwd_bwd(const snb_train_ctx *h_ctx, float step_size, float early_stop_eps, float normal_weight,
                          float mask_weight, float eikonal_weight, snb_stream_t stream);
/* Adam over the MLP/variance parameters and the active levels of the table (exp_runner.py:207) */
int32_t snb_train_optim(const snb_train_ctx *h_ctx, float lr, int32_t step_count, float grad_scale, snb_stream_t stream);

/* The lean iteration (what FusedTrainer runs by default).  Precondition: net.net holds the folded weights of the current
 * parameters and net_grad is zero -- both are left behind by snb_train_tail (or by snb_prep_net + a memset).
 * march_visible -> compact_samples_stats -> sdf_fwd_patch -> render_fused -> sdf_bwd_patch; the gradient w.r.t. the
 * FOLDED MLP weights stays in net_grad, the table gradient in flat_grad + small_pad. */
int32_t snb_train_fwd_bwd_lean(const snb_train_ctx *h_ctx, float step_size, float early_stop_eps, float normal_weight,
                               float mask_weight, float eikonal_weight, snb_stream_t stream);
/* the same iteration with analytic normals (gradient_method = 'ad'): ... -> sdf_fwd_patch -> sdf_grad_patch -> render_fused_ad ->
 * sdf_bwd_patch (alpha path) -> sdf_grad_bwd_patch (normal / eikonal path).  grad, d_grad: f32 [9 * capacity, 3] scratch. */
int32_t snb_train_fwd_bwd_lean_ad(const snb_train_ctx *h_ctx, float step_size, float early_stop_eps, float normal_weight,
                                  float mask_weight, float eikonal_weight, float *grad, float *d_grad, snb_stream_t stream);
/* Everything between the backward of one iteration and the marcher of the next in ONE launch (exp_runner.py:205-207 +
 * the next iteration's models/fields.py:66-67 weight_norm and dataset_loader.py:223-297 gen_random_patches):
 *   block 0: weight-norm backward net_grad -> (v, g, b, variance) gradients (skipped if grads_unfolded != 0: flat_grad
 *            already holds them, e.g. after snb_unfold_grads + a data-parallel allreduce), Adam on that block, fold of
 *            the updated parameters into net.net, net_grad zeroed;
 *   next blocks: snb_sample_patches for (seed, next_step) into h_out_next (skipped if h_ds_next is null);
 *   rest: Adam + gradient zeroing + fp16 refresh over the live table levels.
 * Results are bit-identical to snb_unfold_grads -> snb_train_optim -> snb_prep_net -> snb_sample_patches. */
int32_t snb_train_tail(const snb_train_ctx *h_ctx, float lr, int32_t step_count, float grad_scale, int32_t grads_unfolded,
                       const snb_dataset *h_ds_next, int32_t n_patches_next, uint64_t seed, uint64_t next_step,
                       const snb_batch_out *h_out_next, snb_stream_t stream);

/* Data-parallel step tail over NVLink peer memory (one process per GPU, one box): replaces
 * snb_unfold_grads -> NCCL allreduce of the live gradient range -> snb_train_tail (exp_runner.py:205-207 under DDP-style
 * data parallelism) by ONE kernel per rank.  All ranks' flat parameters / gradients / fp16 tables live in a symmetric,
 * peer-mapped allocation (supernormal_b200/dp.py: PeerGroup).  Table chunks of 1024 float4 are owned round-robin: the owner
 * sums the chunk's gradients over all ranks with peer loads, runs Adam on its shard of the optimizer state, stores the new
 * parameters (+ fp16 copy) into every rank's buffers and zeroes every rank's gradient chunk; the MLP / variance block is
 * reduced in rank order and updated identically on every rank; cross-GPU ordering by two flag barriers inside the kernel.
 * Preconditions: h_ctx->flat_param / flat_grad / net.table_f16 are this rank's entries of the peer group and
 * h_ctx->net_grad == h_ctx->flat_grad (the backward accumulates the folded-weight gradient into flat_grad[0:2432)).
 * flags[r]: >= 17 zero-initialised u32 in rank r's symmetric buffer; counter: local zero-initialised u32.
 * flags[rank][16] != 0 afterwards: a wait timed out (20 s) -- a peer never arrived. */
#define SNB_MAX_PEERS 8
typedef struct snb_peer_group {
    int32_t world, rank;
    float *param[SNB_MAX_PEERS];
    float *grad[SNB_MAX_PEERS];
    void *table_f16[SNB_MAX_PEERS];
    uint32_t *flags[SNB_MAX_PEERS];
    uint32_t *counter;
    int32_t table_f16_only; /* 0: owners broadcast fp32 parameters + fp16 copy and zero the peers' gradient chunks.  1 (what
                             * FusedTrainer uses; validated on 2 GPUs, profiles/r02_dp_peer_check_n2_f16only1.json): owners
                             * broadcast the fp16 copy only (fp32 master parameters stay valid on the owner's chunks alone) and
                             * every rank zeroes its own gradient range after the done barrier. */
    uint32_t epoch;         /* value the two in-kernel barriers wait for: must grow strictly from launch to launch of one peer
                             * group, whatever the optimizer step count does (a resume rewinds step_count, not this).  0: use
                             * step_count (a run that never rewinds). */
    uint64_t *trace;        /* optional (may be NULL): device buffer u64[trace_capacity][4]; block 0 of launch `epoch` stores %globaltimer
                             * (ns, this GPU's clock) at: kernel start | every rank's "backward complete" flag seen | local chunk
                             * blocks done (reduce + Adam + broadcast fenced) | every rank's "stores done" flag seen -- the in-kernel
                             * timeline of the two cross-GPU barriers (scripts/dp_peer_trace.py) */
    int32_t trace_capacity;
} snb_peer_group;
int32_t snb_train_tail_peer(const snb_train_ctx *h_ctx, const snb_peer_group *h_peers, float lr, int32_t step_count,
                            const snb_dataset *h_ds_next, int32_t n_patches_next, uint64_t seed, uint64_t next_step,
                            const snb_batch_out *h_out_next, snb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Mesh extraction.  Replaces extract_fields / extract_geometry (models/renderer.py:9-34): the 64^3-chunked SDF query
 * with a host round trip per chunk, and PyMCubes' CPU mcubes.marching_cubes (create_env.sh:14, models/renderer.py:29).
 * A slab is the lattice planes [x0, x0+nx) x ny x nz of the res^3 grid; its last plane is the halo shared with the
 * next slab.  Vertex numbering: [y/z-edge vertices of planes 0..nx-2 | x-edge vertices | y/z-edge vertices of plane
 * nx-1], lattice order; vertices are in lattice-index coordinates (x + x_offset), the caller applies
 * v/(res-1)*(max-min)+min (models/renderer.py:33).  Winding: normals point from u > iso to u <= iso.
 * ------------------------------------------------------------------------------------- */
/* out[i,j,k] = f(sdf(xs[i], ys[j], zs[k]));  mode as in snb_sdf_eval (2: -sdf, what extract_geometry queries) */
int32_t snb_sdf_grid_query(const float *xs, int32_t nx, const float *ys, int32_t ny, const float *zs, int32_t nz,
                           const snb_net *h_net, int32_t mode, float *out, snb_stream_t stream);
/* host-only: bytes of device workspace snb_mc_count / snb_mc_emit need for an nx*ny*nz slab (0 if any dim < 2) */
int64_t snb_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz);
/* classify + scans.  The first 32 bytes of the workspace then hold int64 {n_vertices, n_main, n_triangles, 0}
 * (n_main = vertices numbered before the last plane's block).  workspace: 256-byte aligned. */
int32_t snb_mc_count(const float *u, int32_t nx, int32_t ny, int32_t nz, float iso, void *workspace, snb_stream_t stream);
/* vertices f32[v_cap,3], triangles i32[t_cap,3] (ids local to the slab); entries beyond the capacities are dropped */
int32_t snb_mc_emit(const float *u, int32_t nx, int32_t ny, int32_t nz, float iso, float x_offset, const void *workspace,
                    int64_t v_cap, int64_t t_cap, float *vertices, int32_t *triangles, snb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB200_H_ */
