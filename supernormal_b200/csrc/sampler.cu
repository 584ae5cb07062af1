// sampler.cu -- device-side random patch sampler (models/dataset_loader.py:223-297) and the single-call
// training step that enqueues the whole iteration from C++ (no per-kernel Python dispatch).
#include "sampler.cuh"

namespace snb {

int32_t adam_launch(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *param_f16, int64_t f16_start, float lr,
                    float beta1, float beta2, float eps, int32_t step_count, float grad_scale, snb_stream_t stream);

__global__ void __launch_bounds__(256) sample_patches_kernel(snb_dataset ds, int n_patches, uint64_t seed, uint64_t step,
                                                             snb_batch_out out) {
    sample_patch_ray(ds, n_patches, seed, step, out, blockIdx.x * blockDim.x + threadIdx.x);
}

// ---- fused occupancy update --------------------------------------------------------------------
// hidden units per pass of the sweep's MLP and CTAs per SM it is compiled for.  Measured per update at iterations 8 / 1000 / 4800 (1 / 3 / 14
// live levels; scripts/sweep_time.py): 64 units, 2 CTAs (round 1) 178 / 160 / 339 us; 32, 3: 174 / 149 / 361; 32, 4: 206 / 144 / 376;
// 16, 4: 167 / 134 / 343 (shipped); 16, 5: 168 / 166 / 605 (spills).
#ifndef SNB_OCC_CH
#define SNB_OCC_CH 16
#endif
#ifndef SNB_OCC_MINB
#define SNB_OCC_MINB 4
#endif
// SDF of one point with the MLP in passes of CH hidden units: the features of the live levels wait in the thread's own column of a shared
// tile, so a pass keeps CH accumulators instead of 64 and the kernel fits 3-4 CTAs per SM instead of 2 (the sweep is latency / MIO bound at
// 16 warps per SM: profiles/r02_sass_hist_occupancy_sweep_warmup.txt).  Same arithmetic per hidden unit as sdf_point, same summation order.
template <int CH>
__device__ __forceinline__ float sdf_point_chunked(float x, float y, float z, const __half2 *__restrict__ table, const LevelCtx *s_lvl,
                                                   uint32_t n_active, const float *s_net, float *s_in) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (uint32_t l = 0; l < n_active; ++l) {
        const LevelCtx c = s_lvl[l];
        Cell cell = cell_of(c, x, y, z);
        const float2 f = __half22float2(interp_level(c, cell, table));
        s_in[(2 * l) * nt + tid] = f.x;
        s_in[(2 * l + 1) * nt + tid] = f.y;
    }
    float s = s_net[kOffB1];
#pragma unroll 1
    for (int h0 = 0; h0 < kH; h0 += CH) {
        float acc[CH];
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4 *>(s_net + kOffB0 + h0 + 4 * q);
            acc[4 * q] = t.x; acc[4 * q + 1] = t.y; acc[4 * q + 2] = t.z; acc[4 * q + 3] = t.w;
        }
        const float xin[3] = {x, y, z};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int q = 0; q < CH / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4 *>(s_net + kOffW0T + r * kH + h0 + 4 * q);
                acc[4 * q] = fmaf(w.x, xin[r], acc[4 * q]); acc[4 * q + 1] = fmaf(w.y, xin[r], acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w.z, xin[r], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w.w, xin[r], acc[4 * q + 3]);
            }
        }
        for (uint32_t jf = 0; jf < 2 * n_active; ++jf) {
            const float v = s_in[jf * nt + tid];
            const float *wr = s_net + kOffW0T + (3 + jf) * kH + h0;
#pragma unroll
            for (int q = 0; q < CH / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4 *>(wr + 4 * q);
                acc[4 * q] = fmaf(w.x, v, acc[4 * q]); acc[4 * q + 1] = fmaf(w.y, v, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w.z, v, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w.w, v, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4 *>(s_net + kOffW1 + h0 + 4 * q);
            s = fmaf(t.x, softplus100(acc[4 * q + 0]), s);
            s = fmaf(t.y, softplus100(acc[4 * q + 1]), s);
            s = fmaf(t.z, softplus100(acc[4 * q + 2]), s);
            s = fmaf(t.w, softplus100(acc[4 * q + 3]), s);
        }
    }
    return s;
}

__global__ void __launch_bounds__(256, SNB_OCC_MINB) occgrid_update_kernel(snb_net net, LevelTable lt, int3 res, const float *__restrict__ roi, int warmup,
                                                             float decay, uint64_t seed, uint64_t step, float *__restrict__ occs,
                                                             const float *__restrict__ occs_prev, const uint8_t *__restrict__ binary,
                                                             const unsigned long long *__restrict__ ws) {
    __shared__ __align__(16) float s_net[kNetFloats];
#if SNB_OCC_CH < 64
    __shared__ float s_in[2 * SNB_MAX_LEVELS * 256];
#endif
    load_net_to_smem(s_net, net.net);
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int64_t num_cells = (int64_t)res.x * res.y * res.z;
    const int64_t n_uniform = warmup ? 0 : num_cells / 4;
    const int64_t total = num_cells + n_uniform;
    const uint2 key = make_uint2((uint32_t)seed ^ 0x0cc61d00u, (uint32_t)(seed >> 32));
    // thinning probability of occupied cells so that ~num_cells/4 of them are re-evaluated (NA/grid.py:187-192)
    float thin = 1.f;
    if (!warmup) {
        unsigned long long n_occ = ws[1];
        if (n_occ > (unsigned long long)(num_cells / 4)) thin = (float)(num_cells / 4) / (float)n_occ;
    }
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        uint4 r = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)w, (uint32_t)(w >> 32) ^ 0x77u), key);
        int64_t c;
        if (w < num_cells) {
            c = w;
            if (!warmup) {
                if (!binary[c]) continue;
                if (thin < 1.f) {
                    uint4 r2 = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)w, 0x1234567u), key);
                    if (u01(r2.x) >= thin) continue;
                }
            }
        } else {
            c = (int64_t)(r.w % (uint32_t)num_cells);
        }
        int cz = (int)(c % res.z), cy = (int)((c / res.z) % res.y), cx = (int)(c / ((int64_t)res.z * res.y));
        int cc[3] = {cx, cy, cz}, rr[3] = {res.x, res.y, res.z};
        float rnd[3] = {u01(r.x), u01(r.y), u01(r.z)}, x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float u = __fdiv_rn(__fadd_rn((float)cc[d], rnd[d]), (float)rr[d]);
            float lo = __ldg(roi + d), hi = __ldg(roi + 3 + d);
            x[d] = __fmaf_rn(u, __fsub_rn(hi, lo), lo);
        }
#if SNB_OCC_CH < 64
        float s = sdf_point_chunked<SNB_OCC_CH>(x[0], x[1], x[2], table, s_lvl, net.n_active, s_net, s_in);
#else
        float s = sdf_point<false>(x[0], x[1], x[2], table, s_lvl, net.n_active, s_net, nullptr);
#endif
        occs[c] = fmaxf(__fmul_rn(occs_prev[c], decay), sigmoidf_(-s * 80.f));
    }
}

// the two full passes over the grid values (mean, then threshold) move 16-byte vectors: 2 M floats are 12 us per pass with scalar loads
// from 150 k threads, 7-8 us with float4 / uchar4 (cold L2, under ncu); the tail and unaligned buffers stay scalar
__global__ void __launch_bounds__(256) occgrid_sum2_kernel(int64_t n, const float *__restrict__ occs, double *__restrict__ sum) {
    double s = 0.0;
    const int64_t n4 = (reinterpret_cast<uintptr_t>(occs) & 15) == 0 ? n >> 2 : 0;
    const float4 *o4 = reinterpret_cast<const float4 *>(occs);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(o4 + i);
        s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += occs[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, s);
}
__global__ void __launch_bounds__(256) occgrid_threshold2_kernel(int64_t n, const float *__restrict__ occs, const double *__restrict__ sum, float thre,
                                                                 uint8_t *__restrict__ binary, unsigned long long *__restrict__ count) {
    const float th = fminf((float)(*sum / (double)n), thre);
    unsigned long long c = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(occs) & 15) | (reinterpret_cast<uintptr_t>(binary) & 3)) == 0;
    const int64_t n4 = vec ? n >> 2 : 0;
    const float4 *o4 = reinterpret_cast<const float4 *>(occs);
    uchar4 *b4 = reinterpret_cast<uchar4 *>(binary);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(o4 + i);
        const uchar4 b = make_uchar4(v.x > th ? 1 : 0, v.y > th ? 1 : 0, v.z > th ? 1 : 0, v.w > th ? 1 : 0);
        b4[i] = b;
        c += b.x + b.y + b.z + b.w;
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool b = occs[i] > th;
        binary[i] = b ? 1 : 0;
        c += b;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_sample_patches(const snb_dataset *ds, int32_t n_patches, uint64_t seed, uint64_t step, const snb_batch_out *out,
                                      snb_stream_t stream) {
    SNB_REQUIRE(ds && out, SNB_ERR_NULL, "sample_patches: null struct");
    SNB_REQUIRE(n_patches >= 0 && ds->W > 3 && ds->H > 3 && ds->n_train > 0 && ds->n_images > 0, SNB_ERR_ARG, "sample_patches: bad sizes");
    if (n_patches == 0) return SNB_OK;
    SNB_REQUIRE(ds->normals && ds->masks && ds->intrinsics_inv && ds->pose && ds->train_ids, SNB_ERR_NULL, "sample_patches: null dataset tensor");
    SNB_REQUIRE(out->rays_o && out->rays_d && out->plane_n && out->near_ && out->far_ && out->v_inv && out->normal_gt && out->mask, SNB_ERR_NULL,
                "sample_patches: null output");
    sample_patches_kernel<<<(unsigned)cdiv((int64_t)n_patches * SNB_PATCH, 256), 256, 0, S(stream)>>>(*ds, n_patches, seed, step, *out);
    SNB_LAUNCH_CHECK("sample_patches");
    return SNB_OK;
}

extern "C" int32_t snb_occgrid_update_fused(const snb_net *net, int32_t rx, int32_t ry, int32_t rz, const float *roi, int32_t warmup,
                                            float ema_decay, float occ_thre, uint64_t seed, uint64_t step, float *occs, float *occs_prev,
                                            uint8_t *binary, void *workspace, snb_stream_t stream) {
    SNB_REQUIRE(net && net->table_f16 && net->net, SNB_ERR_NULL, "occgrid_update_fused: null net");
    SNB_REQUIRE(rx > 0 && ry > 0 && rz > 0, SNB_ERR_ARG, "occgrid_update_fused: bad resolution");
    SNB_REQUIRE(roi && occs && occs_prev && binary && workspace, SNB_ERR_NULL, "occgrid_update_fused: null buffer");
    SNB_REQUIRE(aligned(workspace, 8) && aligned(net->net, 16), SNB_ERR_ALIGN, "occgrid_update_fused: misaligned workspace/net");
    const int64_t n = (int64_t)rx * ry * rz;
    cudaMemcpyAsync(occs_prev, occs, sizeof(float) * n, cudaMemcpyDeviceToDevice, S(stream));
    // (a warp-queue + mma.sync evaluator for this sweep was measured in round 2 and dropped: no gain at <= 4 levels, ~20 us per step
    // slower at 14 levels -- profiles/r02_variants_validation.txt)
    occgrid_update_kernel<<<kNumSMs * 8, 256, 0, S(stream)>>>(*net, make_level_table(net->meta), make_int3(rx, ry, rz), roi, warmup, ema_decay, seed, step, occs, occs_prev,
                                                            binary, (const unsigned long long *)workspace);
    cudaMemsetAsync(workspace, 0, 16, S(stream));
    occgrid_sum2_kernel<<<kNumSMs * 2, 256, 0, S(stream)>>>(n, occs, (double *)workspace);
    occgrid_threshold2_kernel<<<kNumSMs * 2, 256, 0, S(stream)>>>(n, occs, (const double *)workspace, occ_thre, binary,
                                                                 (unsigned long long *)workspace + 1);
    SNB_LAUNCH_CHECK("occgrid_update_fused");
    return SNB_OK;
}

extern "C" int32_t snb_train_fwd_bwd(const snb_train_ctx *c, float step_size, float early_stop_eps, float normal_weight, float mask_weight,
                                     float eikonal_weight, snb_stream_t stream) {
    SNB_REQUIRE(c, SNB_ERR_NULL, "train_fwd_bwd: null ctx");
    SNB_REQUIRE(c->flat_param && c->flat_grad && c->net_grad && c->stats && c->sdf && c->feats && c->d_sdf0 && c->d_sdf1 && c->comp && c->wsum &&
                    c->dcomp && c->dwsum, SNB_ERR_NULL, "train_fwd_bwd: null buffer");
    const int n = c->batch.n_patches;
    int32_t rc;
    float *net_w = const_cast<float *>(c->net.net);
    if ((rc = snb_prep_net(c->n_levels, c->flat_param, net_w, n * SNB_PATCH, c->batch.mask, c->stats, stream))) return rc;
    if ((rc = snb_march_visible(&c->batch, &c->net, c->roi, c->res_x, c->res_y, c->res_z, c->grid_binary, step_size, c->jitter, early_stop_eps,
                                &c->samples, stream))) return rc;
    if ((rc = snb_compact_samples(n, &c->samples, stream))) return rc;
    if ((rc = snb_sdf_fwd_patch(&c->batch, &c->net, &c->samples, c->sdf, c->feats, stream))) return rc;
    if ((rc = snb_render_fused(&c->batch, &c->net, &c->samples, c->sdf, normal_weight, mask_weight, eikonal_weight, c->comp, c->wsum, c->d_sdf0,
                               c->d_sdf1, c->stats, stream))) return rc;
    cudaMemsetAsync(c->net_grad, 0, sizeof(float) * SNB_NET_FLOATS, S(stream));
    if ((rc = snb_sdf_bwd_patch_ws(&c->batch, &c->net, &c->samples, c->feats, c->d_sdf0, c->d_sdf1, c->flat_grad + c->small_pad, c->net_grad,
                                   c->bwd_workspace, c->bwd_workspace_bytes, stream)))
        return rc;
    return snb_unfold_grads(c->n_levels, c->flat_param, c->net_grad, c->stats, c->flat_grad, stream);
}

// The lean iteration: `net` already holds the folded weights of the current parameters and net_grad is zero (both left
// behind by snb_train_tail of the previous iteration), so no prep_net / memset / unfold_grads launches.
extern "C" int32_t snb_train_fwd_bwd_lean(const snb_train_ctx *c, float step_size, float early_stop_eps, float normal_weight, float mask_weight,
                                          float eikonal_weight, snb_stream_t stream) {
    SNB_REQUIRE(c, SNB_ERR_NULL, "train_fwd_bwd_lean: null ctx");
    SNB_REQUIRE(c->flat_param && c->flat_grad && c->net_grad && c->stats && c->sdf && c->feats && c->d_sdf0 && c->d_sdf1 && c->comp && c->wsum,
                SNB_ERR_NULL, "train_fwd_bwd_lean: null buffer");
    const int n = c->batch.n_patches;
    int32_t rc;
    if ((rc = snb_march_visible(&c->batch, &c->net, c->roi, c->res_x, c->res_y, c->res_z, c->grid_binary, step_size, c->jitter, early_stop_eps,
                                &c->samples, stream))) return rc;
    if ((rc = snb_compact_samples_stats(n, &c->samples, n * SNB_PATCH, c->batch.mask, c->stats, stream))) return rc;
    if ((rc = snb_sdf_fwd_patch(&c->batch, &c->net, &c->samples, c->sdf, c->feats, stream))) return rc;
    if ((rc = snb_render_fused(&c->batch, &c->net, &c->samples, c->sdf, normal_weight, mask_weight, eikonal_weight, c->comp, c->wsum, c->d_sdf0,
                               c->d_sdf1, c->stats, stream))) return rc;
    return snb_sdf_bwd_patch_ws(&c->batch, &c->net, &c->samples, c->feats, c->d_sdf0, c->d_sdf1, c->flat_grad + c->small_pad, c->net_grad,
                                c->bwd_workspace, c->bwd_workspace_bytes, stream);
}

extern "C" int32_t snb_train_fwd_bwd_lean_ad(const snb_train_ctx *c, float step_size, float early_stop_eps, float normal_weight, float mask_weight,
                                             float eikonal_weight, float *grad, float *d_grad, snb_stream_t stream) {
    SNB_REQUIRE(c, SNB_ERR_NULL, "train_fwd_bwd_lean_ad: null ctx");
    SNB_REQUIRE(c->flat_param && c->flat_grad && c->net_grad && c->stats && c->sdf && c->feats && c->d_sdf0 && c->d_sdf1 && c->comp && c->wsum && grad && d_grad,
                SNB_ERR_NULL, "train_fwd_bwd_lean_ad: null buffer");
    const int n = c->batch.n_patches;
    int32_t rc;
    if ((rc = snb_march_visible(&c->batch, &c->net, c->roi, c->res_x, c->res_y, c->res_z, c->grid_binary, step_size, c->jitter, early_stop_eps,
                                &c->samples, stream))) return rc;
    if ((rc = snb_compact_samples_stats(n, &c->samples, n * SNB_PATCH, c->batch.mask, c->stats, stream))) return rc;
    if ((rc = snb_sdf_fwd_patch(&c->batch, &c->net, &c->samples, c->sdf, c->feats, stream))) return rc;
    if ((rc = snb_sdf_grad_patch(&c->batch, &c->net, &c->samples, grad, stream))) return rc;
    if ((rc = snb_render_fused_ad(&c->batch, &c->net, &c->samples, c->sdf, grad, normal_weight, mask_weight, eikonal_weight, c->comp, c->wsum,
                                  c->d_sdf0, c->d_sdf1, d_grad, c->stats, stream))) return rc;
    if ((rc = snb_sdf_bwd_patch_ws(&c->batch, &c->net, &c->samples, c->feats, c->d_sdf0, c->d_sdf1, c->flat_grad + c->small_pad, c->net_grad,
                                   c->bwd_workspace, c->bwd_workspace_bytes, stream))) return rc;
    return snb_sdf_grad_bwd_patch(&c->batch, &c->net, &c->samples, c->feats, d_grad, c->flat_grad + c->small_pad, c->net_grad, stream);
}

extern "C" int32_t snb_train_optim(const snb_train_ctx *c, float lr, int32_t step_count, float grad_scale, snb_stream_t stream) {
    SNB_REQUIRE(c, SNB_ERR_NULL, "train_optim: null ctx");
    SNB_REQUIRE(c->flat_param && c->flat_grad && c->exp_avg && c->exp_avg_sq, SNB_ERR_NULL, "train_optim: null buffer");
    // one sweep over [MLP block | live levels of the table]; levels >= n_active have exactly zero gradient and Adam
    // state, so skipping them is exact (no weight decay)
    const int64_t n_live = 2 * (int64_t)c->net.meta.offsets[c->net.n_active];
    return adam_launch(c->small_pad + n_live, c->flat_param, c->flat_grad, c->exp_avg, c->exp_avg_sq, const_cast<void *>(c->net.table_f16),
                       c->small_pad, lr, 0.9f, 0.999f, 1e-8f, step_count, grad_scale, stream);
}
