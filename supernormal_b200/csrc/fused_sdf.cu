// fused_sdf.cu -- SDF queries for the fused training step:
//   sdf_eval        points -> sdf (no grad): occupancy-grid update, mesh extraction, validation
//   sdf_fwd_patch   all 9 plane-projected rays of every visible sample (+ own interval ends): positions
//                   are built in-kernel (models/renderer.py:146-159), encode + MLP fused, features kept
//   sdf_bwd_patch   MLP backward (recomputing layer 0 from the kept features), weight gradients as a
//                   shared-memory-tiled contraction, hash-table scatter with vector fp32 atomics
#include "sdf_mma.cuh"

namespace snb {

constexpr int kFwdWarps = 4;   // warps per CTA of the tensor-core forward kernels
constexpr size_t kFwdSmemBytes = sizeof(float) * kMmaSmemFloats(kFwdWarps);

struct FwdSmem {
    float *net, *whi, *wlo, *xs;
};
__device__ __forceinline__ FwdSmem fwd_smem_setup(float *smem, const float *g_net) {
    FwdSmem s;
    s.net = smem;
    s.whi = s.net + kNetFloats;
    s.wlo = s.whi + kWRows * kWStride;
    s.xs = s.wlo + kWRows * kWStride + (threadIdx.x >> 5) * 32 * kXsStride;
    load_net_to_smem(s.net, g_net);
    stage_w_split(s.net, s.whi, s.wlo);
    return s;
}

__global__ void __launch_bounds__(32 * kFwdWarps, 4) sdf_eval_kernel(int64_t n, const float *__restrict__ x, snb_net net, LevelTable lt, int mode,
                                                                     float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const FwdSmem sh = fwd_smem_setup(smem, net.net);
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const bool valid = i < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        float s = warp_sdf_mma<false, true>(valid, px, py, pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs, nullptr, lane);
        if (valid) out[i] = mode == 1 ? sigmoidf_(-s * 80.f) : (mode == 2 ? -s : s);
    }
}

// sdf and d sdf / d x in one pass (analytic normals, models/fields.py:107-119 without the autograd graph)
constexpr size_t kGradSmemBytes = kFwdSmemBytes + sizeof(float) * kFwdWarps * 3 * 32 * kTsStride;

__global__ void __launch_bounds__(32 * kFwdWarps, 2) sdf_eval_grad_kernel(int64_t n, const float *__restrict__ x, snb_net net, LevelTable lt,
                                                                          float *__restrict__ sdf, float *__restrict__ grad) {
    extern __shared__ __align__(16) float smem[];
    const FwdSmem sh = fwd_smem_setup(smem, net.net);
    float *ts = smem + kMmaSmemFloats(kFwdWarps) + (threadIdx.x >> 5) * 3 * 32 * kTsStride;
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const bool valid = i < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        float g[3];
        float s = warp_sdf_grad_mma<true>(valid, px, py, pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs, ts, lane, g);
        if (valid) {
            if (sdf) sdf[i] = s;
            grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2];
        }
    }
}

struct PointRef {
    int s, k, patch;
    bool is_end;
    float px, py, pz;
};

// point index -> (sample, in-patch ray) and its world position (models/renderer.py:146-156)
__device__ __forceinline__ PointRef decode_point(int64_t p, int S, const snb_patch_batch &b, const snb_samples &sm) {
    PointRef r;
    int64_t q = p / SNB_PATCH;
    r.k = (int)(p - q * SNB_PATCH);
    float t;
    if (q < S) {
        r.s = (int)q;
        r.is_end = false;
        t = __ldg(sm.t0 + r.s);
    } else {
        r.s = __ldg(sm.slot_sample + (q - S));
        r.is_end = true;
        t = __ldg(sm.t1 + r.s);
    }
    r.patch = __ldg(sm.patch_idx + r.s);
    const float *o = b.rays_o + 3 * (int64_t)r.patch;
    const float *n = b.plane_n + 3 * (int64_t)r.patch;
    const float *dk = b.rays_d + ((int64_t)r.patch * SNB_PATCH + r.k) * 3;
    const float *dc = b.rays_d + ((int64_t)r.patch * SNB_PATCH + SNB_PATCH / 2) * 3;
    float nx = __ldg(n), ny = __ldg(n + 1), nz = __ldg(n + 2);
    float dkx = __ldg(dk), dky = __ldg(dk + 1), dkz = __ldg(dk + 2);
    float num = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(dc), nx), __fmul_rn(__ldg(dc + 1), ny)), __fmul_rn(__ldg(dc + 2), nz));
    float den = __fadd_rn(__fadd_rn(__fmul_rn(dkx, nx), __fmul_rn(dky, ny)), __fmul_rn(dkz, nz));
    float tk = __fdiv_rn(__fmul_rn(t, num), den);
    r.px = __fadd_rn(__ldg(o), __fmul_rn(dkx, tk));
    r.py = __fadd_rn(__ldg(o + 1), __fmul_rn(dky, tk));
    r.pz = __fadd_rn(__ldg(o + 2), __fmul_rn(dkz, tk));
    return r;
}

__global__ void __launch_bounds__(32 * kFwdWarps, 4) sdf_fwd_patch_kernel(snb_patch_batch b, snb_net net, LevelTable lt, snb_samples sm,
                                                                          float *__restrict__ sdf, __half2 *__restrict__ feats) {
    extern __shared__ __align__(16) float smem[];
    const FwdSmem sh = fwd_smem_setup(smem, net.net);
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int S = sm.totals[0], E = sm.totals[1];
    const int64_t M = (int64_t)SNB_PATCH * (S + E);
    const uint32_t L = net.meta.n_levels;
    const int lane = threadIdx.x & 31;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < M; p0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + threadIdx.x;
        const bool valid = p < M;
        PointRef r;
        r.px = r.py = r.pz = 0.f;
        if (valid) r = decode_point(p, S, b, sm);
        float s = warp_sdf_mma<true, true>(valid, r.px, r.py, r.pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs,
                                           feats + (valid ? p : 0) * L, lane);
        if (valid) sdf[p] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 128;     // points per CTA iteration == threads per CTA
constexpr int kDzStride = 68;  // floats; 16B-aligned rows, conflict-free 128-bit stores per quarter-warp
constexpr int kXStride = 36;   // constant 1 (bias row) + x_in (<=35)

// dW0T accumulation of one 64-point group: thread owns h in [a4,a4+4) and the NI live columns ib, ib+4, ...
template <int NI>
__device__ __forceinline__ void phase_b(float (&accW)[9][4], const float *__restrict__ xg, const float *__restrict__ dg) {
#pragma unroll 4
    for (int pp = 0; pp < 64; ++pp) {
        float4 d = *reinterpret_cast<const float4 *>(dg + pp * kDzStride);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            float xv = xg[pp * kXStride + 4 * i];
            accW[i][0] = fmaf(xv, d.x, accW[i][0]);
            accW[i][1] = fmaf(xv, d.y, accW[i][1]);
            accW[i][2] = fmaf(xv, d.z, accW[i][2]);
            accW[i][3] = fmaf(xv, d.w, accW[i][3]);
        }
    }
}

// v[64] per lane -> lane l ends with the warp sums of v[2l], v[2l+1] in v[0], v[1]
__device__ __forceinline__ void warp_transpose_reduce64(float (&v)[kH], int lane) {
#pragma unroll
    for (int off = 16, half = 32; off >= 1; off >>= 1, half >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            float send = up ? v[i] : v[i + half];
            float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

__global__ void __launch_bounds__(kTile, 3) sdf_bwd_patch_kernel(snb_patch_batch b, snb_net net, LevelTable lt, snb_samples sm,
                                                                 const __half2 *__restrict__ feats,
                                                                 const float *__restrict__ d_sdf0,
                                                                 const float *__restrict__ d_sdf1,
                                                                 float *__restrict__ table_grad, float *__restrict__ net_grad) {
    extern __shared__ __align__(16) float smem[];
    float *s_net = smem;                      // kNetFloats
    float *s_dz = s_net + kNetFloats;         // kTile * kDzStride
    float *s_x = s_dz + kTile * kDzStride;    // kTile * kXStride
    load_net_to_smem(s_net, net.net);
    const LevelCtx *s_lvl = lt.lv;
    const int tid = threadIdx.x, lane = tid & 31;
    const int S = sm.totals[0], E = sm.totals[1];
    const int64_t M = (int64_t)SNB_PATCH * (S + E);
    const uint32_t L = net.meta.n_levels, n_active = net.n_active;
    const int K = 4 + 2 * (int)n_active;  // live columns of the staged input: [1 | x y z | features of the active levels]

    // phase-B ownership: group g (64 threads) reduces points [64g, 64g+64); thread owns h in [4a,4a+4) and the live
    // columns ib, ib+4, ib+8, ... (interleaved so that early training, with few active levels, still spreads evenly)
    const int grp = tid >> 6, a4 = (tid & 15) * 4, ib = (tid >> 4) & 3;
    const int n_i = (K - ib + 3) / 4;
    float accW[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int h = 0; h < 4; ++h) accW[i][h] = 0.f;
    float accW1a = 0.f, accW1b = 0.f, accB1 = 0.f;

    for (int64_t tile0 = (int64_t)blockIdx.x * kTile; tile0 < M; tile0 += (int64_t)gridDim.x * kTile) {
        const int64_t p = tile0 + tid;
        const bool valid = p < M;
        float dz[kH];
        float dsdf = 0.f;
        PointRef r;
        r.px = r.py = r.pz = 0.f;
        if (valid) {
            r = decode_point(p, S, b, sm);
            if (!r.is_end) {
                dsdf = __ldg(d_sdf0 + (int64_t)r.s * SNB_PATCH + r.k);
                // this start also served as the previous interval's end when that interval had no own end query
                if (r.s > 0 && __ldg(sm.end_slot + r.s - 1) < 0) dsdf += __ldg(d_sdf1 + (int64_t)(r.s - 1) * SNB_PATCH + r.k);
            } else {
                dsdf = __ldg(d_sdf1 + (int64_t)r.s * SNB_PATCH + r.k);
            }
        }
        // ---- phase A: recompute layer 0, dz = dsdf * W1 * softplus'(z); stage dz and x_in in shared memory
        float *xrow = s_x + tid * kXStride;
        if (valid) {
            layer0<false, true>(r.px, r.py, r.pz, nullptr, s_lvl, n_active, s_net, const_cast<__half2 *>(feats + p * L), dz);
            xrow[0] = 1.f; xrow[1] = r.px; xrow[2] = r.py; xrow[3] = r.pz;
            for (uint32_t l = 0; l < n_active; ++l) {
                float2 f = __half22float2(feats[p * L + l]);
                xrow[4 + 2 * l] = f.x;
                xrow[5 + 2 * l] = f.y;
            }
        } else {
#pragma unroll
            for (int h = 0; h < kH; ++h) dz[h] = 0.f;
            for (int i = 0; i < K; ++i) xrow[i] = 0.f;
        }
        float hact[kH];
#pragma unroll
        for (int h = 0; h < kH; ++h) {
            float sp, sg;
            softplus100_both(dz[h], sp, sg);
            hact[h] = dsdf * sp;                       // -> dW1
            dz[h] = dsdf * s_net[kOffW1 + h] * sg;     // -> dz
        }
        float4 *dzrow = reinterpret_cast<float4 *>(s_dz + tid * kDzStride);
#pragma unroll
        for (int q = 0; q < kH / 4; ++q) dzrow[q] = make_float4(dz[4 * q], dz[4 * q + 1], dz[4 * q + 2], dz[4 * q + 3]);

        // dW1 / db1: warp transpose-reduce, lane l keeps rows 2l, 2l+1
        warp_transpose_reduce64(hact, lane);
        accW1a += hact[0];
        accW1b += hact[1];
        float ds = dsdf;
#pragma unroll
        for (int o = 16; o; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
        accB1 += ds;

        // ---- phase C: d(features) = W0[:,3:]^T dz, scattered to the table with the trilinear weights
        if (valid && dsdf != 0.f) {
            for (uint32_t l = 0; l < n_active; ++l) {
                const float4 *w0 = reinterpret_cast<const float4 *>(s_net + kOffW0T + (3 + 2 * l) * kH);
                const float4 *w1 = w0 + kH / 4;
                float g0 = 0.f, g1 = 0.f;
#pragma unroll
                for (int q = 0; q < kH / 4; ++q) {
                    float4 u = w0[q], v = w1[q];
                    g0 = fmaf(u.x, dz[4 * q], g0); g0 = fmaf(u.y, dz[4 * q + 1], g0); g0 = fmaf(u.z, dz[4 * q + 2], g0); g0 = fmaf(u.w, dz[4 * q + 3], g0);
                    g1 = fmaf(v.x, dz[4 * q], g1); g1 = fmaf(v.y, dz[4 * q + 1], g1); g1 = fmaf(v.z, dz[4 * q + 2], g1); g1 = fmaf(v.w, dz[4 * q + 3], g1);
                }
                const LevelCtx c = s_lvl[l];
                Cell cell = cell_of(c, r.px, r.py, r.pz);
                float2 *gt = reinterpret_cast<float2 *>(table_grad) + c.offset;
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    float w = corner_weight(cell, k);
                    atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g0, w * g1));
                }
            }
        }
        __syncthreads();

        // ---- phase B: dW0T[col][h] += sum_p x[p][col] * dz[p][h] over the live columns (col 0: bias)
        {
            const float *xg = s_x + (grp * 64) * kXStride + ib;
            const float *dg = s_dz + (grp * 64) * kDzStride + a4;
            switch (n_i) {
                case 1: phase_b<1>(accW, xg, dg); break;
                case 2: phase_b<2>(accW, xg, dg); break;
                case 3: phase_b<3>(accW, xg, dg); break;
                case 4: phase_b<4>(accW, xg, dg); break;
                case 5: phase_b<5>(accW, xg, dg); break;
                case 6: phase_b<6>(accW, xg, dg); break;
                case 7: phase_b<7>(accW, xg, dg); break;
                case 8: phase_b<8>(accW, xg, dg); break;
                default: phase_b<9>(accW, xg, dg); break;
            }
        }
        __syncthreads();
    }

    // flush: folded-layout gradients
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        int col = ib + 4 * i;
        if (col < K) {
            float *dst = (col == 0) ? net_grad + kOffB0 + a4 : net_grad + kOffW0T + (col - 1) * kH + a4;
#pragma unroll
            for (int h = 0; h < 4; ++h) atomicAdd(dst + h, accW[i][h]);
        }
    }
    atomicAdd(net_grad + kOffW1 + 2 * lane, accW1a);
    atomicAdd(net_grad + kOffW1 + 2 * lane + 1, accW1b);
    if (lane == 0) atomicAdd(net_grad + kOffB1, accB1);
}

static int32_t check_net(const snb_net *net, const char *who) {
    SNB_REQUIRE(net, SNB_ERR_NULL, "%s: null net", who);
    SNB_REQUIRE(net->table_f16 && net->net, SNB_ERR_NULL, "%s: null table/net", who);
    SNB_REQUIRE(net->meta.n_levels >= 1 && net->meta.n_levels <= SNB_MAX_LEVELS && net->n_active <= net->meta.n_levels, SNB_ERR_ARG,
                "%s: bad level counts (%u active of %u)", who, net->n_active, net->meta.n_levels);
    SNB_REQUIRE(aligned(net->net, 16) && aligned(net->table_f16, 4), SNB_ERR_ALIGN, "%s: misaligned net/table", who);
    return SNB_OK;
}

static int32_t check_patch_args(const snb_patch_batch *b, const snb_samples *sm, const char *who) {
    SNB_REQUIRE(b && sm, SNB_ERR_NULL, "%s: null struct", who);
    SNB_REQUIRE(b->n_patches >= 0, SNB_ERR_ARG, "%s: n_patches < 0", who);
    SNB_REQUIRE(b->rays_o && b->rays_d && b->plane_n, SNB_ERR_NULL, "%s: null rays", who);
    SNB_REQUIRE(sm->totals && sm->t0 && sm->t1 && sm->patch_idx && sm->end_slot && sm->slot_sample, SNB_ERR_NULL, "%s: null samples", who);
    return SNB_OK;
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_sdf_eval(int64_t n, const float *x, const snb_net *net, int32_t mode, float *out, snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_eval");
    if (rc) return rc;
    SNB_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, SNB_ERR_ARG, "sdf_eval: bad n/mode");
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(x && out, SNB_ERR_NULL, "sdf_eval: null buffer");
    int64_t blocks = cdiv(n, 32 * kFwdWarps);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes);
        configured = true;
    }
    sdf_eval_kernel<<<(unsigned)blocks, 32 * kFwdWarps, kFwdSmemBytes, S(stream)>>>(n, x, *net, make_level_table(net->meta), mode, out);
    SNB_LAUNCH_CHECK("sdf_eval");
    return SNB_OK;
}

extern "C" int32_t snb_sdf_eval_grad(int64_t n, const float *x, const snb_net *net, float *sdf, float *grad, snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_eval_grad");
    if (rc) return rc;
    SNB_REQUIRE(n >= 0, SNB_ERR_ARG, "sdf_eval_grad: n < 0");
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(x && grad, SNB_ERR_NULL, "sdf_eval_grad: null buffer");
    int64_t blocks = cdiv(n, 32 * kFwdWarps);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_eval_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemBytes);
        configured = true;
    }
    sdf_eval_grad_kernel<<<(unsigned)blocks, 32 * kFwdWarps, kGradSmemBytes, S(stream)>>>(n, x, *net, make_level_table(net->meta), sdf, grad);
    SNB_LAUNCH_CHECK("sdf_eval_grad");
    return SNB_OK;
}

extern "C" int32_t snb_sdf_fwd_patch(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, float *sdf, void *feats,
                                     snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_fwd_patch");
    if (rc) return rc;
    rc = check_patch_args(b, sm, "sdf_fwd_patch");
    if (rc) return rc;
    SNB_REQUIRE(sdf && feats, SNB_ERR_NULL, "sdf_fwd_patch: null output");
    // persistent grid: the point count lives on the device (sm->totals)
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_fwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes);
        configured = true;
    }
    sdf_fwd_patch_kernel<<<kNumSMs * 4, 32 * kFwdWarps, kFwdSmemBytes, S(stream)>>>(*b, *net, make_level_table(net->meta), *sm, sdf, (__half2 *)feats);
    SNB_LAUNCH_CHECK("sdf_fwd_patch");
    return SNB_OK;
}

extern "C" int32_t snb_sdf_bwd_patch(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const void *feats,
                                     const float *d_sdf0, const float *d_sdf1, float *table_grad, float *net_grad,
                                     snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_bwd_patch");
    if (rc) return rc;
    rc = check_patch_args(b, sm, "sdf_bwd_patch");
    if (rc) return rc;
    SNB_REQUIRE(feats && d_sdf0 && d_sdf1 && table_grad && net_grad, SNB_ERR_NULL, "sdf_bwd_patch: null buffer");
    SNB_REQUIRE(aligned(table_grad, 8), SNB_ERR_ALIGN, "sdf_bwd_patch: table_grad must be 8-byte aligned");
    static const size_t smem = sizeof(float) * (kNetFloats + kTile * kDzStride + kTile * kXStride);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_bwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    sdf_bwd_patch_kernel<<<kNumSMs * 3, kTile, smem, S(stream)>>>(*b, *net, make_level_table(net->meta), *sm, (const __half2 *)feats, d_sdf0, d_sdf1, table_grad, net_grad);
    SNB_LAUNCH_CHECK("sdf_bwd_patch");
    return SNB_OK;
}
