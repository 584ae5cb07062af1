// fused_sdf.cu -- SDF queries for the fused training step:
//   sdf_eval        points -> sdf (no grad): occupancy-grid update, mesh extraction, validation
//   sdf_fwd_patch   all 9 plane-projected rays of every visible sample (+ own interval ends): positions
//                   are built in-kernel (models/renderer.py:146-159), encode + MLP fused, features kept
//   sdf_bwd_patch   MLP backward (recomputing layer 0 from the kept features), weight gradients as a
//                   shared-memory-tiled contraction, hash-table scatter with vector fp32 atomics
#include "sdf_mma.cuh"
#include "sampler.cuh"
#include "umma.cuh"

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
        float s = warp_sdf_mma<false>(valid, px, py, pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs, nullptr, lane);
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
        float s = warp_sdf_grad_mma(valid, px, py, pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs, ts, lane, g);
        if (valid) {
            if (sdf) sdf[i] = s;
            grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2];
        }
    }
}

// row stride (in half2) of the kept-feature buffer: the live levels rounded up to a power of two (to an even count beyond 8), so that early in the schedule (1-2 live
// levels) a point's row is 4-8 bytes instead of 4 * n_levels -- with the fixed 56-byte rows the backward read one 32-byte sector per point
// for 4 useful bytes (15.9 MB DRAM per launch against 6.9 MB algorithmic, profiles/roofline_traffic.json of round 1)
__host__ __device__ __forceinline__ uint32_t feat_row_stride(uint32_t n_active) {
    return n_active <= 1 ? 1u : n_active <= 2 ? 2u : n_active <= 4 ? 4u : n_active <= 8 ? 8u : ((n_active + 1u) & ~1u);
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
    const uint32_t L = feat_row_stride(net.n_active);
    const int lane = threadIdx.x & 31;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < M; p0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + threadIdx.x;
        const bool valid = p < M;
        PointRef r;
        r.px = r.py = r.pz = 0.f;
        if (valid) r = decode_point(p, S, b, sm);
        float s = warp_sdf_mma<true>(valid, r.px, r.py, r.pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs,
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
    const uint32_t n_active = net.n_active, L = feat_row_stride(n_active);
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

// ---------------------------------------------------------------------------------------------
// analytic normals in the fused step (gradient_method = 'ad'; models/renderer.py:225-226, models/fields.py:107-119)
// ---------------------------------------------------------------------------------------------
// forward: d sdf / d x at every sample START of every in-patch ray (point p = 9 s + k < 9 S), the one-pass tensor-core evaluator of
// sdf_eval_grad_kernel on the in-kernel positions of decode_point.  grad: f32 [9 S, 3].
__global__ void __launch_bounds__(32 * kFwdWarps, 2) sdf_grad_patch_kernel(snb_patch_batch b, snb_net net, LevelTable lt, snb_samples sm,
                                                                           float *__restrict__ grad) {
    extern __shared__ __align__(16) float smem[];
    const FwdSmem sh = fwd_smem_setup(smem, net.net);
    float *ts = smem + kMmaSmemFloats(kFwdWarps) + (threadIdx.x >> 5) * 3 * 32 * kTsStride;
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int S = sm.totals[0];
    const int64_t M = (int64_t)SNB_PATCH * S;
    const int lane = threadIdx.x & 31;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < M; p0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + threadIdx.x;
        const bool valid = p < M;
        PointRef r;
        r.px = r.py = r.pz = 0.f;
        if (valid) r = decode_point(p, S, b, sm);
        float g[3];
        warp_sdf_grad_mma(valid, r.px, r.py, r.pz, table, s_lvl, net.n_active, sh.net, sh.whi, sh.wlo, sh.xs, ts, lane, g);
        if (valid) { grad[3 * p] = g[0]; grad[3 * p + 1] = g[1]; grad[3 * p + 2] = g[2]; }
    }
}

// backward of that gradient: given dg = d loss / d (d sdf / d x) per start point, the parameter gradients of
//     L_g = dg . grad_x sdf = sum_h W1_h s_h u_h,     s = sigmoid(100 z),  z = W0 x_in + b0,  u = W0 r,  r_i = sum_d dg_d d x_in_i / d x_d
// (r = dg on the three position inputs, dg . grad_x feat_i on a feature input).  With s' = 100 s (1 - s):
//     dW1_h += s_h u_h
//     dz_h = W1_h u_h s'_h :   db0 += dz,   dW0 += dz (x) x_in,   d feat_i = sum_h W0[h][i] dz_h      -> table, trilinear VALUE weights
//     v_h  = W1_h s_h      :   dW0 += v (x) r,                    q_i = sum_h W0[h][i] v_h           -> table, DERIVATIVE weights
//                                                                 scale * sum_d dg_d (+-1)_d(corner) prod_{e != d} w_e(corner)
// i.e. the double backward torch.autograd runs for create_graph = True, as two passes of the thread-per-point machinery of
// sdf_bwd_patch_kernel: the table gets both terms of a corner in one reduction (projections from registers); pass A stages (r, v), pass B
// stages (x_in, dz), and each pass adds its outer product to the persistent dW0 accumulators.  Thread = point; FMA pipe (this path is not the shipped default; correctness and no host
// round trip are the point -- the autograd route it replaces takes 25.7 ms per step).
__global__ void __launch_bounds__(kTile, 1) sdf_grad_bwd_patch_kernel(snb_patch_batch b, snb_net net, LevelTable lt, snb_samples sm,
                                                                      const __half2 *__restrict__ feats, const float *__restrict__ d_grad,
                                                                      float *__restrict__ table_grad, float *__restrict__ net_grad) {
    extern __shared__ __align__(16) float smem[];
    float *s_net = smem;                      // kNetFloats
    float *s_dz = s_net + kNetFloats;         // kTile * kDzStride   (v in pass A, dz in pass B)
    float *s_x = s_dz + kTile * kDzStride;    // kTile * kXStride    (r in pass A, x_in in pass B)
    load_net_to_smem(s_net, net.net);
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int tid = threadIdx.x, lane = tid & 31;
    const int S = sm.totals[0];
    const int64_t M = (int64_t)SNB_PATCH * S;
    const uint32_t n_active = net.n_active, L = feat_row_stride(n_active);
    const int K = 4 + 2 * (int)n_active;
    const int grp = tid >> 6, a4 = (tid & 15) * 4, ib = (tid >> 4) & 3;
    const int n_i = (K - ib + 3) / 4;
    float accW[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int h = 0; h < 4; ++h) accW[i][h] = 0.f;
    float accW1a = 0.f, accW1b = 0.f;

    auto outer = [&]() {   // dW0T[col][h] += sum_p s_x[p][col] * s_dz[p][h] over this thread's columns (phase B of sdf_bwd_patch_kernel)
        const float *xg = s_x + (grp * 64) * kXStride + ib;
        const float *dg_ = s_dz + (grp * 64) * kDzStride + a4;
        switch (n_i) {
            case 1: phase_b<1>(accW, xg, dg_); break;
            case 2: phase_b<2>(accW, xg, dg_); break;
            case 3: phase_b<3>(accW, xg, dg_); break;
            case 4: phase_b<4>(accW, xg, dg_); break;
            case 5: phase_b<5>(accW, xg, dg_); break;
            case 6: phase_b<6>(accW, xg, dg_); break;
            case 7: phase_b<7>(accW, xg, dg_); break;
            case 8: phase_b<8>(accW, xg, dg_); break;
            default: phase_b<9>(accW, xg, dg_); break;
        }
    };
    for (int64_t tile0 = (int64_t)blockIdx.x * kTile; tile0 < M; tile0 += (int64_t)gridDim.x * kTile) {
        const int64_t p = tile0 + tid;
        const bool valid = p < M;
        PointRef r;
        r.px = r.py = r.pz = 0.f;
        float dg[3] = {0.f, 0.f, 0.f};
        if (valid) {
            r = decode_point(p, S, b, sm);
            dg[0] = __ldg(d_grad + 3 * p); dg[1] = __ldg(d_grad + 3 * p + 1); dg[2] = __ldg(d_grad + 3 * p + 2);
        }
        const bool live = valid && (dg[0] != 0.f || dg[1] != 0.f || dg[2] != 0.f);
        float *xrow = s_x + tid * kXStride;
        float4 *dzrow = reinterpret_cast<float4 *>(s_dz + tid * kDzStride);
        // ---- z (kept features) and u = W0 r, r staged as pass A's input row: [0 | dg | dg . grad feat_l]
        float z[kH], u[kH];
#pragma unroll
        for (int h = 0; h < kH; ++h) u[h] = 0.f;
        if (valid) {
            layer0<false, true>(r.px, r.py, r.pz, nullptr, s_lvl, n_active, s_net, const_cast<__half2 *>(feats + p * L), z);
        } else {
#pragma unroll
            for (int h = 0; h < kH; ++h) z[h] = 0.f;
        }
        xrow[0] = 0.f; xrow[1] = dg[0]; xrow[2] = dg[1]; xrow[3] = dg[2];
        rank1_update(u, s_net + kOffW0T + 0 * kH, dg[0]);
        rank1_update(u, s_net + kOffW0T + 1 * kH, dg[1]);
        rank1_update(u, s_net + kOffW0T + 2 * kH, dg[2]);
        for (uint32_t l = 0; l < n_active; ++l) {
            float rx = 0.f, ry = 0.f;
            if (live) {
                const LevelCtx c = s_lvl[l];
                const Cell cell = cell_of(c, r.px, r.py, r.pz);
                float2 dv[3];
                interp_level_grad(c, cell, table, dv);
                rx = dg[0] * dv[0].x + dg[1] * dv[1].x + dg[2] * dv[2].x;
                ry = dg[0] * dv[0].y + dg[1] * dv[1].y + dg[2] * dv[2].y;
            }
            xrow[4 + 2 * l] = rx;
            xrow[5 + 2 * l] = ry;
            rank1_update(u, s_net + kOffW0T + (3 + 2 * l) * kH, rx);
            rank1_update(u, s_net + kOffW0T + (4 + 2 * l) * kH, ry);
        }
        // ---- s, s'; pass A row v = W1 s; dz = W1 u s' kept in u[]; dW1 += s u
        {
            float hact[kH];
#pragma unroll
            for (int h = 0; h < kH; ++h) {
                float sp, sg;
                softplus100_both(z[h], sp, sg);
                const float w1 = s_net[kOffW1 + h];
                hact[h] = live ? sg * u[h] : 0.f;
                z[h] = live ? w1 * sg : 0.f;                                  // v
                u[h] = live ? w1 * u[h] * 100.f * sg * (1.f - sg) : 0.f;       // dz
            }
#pragma unroll
            for (int q = 0; q < kH / 4; ++q) dzrow[q] = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
            warp_transpose_reduce64(hact, lane);
            accW1a += hact[0];
            accW1b += hact[1];
        }
        // ---- table: both terms of a corner in ONE reduction -- value weight x d feat (from dz) + derivative weight x q (from v), the two
        // projections taken straight from the registers
        if (live) {
            for (uint32_t l = 0; l < n_active; ++l) {
                const float4 *w0 = reinterpret_cast<const float4 *>(s_net + kOffW0T + (3 + 2 * l) * kH);
                const float4 *w1 = w0 + kH / 4;
                float g0 = 0.f, g1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int q = 0; q < kH / 4; ++q) {
                    const float4 a = w0[q], bq = w1[q];
                    g0 = fmaf(a.x, u[4 * q], g0); g0 = fmaf(a.y, u[4 * q + 1], g0); g0 = fmaf(a.z, u[4 * q + 2], g0); g0 = fmaf(a.w, u[4 * q + 3], g0);
                    g1 = fmaf(bq.x, u[4 * q], g1); g1 = fmaf(bq.y, u[4 * q + 1], g1); g1 = fmaf(bq.z, u[4 * q + 2], g1); g1 = fmaf(bq.w, u[4 * q + 3], g1);
                    q0 = fmaf(a.x, z[4 * q], q0); q0 = fmaf(a.y, z[4 * q + 1], q0); q0 = fmaf(a.z, z[4 * q + 2], q0); q0 = fmaf(a.w, z[4 * q + 3], q0);
                    q1 = fmaf(bq.x, z[4 * q], q1); q1 = fmaf(bq.y, z[4 * q + 1], q1); q1 = fmaf(bq.z, z[4 * q + 2], q1); q1 = fmaf(bq.w, z[4 * q + 3], q1);
                }
                const LevelCtx c = s_lvl[l];
                const Cell cell = cell_of(c, r.px, r.py, r.pz);
                float2 *gt = reinterpret_cast<float2 *>(table_grad) + c.offset;
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    float wd = 0.f;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int e0 = (d + 1) % 3, e1 = (d + 2) % 3;
                        const float a0 = ((k >> e0) & 1u) ? cell.w[e0] : 1.f - cell.w[e0];
                        const float a1 = ((k >> e1) & 1u) ? cell.w[e1] : 1.f - cell.w[e1];
                        wd += (((k >> d) & 1u) ? dg[d] : -dg[d]) * a0 * a1;
                    }
                    wd *= c.scale;
                    const float w = corner_weight(cell, k);
                    atomicAdd(gt + corner_index(c, cell, k), make_float2(fmaf(w, g0, wd * q0), fmaf(w, g1, wd * q1)));
                }
            }
        }
        __syncthreads();
        // ---- pass A: dW0 += v (x) r
        outer();
        __syncthreads();
        // ---- pass B: rows x_in and dz;  dW0 += dz (x) x_in (column 0: db0)
        if (valid) {
            xrow[0] = 1.f; xrow[1] = r.px; xrow[2] = r.py; xrow[3] = r.pz;
            for (uint32_t l = 0; l < n_active; ++l) {
                const float2 f = __half22float2(feats[p * L + l]);
                xrow[4 + 2 * l] = f.x;
                xrow[5 + 2 * l] = f.y;
            }
        } else {
            for (int i = 0; i < K; ++i) xrow[i] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < kH / 4; ++q) dzrow[q] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
        __syncthreads();
        outer();
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
}

// ---------------------------------------------------------------------------------------------
// backward with tcgen05 (5th-gen tensor cores, TMEM accumulators)
// ---------------------------------------------------------------------------------------------
// CTA = 128 threads = one tile of 128 points; thread t owns point t, and TMEM lane t is row t of every accumulator, so
// tcgen05.ld hands each thread exactly its point's 64 pre-activations / 32 feature gradients (no layout shuffling).
//   (1) Z[128 x 64]  = X1[128 x 40] * B1^T        tcgen05.mma kind::tf32, 5 k-steps x (hi, lo) weights
//         X1 = [ fp16 features (exact in TF32) | x_hi | x_lo | 1 | 0 ],   B1 = [ W0feat | W0xyz | W0xyz | b0 | 0 ] split hi + lo:
//         products exact to 2^-22 (the recomputed z feeds sigmoid(100 z), it must match the forward)
//   (2) dz = dsdf * W1 * sigmoid(100 z) per thread -> TF32-rounded into a K-major tile;   dW1 / db1 by warp reductions
//   (3) U[128 x 32]  = dz[128 x 64] * W0feat (hi)  tcgen05.mma, 8 k-steps  -> d loss / d features, scattered per thread
//         (single TF32 pass: 2^-11 relative, the precision class of tiny-cuda-nn's fp16 gradient path)
//   (4) dW0T[40 x 64] += X1^T dz                     warp-level mma.sync m16n8k8 on the same two tiles (contraction over the
//         points; tcgen05 cannot take MN-major TF32 operands without the 128B_BASE32B swizzle), persistent accumulators.
//         dz enters as the TF32 tile (rounded to nearest: unbiased, 2^-12 rms per element, averaged over the points); a bf16
//         side tile for the residual was measured (+14 % kernel time for errors far below the minibatch noise) and dropped.
// (3) runs on the tensor pipe while the warps do (4).
// KF = feature slots of the tiles: 32 (up to 16 levels; 80 KB of shared memory, 2 CTAs/SM) or 8 (up to 4 levels; 52 KB, 4 CTAs/SM --
// the first ~1 400 iterations of the schedule, where a tile's dependent phases are short and latency, not work, sets the pace).
template <int KF>
struct BwdCfg {
    static constexpr int kK1 = KF + 8;                       // columns of X1: KF feature slots | x_hi (3) | x_lo (3) | 1 | pad
    static constexpr int kColXhi = KF, kColXlo = KF + 3, kColOne = KF + 6;
    static constexpr int kN2 = KF < 16 ? 16 : KF;            // N of the feature-gradient MMA (multiple of 16 for M = 128)
    static constexpr int kNMT = (kK1 + 15) / 16;             // 16-column tiles of the weight-gradient mma.sync
    static constexpr uint32_t kTmemCols = 128;               // Z: columns 0..63, U: 64..64+kN2
    static constexpr size_t kSmem = umma::tile_bytes(64, kK1) * 2 + umma::tile_bytes(kN2, 64) + umma::tile_bytes(128, kK1) + umma::tile_bytes(128, 64);
    static constexpr int kMinBlocks = KF <= 8 ? 3 : 2;
};

// -DSNB_BWD_DEBUG: per-phase clock64 cycles of thread 0 of every CTA, summed over the tiles (scripts/bwd_phase_times.py,
// profiles/r02_bwd_phase_times.txt).  At 1 live level a 128-point tile costs ~11.5 k cycles: elementwise block 3.9 k, dW0 mma.sync loop
// 2.7 k, scatter 1.6 k (+2.4 k per further level), issue of the next tile's loads 1.4 k, Z MMA 0.6 k, stage 0.4 k, the three barriers 0.5 k.
#ifdef SNB_BWD_DEBUG
__device__ unsigned long long g_bwd_dbg[16];
#define BWD_T(i) do { if (tid == 0) { const long long _n = clock64(); atomicAdd(&g_bwd_dbg[i], (unsigned long long)(_n - dbg_t)); dbg_t = _n; } } while (0)
#else
#define BWD_T(i) do { } while (0)
#endif

template <int KF>
__global__ void __launch_bounds__(128, BwdCfg<KF>::kMinBlocks) sdf_bwd_patch_umma_kernel(snb_patch_batch b, snb_net net, LevelTable lt, snb_samples sm,
                                                                    const __half2 *__restrict__ feats,
                                                                    const float *__restrict__ d_sdf0,
                                                                    const float *__restrict__ d_sdf1,
                                                                    float *__restrict__ table_grad, float *__restrict__ net_grad,
                                                                    int *__restrict__ err_flag) {
    using Cfg = BwdCfg<KF>;
    constexpr int kK1 = Cfg::kK1, kColXhi = Cfg::kColXhi, kColXlo = Cfg::kColXlo, kColOne = Cfg::kColOne, kN2 = Cfg::kN2, kNMT = Cfg::kNMT;
    constexpr uint32_t kBwdTmemCols = Cfg::kTmemCols;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_z, bar_u;
    __shared__ uint32_t s_tmem;
    __shared__ float s_w1[kH];
    __shared__ LevelCtx s_lvl[SNB_MAX_LEVELS];
    uint8_t *b1hi = smem_raw, *b1lo = b1hi + umma::tile_bytes(64, kK1), *b2 = b1lo + umma::tile_bytes(64, kK1);
    uint8_t *a1 = b2 + umma::tile_bytes(kN2, 64), *a2 = a1 + umma::tile_bytes(128, kK1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t n_active = net.n_active, L = feat_row_stride(n_active);

    // ---- one-time staging: weights (TF32 hi / lo), level table, TMEM, barriers
    {
        const float *W = net.net;
        for (int e = tid; e < 64 * kK1; e += 128) {
            const int h = e / kK1, k = e % kK1;
            float w = 0.f;
            bool lo_zero = false;
            if (k < KF) w = __ldg(W + kOffW0T + (3 + k) * kH + h);
            else if (k < kColXlo) w = __ldg(W + kOffW0T + (k - kColXhi) * kH + h);
            else if (k < kColOne) { w = __ldg(W + kOffW0T + (k - kColXlo) * kH + h); lo_zero = true; }   // x_lo * W_lo is below 2^-22
            else if (k == kColOne) w = __ldg(W + kOffB0 + h);
            const float hi = __uint_as_float(to_tf32(w));
            *reinterpret_cast<float *>(b1hi + umma::kmajor_off(h, k, kK1)) = hi;
            *reinterpret_cast<float *>(b1lo + umma::kmajor_off(h, k, kK1)) = lo_zero ? 0.f : __uint_as_float(to_tf32(w - hi));
        }
        for (int e = tid; e < kN2 * 64; e += 128) {
            const int j = e / 64, h = e % 64;
            *reinterpret_cast<float *>(b2 + umma::kmajor_off(j, h, 64)) = j < KF ? __uint_as_float(to_tf32(__ldg(W + kOffW0T + (3 + j) * kH + h))) : 0.f;
        }
        if (tid < kH) s_w1[tid] = __ldg(W + kOffW1 + tid);
        if (tid < SNB_MAX_LEVELS) s_lvl[tid] = lt.lv[tid];
        if (warp == 0) umma::tmem_alloc(&s_tmem, kBwdTmemCols);
        if (tid == 0) { umma::mbar_init(&bar_z, 1); umma::mbar_init(&bar_u, 1); }
        umma::fence_smem_to_async_proxy();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }
    const uint32_t tmem = s_tmem;
    const uint32_t tmem_lane = tmem + ((uint32_t)(32 * warp) << 16);
    const uint32_t a1_addr = umma::smem_u32(a1), a2_addr = umma::smem_u32(a2);
    const uint32_t b1hi_addr = umma::smem_u32(b1hi), b1lo_addr = umma::smem_u32(b1lo), b2_addr = umma::smem_u32(b2);
    const uint32_t idesc_z = umma::idesc_tf32(128, 64), idesc_u = umma::idesc_tf32(128, kN2);

    const int S = sm.totals[0], E = sm.totals[1];
    const int64_t Q = (int64_t)S + E;                            // sample starts + own interval ends; point p = 9 q + k
    // Tile = kTileQ consecutive q x 9 rays, RAY-major over the threads: thread t -> ray k = t / kTileQ, q = q0 + t % kTileQ, so that
    // consecutive lanes walk ALONG a ray (spacing = the marching step, far below the cell size of most levels) and the
    // segmented reduction of the scatter below finds long runs of lanes in the same grid cell.
    constexpr int kTileQ = 14;
    const int tk = tid / kTileQ, tj = tid % kTileQ;
    const int n_feat_steps = (int)(2 * n_active + 7) >> 3;     // feature k-steps of (1) that can be non-zero
    // (4): output tiles [3 m-tiles (input columns 0..15, 16..31, 32..47) x 8 n-tiles (hidden)] dealt to the 4 warps
    float wacc[2 * kNMT][4];
#pragma unroll
    for (int i = 0; i < 2 * kNMT; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) wacc[i][c] = 0.f;
    float accW1a = 0.f, accW1b = 0.f, accB1 = 0.f;
    uint32_t phase = 0;
    bool failed = false;

    // Point data of a tile (decode chain: sample -> patch -> rays, seeds, kept features) is fetched one tile AHEAD into registers,
    // so its dependent global loads overlap the previous tile's MMAs / scatter instead of heading every tile.
    struct TileRegs {
        bool valid;
        float dsdf, px, py, pz;
        __half2 f[KF / 2];
    };
    auto fetch = [&](int64_t q0, TileRegs &o) {
        const int64_t p = (q0 + tj) * SNB_PATCH + tk;
        o.valid = tk < SNB_PATCH && q0 + tj < Q;
        o.dsdf = 0.f;
        o.px = o.py = o.pz = 0.f;
#pragma unroll
        for (int l = 0; l < KF / 2; ++l) o.f[l] = __float2half2_rn(0.f);
        if (o.valid) {
            const PointRef r = decode_point(p, S, b, sm);
            o.px = r.px; o.py = r.py; o.pz = r.pz;
            if (!r.is_end) {
                o.dsdf = __ldg(d_sdf0 + (int64_t)r.s * SNB_PATCH + r.k);
                // this start also served as the previous interval's end when that interval had no own end query
                if (r.s > 0 && __ldg(sm.end_slot + r.s - 1) < 0) o.dsdf += __ldg(d_sdf1 + (int64_t)(r.s - 1) * SNB_PATCH + r.k);
            } else {
                o.dsdf = __ldg(d_sdf1 + (int64_t)r.s * SNB_PATCH + r.k);
            }
            const __half2 *fr = feats + p * L;
#pragma unroll
            for (int l = 0; l < KF / 2; ++l)
                if (l < (int)n_active) o.f[l] = fr[l];
        }
    };
    TileRegs nxt;
    fetch((int64_t)blockIdx.x * kTileQ, nxt);
#ifdef SNB_BWD_DEBUG
    long long dbg_t = clock64();
#endif

    for (int64_t q0 = (int64_t)blockIdx.x * kTileQ; q0 < Q; q0 += (int64_t)gridDim.x * kTileQ) {
        // ---- stage this thread's point: row tid of X1
        const bool valid = nxt.valid;
        const float dsdf = nxt.dsdf;
        PointRef r;
        r.px = nxt.px; r.py = nxt.py; r.pz = nxt.pz;
        {
#pragma unroll
            for (int q = 0; q < KF / 4; ++q) {  // chunks of 4 feature columns (2 levels each)
                const float2 f0 = __half22float2(nxt.f[2 * q]), f1 = __half22float2(nxt.f[2 * q + 1]);
                *reinterpret_cast<float4 *>(a1 + umma::kmajor_off(tid, 4 * q, kK1)) = make_float4(f0.x, f0.y, f1.x, f1.y);
            }
            const float xh = __uint_as_float(to_tf32(r.px)), yh = __uint_as_float(to_tf32(r.py)), zh = __uint_as_float(to_tf32(r.pz));
            const float xl = __uint_as_float(to_tf32(r.px - xh)), yl = __uint_as_float(to_tf32(r.py - yh)), zl = __uint_as_float(to_tf32(r.pz - zh));
            *reinterpret_cast<float4 *>(a1 + umma::kmajor_off(tid, KF, kK1)) = make_float4(xh, yh, zh, xl);
            *reinterpret_cast<float4 *>(a1 + umma::kmajor_off(tid, KF + 4, kK1)) = make_float4(yl, zl, 1.f, 0.f);
        }
        BWD_T(0);   // stage
        fetch(q0 + (int64_t)gridDim.x * kTileQ, nxt);      // next tile's loads are in flight from here on
        BWD_T(9);   // issue of the next fetch
        umma::fence_smem_to_async_proxy();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        BWD_T(1);   // barrier 1

        // ---- (1) Z = X1 B1^T
        if (tid == 0) {
            uint32_t acc = 0;
            for (int s = 0; s < kK1 / 8; ++s) {
                if (s < KF / 8 && s >= n_feat_steps) continue;   // all-zero feature columns
                const uint64_t ad = umma::kmajor_desc(a1_addr + 256u * s, kK1);
                umma::mma_tf32(tmem, ad, umma::kmajor_desc(b1hi_addr + 256u * s, kK1), idesc_z, acc);
                umma::mma_tf32(tmem, ad, umma::kmajor_desc(b1lo_addr + 256u * s, kK1), idesc_z, 1);
                acc = 1;
            }
            umma::commit(&bar_z);
        }
        if (!umma::mbar_wait(&bar_z, phase)) failed = true;
        umma::fence_after_sync();
        BWD_T(2);   // MMA Z issue + wait

        // ---- (2) dz, dW1 / db1
        {
            float hact[kH];
#pragma unroll
            for (int c0 = 0; c0 < kH; c0 += 16) {
                float z[16];
                umma::tmem_ld16(tmem_lane + (uint32_t)c0, z);
                umma::tmem_ld_wait();
                float dzv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float sp, sg;
                    softplus100_both_lg2(z[i], sp, sg);
                    hact[c0 + i] = dsdf * sp;
                    dzv[i] = __uint_as_float(to_tf32(dsdf * s_w1[c0 + i] * sg));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4 *>(a2 + umma::kmajor_off(tid, c0 + 4 * q, 64)) = make_float4(dzv[4 * q], dzv[4 * q + 1], dzv[4 * q + 2], dzv[4 * q + 3]);

            }
            warp_transpose_reduce64(hact, lane);
            accW1a += hact[0];
            accW1b += hact[1];
            float ds = dsdf;
#pragma unroll
            for (int o = 16; o; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
            accB1 += ds;
        }
        BWD_T(3);   // elementwise block + transpose-reduce
        umma::fence_smem_to_async_proxy();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        BWD_T(4);   // barrier 2

        // ---- (3) U = dz W0feat  (async on the tensor pipe)
        if (tid == 0 && n_active > 0) {
            for (int s = 0; s < 8; ++s)
                umma::mma_tf32(tmem + 64u, umma::kmajor_desc(a2_addr + 256u * s, 64), umma::kmajor_desc(b2_addr + 256u * s, 64), idesc_u, s > 0);
            umma::commit(&bar_u);
        }

        // ---- (4) dW0T[col][h] += sum_p X1[p][col] dz[p][h]   (mma.sync; A = X1^T, B = dz read from the K-major tiles)
        // warp owns the hidden tiles nt = warp, warp + 4 and all three column tiles: wacc[2 mt + j].  Fragment addresses in the
        // core-matrix layout: point 8 ks + t (+4) -> ks * (K/4) * 128 + t * 16 (+64); column c -> (c / 4) * 128 + (c % 4) * 4.
        {
            const uint8_t *ab = a1 + (g >> 2) * 128 + (g & 3) * 4 + t * 16;              // column g of m-tile 0; m-tile mt: + mt * 512
            const uint8_t *bb0 = a2 + (2 * warp + (g >> 2)) * 128 + (g & 3) * 4 + t * 16;  // hidden 8 warp + g; second tile: + 8 * 128
#pragma unroll 2
            for (int ks = 0; ks < 16; ++ks) {
                const uint8_t *ak = ab + ks * (kK1 / 4) * 128, *bk = bb0 + ks * (64 / 4) * 128;
                const uint32_t b00 = *reinterpret_cast<const uint32_t *>(bk), b01 = *reinterpret_cast<const uint32_t *>(bk + 64);
                const uint32_t b10 = *reinterpret_cast<const uint32_t *>(bk + 1024), b11 = *reinterpret_cast<const uint32_t *>(bk + 1024 + 64);
#pragma unroll
                for (int mt = 0; mt < kNMT; ++mt) {
                    // a feature-only tile whose columns are all beyond the active levels contributes nothing
                    if (16 * mt + 16 <= KF && 16 * mt >= 2 * (int)n_active) continue;
                    const bool hi_half = 16 * mt + 8 < kK1;          // columns 16 mt + 8 .. + 15 exist
                    uint32_t a[4];
                    a[0] = *reinterpret_cast<const uint32_t *>(ak + mt * 512);
                    a[2] = *reinterpret_cast<const uint32_t *>(ak + mt * 512 + 64);
                    a[1] = hi_half ? *reinterpret_cast<const uint32_t *>(ak + mt * 512 + 256) : 0u;
                    a[3] = hi_half ? *reinterpret_cast<const uint32_t *>(ak + mt * 512 + 256 + 64) : 0u;
                    mma_tf32(wacc[2 * mt], a, b00, b01);
                    mma_tf32(wacc[2 * mt + 1], a, b10, b11);
                }
            }
        }

        BWD_T(5);   // MMA U issue + dW0 mma.sync loop
        // ---- scatter d loss / d features (thread-per-point, straight from TMEM)
        if (n_active > 0) {
            if (!umma::mbar_wait(&bar_u, phase)) failed = true;
            umma::fence_after_sync();
            BWD_T(6);   // wait U
            // The table scatter is bound by the L2 atomic units (~64 red.v2.f32 lanes per clock chip-wide: 67 M of them are the whole
            // 530 us of the FMA kernel at 14 levels), so contributions to the same grid cell are summed in the warp first:
            // segmented inclusive scan over runs of consecutive lanes with the same cell (16 values: 8 corners x 2 features),
            // the last lane of each run issues the run's 8 atomics.  Levels too fine to form runs take the direct path.
            const bool live = valid && dsdf != 0.f;
#pragma unroll 1
            for (uint32_t l = 0; l < n_active; ++l) {
                float g0, g1;
                umma::tmem_ld2(tmem_lane + 64u + 2u * l, g0, g1);    // warp-collective: every lane takes part
                umma::tmem_ld_wait();
                const LevelCtx c = s_lvl[l];
                const Cell cell = cell_of(c, r.px, r.py, r.pz);
                const unsigned long long key = live ? (((unsigned long long)(cell.g[0] & 0x1FFFFFu)) | ((unsigned long long)(cell.g[1] & 0x1FFFFFu) << 21) |
                                                       ((unsigned long long)(cell.g[2] & 0x1FFFFFu) << 42))
                                                    : ~0ull;
                const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
                float2 *gt = reinterpret_cast<float2 *>(table_grad) + c.offset;
                if (__popc(heads) > 20) {                             // (almost) no two neighbours share a cell
                    if (live) {
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k) {
                            float w = corner_weight(cell, k);
                            atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g0, w * g1));
                        }
                    }
                    continue;
                }
                const int run_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
                float vx[8], vy[8];
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const float w = live ? corner_weight(cell, k) : 0.f;
                    vx[k] = w * g0;
                    vy[k] = w * g1;
                }
                // a run longer than 16 lanes needs the last doubling step: some lane >= 16 whose run starts at or before lane - 16
                const bool long_run = __any_sync(0xffffffffu, lane - 16 >= run_start);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    if (d == 16 && !long_run) break;
                    const bool take = lane - d >= run_start;
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        const float ux = __shfl_up_sync(0xffffffffu, vx[k], d), uy = __shfl_up_sync(0xffffffffu, vy[k], d);
                        if (take) { vx[k] += ux; vy[k] += uy; }
                    }
                }
                if (tail && live) {
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) atomicAdd(gt + corner_index(c, cell, k), make_float2(vx[k], vy[k]));
                }
            }
        }
        BWD_T(7);   // scatter
        phase ^= 1u;
        umma::fence_before_sync();
        __syncthreads();          // tiles a1 / a2 and both accumulators are free again
        umma::fence_after_sync();
        BWD_T(8);   // barrier 3
#ifdef SNB_BWD_DEBUG
        if (tid == 0) atomicAdd(&g_bwd_dbg[15], 1ull);
#endif
    }

    // ---- flush
#pragma unroll
    for (int i = 0; i < 2 * kNMT; ++i) {
        const int mt = i >> 1, nt = warp + 4 * (i & 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int col = 16 * mt + g + 8 * (c >> 1), h = 8 * nt + 2 * t + (c & 1);
            float *dst = nullptr;
            if (col < KF) { if (col < 2 * (int)n_active) dst = net_grad + kOffW0T + (3 + col) * kH + h; }
            else if (col < kColXlo) dst = net_grad + kOffW0T + (col - kColXhi) * kH + h;
            else if (col < kColOne) dst = net_grad + kOffW0T + (col - kColXlo) * kH + h;
            else if (col == kColOne) dst = net_grad + kOffB0 + h;
            if (dst && wacc[i][c] != 0.f) atomicAdd(dst, wacc[i][c]);
        }
    }
    atomicAdd(net_grad + kOffW1 + 2 * lane, accW1a);
    atomicAdd(net_grad + kOffW1 + 2 * lane + 1, accW1b);
    if (lane == 0) atomicAdd(net_grad + kOffB1, accB1);
    if (failed && tid == 0 && err_flag) atomicExch(err_flag, 1);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kBwdTmemCols);
}

// ---------------------------------------------------------------------------------------------
// MLP backward of the split form (more than 4 live levels): same contractions as sdf_bwd_patch_umma_kernel<32>, no scatter
// ---------------------------------------------------------------------------------------------
// CTA = 256 threads on ONE 128-point tile: warps w and w + 4 share TMEM lanes 32 (w & 3) .. + 31 (a warp may touch the lane
// quarter warp_id % 4) and split the 64 hidden units / the 16 level slots between them, i.e. TWO threads per point.  The
// thread-per-point kernel above keeps 64 pre-activations + 64 dz in one thread (163 registers, 8 warps per SM: 12 % of the warp
// slots, 33 % issue-active -- profiles/r02_ncu_bwd_split_v1_it4800.txt); halving the per-thread slice doubles the resident warps at the
// same shared-memory footprint and halves every dependent chain (softplus block, transpose-reduce, weight-gradient MMAs).
//   half 0 (warps 0-3): decodes the point, stages feature slots 0..15 + position columns, hidden units 0..31, levels 0..7 of the output
//   half 1 (warps 4-7): stages feature slots 16..31, hidden units 32..63, levels 8..15
// Output (instead of the table scatter): ws_dfeat[(level * 9 + ray) * qcap + q] = d loss / d feature (float2), ws_pos[ray * qcap + q]
// = (x, y, z, live) -- consumed by hash_scatter_kernel.
constexpr int kMlpThreads = 256;

// v[32] per lane -> lane l ends with the warp sum of v[l] in v[0]
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, half = 16; off >= 1; off >>= 1, half >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            float send = up ? v[i] : v[i + half];
            float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// LBO: leading-dimension byte offset of the two point tiles (X1 and dz): 192 = padded cores, conflict-free fragment loads in (4)
template <int LBO>
constexpr size_t mlp_smem_bytes() {
    return umma::tile_bytes(64, BwdCfg<32>::kK1) * 2 + umma::tile_bytes(BwdCfg<32>::kN2, 64) + umma::tile_bytes_lbo<LBO>(128, BwdCfg<32>::kK1) +
           umma::tile_bytes_lbo<LBO>(128, 64);
}

template <int LBO>
__global__ void __launch_bounds__(kMlpThreads, 2) sdf_bwd_mlp_umma_kernel(snb_patch_batch b, snb_net net, snb_samples sm,
                                                                          const __half2 *__restrict__ feats,
                                                                          const float *__restrict__ d_sdf0,
                                                                          const float *__restrict__ d_sdf1, float *__restrict__ net_grad,
                                                                          int *__restrict__ err_flag, float2 *__restrict__ ws_dfeat,
                                                                          float4 *__restrict__ ws_pos, int64_t qcap) {
    constexpr int KF = 32;
    using Cfg = BwdCfg<KF>;
    constexpr int kK1 = Cfg::kK1, kColXhi = Cfg::kColXhi, kColXlo = Cfg::kColXlo, kColOne = Cfg::kColOne, kN2 = Cfg::kN2, kNMT = Cfg::kNMT;
    constexpr uint32_t kBwdTmemCols = Cfg::kTmemCols;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_z, bar_u;
    __shared__ uint32_t s_tmem;
    __shared__ float s_w1[kH];
    uint8_t *b1hi = smem_raw, *b1lo = b1hi + umma::tile_bytes(64, kK1), *b2 = b1lo + umma::tile_bytes(64, kK1);
    uint8_t *a1 = b2 + umma::tile_bytes(kN2, 64), *a2 = a1 + umma::tile_bytes_lbo<LBO>(128, kK1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = warp >> 2, row = 32 * (warp & 3) + lane;      // this thread's point (row of the tiles, TMEM lane) and its half
    const int g = lane >> 2, t = lane & 3;
    const uint32_t n_active = net.n_active, L = feat_row_stride(n_active);

    // ---- one-time staging: weights (TF32 hi / lo), TMEM, barriers
    {
        const float *W = net.net;
        for (int e = tid; e < 64 * kK1; e += kMlpThreads) {
            const int h = e / kK1, k = e % kK1;
            float w = 0.f;
            bool lo_zero = false;
            if (k < KF) w = __ldg(W + kOffW0T + (3 + k) * kH + h);
            else if (k < kColXlo) w = __ldg(W + kOffW0T + (k - kColXhi) * kH + h);
            else if (k < kColOne) { w = __ldg(W + kOffW0T + (k - kColXlo) * kH + h); lo_zero = true; }   // x_lo * W_lo is below 2^-22
            else if (k == kColOne) w = __ldg(W + kOffB0 + h);
            const float hi = __uint_as_float(to_tf32(w));
            *reinterpret_cast<float *>(b1hi + umma::kmajor_off(h, k, kK1)) = hi;
            *reinterpret_cast<float *>(b1lo + umma::kmajor_off(h, k, kK1)) = lo_zero ? 0.f : __uint_as_float(to_tf32(w - hi));
        }
        for (int e = tid; e < kN2 * 64; e += kMlpThreads) {
            const int j = e / 64, h = e % 64;
            *reinterpret_cast<float *>(b2 + umma::kmajor_off(j, h, 64)) = j < KF ? __uint_as_float(to_tf32(__ldg(W + kOffW0T + (3 + j) * kH + h))) : 0.f;
        }
        if (tid < kH) s_w1[tid] = __ldg(W + kOffW1 + tid);
        if (warp == 0) umma::tmem_alloc(&s_tmem, kBwdTmemCols);
        if (tid == 0) { umma::mbar_init(&bar_z, 1); umma::mbar_init(&bar_u, 1); }
        umma::fence_smem_to_async_proxy();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }
    const uint32_t tmem = s_tmem;
    const uint32_t tmem_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t a1_addr = umma::smem_u32(a1), a2_addr = umma::smem_u32(a2);
    const uint32_t b1hi_addr = umma::smem_u32(b1hi), b1lo_addr = umma::smem_u32(b1lo), b2_addr = umma::smem_u32(b2);
    const uint32_t idesc_z = umma::idesc_tf32(128, 64), idesc_u = umma::idesc_tf32(128, kN2);

    const int S = sm.totals[0], E = sm.totals[1];
    const int64_t Q = (int64_t)S + E;                            // sample starts + own interval ends; point p = 9 q + k
    // tile = kTileQ consecutive q x 9 rays, ray-major over the rows (row r -> ray r / kTileQ, q = q0 + r % kTileQ): the layout
    // hash_scatter_kernel reads back along q
    constexpr int kTileQ = 14;
    const int tk = row / kTileQ, tj = row % kTileQ;
    const int n_feat_steps = (int)(2 * n_active + 7) >> 3;     // feature k-steps of (1) that can be non-zero
    // (4): warp w owns hidden n-tile w (8 columns) and all kNMT 16-row column tiles of X1
    float wacc[kNMT][4];
#pragma unroll
    for (int i = 0; i < kNMT; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) wacc[i][c] = 0.f;
    // dW1 partial sums of this thread's point rows over all its tiles, per hidden unit of its half (transpose-reduced over the warp once,
    // at the end: a per-tile reduction costs 31 shuffles + ~90 selects / adds per thread and tile)
    float accH[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) accH[i] = 0.f;
    float accB1 = 0.f;
    uint32_t phase = 0;
    bool failed = false;

    // point data one tile ahead in registers (see sdf_bwd_patch_umma_kernel): half 0 carries the position, each half its 8 level slots
    struct TileRegs {
        bool valid;
        float dsdf, px, py, pz;
        __half2 f[8];
    };
    auto fetch = [&](int64_t q0, TileRegs &o) {
        const int64_t q = q0 + tj, p = q * SNB_PATCH + tk;
        o.valid = tk < SNB_PATCH && q < Q;
        o.dsdf = 0.f;
        o.px = o.py = o.pz = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) o.f[l] = __float2half2_rn(0.f);
        if (o.valid) {
            int s;
            bool is_end;
            if (half == 0) {
                const PointRef r = decode_point(p, S, b, sm);
                o.px = r.px; o.py = r.py; o.pz = r.pz;
                s = r.s; is_end = r.is_end;
            } else {
                is_end = q >= S;
                s = is_end ? __ldg(sm.slot_sample + (q - S)) : (int)q;
            }
            if (!is_end) {
                o.dsdf = __ldg(d_sdf0 + (int64_t)s * SNB_PATCH + tk);
                // this start also served as the previous interval's end when that interval had no own end query
                if (s > 0 && __ldg(sm.end_slot + s - 1) < 0) o.dsdf += __ldg(d_sdf1 + (int64_t)(s - 1) * SNB_PATCH + tk);
            } else {
                o.dsdf = __ldg(d_sdf1 + (int64_t)s * SNB_PATCH + tk);
            }
            const __half2 *fr = feats + p * L + 8 * half;
#pragma unroll
            for (int l = 0; l < 8; ++l)
                if (8 * half + l < (int)n_active) o.f[l] = fr[l];
        }
    };
    TileRegs nxt;
    fetch((int64_t)blockIdx.x * kTileQ, nxt);

    for (int64_t q0 = (int64_t)blockIdx.x * kTileQ; q0 < Q; q0 += (int64_t)gridDim.x * kTileQ) {
        // ---- stage row `row` of X1: this half's 16 feature columns; half 0 also the position columns
        const bool valid = nxt.valid;
        const float dsdf = nxt.dsdf;
        const float px = nxt.px, py = nxt.py, pz = nxt.pz;
        {
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // chunks of 4 feature columns (2 levels each)
                const float2 f0 = __half22float2(nxt.f[2 * q]), f1 = __half22float2(nxt.f[2 * q + 1]);
                *reinterpret_cast<float4 *>(a1 + umma::kmajor_off_lbo<LBO>(row, 16 * half + 4 * q, kK1)) = make_float4(f0.x, f0.y, f1.x, f1.y);
            }
            if (half == 0) {
                const float xh = __uint_as_float(to_tf32(px)), yh = __uint_as_float(to_tf32(py)), zh = __uint_as_float(to_tf32(pz));
                const float xl = __uint_as_float(to_tf32(px - xh)), yl = __uint_as_float(to_tf32(py - yh)), zl = __uint_as_float(to_tf32(pz - zh));
                *reinterpret_cast<float4 *>(a1 + umma::kmajor_off_lbo<LBO>(row, KF, kK1)) = make_float4(xh, yh, zh, xl);
                *reinterpret_cast<float4 *>(a1 + umma::kmajor_off_lbo<LBO>(row, KF + 4, kK1)) = make_float4(yl, zl, 1.f, 0.f);
            }
        }
        fetch(q0 + (int64_t)gridDim.x * kTileQ, nxt);      // next tile's loads are in flight from here on
        umma::fence_smem_to_async_proxy();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();

        // ---- (1) Z = X1 B1^T
        if (tid == 0) {
            uint32_t acc = 0;
            for (int s = 0; s < kK1 / 8; ++s) {
                if (s < KF / 8 && s >= n_feat_steps) continue;   // all-zero feature columns
                const uint64_t ad = umma::kmajor_desc_lbo<LBO>(a1_addr + 2u * LBO * s, kK1);
                umma::mma_tf32(tmem, ad, umma::kmajor_desc(b1hi_addr + 256u * s, kK1), idesc_z, acc);
                umma::mma_tf32(tmem, ad, umma::kmajor_desc(b1lo_addr + 256u * s, kK1), idesc_z, 1);
                acc = 1;
            }
            umma::commit(&bar_z);
        }
        if (!umma::mbar_wait(&bar_z, phase)) failed = true;
        umma::fence_after_sync();

        // ---- (2) dz, dW1 / db1 for this thread's 32 hidden units
        {
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                const int h0 = 32 * half + c0;
                float z[16];
                umma::tmem_ld16(tmem_lane + (uint32_t)h0, z);
                umma::tmem_ld_wait();
                float dzv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float sp, sg;
                    softplus100_both_lg2(z[i], sp, sg);
                    accH[c0 + i] = fmaf(dsdf, sp, accH[c0 + i]);
                    dzv[i] = __uint_as_float(to_tf32(dsdf * s_w1[h0 + i] * sg));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4 *>(a2 + umma::kmajor_off_lbo<LBO>(row, h0 + 4 * q, 64)) = make_float4(dzv[4 * q], dzv[4 * q + 1], dzv[4 * q + 2], dzv[4 * q + 3]);
            }
            if (half == 0) {
                float ds = dsdf;
#pragma unroll
                for (int o = 16; o; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
                accB1 += ds;
            }
        }
        umma::fence_smem_to_async_proxy();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();

        // ---- (3) U = dz W0feat  (async on the tensor pipe)
        if (tid == 0) {
            for (int s = 0; s < 8; ++s)
                umma::mma_tf32(tmem + 64u, umma::kmajor_desc_lbo<LBO>(a2_addr + 2u * LBO * s, 64), umma::kmajor_desc(b2_addr + 256u * s, 64), idesc_u, s > 0);
            umma::commit(&bar_u);
        }

        // ---- (4) dW0T[col][h] += sum_p X1[p][col] dz[p][h]   (mma.sync; A = X1^T, B = dz read from the K-major tiles); hidden tile = warp
        {
            // fragment addresses in the core-matrix layout: point 8 ks + t (+4) -> ks * (K/4) * LBO + t * 16 (+64); column c -> (c / 4) * LBO + (c % 4) * 4.
            // The MMA's row slots (g, g + 8) of an m-tile carry the ADJACENT columns (2g, 2g + 1) of X1 (any bijection works as long as
            // the flush below undoes it), so a0 / a1 and a2 / a3 come out of one 64-bit load each.
            const uint8_t *ab = a1 + (g >> 1) * LBO + (g & 1) * 8 + t * 16;               // columns 2g, 2g + 1 of m-tile 0; m-tile mt: + mt * 4 * LBO
            const uint8_t *bb0 = a2 + (2 * warp + (g >> 2)) * LBO + (g & 3) * 4 + t * 16;   // hidden 8 warp + g
#pragma unroll 2
            for (int ks = 0; ks < 16; ++ks) {
                const uint8_t *ak = ab + ks * (kK1 / 4) * LBO, *bk = bb0 + ks * (64 / 4) * LBO;
                const uint32_t b00 = *reinterpret_cast<const uint32_t *>(bk), b01 = *reinterpret_cast<const uint32_t *>(bk + 64);
#pragma unroll
                for (int mt = 0; mt < kNMT; ++mt) {
                    if (16 * mt + 16 <= KF && 16 * mt >= 2 * (int)n_active) continue;   // feature-only tile beyond the active levels
                    const bool in_tile = 16 * mt + 2 * g < kK1;      // the last m-tile holds kK1 - 32 = 8 columns: row slots g >= 4 are empty
                    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
                    if (in_tile) {
                        lo = *reinterpret_cast<const uint2 *>(ak + mt * 4 * LBO);          // point k-slot t:     columns 2g, 2g + 1
                        hi = *reinterpret_cast<const uint2 *>(ak + mt * 4 * LBO + 64);     // point k-slot t + 4
                    }
                    const uint32_t a[4] = {lo.x, lo.y, hi.x, hi.y};
                    mma_tf32(wacc[mt], a, b00, b01);
                }
            }
        }

        // ---- hand d loss / d features (this half's 8 level slots) and the point to the scatter kernel
        if (!umma::mbar_wait(&bar_u, phase)) failed = true;
        umma::fence_after_sync();
        {
            const int64_t q = q0 + tj;
            if (half == 0 && valid) ws_pos[(int64_t)tk * qcap + q] = make_float4(px, py, pz, dsdf != 0.f ? 1.f : 0.f);
            if (8 * half < (int)n_active) {                           // warp-uniform
                float u[16];
                umma::tmem_ld16(tmem_lane + 64u + 16u * (uint32_t)half, u);
                umma::tmem_ld_wait();
#pragma unroll
                for (int l2 = 0; l2 < 8; ++l2) {
                    const int l = 8 * half + l2;
                    if (valid && l < (int)n_active) ws_dfeat[((int64_t)l * SNB_PATCH + tk) * qcap + q] = make_float2(u[2 * l2], u[2 * l2 + 1]);
                }
            }
        }
        phase ^= 1u;
        umma::fence_before_sync();
        __syncthreads();          // tiles a1 / a2 and both accumulators are free again
        umma::fence_after_sync();
    }

    // ---- flush
#pragma unroll
    for (int mt = 0; mt < kNMT; ++mt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int col = 16 * mt + 2 * g + (c >> 1), h = 8 * warp + 2 * t + (c & 1);   // row slots (g, g + 8) = columns (2g, 2g + 1)
            float *dst = nullptr;
            if (col < KF) { if (col < 2 * (int)n_active) dst = net_grad + kOffW0T + (3 + col) * kH + h; }
            else if (col < kColXlo) dst = net_grad + kOffW0T + (col - kColXhi) * kH + h;
            else if (col < kColOne) dst = net_grad + kOffW0T + (col - kColXlo) * kH + h;
            else if (col == kColOne) dst = net_grad + kOffB0 + h;
            if (dst && wacc[mt][c] != 0.f) atomicAdd(dst, wacc[mt][c]);
        }
    }
    warp_transpose_reduce32(accH, lane);
    atomicAdd(net_grad + kOffW1 + 32 * half + lane, accH[0]);
    if (half == 0 && lane == 0) atomicAdd(net_grad + kOffB1, accB1);
    if (failed && tid == 0 && err_flag) atomicExch(err_flag, 1);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kBwdTmemCols);
}

// ---------------------------------------------------------------------------------------------
// hash-table scatter of the split backward
// ---------------------------------------------------------------------------------------------
// table_grad[level][corner entry] += trilinear weight * d loss / d feature, for every (point, level, corner).  What bounds it is the
// L2 reduction path: ~97 reduction SECTORS (32 B) per clock chip-wide = 190 G/s at 1.965 GHz, independent of the operand width and of
// how many lanes of one instruction fall into the sector (scripts/micro/red_rate.cu, profiles/r02_red_rate_microbench.txt).  So:
//   * a warp takes 32 consecutive samples q of ONE in-patch ray (neighbours along the ray share grid cells on all but the finest
//     levels) and loops over the levels;
//   * writer role: lane = point; it computes the cell and its 8 weighted contributions (float2 each) and parks them with the cell
//     coordinates in shared memory;
//   * reader role: lane = (quarter of the warp, corner c); it walks the 8 points of its quarter, sums corner c's contribution while
//     the cell stays the same and issues ONE red.global.add.v2.f32 per run.  Coarse levels collapse to one reduction per corner and
//     quarter; on fine levels every point flushes, but one instruction then carries the 8 corners of 4 points in adjacent lanes and
//     x-neighbour corners (entry h and h ^ 1 / h ^ 3 for a hashed level, idx and idx + 1 for a dense one) share a 32-byte sector
//     in 3 of 4 cases -> 5 instead of 8 sector requests per point and level.
// Shared memory is double-buffered over the levels, so one __syncwarp per level suffices.
constexpr int kScatWarps = 8;
constexpr int kScatRow = 9;     // float2 per point row (8 corners + 1 pad): 18-word stride, conflict-free 64-bit accesses

__global__ void __launch_bounds__(32 * kScatWarps, 4) hash_scatter_kernel(LevelTable lt, uint32_t n_active, const int *__restrict__ totals, int64_t qcap,
                                                                          const float4 *__restrict__ ws_pos, const float2 *__restrict__ ws_dfeat,
                                                                          float *__restrict__ table_grad) {
    __shared__ __align__(16) float2 s_val[kScatWarps][2][32 * kScatRow];
    __shared__ __align__(16) uint4 s_cell[kScatWarps][2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qd8 = lane & ~7, cn = lane & 7;
    const int64_t Q = (int64_t)totals[0] + totals[1];
    const int64_t n_chunks = (Q + 31) >> 5, n_tasks = n_chunks * SNB_PATCH;
    for (int64_t task = (int64_t)blockIdx.x * kScatWarps + warp; task < n_tasks; task += (int64_t)gridDim.x * kScatWarps) {
        const int k = (int)(task % SNB_PATCH);
        const int64_t q = (task / SNB_PATCH) * 32 + lane;
        float4 pw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < Q) pw = __ldg(ws_pos + (int64_t)k * qcap + q);
        const bool live = pw.w != 0.f;
        const float2 *gp = ws_dfeat + (int64_t)k * qcap + q;
        float2 g_next = make_float2(0.f, 0.f);
        if (live && n_active > 0) g_next = __ldg(gp);
        for (uint32_t l = 0; l < n_active; ++l) {
            const float2 g = g_next;
            if (live && l + 1 < n_active) g_next = __ldg(gp + (int64_t)(l + 1) * SNB_PATCH * qcap);
            const LevelCtx c = lt.lv[l];
            const int buf = (int)(l & 1u);
            // ---- writer: this lane's point.  Run structure of the warp as ONE bit mask: bit p set <=> point p starts a run (its cell
            // differs from its left neighbour's, or it opens a quarter).  Dead lanes (g = 0) contribute zeros and may sit inside any run.
            const Cell cell = cell_of(c, pw.x, pw.y, pw.z);
            const uint32_t px = __shfl_up_sync(0xffffffffu, cell.g[0], 1), py = __shfl_up_sync(0xffffffffu, cell.g[1], 1),
                           pz = __shfl_up_sync(0xffffffffu, cell.g[2], 1);
            const bool head = (lane & 7) == 0 || cell.g[0] != px || cell.g[1] != py || cell.g[2] != pz;
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            s_cell[warp][buf][lane] = make_uint4(cell.g[0], cell.g[1], cell.g[2], 0u);
            float2 *row = &s_val[warp][buf][lane * kScatRow];
#pragma unroll
            for (uint32_t cc = 0; cc < 8; ++cc) {
                const float w = corner_weight(cell, cc);
                row[cc] = make_float2(w * g.x, w * g.y);
            }
            __syncwarp();
            // ---- reader: corner cn of the 8 points of this quarter; one reduction per run, issued at the run's LAST point (bit p + 1
            // of `heads` set, or end of the quarter) with that point's own cell -- every point of a run has the same one
            float2 *gt = reinterpret_cast<float2 *>(table_grad) + c.offset;
            const uint4 *cq = &s_cell[warp][buf][qd8];
            const float2 *vq = &s_val[warp][buf][qd8 * kScatRow + cn];
            const uint32_t tails = ((heads >> qd8) >> 1) | 0x80u;      // bit i: point i closes a run
            const uint32_t bx = cn & 1u, by = (cn >> 1) & 1u, bz = (uint32_t)cn >> 2;
            float ax = 0.f, ay = 0.f;
            if (c.hashed) {     // level-uniform: idx = x ^ y * 2654435761 ^ z * 805459861 mod size (size a power of two for every hashed level)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 v = vq[i * kScatRow];
                    ax += v.x; ay += v.y;
                    if ((tails >> i) & 1u) {
                        if (ax != 0.f || ay != 0.f) {
                            const uint4 cg = cq[i];
                            const uint32_t idx = (cg.x + bx) ^ ((cg.y + by) * 2654435761u) ^ ((cg.z + bz) * 805459861u);
                            atomicAdd(gt + level_mod(c, idx), make_float2(ax, ay));
                        }
                        ax = 0.f; ay = 0.f;
                    }
                }
            } else {            // dense level: the generic index (stride loop of tiny-cuda-nn's grid_index)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 v = vq[i * kScatRow];
                    ax += v.x; ay += v.y;
                    if ((tails >> i) & 1u) {
                        if (ax != 0.f || ay != 0.f) {
                            const uint4 cg = cq[i];
                            Cell kc;
                            kc.g[0] = cg.x; kc.g[1] = cg.y; kc.g[2] = cg.z;
                            atomicAdd(gt + corner_index(c, kc, (uint32_t)cn), make_float2(ax, ay));
                        }
                        ax = 0.f; ay = 0.f;
                    }
                }
            }
        }
        __syncwarp();   // the next task's first level reuses buffer 0 (or 1): every lane must be done reading
    }
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

// workspace of the split backward: [ ws_pos float4[9 * qcap] | ws_dfeat float2[n_levels * 9 * qcap] ],  qcap = capacity + end_capacity
extern "C" int64_t snb_sdf_bwd_workspace_bytes(int32_t n_levels, int64_t capacity, int64_t end_capacity) {
    if (n_levels < 1 || n_levels > SNB_MAX_LEVELS || capacity < 0 || end_capacity < 0) return -1;
    const int64_t qcap = capacity + end_capacity;
    return (int64_t)SNB_PATCH * qcap * (int64_t)(sizeof(float4) + (size_t)n_levels * sizeof(float2));
}

extern "C" int32_t snb_sdf_bwd_patch_ws(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const void *feats,
                                        const float *d_sdf0, const float *d_sdf1, float *table_grad, float *net_grad, void *workspace,
                                        int64_t workspace_bytes, snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_bwd_patch");
    if (rc) return rc;
    rc = check_patch_args(b, sm, "sdf_bwd_patch");
    if (rc) return rc;
    SNB_REQUIRE(feats && d_sdf0 && d_sdf1 && table_grad && net_grad, SNB_ERR_NULL, "sdf_bwd_patch: null buffer");
    SNB_REQUIRE(aligned(table_grad, 8), SNB_ERR_ALIGN, "sdf_bwd_patch: table_grad must be 8-byte aligned");
    SNB_REQUIRE(!workspace || aligned(workspace, 16), SNB_ERR_ALIGN, "sdf_bwd_patch: workspace must be 16-byte aligned");
    static const size_t smem = sizeof(float) * (kNetFloats + kTile * kDzStride + kTile * kXStride);
    // tcgen05 kernel (KF = 8 tiles up to 4 active levels, KF = 32 beyond): 140 vs 178 us at 1 level, 75 vs 113 us at 4, 370 vs 532 us
    // at 14 against the FMA kernel, which SNB_BWD_UMMA=0 still selects for cross-checks (tests/test_gpu_fused.py).  Beyond 4 levels the
    // scatter moves into its own high-occupancy kernel when the caller provides the workspace (SNB_BWD_SPLIT=0: keep it fused).
    static const int use_umma = getenv("SNB_BWD_UMMA") ? atoi(getenv("SNB_BWD_UMMA")) : 1;
    static const int use_split = getenv("SNB_BWD_SPLIT") ? atoi(getenv("SNB_BWD_SPLIT")) : 1;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_bwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(sdf_bwd_patch_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg<32>::kSmem);
        cudaFuncSetAttribute(sdf_bwd_mlp_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_smem_bytes<128>());
        cudaFuncSetAttribute(sdf_bwd_mlp_umma_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_smem_bytes<192>());
        cudaFuncSetAttribute(sdf_bwd_patch_umma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg<8>::kSmem);
        configured = true;
    }
    const LevelTable ltab = make_level_table(net->meta);
    const int64_t qcap = sm->capacity + sm->end_capacity;
    static const int split_min = getenv("SNB_BWD_SPLIT_MIN") ? atoi(getenv("SNB_BWD_SPLIT_MIN")) : 5;
    const bool split = use_umma && use_split && (int)net->n_active >= split_min && workspace &&
                       workspace_bytes >= snb_sdf_bwd_workspace_bytes((int32_t)net->meta.n_levels, sm->capacity, sm->end_capacity);
    if (split) {
        float4 *ws_pos = reinterpret_cast<float4 *>(workspace);
        float2 *ws_dfeat = reinterpret_cast<float2 *>(ws_pos + (int64_t)SNB_PATCH * qcap);
        static const int lbo = getenv("SNB_BWD_LBO") ? atoi(getenv("SNB_BWD_LBO")) : 192;
        if (lbo == 128)
            sdf_bwd_mlp_umma_kernel<128><<<kNumSMs * 2, kMlpThreads, mlp_smem_bytes<128>(), S(stream)>>>(
                *b, *net, *sm, (const __half2 *)feats, d_sdf0, d_sdf1, net_grad, sm->totals + 2, ws_dfeat, ws_pos, qcap);
        else
            sdf_bwd_mlp_umma_kernel<192><<<kNumSMs * 2, kMlpThreads, mlp_smem_bytes<192>(), S(stream)>>>(
                *b, *net, *sm, (const __half2 *)feats, d_sdf0, d_sdf1, net_grad, sm->totals + 2, ws_dfeat, ws_pos, qcap);
        SNB_LAUNCH_CHECK("sdf_bwd_patch (mlp)");
        hash_scatter_kernel<<<kNumSMs * 4, 32 * kScatWarps, 0, S(stream)>>>(ltab, net->n_active, sm->totals, qcap, ws_pos, ws_dfeat, table_grad);
        SNB_LAUNCH_CHECK("sdf_bwd_patch (scatter)");
        return SNB_OK;
    }
    if (use_umma && net->n_active <= 4)
        sdf_bwd_patch_umma_kernel<8><<<kNumSMs * BwdCfg<8>::kMinBlocks, 128, BwdCfg<8>::kSmem, S(stream)>>>(
            *b, *net, ltab, *sm, (const __half2 *)feats, d_sdf0, d_sdf1, table_grad, net_grad, sm->totals + 2);
    else if (use_umma)
        sdf_bwd_patch_umma_kernel<32><<<kNumSMs * BwdCfg<32>::kMinBlocks, 128, BwdCfg<32>::kSmem, S(stream)>>>(
            *b, *net, ltab, *sm, (const __half2 *)feats, d_sdf0, d_sdf1, table_grad, net_grad, sm->totals + 2);
    else
        sdf_bwd_patch_kernel<<<kNumSMs * 3, kTile, smem, S(stream)>>>(*b, *net, ltab, *sm, (const __half2 *)feats, d_sdf0, d_sdf1, table_grad, net_grad);
    SNB_LAUNCH_CHECK("sdf_bwd_patch");
    return SNB_OK;
}

#ifdef SNB_BWD_DEBUG
extern "C" int32_t snb_debug_bwd_phases(unsigned long long *host_out, int32_t reset) {
    int32_t rc = (int32_t)cudaMemcpyFromSymbol(host_out, g_bwd_dbg, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_bwd_dbg, z, sizeof(z)); }
    return rc;
}
#endif

extern "C" int32_t snb_sdf_bwd_patch(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const void *feats,
                                     const float *d_sdf0, const float *d_sdf1, float *table_grad, float *net_grad,
                                     snb_stream_t stream) {
    return snb_sdf_bwd_patch_ws(b, net, sm, feats, d_sdf0, d_sdf1, table_grad, net_grad, nullptr, 0, stream);
}

/* gradient_method = 'ad': analytic SDF gradient at every sample start of every in-patch ray (points p = 9 s + k < 9 S).  grad: f32 [9 * capacity, 3] */
extern "C" int32_t snb_sdf_grad_patch(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, float *grad, snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_grad_patch");
    if (rc) return rc;
    rc = check_patch_args(b, sm, "sdf_grad_patch");
    if (rc) return rc;
    SNB_REQUIRE(grad, SNB_ERR_NULL, "sdf_grad_patch: null output");
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_grad_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemBytes);
        configured = true;
    }
    sdf_grad_patch_kernel<<<kNumSMs * 2, 32 * kFwdWarps, kGradSmemBytes, S(stream)>>>(*b, *net, make_level_table(net->meta), *sm, grad);
    SNB_LAUNCH_CHECK("sdf_grad_patch");
    return SNB_OK;
}

/* ... and its backward (the double backward of the reference's create_graph = True route): d_grad = d loss / d grad of snb_sdf_grad_patch;
 * table_grad / net_grad += like snb_sdf_bwd_patch (net_grad w.r.t. the FOLDED weights).  feats: the rows snb_sdf_fwd_patch kept. */
extern "C" int32_t snb_sdf_grad_bwd_patch(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const void *feats,
                                          const float *d_grad, float *table_grad, float *net_grad, snb_stream_t stream) {
    int32_t rc = check_net(net, "sdf_grad_bwd_patch");
    if (rc) return rc;
    rc = check_patch_args(b, sm, "sdf_grad_bwd_patch");
    if (rc) return rc;
    SNB_REQUIRE(feats && d_grad && table_grad && net_grad, SNB_ERR_NULL, "sdf_grad_bwd_patch: null buffer");
    SNB_REQUIRE(aligned(table_grad, 8), SNB_ERR_ALIGN, "sdf_grad_bwd_patch: table_grad must be 8-byte aligned");
    static const size_t smem = sizeof(float) * (kNetFloats + kTile * kDzStride + kTile * kXStride);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_grad_bwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    sdf_grad_bwd_patch_kernel<<<kNumSMs * 3, kTile, smem, S(stream)>>>(*b, *net, make_level_table(net->meta), *sm, (const __half2 *)feats, d_grad,
                                                                      table_grad, net_grad);
    SNB_LAUNCH_CHECK("sdf_grad_bwd_patch");
    return SNB_OK;
}
