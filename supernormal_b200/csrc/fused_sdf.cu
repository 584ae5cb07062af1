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
                                                                 float *__restrict__ table_grad, float *__restrict__ net_grad, int dbg) {
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
        if (valid && dsdf != 0.f && !(dbg & 2)) {
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
                    if (!(dbg & 1)) atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g0, w * g1));
                    else if (w * g0 == 123.f) atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g0, w * g1));
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
// backward on the tensor cores
// ---------------------------------------------------------------------------------------------
// All three contractions of the MLP backward are mma.sync m16n8k8 TF32 with fp32 accumulation and hi/lo-split fp32
// operands (products exact to 2^-22, see sdf_mma.cuh):
//   (1) z  = X W0          warp-local [32 pts x K] x [K x 64]          (recompute from the kept fp16 features)
//   (2) u  = dz W0feat^T   warp-local [32 pts x 64] x [64 x 2L]        (d loss / d features -> table scatter)
//   (3) dW0 += X^T dz      CTA-wide   [K x P] x [P x 64], P = 32 * warps points per round; the 3 x 8 output tiles are
//                          dealt to the warps, so the persistent accumulators cost 8 registers per thread instead of 96.
// dz is handed from the accumulator layout of (1) to the operand layouts of (2)/(3) through a shared-memory tile.
constexpr int kDzS = 72;        // floats per point row of the dz tile

struct BwdSmem {
    float *net, *whi, *wlo, *xs, *dz;
    LevelCtx *lvl;
};

__global__ void __launch_bounds__(384, 1) sdf_bwd_patch_mma_kernel(snb_patch_batch b, snb_net net, LevelTable lt, snb_samples sm,
                                                                   const __half2 *__restrict__ feats,
                                                                   const float *__restrict__ d_sdf0,
                                                                   const float *__restrict__ d_sdf1,
                                                                   float *__restrict__ table_grad, float *__restrict__ net_grad, int dbg) {
    extern __shared__ __align__(16) float smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int P = 32 * nwarps;                       // points per CTA round
    BwdSmem sh;
    sh.net = smem;
    sh.whi = sh.net + kNetFloats;
    sh.wlo = sh.whi + kWRows * kWStride;
    sh.xs = sh.wlo + kWRows * kWStride;              // [P][kXsStride] (+16 floats of slack for the transposed over-read)
    sh.dz = sh.xs + P * kXsStride + 16;              // [P][kDzS]
    sh.lvl = reinterpret_cast<LevelCtx *>(sh.dz + P * kDzS);
    load_net_to_smem(sh.net, net.net);
    stage_w_split(sh.net, sh.whi, sh.wlo);
    if (threadIdx.x < SNB_MAX_LEVELS) sh.lvl[threadIdx.x] = lt.lv[threadIdx.x];
    __syncthreads();

    const int S = sm.totals[0], E = sm.totals[1];
    const int64_t M = (int64_t)SNB_PATCH * (S + E);
    const uint32_t L = net.meta.n_levels, n_active = net.n_active;
    const int ksteps = (int)(2 * n_active + 7) >> 3;          // feature k-steps of (1) == feature n-tiles of (2)
    float *xs = sh.xs + warp * 32 * kXsStride;                 // this warp's rows of the staging tile
    float *dzt = sh.dz + warp * 32 * kDzS;

    // (3): output tiles owned by this warp.  Tile id -> (mt, nt): mt 0,1 = feature columns 0..15 / 16..31, mt 2 = x y z 1.
    const int n_mt_feat = (int)(2 * n_active + 15) >> 4;
    const int n_tiles = (n_mt_feat + 1) * 8;
    float wacc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) wacc[i][c] = 0.f;
    float accW1[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) accW1[nt][0] = accW1[nt][1] = 0.f;
    float accB1 = 0.f;

    for (int64_t round0 = (int64_t)blockIdx.x * P; round0 < M; round0 += (int64_t)gridDim.x * P) {
        // ---- stage the lane's point: kept features, position, dsdf
        const int64_t p = round0 + warp * 32 + lane;
        const bool valid = p < M;
        float dsdf = 0.f;
        PointRef r;
        r.px = r.py = r.pz = 0.f;
        float *row = xs + lane * kXsStride;
        if (valid) {
            r = decode_point(p, S, b, sm);
            if (!r.is_end) {
                dsdf = __ldg(d_sdf0 + (int64_t)r.s * SNB_PATCH + r.k);
                // this start also served as the previous interval's end when that interval had no own end query
                if (r.s > 0 && __ldg(sm.end_slot + r.s - 1) < 0) dsdf += __ldg(d_sdf1 + (int64_t)(r.s - 1) * SNB_PATCH + r.k);
            } else {
                dsdf = __ldg(d_sdf1 + (int64_t)r.s * SNB_PATCH + r.k);
            }
            const __half2 *fr = feats + p * L;
            for (uint32_t l = 0; l < n_active; ++l) *reinterpret_cast<float2 *>(row + 2 * l) = __half22float2(fr[l]);
        } else {
            for (uint32_t l = 0; l < n_active; ++l) *reinterpret_cast<float2 *>(row + 2 * l) = make_float2(0.f, 0.f);
        }
        for (int c = 2 * (int)n_active; c < 16 * n_mt_feat; ++c) row[c] = 0.f;
        stage_point(xs, lane, r.px, r.py, r.pz);
        row[kXsTmp] = dsdf;
        __syncwarp();

        // ---- (1) z, then dz = dsdf * W1 * sigmoid(100 z) -> shared tile; dW1 / db1 partial sums stay in registers
        {
            float acc[2][8][4];
            warp_layer0_mma(acc, xs, sh.net, sh.whi, sh.wlo, ksteps, lane);
            float ds[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) ds[rr] = xs[(8 * rr + g) * kXsStride + kXsTmp];
            if (t == 0) accB1 += (ds[0] + ds[1]) + (ds[2] + ds[3]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 w1 = *reinterpret_cast<const float2 *>(sh.net + kOffW1 + 8 * nt + 2 * t);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    float sp0, sg0, sp1, sg1;
                    softplus100_both(acc[rr >> 1][nt][2 * (rr & 1)], sp0, sg0);
                    softplus100_both(acc[rr >> 1][nt][2 * (rr & 1) + 1], sp1, sg1);
                    accW1[nt][0] = fmaf(ds[rr], sp0, accW1[nt][0]);
                    accW1[nt][1] = fmaf(ds[rr], sp1, accW1[nt][1]);
                    *reinterpret_cast<float2 *>(dzt + (8 * rr + g) * kDzS + 8 * nt + 2 * t) = make_float2(ds[rr] * w1.x * sg0, ds[rr] * w1.y * sg1);
                }
            }
        }
        __syncwarp();

        // ---- (2) u = dz W0feat^T  (rows: points, cols: feature pairs) and the table scatter straight from the fragments
        if (ksteps > 0 && !(dbg & 2)) {
            float u[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int c = 0; c < 4; ++c) u[mt][nt][c] = 0.f;
#pragma unroll 2
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float *r0 = dzt + (16 * mt + g) * kDzS + 8 * ks + t;
                    const float v[4] = {r0[0], r0[8 * kDzS], r0[4], r0[8 * kDzS + 4]};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ahi[mt][e] = to_tf32(v[e]);
                        alo[mt][e] = to_tf32(v[e] - __uint_as_float(ahi[mt][e]));
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    if (nt < ksteps) {
                        // B[k = h][n = feature j] = W0T[3 + j][h]
                        const float *wh = sh.whi + (8 * nt + g) * kWStride + 8 * ks + t, *wl = sh.wlo + (8 * nt + g) * kWStride + 8 * ks + t;
                        const uint32_t h0 = __float_as_uint(wh[0]), h1 = __float_as_uint(wh[4]);
                        const uint32_t l0 = __float_as_uint(wl[0]), l1 = __float_as_uint(wl[4]);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            mma_tf32(u[mt][nt], ahi[mt], h0, h1);
                            mma_tf32(u[mt][nt], alo[mt], h0, h1);
                            mma_tf32(u[mt][nt], ahi[mt], l0, l1);
                        }
                    }
                }
            }
            // thread holds, for points 8 rr + g, the feature pair of level 4 nt + t
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const float *prow = xs + (8 * rr + g) * kXsStride;
                if (prow[kXsTmp] == 0.f) continue;       // dsdf == 0 (also: rows past M)
                const float px = prow[kXsXyz], py = prow[kXsXyz + 1], pz = prow[kXsXyz + 2];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const uint32_t l = 4 * nt + t;
                    if (l < n_active) {
                        const float g0 = u[rr >> 1][nt][2 * (rr & 1)], g1 = u[rr >> 1][nt][2 * (rr & 1) + 1];
                        const LevelCtx c = sh.lvl[l];
                        Cell cell = cell_of(c, px, py, pz);
                        float2 *gt = reinterpret_cast<float2 *>(table_grad) + c.offset;
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k) {
                            float w = corner_weight(cell, k);
                            if (!(dbg & 1)) atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g0, w * g1));
                            else if (w * g0 == 123.f) atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g0, w * g1));
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (dbg & 4) continue;

        // ---- (3) dW0T[col][h] += sum over the round's P points of x[p][col] * dz[p][h]
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int tile = warp + i * nwarps;
            if (tile < n_tiles) {
                const int mtf = tile >> 3, nt = tile & 7;
                const bool xyz = mtf == n_mt_feat;                       // last m-tile: x y z 1 (fp32, needs its own hi/lo split)
                const int col0 = xyz ? kXsXyz : 16 * mtf;
                for (int ks = 0; ks < 4 * nwarps; ++ks) {
                    // A[m = column][k = point] read transposed from the staging tile; B[k = point][n = h] from the dz tile
                    const float *xa = sh.xs + (8 * ks + t) * kXsStride + col0 + g;
                    const float av[4] = {xa[0], xa[8], xa[4 * kXsStride], xa[4 * kXsStride + 8]};
                    const float *db = sh.dz + (8 * ks + t) * kDzS + 8 * nt + g;
                    const float bv0 = db[0], bv1 = db[4 * kDzS];
                    uint32_t a[4], bh0 = to_tf32(bv0), bh1 = to_tf32(bv1);
                    const uint32_t bl0 = to_tf32(bv0 - __uint_as_float(bh0)), bl1 = to_tf32(bv1 - __uint_as_float(bh1));
#pragma unroll
                    for (int e = 0; e < 4; ++e) a[e] = xyz ? to_tf32(av[e]) : __float_as_uint(av[e]);
                    mma_tf32(wacc[i], a, bh0, bh1);
                    mma_tf32(wacc[i], a, bl0, bl1);
                    if (xyz) {
                        uint32_t al[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) al[e] = to_tf32(av[e] - __uint_as_float(a[e]));
                        mma_tf32(wacc[i], al, bh0, bh1);
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- flush
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int tile = warp + i * nwarps;
        if (tile < n_tiles) {
            const int mtf = tile >> 3, nt = tile & 7;
            const bool xyz = mtf == n_mt_feat;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int rowi = g + 8 * (c >> 1), h = 8 * nt + 2 * t + (c & 1);
                float *dst = nullptr;
                if (xyz) {
                    if (rowi < 3) dst = net_grad + kOffW0T + rowi * kH + h;
                    else if (rowi == 3) dst = net_grad + kOffB0 + h;
                } else {
                    const int j = 16 * mtf + rowi;
                    if (j < 2 * (int)n_active) dst = net_grad + kOffW0T + (3 + j) * kH + h;
                }
                if (dst) atomicAdd(dst, wacc[i][c]);
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float v = accW1[nt][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) atomicAdd(net_grad + kOffW1 + 8 * nt + 2 * t + e, v);
        }
    accB1 += __shfl_xor_sync(0xffffffffu, accB1, 4);
    accB1 += __shfl_xor_sync(0xffffffffu, accB1, 8);
    accB1 += __shfl_xor_sync(0xffffffffu, accB1, 16);
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
    static const int dbg = getenv("SNB_BWD_DBG") ? atoi(getenv("SNB_BWD_DBG")) : 0;   // timing experiments only: 1 no atomics, 2 no phase (2), 4 no phase (3)
    static const int bwd_warps = getenv("SNB_BWD_WARPS") ? atoi(getenv("SNB_BWD_WARPS")) : 12;   // 0: scalar-FMA kernel (cross-checks)
    static const size_t smem_mma = sizeof(float) * (kNetFloats + 2 * kWRows * kWStride + 32 * bwd_warps * (kXsStride + kDzS) + 16) + sizeof(LevelCtx) * SNB_MAX_LEVELS;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sdf_bwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(sdf_bwd_patch_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
        configured = true;
    }
    if (bwd_warps > 0) {
        SNB_REQUIRE(bwd_warps <= 12 && bwd_warps >= 2, SNB_ERR_ARG, "sdf_bwd_patch: SNB_BWD_WARPS must be in [2, 12]");
        const int ctas_per_sm = smem_mma <= 110 * 1024 ? 2 : 1;
        sdf_bwd_patch_mma_kernel<<<kNumSMs * ctas_per_sm, 32 * bwd_warps, smem_mma, S(stream)>>>(*b, *net, make_level_table(net->meta), *sm, (const __half2 *)feats, d_sdf0, d_sdf1, table_grad, net_grad, dbg);
    } else {
        sdf_bwd_patch_kernel<<<kNumSMs * 3, kTile, smem, S(stream)>>>(*b, *net, make_level_table(net->meta), *sm, (const __half2 *)feats, d_sdf0, d_sdf1, table_grad, net_grad, dbg);
    }
    SNB_LAUNCH_CHECK("sdf_bwd_patch");
    return SNB_OK;
}
