// sdf_core.cuh -- per-thread fused "hash-grid encode -> 64-wide SDF MLP" used by every fused kernel.
//
// Restates models/fields.py:76-99 (SDFNetwork.forward): x_in = [x, enc(x)*level_mask];
// h = softplus_100(W0 x_in + b0); sdf = W1 h + b1, with W = g*v/||v|| (weight_norm) folded on the
// device once per step by prep_net (optim.cu).  The encoding half is fp16-faithful tiny-cuda-nn
// (hashgrid.cuh); the MLP is fp32 like the reference's torch.nn.Linear.
//
// Layout of the folded network ("net" buffer, floats), input-major so that one input feature's 64
// weights are a contiguous, warp-broadcast float4 stream from shared memory:
//   W0T[DIN_MAX][64] | b0[64] | W1[64] | b1 | inv_s | pad
#pragma once
#include "hashgrid.cuh"

namespace snb {

constexpr int kH = SNB_HIDDEN;          // 64
constexpr int kDinMax = 3 + 2 * SNB_MAX_LEVELS;  // 35
constexpr int kOffW0T = 0;
constexpr int kOffB0 = kDinMax * kH;    // 2240
constexpr int kOffW1 = kOffB0 + kH;     // 2304
constexpr int kOffB1 = kOffW1 + kH;     // 2368
constexpr int kOffInvS = kOffB1 + 1;    // 2369
constexpr int kNetFloats = 2432;        // padded to a multiple of 64

__device__ __forceinline__ void load_net_to_smem(float *s_net, const float *__restrict__ g_net) {
    const float4 *src = reinterpret_cast<const float4 *>(g_net);
    float4 *dst = reinterpret_cast<float4 *>(s_net);
    for (int i = threadIdx.x; i < kNetFloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// softplus(beta=100, threshold=20) and its derivative sigmoid(100 z)   (models/fields.py:70), in the overflow-free form
// softplus(z) = max(z, 0) + log1p(exp(-|100 z|)) / 100 = max(z, 0) + (ln 2 / 100) lg2(1 + u),  u = ex2(-|100 z| / ln 2):
// MUFU.EX2 + FADD + MUFU.LG2 + FFMA (+ one MUFU.RCP for the derivative).  Absolute error of softplus ~2e-9 (lg2.approx: 2^-22 absolute
// near 1; the hidden activations are O(0.1)), relative error of the derivative ~2^-21: far below the fp32 noise of the 64-term dot
// products that consume them.
// History: round 1 replaced the LG2 by a degree-7 polynomial log1p (7 FMAs) because the XU pipe (16 lanes/clk/SM) was the scarce one
// in the thread-per-point FMA kernels of the time, and voted per warp to skip saturated units (|100 z| >= 17).  With the .ftz MUFU forms
// and the tensor-core layer 0 the issue slots are the scarce resource, and the vote never fires on the diligent schedule (exactly 64
// EX2 per point in every profile, profiles/r02_sass_hist_step_it4800.txt): LG2 without the vote is 3 % of the step faster at
// iterations 15 / 1000, 1.6 % at 4800 (profiles/r02_softplus_variants.txt).

// MUFU ops as their .ftz PTX forms (see softplus100_both_lg2 below for why)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float softplus100_tail(float u, float z) { return fmaf(0.0069314718056f, lg2_ftz(1.f + u), fmaxf(z, 0.f)); }
__device__ __forceinline__ float softplus100(float z) { return softplus100_tail(ex2_ftz(fabsf(z) * -144.26950408889634f), z); }
// softplus and its derivative sigmoid(100 z): one EX2 and one RCP
__device__ __forceinline__ void softplus100_both(float z, float &sp, float &sg) {
    const float u = ex2_ftz(fabsf(z) * -144.26950408889634f);     // the same u as softplus100: the two forms agree bit for bit on sp
    sp = softplus100_tail(u, z);
    const float r = rcp_ftz(1.f + u);
    sg = z >= 0.f ? r : u * r;
}

// the same pair with the logarithm on the XU pipe: log1p(u) = -ln(r), r = 1 / (1 + u) = sigmoid(|100 z|) in [0.5, 1], so
// softplus = max(z, 0) - (ln 2 / 100) lg2(r): EX2 + RCP + LG2 and 7 FMAs fewer than the polynomial.  For kernels whose issue slots, not
// the XU pipe, are the scarce resource (the 256-thread MLP backward: 19 % XU utilisation).  MUFU.LG2's absolute error near r = 1 is
// 2^-22, i.e. 1.6e-9 on softplus -- the class of the polynomial form.  The three MUFU ops are issued as their .ftz PTX forms: the
// non-ftz intrinsics (__expf, __log2f, __fdividef) wrap every MUFU in a denormal rescue (FSETP + FMUL by 2^24 + FSEL + FADD -24, plus
// predicate spills through LOP3) that tripled the per-value instruction count (profiles/r02_ncu_bwd_split_v4_it4800.txt); here u
// underflowing to 0 is exact (|100 z| > 87: softplus = max(z, 0), derivative 0 / 1) and r is never denormal.
__device__ __forceinline__ void softplus100_both_lg2(float z, float &sp, float &sg) {
    const float u = ex2_ftz(fabsf(z) * -144.26950408889634f);     // exp(-|100 z|)
    const float r = rcp_ftz(1.f + u);
    sp = fmaf(-0.0069314718056f, lg2_ftz(r), fmaxf(z, 0.f));
    sg = z >= 0.f ? r : u * r;                                     // sigmoid(100 z)
}

__device__ __forceinline__ void rank1_update(float (&acc)[kH], const float *__restrict__ w_row, float v) {
    const float4 *w = reinterpret_cast<const float4 *>(w_row);
#pragma unroll
    for (int q = 0; q < kH / 4; ++q) {
        float4 t = w[q];
        acc[4 * q + 0] = fmaf(t.x, v, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(t.y, v, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(t.z, v, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(t.w, v, acc[4 * q + 3]);
    }
}

// Pre-activations z[64] of layer 0 for one point.  feat_row (optional, global, half2[L]) receives the
// encoded features of the active levels so the backward can skip the gathers.
template <bool SAVE_FEAT, bool LOAD_FEAT>
__device__ __forceinline__ void layer0(float x, float y, float z, const __half2 *__restrict__ table,
                                       const LevelCtx *s_lvl, uint32_t n_active, const float *s_net,
                                       __half2 *feat_row, float (&acc)[kH]) {
    const float4 *b0 = reinterpret_cast<const float4 *>(s_net + kOffB0);
#pragma unroll
    for (int q = 0; q < kH / 4; ++q) {
        float4 t = b0[q];
        acc[4 * q] = t.x; acc[4 * q + 1] = t.y; acc[4 * q + 2] = t.z; acc[4 * q + 3] = t.w;
    }
    rank1_update(acc, s_net + kOffW0T + 0 * kH, x);
    rank1_update(acc, s_net + kOffW0T + 1 * kH, y);
    rank1_update(acc, s_net + kOffW0T + 2 * kH, z);
    for (uint32_t l = 0; l < n_active; ++l) {   // (unrolling by 2 for more gathers in flight spills: 64 accumulators per thread)
        __half2 f;
        if (LOAD_FEAT) {
            f = feat_row[l];
        } else {
            const LevelCtx c = s_lvl[l];
            Cell cell = cell_of(c, x, y, z);
            f = interp_level(c, cell, table);
            if (SAVE_FEAT) feat_row[l] = f;
        }
        float2 ff = __half22float2(f);
        rank1_update(acc, s_net + kOffW0T + (3 + 2 * l) * kH, ff.x);
        rank1_update(acc, s_net + kOffW0T + (4 + 2 * l) * kH, ff.y);
    }
}

__device__ __forceinline__ float layer1(const float (&acc)[kH], const float *s_net) {
    float s = s_net[kOffB1];
    const float4 *w1 = reinterpret_cast<const float4 *>(s_net + kOffW1);
#pragma unroll
    for (int q = 0; q < kH / 4; ++q) {
        float4 t = w1[q];
        s = fmaf(t.x, softplus100(acc[4 * q + 0]), s);
        s = fmaf(t.y, softplus100(acc[4 * q + 1]), s);
        s = fmaf(t.z, softplus100(acc[4 * q + 2]), s);
        s = fmaf(t.w, softplus100(acc[4 * q + 3]), s);
    }
    return s;
}

template <bool SAVE_FEAT>
__device__ __forceinline__ float sdf_point(float x, float y, float z, const __half2 *__restrict__ table,
                                           const LevelCtx *s_lvl, uint32_t n_active, const float *s_net,
                                           __half2 *feat_row) {
    float acc[kH];
    layer0<SAVE_FEAT, false>(x, y, z, table, s_lvl, n_active, s_net, feat_row, acc);
    return layer1(acc, s_net);
}

// NeuS opacity of an interval from the SDF at its two ends (models/renderer.py:173-179)
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float neus_alpha(float s0, float s1, float inv_s) {
    float c = sigmoidf_(s0 * inv_s), n = sigmoidf_(s1 * inv_s);
    float a = (c - n + 1e-5f) / (c + 1e-5f);
    return fminf(fmaxf(a, 0.f), 1.f);
}

}  // namespace snb
