// sdf_mma.cuh -- the 64-wide SDF MLP as a warp-cooperative tensor-core contraction (north star item 3).
//
// models/fields.py:88-99: z = W0 [x, enc(x)] + b0 (31..35 -> 64), h = softplus_100(z), sdf = W1 h + b1.  A warp owns 32
// points.  Each lane gathers and interpolates the hash-grid features of ITS point (that part is per-point gather work),
// drops them into a shared-memory tile, and the feature part of layer 0 -- the only part that is a real contraction,
// [32 points x 2L] x [2L x 64] -- runs on the tensor cores as mma.sync.m16n8k8 TF32 with fp32 accumulation:
//   * features are fp16 values (tiny-cuda-nn's output type), hence EXACT in TF32 (same 10-bit mantissa);
//   * the fp32 weights are split W = hi + lo (two TF32 terms, 2^-22 relative), two MMAs per tile, so the products are
//     exact and only fp32 accumulation rounding remains -- the plain-fp32 accuracy the finite-difference normals need
//     (a single TF32 pass would put ~1e-4 absolute noise on the SDF, i.e. ~0.1 rad on a dfd normal);
//   * the three position columns and the bias stay fp32 FMAs (positions are not TF32-representable).
// The accumulator fragment leaves every thread with 16 hidden units of 4 points; softplus, layer 1 and (backward) the
// d z computation are elementwise / row-reductions in that layout, so nothing is transposed back.
//
// Fragment layouts (PTX ISA, mma.m16n8k8 .tf32; g = lane >> 2, t = lane & 3):
//   A 16x8 : a0 (g, t)   a1 (g+8, t)   a2 (g, t+4)   a3 (g+8, t+4)
//   B  8x8 : b0 (k = t, n = g)         b1 (k = t+4, n = g)
//   C 16x8 : c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
#pragma once
#include "sdf_core.cuh"

namespace snb {

constexpr int kXsStride = 40;   // floats per point row of the staging tile: 32 feature columns | x y z 1 | 4 spare
constexpr int kXsXyz = 32;      // columns 32..34: position, 35: constant 1 (bias column of the weight-gradient GEMM)
constexpr int kXsTmp = 36;      // scratch column (per-point scalars travelling between fragment and lane layouts)
constexpr int kWStride = 72;    // floats per feature row of the split weight tiles (conflict-free B-fragment loads)
constexpr int kWRows = 2 * SNB_MAX_LEVELS;

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Split the feature rows of the folded layer-0 weights (s_net: W0T[in][64], rows 3..) into TF32 hi / lo tiles.
__device__ __forceinline__ void stage_w_split(const float *s_net, float *s_whi, float *s_wlo) {
    for (int e = threadIdx.x; e < kWRows * kH; e += blockDim.x) {
        const int k = e / kH, n = e % kH;
        const float w = s_net[kOffW0T + (3 + k) * kH + n];
        const float hi = __uint_as_float(to_tf32(w));
        s_whi[k * kWStride + n] = hi;
        s_wlo[k * kWStride + n] = __uint_as_float(to_tf32(w - hi));
    }
    __syncthreads();
}

// Lane `lane` publishes its point: feature columns [0, 8*ksteps), position, the constant 1.
__device__ __forceinline__ void stage_point(float *xs, int lane, float x, float y, float z) {
    float *row = xs + lane * kXsStride;
    row[kXsXyz] = x; row[kXsXyz + 1] = y; row[kXsXyz + 2] = z; row[kXsXyz + 3] = 1.f;
}

// z[32 points x 64] for the warp's tile, in accumulator-fragment layout acc[mt][nt][c]:
// point row = 16 mt + g + 8 (c >> 1), hidden column = 8 nt + 2 t + (c & 1).
__device__ __forceinline__ void warp_layer0_mma(float (&acc)[2][8][4], const float *xs, const float *s_net, const float *s_whi,
                                                const float *s_wlo, int ksteps, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // bias + position part in fp32 (same operation order as the scalar path: b0, then x, y, z)
    float px[4], py[4], pz[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float *row = xs + (8 * r + g) * kXsStride + kXsXyz;   // rows g, g+8, g+16, g+24  <->  (mt, c>>1) = (r>>1, r&1)
        px[r] = row[0]; py[r] = row[1]; pz[r] = row[2];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = 8 * nt + 2 * t;
        const float2 b = *reinterpret_cast<const float2 *>(s_net + kOffB0 + col);
        const float2 wx = *reinterpret_cast<const float2 *>(s_net + kOffW0T + 0 * kH + col);
        const float2 wy = *reinterpret_cast<const float2 *>(s_net + kOffW0T + 1 * kH + col);
        const float2 wz = *reinterpret_cast<const float2 *>(s_net + kOffW0T + 2 * kH + col);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float v0 = fmaf(wx.x, px[r], b.x), v1 = fmaf(wx.y, px[r], b.y);
            v0 = fmaf(wy.x, py[r], v0); v1 = fmaf(wy.y, py[r], v1);
            v0 = fmaf(wz.x, pz[r], v0); v1 = fmaf(wz.y, pz[r], v1);
            acc[r >> 1][nt][2 * (r & 1)] = v0;
            acc[r >> 1][nt][2 * (r & 1) + 1] = v1;
        }
    }
    // feature part on the tensor cores: exact TF32 features x (hi + lo) weights
    for (int ks = 0; ks < ksteps; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float *r0 = xs + (16 * mt + g) * kXsStride + 8 * ks + t;
            a[mt][0] = __float_as_uint(r0[0]);
            a[mt][1] = __float_as_uint(r0[8 * kXsStride]);
            a[mt][2] = __float_as_uint(r0[4]);
            a[mt][3] = __float_as_uint(r0[8 * kXsStride + 4]);
        }
        const float *wh = s_whi + (8 * ks + t) * kWStride + g, *wl = s_wlo + (8 * ks + t) * kWStride + g;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uint32_t h0 = __float_as_uint(wh[8 * nt]), h1 = __float_as_uint(wh[4 * kWStride + 8 * nt]);
            const uint32_t l0 = __float_as_uint(wl[8 * nt]), l1 = __float_as_uint(wl[4 * kWStride + 8 * nt]);
            mma_tf32(acc[0][nt], a[0], h0, h1);
            mma_tf32(acc[1][nt], a[1], h0, h1);
            mma_tf32(acc[0][nt], a[0], l0, l1);
            mma_tf32(acc[1][nt], a[1], l0, l1);
        }
    }
}

// sdf of the lane's own point from the fragment-layout pre-activations (softplus + layer 1 + row reduction).
__device__ __forceinline__ float warp_layer1(const float (&acc)[2][8][4], float *xs, const float *s_net, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const float2 w1 = *reinterpret_cast<const float2 *>(s_net + kOffW1 + 8 * nt + 2 * t);
        // all 8 softplus of the group branch-free, back to back, so their MUFU latencies overlap
        float sp[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            sp[2 * r] = softplus100(acc[r >> 1][nt][2 * (r & 1)]);
            sp[2 * r + 1] = softplus100(acc[r >> 1][nt][2 * (r & 1) + 1]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            part[r] = fmaf(w1.x, sp[2 * r], part[r]);
            part[r] = fmaf(w1.y, sp[2 * r + 1], part[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        part[r] += __shfl_xor_sync(0xffffffffu, part[r], 1);
        part[r] += __shfl_xor_sync(0xffffffffu, part[r], 2);
    }
    __syncwarp();
    if (t == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) xs[(8 * r + g) * kXsStride + kXsTmp] = part[r];
    }
    __syncwarp();
    return xs[lane * kXsStride + kXsTmp] + s_net[kOffB1];
}

// Encode the lane's point (gathers + fp16-faithful interpolation), publish it, and evaluate the MLP for the whole warp.
// All 32 lanes must call this together; lanes without a point pass valid = false (their result is garbage, never NaN-safe).
template <bool SAVE_FEAT>
__device__ __forceinline__ float warp_sdf_mma(bool valid, float x, float y, float z, const __half2 *__restrict__ table,
                                              const LevelCtx *lvl, uint32_t n_active, const float *s_net, const float *s_whi,
                                              const float *s_wlo, float *xs, __half2 *feat_row, int lane) {
    const int ksteps = (int)(2 * n_active + 7) >> 3;
    float *row = xs + lane * kXsStride;
    __syncwarp();
    // two levels per iteration: 16 independent table gathers in flight per lane instead of 8 (the kernel is gather-latency bound:
    // 29 % of its stall samples sit on the HFMA2 that consumes the loads, profiles/r02_sass_hist_step_it4800.txt)
#pragma unroll 2
    for (uint32_t l = 0; l < n_active; ++l) {
        float2 ff = make_float2(0.f, 0.f);
        if (valid) {
            const LevelCtx c = lvl[l];
            Cell cell = cell_of(c, x, y, z);
            __half2 f = interp_level(c, cell, table);
            if (SAVE_FEAT) feat_row[l] = f;
            ff = __half22float2(f);
        }
        *reinterpret_cast<float2 *>(row + 2 * l) = ff;
    }
    for (int cidx = 2 * (int)n_active; cidx < 8 * ksteps; ++cidx) row[cidx] = 0.f;
    stage_point(xs, lane, valid ? x : 0.f, valid ? y : 0.f, valid ? z : 0.f);
    __syncwarp();
    float acc[2][8][4];
    warp_layer0_mma(acc, xs, s_net, s_whi, s_wlo, ksteps, lane);
    return warp_layer1(acc, xs, s_net, lane);
}

// ---------------------------------------------------------------------------------------------
// SDF and its analytic gradient in ONE pass (north star item 3): forward-mode through the MLP.
//   z = W0 [x, f(x)] + b0,  sdf = W1 softplus(z) + b1,
//   d sdf / d x_d = sum_h W1[h] sigmoid(100 z_h) * ( W0[h][d] + sum_j W0feat[h][j] * d f_j / d x_d ).
// The three tangent contractions [32 points x 2L] x [2L x 64] run on the tensor cores like the value one.  Tangents are
// fp32 numbers (not fp16-exact), so they are split hi + lo in registers when the A fragments are loaded and three MMAs
// (hi*hi, lo*hi, hi*lo) keep the products exact to 2^-22: a single TF32 pass would cost ~5e-4 relative on the normal.
constexpr int kTsStride = 36;   // floats per point row of a tangent tile (32 feature columns; conflict-free A-fragment loads)

__device__ __forceinline__ float warp_sdf_grad_mma(bool valid, float x, float y, float z, const __half2 *__restrict__ table,
                                                   const LevelCtx *lvl, uint32_t n_active, const float *s_net, const float *s_whi,
                                                   const float *s_wlo, float *xs, float *ts, int lane, float (&grad)[3]) {
    const int ksteps = (int)(2 * n_active + 7) >> 3;
    const int g = lane >> 2, t = lane & 3;
    float *row = xs + lane * kXsStride;
    __syncwarp();
    for (uint32_t l = 0; l < n_active; ++l) {
        float2 ff = make_float2(0.f, 0.f);
        float2 dv[3] = {ff, ff, ff};
        if (valid) {
            const LevelCtx c = lvl[l];
            Cell cell = cell_of(c, x, y, z);
            ff = __half22float2(interp_level_grad(c, cell, table, dv));
        }
        *reinterpret_cast<float2 *>(row + 2 * l) = ff;
#pragma unroll
        for (int d = 0; d < 3; ++d) *reinterpret_cast<float2 *>(ts + (d * 32 + lane) * kTsStride + 2 * l) = dv[d];
    }
    for (int cidx = 2 * (int)n_active; cidx < 8 * ksteps; ++cidx) {
        row[cidx] = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) ts[(d * 32 + lane) * kTsStride + cidx] = 0.f;
    }
    stage_point(xs, lane, valid ? x : 0.f, valid ? y : 0.f, valid ? z : 0.f);
    __syncwarp();

    float acc[2][8][4];
    warp_layer0_mma(acc, xs, s_net, s_whi, s_wlo, ksteps, lane);
    // softplus / sigmoid in fragment layout: acc <- W1[h] * sigmoid(100 z_h); sdf partial sums on the side
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const float2 w1 = *reinterpret_cast<const float2 *>(s_net + kOffW1 + 8 * nt + 2 * t);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float sp, sg;
            softplus100_both(acc[r >> 1][nt][2 * (r & 1)], sp, sg);
            part[r] = fmaf(w1.x, sp, part[r]);
            acc[r >> 1][nt][2 * (r & 1)] = w1.x * sg;
            softplus100_both(acc[r >> 1][nt][2 * (r & 1) + 1], sp, sg);
            part[r] = fmaf(w1.y, sp, part[r]);
            acc[r >> 1][nt][2 * (r & 1) + 1] = w1.y * sg;
        }
    }
    float gpart[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float tacc[2][8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 wd = *reinterpret_cast<const float2 *>(s_net + kOffW0T + d * kH + 8 * nt + 2 * t);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                tacc[mt][nt][0] = wd.x; tacc[mt][nt][1] = wd.y; tacc[mt][nt][2] = wd.x; tacc[mt][nt][3] = wd.y;
            }
        }
        const float *tile = ts + d * 32 * kTsStride;
        for (int ks = 0; ks < ksteps; ++ks) {
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float *r0 = tile + (16 * mt + g) * kTsStride + 8 * ks + t;
                const float v[4] = {r0[0], r0[8 * kTsStride], r0[4], r0[8 * kTsStride + 4]};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ahi[mt][e] = to_tf32(v[e]);
                    alo[mt][e] = to_tf32(v[e] - __uint_as_float(ahi[mt][e]));
                }
            }
            const float *wh = s_whi + (8 * ks + t) * kWStride + g, *wl = s_wlo + (8 * ks + t) * kWStride + g;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const uint32_t h0 = __float_as_uint(wh[8 * nt]), h1 = __float_as_uint(wh[4 * kWStride + 8 * nt]);
                const uint32_t l0 = __float_as_uint(wl[8 * nt]), l1 = __float_as_uint(wl[4 * kWStride + 8 * nt]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_tf32(tacc[mt][nt], ahi[mt], h0, h1);
                    mma_tf32(tacc[mt][nt], alo[mt], h0, h1);
                    mma_tf32(tacc[mt][nt], ahi[mt], l0, l1);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s = fmaf(acc[r >> 1][nt][2 * (r & 1)], tacc[r >> 1][nt][2 * (r & 1)], s);
                s = fmaf(acc[r >> 1][nt][2 * (r & 1) + 1], tacc[r >> 1][nt][2 * (r & 1) + 1], s);
            }
            gpart[d][r] = s;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        part[r] += __shfl_xor_sync(0xffffffffu, part[r], 1);
        part[r] += __shfl_xor_sync(0xffffffffu, part[r], 2);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            gpart[d][r] += __shfl_xor_sync(0xffffffffu, gpart[d][r], 1);
            gpart[d][r] += __shfl_xor_sync(0xffffffffu, gpart[d][r], 2);
        }
    }
    __syncwarp();
    if (t == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float *o = xs + (8 * r + g) * kXsStride + kXsTmp;
            o[0] = part[r]; o[1] = gpart[0][r]; o[2] = gpart[1][r]; o[3] = gpart[2][r];
        }
    }
    __syncwarp();
    const float4 o = *reinterpret_cast<const float4 *>(xs + lane * kXsStride + kXsTmp);
    grad[0] = o.y; grad[1] = o.z; grad[2] = o.w;
    return o.x + s_net[kOffB1];
}

constexpr int kMmaSmemFloats(int warps) { return kNetFloats + 2 * kWRows * kWStride + warps * 32 * kXsStride; }

}  // namespace snb
