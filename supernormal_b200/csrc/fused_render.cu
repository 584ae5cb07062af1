// fused_render.cu -- NeuS alpha, patch-based transmittance scan, directional-finite-difference normals,
// accumulation along rays and the training losses, forward and backward.
//
// Reference sequence: models/renderer.py:164-267 (≈40 ATen launches + the two nerfacc kernels
// CS/render_weight.cu:87-118 / :298-340 + three scatter_add_) and exp_runner.py:191-203.
// Here one warp owns one patch: lanes 0..26 = 3 consecutive samples x 9 in-patch rays.  The 3x3
// neighbourhood a dfd normal needs is exchanged with warp shuffles, the transmittance product and the
// backward suffix sum are carried along the samples in the reference's serial order
// (w = a*T; T = T*(1-a)), and nothing per-sample is written except what the next kernel reads.
#include "sdf_core.cuh"

namespace snb {

constexpr unsigned kFull = 0xffffffffu;

struct LaneGeom {     // per (patch, in-patch ray k) constants
    float o[3], d[3], num, den;
    int k, r, c, jj;
    bool active;
};

__device__ __forceinline__ LaneGeom lane_geom(const snb_patch_batch &b, int patch, int lane) {
    LaneGeom g;
    g.active = lane < 27;
    g.jj = g.active ? lane / 9 : 0;
    g.k = g.active ? lane % 9 : 0;
    g.r = g.k / 3;
    g.c = g.k % 3;
    const float *o = b.rays_o + 3 * (int64_t)patch;
    const float *n = b.plane_n + 3 * (int64_t)patch;
    const float *dk = b.rays_d + ((int64_t)patch * SNB_PATCH + g.k) * 3;
    const float *dc = b.rays_d + ((int64_t)patch * SNB_PATCH + SNB_PATCH / 2) * 3;
    float nx = __ldg(n), ny = __ldg(n + 1), nz = __ldg(n + 2);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        g.o[a] = __ldg(o + a);
        g.d[a] = __ldg(dk + a);
    }
    g.num = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(dc), nx), __fmul_rn(__ldg(dc + 1), ny)), __fmul_rn(__ldg(dc + 2), nz));
    g.den = __fadd_rn(__fadd_rn(__fmul_rn(g.d[0], nx), __fmul_rn(g.d[1], ny)), __fmul_rn(g.d[2], nz));
    return g;
}

struct SampleVals {  // everything the forward and the backward recompute identically for one (sample, ray)
    float s0, s1, dt, alpha, c, n, raw;
    float dl, dr, du, dd;       // distances to the left / right / upper / lower neighbour on the marching plane
    float proj[3];              // (df_dt, df_dx, df_dy)
    float g[3];                 // V_inverse @ proj
};

__device__ __forceinline__ float dist3(float ax, float ay, float az, float bx, float by, float bz) {
    float x = ax - bx, y = ay - by, z = az - bz;
    return sqrtf(x * x + y * y + z * z);
}

// Raw per-(sample, ray) inputs, software-pipelined one chunk ahead so their latency overlaps the arithmetic.
struct RawChunk {
    float t0, t1, s0, s_end;  // s_end: SDF of the interval's own end query (only meaningful when slot >= 0)
    int slot;
    bool valid;
};

__device__ __forceinline__ RawChunk load_chunk(const LaneGeom &lg, int base, int n, int j0, int S, const snb_samples &sm,
                                               const float *__restrict__ sdf) {
    RawChunk c;
    int j = j0 + lg.jj;
    c.valid = lg.active && j < n;
    c.t0 = 0.f; c.t1 = 1.f; c.s0 = 0.f; c.s_end = 0.f; c.slot = -1;
    if (c.valid) {
        int s = base + j;
        c.t0 = __ldg(sm.t0 + s);
        c.t1 = __ldg(sm.t1 + s);
        c.s0 = __ldg(sdf + (int64_t)s * SNB_PATCH + lg.k);
        c.slot = __ldg(sm.end_slot + s);
        if (c.slot >= 0) c.s_end = __ldg(sdf + ((int64_t)S + c.slot) * SNB_PATCH + lg.k);
        else if (j == n - 1) c.s_end = (s + 1 < S) ? __ldg(sdf + (int64_t)(s + 1) * SNB_PATCH + lg.k) : c.s0;  // overflow fallback only
    }
    return c;
}

// models/renderer.py:164-169: the SDF at an interval's end is the next interval's start value unless the two are not
// contiguous.  The next sample of the same ray sits 9 lanes up (same chunk) or in lane k of the next chunk.
__device__ __forceinline__ float end_sdf(const LaneGeom &lg, const RawChunk &cur, const RawChunk &nxt, int j, int n, int lane) {
    float up = __shfl_sync(kFull, cur.s0, (lane + 9) & 31);
    float nx = __shfl_sync(kFull, nxt.s0, lg.k);
    float contiguous = lg.jj < 2 ? up : nx;
    return (cur.slot >= 0 || j == n - 1) ? cur.s_end : contiguous;
}

// All 32 lanes must call this (shuffles); `valid` lanes get meaningful values.
__device__ __forceinline__ SampleVals sample_vals(const LaneGeom &lg, bool valid, float t0, float t1, float s0, float s1,
                                                  const float *vinv, float inv_s, int lane) {
    SampleVals v;
    v.s0 = s0;
    v.s1 = s1;
    float t0k = __fdiv_rn(__fmul_rn(t0, lg.num), lg.den), t1k = __fdiv_rn(__fmul_rn(t1, lg.num), lg.den);
    float px = __fadd_rn(lg.o[0], __fmul_rn(lg.d[0], t0k));
    float py = __fadd_rn(lg.o[1], __fmul_rn(lg.d[1], t0k));
    float pz = __fadd_rn(lg.o[2], __fmul_rn(lg.d[2], t0k));
    v.dt = t1k - t0k;
    // 3x3 neighbourhood of the same sample lives in lanes jj*9 + k'
    int ll = lg.c > 0 ? lane - 1 : lane, lr = lg.c < 2 ? lane + 1 : lane;
    int lu = lg.r > 0 ? lane - 3 : lane, ld = lg.r < 2 ? lane + 3 : lane;
    float sl = __shfl_sync(kFull, v.s0, ll), sr = __shfl_sync(kFull, v.s0, lr);
    float su = __shfl_sync(kFull, v.s0, lu), sd = __shfl_sync(kFull, v.s0, ld);
    v.dl = dist3(px, py, pz, __shfl_sync(kFull, px, ll), __shfl_sync(kFull, py, ll), __shfl_sync(kFull, pz, ll));
    v.dr = dist3(__shfl_sync(kFull, px, lr), __shfl_sync(kFull, py, lr), __shfl_sync(kFull, pz, lr), px, py, pz);
    v.du = dist3(px, py, pz, __shfl_sync(kFull, px, lu), __shfl_sync(kFull, py, lu), __shfl_sync(kFull, pz, lu));
    v.dd = dist3(__shfl_sync(kFull, px, ld), __shfl_sync(kFull, py, ld), __shfl_sync(kFull, pz, ld), px, py, pz);
    // models/renderer.py:194-212: forward difference along t, central / one-sided differences in the patch
    v.proj[0] = (v.s1 - v.s0) / v.dt;
    v.proj[1] = lg.c == 0 ? (sr - v.s0) / v.dr : (lg.c == 1 ? (sr - sl) / (v.dl + v.dr) : (v.s0 - sl) / v.dl);
    v.proj[2] = lg.r == 0 ? (sd - v.s0) / v.dd : (lg.r == 1 ? (sd - su) / (v.dd + v.du) : (v.s0 - su) / v.du);
#pragma unroll
    for (int a = 0; a < 3; ++a) v.g[a] = vinv[3 * a] * v.proj[0] + vinv[3 * a + 1] * v.proj[1] + vinv[3 * a + 2] * v.proj[2];
    v.c = sigmoidf_(v.s0 * inv_s);
    v.n = sigmoidf_(v.s1 * inv_s);
    v.raw = (v.c - v.n + 1e-5f) / (v.c + 1e-5f);
    v.alpha = fminf(fmaxf(v.raw, 0.f), 1.f);
    if (!valid) {
        v.alpha = 0.f;
        v.g[0] = v.g[1] = v.g[2] = 0.f;
    }
    return v;
}

// serial chain over the three sample phases: returns this lane's value-before and updates the carry
__device__ __forceinline__ float chain_mul(float &carry, float factor, const LaneGeom &lg) {
    float f0 = __shfl_sync(kFull, factor, lg.k), f1 = __shfl_sync(kFull, factor, 9 + lg.k), f2 = __shfl_sync(kFull, factor, 18 + lg.k);
    float T0 = carry, T1 = __fmul_rn(T0, f0), T2 = __fmul_rn(T1, f1);
    carry = __fmul_rn(T2, f2);
    return lg.jj == 0 ? T0 : (lg.jj == 1 ? T1 : T2);
}
__device__ __forceinline__ float chain_sub(float &carry, float term, const LaneGeom &lg) {
    float f0 = __shfl_sync(kFull, term, lg.k), f1 = __shfl_sync(kFull, term, 9 + lg.k), f2 = __shfl_sync(kFull, term, 18 + lg.k);
    float A0 = carry, A1 = A0 - f0, A2 = A1 - f1;
    carry = A2 - f2;
    return lg.jj == 0 ? A0 : (lg.jj == 1 ? A1 : A2);
}

__global__ void __launch_bounds__(128) render_fwd_kernel(snb_patch_batch b, const float *__restrict__ net, snb_samples sm,
                                                         const float *__restrict__ sdf, float *__restrict__ comp,
                                                         float *__restrict__ wsum, float *__restrict__ gradients,
                                                         float *__restrict__ weights, float *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int patch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (patch >= b.n_patches) return;
    const float inv_s = __ldg(net + kOffInvS);
    const int S = sm.totals[0];
    LaneGeom lg = lane_geom(b, patch, lane);
    float vinv[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) vinv[a] = __ldg(b.v_inv + ((int64_t)patch * SNB_PATCH + lg.k) * 9 + a);
    const int base = sm.packed_info[2 * patch], n = sm.packed_info[2 * patch + 1];
    float Tc = 1.f, cn[3] = {0.f, 0.f, 0.f}, ws = 0.f, eik = 0.f;
    RawChunk nxt = load_chunk(lg, base, n, 0, S, sm, sdf);
    for (int j0 = 0; j0 < n; j0 += 3) {
        RawChunk cur = nxt;
        nxt = load_chunk(lg, base, n, j0 + 3, S, sm, sdf);
        int j = j0 + lg.jj;
        bool valid = cur.valid;
        int s = base + j;
        float s1 = end_sdf(lg, cur, nxt, j, n, lane);
        SampleVals v = sample_vals(lg, valid, cur.t0, cur.t1, cur.s0, s1, vinv, inv_s, lane);
        float T = chain_mul(Tc, __fsub_rn(1.f, v.alpha), lg);
        float w = __fmul_rn(v.alpha, T);
        if (valid) {
            cn[0] += w * v.g[0]; cn[1] += w * v.g[1]; cn[2] += w * v.g[2];
            ws += w;
            float nrm = sqrtf(v.g[0] * v.g[0] + v.g[1] * v.g[1] + v.g[2] * v.g[2]);
            eik += (nrm - 1.f) * (nrm - 1.f);
            if (gradients) {
                float *go = gradients + ((int64_t)s * SNB_PATCH + lg.k) * 3;
                go[0] = v.g[0]; go[1] = v.g[1]; go[2] = v.g[2];
            }
            if (weights) weights[(int64_t)s * SNB_PATCH + lg.k] = w;
        }
    }
    // fold the three sample phases: lanes 0..8 end with the per-ray totals
#pragma unroll
    for (int a = 0; a < 3; ++a) cn[a] += __shfl_down_sync(kFull, cn[a], 9) + __shfl_down_sync(kFull, cn[a], 18);
    ws += __shfl_down_sync(kFull, ws, 9) + __shfl_down_sync(kFull, ws, 18);
    if (lane < 9) {
        float *co = comp + ((int64_t)patch * SNB_PATCH + lane) * 3;
        co[0] = cn[0]; co[1] = cn[1]; co[2] = cn[2];
        wsum[(int64_t)patch * SNB_PATCH + lane] = ws;
    }
    if (!lg.active) eik = 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) eik += __shfl_xor_sync(kFull, eik, o);
    if (lane == 0 && stats && eik != 0.f) atomicAdd(stats + 3, eik);
}

// exp_runner.py:169-203: masked L2 normal loss / mask_sum, BCE on the rendered opacity; emits the seeds.
__global__ void __launch_bounds__(256) patch_loss_kernel(int n_rays, const float *__restrict__ comp, const float *__restrict__ wsum,
                                                         const float *__restrict__ gt, const float *__restrict__ mask,
                                                         float normal_w, float mask_w, float *__restrict__ stats,
                                                         float *__restrict__ dcomp, float *__restrict__ dwsum) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float nsq = 0.f, bce = 0.f;
    if (i < n_rays) {
        float m = mask_w > 0.f ? (__ldg(mask + i) > 0.5f ? 1.f : 0.f) : 1.f;
        float mask_sum = stats[0];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float e = (comp[3 * i + a] - __ldg(gt + 3 * i + a)) * m;
            nsq += e * e;
            dcomp[3 * i + a] = normal_w * 2.f * e * m / mask_sum;
        }
        float w = wsum[i];
        float c = fminf(fmaxf(w, 1e-5f), 1.f - 1e-5f);
        bce = -(m * logf(c) + (1.f - m) * logf(1.f - c));
        bool pass = w >= 1e-5f && w <= 1.f - 1e-5f;
        dwsum[i] = pass ? mask_w * (c - m) / (c * (1.f - c)) / (float)n_rays : 0.f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        nsq += __shfl_xor_sync(kFull, nsq, o);
        bce += __shfl_xor_sync(kFull, bce, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 1, nsq);
        atomicAdd(stats + 2, bce);
    }
}

__global__ void __launch_bounds__(128) render_bwd_kernel(snb_patch_batch b, const float *__restrict__ net, snb_samples sm,
                                                         const float *__restrict__ sdf, const float *__restrict__ comp,
                                                         const float *__restrict__ wsum, const float *__restrict__ dcomp,
                                                         const float *__restrict__ dwsum, const float *__restrict__ dgrad,
                                                         float eik_w, float *__restrict__ d_sdf0, float *__restrict__ d_sdf1,
                                                         float *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int patch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (patch >= b.n_patches) return;
    const float inv_s = __ldg(net + kOffInvS);
    const int S = sm.totals[0];
    LaneGeom lg = lane_geom(b, patch, lane);
    float vinv[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) vinv[a] = __ldg(b.v_inv + ((int64_t)patch * SNB_PATCH + lg.k) * 9 + a);
    const int base = sm.packed_info[2 * patch], n = sm.packed_info[2 * patch + 1];
    const int64_t rk = (int64_t)patch * SNB_PATCH + lg.k;
    const float dc3[3] = {__ldg(dcomp + 3 * rk), __ldg(dcomp + 3 * rk + 1), __ldg(dcomp + 3 * rk + 2)};
    const float dws = __ldg(dwsum + rk);
    // sum_j gw_j w_j == <dcomp, comp> + dwsum * wsum   (gw_j = <dcomp, g_j> + dwsum)
    float Ac = dc3[0] * __ldg(comp + 3 * rk) + dc3[1] * __ldg(comp + 3 * rk + 1) + dc3[2] * __ldg(comp + 3 * rk + 2) + dws * __ldg(wsum + rk);
    float Tc = 1.f, dinv = 0.f;
    const float eik_scale = S > 0 ? eik_w * 2.f / ((float)S * SNB_PATCH * 1.0f) : 0.f;
    const int rowbase = lane - lg.c, colbase = lane - 3 * lg.r;
    RawChunk nxt = load_chunk(lg, base, n, 0, S, sm, sdf);
    for (int j0 = 0; j0 < n; j0 += 3) {
        RawChunk cur = nxt;
        nxt = load_chunk(lg, base, n, j0 + 3, S, sm, sdf);
        int j = j0 + lg.jj;
        bool valid = cur.valid;
        int s = base + j;
        float s1 = end_sdf(lg, cur, nxt, j, n, lane);
        SampleVals v = sample_vals(lg, valid, cur.t0, cur.t1, cur.s0, s1, vinv, inv_s, lane);
        float T = chain_mul(Tc, __fsub_rn(1.f, v.alpha), lg);
        float w = __fmul_rn(v.alpha, T);
        float gw = dc3[0] * v.g[0] + dc3[1] * v.g[1] + dc3[2] * v.g[2] + dws;
        float A = chain_sub(Ac, valid ? gw * w : 0.f, lg);
        // CS/render_weight.cu:323-338
        float dalpha = (gw * T - A) / fmaxf(1.f - v.alpha, 1e-10f);
        // d L / d g
        float dg[3];
        float nrm = sqrtf(v.g[0] * v.g[0] + v.g[1] * v.g[1] + v.g[2] * v.g[2]);
        float ek = nrm > 0.f ? eik_scale * (nrm - 1.f) / nrm : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) dg[a] = w * dc3[a] + ek * v.g[a];
        if (dgrad && valid) {
            const float *ge = dgrad + ((int64_t)s * SNB_PATCH + lg.k) * 3;
            dg[0] += __ldg(ge); dg[1] += __ldg(ge + 1); dg[2] += __ldg(ge + 2);
        }
        float q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) q[a] = vinv[a] * dg[0] + vinv[3 + a] * dg[1] + vinv[6 + a] * dg[2];  // V^T dg
        // alpha backward (clip passes the gradient on the closed interval, like torch.clamp)
        float ds0 = 0.f, ds1 = 0.f;
        if (v.raw >= 0.f && v.raw <= 1.f) {
            float ce = v.c + 1e-5f;
            float dcdf = dalpha * v.n / (ce * ce), dndf = -dalpha / ce;
            float gc = v.c * (1.f - v.c), gn = v.n * (1.f - v.n);
            ds0 = dcdf * gc * inv_s;
            ds1 = dndf * gn * inv_s;
            if (valid) dinv += dcdf * gc * v.s0 + dndf * gn * v.s1;
        }
        // dfd backward
        float ut = q[0] / v.dt;
        ds1 += ut;
        ds0 -= ut;
        float ux = q[1] / (lg.c == 0 ? v.dr : (lg.c == 1 ? v.dl + v.dr : v.dl));
        float uy = q[2] / (lg.r == 0 ? v.dd : (lg.r == 1 ? v.dd + v.du : v.du));
        if (!valid) ux = uy = 0.f;
        float ux0 = __shfl_sync(kFull, ux, rowbase), ux1 = __shfl_sync(kFull, ux, rowbase + 1), ux2 = __shfl_sync(kFull, ux, rowbase + 2);
        float uy0 = __shfl_sync(kFull, uy, colbase), uy1 = __shfl_sync(kFull, uy, colbase + 3), uy2 = __shfl_sync(kFull, uy, colbase + 6);
        ds0 += lg.c == 0 ? (-ux0 - ux1) : (lg.c == 1 ? (ux0 - ux2) : (ux1 + ux2));
        ds0 += lg.r == 0 ? (-uy0 - uy1) : (lg.r == 1 ? (uy0 - uy2) : (uy1 + uy2));
        if (valid) {
            d_sdf0[(int64_t)s * SNB_PATCH + lg.k] = ds0;
            d_sdf1[(int64_t)s * SNB_PATCH + lg.k] = ds1;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dinv += __shfl_xor_sync(kFull, dinv, o);
    if (lane == 0 && dinv != 0.f) atomicAdd(stats + 4, dinv);
}

}  // namespace snb
using namespace snb;

static int32_t check_render(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const char *who) {
    SNB_REQUIRE(b && net && sm, SNB_ERR_NULL, "%s: null struct", who);
    SNB_REQUIRE(b->n_patches >= 0, SNB_ERR_ARG, "%s: n_patches < 0", who);
    SNB_REQUIRE(b->rays_o && b->rays_d && b->plane_n && b->v_inv && net->net, SNB_ERR_NULL, "%s: null batch buffer", who);
    SNB_REQUIRE(sm->totals && sm->t0 && sm->t1 && sm->end_slot && sm->packed_info, SNB_ERR_NULL, "%s: null samples", who);
    return SNB_OK;
}

extern "C" int32_t snb_render_fwd(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const float *sdf, float *comp,
                                  float *wsum, float *gradients, float *weights, float *stats, snb_stream_t stream) {
    int32_t rc = check_render(b, net, sm, "render_fwd");
    if (rc) return rc;
    if (b->n_patches == 0) return SNB_OK;
    SNB_REQUIRE(sdf && comp && wsum, SNB_ERR_NULL, "render_fwd: null buffer");
    render_fwd_kernel<<<(unsigned)cdiv(b->n_patches, 4), 128, 0, S(stream)>>>(*b, net->net, *sm, sdf, comp, wsum, gradients, weights, stats);
    SNB_LAUNCH_CHECK("render_fwd");
    return SNB_OK;
}

extern "C" int32_t snb_patch_loss(const snb_patch_batch *b, const float *comp, const float *wsum, float normal_weight,
                                  float mask_weight, float *stats, float *dcomp, float *dwsum, snb_stream_t stream) {
    SNB_REQUIRE(b, SNB_ERR_NULL, "patch_loss: null batch");
    if (b->n_patches == 0) return SNB_OK;
    SNB_REQUIRE(comp && wsum && b->normal_gt && b->mask && stats && dcomp && dwsum, SNB_ERR_NULL, "patch_loss: null buffer");
    int n = b->n_patches * SNB_PATCH;
    patch_loss_kernel<<<(unsigned)cdiv(n, 256), 256, 0, S(stream)>>>(n, comp, wsum, b->normal_gt, b->mask, normal_weight, mask_weight, stats, dcomp, dwsum);
    SNB_LAUNCH_CHECK("patch_loss");
    return SNB_OK;
}

extern "C" int32_t snb_render_bwd(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const float *sdf,
                                  const float *comp, const float *wsum, const float *dcomp, const float *dwsum, const float *dgrad,
                                  float eikonal_weight, float *d_sdf0, float *d_sdf1, float *stats, snb_stream_t stream) {
    int32_t rc = check_render(b, net, sm, "render_bwd");
    if (rc) return rc;
    if (b->n_patches == 0) return SNB_OK;
    SNB_REQUIRE(sdf && comp && wsum && dcomp && dwsum && d_sdf0 && d_sdf1 && stats, SNB_ERR_NULL, "render_bwd: null buffer");
    render_bwd_kernel<<<(unsigned)cdiv(b->n_patches, 4), 128, 0, S(stream)>>>(*b, net->net, *sm, sdf, comp, wsum, dcomp, dwsum, dgrad,
                                                                            eikonal_weight, d_sdf0, d_sdf1, stats);
    SNB_LAUNCH_CHECK("render_bwd");
    return SNB_OK;
}

// =================================================================================================
// Single-kernel forward + loss + backward of the render stage (training path).
//
// render_fwd / patch_loss / render_bwd above walk a patch three samples at a time inside ONE warp: a serial
// chain of ~n/3 iterations of dependent loads, shuffles, divisions and square roots per patch, three launches,
// and the forward recomputed from scratch in the backward.  Here a CTA of nine warps owns a patch: warp k is
// in-patch ray k, lane j is sample j of a 32-sample chunk.  The 3x3 neighbourhood a dfd normal needs is exchanged
// through shared memory, the transmittance product and the backward suffix sum are warp scans along the lanes,
// the losses (exp_runner.py:191-203) are evaluated between the two sweeps without leaving the kernel, and patches
// that fit one chunk (the common case) keep their forward values in registers for the backward.
// The product order of a scan differs from the serial w = a*T; T *= 1-a of CS/render_weight.cu:107-116 by
// fp32 rounding only (the bit-exact serial kernels stay behind snb_weight_from_alpha_patch_*).
// =================================================================================================
namespace snb {

constexpr int kRays = SNB_PATCH;        // warps per CTA
constexpr int kChunk = 32;

struct RayConst {   // per thread, uniform within warp k
    float o[3], d[3], num, den, vinv[9];
};

struct Pt {         // forward values of one (sample, ray)
    float s0, s1, dt, alpha, c, n, raw, T, w;
    float dl, dr, du, dd;
    float g[3];
};

struct RenderSmem {
    float s0[2][kRays][kChunk];
    float px[2][kRays][kChunk], py[2][kRays][kChunk], pz[2][kRays][kChunk];
    float ux[kRays][kChunk], uy[kRays][kChunk];
    float red[kRays][4];
};

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v *= u;
    }
    return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += u;
    }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// F = true (the only instantiation launched): approximate division / reciprocal square root / exp in the render stage (2-ulp MUFU
// forms instead of the IEEE sequences: ~12 divisions, 5 square roots and 2 sigmoids per ray sample).  The stage is issue-bound (56 %
// issue-active, profiles/r01_ncu_render_fused_kernel_v1.txt) and its parity bar is a tolerance (3e-3 relative on the rendered
// normals), not bits.  Validated on a B200 in round 2 (all fused-vs-legacy / oracle tolerance tests green with it; 67 -> 48 us at
// iteration 100, 41 -> 30 us at 1000, 84 -> 62 us at 4800: profiles/r02_variants_validation.txt); the IEEE instantiation was dropped.
template <bool F> __device__ __forceinline__ float fdiv_(float a, float b) { return F ? __fdividef(a, b) : a / b; }
template <bool F> __device__ __forceinline__ float fsqrt_(float a) { return F ? (a > 0.f ? a * rsqrtf(a) : 0.f) : sqrtf(a); }
template <bool F> __device__ __forceinline__ float fsigmoid_(float x) { return F ? __fdividef(1.f, 1.f + __expf(-x)) : sigmoidf_(x); }
template <bool F> __device__ __forceinline__ float dist3_(float ax, float ay, float az, float bx, float by, float bz) {
    if (!F) return dist3(ax, ay, az, bx, by, bz);
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fsqrt_<true>(dx * dx + dy * dy + dz * dz);
}

// phase 1 of a chunk: load this (sample, ray), build its position, publish s0 / position for the neighbours
template <bool F>
__device__ __forceinline__ void pt_stage(Pt &p, const RayConst &rc, RenderSmem &sh, int buf, int k, int lane, bool valid, int s, int S,
                                         const snb_samples &sm, const float *__restrict__ sdf) {
    float t0 = 0.f, t1 = 1.f;
    p.s0 = p.s1 = 0.f;
    if (valid) {
        t0 = __ldg(sm.t0 + s);
        t1 = __ldg(sm.t1 + s);
        const int slot = __ldg(sm.end_slot + s);
        // models/renderer.py:164-169: SDF at the interval end = next start unless the intervals are not contiguous
        const float *p0 = sdf + (int64_t)s * SNB_PATCH + k;
        const float *p1 = slot >= 0 ? sdf + ((int64_t)S + slot) * SNB_PATCH + k : ((s + 1 < S) ? p0 + SNB_PATCH : p0);
        p.s0 = __ldg(p0);
        p.s1 = __ldg(p1);
    }
    float t0k = F ? __fdividef(__fmul_rn(t0, rc.num), rc.den) : __fdiv_rn(__fmul_rn(t0, rc.num), rc.den);
    float t1k = F ? __fdividef(__fmul_rn(t1, rc.num), rc.den) : __fdiv_rn(__fmul_rn(t1, rc.num), rc.den);
    p.dt = t1k - t0k;
    sh.s0[buf][k][lane] = p.s0;
    sh.px[buf][k][lane] = __fadd_rn(rc.o[0], __fmul_rn(rc.d[0], t0k));
    sh.py[buf][k][lane] = __fadd_rn(rc.o[1], __fmul_rn(rc.d[1], t0k));
    sh.pz[buf][k][lane] = __fadd_rn(rc.o[2], __fmul_rn(rc.d[2], t0k));
}

// phase 2 (after a CTA barrier): dfd normal (models/renderer.py:187-223) and NeuS alpha (:171-179)
template <bool F>
__device__ __forceinline__ void pt_finish(Pt &p, const RayConst &rc, const RenderSmem &sh, int buf, int k, int lane, bool valid, float inv_s) {
    const int r = k / 3, c = k % 3;
    const int kl = c > 0 ? k - 1 : k, kr = c < 2 ? k + 1 : k, ku = r > 0 ? k - 3 : k, kd = r < 2 ? k + 3 : k;
    const float x = sh.px[buf][k][lane], y = sh.py[buf][k][lane], z = sh.pz[buf][k][lane];
    p.dl = dist3_<F>(x, y, z, sh.px[buf][kl][lane], sh.py[buf][kl][lane], sh.pz[buf][kl][lane]);
    p.dr = dist3_<F>(sh.px[buf][kr][lane], sh.py[buf][kr][lane], sh.pz[buf][kr][lane], x, y, z);
    p.du = dist3_<F>(x, y, z, sh.px[buf][ku][lane], sh.py[buf][ku][lane], sh.pz[buf][ku][lane]);
    p.dd = dist3_<F>(sh.px[buf][kd][lane], sh.py[buf][kd][lane], sh.pz[buf][kd][lane], x, y, z);
    const float sl = sh.s0[buf][kl][lane], sr = sh.s0[buf][kr][lane], su = sh.s0[buf][ku][lane], sd = sh.s0[buf][kd][lane];
    float proj0 = fdiv_<F>(p.s1 - p.s0, p.dt);
    float proj1 = c == 0 ? fdiv_<F>(sr - p.s0, p.dr) : (c == 1 ? fdiv_<F>(sr - sl, p.dl + p.dr) : fdiv_<F>(p.s0 - sl, p.dl));
    float proj2 = r == 0 ? fdiv_<F>(sd - p.s0, p.dd) : (r == 1 ? fdiv_<F>(sd - su, p.dd + p.du) : fdiv_<F>(p.s0 - su, p.du));
#pragma unroll
    for (int a = 0; a < 3; ++a) p.g[a] = valid ? rc.vinv[3 * a] * proj0 + rc.vinv[3 * a + 1] * proj1 + rc.vinv[3 * a + 2] * proj2 : 0.f;
    p.c = fsigmoid_<F>(p.s0 * inv_s);
    p.n = fsigmoid_<F>(p.s1 * inv_s);
    p.raw = fdiv_<F>(p.c - p.n + 1e-5f, p.c + 1e-5f);
    p.alpha = valid ? fminf(fmaxf(p.raw, 0.f), 1.f) : 0.f;
}

// transmittance at the lane's sample + carry update (warp scan of 1 - alpha)
__device__ __forceinline__ void pt_weight(Pt &p, float &Tcarry, int lane) {
    float incl = warp_incl_prod(1.f - p.alpha, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    p.T = Tcarry * (lane == 0 ? 1.f : excl);
    Tcarry *= __shfl_sync(kFull, incl, 31);
    p.w = p.alpha * p.T;
}

// AD (gradient_method = 'ad', models/renderer.py:225-226): the per-sample normal is not the directional finite difference but the ANALYTIC
// gradient of the SDF at the sample start, read from grad_in[(s * 9 + k) * 3 ..] (snb_sdf_grad_patch); the backward then hands
// d loss / d gradient = w * d loss / d comp_normal + eikonal term to d_grad in the same layout instead of pushing it through V^-1 and the
// in-patch differences -- d_sdf0 / d_sdf1 carry the alpha path only.
template <bool F, bool AD>
__global__ void __launch_bounds__(32 * kRays) render_fused_kernel(snb_patch_batch b, const float *__restrict__ net, snb_samples sm,
                                                                  const float *__restrict__ sdf, float normal_w, float mask_w, float eik_w,
                                                                  float *__restrict__ comp, float *__restrict__ wsum,
                                                                  float *__restrict__ d_sdf0, float *__restrict__ d_sdf1,
                                                                  float *__restrict__ stats, const float *__restrict__ grad_in,
                                                                  float *__restrict__ d_grad) {
    __shared__ RenderSmem sh;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    // longest patches first when the compaction left a launch order (the kernel ends with its longest patch: up to 13 chunks late in the schedule)
    const int patch = sm.launch_order ? sm.launch_order[blockIdx.x] : (int)blockIdx.x;
    const float inv_s = __ldg(net + kOffInvS);
    const int S = sm.totals[0];
    const int base = sm.packed_info[2 * patch], n = sm.packed_info[2 * patch + 1];
    const int64_t rk = (int64_t)patch * SNB_PATCH + k;
    RayConst rc;
    {
        const float *o = b.rays_o + 3 * (int64_t)patch, *nn = b.plane_n + 3 * (int64_t)patch;
        const float *dk = b.rays_d + rk * 3, *dc = b.rays_d + ((int64_t)patch * SNB_PATCH + SNB_PATCH / 2) * 3;
        float nx = __ldg(nn), ny = __ldg(nn + 1), nz = __ldg(nn + 2);
#pragma unroll
        for (int a = 0; a < 3; ++a) { rc.o[a] = __ldg(o + a); rc.d[a] = __ldg(dk + a); }
        rc.num = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(dc), nx), __fmul_rn(__ldg(dc + 1), ny)), __fmul_rn(__ldg(dc + 2), nz));
        rc.den = __fadd_rn(__fadd_rn(__fmul_rn(rc.d[0], nx), __fmul_rn(rc.d[1], ny)), __fmul_rn(rc.d[2], nz));
#pragma unroll
        for (int a = 0; a < 9; ++a) rc.vinv[a] = __ldg(b.v_inv + rk * 9 + a);
    }
    const bool single = n <= kChunk;   // forward values stay in registers for the backward

    // ---------------- forward sweep ----------------
    float Tc = 1.f, cn[3] = {0.f, 0.f, 0.f}, ws = 0.f, eik = 0.f;
    Pt p;
    for (int j0 = 0, buf = 0; j0 < n; j0 += kChunk, buf ^= 1) {
        const bool valid = j0 + lane < n;
        pt_stage<F>(p, rc, sh, buf, k, lane, valid, base + j0 + lane, S, sm, sdf);
        __syncthreads();
        pt_finish<F>(p, rc, sh, buf, k, lane, valid, inv_s);
        if (AD) {
            const float *gi = grad_in + ((int64_t)(base + j0 + lane) * SNB_PATCH + k) * 3;
#pragma unroll
            for (int a = 0; a < 3; ++a) p.g[a] = valid ? __ldg(gi + a) : 0.f;
        }
        pt_weight(p, Tc, lane);
        cn[0] += p.w * p.g[0]; cn[1] += p.w * p.g[1]; cn[2] += p.w * p.g[2];
        ws += p.w;
        if (valid) {
            float nrm = fsqrt_<F>(p.g[0] * p.g[0] + p.g[1] * p.g[1] + p.g[2] * p.g[2]);
            eik += (nrm - 1.f) * (nrm - 1.f);
        }
    }
    ws = warp_sum(ws);
#pragma unroll
    for (int a = 0; a < 3; ++a) cn[a] = warp_sum(cn[a]);
    eik = warp_sum(eik);

    // ---------------- losses and their seeds for ray k (exp_runner.py:169-203) ----------------
    const int n_rays = b.n_patches * SNB_PATCH;
    const float mask_sum = stats[0];
    const float m = mask_w > 0.f ? (__ldg(b.mask + rk) > 0.5f ? 1.f : 0.f) : 1.f;
    float dc3[3], nsq = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float e = (cn[a] - __ldg(b.normal_gt + 3 * rk + a)) * m;
        nsq += e * e;
        dc3[a] = normal_w * 2.f * e * m / mask_sum;
    }
    const float cw = fminf(fmaxf(ws, 1e-5f), 1.f - 1e-5f);
    const float bce = -(m * logf(cw) + (1.f - m) * logf(1.f - cw));
    const float dws = (ws >= 1e-5f && ws <= 1.f - 1e-5f) ? mask_w * (cw - m) / (cw * (1.f - cw)) / (float)n_rays : 0.f;
    if (lane == 0) {
        float *co = comp + rk * 3;
        co[0] = cn[0]; co[1] = cn[1]; co[2] = cn[2];
        wsum[rk] = ws;
        sh.red[k][0] = nsq; sh.red[k][1] = bce; sh.red[k][2] = eik;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < kRays; ++q) t += sh.red[q][threadIdx.x];
        if (t != 0.f) atomicAdd(stats + 1 + threadIdx.x, t);
    }
    if (n == 0 || d_sdf0 == nullptr) return;

    // ---------------- backward sweep ----------------
    const int r = k / 3, c = k % 3;
    const float eik_scale = S > 0 ? eik_w * 2.f / ((float)S * SNB_PATCH) : 0.f;
    float Ac = dc3[0] * cn[0] + dc3[1] * cn[1] + dc3[2] * cn[2] + dws * ws;   // sum_j gw_j w_j
    float dinv = 0.f;
    Tc = 1.f;
    for (int j0 = 0, buf = 0; j0 < n; j0 += kChunk, buf ^= 1) {
        const bool valid = j0 + lane < n;
        const int s = base + j0 + lane;
        if (!single) {
            pt_stage<F>(p, rc, sh, buf, k, lane, valid, s, S, sm, sdf);
            __syncthreads();
            pt_finish<F>(p, rc, sh, buf, k, lane, valid, inv_s);
            if (AD) {
                const float *gi = grad_in + ((int64_t)s * SNB_PATCH + k) * 3;
#pragma unroll
                for (int a = 0; a < 3; ++a) p.g[a] = valid ? __ldg(gi + a) : 0.f;
            }
            pt_weight(p, Tc, lane);
        }
        float gw = dc3[0] * p.g[0] + dc3[1] * p.g[1] + dc3[2] * p.g[2] + dws;
        float term = valid ? gw * p.w : 0.f;
        float pin = warp_incl_sum(term, lane);
        float A = Ac - (pin - term);                         // sum over samples >= j of gw*w  (CS/render_weight.cu:323-338)
        Ac -= __shfl_sync(kFull, pin, 31);
        float dalpha = fdiv_<F>(gw * p.T - A, fmaxf(1.f - p.alpha, 1e-10f));
        float nrm = fsqrt_<F>(p.g[0] * p.g[0] + p.g[1] * p.g[1] + p.g[2] * p.g[2]);
        float ek = nrm > 0.f ? (F ? __fdividef(eik_scale * (nrm - 1.f), nrm) : eik_scale * (nrm - 1.f) / nrm) : 0.f;
        float dg[3], q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) dg[a] = p.w * dc3[a] + ek * p.g[a];
#pragma unroll
        for (int a = 0; a < 3; ++a) q[a] = AD ? 0.f : rc.vinv[a] * dg[0] + rc.vinv[3 + a] * dg[1] + rc.vinv[6 + a] * dg[2];  // V^T dg
        if (AD && valid) {
            float *go = d_grad + ((int64_t)s * SNB_PATCH + k) * 3;
            go[0] = dg[0]; go[1] = dg[1]; go[2] = dg[2];
        }
        float ds0 = 0.f, ds1 = 0.f;
        if (p.raw >= 0.f && p.raw <= 1.f) {   // clip passes the gradient on the closed interval, like torch.clamp
            float ce = p.c + 1e-5f;
            float dcdf = F ? __fdividef(dalpha * p.n, ce * ce) : dalpha * p.n / (ce * ce), dndf = F ? __fdividef(-dalpha, ce) : -dalpha / ce;
            float gc = p.c * (1.f - p.c), gn = p.n * (1.f - p.n);
            ds0 = dcdf * gc * inv_s;
            ds1 = dndf * gn * inv_s;
            if (valid) dinv += dcdf * gc * p.s0 + dndf * gn * p.s1;
        }
        float ut = fdiv_<F>(q[0], p.dt);
        ds1 += ut;
        ds0 -= ut;
        float ux = fdiv_<F>(q[1], c == 0 ? p.dr : (c == 1 ? p.dl + p.dr : p.dl));
        float uy = fdiv_<F>(q[2], r == 0 ? p.dd : (r == 1 ? p.dd + p.du : p.du));
        sh.ux[k][lane] = valid ? ux : 0.f;
        sh.uy[k][lane] = valid ? uy : 0.f;
        __syncthreads();
        const int rb = 3 * r;
        float ux0 = sh.ux[rb][lane], ux1 = sh.ux[rb + 1][lane], ux2 = sh.ux[rb + 2][lane];
        float uy0 = sh.uy[c][lane], uy1 = sh.uy[c + 3][lane], uy2 = sh.uy[c + 6][lane];
        ds0 += c == 0 ? (-ux0 - ux1) : (c == 1 ? (ux0 - ux2) : (ux1 + ux2));
        ds0 += r == 0 ? (-uy0 - uy1) : (r == 1 ? (uy0 - uy2) : (uy1 + uy2));
        if (valid) {
            d_sdf0[(int64_t)s * SNB_PATCH + k] = ds0;
            d_sdf1[(int64_t)s * SNB_PATCH + k] = ds1;
        }
        if (!single) __syncthreads();   // ux/uy are rewritten by the next chunk
    }
    dinv = warp_sum(dinv);
    if (lane == 0) sh.red[k][3] = dinv;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < kRays; ++q) t += sh.red[q][3];
        if (t != 0.f) atomicAdd(stats + 4, t);
    }
}

}  // namespace snb

/* forward + losses + backward of the render stage in one launch (same results as snb_render_fwd -> snb_patch_loss ->
 * snb_render_bwd up to fp32 summation order).  d_sdf0/d_sdf1 null: forward + losses only. */
extern "C" int32_t snb_render_fused(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const float *sdf,
                                    float normal_weight, float mask_weight, float eikonal_weight, float *comp, float *wsum,
                                    float *d_sdf0, float *d_sdf1, float *stats, snb_stream_t stream) {
    int32_t rc = check_render(b, net, sm, "render_fused");
    if (rc) return rc;
    if (b->n_patches == 0) return SNB_OK;
    SNB_REQUIRE(sdf && comp && wsum && stats && b->normal_gt && b->mask, SNB_ERR_NULL, "render_fused: null buffer");
    SNB_REQUIRE((d_sdf0 == nullptr) == (d_sdf1 == nullptr), SNB_ERR_NULL, "render_fused: d_sdf0/d_sdf1 must both be given or both be null");
    render_fused_kernel<true, false><<<(unsigned)b->n_patches, 32 * kRays, 0, S(stream)>>>(*b, net->net, *sm, sdf, normal_weight, mask_weight,
                                                                                         eikonal_weight, comp, wsum, d_sdf0, d_sdf1, stats, nullptr, nullptr);
    SNB_LAUNCH_CHECK("render_fused");
    return SNB_OK;
}

/* the same stage for gradient_method = 'ad': per-sample normals = grad_in (analytic SDF gradients, f32[9 * capacity, 3] at (s * 9 + k)), seeds
 * d_grad = d loss / d grad_in in the same layout; d_sdf0 / d_sdf1 carry the alpha path only.  d_sdf0 / d_sdf1 / d_grad all null: forward only. */
extern "C" int32_t snb_render_fused_ad(const snb_patch_batch *b, const snb_net *net, const snb_samples *sm, const float *sdf, const float *grad_in,
                                       float normal_weight, float mask_weight, float eikonal_weight, float *comp, float *wsum,
                                       float *d_sdf0, float *d_sdf1, float *d_grad, float *stats, snb_stream_t stream) {
    int32_t rc = check_render(b, net, sm, "render_fused_ad");
    if (rc) return rc;
    if (b->n_patches == 0) return SNB_OK;
    SNB_REQUIRE(sdf && grad_in && comp && wsum && stats && b->normal_gt && b->mask, SNB_ERR_NULL, "render_fused_ad: null buffer");
    SNB_REQUIRE((d_sdf0 == nullptr) == (d_sdf1 == nullptr) && (d_sdf0 == nullptr) == (d_grad == nullptr), SNB_ERR_NULL,
                "render_fused_ad: d_sdf0 / d_sdf1 / d_grad must all be given or all be null");
    render_fused_kernel<true, true><<<(unsigned)b->n_patches, 32 * kRays, 0, S(stream)>>>(*b, net->net, *sm, sdf, normal_weight, mask_weight,
                                                                                        eikonal_weight, comp, wsum, d_sdf0, d_sdf1, stats, grad_in, d_grad);
    SNB_LAUNCH_CHECK("render_fused_ad");
    return SNB_OK;
}
