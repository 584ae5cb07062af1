// render_ops.cu -- packing, transmittance, patch-based rendering weights (fwd/bwd) and
// accumulation: the nerfacc operators SuperNormal's renderer calls.
//
// Rounding contract (bit-exact with the reference kernels as compiled for sm_100a, SURVEY.md
// Appendix C): forward  w = a*T; T = T*(1-a)  with separately rounded ops; backward uses
// fma(g,w,accum), fma(g,T,-accum), IEEE division, fma(-g,w,accum).
//
// Thread mapping: the reference uses a 16x16 block with x = patch, y = in-patch ray (7 of 16 y-lanes
// idle, stride-P addresses across x).  Here thread = patch*P + k, so the P rays of a patch sit in
// adjacent lanes and every iteration j reads one contiguous P*4-byte segment per patch.
#include "common.cuh"

namespace snb {

__global__ void weight_patch_fwd_kernel(int32_t n_patches, int32_t P, const int32_t *__restrict__ packed,
                                        const float *__restrict__ alphas, float *__restrict__ weights) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)n_patches * P) return;
    int i = (int)(tid / P), k = (int)(tid % P);
    int base = packed[2 * i], steps = packed[2 * i + 1];
    const float *a = alphas + (int64_t)base * P + k;
    float *w = weights + (int64_t)base * P + k;
    float T = 1.f;
    for (int j = 0; j < steps; ++j) {
        float al = a[(int64_t)j * P];
        w[(int64_t)j * P] = __fmul_rn(al, T);
        T = __fmul_rn(T, __fsub_rn(1.f, al));
    }
}

__global__ void weight_patch_bwd_kernel(int32_t n_patches, int32_t P, const int32_t *__restrict__ packed,
                                        const float *__restrict__ alphas, const float *__restrict__ weights,
                                        const float *__restrict__ gw, float *__restrict__ ga) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)n_patches * P) return;
    int i = (int)(tid / P), k = (int)(tid % P);
    int base = packed[2 * i], steps = packed[2 * i + 1];
    int64_t o = (int64_t)base * P + k;
    float accum = 0.f;
    for (int j = 0; j < steps; ++j) accum = __fmaf_rn(gw[o + (int64_t)j * P], weights[o + (int64_t)j * P], accum);
    float T = 1.f;
    for (int j = 0; j < steps; ++j) {
        int64_t id = o + (int64_t)j * P;
        float al = alphas[id], g = gw[id], w = weights[id];
        float num = __fmaf_rn(g, T, -accum);
        ga[id] = __fdiv_rn(num, fmaxf(__fsub_rn(1.f, al), 1e-10f));
        accum = __fmaf_rn(-g, w, accum);
        T = __fmul_rn(T, __fsub_rn(1.f, al));
    }
}

__global__ void transmittance_kernel(int32_t n_rays, const int32_t *__restrict__ packed,
                                     const float *__restrict__ alphas, float *__restrict__ trans) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    int base = packed[2 * i], steps = packed[2 * i + 1];
    float T = 1.f;
    for (int j = 0; j < steps; ++j) {
        trans[base + j] = T;
        T = __fmul_rn(T, __fsub_rn(1.f, alphas[base + j]));
    }
}

__global__ void count_by_ray_kernel(int64_t n, const int64_t *__restrict__ idx, int32_t n_rays, int32_t *__restrict__ counts) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int64_t r = idx[s];
    if (r >= 0 && r < n_rays) atomicAdd(counts + r, 1);
}

__global__ void max_i64_kernel(int64_t n, const int64_t *__restrict__ v, long long *out) {
    long long m = -1;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
        m = max(m, (long long)v[s]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
__global__ void set_i64_kernel(long long *p, long long v) { *p = v; }

__global__ void accumulate_fwd_kernel(int64_t n_elems, int32_t P, int32_t D, const float *__restrict__ w,
                                      const float *__restrict__ vals, const int64_t *__restrict__ idx, int32_t n_out,
                                      float *__restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over S*P*D
    if (e >= n_elems) return;
    int d = (int)(e % D);
    int64_t sk = e / D;
    int64_t s = sk / P;
    int k = (int)(sk % P);
    int64_t r = idx[s];
    if (r < 0 || r >= n_out) return;
    float v = w[sk];
    if (vals) v = __fmul_rn(v, vals[e]);
    atomicAdd(out + ((int64_t)r * P + k) * D + d, v);
}

__global__ void accumulate_bwd_kernel(int64_t n_sk, int32_t P, int32_t D, const float *__restrict__ w,
                                      const float *__restrict__ vals, const int64_t *__restrict__ idx,
                                      const float *__restrict__ go, float *__restrict__ gw, float *__restrict__ gv) {
    int64_t sk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over S*P
    if (sk >= n_sk) return;
    int64_t s = sk / P;
    int k = (int)(sk % P);
    const float *g = go + ((int64_t)idx[s] * P + k) * D;
    float acc = 0.f, ww = w[sk];
    for (int d = 0; d < D; ++d) {
        float gd = g[d];
        acc += vals ? gd * vals[sk * D + d] : gd;
        if (gv) gv[sk * D + d] = gd * ww;
    }
    if (gw) gw[sk] = acc;
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_count_by_ray(int64_t n, const int64_t *ray_indices, int32_t n_rays, int32_t *num_steps, snb_stream_t stream) {
    SNB_REQUIRE(n >= 0 && n_rays >= 0, SNB_ERR_ARG, "count_by_ray: negative size");
    if (n_rays == 0) return SNB_OK;
    SNB_REQUIRE(num_steps && (n == 0 || ray_indices), SNB_ERR_NULL, "count_by_ray: null buffer");
    cudaMemsetAsync(num_steps, 0, sizeof(int32_t) * n_rays, S(stream));
    if (n) count_by_ray_kernel<<<(unsigned)cdiv(n, 256), 256, 0, S(stream)>>>(n, ray_indices, n_rays, num_steps);
    SNB_LAUNCH_CHECK("count_by_ray");
    return SNB_OK;
}

extern "C" int32_t snb_max_i64(int64_t n, const int64_t *v, int64_t *out, snb_stream_t stream) {
    SNB_REQUIRE(out && (n == 0 || v), SNB_ERR_NULL, "max_i64: null buffer");
    set_i64_kernel<<<1, 1, 0, S(stream)>>>((long long *)out, -1);
    if (n) max_i64_kernel<<<(unsigned)(cdiv(n, 256) < 1184 ? cdiv(n, 256) : 1184), 256, 0, S(stream)>>>(n, v, (long long *)out);
    SNB_LAUNCH_CHECK("max_i64");
    return SNB_OK;
}

extern "C" int32_t snb_transmittance_from_alpha(int32_t n_rays, const int32_t *packed_info, const float *alphas,
                                                float *transmittance, snb_stream_t stream) {
    SNB_REQUIRE(n_rays >= 0, SNB_ERR_ARG, "transmittance: n_rays < 0");
    if (n_rays == 0) return SNB_OK;
    SNB_REQUIRE(packed_info, SNB_ERR_NULL, "transmittance: null packed_info");
    transmittance_kernel<<<(unsigned)cdiv(n_rays, 128), 128, 0, S(stream)>>>(n_rays, packed_info, alphas, transmittance);
    SNB_LAUNCH_CHECK("transmittance");
    return SNB_OK;
}

extern "C" int32_t snb_weight_from_alpha_patch_fwd(int32_t n_patches, int32_t P, const int32_t *packed_info,
                                                   const float *alphas, float *weights, snb_stream_t stream) {
    SNB_REQUIRE(n_patches >= 0 && P >= 1, SNB_ERR_ARG, "weight_fwd: bad sizes n_patches=%d P=%d", n_patches, P);
    if (n_patches == 0) return SNB_OK;
    SNB_REQUIRE(packed_info, SNB_ERR_NULL, "weight_fwd: null packed_info");
    weight_patch_fwd_kernel<<<(unsigned)cdiv((int64_t)n_patches * P, 128), 128, 0, S(stream)>>>(n_patches, P, packed_info, alphas, weights);
    SNB_LAUNCH_CHECK("weight_fwd");
    return SNB_OK;
}

extern "C" int32_t snb_weight_from_alpha_patch_bwd(int32_t n_patches, int32_t P, const int32_t *packed_info,
                                                   const float *alphas, const float *weights, const float *grad_weights,
                                                   float *grad_alphas, snb_stream_t stream) {
    SNB_REQUIRE(n_patches >= 0 && P >= 1, SNB_ERR_ARG, "weight_bwd: bad sizes n_patches=%d P=%d", n_patches, P);
    if (n_patches == 0) return SNB_OK;
    SNB_REQUIRE(packed_info, SNB_ERR_NULL, "weight_bwd: null packed_info");
    weight_patch_bwd_kernel<<<(unsigned)cdiv((int64_t)n_patches * P, 128), 128, 0, S(stream)>>>(n_patches, P, packed_info, alphas,
                                                                                               weights, grad_weights, grad_alphas);
    SNB_LAUNCH_CHECK("weight_bwd");
    return SNB_OK;
}

extern "C" int32_t snb_accumulate_fwd(int64_t n_samples, int32_t P, int32_t D, const float *weights, const float *values,
                                      const int64_t *ray_indices, int32_t n_out, float *out, snb_stream_t stream) {
    SNB_REQUIRE(n_samples >= 0 && P >= 1 && D >= 1 && n_out >= 0, SNB_ERR_ARG, "accumulate_fwd: bad sizes");
    if (n_out == 0) return SNB_OK;
    SNB_REQUIRE(out, SNB_ERR_NULL, "accumulate_fwd: null out");
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_out * P * D, S(stream));
    if (n_samples == 0) return SNB_OK;
    SNB_REQUIRE(weights && ray_indices, SNB_ERR_NULL, "accumulate_fwd: null input");
    int64_t n = n_samples * P * D;
    accumulate_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, S(stream)>>>(n, P, D, weights, values, ray_indices, n_out, out);
    SNB_LAUNCH_CHECK("accumulate_fwd");
    return SNB_OK;
}

extern "C" int32_t snb_accumulate_bwd(int64_t n_samples, int32_t P, int32_t D, const float *weights, const float *values,
                                      const int64_t *ray_indices, const float *grad_out, float *grad_weights,
                                      float *grad_values, snb_stream_t stream) {
    SNB_REQUIRE(n_samples >= 0 && P >= 1 && D >= 1, SNB_ERR_ARG, "accumulate_bwd: bad sizes");
    if (n_samples == 0) return SNB_OK;
    SNB_REQUIRE(weights && ray_indices && grad_out, SNB_ERR_NULL, "accumulate_bwd: null input");
    int64_t n = n_samples * P;
    accumulate_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, S(stream)>>>(n, P, D, weights, values, ray_indices, grad_out,
                                                                        grad_weights, grad_values);
    SNB_LAUNCH_CHECK("accumulate_bwd");
    return SNB_OK;
}
