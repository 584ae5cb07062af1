// optim.cu -- parameter plumbing of the fused step: weight-norm folding (models/fields.py:66-67),
// variance network (models/fields.py:133-139), their backward, and Adam (exp_runner.py:97,205-207).
#include "sdf_core.cuh"

namespace snb {

constexpr int kPrepThreads = 1024;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the whole CTA (kPrepThreads threads); every thread gets the result
__device__ __forceinline__ float block_sum(float v, float *s_red) {
    v = warp_sum_f(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x & 31) < (kPrepThreads / 32) ? s_red[threadIdx.x & 31] : 0.f;
    return warp_sum_f(t);
}

// one CTA, 1024 threads: warp w owns rows 2w, 2w+1 of lin0 (lanes over the input features)
__global__ void __launch_bounds__(kPrepThreads) prep_net_kernel(int n_levels, const float *__restrict__ small, float *__restrict__ net,
                                                                int n_mask, const float *__restrict__ mask, float *__restrict__ stats) {
    __shared__ float s_norm[kH];
    __shared__ float s_red[32];
    const int d_in = 3 + 2 * n_levels;
    const float *v0 = small, *g0 = v0 + kH * d_in, *b0 = g0 + kH, *v1 = b0 + kH, *g1 = v1 + kH, *b1 = g1 + 1, *var = b1 + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int row = 2 * warp + q;
        float ss = 0.f;
        for (int i = lane; i < d_in; i += 32) ss += v0[row * d_in + i] * v0[row * d_in + i];
        ss = warp_sum_f(ss);
        if (lane == 0) s_norm[row] = sqrtf(ss);
    }
    float part = tid < kH ? v1[tid] * v1[tid] : 0.f;
    float n1 = sqrtf(block_sum(part, s_red));   // also orders s_norm
    for (int e = tid; e < kDinMax * kH; e += kPrepThreads) {
        int i = e / kH, h = e % kH;
        net[kOffW0T + e] = i < d_in ? g0[h] * v0[h * d_in + i] / s_norm[h] : 0.f;
    }
    if (tid < kH) {
        net[kOffB0 + tid] = b0[tid];
        net[kOffW1 + tid] = g1[0] * v1[tid] / n1;
    }
    if (tid == 0) {
        net[kOffB1] = b1[0];
        net[kOffInvS] = fminf(fmaxf(expf(var[0] * 10.f), 1e-6f), 1e6f);
    }
    for (int e = kOffInvS + 1 + tid; e < kNetFloats; e += kPrepThreads) net[e] = 0.f;
    if (stats) {
        float m = 0.f;
        for (int e = tid; e < n_mask; e += kPrepThreads) m += mask[e] > 0.5f ? 1.f : 0.f;
        m = block_sum(m, s_red);
        if (tid == 0) stats[0] = m + 1e-5f;
        if (tid >= 1 && tid < 8) stats[tid] = 0.f;
    }
}

// one CTA, 1024 threads: gradients w.r.t. the folded weights -> gradients w.r.t. (v, g, b, variance)
__global__ void __launch_bounds__(kPrepThreads) unfold_grads_kernel(int n_levels, const float *__restrict__ small,
                                                                    const float *__restrict__ net_grad, const float *__restrict__ stats,
                                                                    float *__restrict__ small_grad) {
    __shared__ float s_red[32];
    const int d_in = 3 + 2 * n_levels;
    const int o_g0 = kH * d_in, o_b0 = o_g0 + kH, o_v1 = o_b0 + kH, o_g1 = o_v1 + kH, o_b1 = o_g1 + 1, o_var = o_b1 + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // lin0 row:  W = g v/|v|  ->  dg = <dW,v>/|v|,  dv = g/|v| (dW - <dW,v> v/|v|^2)
        const int row = 2 * warp + q;
        const float *v = small + row * d_in;
        float ss = 0.f, dot = 0.f;
        for (int i = lane; i < d_in; i += 32) {
            float vi = v[i];
            ss += vi * vi;
            dot += net_grad[kOffW0T + i * kH + row] * vi;
        }
        ss = warp_sum_f(ss);
        dot = warp_sum_f(dot);
        float nrm = sqrtf(ss), g = small[o_g0 + row];
        for (int i = lane; i < d_in; i += 32) small_grad[row * d_in + i] = g / nrm * (net_grad[kOffW0T + i * kH + row] - dot * v[i] / ss);
        if (lane == 0) {
            small_grad[o_g0 + row] = dot / nrm;
            small_grad[o_b0 + row] = net_grad[kOffB0 + row];
        }
    }
    float v1 = tid < kH ? small[o_v1 + tid] : 0.f, dw1 = tid < kH ? net_grad[kOffW1 + tid] : 0.f;
    float ss1 = block_sum(v1 * v1, s_red);
    float dot1 = block_sum(dw1 * v1, s_red);
    float n1 = sqrtf(ss1), g1 = small[o_g1];
    if (tid < kH) small_grad[o_v1 + tid] = g1 / n1 * (dw1 - dot1 * v1 / ss1);
    if (tid == 0) {
        small_grad[o_g1] = dot1 / n1;
        small_grad[o_b1] = net_grad[kOffB1];
        float e = expf(small[o_var] * 10.f);
        small_grad[o_var] = (e >= 1e-6f && e <= 1e6f) ? stats[4] * 10.f * e : 0.f;
    }
}

// torch.optim.Adam (no amsgrad / weight decay): p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_kernel(int64_t n, float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, __half *__restrict__ p16, int64_t f16_start, float lr,
                                                   float b1, float b2, float eps, float bc1, float rsqrt_bc2, float gscale) {
    const int64_t n4 = n >> 2;
    const float step = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4 *>(p)[i], G = reinterpret_cast<float4 *>(g)[i];
        float4 M = reinterpret_cast<float4 *>(m)[i], V = reinterpret_cast<float4 *>(v)[i];
        float *pp = &P.x, *gg = &G.x, *mm = &M.x, *vv = &V.x;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float gr = gg[a] * gscale;
            mm[a] = b1 * mm[a] + (1.f - b1) * gr;
            vv[a] = b2 * vv[a] + (1.f - b2) * gr * gr;
            pp[a] -= step * mm[a] / (sqrtf(vv[a]) * rsqrt_bc2 + eps);
        }
        reinterpret_cast<float4 *>(p)[i] = P;
        reinterpret_cast<float4 *>(m)[i] = M;
        reinterpret_cast<float4 *>(v)[i] = V;
        reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p16 && 4 * i >= f16_start) {   // fp16 copy of the parameters from f16_start on (the hash table behind the MLP block)
            __half2 *q = reinterpret_cast<__half2 *>(p16 + (4 * i - f16_start));
            q[0] = __floats2half2_rn(P.x, P.y);
            q[1] = __floats2half2_rn(P.z, P.w);
        }
    }
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gr = g[i] * gscale;
        float mi = b1 * m[i] + (1.f - b1) * gr, vi = b2 * v[i] + (1.f - b2) * gr * gr;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step * mi / (sqrtf(vi) * rsqrt_bc2 + eps);
        g[i] = 0.f;
        if (p16 && i >= f16_start) p16[i - f16_start] = __float2half_rn(p[i]);
    }
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_prep_net(int32_t n_levels, const float *small, float *net, int32_t n_mask, const float *mask, float *stats,
                                snb_stream_t stream) {
    SNB_REQUIRE(n_levels >= 1 && n_levels <= SNB_MAX_LEVELS && n_mask >= 0, SNB_ERR_ARG, "prep_net: bad n_levels/n_mask");
    SNB_REQUIRE(small && net, SNB_ERR_NULL, "prep_net: null buffer");
    SNB_REQUIRE(!stats || n_mask == 0 || mask, SNB_ERR_NULL, "prep_net: null mask");
    prep_net_kernel<<<1, kPrepThreads, 0, S(stream)>>>(n_levels, small, net, n_mask, mask, stats);
    SNB_LAUNCH_CHECK("prep_net");
    return SNB_OK;
}

extern "C" int32_t snb_unfold_grads(int32_t n_levels, const float *small, const float *net_grad, const float *stats, float *small_grad,
                                    snb_stream_t stream) {
    SNB_REQUIRE(n_levels >= 1 && n_levels <= SNB_MAX_LEVELS, SNB_ERR_ARG, "unfold_grads: bad n_levels");
    SNB_REQUIRE(small && net_grad && stats && small_grad, SNB_ERR_NULL, "unfold_grads: null buffer");
    unfold_grads_kernel<<<1, kPrepThreads, 0, S(stream)>>>(n_levels, small, net_grad, stats, small_grad);
    SNB_LAUNCH_CHECK("unfold_grads");
    return SNB_OK;
}

namespace snb {
int32_t adam_launch(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *param_f16, int64_t f16_start, float lr,
                    float beta1, float beta2, float eps, int32_t step_count, float grad_scale, snb_stream_t stream) {
    SNB_REQUIRE(n >= 0 && step_count >= 1, SNB_ERR_ARG, "adam_step: bad n/step_count");
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(param && grad && exp_avg && exp_avg_sq, SNB_ERR_NULL, "adam_step: null buffer");
    SNB_REQUIRE(aligned(param, 16) && aligned(grad, 16) && aligned(exp_avg, 16) && aligned(exp_avg_sq, 16) && aligned(param_f16, 8) && f16_start % 4 == 0,
                SNB_ERR_ALIGN, "adam_step: buffers must be 16-byte aligned");
    double bc1 = 1.0 - pow((double)beta1, step_count), bc2 = 1.0 - pow((double)beta2, step_count);
    int64_t blocks = cdiv(n / 4 + 1, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    adam_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(n, param, grad, exp_avg, exp_avg_sq, (__half *)param_f16, f16_start, lr, beta1, beta2,
                                                        eps, (float)bc1, (float)(1.0 / sqrt(bc2)), grad_scale);
    SNB_LAUNCH_CHECK("adam_step");
    return SNB_OK;
}
}  // namespace snb

extern "C" int32_t snb_adam_step(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *param_f16, float lr,
                                 float beta1, float beta2, float eps, int32_t step_count, float grad_scale, snb_stream_t stream) {
    return snb::adam_launch(n, param, grad, exp_avg, exp_avg_sq, param_f16, 0, lr, beta1, beta2, eps, step_count, grad_scale, stream);
}
