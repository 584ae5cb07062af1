// optim.cu -- parameter plumbing of the fused step: weight-norm folding (models/fields.py:66-67),
// variance network (models/fields.py:133-139), their backward, and Adam (exp_runner.py:97,205-207).
#include "sampler.cuh"

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

// ---- building blocks shared by the stand-alone kernels and the step-tail kernel (generic pointers: global or shared) ----

struct SmallLayout {   // offsets inside `small` = v0[64, d_in] | g0[64] | b0[64] | v1[64] | g1 | b1 | variance
    int d_in, g0, b0, v1, g1, b1, var, n;
    __device__ __forceinline__ explicit SmallLayout(int n_levels) {
        d_in = 3 + 2 * n_levels;
        g0 = kH * d_in; b0 = g0 + kH; v1 = b0 + kH; g1 = v1 + kH; b1 = g1 + 1; var = b1 + 1; n = var + 1;
    }
};

// |v0[row]| of one lin0 row, lanes over the input features; every lane gets the result
__device__ __forceinline__ float row_sumsq(const float *v, int d_in, int lane) {
    float ss = 0.f;
    for (int i = lane; i < d_in; i += 32) ss += v[i] * v[i];
    return warp_sum_f(ss);
}

// folded lin0 / lin1 / variance -> net (models/fields.py:66-67,133-139); s_norm[64] = |v0[row]|, n1 = |v1|
__device__ __forceinline__ void fold_store(const SmallLayout &L, const float *small, const float *s_norm, float n1, float *__restrict__ net,
                                           int tid, int nthreads) {
    const float *v0 = small, *g0 = small + L.g0, *b0 = small + L.b0, *v1 = small + L.v1;
    for (int e = tid; e < kDinMax * kH; e += nthreads) {
        int i = e / kH, h = e % kH;
        net[kOffW0T + e] = i < L.d_in ? g0[h] * v0[h * L.d_in + i] / s_norm[h] : 0.f;
    }
    if (tid < kH) {
        net[kOffB0 + tid] = b0[tid];
        net[kOffW1 + tid] = small[L.g1] * v1[tid] / n1;
    }
    if (tid == 0) {
        net[kOffB1] = small[L.b1];
        net[kOffInvS] = fminf(fmaxf(expf(small[L.var] * 10.f), 1e-6f), 1e6f);
    }
    for (int e = kOffInvS + 1 + tid; e < kNetFloats; e += nthreads) net[e] = 0.f;
}

// one lin0 row of the weight-norm backward:  W = g v/|v|  ->  dg = <dW,v>/|v|,  dv = g/|v| (dW - <dW,v> v/|v|^2)
__device__ __forceinline__ void unfold_row(const SmallLayout &L, int row, int lane, const float *small, const float *net_grad, float *small_grad) {
    const float *v = small + row * L.d_in;
    float ss = 0.f, dot = 0.f;
    for (int i = lane; i < L.d_in; i += 32) {
        float vi = v[i];
        ss += vi * vi;
        dot += net_grad[kOffW0T + i * kH + row] * vi;
    }
    ss = warp_sum_f(ss);
    dot = warp_sum_f(dot);
    float nrm = sqrtf(ss), g = small[L.g0 + row];
    for (int i = lane; i < L.d_in; i += 32) small_grad[row * L.d_in + i] = g / nrm * (net_grad[kOffW0T + i * kH + row] - dot * v[i] / ss);
    if (lane == 0) {
        small_grad[L.g0 + row] = dot / nrm;
        small_grad[L.b0 + row] = net_grad[kOffB0 + row];
    }
}

// lin1 (one row of 64), biases, variance: ss1 = |v1|^2, dot1 = <dW1, v1> (block-wide sums, identical in every thread)
__device__ __forceinline__ void unfold_tail(const SmallLayout &L, int tid, float ss1, float dot1, const float *small, const float *net_grad,
                                            float d_inv_s, float *small_grad) {
    float n1 = sqrtf(ss1), g1 = small[L.g1];
    if (tid < kH) small_grad[L.v1 + tid] = g1 / n1 * (net_grad[kOffW1 + tid] - dot1 * small[L.v1 + tid] / ss1);
    if (tid == 0) {
        small_grad[L.g1] = dot1 / n1;
        small_grad[L.b1] = net_grad[kOffB1];
        float e = expf(small[L.var] * 10.f);
        small_grad[L.var] = (e >= 1e-6f && e <= 1e6f) ? d_inv_s * 10.f * e : 0.f;
    }
}

// torch.optim.Adam (no amsgrad / weight decay): p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
struct AdamCoef { float b1, b2, eps, step, rsqrt_bc2, gscale; };
__device__ __forceinline__ void adam_elem(const AdamCoef &c, float &p, float g, float &m, float &v) {
    // explicit roundings: the same bits wherever this is inlined (the compiler may otherwise fuse either product of a*b + c*d)
    const float gr = __fmul_rn(g, c.gscale);
    m = __fmaf_rn(c.b1, m, __fmul_rn(1.f - c.b1, gr));
    v = __fmaf_rn(c.b2, v, __fmul_rn(__fmul_rn(1.f - c.b2, gr), gr));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(c.step, m), __fmaf_rn(sqrtf(v), c.rsqrt_bc2, c.eps)));
}

// Adam over float4 elements [i0, i1) of the flat buffers + gradient zeroing + fp16 refresh from float f16_start on
__device__ __forceinline__ void adam_sweep4(const AdamCoef &c, int64_t i0, int64_t i1, int64_t first, int64_t stride, float *__restrict__ p,
                                            float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, __half *__restrict__ p16,
                                            int64_t f16_start) {
    for (int64_t i = i0 + first; i < i1; i += stride) {
        float4 P = reinterpret_cast<float4 *>(p)[i], G = reinterpret_cast<float4 *>(g)[i];
        float4 M = reinterpret_cast<float4 *>(m)[i], V = reinterpret_cast<float4 *>(v)[i];
        adam_elem(c, P.x, G.x, M.x, V.x);
        adam_elem(c, P.y, G.y, M.y, V.y);
        adam_elem(c, P.z, G.z, M.z, V.z);
        adam_elem(c, P.w, G.w, M.w, V.w);
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
}

// one CTA, 1024 threads: warp w owns rows 2w, 2w+1 of lin0 (lanes over the input features)
__global__ void __launch_bounds__(kPrepThreads) prep_net_kernel(int n_levels, const float *__restrict__ small, float *__restrict__ net,
                                                                int n_mask, const float *__restrict__ mask, float *__restrict__ stats) {
    __shared__ float s_norm[kH];
    __shared__ float s_red[32];
    const SmallLayout L(n_levels);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int row = 2 * warp + q;
        float ss = row_sumsq(small + row * L.d_in, L.d_in, lane);
        if (lane == 0) s_norm[row] = sqrtf(ss);
    }
    float part = tid < kH ? small[L.v1 + tid] * small[L.v1 + tid] : 0.f;
    float n1 = sqrtf(block_sum(part, s_red));   // also orders s_norm
    fold_store(L, small, s_norm, n1, net, tid, kPrepThreads);
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
    const SmallLayout L(n_levels);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q) unfold_row(L, 2 * warp + q, lane, small, net_grad, small_grad);
    float v1 = tid < kH ? small[L.v1 + tid] : 0.f, dw1 = tid < kH ? net_grad[kOffW1 + tid] : 0.f;
    float ss1 = block_sum(v1 * v1, s_red);
    float dot1 = block_sum(dw1 * v1, s_red);
    unfold_tail(L, tid, ss1, dot1, small, net_grad, stats[4], small_grad);
}

__global__ void __launch_bounds__(256) adam_kernel(int64_t n, float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, __half *__restrict__ p16, int64_t f16_start, AdamCoef c) {
    const int64_t n4 = n >> 2;
    adam_sweep4(c, 0, n4, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, p, g, m, v, p16, f16_start);
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_elem(c, pi, g[i], mi, vi);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        g[i] = 0.f;
        if (p16 && i >= f16_start) p16[i - f16_start] = __float2half_rn(pi);
    }
}

// ---- step tail: everything between the backward of iteration `it` and the marcher of iteration `it + 1` in ONE launch ----
//   block 0                 weight-norm backward (net_grad -> small_grad), Adam on the MLP / variance block, weight-norm fold of the
//                           UPDATED parameters into `net` for the next forward, net_grad zeroed -- all through shared memory
//   blocks 1 .. n_sampler   device patch sampler for iteration it + 1 (independent of the parameters)
//   remaining blocks        Adam over the live hash-table levels + gradient zeroing + fp16 table refresh
// Replaces unfold_grads (1 CTA) -> adam -> prep_net (1 CTA) -> sample_patches: three serial single-CTA latency chains and a
// launch overlap with the HBM sweep instead of preceding / following it.
struct TailArgs {
    AdamCoef coef;
    float *p, *g, *m, *v;
    __half *p16;
    int64_t small_pad, n_live;   // floats; n_live = live table floats (multiple of 4)
    int n_levels, do_unfold;
    float *net, *net_grad;
    const float *stats;
    int n_sampler_blocks, n_patches;
    uint64_t seed, step;
    snb_dataset ds;
    snb_batch_out out;
};
constexpr int kTailThreads = 1024;
constexpr int kSmallMax = kH * kDinMax + 3 * kH + 3;   // 2435
constexpr int kSmallIters = (kSmallMax + kTailThreads - 1) / kTailThreads;   // 3

__global__ void __launch_bounds__(kTailThreads) train_tail_kernel(const __grid_constant__ TailArgs a) {
    if (blockIdx.x > (unsigned)a.n_sampler_blocks) {   // ---- table Adam ----
        const int64_t nb = gridDim.x - 1 - a.n_sampler_blocks, b = blockIdx.x - 1 - a.n_sampler_blocks;
        const int64_t i0 = a.small_pad >> 2;
        adam_sweep4(a.coef, i0, i0 + (a.n_live >> 2), b * kTailThreads + threadIdx.x, nb * kTailThreads, a.p, a.g, a.m, a.v, a.p16,
                    a.small_pad);
        return;
    }
    if (blockIdx.x >= 1) {   // ---- sampler for the next iteration ----
        sample_patch_ray(a.ds, a.n_patches, a.seed, a.step, a.out, (blockIdx.x - 1) * kTailThreads + threadIdx.x);
        return;
    }
    // ---- block 0: MLP / variance block.  One global round trip: every load is issued before the first use. ----
    __shared__ float s_p[kSmallIters * kTailThreads], s_g[kSmallIters * kTailThreads], s_ng[kSmallIters * kTailThreads];
    __shared__ float s_norm[kH], s_red[4];
    const SmallLayout L(a.n_levels);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float pr[kSmallIters], mr[kSmallIters], vr[kSmallIters], xr[kSmallIters];   // xr: net_grad (to unfold) or the unfolded gradient
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        const bool in = e < L.n;
        pr[k] = in ? a.p[e] : 0.f;
        mr[k] = in ? a.m[e] : 0.f;
        vr[k] = in ? a.v[e] : 0.f;
        xr[k] = a.do_unfold ? (e < kNetFloats ? a.net_grad[e] : 0.f) : (in ? a.g[e] : 0.f);
    }
    const float d_inv_s = a.do_unfold ? a.stats[4] : 0.f;
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        s_p[e] = pr[k];
        (a.do_unfold ? s_ng : s_g)[e] = xr[k];
    }
    __syncthreads();
    for (int e = tid; e < kNetFloats; e += kTailThreads) a.net_grad[e] = 0.f;   // ready for the next backward
    if (a.do_unfold) {
        for (int row = warp; row < kH; row += kTailThreads / 32) unfold_row(L, row, lane, s_p, s_ng, s_g);
        float v1 = tid < kH ? s_p[L.v1 + tid] : 0.f, dw1 = tid < kH ? s_ng[kOffW1 + tid] : 0.f;
        float ss1 = warp_sum_f(v1 * v1), dot1 = warp_sum_f(dw1 * v1);
        if (warp < 2 && lane == 0) { s_red[warp] = ss1; s_red[2 + warp] = dot1; }
        __syncthreads();
        unfold_tail(L, tid, s_red[0] + s_red[1], s_red[2] + s_red[3], s_p, s_ng, d_inv_s, s_g);
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        if (e < L.n) {
            adam_elem(a.coef, pr[k], s_g[e], mr[k], vr[k]);
            s_p[e] = pr[k];
            a.p[e] = pr[k];
            a.m[e] = mr[k];
            a.v[e] = vr[k];
            a.g[e] = 0.f;
        }
    }
    __syncthreads();
    for (int row = warp; row < kH; row += kTailThreads / 32) {
        float ss = row_sumsq(s_p + row * L.d_in, L.d_in, lane);
        if (lane == 0) s_norm[row] = sqrtf(ss);
    }
    float v1n = tid < kH ? s_p[L.v1 + tid] : 0.f;
    float ssn = warp_sum_f(v1n * v1n);
    if (warp < 2 && lane == 0) s_red[warp] = ssn;
    __syncthreads();
    fold_store(L, s_p, s_norm, sqrtf(s_red[0] + s_red[1]), a.net, tid, kTailThreads);
}

// ---- data-parallel step tail over NVLink peer memory: gradient reduction + sharded Adam + parameter broadcast in ONE kernel ----
// Every rank launches this kernel after its backward.  Flat parameters / gradients / fp16 table of all ranks live in a symmetric
// (peer-mapped) allocation.  Chunk c (1024 float4) of the live table range is OWNED by rank c % world, statically, so Adam state is
// sharded and never moves:
//   owner: g = sum over ranks of grad_r[chunk] (peer loads, rank order), Adam on its own m / v / p, then stores the updated fp32
//          parameters and the fp16 copy into EVERY rank's buffers and zeroes every rank's gradient chunk (it is its only reader);
//   block 0 of every rank: sums all ranks' gradients w.r.t. the folded MLP weights (grad[0 : 2432), the inv_s gradient in slot
//          kOffInvS) in rank order -- identical bits everywhere -- and runs the usual weight-norm backward / Adam / fold locally;
//   sampler blocks as in train_tail_kernel.
// Two cross-GPU barriers through flags in the symmetric buffer: "my backward is done" (written by block 0 at kernel start, awaited
// by every reading block) and "all my remote stores are done" (written by block 0 after the local blocks have fenced and counted
// in; block 0 leaves only when every rank has said so, which is what orders the next kernel behind the peers' parameter stores).
// Replaces snb_unfold_grads -> NCCL allreduce (ring/tree over the whole range, every rank then repeating the full Adam sweep) ->
// snb_train_tail.  Per GPU and step it moves (world-1)/world of the live range once in (gradients) and 2.5x that out
// (parameters, fp16 copy, zeros) instead of the allreduce's 2x (world-1)/world in + out plus a full local Adam sweep.
constexpr int kMaxPeers = SNB_MAX_PEERS;
struct PeerArgs {
    int world, rank;
    float *param[kMaxPeers];
    float *grad[kMaxPeers];
    __half *f16[kMaxPeers];
    uint32_t *flags[kMaxPeers];   // per rank: [0, 8) start flags, [8, 16) done flags, [16] error code
    uint32_t *counter;            // local
    uint32_t epoch;
    int n_adam_blocks;
    unsigned long long *trace;   // optional u64[trace_cap][4] timeline of block 0 (snb_peer_group.trace)
    int trace_cap;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float *p) {   // system-scope load: never served from a stale L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// threads 0 .. world-1 each wait for one rank's flag, then the CTA joins.  A wait longer than 20 s (a rank died) records an error
// code instead of hanging the GPU; the host checks it when it reads the loss terms.
__device__ __forceinline__ void wait_flags(const uint32_t *flags, uint32_t *err, int world, uint32_t epoch, uint32_t code) {
    if ((int)threadIdx.x < world) {
        const unsigned long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            if (global_ns() - t0 > 20000000000ull) { atomicExch(err, code + threadIdx.x); break; }
        }
    }
    __syncthreads();
}

// kF16Only (snb_peer_group.table_f16_only, FusedTrainer's default since it was validated on 2 GPUs -- DESIGN.md section 6): the owner keeps the fp32
// master parameters of its chunks to itself and broadcasts only the fp16 copy the forward gathers from, and every rank zeroes its own
// gradient range after the done barrier instead of the owner zeroing it remotely: 2 instead of 14 bytes per owned parameter leave the GPU.
template <bool kF16Only>
__global__ void __launch_bounds__(kTailThreads, 1) train_tail_peer_kernel(const __grid_constant__ TailArgs a, const __grid_constant__ PeerArgs pg) {
    const int tid = threadIdx.x;
    uint32_t *my_flags = pg.flags[pg.rank];
    if (blockIdx.x > (unsigned)a.n_sampler_blocks) {   // ---- owned table chunks: reduce, Adam, broadcast ----
        wait_flags(my_flags, my_flags + 16, pg.world, pg.epoch, 100);
        const int64_t b = blockIdx.x - 1 - a.n_sampler_blocks;
        const int64_t base4 = a.small_pad >> 2, n4 = a.n_live >> 2, n_chunks = (n4 + kTailThreads - 1) / kTailThreads;
        for (int64_t q = b;; q += pg.n_adam_blocks) {
            const int64_t c = q * pg.world + pg.rank;
            if (c >= n_chunks) break;
            const int64_t j = c * kTailThreads + tid;   // float4 index inside the live table range
            if (j >= n4) continue;
            const int64_t i = base4 + j;
            float4 gr[kMaxPeers];
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (r < pg.world) gr[r] = ld_peer_f4(pg.grad[r] + 4 * i);
            float4 P = reinterpret_cast<float4 *>(a.p)[i], M = reinterpret_cast<float4 *>(a.m)[i], V = reinterpret_cast<float4 *>(a.v)[i];
            float4 G = gr[0];
#pragma unroll
            for (int r = 1; r < kMaxPeers; ++r)
                if (r < pg.world) { G.x += gr[r].x; G.y += gr[r].y; G.z += gr[r].z; G.w += gr[r].w; }
            adam_elem(a.coef, P.x, G.x, M.x, V.x);
            adam_elem(a.coef, P.y, G.y, M.y, V.y);
            adam_elem(a.coef, P.z, G.z, M.z, V.z);
            adam_elem(a.coef, P.w, G.w, M.w, V.w);
            reinterpret_cast<float4 *>(a.m)[i] = M;
            reinterpret_cast<float4 *>(a.v)[i] = V;
            const __half2 h0 = __floats2half2_rn(P.x, P.y), h1 = __floats2half2_rn(P.z, P.w);
            uint2 hp;
            hp.x = *reinterpret_cast<const uint32_t *>(&h0);
            hp.y = *reinterpret_cast<const uint32_t *>(&h1);
            if (kF16Only) reinterpret_cast<float4 *>(a.p)[i] = P;
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (r < pg.world) {
                    if (!kF16Only) reinterpret_cast<float4 *>(pg.param[r])[i] = P;
                    reinterpret_cast<uint2 *>(pg.f16[r])[j] = hp;
                    if (!kF16Only) reinterpret_cast<float4 *>(pg.grad[r])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) atomicAdd(pg.counter, 1u);
        if (kF16Only) {   // every rank has finished reading this rank's gradients once all done flags are up: zero them locally
            wait_flags(my_flags + kMaxPeers, my_flags + 16, pg.world, pg.epoch, 500);
            for (int64_t j = b * kTailThreads + tid; j < n4; j += (int64_t)pg.n_adam_blocks * kTailThreads)
                reinterpret_cast<float4 *>(a.g)[base4 + j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    if (blockIdx.x >= 1) {   // ---- sampler for the next iteration ----
        sample_patch_ray(a.ds, a.n_patches, a.seed, a.step, a.out, (blockIdx.x - 1) * kTailThreads + tid);
        return;
    }
    // ---- block 0: barrier master + MLP / variance block ----
    __shared__ float s_p[kSmallIters * kTailThreads], s_g[kSmallIters * kTailThreads], s_ng[kSmallIters * kTailThreads];
    __shared__ float s_norm[kH], s_red[4];
    const SmallLayout L(a.n_levels);
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long *tr = (pg.trace && pg.trace_cap > 0) ? pg.trace + 4ull * (pg.epoch % (uint32_t)pg.trace_cap) : nullptr;
    if (tid == 0 && tr) tr[0] = global_ns();
    if (tid == 0) {
        a.g[kOffInvS] = a.stats[4];   // the inv_s gradient of this rank travels in the free slot of the folded-gradient block
        __threadfence_system();
    }
    __syncthreads();
    if (tid < pg.world) st_release_sys(pg.flags[tid] + pg.rank, pg.epoch);   // "my backward is complete"
    float pr[kSmallIters], mr[kSmallIters], vr[kSmallIters], xr[kSmallIters];
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        const bool in = e < L.n;
        pr[k] = in ? a.p[e] : 0.f;
        mr[k] = in ? a.m[e] : 0.f;
        vr[k] = in ? a.v[e] : 0.f;
    }
    wait_flags(my_flags, my_flags + 16, pg.world, pg.epoch, 200);
    if (tid == 0 && tr) tr[1] = global_ns();
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        float part[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < pg.world && e < kNetFloats) part[r] = ld_peer_f(pg.grad[r] + e);
        float acc = 0.f;
        if (e < kNetFloats) {
            acc = part[0];
#pragma unroll
            for (int r = 1; r < kMaxPeers; ++r)
                if (r < pg.world) acc += part[r];
        }
        xr[k] = acc;
    }
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        s_p[e] = pr[k];
        s_ng[e] = xr[k];
    }
    __syncthreads();
    const float d_inv_s = s_ng[kOffInvS];
    for (int row = warp; row < kH; row += kTailThreads / 32) unfold_row(L, row, lane, s_p, s_ng, s_g);
    {
        float v1 = tid < kH ? s_p[L.v1 + tid] : 0.f, dw1 = tid < kH ? s_ng[kOffW1 + tid] : 0.f;
        float ss1 = warp_sum_f(v1 * v1), dot1 = warp_sum_f(dw1 * v1);
        if (warp < 2 && lane == 0) { s_red[warp] = ss1; s_red[2 + warp] = dot1; }
    }
    __syncthreads();
    unfold_tail(L, tid, s_red[0] + s_red[1], s_red[2] + s_red[3], s_p, s_ng, d_inv_s, s_g);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSmallIters; ++k) {
        const int e = tid + k * kTailThreads;
        if (e < L.n) {
            adam_elem(a.coef, pr[k], s_g[e], mr[k], vr[k]);
            s_p[e] = pr[k];
            a.p[e] = pr[k];
            a.m[e] = mr[k];
            a.v[e] = vr[k];
        }
    }
    __syncthreads();
    for (int row = warp; row < kH; row += kTailThreads / 32) {
        float ss = row_sumsq(s_p + row * L.d_in, L.d_in, lane);
        if (lane == 0) s_norm[row] = sqrtf(ss);
    }
    {
        float v1n = tid < kH ? s_p[L.v1 + tid] : 0.f;
        float ssn = warp_sum_f(v1n * v1n);
        if (warp < 2 && lane == 0) s_red[warp] = ssn;
    }
    __syncthreads();
    fold_store(L, s_p, s_norm, sqrtf(s_red[0] + s_red[1]), a.net, tid, kTailThreads);
    // ---- all local chunk blocks have fenced their remote stores -> tell every rank; leave when every rank has told us ----
    if (tid == 0) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_gpu(pg.counter) < (uint32_t)pg.n_adam_blocks) {
            if (global_ns() - t0 > 20000000000ull) { atomicExch(my_flags + 16, 300u); break; }
        }
        *pg.counter = 0u;
        __threadfence_system();
        if (tr) tr[2] = global_ns();
    }
    __syncthreads();
    if (tid < pg.world) st_release_sys(pg.flags[tid] + kMaxPeers + pg.rank, pg.epoch);
    wait_flags(my_flags + kMaxPeers, my_flags + 16, pg.world, pg.epoch, 400);
    if (tid == 0 && tr) tr[3] = global_ns();
    for (int e = tid; e < kNetFloats; e += kTailThreads) a.g[e] = 0.f;   // every rank has read this block: ready for the next backward
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
    const AdamCoef coef{beta1, beta2, eps, lr / (float)bc1, (float)(1.0 / sqrt(bc2)), grad_scale};
    adam_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(n, param, grad, exp_avg, exp_avg_sq, (__half *)param_f16, f16_start, coef);
    SNB_LAUNCH_CHECK("adam_step");
    return SNB_OK;
}
}  // namespace snb

extern "C" int32_t snb_adam_step(int64_t n, float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *param_f16, float lr,
                                 float beta1, float beta2, float eps, int32_t step_count, float grad_scale, snb_stream_t stream) {
    return snb::adam_launch(n, param, grad, exp_avg, exp_avg_sq, param_f16, 0, lr, beta1, beta2, eps, step_count, grad_scale, stream);
}

extern "C" int32_t snb_train_tail(const snb_train_ctx *c, float lr, int32_t step_count, float grad_scale, int32_t grads_unfolded,
                                  const snb_dataset *ds_next, int32_t n_patches_next, uint64_t seed, uint64_t next_step,
                                  const snb_batch_out *out_next, snb_stream_t stream) {
    SNB_REQUIRE(c, SNB_ERR_NULL, "train_tail: null ctx");
    SNB_REQUIRE(c->flat_param && c->flat_grad && c->exp_avg && c->exp_avg_sq && c->net_grad && c->stats && c->net.net && c->net.table_f16,
                SNB_ERR_NULL, "train_tail: null buffer");
    SNB_REQUIRE(step_count >= 1 && c->n_levels >= 1 && c->n_levels <= SNB_MAX_LEVELS && c->net.n_active <= (uint32_t)c->n_levels, SNB_ERR_ARG,
                "train_tail: bad step_count / level counts");
    SNB_REQUIRE(c->small_pad % 4 == 0 && c->small_pad >= kSmallMax, SNB_ERR_ARG, "train_tail: small_pad must be a multiple of 4 and >= 2435");
    SNB_REQUIRE(aligned(c->flat_param, 16) && aligned(c->flat_grad, 16) && aligned(c->exp_avg, 16) && aligned(c->exp_avg_sq, 16) &&
                    aligned(c->net.table_f16, 8), SNB_ERR_ALIGN, "train_tail: buffers must be 16-byte aligned");
    TailArgs a{};
    const float beta1 = 0.9f, beta2 = 0.999f;   // bias corrections from the fp32 betas, like snb_adam_step / adam_launch
    const double bc1 = 1.0 - pow((double)beta1, step_count), bc2 = 1.0 - pow((double)beta2, step_count);
    a.coef = AdamCoef{beta1, beta2, 1e-8f, lr / (float)bc1, (float)(1.0 / sqrt(bc2)), grad_scale};
    a.p = c->flat_param; a.g = c->flat_grad; a.m = c->exp_avg; a.v = c->exp_avg_sq;
    a.p16 = (__half *)const_cast<void *>(c->net.table_f16);
    a.small_pad = c->small_pad;
    a.n_live = 2 * (int64_t)c->net.meta.offsets[c->net.n_active];   // levels >= n_active: zero gradient and state, skipping is exact
    a.n_levels = c->n_levels;
    a.do_unfold = grads_unfolded ? 0 : 1;
    a.net = const_cast<float *>(c->net.net);
    a.net_grad = c->net_grad;
    a.stats = c->stats;
    if (ds_next && n_patches_next > 0) {
        SNB_REQUIRE(out_next, SNB_ERR_NULL, "train_tail: null sampler output");
        SNB_REQUIRE(ds_next->W > 3 && ds_next->H > 3 && ds_next->n_train > 0 && ds_next->n_images > 0, SNB_ERR_ARG, "train_tail: bad dataset sizes");
        SNB_REQUIRE(ds_next->normals && ds_next->masks && ds_next->intrinsics_inv && ds_next->pose && ds_next->train_ids,
                    SNB_ERR_NULL, "train_tail: null dataset tensor");
        SNB_REQUIRE(out_next->rays_o && out_next->rays_d && out_next->plane_n && out_next->near_ && out_next->far_ && out_next->v_inv &&
                        out_next->normal_gt && out_next->mask, SNB_ERR_NULL, "train_tail: null sampler output tensor");
        a.ds = *ds_next;
        a.out = *out_next;
        a.n_patches = n_patches_next;
        a.seed = seed;
        a.step = next_step;
        a.n_sampler_blocks = (int)cdiv((int64_t)n_patches_next * SNB_PATCH, kTailThreads);
    }
    int64_t adam_blocks = cdiv(a.n_live / 4, kTailThreads);
    if (adam_blocks > kNumSMs * 2) adam_blocks = kNumSMs * 2;
    train_tail_kernel<<<(unsigned)(1 + a.n_sampler_blocks + adam_blocks), kTailThreads, 0, S(stream)>>>(a);
    SNB_LAUNCH_CHECK("train_tail");
    return SNB_OK;
}

extern "C" int32_t snb_train_tail_peer(const snb_train_ctx *c, const snb_peer_group *pgrp, float lr, int32_t step_count,
                                       const snb_dataset *ds_next, int32_t n_patches_next, uint64_t seed, uint64_t next_step,
                                       const snb_batch_out *out_next, snb_stream_t stream) {
    SNB_REQUIRE(c && pgrp, SNB_ERR_NULL, "train_tail_peer: null ctx / peer group");
    SNB_REQUIRE(pgrp->world >= 1 && pgrp->world <= SNB_MAX_PEERS && pgrp->rank >= 0 && pgrp->rank < pgrp->world, SNB_ERR_ARG,
                "train_tail_peer: bad world / rank");
    SNB_REQUIRE(c->flat_param && c->flat_grad && c->exp_avg && c->exp_avg_sq && c->stats && c->net.net && c->net.table_f16 && pgrp->counter,
                SNB_ERR_NULL, "train_tail_peer: null buffer");
    for (int r = 0; r < pgrp->world; ++r)
        SNB_REQUIRE(pgrp->param[r] && pgrp->grad[r] && pgrp->table_f16[r] && pgrp->flags[r], SNB_ERR_NULL, "train_tail_peer: null peer pointer");
    SNB_REQUIRE(pgrp->param[pgrp->rank] == c->flat_param && pgrp->grad[pgrp->rank] == c->flat_grad && c->net_grad == c->flat_grad &&
                    pgrp->table_f16[pgrp->rank] == c->net.table_f16, SNB_ERR_ARG,
                "train_tail_peer: the context must use this rank's symmetric buffers, with net_grad aliased to flat_grad");
    SNB_REQUIRE(step_count >= 1 && c->n_levels >= 1 && c->n_levels <= SNB_MAX_LEVELS && c->net.n_active <= (uint32_t)c->n_levels, SNB_ERR_ARG,
                "train_tail_peer: bad step_count / level counts");
    SNB_REQUIRE(c->small_pad % 4 == 0 && c->small_pad >= kSmallMax, SNB_ERR_ARG, "train_tail_peer: small_pad must be a multiple of 4 and >= 2435");
    for (int r = 0; r < pgrp->world; ++r)
        SNB_REQUIRE(aligned(pgrp->param[r], 16) && aligned(pgrp->grad[r], 16) && aligned(pgrp->table_f16[r], 8) && aligned(pgrp->flags[r], 4),
                    SNB_ERR_ALIGN, "train_tail_peer: peer buffers must be 16-byte aligned");
    SNB_REQUIRE(aligned(c->exp_avg, 16) && aligned(c->exp_avg_sq, 16), SNB_ERR_ALIGN, "train_tail_peer: buffers must be 16-byte aligned");
    TailArgs a{};
    const float beta1 = 0.9f, beta2 = 0.999f;
    const double bc1 = 1.0 - pow((double)beta1, step_count), bc2 = 1.0 - pow((double)beta2, step_count);
    a.coef = AdamCoef{beta1, beta2, 1e-8f, lr / (float)bc1, (float)(1.0 / sqrt(bc2)), 1.0f / (float)pgrp->world};
    a.p = c->flat_param; a.g = c->flat_grad; a.m = c->exp_avg; a.v = c->exp_avg_sq;
    a.p16 = (__half *)const_cast<void *>(c->net.table_f16);
    a.small_pad = c->small_pad;
    a.n_live = 2 * (int64_t)c->net.meta.offsets[c->net.n_active];
    a.n_levels = c->n_levels;
    a.do_unfold = 1;
    a.net = const_cast<float *>(c->net.net);
    a.net_grad = c->net_grad;
    a.stats = c->stats;
    if (ds_next && n_patches_next > 0) {
        SNB_REQUIRE(out_next, SNB_ERR_NULL, "train_tail_peer: null sampler output");
        SNB_REQUIRE(ds_next->W > 3 && ds_next->H > 3 && ds_next->n_train > 0 && ds_next->n_images > 0, SNB_ERR_ARG, "train_tail_peer: bad dataset sizes");
        SNB_REQUIRE(ds_next->normals && ds_next->masks && ds_next->intrinsics_inv && ds_next->pose && ds_next->train_ids,
                    SNB_ERR_NULL, "train_tail_peer: null dataset tensor");
        SNB_REQUIRE(out_next->rays_o && out_next->rays_d && out_next->plane_n && out_next->near_ && out_next->far_ && out_next->v_inv &&
                        out_next->normal_gt && out_next->mask, SNB_ERR_NULL, "train_tail_peer: null sampler output tensor");
        a.ds = *ds_next;
        a.out = *out_next;
        a.n_patches = n_patches_next;
        a.seed = seed;
        a.step = next_step;
        a.n_sampler_blocks = (int)cdiv((int64_t)n_patches_next * SNB_PATCH, kTailThreads);
    }
    SNB_REQUIRE(a.n_sampler_blocks < kNumSMs - 8, SNB_ERR_ARG, "train_tail_peer: too many patches for the in-kernel sampler");
    PeerArgs pa{};
    pa.world = pgrp->world;
    pa.rank = pgrp->rank;
    for (int r = 0; r < pgrp->world; ++r) {
        pa.param[r] = pgrp->param[r];
        pa.grad[r] = pgrp->grad[r];
        pa.f16[r] = (__half *)pgrp->table_f16[r];
        pa.flags[r] = pgrp->flags[r];
    }
    pa.counter = pgrp->counter;
    pa.epoch = pgrp->epoch ? pgrp->epoch : (uint32_t)step_count;
    pa.trace = (unsigned long long *)pgrp->trace;
    pa.trace_cap = pgrp->trace_capacity;
    // one CTA per SM at most (every waiting CTA is resident): block 0 + sampler blocks + chunk blocks <= 148
    const int64_t n_chunks = cdiv(a.n_live / 4, (int64_t)kTailThreads);
    int64_t own_chunks = cdiv(n_chunks, (int64_t)pgrp->world);
    int64_t adam_blocks = own_chunks < 1 ? 1 : own_chunks;
    if (adam_blocks > kNumSMs - 1 - a.n_sampler_blocks) adam_blocks = kNumSMs - 1 - a.n_sampler_blocks;
    pa.n_adam_blocks = (int)adam_blocks;
    if (pgrp->table_f16_only)
        train_tail_peer_kernel<true><<<(unsigned)(1 + a.n_sampler_blocks + adam_blocks), kTailThreads, 0, S(stream)>>>(a, pa);
    else
        train_tail_peer_kernel<false><<<(unsigned)(1 + a.n_sampler_blocks + adam_blocks), kTailThreads, 0, S(stream)>>>(a, pa);
    SNB_LAUNCH_CHECK("train_tail_peer");
    return SNB_OK;
}
