// umma_test.cu -- self-test of the tcgen05 plumbing in umma.cuh: D[128 x N] = A[128 x K] * B[N x K]^T with kind::tf32,
// operands written to shared memory by the threads (no TMA), accumulator read back thread-per-row.
#include "common.cuh"
#include "umma.cuh"

namespace snb {

// LBOA: leading-dimension byte offset of the A tile (128 = dense cores, 192 = the padded layout of sdf_bwd_mlp_umma_kernel)
template <int LBOA>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D,
                                                               int K, int N, int *__restrict__ err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *a_s = smem, *b_s = smem + umma::tile_bytes_lbo<LBOA>(128, K);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 128 * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float *>(a_s + umma::kmajor_off_lbo<LBOA>(r, k, K)) = A[e];
    }
    for (int e = tid; e < N * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float *>(b_s + umma::kmajor_off(r, k, K)) = B[e];
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_smem_to_async_proxy();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t a0 = umma::smem_u32(a_s), b0 = umma::smem_u32(b_s);
        for (int s = 0; s < K / 8; ++s)
            umma::mma_tf32(tmem, umma::kmajor_desc_lbo<LBOA>(a0 + 2u * LBOA * s, K), umma::kmajor_desc(b0 + 256u * s, K), idesc, s > 0);
        umma::commit(&bar);
    }
    const bool ok = umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (!ok) {
        if (tid == 0) atomicExch(err, 1);
    } else {
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            umma::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
            umma::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace snb
using namespace snb;

static int32_t selftest_launch(const float *A, const float *B, float *D, int32_t K, int32_t N, int32_t lbo_a, int32_t *err, snb_stream_t stream) {
    SNB_REQUIRE(A && B && D && err, SNB_ERR_NULL, "umma_selftest: null buffer");
    SNB_REQUIRE(K > 0 && K % 8 == 0 && K <= 128 && (N == 32 || N == 64), SNB_ERR_ARG, "umma_selftest: K must be a multiple of 8 (<= 128), N 32 or 64");
    SNB_REQUIRE(lbo_a == 128 || lbo_a == 192, SNB_ERR_ARG, "umma_selftest: lbo_a must be 128 or 192");
    if (lbo_a == 128) {
        const size_t smem = umma::tile_bytes(128, K) + umma::tile_bytes(N, K);
        cudaFuncSetAttribute(umma_selftest_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        umma_selftest_kernel<128><<<1, 128, smem, S(stream)>>>(A, B, D, K, N, err);
    } else {
        const size_t smem = umma::tile_bytes_lbo<192>(128, K) + umma::tile_bytes(N, K);
        cudaFuncSetAttribute(umma_selftest_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        umma_selftest_kernel<192><<<1, 128, smem, S(stream)>>>(A, B, D, K, N, err);
    }
    SNB_LAUNCH_CHECK("umma_selftest");
    return SNB_OK;
}

extern "C" int32_t snb_umma_selftest(const float *A, const float *B, float *D, int32_t K, int32_t N, int32_t *err, snb_stream_t stream) {
    return selftest_launch(A, B, D, K, N, 128, err, stream);
}

extern "C" int32_t snb_umma_selftest_lbo(const float *A, const float *B, float *D, int32_t K, int32_t N, int32_t lbo_a, int32_t *err,
                                         snb_stream_t stream) {
    return selftest_launch(A, B, D, K, N, lbo_a, err, stream);
}
