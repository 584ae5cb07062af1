// occgrid.cu -- occupancy-grid EMA update (NA/grid.py:197-239) as three small kernels:
// jittered cell points (with the AABB contract_inv of CS/include/helpers_contraction.h:23-28 fused
// in), EMA-max scatter, and mean/threshold binarisation.
#include "common.cuh"

namespace snb {

__global__ void occgrid_points_kernel(int64_t n, const int64_t *__restrict__ indices, const float *__restrict__ rnd,
                                      int3 res, const float *__restrict__ roi, float *__restrict__ x) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c = indices ? indices[i] : i;
    int cz = (int)(c % res.z);
    int cy = (int)((c / res.z) % res.y);
    int cx = (int)(c / ((int64_t)res.z * res.y));
    int cc[3] = {cx, cy, cz};
    int rr[3] = {res.x, res.y, res.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // (coord + rand) / res, then x*(max-min)+min -- separately rounded like the ATen ops + contract_inv
        float u = __fdiv_rn(__fadd_rn((float)cc[d], rnd[3 * i + d]), (float)rr[d]);
        float lo = __ldg(roi + d), hi = __ldg(roi + 3 + d);
        x[3 * i + d] = __fmaf_rn(u, __fsub_rn(hi, lo), lo);
    }
}

__global__ void occgrid_ema_kernel(int64_t n, const int64_t *__restrict__ indices, const float *__restrict__ occ,
                                   float decay, float *__restrict__ occs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c = indices ? indices[i] : i;
    // duplicates in `indices`: last writer wins, as in the reference's index assignment (NA/grid.py:232)
    occs[c] = fmaxf(__fmul_rn(occs[c], decay), occ[i]);
}

__global__ void occgrid_sum_kernel(int64_t n, const float *__restrict__ occs, double *__restrict__ sum) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += occs[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
        atomicAdd(sum, t);
    }
}

__global__ void occgrid_threshold_kernel(int64_t n, const float *__restrict__ occs, const double *__restrict__ sum,
                                         float thre, uint8_t *__restrict__ binary) {
    float mean = (float)(*sum / (double)n);
    float th = fminf(mean, thre);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        binary[i] = occs[i] > th ? 1 : 0;
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_occgrid_points(int64_t n, const int64_t *indices, const float *rnd, int32_t rx, int32_t ry, int32_t rz,
                                      const float *roi, float *x, snb_stream_t stream) {
    SNB_REQUIRE(n >= 0 && rx > 0 && ry > 0 && rz > 0, SNB_ERR_ARG, "occgrid_points: bad sizes");
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(rnd && roi && x, SNB_ERR_NULL, "occgrid_points: null buffer");
    occgrid_points_kernel<<<(unsigned)cdiv(n, 256), 256, 0, S(stream)>>>(n, indices, rnd, make_int3(rx, ry, rz), roi, x);
    SNB_LAUNCH_CHECK("occgrid_points");
    return SNB_OK;
}

extern "C" int32_t snb_occgrid_ema(int64_t n, const int64_t *indices, const float *occ, float decay, float *occs, snb_stream_t stream) {
    SNB_REQUIRE(n >= 0, SNB_ERR_ARG, "occgrid_ema: n < 0");
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(occ && occs, SNB_ERR_NULL, "occgrid_ema: null buffer");
    occgrid_ema_kernel<<<(unsigned)cdiv(n, 256), 256, 0, S(stream)>>>(n, indices, occ, decay, occs);
    SNB_LAUNCH_CHECK("occgrid_ema");
    return SNB_OK;
}

extern "C" int32_t snb_occgrid_binarize(int64_t num_cells, const float *occs, float occ_thre, uint8_t *binary, void *workspace,
                                        snb_stream_t stream) {
    SNB_REQUIRE(num_cells > 0, SNB_ERR_ARG, "occgrid_binarize: num_cells <= 0");
    SNB_REQUIRE(occs && binary && workspace, SNB_ERR_NULL, "occgrid_binarize: null buffer");
    SNB_REQUIRE(aligned(workspace, 8), SNB_ERR_ALIGN, "occgrid_binarize: workspace must be 8-byte aligned");
    cudaMemsetAsync(workspace, 0, 8, S(stream));
    occgrid_sum_kernel<<<kNumSMs * 4, 256, 0, S(stream)>>>(num_cells, occs, (double *)workspace);
    occgrid_threshold_kernel<<<kNumSMs * 4, 256, 0, S(stream)>>>(num_cells, occs, (const double *)workspace, occ_thre, binary);
    SNB_LAUNCH_CHECK("occgrid_binarize");
    return SNB_OK;
}
