// hashgrid.cu -- drop-in multiresolution hash-grid encoding: forward, table gradient, input
// gradient and the double backward that analytic SDF normals need (boundary B).
//
// One thread owns one point and walks the active levels; the 8 corner loads of a level are issued
// back to back (independent addresses) so each thread keeps 8+ L2 requests in flight, and the whole
// fp16 output row is written with 8-byte stores.  The table (<= 28 MB for T=2^19) is L2-resident on
// B200 (126 MB L2), so the gathers are L2 traffic, not HBM.
#include "hashgrid.cuh"

namespace snb {

template <bool OUT_F32>
__global__ void __launch_bounds__(256) hashgrid_fwd_kernel(int64_t n, const float *__restrict__ x,
                                                           const __half2 *__restrict__ table, snb_hashgrid_meta m,
                                                           uint32_t n_active, void *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
    const uint32_t L = m.n_levels;
    for (uint32_t l = 0; l < L; ++l) {
        __half2 r = __float2half2_rn(0.f);
        if (l < n_active) {
            LevelCtx c = level_ctx(m, l);
            Cell cell = cell_of(c, px, py, pz);
            r = interp_level(c, cell, table);
        }
        if (OUT_F32) {
            reinterpret_cast<float2 *>(out)[i * L + l] = __half22float2(r);
        } else {
            reinterpret_cast<__half2 *>(out)[i * L + l] = r;
        }
    }
}

template <bool DY_F32>
__device__ __forceinline__ float2 load_dy(const void *dy, int64_t i, uint32_t L, uint32_t l) {
    if (DY_F32) return __ldg(reinterpret_cast<const float2 *>(dy) + i * L + l);
    return __half22float2(__ldg(reinterpret_cast<const __half2 *>(dy) + i * L + l));
}

template <bool DY_F32>
__global__ void __launch_bounds__(256) hashgrid_bwd_table_kernel(int64_t n, const float *__restrict__ x,
                                                                 const void *__restrict__ dy, float dy_scale,
                                                                 snb_hashgrid_meta m, uint32_t n_active,
                                                                 float *__restrict__ grad) {
    // thread = (point, level), level-major blocks (blockIdx.y = level) so one block hits one level's table
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t l = blockIdx.y;
    if (i >= n || l >= n_active) return;
    float2 g = load_dy<DY_F32>(dy, i, m.n_levels, l);
    g.x *= dy_scale;
    g.y *= dy_scale;
    if (g.x == 0.f && g.y == 0.f) return;
    LevelCtx c = level_ctx(m, l);
    Cell cell = cell_of(c, __ldg(x + 3 * i), __ldg(x + 3 * i + 1), __ldg(x + 3 * i + 2));
    float2 *gt = reinterpret_cast<float2 *>(grad) + c.offset;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        float w = corner_weight(cell, k);
        atomicAdd(gt + corner_index(c, cell, k), make_float2(w * g.x, w * g.y));
    }
}

// sign of corner k along dim d, and the product of the other two dims' weights
__device__ __forceinline__ float sgn(uint32_t k, int d) { return ((k >> d) & 1u) ? 1.f : -1.f; }
__device__ __forceinline__ float a_of(const Cell &cell, uint32_t k, int d) { return ((k >> d) & 1u) ? cell.w[d] : 1.f - cell.w[d]; }

template <bool DY_F32>
__global__ void __launch_bounds__(256) hashgrid_bwd_input_kernel(int64_t n, const float *__restrict__ x,
                                                                 const void *__restrict__ dy,
                                                                 const __half2 *__restrict__ table, snb_hashgrid_meta m,
                                                                 uint32_t n_active, float *__restrict__ dx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
    float r[3] = {0.f, 0.f, 0.f};
    for (uint32_t l = 0; l < n_active && l < m.n_levels; ++l) {
        float2 g = load_dy<DY_F32>(dy, i, m.n_levels, l);
        LevelCtx c = level_ctx(m, l);
        Cell cell = cell_of(c, px, py, pz);
        const __half2 *t = table + c.offset;
        float2 v[8];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) v[k] = __half22float2(__ldg(t + corner_index(c, cell, k)));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int e0 = (d + 1) % 3, e1 = (d + 2) % 3;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                float wk = sgn(k, d) * a_of(cell, k, e0) * a_of(cell, k, e1);
                s0 += wk * v[k].x;
                s1 += wk * v[k].y;
            }
            r[d] += c.scale * (g.x * s0 + g.y * s1);
        }
    }
    dx[3 * i] = r[0];
    dx[3 * i + 1] = r[1];
    dx[3 * i + 2] = r[2];
}

template <bool DY_F32>
__global__ void __launch_bounds__(256) hashgrid_bwd_bwd_input_kernel(int64_t n, const float *__restrict__ x,
                                                                     const float *__restrict__ g2,
                                                                     const void *__restrict__ dy,
                                                                     const __half2 *__restrict__ table,
                                                                     snb_hashgrid_meta m, uint32_t n_active,
                                                                     float *__restrict__ grad, float *__restrict__ d_dy,
                                                                     float *__restrict__ dx2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
    float h[3] = {__ldg(g2 + 3 * i), __ldg(g2 + 3 * i + 1), __ldg(g2 + 3 * i + 2)};
    float r2[3] = {0.f, 0.f, 0.f};
    const uint32_t L = m.n_levels;
    for (uint32_t l = 0; l < L; ++l) {
        if (l >= n_active) {
            if (d_dy) reinterpret_cast<float2 *>(d_dy)[i * L + l] = make_float2(0.f, 0.f);
            continue;
        }
        float2 g = load_dy<DY_F32>(dy, i, L, l);
        LevelCtx c = level_ctx(m, l);
        Cell cell = cell_of(c, px, py, pz);
        const __half2 *t = table + c.offset;
        float2 *gt = reinterpret_cast<float2 *>(grad) + c.offset;
        float2 ddy = make_float2(0.f, 0.f);
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            uint32_t idx = corner_index(c, cell, k);
            float2 v = __half22float2(__ldg(t + idx));
            float a[3] = {a_of(cell, k, 0), a_of(cell, k, 1), a_of(cell, k, 2)};
            float s[3] = {sgn(k, 0), sgn(k, 1), sgn(k, 2)};
            // first-order coefficient of this corner along g2: sum_d h_d * scale * s_d * prod_{e!=d} a_e
            float c1 = c.scale * (h[0] * s[0] * a[1] * a[2] + h[1] * s[1] * a[0] * a[2] + h[2] * s[2] * a[0] * a[1]);
            ddy.x += c1 * v.x;
            ddy.y += c1 * v.y;
            if (grad && (g.x != 0.f || g.y != 0.f)) atomicAdd(gt + idx, make_float2(c1 * g.x, c1 * g.y));
            if (dx2) {
                float gv = g.x * v.x + g.y * v.y;
                float s2 = c.scale * c.scale * gv;
                // d/dx_e of sum_d h_d dy/dx_d, cross terms only (linear interpolation)
                r2[0] += s2 * s[0] * (h[1] * s[1] * a[2] + h[2] * s[2] * a[1]);
                r2[1] += s2 * s[1] * (h[0] * s[0] * a[2] + h[2] * s[2] * a[0]);
                r2[2] += s2 * s[2] * (h[0] * s[0] * a[1] + h[1] * s[1] * a[0]);
            }
        }
        if (d_dy) reinterpret_cast<float2 *>(d_dy)[i * L + l] = ddy;
    }
    if (dx2) {
        dx2[3 * i] = r2[0];
        dx2[3 * i + 1] = r2[1];
        dx2[3 * i + 2] = r2[2];
    }
}

__global__ void cast_f32_f16_kernel(int64_t n2, const float2 *__restrict__ src, __half2 *__restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __float22half2_rn(src[i]);
}

static int32_t check_grid(int64_t n, const void *x, const snb_hashgrid_meta *m, const char *who) {
    SNB_REQUIRE(n >= 0, SNB_ERR_ARG, "%s: n < 0", who);
    SNB_REQUIRE(m, SNB_ERR_NULL, "%s: null meta", who);
    SNB_REQUIRE(m->n_levels >= 1 && m->n_levels <= SNB_MAX_LEVELS, SNB_ERR_ARG, "%s: n_levels=%u out of [1,%d]", who, m->n_levels, SNB_MAX_LEVELS);
    SNB_REQUIRE(n == 0 || x, SNB_ERR_NULL, "%s: null positions", who);
    return SNB_OK;
}

}  // namespace snb
using namespace snb;

extern "C" uint32_t snb_hashgrid_make_meta(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution,
                                           float per_level_scale, snb_hashgrid_meta *h_meta) {
    if (!h_meta || n_levels < 1 || n_levels > SNB_MAX_LEVELS || log2_hashmap_size > 30) {
        set_error("hashgrid_make_meta: bad arguments (n_levels=%u log2T=%u)", n_levels, log2_hashmap_size);
        return 0;
    }
    // GridEncodingTemplated ctor + grid_scale/grid_resolution (tcnn), host libm
    const float l2 = log2f(per_level_scale);
    uint32_t offset = 0;
    h_meta->n_levels = n_levels;
    for (uint32_t i = 0; i < n_levels; ++i) {
        float scale = exp2f((float)i * l2) * (float)base_resolution - 1.0f;
        uint32_t res = (uint32_t)ceilf(scale) + 1u;
        uint32_t max_params = UINT32_MAX / 2;
        uint32_t np = powf((float)res, 3.0f) > (float)max_params ? max_params : res * res * res;
        np = (np + 7u) / 8u * 8u;
        uint32_t cap = 1u << log2_hashmap_size;
        if (np > cap) np = cap;
        h_meta->offsets[i] = offset;
        h_meta->scales[i] = scale;
        h_meta->resolutions[i] = res;
        offset += np;
    }
    for (uint32_t i = n_levels; i <= SNB_MAX_LEVELS; ++i) h_meta->offsets[i] = offset;
    return offset;
}

extern "C" int32_t snb_cast_f32_to_f16(int64_t n, const float *src, void *dst, snb_stream_t stream) {
    SNB_REQUIRE(n >= 0 && (n % 2) == 0, SNB_ERR_ARG, "cast_f32_to_f16: n must be even and >= 0");
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(src && dst, SNB_ERR_NULL, "cast_f32_to_f16: null buffer");
    SNB_REQUIRE(aligned(src, 8) && aligned(dst, 4), SNB_ERR_ALIGN, "cast_f32_to_f16: misaligned buffer");
    cast_f32_f16_kernel<<<kNumSMs * 8, 256, 0, S(stream)>>>(n / 2, (const float2 *)src, (__half2 *)dst);
    SNB_LAUNCH_CHECK("cast_f32_to_f16");
    return SNB_OK;
}

extern "C" int32_t snb_hashgrid_fwd(int64_t n, const float *x, const void *table, const snb_hashgrid_meta *m, uint32_t n_active,
                                    void *out, int32_t out_is_f32, snb_stream_t stream) {
    int32_t rc = check_grid(n, x, m, "hashgrid_fwd");
    if (rc) return rc;
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(table && out, SNB_ERR_NULL, "hashgrid_fwd: null table/out");
    SNB_REQUIRE(aligned(table, 4) && aligned(out, out_is_f32 ? 8 : 4), SNB_ERR_ALIGN, "hashgrid_fwd: misaligned buffer");
    unsigned blocks = (unsigned)cdiv(n, 256);
    if (out_is_f32)
        hashgrid_fwd_kernel<true><<<blocks, 256, 0, S(stream)>>>(n, x, (const __half2 *)table, *m, n_active, out);
    else
        hashgrid_fwd_kernel<false><<<blocks, 256, 0, S(stream)>>>(n, x, (const __half2 *)table, *m, n_active, out);
    SNB_LAUNCH_CHECK("hashgrid_fwd");
    return SNB_OK;
}

extern "C" int32_t snb_hashgrid_bwd_table(int64_t n, const float *x, const void *dL_dy, int32_t dy_is_f32, float dy_scale,
                                          const snb_hashgrid_meta *m, uint32_t n_active, float *table_grad, snb_stream_t stream) {
    int32_t rc = check_grid(n, x, m, "hashgrid_bwd_table");
    if (rc) return rc;
    if (n == 0 || n_active == 0) return SNB_OK;
    SNB_REQUIRE(dL_dy && table_grad, SNB_ERR_NULL, "hashgrid_bwd_table: null buffer");
    SNB_REQUIRE(aligned(table_grad, 8) && aligned(dL_dy, dy_is_f32 ? 8 : 4), SNB_ERR_ALIGN, "hashgrid_bwd_table: misaligned buffer");
    dim3 grid((unsigned)cdiv(n, 256), n_active < m->n_levels ? n_active : m->n_levels);
    if (dy_is_f32)
        hashgrid_bwd_table_kernel<true><<<grid, 256, 0, S(stream)>>>(n, x, dL_dy, dy_scale, *m, n_active, table_grad);
    else
        hashgrid_bwd_table_kernel<false><<<grid, 256, 0, S(stream)>>>(n, x, dL_dy, dy_scale, *m, n_active, table_grad);
    SNB_LAUNCH_CHECK("hashgrid_bwd_table");
    return SNB_OK;
}

extern "C" int32_t snb_hashgrid_bwd_input(int64_t n, const float *x, const void *dL_dy, int32_t dy_is_f32, const void *table,
                                          const snb_hashgrid_meta *m, uint32_t n_active, float *dL_dx, snb_stream_t stream) {
    int32_t rc = check_grid(n, x, m, "hashgrid_bwd_input");
    if (rc) return rc;
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(dL_dy && table && dL_dx, SNB_ERR_NULL, "hashgrid_bwd_input: null buffer");
    unsigned blocks = (unsigned)cdiv(n, 256);
    if (dy_is_f32)
        hashgrid_bwd_input_kernel<true><<<blocks, 256, 0, S(stream)>>>(n, x, dL_dy, (const __half2 *)table, *m, n_active, dL_dx);
    else
        hashgrid_bwd_input_kernel<false><<<blocks, 256, 0, S(stream)>>>(n, x, dL_dy, (const __half2 *)table, *m, n_active, dL_dx);
    SNB_LAUNCH_CHECK("hashgrid_bwd_input");
    return SNB_OK;
}

extern "C" int32_t snb_hashgrid_bwd_bwd_input(int64_t n, const float *x, const float *g2, const void *dL_dy, int32_t dy_is_f32,
                                              const void *table, const snb_hashgrid_meta *m, uint32_t n_active,
                                              float *table_grad, float *d_dL_dy, float *dx2, snb_stream_t stream) {
    int32_t rc = check_grid(n, x, m, "hashgrid_bwd_bwd_input");
    if (rc) return rc;
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(g2 && dL_dy && table, SNB_ERR_NULL, "hashgrid_bwd_bwd_input: null buffer");
    unsigned blocks = (unsigned)cdiv(n, 256);
    if (dy_is_f32)
        hashgrid_bwd_bwd_input_kernel<true><<<blocks, 256, 0, S(stream)>>>(n, x, g2, dL_dy, (const __half2 *)table, *m, n_active,
                                                                          table_grad, d_dL_dy, dx2);
    else
        hashgrid_bwd_bwd_input_kernel<false><<<blocks, 256, 0, S(stream)>>>(n, x, g2, dL_dy, (const __half2 *)table, *m, n_active,
                                                                           table_grad, d_dL_dy, dx2);
    SNB_LAUNCH_CHECK("hashgrid_bwd_bwd_input");
    return SNB_OK;
}
