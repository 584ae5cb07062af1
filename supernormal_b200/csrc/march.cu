// march.cu -- occupancy-grid ray marching, warp-per-ray, bit-exact with the reference kernel
// (CS/ray_marching.cu:81-192 as compiled by nvcc for sm_100a; rounding recipe in SURVEY.md
// Appendix C).  Every floating-point operation is an explicit round-to-nearest intrinsic so the
// result does not depend on this translation unit's contraction decisions.
//
// Parallelisation: the reference runs one thread per ray and walks it serially.  Here a warp owns
// the ray.  Lane l replays the (cheap, serial) t0/t1 recurrence l steps ahead, all 32 lanes fetch
// their occupancy byte concurrently (one L2 round trip per 32 candidate samples instead of 32),
// a ballot finds the first candidate that is not an occupied in-range sample, and the occupied
// prefix is stored with one coalesced write.  Only the empty-space skip (whose repeated
// `t += dt` must be replayed add-by-add to stay bit-exact) is serial.
#include "march_common.cuh"

namespace snb {

struct MarchArgs {
    int32_t n_rays;
    const float *rays_o, *rays_d, *t_min, *t_max, *roi;
    int3 res;
    const uint8_t *grid;
    float step, cone;
};

template <bool EMIT>
__global__ void __launch_bounds__(128) march_kernel(MarchArgs a, const int32_t *__restrict__ packed_info,
                                                    int64_t capacity, int32_t *__restrict__ num_steps,
                                                    int64_t *__restrict__ ridx64, int32_t *__restrict__ ridx32,
                                                    float *__restrict__ t_starts, float *__restrict__ t_ends) {
    const int lane = threadIdx.x & 31;
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.n_rays) return;
    const RoiCtx rc = make_roi_ctx(a.roi, a.res);
    const float o[3] = {__ldg(a.rays_o + 3 * ray), __ldg(a.rays_o + 3 * ray + 1), __ldg(a.rays_o + 3 * ray + 2)};
    const float d[3] = {__ldg(a.rays_d + 3 * ray), __ldg(a.rays_d + 3 * ray + 1), __ldg(a.rays_d + 3 * ray + 2)};
    const float inv_d[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
    const float near = __ldg(a.t_min + ray), far = __ldg(a.t_max + ray);
    const float dt_min = a.step;

    int64_t base = 0;
    if (EMIT) base = packed_info[2 * ray];

    int j = 0;
    float t0 = near;
    float t1 = __fadd_rn(t0, march_dt(t0, a.cone, dt_min));
    float t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);

    while (t_mid < far) {  // warp-uniform state
        // Empty space: probe only the current candidate (warp-uniform, broadcast load) and DDA-skip; the 32-wide
        // speculation below is only worth its cost once a sample has been found.
        {
            float cx = __fmaf_rn(t_mid, d[0], o[0]), cy = __fmaf_rn(t_mid, d[1], o[1]), cz = __fmaf_rn(t_mid, d[2], o[2]);
            if (!march_occupied(rc, cx, cy, cz, a.grid)) {
                t_mid = march_skip(rc, t_mid, dt_min, cx, cy, cz, d, inv_d, far);
                float dt = march_dt(t_mid, a.cone, dt_min);
                t0 = __fmaf_rn(dt, -0.5f, t_mid);
                t1 = __fmaf_rn(dt, 0.5f, t_mid);
                continue;
            }
        }
        // lane l: state after l consecutive occupied samples
        float l0 = t0, l1 = t1;
        for (int s = 0; s < lane; ++s) {
            l0 = l1;
            l1 = __fadd_rn(l0, march_dt(l0, a.cone, dt_min));
        }
        float lm = (lane == 0) ? t_mid : __fmul_rn(__fadd_rn(l0, l1), 0.5f);
        float px = __fmaf_rn(lm, d[0], o[0]), py = __fmaf_rn(lm, d[1], o[1]), pz = __fmaf_rn(lm, d[2], o[2]);
        bool in_range = lm < far;
        bool occ = in_range && march_occupied(rc, px, py, pz, a.grid);
        unsigned stop = ~__ballot_sync(0xffffffffu, occ);
        int f = stop ? (__ffs(stop) - 1) : 32;  // lanes [0,f) are emitted samples
        if (EMIT && lane < f) {
            int64_t off = base + j + lane;
            if (off < capacity) {
                t_starts[off] = l0;
                t_ends[off] = l1;
                if (ridx64) ridx64[off] = ray;
                if (ridx32) ridx32[off] = ray;
            }
        }
        j += f;
        if (f == 32) {  // continue after lane 31's sample
            t0 = __shfl_sync(0xffffffffu, l1, 31);
            t1 = __fadd_rn(t0, march_dt(t0, a.cone, dt_min));
            t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
            continue;
        }
        // state of the first non-sample candidate
        float s_mid = __shfl_sync(0xffffffffu, lm, f);
        bool s_in = __shfl_sync(0xffffffffu, (int)in_range, f) != 0;
        if (!s_in) break;  // t_mid >= far (or NaN): the reference loop exits here
        float sx = __shfl_sync(0xffffffffu, px, f), sy = __shfl_sync(0xffffffffu, py, f), sz = __shfl_sync(0xffffffffu, pz, f);
        t_mid = march_skip(rc, s_mid, dt_min, sx, sy, sz, d, inv_d, far);
        float dt = march_dt(t_mid, a.cone, dt_min);
        t0 = __fmaf_rn(dt, -0.5f, t_mid);
        t1 = __fmaf_rn(dt, 0.5f, t_mid);
    }
    if (!EMIT && lane == 0) num_steps[ray] = j;
}

// Exclusive scan of per-ray counts in one CTA (n is a few thousand on the training path; the loop
// handles any n).  packed_info[i] = (offset_i, count_i); *total = sum.
__global__ void __launch_bounds__(1024) packed_info_kernel(int32_t n, const int32_t *__restrict__ counts,
                                                           int32_t *__restrict__ packed, int32_t *__restrict__ total) {
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int32_t start = 0; start < n; start += 1024) {
        int32_t i = start + threadIdx.x;
        int32_t c = (i < n) ? counts[i] : 0;
        int32_t v = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        int32_t carry = carry_s;
        int32_t incl = v + (warp ? warp_sums[warp - 1] : 0) + carry;
        if (i < n) {
            packed[2 * i] = incl - c;
            packed[2 * i + 1] = c;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

static int32_t check_march(int32_t n_rays, const void *o, const void *d, const void *tmin, const void *tmax,
                           const void *roi, int rx, int ry, int rz, const void *grid, float step) {
    SNB_REQUIRE(n_rays >= 0, SNB_ERR_ARG, "ray_marching: n_rays < 0");
    SNB_REQUIRE(n_rays == 0 || (o && d && tmin && tmax), SNB_ERR_NULL, "ray_marching: null ray buffer");
    SNB_REQUIRE(roi && grid, SNB_ERR_NULL, "ray_marching: null roi/grid");
    SNB_REQUIRE(step > 0.0f, SNB_ERR_ARG, "ray_marching: step_size must be > 0");
    SNB_REQUIRE(rx > 0 && ry > 0 && rz > 0 && (int64_t)rx * ry * rz < (1ll << 31), SNB_ERR_ARG,
                "ray_marching: bad grid resolution %d %d %d", rx, ry, rz);
    return SNB_OK;
}

}  // namespace snb

using namespace snb;

extern "C" int32_t snb_march_count(int32_t n_rays, const float *rays_o, const float *rays_d, const float *t_min,
                                   const float *t_max, const float *roi, int32_t res_x, int32_t res_y, int32_t res_z,
                                   const uint8_t *grid_binary, float step_size, float cone_angle, int32_t *num_steps,
                                   snb_stream_t stream) {
    int32_t rc = check_march(n_rays, rays_o, rays_d, t_min, t_max, roi, res_x, res_y, res_z, grid_binary, step_size);
    if (rc) return rc;
    if (n_rays == 0) return SNB_OK;
    SNB_REQUIRE(num_steps, SNB_ERR_NULL, "march_count: null num_steps");
    MarchArgs a{n_rays, rays_o, rays_d, t_min, t_max, roi, make_int3(res_x, res_y, res_z), grid_binary, step_size, cone_angle};
    march_kernel<false><<<(unsigned)cdiv(n_rays, 4), 128, 0, S(stream)>>>(a, nullptr, 0, num_steps, nullptr, nullptr, nullptr, nullptr);
    SNB_LAUNCH_CHECK("march_count");
    return SNB_OK;
}

extern "C" int32_t snb_packed_info_from_counts(int32_t n_rays, const int32_t *num_steps, int32_t *packed_info,
                                               int32_t *total, snb_stream_t stream) {
    SNB_REQUIRE(n_rays >= 0, SNB_ERR_ARG, "packed_info: n_rays < 0");
    SNB_REQUIRE(n_rays == 0 || (num_steps && packed_info), SNB_ERR_NULL, "packed_info: null buffer");
    packed_info_kernel<<<1, 1024, 0, S(stream)>>>(n_rays, num_steps, packed_info, total);
    SNB_LAUNCH_CHECK("packed_info");
    return SNB_OK;
}

extern "C" int32_t snb_march_emit(int32_t n_rays, const float *rays_o, const float *rays_d, const float *t_min,
                                  const float *t_max, const float *roi, int32_t res_x, int32_t res_y, int32_t res_z,
                                  const uint8_t *grid_binary, float step_size, float cone_angle,
                                  const int32_t *packed_info, int64_t capacity, int64_t *ray_indices_i64,
                                  int32_t *ray_indices_i32, float *t_starts, float *t_ends, snb_stream_t stream) {
    int32_t rc = check_march(n_rays, rays_o, rays_d, t_min, t_max, roi, res_x, res_y, res_z, grid_binary, step_size);
    if (rc) return rc;
    if (n_rays == 0 || capacity == 0) return SNB_OK;
    SNB_REQUIRE(packed_info && t_starts && t_ends, SNB_ERR_NULL, "march_emit: null output");
    MarchArgs a{n_rays, rays_o, rays_d, t_min, t_max, roi, make_int3(res_x, res_y, res_z), grid_binary, step_size, cone_angle};
    march_kernel<true><<<(unsigned)cdiv(n_rays, 4), 128, 0, S(stream)>>>(a, packed_info, capacity, nullptr, ray_indices_i64,
                                                                        ray_indices_i32, t_starts, t_ends);
    SNB_LAUNCH_CHECK("march_emit");
    return SNB_OK;
}
