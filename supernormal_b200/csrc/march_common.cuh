// march_common.cuh -- bit-exact building blocks of the reference ray marcher (CS/ray_marching.cu:9-75 as nvcc
// compiles it for sm_100a), shared by march.cu and fused_march.cu.
//
// Every division of the reference is an IEEE division by (roi_max - roi_min) or by float(resolution).  When the
// divisor is a power of two -- the shipped configuration: roi [-1,1]^3 -> 2.0, resolution 128 -- x / d and
// x * (1/d) are the same real number, both correctly rounded, hence bit-identical; the multiply replaces the
// ~10-instruction division sequence on the marcher's serial critical path.
#pragma once
#include "common.cuh"

namespace snb {

struct RoiCtx {
    float rmin[3], rmax[3], ext[3], inv_ext[3], resf[3], inv_resf[3];
    int res[3];
    bool ext_pow2[3], res_pow2[3];
};

__device__ __forceinline__ bool is_pow2f(float v) {
    unsigned u = __float_as_uint(v);
    unsigned e = (u >> 23) & 0xffu;
    return v > 0.f && (u & 0x007fffffu) == 0u && e > 1u && e < 253u;  // normal, and 1/v is normal too
}

__device__ __forceinline__ RoiCtx make_roi_ctx(const float *__restrict__ roi, int3 res) {
    RoiCtx c;
    c.res[0] = res.x; c.res[1] = res.y; c.res[2] = res.z;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c.rmin[a] = __ldg(roi + a);
        c.rmax[a] = __ldg(roi + 3 + a);
        c.ext[a] = __fsub_rn(c.rmax[a], c.rmin[a]);
        c.ext_pow2[a] = is_pow2f(c.ext[a]);
        c.inv_ext[a] = c.ext_pow2[a] ? __fdiv_rn(1.0f, c.ext[a]) : 0.f;
        c.resf[a] = (float)c.res[a];
        c.res_pow2[a] = is_pow2f(c.resf[a]);
        c.inv_resf[a] = c.res_pow2[a] ? __fdiv_rn(1.0f, c.resf[a]) : 0.f;
    }
    return c;
}

__device__ __forceinline__ float div_ext(const RoiCtx &c, int a, float x) { return c.ext_pow2[a] ? __fmul_rn(x, c.inv_ext[a]) : __fdiv_rn(x, c.ext[a]); }
__device__ __forceinline__ float div_res(const RoiCtx &c, int a, float x) { return c.res_pow2[a] ? __fmul_rn(x, c.inv_resf[a]) : __fdiv_rn(x, c.resf[a]); }

// clamp(t*cone, dt_min, 1e10) == fmaxf(dt_min, fminf(t*cone, 1e10))  (helpers_math.h:1167)
__device__ __forceinline__ float march_dt(float t, float cone, float dt_min) { return fmaxf(dt_min, fminf(__fmul_rn(t, cone), 1e10f)); }

// grid_occupied_at, CS/ray_marching.cu:27-45 (AABB)
__device__ __forceinline__ bool march_occupied(const RoiCtx &c, float x, float y, float z, const uint8_t *__restrict__ grid) {
    if (x < c.rmin[0] || x > c.rmax[0] || y < c.rmin[1] || y > c.rmax[1] || z < c.rmin[2] || z > c.rmax[2]) return false;
    float p[3] = {x, y, z};
    int ix[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float u = div_ext(c, a, __fsub_rn(p[a], c.rmin[a]));
        ix[a] = min(max(__float2int_rz(__fmul_rn(u, c.resf[a])), 0), c.res[a] - 1);
    }
    return __ldg(grid + ((ix[0] * c.res[1] + ix[1]) * c.res[2] + ix[2])) != 0;
}

// distance_to_next_voxel, CS/ray_marching.cu:48-57:
// ((floorf(_x + 0.5 + 0.5*sign(dir)) - _x) * inv_dir) / res * (roi_max-roi_min), _x = u*res never rounded on its own
__device__ __forceinline__ float march_axis_dist(const RoiCtx &c, int a, float p, float dir, float inv_dir) {
    float u = div_ext(c, a, __fsub_rn(p, c.rmin[a]));
    float fl = floorf(__fmaf_rn(copysignf(1.0f, dir), 0.5f, __fmaf_rn(c.resf[a], u, 0.5f)));
    float diff = __fmaf_rn(c.resf[a], -u, fl);
    return __fmul_rn(div_res(c, a, __fmul_rn(diff, inv_dir)), c.ext[a]);
}

// advance_to_next_voxel, CS/ray_marching.cu:59-75 (the repeated add must be replayed serially)
__device__ __forceinline__ float march_skip(const RoiCtx &c, float t_mid, float dt_min, float px, float py, float pz, const float *d,
                                            const float *inv_d, float far) {
    float tx = march_axis_dist(c, 0, px, d[0], inv_d[0]);
    float ty = march_axis_dist(c, 1, py, d[1], inv_d[1]);
    float tz = march_axis_dist(c, 2, pz, d[2], inv_d[2]);
    float t_target = fminf(__fadd_rn(t_mid, fmaxf(fminf(fminf(tx, ty), tz), 0.0f)), far);
    float t = t_mid;
    do {
        t = __fadd_rn(t, dt_min);
    } while (t < t_target);
    return t;
}

// same, also counting the lattice steps taken (used by the speculative empty-space windows of fused_march.cu)
__device__ __forceinline__ float march_skip_count(const RoiCtx &c, float t_mid, float dt_min, float px, float py, float pz, const float *d,
                                                  const float *inv_d, float far, int &steps) {
    float tx = march_axis_dist(c, 0, px, d[0], inv_d[0]);
    float ty = march_axis_dist(c, 1, py, d[1], inv_d[1]);
    float tz = march_axis_dist(c, 2, pz, d[2], inv_d[2]);
    float t_target = fminf(__fadd_rn(t_mid, fmaxf(fminf(fminf(tx, ty), tz), 0.0f)), far);
    float t = t_mid;
    int n = 0;
    do {
        t = __fadd_rn(t, dt_min);
        ++n;
    } while (t < t_target);
    steps = n;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// Closed form of the reference's repeated rounded adds  t (+) dt (+) dt ...  (advance_to_next_voxel's do-while and
// the t0/t1 stepping of the march loop).  While the run stays inside the binade of its base value, every t is an
// integer multiple T of u = ulp(base), and one rounded add maps T to rne(T + dt/u) = T + rne(dt/u): the same integer
// increment Q every time -- unless dt/u has a fractional part of exactly 1/2, where round-half-even would depend on
// the parity of T.  Then: value after k adds = (T0 + k Q) u, and the number of adds the do-while performs to reach
// t_target is max(1, ceil((T_target - T) / Q)) -- integer arithmetic, bit-identical to replaying the adds.
// Q == 0 marks "closed form not available" (base not a positive normal, tie case, dt outside [u, 2^22 u]); callers then
// replay the adds serially, as they do when a run leaves the binade.
struct AddChain {
    uint32_t T0;   // 24-bit significand of the base (hidden bit set)
    uint32_t hi;   // exponent bits of the base, in place
    uint32_t Q;    // significand increment per add; 0: unavailable
};

__device__ __forceinline__ AddChain make_add_chain(float base, float dt) {
    AddChain c;
    const uint32_t b = __float_as_uint(base), e = b >> 23;   // callers pass base > 0: sign bit clear
    c.T0 = (b & 0x007fffffu) | 0x00800000u;
    c.hi = e << 23;
    c.Q = 0;
    const int sh = 150 - (int)e;                              // dt / u = dt * 2^sh
    if (!(base > 0.f) || e == 0u || e >= 255u || sh > 127 || sh < -126) return c;
    const float D = __fmul_rn(dt, __uint_as_float((uint32_t)(sh + 127) << 23));   // exact: power-of-two scaling
    if (!(D >= 1.f && D <= 4194304.f)) return c;
    const float Di = floorf(D), f = __fsub_rn(D, Di);         // exact: D < 2^23
    if (f == 0.5f) return c;
    c.Q = (uint32_t)Di + (f > 0.5f ? 1u : 0u);
    return c;
}

// value after k adds; false when the closed form does not apply (caller falls back to the serial replay)
__device__ __forceinline__ bool add_chain_at(const AddChain &c, uint32_t k, float &out) {
    const uint64_t T = (uint64_t)c.T0 + (uint64_t)k * c.Q;
    if (c.Q == 0u || T >= 0x01000000ull) return false;
    out = __uint_as_float(c.hi | ((uint32_t)T & 0x007fffffu));
    return true;
}

// do { t += dt; ++n; } while (t < t_target) starting from the chain point with significand T; false -> fall back
__device__ __forceinline__ bool add_chain_skip(const AddChain &c, uint32_t T, float t_target, float &t_out, int &n_out) {
    const uint32_t tb = __float_as_uint(t_target);
    if (c.Q == 0u || (tb & 0xff800000u) != c.hi) return false;   // other binade, negative, inf or NaN target
    const uint32_t TT = (tb & 0x007fffffu) | 0x00800000u;
    uint32_t n = TT > T ? (TT - T + c.Q - 1u) / c.Q : 1u;
    if (n == 0u) n = 1u;
    const uint64_t Tl = (uint64_t)T + (uint64_t)n * c.Q;
    if (Tl >= 0x01000000ull) return false;
    t_out = __uint_as_float(c.hi | ((uint32_t)Tl & 0x007fffffu));
    n_out = (int)n;
    return true;
}

// march_skip_count with the closed form when it applies (T = significand of t_mid on chain c)
__device__ __forceinline__ float march_skip_count_fast(const RoiCtx &c, const AddChain &ch, uint32_t T, float t_mid, float dt_min, float px,
                                                       float py, float pz, const float *d, const float *inv_d, float far, int &steps) {
    float tx = march_axis_dist(c, 0, px, d[0], inv_d[0]);
    float ty = march_axis_dist(c, 1, py, d[1], inv_d[1]);
    float tz = march_axis_dist(c, 2, pz, d[2], inv_d[2]);
    float t_target = fminf(__fadd_rn(t_mid, fmaxf(fminf(fminf(tx, ty), tz), 0.0f)), far);
    float t;
    if (add_chain_skip(ch, T, t_target, t, steps)) return t;
    t = t_mid;
    int n = 0;
    do {
        t = __fadd_rn(t, dt_min);
        ++n;
    } while (t < t_target);
    steps = n;
    return t;
}

}  // namespace snb
