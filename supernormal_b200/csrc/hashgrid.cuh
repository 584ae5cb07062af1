// hashgrid.cuh -- device helpers shared by the drop-in hash-grid kernels (hashgrid.cu) and the
// fused SDF kernels (sdf_fused.cu).  Semantics restate tiny-cuda-nn's grid encoding
// (encodings/grid.h, common_device.h @2ec562e) -- source absent from the
// reference tree, so parity is pinned only against oracle/c/oracle.c.
#pragma once
#include "common.cuh"

namespace snb {

struct LevelCtx {
    uint32_t offset, size, res;
    float scale;
    bool hashed;     // size < res^3 (index via coherent prime hash)
    bool pow2;       // size is a power of two
    uint32_t magic;  // floor(2^32 / size) for the multiply-high modulo of non-power-of-two levels; 0 = use %
};

__device__ __forceinline__ LevelCtx level_ctx(const snb_hashgrid_meta &m, uint32_t l) {
    LevelCtx c;
    c.offset = m.offsets[l];
    c.size = m.offsets[l + 1] - m.offsets[l];
    c.res = m.resolutions[l];
    c.scale = m.scales[l];
    // tcnn grid_index: for (dim<3 && stride<=size) stride*=res;  hashed iff size < stride afterwards
    uint32_t stride = 1;
    for (int d = 0; d < 3 && stride <= c.size; ++d) stride *= c.res;
    c.hashed = c.size < stride;
    c.pow2 = (c.size & (c.size - 1)) == 0;
    c.magic = 0;
    return c;
}

// Fused kernels receive the per-level constants as a by-value kernel parameter built on the host (constant bank:
// uniform operands, no per-point recomputation of the stride loop, and the division behind `magic` leaves the device).
struct LevelTable {
    LevelCtx lv[SNB_MAX_LEVELS];
};

static inline LevelTable make_level_table(const snb_hashgrid_meta &m) {
    LevelTable t;
    for (uint32_t l = 0; l < SNB_MAX_LEVELS; ++l) {
        LevelCtx c = {};
        if (l < m.n_levels) {
            c.offset = m.offsets[l];
            c.size = m.offsets[l + 1] - m.offsets[l];
            c.res = m.resolutions[l];
            c.scale = m.scales[l];
            uint64_t stride = 1;
            for (int d = 0; d < 3 && stride <= c.size; ++d) stride *= c.res;
            c.hashed = c.size < stride;
            c.pow2 = (c.size & (c.size - 1)) == 0;
            c.magic = (c.pow2 || c.size == 0) ? 0u : (uint32_t)(0x100000000ull / c.size);
        }
        t.lv[l] = c;
    }
    return t;
}

// idx mod size.  Dense levels are not powers of two and SuperNormal feeds x in [-1,1], so negative cells wrap through
// uint32 and the modulo is live on most points: q' = umulhi(idx, floor(2^32/size)) is floor(idx/size) or one less.
__device__ __forceinline__ uint32_t level_mod(const LevelCtx &c, uint32_t idx) {
    if (c.pow2) return idx & (c.size - 1);
    if (c.magic == 0) return idx % c.size;
    uint32_t r = idx - __umulhi(idx, c.magic) * c.size;
    return r >= c.size ? r - c.size : r;
}

struct Cell {
    uint32_t g[3];  // (uint32_t)(int)floor(pos)
    float w[3];     // pos - floor(pos)
};

__device__ __forceinline__ Cell cell_of(const LevelCtx &c, float x, float y, float z) {
    Cell r;
    float p[3] = {__fmaf_rn(c.scale, x, 0.5f), __fmaf_rn(c.scale, y, 0.5f), __fmaf_rn(c.scale, z, 0.5f)};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float fl = floorf(p[d]);
        r.g[d] = (uint32_t)(int)fl;
        r.w[d] = __fsub_rn(p[d], fl);
    }
    return r;
}

// entry index (without level offset) of corner `corner` (bit d set -> g[d]+1)
__device__ __forceinline__ uint32_t corner_index(const LevelCtx &c, const Cell &cell, uint32_t corner) {
    uint32_t px = cell.g[0] + (corner & 1u), py = cell.g[1] + ((corner >> 1) & 1u), pz = cell.g[2] + ((corner >> 2) & 1u);
    uint32_t idx;
    if (c.hashed) {
        idx = px ^ (py * 2654435761u) ^ (pz * 805459861u);
    } else {
        // dense: the reference's stride loop stops adding a dimension once stride > size
        idx = px;
        uint32_t stride = c.res;
        if (stride <= c.size) {
            idx += py * stride;
            stride *= c.res;
            if (stride <= c.size) idx += pz * stride;
        }
    }
    return level_mod(c, idx);
}

// trilinear weight in the reference's multiplication order ((1*a0)*a1)*a2
__device__ __forceinline__ float corner_weight(const Cell &cell, uint32_t corner) {
    float w = (corner & 1u) ? cell.w[0] : __fsub_rn(1.f, cell.w[0]);
    w = __fmul_rn(w, (corner & 2u) ? cell.w[1] : __fsub_rn(1.f, cell.w[1]));
    w = __fmul_rn(w, (corner & 4u) ? cell.w[2] : __fsub_rn(1.f, cell.w[2]));
    return w;
}

// fp16-faithful interpolation of one level: result = fma((half)w, table[idx], result) over corners 0..7
__device__ __forceinline__ __half2 interp_level(const LevelCtx &c, const Cell &cell, const __half2 *__restrict__ table) {
    const __half2 *t = table + c.offset;
    __half2 v[8];
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) v[k] = __ldg(t + corner_index(c, cell, k));
    __half2 acc = __float2half2_rn(0.f);
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) acc = __hfma2(__float2half2_rn(corner_weight(cell, k)), v[k], acc);
    return acc;
}

// value (fp16-faithful, as interp_level) and d(value)/dx of one level sharing the 8 corner loads.  The derivative is
// tiny-cuda-nn's dy_dx for linear interpolation (encodings/grid.h, kernel_grid with dy_dx != nullptr): fp32,
//   d/dx_d = scale * sum over the 4 corner pairs along d of  a_e0 * a_e1 * (v_right - v_left).
__device__ __forceinline__ __half2 interp_level_grad(const LevelCtx &c, const Cell &cell, const __half2 *__restrict__ table,
                                                     float2 (&dv)[3]) {
    const __half2 *t = table + c.offset;
    __half2 v[8];
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) v[k] = __ldg(t + corner_index(c, cell, k));
    __half2 acc = __float2half2_rn(0.f);
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) acc = __hfma2(__float2half2_rn(corner_weight(cell, k)), v[k], acc);
    float2 vf[8];
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) vf[k] = __half22float2(v[k]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int e0 = (d + 1) % 3, e1 = (d + 2) % 3;
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            if ((k >> d) & 1u) continue;
            const float a0 = ((k >> e0) & 1u) ? cell.w[e0] : 1.f - cell.w[e0];
            const float a1 = ((k >> e1) & 1u) ? cell.w[e1] : 1.f - cell.w[e1];
            const float w = a0 * a1;
            sx = fmaf(w, vf[k | (1u << d)].x - vf[k].x, sx);
            sy = fmaf(w, vf[k | (1u << d)].y - vf[k].y, sy);
        }
        dv[d] = make_float2(c.scale * sx, c.scale * sy);
    }
    return acc;
}

}  // namespace snb
