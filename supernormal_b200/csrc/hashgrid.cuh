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
    return c;
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
    return c.pow2 ? (idx & (c.size - 1)) : (idx % c.size);
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

}  // namespace snb
