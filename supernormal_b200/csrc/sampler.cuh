// sampler.cuh -- counter-based RNG and the per-ray body of the device patch sampler (models/dataset_loader.py:223-297),
// shared by sampler.cu (stand-alone sampler, occupancy update) and optim.cu (step-tail kernel that pre-samples the next batch).
#pragma once
#include "sdf_core.cuh"

namespace snb {

// Philox4x32-10 (Salmon et al. 2011), counter-based: ctr = (step lo, step hi, index, stream), key = seed
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }  // [0,1), 24 bits like torch.rand

// one thread per patch ray (tid in [0, n_patches * 9)); shared by sample_patches_kernel and the step-tail kernel (optim.cu)
__device__ __forceinline__ void sample_patch_ray(const snb_dataset &ds, int n_patches, uint64_t seed, uint64_t step, const snb_batch_out &out,
                                                 int tid) {
    if (tid >= n_patches * SNB_PATCH) return;
    int i = tid / SNB_PATCH, k = tid % SNB_PATCH;
    uint4 r = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)i, 0x5a4dce11u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    int cx = 1 + (int)(r.x % (uint32_t)(ds.W - 3));   // randint(low=1, high=W-2)
    int cy = 1 + (int)(r.y % (uint32_t)(ds.H - 3));
    int view = __ldg(ds.train_ids + (r.z % (uint32_t)ds.n_train));
    int px = cx + (k % 3) - 1, py = cy + (k / 3) - 1;
    const float *Ki = ds.intrinsics_inv + view * 16, *Pm = ds.pose + view * 16;
    float fx = (float)px, fy = (float)py;
    float p[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __ldg(Ki + 4 * a) * fx + __ldg(Ki + 4 * a + 1) * fy + __ldg(Ki + 4 * a + 2);
    float inv = 1.f / sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    p[0] *= inv; p[1] *= inv; p[2] *= inv;
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = __ldg(Pm + 4 * a) * p[0] + __ldg(Pm + 4 * a + 1) * p[1] + __ldg(Pm + 4 * a + 2) * p[2];
    float o[3] = {__ldg(Pm + 3), __ldg(Pm + 7), __ldg(Pm + 11)};
    int64_t rk = (int64_t)i * SNB_PATCH + k;
    out.rays_d[3 * rk] = d[0]; out.rays_d[3 * rk + 1] = d[1]; out.rays_d[3 * rk + 2] = d[2];
    int64_t pix = ((int64_t)view * ds.H + py) * ds.W + px;
#pragma unroll
    for (int a = 0; a < 3; ++a) out.normal_gt[3 * rk + a] = __ldg(ds.normals + 3 * pix + a);
    out.mask[rk] = __ldg(ds.masks + pix);
    if (ds.v_inverse) {   // precomputed table (Dataset.V_inverse_all, models/dataset_loader.py:114-137): 36 B per pixel, 226 MB for DiLiGenT-MV
#pragma unroll
        for (int a = 0; a < 9; ++a) out.v_inv[9 * rk + a] = __ldg(ds.v_inverse + 9 * pix + a);
    } else {
        // closed form of the same inverse.  V = [v; r; d] (rows) with v = R p (the unit ray direction), r = R e_x, d = R e_y, i.e.
        // V = A R^T with A = [p; e_x; e_y] in the camera frame, so V^-1 = R A^-1 (R orthonormal: load_K_Rt_from_P transposes the rotation
        // of cv2.decomposeProjectionMatrix) and A^-1 = [[0, 1, 0], [0, 0, 1], [1/pz, -px/pz, -py/pz]]:
        //   V^-1[:, 0] = R[:, 2] / pz,   V^-1[:, 1] = R[:, 0] - R[:, 2] px / pz,   V^-1[:, 2] = R[:, 1] - R[:, 2] py / pz
        const float ipz = 1.f / p[2], qx = p[0] * ipz, qy = p[1] * ipz;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float r0 = __ldg(Pm + 4 * a), r1 = __ldg(Pm + 4 * a + 1), r2 = __ldg(Pm + 4 * a + 2);
            out.v_inv[9 * rk + 3 * a] = r2 * ipz;
            out.v_inv[9 * rk + 3 * a + 1] = fmaf(-r2, qx, r0);
            out.v_inv[9 * rk + 3 * a + 2] = fmaf(-r2, qy, r1);
        }
    }
    if (k == SNB_PATCH / 2) {
        out.rays_o[3 * i] = o[0]; out.rays_o[3 * i + 1] = o[1]; out.rays_o[3 * i + 2] = o[2];
        out.plane_n[3 * i] = __ldg(Pm + 2); out.plane_n[3 * i + 1] = __ldg(Pm + 6); out.plane_n[3 * i + 2] = __ldg(Pm + 10);
        // near_far_from_sphere, models/dataset_loader.py:279-297
        float a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        float b = 2.0f * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
        float c = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - 1.0f;
        float mid = 0.5f * (-b) / a;
        float root = sqrtf(b * b - 4.f * a * c) / (2.f * a);  // NaN if the ray misses the unit sphere
        out.near_[i] = mid - root;
        out.far_[i] = mid + root;
        if (out.jitter) out.jitter[i] = u01(r.w);
    }
}

}  // namespace snb
