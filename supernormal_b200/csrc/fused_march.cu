// fused_march.cu -- occupancy-grid marching of the patch-centre rays fused with the NeuS visibility cut.
//
// Reference sequence (NA/ray_marching.py:157-220 driven by models/renderer.py:80-135): march every
// centre ray to `far` (two kernel passes + cumsum + .item()), query the SDF at ALL S0 samples, CUB
// exclusive product of (1-alpha), keep samples with T >= early_stop_eps, boolean-mask compaction
// (nonzero -> another host sync).  Because T is non-increasing along a ray the kept samples are a
// PREFIX of each ray's samples, so here one warp marches a ray in batches of 31 candidate samples
// (same bit-exact stepping as march.cu), evaluates the SDF of that batch in-lane (encode + MLP, no
// HBM round trip), multiplies T serially, and stops marching the moment T < eps: late in training
// this evaluates ~S instead of ~S0 >> S points.  Samples go to a per-ray scratch; a device scan and
// a copy kernel pack them by patch with no host round trip.
#include "march_common.cuh"
#include "sdf_core.cuh"

namespace snb {

// Empty space.  In the reference an unoccupied probe at t_mid jumps to the first point of the lattice
//     t_mid (+) dt (+) dt (+) ...            ((+) = one rounded fp32 add, advance_to_next_voxel's do-while)
// at or beyond the voxel exit, so EVERY point the serial marcher visits in an empty stretch lies on that one
// lattice.  The CTA therefore evaluates 32 W consecutive lattice points at once -- thread k takes the point after
// k adds (closed form, march_common.cuh), probes the grid and runs the reference's own skip arithmetic as if it
// were visited -- and then follows the visited chain 0 -> 0+J(0) -> ... through shared memory: the per-voxel
// dependent load + division chain of the serial marcher becomes one parallel step plus log2(32 W) pointer-doubling rounds, and
// the visited points, hence the emitted samples, are bit-identical.
//
// Occupied stretches.  W warps of a CTA own ONE ray.  A lone warp needs ~40 k cycles for the SDF of a 31-sample batch at
// 14 levels (in-order issue: it cannot overlap its own gathers with its own MLP), and the longest rays chain 10-20 such
// batches while the SM idles (measured per-ray with clock64; profiles/README.md).  So the batch's 32 points are
// evaluated by the whole CTA: warp w gathers and blends levels w, w+W, ... into a shared feature tile, then computes
// hidden units [64 w / W, 64 (w+1) / W) of the MLP for all 32 points; the partial sums meet in shared memory.  No work
// is duplicated or speculated; the batch latency drops ~W-fold.  Control state (t0, t1, t_mid, T, j, ...) is kept
// redundantly in every thread, so the control flow is CTA-uniform and the barriers are unconditional.
constexpr int kFeatStride = 2 * SNB_MAX_LEVELS + 1;   // odd: conflict-free row-per-lane reads

// SDF of the warp-lane's point, computed cooperatively by the W warps of the CTA (all threads must call; every warp
// passes the same 32 points).  s_feat: [32][kFeatStride], s_part: [W][32].
template <int W>
__device__ __forceinline__ float cta_sdf(bool valid, float x, float y, float z, const __half2 *__restrict__ table, const LevelCtx *lvl,
                                         uint32_t n_active, const float *s_net, float *s_feat, float *s_part, int warp, int lane) {
    constexpr int HU = kH / W;   // hidden units per warp
    for (uint32_t l = warp; l < n_active; l += W) {
        float2 ff = make_float2(0.f, 0.f);
        if (valid) {
            const LevelCtx c = lvl[l];
            Cell cell = cell_of(c, x, y, z);
            ff = __half22float2(interp_level(c, cell, table));
        }
        s_feat[lane * kFeatStride + 2 * l] = ff.x;
        s_feat[lane * kFeatStride + 2 * l + 1] = ff.y;
    }
    __syncthreads();
    const int h0 = HU * warp;
    float acc[HU];
#pragma unroll
    for (int q = 0; q < HU / 4; ++q) {
        const float4 bq = *reinterpret_cast<const float4 *>(s_net + kOffB0 + h0 + 4 * q);
        acc[4 * q] = bq.x; acc[4 * q + 1] = bq.y; acc[4 * q + 2] = bq.z; acc[4 * q + 3] = bq.w;
    }
    const float xin[3] = {x, y, z};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int q = 0; q < HU / 4; ++q) {
            const float4 wq = *reinterpret_cast<const float4 *>(s_net + kOffW0T + r * kH + h0 + 4 * q);
            acc[4 * q] = fmaf(wq.x, xin[r], acc[4 * q]); acc[4 * q + 1] = fmaf(wq.y, xin[r], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(wq.z, xin[r], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(wq.w, xin[r], acc[4 * q + 3]);
        }
    }
    const float *frow = s_feat + lane * kFeatStride;
    for (uint32_t jf = 0; jf < 2 * n_active; ++jf) {
        const float v = frow[jf];
        const float *wr = s_net + kOffW0T + (3 + jf) * kH + h0;
#pragma unroll
        for (int q = 0; q < HU / 4; ++q) {
            const float4 wq = *reinterpret_cast<const float4 *>(wr + 4 * q);
            acc[4 * q] = fmaf(wq.x, v, acc[4 * q]); acc[4 * q + 1] = fmaf(wq.y, v, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(wq.z, v, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(wq.w, v, acc[4 * q + 3]);
        }
    }
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < HU; ++i) part = fmaf(s_net[kOffW1 + h0 + i], softplus100(acc[i]), part);
    s_part[warp * 32 + lane] = part;
    __syncthreads();
    float s = s_net[kOffB1];
#pragma unroll
    for (int w = 0; w < W; ++w) s += s_part[w * 32 + lane];
    return s;
}

#ifdef SNB_MARCH_DEBUG
__device__ long long g_march_dbg[8192 * 4];   // per ray: cycles, windows, batches, samples (scripts/march_ray_times.py)
#endif

template <int W, int MB>
__global__ void __launch_bounds__(32 * W, MB) march_visible_kernel(snb_patch_batch b, snb_net net, LevelTable lt, const float *__restrict__ roi,
                                                                       int3 res, const uint8_t *__restrict__ grid, float step,
                                                                       const float *__restrict__ jitter, float eps, snb_samples sm) {
    __shared__ __align__(16) float s_net[kNetFloats];
    __shared__ float s_land[32 * W], s_tm[32 * W];
    __shared__ int s_J[32 * W];
    __shared__ uint32_t s_hop[2][32 * W + 1];   // pointer doubling over the window's hop chain: end node | (last hop source + 1) << 16
    __shared__ float s_feat[32 * kFeatStride], s_part[32 * W];
#ifdef SNB_MARCH_DEBUG
    const long long dbg_t0 = clock64();
    int dbg_windows = 0, dbg_batches = 0;
#endif
    const LevelCtx *s_lvl = lt.lv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ray = blockIdx.x;
    if (ray >= b.n_patches) return;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    // the MLP weights are staged on the ray's first occupied batch: half to three quarters of a batch's rays never meet an occupied
    // cell (background patches; per-ray clock64 profile in profiles/r02_march_ray_times.txt) and leave without them
    bool net_ready = false;
    float inv_s = 0.f;
    if (threadIdx.x == 0) s_hop[0][32 * W] = s_hop[1][32 * W] = 32u * W;   // the node behind the window absorbs
    constexpr unsigned kAll = 0xffffffffu;

    const RoiCtx rc = make_roi_ctx(roi, res);
    const float o[3] = {__ldg(b.rays_o + 3 * ray), __ldg(b.rays_o + 3 * ray + 1), __ldg(b.rays_o + 3 * ray + 2)};
    const float *dc = b.rays_d + ((int64_t)ray * SNB_PATCH + SNB_PATCH / 2) * 3;  // centre ray of the patch
    const float d[3] = {__ldg(dc), __ldg(dc + 1), __ldg(dc + 2)};
    const float inv_d[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
    float near = __ldg(b.near_ + ray);
    const float far = __ldg(b.far_ + ray);
    if (jitter) near = __fadd_rn(near, __fmul_rn(__ldg(jitter + ray), step));  // NA/ray_marching.py:158

    float *sc0 = sm.scratch_t0 + (int64_t)ray * sm.scratch_stride;
    float *sc1 = sm.scratch_t1 + (int64_t)ray * sm.scratch_stride;

    int j = 0, runs = 0;
    bool chain_open = false, overflow = false, occ_mode = false;
    float T = 1.f;
    // cone_angle == 0: calc_dt (CS/ray_marching.cu:9-14) is `step` for every finite t
    float t0 = near;
    float t1 = __fadd_rn(t0, step);
    float t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);

    while (t_mid < far) {
        if (!occ_mode) {
            // ---- speculative window over 32 W points of the empty-space lattice
#ifdef SNB_MARCH_DEBUG
            ++dbg_windows;
#endif
            const int idx = 32 * warp + lane;
            AddChain ch = make_add_chain(t_mid, step);
            float tm;
            if (add_chain_at(ch, 32u * W - 1u, tm)) {   // the closed form covers the last point of the window, hence all of it
                add_chain_at(ch, (uint32_t)idx, tm);
            } else {                                    // serial replay
                ch.Q = 0u;
                tm = t_mid;
                for (int s = 0; s < idx; ++s) tm = __fadd_rn(tm, step);
            }
            const bool inr = tm < far;
            const float px = __fmaf_rn(tm, d[0], o[0]), py = __fmaf_rn(tm, d[1], o[1]), pz = __fmaf_rn(tm, d[2], o[2]);
            const bool occ = inr && march_occupied(rc, px, py, pz, grid);
            int J = 1;
            float land = tm;
            if (inr && !occ) land = march_skip_count_fast(rc, ch, ch.T0 + (uint32_t)idx * ch.Q, tm, step, px, py, pz, d, inv_d, far, J);
            s_land[idx] = land;
            s_tm[idx] = tm;
            const int hop = !inr ? -1 : (occ ? 0 : J);   // hop length; 0: occupied, -1: beyond far
            s_J[idx] = hop;
            // Follow the visited chain 0 -> J(0) -> ... to its first occupied / beyond-far node or out of the window.  With a marching
            // step near the cell size every hop is 1-3 points long, and walking up to 32 W dependent shared-memory reads serially was
            // ~60 % of a window (5.7 k cycles per window at iteration 1000).  Pointer doubling instead: node i holds the node reached after
            // 2^r hops (terminal nodes and the node behind the window absorb) and the source of the last hop made; log2(32 W) rounds of one
            // shared-memory read each, cut short as soon as node 0's end stops moving.  Same cur / prev as the serial walk.
            uint32_t e = hop <= 0 ? (uint32_t)idx : (uint32_t)min(idx + hop, 32 * W);
            uint32_t pv = hop <= 0 ? 0u : (uint32_t)idx + 1u;
            int hb = 0;
            s_hop[0][idx] = e | (pv << 16);
            __syncthreads();
            uint32_t v0 = s_hop[0][0];
#pragma unroll 1
            for (int r = 1; r < 32 * W; r <<= 1) {
                const uint32_t v2 = s_hop[hb][e];
                e = v2 & 0xffffu;
                if (v2 >> 16) pv = v2 >> 16;
                hb ^= 1;
                s_hop[hb][idx] = e | (pv << 16);
                __syncthreads();
                const uint32_t v0n = s_hop[hb][0];
                const bool settled = (v0n & 0xffffu) == (v0 & 0xffffu);
                v0 = v0n;
                if (settled) break;
            }
            const int cur = (int)(v0 & 0xffffu), prev = (int)(v0 >> 16) - 1;
            const int c_end = cur < 32 * W ? s_J[cur] : 1;
            const bool done = c_end < 0, found = c_end == 0;
            const float last_land = prev >= 0 ? s_land[prev] : t_mid;
            const float tm_found = found ? s_tm[cur] : 0.f;
            __syncthreads();   // the window arrays are rewritten by the next window
            if (done) break;
            if (found) {
                if (cur > 0) {   // reached by a skip: t0/t1 are re-centred on t_mid (CS/ray_marching.cu:173-174)
                    t_mid = tm_found;
                    t0 = __fmaf_rn(step, -0.5f, t_mid);
                    t1 = __fmaf_rn(step, 0.5f, t_mid);
                    chain_open = false;
                }
                occ_mode = true;
                continue;
            }
            t_mid = last_land;
            t0 = __fmaf_rn(step, -0.5f, t_mid);
            t1 = __fmaf_rn(step, 0.5f, t_mid);
            chain_open = false;
            continue;
        }
        // ---- occupied stretch: 31 candidate samples on the lattice t0, t1, t1 (+) dt, ...; candidate 0 is known to be occupied.
        // Every warp holds the same batch: lane i spans [l0, l1] = [t1 after i-1 adds, t1 after i adds] (i = 0: [t0, t1]).
#ifdef SNB_MARCH_DEBUG
        ++dbg_batches;
#endif
        if (!net_ready) {   // CTA-uniform
            load_net_to_smem(s_net, net.net);
            inv_s = s_net[kOffInvS];
            net_ready = true;
        }
        float l0 = t0, l1 = t1;
        {
            const AddChain c1 = make_add_chain(t1, step);
            float a = t1, bprev = t0;
            bool okc = add_chain_at(c1, (uint32_t)lane, a);
            if (lane > 0) okc = add_chain_at(c1, (uint32_t)lane - 1u, bprev) && okc;
            if (__all_sync(kAll, okc)) {
                l1 = a;
                l0 = bprev;
            } else {
                for (int s = 0; s < lane; ++s) {
                    l0 = l1;
                    l1 = __fadd_rn(l0, step);
                }
            }
        }
        const float lm = (lane == 0) ? t_mid : __fmul_rn(__fadd_rn(l0, l1), 0.5f);
        const float px = __fmaf_rn(lm, d[0], o[0]), py = __fmaf_rn(lm, d[1], o[1]), pz = __fmaf_rn(lm, d[2], o[2]);
        const bool in_range = lm < far;
        const bool occ = (lane < 31) && in_range && march_occupied(rc, px, py, pz, grid);
        const unsigned stop = ~__ballot_sync(kAll, occ);  // bit 31 always set: lane 31 only evaluates an end point
        const int f = __ffs(stop) - 1;                    // lanes [0,f) are samples, 1 <= f <= 31
        // SDF at the start of every sample and (lane f) at the end of the last one; positions as the
        // reference builds them: t_origins + t_dirs * t (models/renderer.py:84-86), separately rounded
        const float sdf = cta_sdf<W>(lane <= f, __fadd_rn(o[0], __fmul_rn(d[0], l0)), __fadd_rn(o[1], __fmul_rn(d[1], l0)),
                                     __fadd_rn(o[2], __fmul_rn(d[2], l0)), table, s_lvl, net.n_active, s_net, s_feat, s_part, warp, lane);
        const float sdf_next = __shfl_down_sync(kAll, sdf, 1);
        float fac = lane < f ? __fsub_rn(1.f, neus_alpha(sdf, sdf_next, inv_s)) : 1.f;
        // transmittance in front of every candidate: T * prod_{i<lane} (1 - alpha_i); NA/vol_rendering.py:730-748 keeps T >= eps
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            float u = __shfl_up_sync(kAll, fac, off);
            if (lane >= off) fac *= u;
        }
        const float excl = __shfl_up_sync(kAll, fac, 1);
        const float Tb = lane == 0 ? T : T * excl;
        const unsigned invisible = __ballot_sync(kAll, lane < f && !(Tb >= eps));
        int nvis = invisible ? __ffs(invisible) - 1 : f;
        T = __shfl_sync(kAll, T * fac, f - 1);
        bool ray_done = nvis < f;
        if (j + nvis > sm.scratch_stride) {
            nvis = sm.scratch_stride - j;
            overflow = true;
            ray_done = true;
        }
        if (warp == 0 && lane < nvis) {
            sc0[j + lane] = l0;
            sc1[j + lane] = l1;
        }
        if (nvis > 0 && !chain_open) ++runs;
        j += nvis;
        chain_open = (nvis == 31);
        if (ray_done) break;
        if (f == 31) {  // all 31 candidates were visible samples: continue contiguously
            t0 = __shfl_sync(kAll, l1, 30);
            t1 = __fadd_rn(t0, step);
            t_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
            float cx = __fmaf_rn(t_mid, d[0], o[0]), cy = __fmaf_rn(t_mid, d[1], o[1]), cz = __fmaf_rn(t_mid, d[2], o[2]);
            if (t_mid < far && !march_occupied(rc, cx, cy, cz, grid)) occ_mode = false;   // lane 0 of a batch must be occupied
            continue;
        }
        // candidate f is unoccupied (or beyond far: the loop condition ends the ray): back to the empty-space lattice
        t_mid = __shfl_sync(kAll, lm, f);
        occ_mode = false;
        chain_open = false;
    }
    if (threadIdx.x == 0) {
        sm.counts[ray] = j;
        sm.end_counts[ray] = runs;
        if (overflow) atomicExch(sm.totals + 2, 1);
#ifdef SNB_MARCH_DEBUG
        if (ray < 8192) {
            g_march_dbg[4 * ray] = clock64() - dbg_t0;
            g_march_dbg[4 * ray + 1] = dbg_windows;
            g_march_dbg[4 * ray + 2] = dbg_batches;
            g_march_dbg[4 * ray + 3] = j;
        }
#endif
    }
}

// One CTA: exclusive scans of the sample counts and the end counts, clipped to the capacities.
// With `stats`: also what prep_net does for the loss accumulators -- stats[0] = number of foreground pixels of the batch + 1e-5
// (the normal-loss normaliser, exp_runner.py:184-186), stats[1..7] = 0.
__global__ void __launch_bounds__(1024) scan_counts_kernel(int32_t n, snb_samples sm, int32_t n_mask, const float *__restrict__ mask,
                                                           float *__restrict__ stats) {
    __shared__ int32_t wsum[2][32];
    __shared__ float msum[32];
    float mcount = 0.f;
    if (stats)
        for (int e = threadIdx.x; e < n_mask; e += 1024) mcount += mask[e] > 0.5f ? 1.f : 0.f;
    __shared__ int32_t carry[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int32_t start = 0; start < n; start += 1024) {
        int32_t i = start + threadIdx.x;
        int32_t c[2] = {i < n ? sm.counts[i] : 0, i < n ? sm.end_counts[i] : 0};
        int32_t v[2] = {c[0], c[1]};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t u = __shfl_up_sync(0xffffffffu, v[a], o);
                if (lane >= o) v[a] += u;
            }
            if (lane == 31) wsum[a][warp] = v[a];
        }
        __syncthreads();
        if (warp < 2) {
            int32_t w = wsum[warp][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            wsum[warp][lane] = w;
        }
        __syncthreads();
        int32_t incl[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) incl[a] = v[a] + (warp ? wsum[a][warp - 1] : 0) + carry[a];
        if (i < n) {
            int64_t cap[2] = {sm.capacity, sm.end_capacity};
            int32_t *dst[2] = {sm.packed_info, sm.end_packed};
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                int64_t off = incl[a] - c[a];
                int64_t cnt = c[a];
                if (off >= cap[a]) { off = cap[a]; cnt = 0; }
                else if (off + cnt > cap[a]) cnt = cap[a] - off;
                dst[a][2 * i] = (int32_t)off;
                dst[a][2 * i + 1] = (int32_t)cnt;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry[0] = incl[0]; carry[1] = incl[1]; }
        __syncthreads();
    }
    if (stats) {   // 0/1 counts: exact in fp32 in any order (n_mask < 2^24)
#pragma unroll
        for (int o = 16; o; o >>= 1) mcount += __shfl_xor_sync(0xffffffffu, mcount, o);
        if (lane == 0) msum[warp] = mcount;
        __syncthreads();
        if (warp == 0) {
            float t = msum[lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) stats[0] = t + 1e-5f;
            else if (lane < 8) stats[lane] = 0.f;
        }
    }
    if (threadIdx.x == 0) {
        if (carry[0] > sm.capacity || carry[1] > sm.end_capacity) sm.totals[2] = 1;
        sm.totals[0] = (int32_t)min((int64_t)carry[0], sm.capacity);
        sm.totals[1] = (int32_t)min((int64_t)carry[1], sm.end_capacity);
    }
}

// warp per ray: scratch -> packed (t0,t1,patch id), end-slot assignment.
// An interval needs its own end query iff t1[i] != t0[i+1] (models/renderer.py:152-154); the last sample of a
// ray always does (the reference compares it with the next ray's first start, which never matches).
__device__ __forceinline__ void compact_ray(const snb_samples &sm, int ray, int lane, int base, int cnt, int ebase, int ecnt) {
    const float *sc0 = sm.scratch_t0 + (int64_t)ray * sm.scratch_stride;
    const float *sc1 = sm.scratch_t1 + (int64_t)ray * sm.scratch_stride;
    int eused = 0;
    for (int j0 = 0; j0 < cnt; j0 += 32) {
        int j = j0 + lane;
        bool valid = j < cnt;
        float a0 = valid ? sc0[j] : 0.f, a1 = valid ? sc1[j] : 0.f;
        float nxt = (valid && j + 1 < cnt) ? sc0[j + 1] : 0.f;
        bool diff = valid && (j + 1 >= cnt || a1 != nxt);
        unsigned dm = __ballot_sync(0xffffffffu, diff);
        int slot = -1;
        if (diff) {
            int e = eused + __popc(dm & ((1u << lane) - 1u));
            if (e < ecnt) slot = ebase + e;
        }
        eused += __popc(dm);
        if (valid) {
            sm.t0[base + j] = a0;
            sm.t1[base + j] = a1;
            sm.patch_idx[base + j] = ray;
            sm.end_slot[base + j] = slot;
            if (slot >= 0) sm.slot_sample[slot] = base + j;
        }
    }
}

__global__ void __launch_bounds__(256) compact_kernel(int32_t n, snb_samples sm) {
    const int lane = threadIdx.x & 31;
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= n) return;
    if (sm.launch_order && lane == 0) sm.launch_order[ray] = ray;   // (the one-launch form below sorts; this form is the any-n fallback)
    compact_ray(sm, ray, lane, sm.packed_info[2 * ray], sm.packed_info[2 * ray + 1], sm.end_packed[2 * ray], sm.end_packed[2 * ray + 1]);
}

// scan + compaction + loss-accumulator reset in ONE launch (n <= kCompactOneMax rays): every CTA (8 rays, warp per ray) sums the
// counts of the rays in front of it itself -- n/8 CTAs x n counts from L2 instead of a single-CTA scan kernel and a second launch --
// and one extra CTA counts the foreground pixels (stats[0]) and zeroes stats[1..7].  Same outputs as scan_counts_kernel + compact_kernel.
constexpr int kCompactOneMax = 4096;
constexpr int kOrderBins = 8;   // size classes of the render launch order: 0 samples, 1, 2, ... 6 chunks of 32, more
__device__ __forceinline__ int order_class(int count) { return count <= 0 ? 0 : min((count + 31) >> 5, kOrderBins - 1); }
__global__ void __launch_bounds__(256) compact_one_kernel(int32_t n, snb_samples sm, int32_t n_mask, const float *__restrict__ mask,
                                                          float *__restrict__ stats) {
    __shared__ int32_t s_part[2][8];
    __shared__ int32_t s_cnt[2][8];
    __shared__ int32_t s_hist[2][8][kOrderBins];
    __shared__ float s_m[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_ctas = (n + 7) >> 3;
    if ((int)blockIdx.x == n_ctas) {   // ---- stats CTA ----
        if (!stats) return;
        float c = 0.f;
        if ((reinterpret_cast<uintptr_t>(mask) & 15) == 0) {
            const float4 *m4 = reinterpret_cast<const float4 *>(mask);
            for (int e = threadIdx.x; e < (n_mask >> 2); e += 256) {
                float4 v = __ldg(m4 + e);
                c += (v.x > 0.5f ? 1.f : 0.f) + (v.y > 0.5f ? 1.f : 0.f) + (v.z > 0.5f ? 1.f : 0.f) + (v.w > 0.5f ? 1.f : 0.f);
            }
            for (int e = (n_mask & ~3) + threadIdx.x; e < n_mask; e += 256) c += mask[e] > 0.5f ? 1.f : 0.f;
        } else {
            for (int e = threadIdx.x; e < n_mask; e += 256) c += mask[e] > 0.5f ? 1.f : 0.f;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);   // 0/1 counts: exact in fp32 in any order
        if (lane == 0) s_m[warp] = c;
        __syncthreads();
        if (warp == 0) {
            float t = lane < 8 ? s_m[lane] : 0.f;
#pragma unroll
            for (int o = 4; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) stats[0] = t + 1e-5f;
            else if (lane < 8) stats[lane] = 0.f;
        }
        return;
    }
    // ---- sum of the counts in front of this CTA's 8 rays, and the 8 own counts ----
    const int first = blockIdx.x * 8;
    int32_t a0 = 0, a1 = 0;
    // launch order of the render stage (longest patches first): rank of a ray = rays in a higher size class (kOrderBins classes by
    // 32-sample chunks) + rays of its own class in front of it.  Class histograms of ALL rays and of the rays in front of this CTA,
    // by ballots while the counts stream past; class 0 (no sample: half to three quarters of a batch) follows from the totals.
    const bool want_order = sm.launch_order != nullptr;
    int h_all[kOrderBins], h_front[kOrderBins];
#pragma unroll
    for (int q = 0; q < kOrderBins; ++q) h_all[q] = h_front[q] = 0;
    const int i_end = want_order ? n : first;
    for (int i0 = 0; i0 < i_end; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const int c = i < n ? sm.counts[i] : 0;
        if (i < first) { a0 += c; a1 += sm.end_counts[i]; }
        if (want_order) {
            const int cls = order_class(c);
#pragma unroll
            for (int q = 1; q < kOrderBins; ++q) {
                h_all[q] += __popc(__ballot_sync(0xffffffffu, cls == q));
                h_front[q] += __popc(__ballot_sync(0xffffffffu, cls == q && i < first));
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    const int ray = first + warp;
    if (lane == 0) {
        s_part[0][warp] = a0; s_part[1][warp] = a1;
        s_cnt[0][warp] = ray < n ? sm.counts[ray] : 0;
        s_cnt[1][warp] = ray < n ? sm.end_counts[ray] : 0;
#pragma unroll
        for (int q = 1; q < kOrderBins; ++q) { s_hist[0][warp][q] = h_all[q]; s_hist[1][warp][q] = h_front[q]; }
    }
    __syncthreads();
    if (want_order && lane == 0 && ray < n) {
        const int mine = order_class(s_cnt[0][warp]);
        int rank = 0, nonempty_all = 0, nonempty_front = 0;
#pragma unroll
        for (int q = 1; q < kOrderBins; ++q) {
            int all_q = 0, front_q = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { all_q += s_hist[0][w][q]; front_q += s_hist[1][w][q]; }
            nonempty_all += all_q;
            nonempty_front += front_q;
            if (q > mine) rank += all_q;
            if (q == mine) rank += front_q;
        }
        if (mine == 0) rank = nonempty_all + (first - nonempty_front);
        for (int w = 0; w < warp; ++w) rank += order_class(s_cnt[0][w]) == mine ? 1 : 0;
        sm.launch_order[rank] = ray;
    }
    int64_t off[2] = {0, 0};
    int32_t own[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int w = 0; w < 8; ++w) off[a] += s_part[a][w] + (w < warp ? s_cnt[a][w] : 0);
        own[a] = s_cnt[a][warp];
    }
    if ((int)blockIdx.x == n_ctas - 1 && warp == 7 && lane == 0) {   // totals (off + own of the last warp slot = grand totals)
        int64_t t0 = off[0] + own[0], t1 = off[1] + own[1];
        if (t0 > sm.capacity || t1 > sm.end_capacity) sm.totals[2] = 1;
        sm.totals[0] = (int32_t)min(t0, sm.capacity);
        sm.totals[1] = (int32_t)min(t1, sm.end_capacity);
    }
    if (ray >= n) return;
    const int64_t cap[2] = {sm.capacity, sm.end_capacity};
    int32_t boff[2], bcnt[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {   // clip to the capacities exactly like scan_counts_kernel
        int64_t o = off[a], c = own[a];
        if (o >= cap[a]) { o = cap[a]; c = 0; }
        else if (o + c > cap[a]) c = cap[a] - o;
        boff[a] = (int32_t)o; bcnt[a] = (int32_t)c;
    }
    if (lane == 0) {
        sm.packed_info[2 * ray] = boff[0]; sm.packed_info[2 * ray + 1] = bcnt[0];
        sm.end_packed[2 * ray] = boff[1]; sm.end_packed[2 * ray + 1] = bcnt[1];
    }
    compact_ray(sm, ray, lane, boff[0], bcnt[0], boff[1], bcnt[1]);
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_march_visible(const snb_patch_batch *b, const snb_net *net, const float *roi, int32_t rx, int32_t ry,
                                     int32_t rz, const uint8_t *grid, float step, const float *jitter, float eps,
                                     const snb_samples *sm, snb_stream_t stream) {
    SNB_REQUIRE(b && net && sm, SNB_ERR_NULL, "march_visible: null struct");
    SNB_REQUIRE(b->n_patches >= 0 && step > 0.f && rx > 0 && ry > 0 && rz > 0, SNB_ERR_ARG, "march_visible: bad sizes/step");
    SNB_REQUIRE(net->n_active <= net->meta.n_levels && net->meta.n_levels <= SNB_MAX_LEVELS, SNB_ERR_ARG, "march_visible: bad level counts");
    if (b->n_patches == 0) return SNB_OK;
    SNB_REQUIRE(b->rays_o && b->rays_d && b->near_ && b->far_ && roi && grid && net->table_f16 && net->net, SNB_ERR_NULL, "march_visible: null input");
    SNB_REQUIRE(sm->counts && sm->end_counts && sm->totals && sm->scratch_t0 && sm->scratch_t1 && sm->scratch_stride > 0, SNB_ERR_NULL, "march_visible: null scratch");
    SNB_REQUIRE(aligned(net->net, 16), SNB_ERR_ALIGN, "march_visible: net must be 16-byte aligned");
    cudaMemsetAsync(sm->totals, 0, 4 * sizeof(int32_t), S(stream));
    // CTA per ray, 4 warps, compiled for 6 CTAs per SM (80 registers): measured best of W in {1,2,4,8} x {4,6,8} CTAs/SM over the schedule
    march_visible_kernel<4, 6><<<(unsigned)b->n_patches, 128, 0, S(stream)>>>(*b, *net, make_level_table(net->meta), roi, make_int3(rx, ry, rz), grid, step, jitter, eps, *sm);
    SNB_LAUNCH_CHECK("march_visible");
    return SNB_OK;
}

#ifdef SNB_MARCH_DEBUG
extern "C" int32_t snb_debug_march_stats(long long *host_out, int32_t n_rays) {
    return (int32_t)cudaMemcpyFromSymbol(host_out, g_march_dbg, sizeof(long long) * 4 * (size_t)(n_rays < 8192 ? n_rays : 8192));
}
#endif

extern "C" int32_t snb_compact_samples_stats(int32_t n, const snb_samples *sm, int32_t n_mask, const float *mask, float *stats,
                                             snb_stream_t stream) {
    SNB_REQUIRE(sm, SNB_ERR_NULL, "compact_samples: null struct");
    SNB_REQUIRE(n >= 0 && n_mask >= 0 && n_mask < (1 << 24), SNB_ERR_ARG, "compact_samples: bad n / n_mask");
    SNB_REQUIRE(sm->packed_info && sm->end_packed && sm->t0 && sm->t1 && sm->patch_idx && sm->end_slot && sm->slot_sample, SNB_ERR_NULL, "compact_samples: null buffer");
    SNB_REQUIRE(!stats || n_mask == 0 || mask, SNB_ERR_NULL, "compact_samples: null mask");
    if (n > 0 && n <= kCompactOneMax) {   // one launch: per-CTA prefix sums (n^2/8 count reads from L2) + one stats CTA
        compact_one_kernel<<<(unsigned)cdiv(n, 8) + 1, 256, 0, S(stream)>>>(n, *sm, n_mask, mask, stats);
    } else {
        scan_counts_kernel<<<1, 1024, 0, S(stream)>>>(n, *sm, n_mask, mask, stats);
        if (n) compact_kernel<<<(unsigned)cdiv(n, 8), 256, 0, S(stream)>>>(n, *sm);
    }
    SNB_LAUNCH_CHECK("compact_samples");
    return SNB_OK;
}

// the two-launch form (single-CTA scan, then compaction), any n
extern "C" int32_t snb_compact_samples(int32_t n, const snb_samples *sm, snb_stream_t stream) {
    SNB_REQUIRE(sm, SNB_ERR_NULL, "compact_samples: null struct");
    SNB_REQUIRE(n >= 0, SNB_ERR_ARG, "compact_samples: n < 0");
    SNB_REQUIRE(sm->packed_info && sm->end_packed && sm->t0 && sm->t1 && sm->patch_idx && sm->end_slot && sm->slot_sample, SNB_ERR_NULL, "compact_samples: null buffer");
    scan_counts_kernel<<<1, 1024, 0, S(stream)>>>(n, *sm, 0, nullptr, nullptr);
    if (n) compact_kernel<<<(unsigned)cdiv(n, 8), 256, 0, S(stream)>>>(n, *sm);
    SNB_LAUNCH_CHECK("compact_samples");
    return SNB_OK;
}
