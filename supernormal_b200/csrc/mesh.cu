// mesh.cu -- SDF lattice query + marching cubes on the device (BASELINE config 5).
//
// Reference: models/renderer.py:9-34 -- extract_fields evaluates -sdf on a res^3 lattice in 64^3 chunks with a
// .cpu().numpy() round trip per chunk (512 syncs at 512^3), then PyMCubes runs marching cubes on the HOST.  Here the
// lattice is evaluated by the fused encode+MLP kernel straight into HBM (one launch per x-slab), and marching cubes
// runs where the data is: classify -> two device scans -> emit, with one shared vertex per crossed lattice edge
// (like PyMCubes) and no host round trip except reading three totals to size the outputs.
// Slabs: a rank owns lattice planes [x0, x0+nx) whose LAST plane is the halo shared with the next rank; vertex
// numbering (contract shared with oracle/mc.py and dp.merge_slab_meshes):
//   [ y/z-edge vertices of planes 0..nx-2 | x-edge vertices | y/z-edge vertices of plane nx-1 ]
#include "sdf_core.cuh"
#include "mc_tables.cuh"

namespace snb {

__global__ void __launch_bounds__(256, 2) sdf_grid_query_kernel(const float *__restrict__ xs, int nx, const float *__restrict__ ys, int ny,
                                                             const float *__restrict__ zs, int nz, snb_net net, LevelTable lt, int mode,
                                                             float *__restrict__ out) {
    __shared__ __align__(16) float s_net[kNetFloats];
    load_net_to_smem(s_net, net.net);
    const LevelCtx *s_lvl = lt.lv;
    const __half2 *table = reinterpret_cast<const __half2 *>(net.table_f16);
    const int64_t plane = (int64_t)ny * nz, n = plane * nx;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / plane);
        const int64_t r = p - (int64_t)i * plane;
        const int j = (int)(r / nz), k = (int)(r - (int64_t)j * nz);
        float s = sdf_point<false>(__ldg(xs + i), __ldg(ys + j), __ldg(zs + k), table, s_lvl, net.n_active, s_net, nullptr);
        out[p] = mode == 1 ? sigmoidf_(-s * 80.f) : (mode == 2 ? -s : s);
    }
}

struct McDims {
    int nx, ny, nz;
    int64_t plane, nE, nT;   // nE = 2*(nx-1)*plane + plane vertex-count slots, nT = (nx-1)*plane cell slots
};

__host__ __device__ inline McDims mc_dims(int nx, int ny, int nz) {
    McDims d;
    d.nx = nx; d.ny = ny; d.nz = nz;
    d.plane = (int64_t)ny * nz;
    d.nE = 2 * (int64_t)(nx - 1) * d.plane + d.plane;
    d.nT = (int64_t)(nx - 1) * d.plane;
    return d;
}

__device__ __forceinline__ int64_t slotA(const McDims &d, int i, int64_t jk) {
    return i < d.nx - 1 ? (int64_t)i * d.plane + jk : 2 * (int64_t)(d.nx - 1) * d.plane + jk;
}
__device__ __forceinline__ int64_t slotB(const McDims &d, int i, int64_t jk) { return (int64_t)(d.nx - 1 + i) * d.plane + jk; }

// per lattice point: how many vertices it owns (crossed +y/+z edges -> slot A, crossed +x edge -> slot B) and how many
// triangles its cell emits
__global__ void __launch_bounds__(256) mc_classify_kernel(const float *__restrict__ u, McDims d, float iso, uint8_t *__restrict__ cntE,
                                                          uint8_t *__restrict__ cntT) {
    const int64_t n = d.plane * d.nx;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / d.plane);
        const int64_t jk = p - (int64_t)i * d.plane;
        const int j = (int)(jk / d.nz), k = (int)(jk - (int64_t)j * d.nz);
        const bool hx = i + 1 < d.nx, hy = j + 1 < d.ny, hz = k + 1 < d.nz;
        const bool in0 = __ldg(u + p) > iso;
        const bool inx = hx && (__ldg(u + p + d.plane) > iso), iny = hy && (__ldg(u + p + d.nz) > iso), inz = hz && (__ldg(u + p + 1) > iso);
        cntE[slotA(d, i, jk)] = (uint8_t)((hy && iny != in0) + (hz && inz != in0));
        if (hx) {
            cntE[slotB(d, i, jk)] = (uint8_t)(inx != in0);
            int nt = 0;
            if (hy && hz) {
                int c = (int)in0 | ((int)inx << 1) | ((int)iny << 2) | ((int)inz << 4);
                c |= (int)(__ldg(u + p + d.plane + d.nz) > iso) << 3;
                c |= (int)(__ldg(u + p + d.plane + 1) > iso) << 5;
                c |= (int)(__ldg(u + p + d.nz + 1) > iso) << 6;
                c |= (int)(__ldg(u + p + d.plane + d.nz + 1) > iso) << 7;
                nt = kMcNumTris[c];
            }
            cntT[p] = (uint8_t)nt;
        }
    }
}

// ---- exclusive scan of a uint8 array into int32 offsets: block sums -> scan of the sums -> offsets ------------
constexpr int kScanItems = 16, kScanThreads = 256, kScanBlock = kScanItems * kScanThreads;

__device__ __forceinline__ int block_incl_scan(int v, int *s_w) {   // 256 threads; returns the inclusive scan, s_w[8] = total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) { int t = s_w[w]; s_w[w] = acc; acc += t; }
        s_w[8] = acc;
    }
    __syncthreads();
    return v + s_w[warp];
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(const uint8_t *__restrict__ in, int64_t n, int *__restrict__ sums) {
    __shared__ int s_w[9];
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
    int v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q)
        if (base + q < n) v += in[base + q];
    block_incl_scan(v, s_w);
    if (threadIdx.x == 0) sums[blockIdx.x] = s_w[8];
}

// single CTA: sums -> exclusive offsets (in place), grand total -> *total
__global__ void __launch_bounds__(1024) scan_top_kernel(int *__restrict__ sums, int nb, long long *__restrict__ total) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int own = b < nb ? sums[b] : 0;
        int v = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) s_w[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const int incl = v + (warp ? s_w[warp - 1] : 0) + s_carry;
        if (b < nb) sums[b] = incl - own;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_write_kernel(const uint8_t *__restrict__ in, int64_t n, const int *__restrict__ sums,
                                                                  int *__restrict__ out) {
    __shared__ int s_w[9];
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
    int item[kScanItems], v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        item[q] = base + q < n ? in[base + q] : 0;
        v += item[q];
    }
    int excl = block_incl_scan(v, s_w) - v + sums[blockIdx.x];
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        if (base + q < n) out[base + q] = excl;
        excl += item[q];
    }
}

__global__ void __launch_bounds__(256) mc_emit_kernel(const float *__restrict__ u, McDims d, float iso, float x_offset,
                                                      const int *__restrict__ offE, const int *__restrict__ offT, int64_t v_cap,
                                                      int64_t t_cap, float *__restrict__ verts, int *__restrict__ tris) {
    const int64_t n = d.plane * d.nx;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / d.plane);
        const int64_t jk = p - (int64_t)i * d.plane;
        const int j = (int)(jk / d.nz), k = (int)(jk - (int64_t)j * d.nz);
        const bool hx = i + 1 < d.nx, hy = j + 1 < d.ny, hz = k + 1 < d.nz;
        const float v0 = __ldg(u + p);
        const bool in0 = v0 > iso;
        const float vx = hx ? __ldg(u + p + d.plane) : v0, vy = hy ? __ldg(u + p + d.nz) : v0, vz = hz ? __ldg(u + p + 1) : v0;
        const bool cx = hx && ((vx > iso) != in0), cy = hy && ((vy > iso) != in0), cz = hz && ((vz > iso) != in0);
        const float fi = (float)i + x_offset, fj = (float)j, fk = (float)k;
        if (cy | cz) {
            int64_t id = offE[slotA(d, i, jk)];
            if (cy) {
                if (id < v_cap) { verts[3 * id] = fi; verts[3 * id + 1] = __fadd_rn(fj, __fdiv_rn(__fsub_rn(iso, v0), __fsub_rn(vy, v0))); verts[3 * id + 2] = fk; }
                ++id;
            }
            if (cz && id < v_cap) { verts[3 * id] = fi; verts[3 * id + 1] = fj; verts[3 * id + 2] = __fadd_rn(fk, __fdiv_rn(__fsub_rn(iso, v0), __fsub_rn(vz, v0))); }
        }
        if (cx) {
            int64_t id = offE[slotB(d, i, jk)];
            if (id < v_cap) { verts[3 * id] = __fadd_rn(fi, __fdiv_rn(__fsub_rn(iso, v0), __fsub_rn(vx, v0))); verts[3 * id + 1] = fj; verts[3 * id + 2] = fk; }
        }
        if (!(hx && hy && hz)) continue;
        int c = (int)in0 | ((int)(vx > iso) << 1) | ((int)(vy > iso) << 2) | ((int)(vz > iso) << 4);
        c |= (int)(__ldg(u + p + d.plane + d.nz) > iso) << 3;
        c |= (int)(__ldg(u + p + d.plane + 1) > iso) << 5;
        c |= (int)(__ldg(u + p + d.nz + 1) > iso) << 6;
        c |= (int)(__ldg(u + p + d.plane + d.nz + 1) > iso) << 7;
        const int nt = kMcNumTris[c];
        if (nt == 0) continue;
        int64_t t0 = offT[p];
        for (int t = 0; t < nt; ++t, ++t0) {
            if (t0 >= t_cap) break;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int e = kMcTriTable[c][3 * t + a];
                const int axis = e >> 2, lo = e & 1, hi = (e >> 1) & 1;
                const int oi = axis == 0 ? 0 : lo, oj = axis == 0 ? lo : (axis == 1 ? 0 : hi), ok = axis == 2 ? 0 : hi;
                const int pi = i + oi;
                const int64_t pjk = jk + (int64_t)oj * d.nz + ok;
                int id;
                if (axis == 0) {
                    id = offE[slotB(d, pi, pjk)];
                } else {
                    id = offE[slotA(d, pi, pjk)];
                    if (axis == 2) {   // the z-edge vertex follows the owner's y-edge vertex, if it has one
                        const int64_t q = (int64_t)pi * d.plane + pjk;
                        const bool has_y = (pjk / d.nz) + 1 < d.ny && ((__ldg(u + q + d.nz) > iso) != (__ldg(u + q) > iso));
                        id += has_y;
                    }
                }
                tris[3 * t0 + a] = id;
            }
        }
    }
}

static unsigned grid_for(int64_t n, int threads, int per_sm) {
    int64_t b = cdiv(n, threads);
    int64_t cap = (int64_t)kNumSMs * per_sm;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

struct McWorkspace {
    uint8_t *cntE, *cntT;
    int *offE, *offT, *sumsE, *sumsT;
    long long *totals;   // [4]: n_vertices, n_main, n_triangles, 0
    size_t bytes;
};

static McWorkspace mc_layout(void *base, const McDims &d) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    McWorkspace w;
    size_t o = 0;
    uint8_t *b = (uint8_t *)base;
    w.totals = (long long *)(b + o); o += up(4 * sizeof(long long));
    w.cntE = b + o; o += up((size_t)d.nE);
    w.cntT = b + o; o += up((size_t)d.nT);
    w.offE = (int *)(b + o); o += up((size_t)(d.nE + 1) * 4);
    w.offT = (int *)(b + o); o += up((size_t)(d.nT + 1) * 4);
    w.sumsE = (int *)(b + o); o += up((size_t)cdiv(d.nE, kScanBlock) * 4 + 4);
    w.sumsT = (int *)(b + o); o += up((size_t)cdiv(d.nT, kScanBlock) * 4 + 4);
    w.bytes = o;
    return w;
}

__global__ void mc_totals_kernel(const int *__restrict__ offE, const uint8_t *__restrict__ cntE, McDims d, long long *totals) {
    // n_main = vertices numbered before the last plane's y/z block
    totals[1] = offE[2 * (int64_t)(d.nx - 1) * d.plane];
    totals[3] = 0;
}

}  // namespace snb
using namespace snb;

extern "C" int32_t snb_sdf_grid_query(const float *xs, int32_t nx, const float *ys, int32_t ny, const float *zs, int32_t nz,
                                      const snb_net *net, int32_t mode, float *out, snb_stream_t stream) {
    SNB_REQUIRE(net && net->table_f16 && net->net, SNB_ERR_NULL, "sdf_grid_query: null net");
    SNB_REQUIRE(nx >= 0 && ny >= 0 && nz >= 0 && mode >= 0 && mode <= 2, SNB_ERR_ARG, "sdf_grid_query: bad sizes/mode");
    SNB_REQUIRE(net->n_active <= net->meta.n_levels && net->meta.n_levels <= SNB_MAX_LEVELS, SNB_ERR_ARG, "sdf_grid_query: bad level counts");
    const int64_t n = (int64_t)nx * ny * nz;
    if (n == 0) return SNB_OK;
    SNB_REQUIRE(xs && ys && zs && out, SNB_ERR_NULL, "sdf_grid_query: null buffer");
    SNB_REQUIRE(aligned(net->net, 16), SNB_ERR_ALIGN, "sdf_grid_query: net must be 16-byte aligned");
    sdf_grid_query_kernel<<<grid_for(n, 256, 16), 256, 0, S(stream)>>>(xs, nx, ys, ny, zs, nz, *net, make_level_table(net->meta), mode, out);
    SNB_LAUNCH_CHECK("sdf_grid_query");
    return SNB_OK;
}

extern "C" int64_t snb_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz) {
    if (nx < 2 || ny < 2 || nz < 2) return 0;
    return (int64_t)mc_layout(nullptr, mc_dims(nx, ny, nz)).bytes;
}

extern "C" int32_t snb_mc_count(const float *u, int32_t nx, int32_t ny, int32_t nz, float iso, void *workspace, snb_stream_t stream) {
    SNB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, SNB_ERR_ARG, "mc_count: the lattice needs at least 2 points per axis");
    SNB_REQUIRE((int64_t)nx * ny * nz * 3 < 2147483647LL, SNB_ERR_ARG, "mc_count: slab too large for int32 vertex ids (shard along x)");
    SNB_REQUIRE(u && workspace, SNB_ERR_NULL, "mc_count: null buffer");
    SNB_REQUIRE(aligned(workspace, 256), SNB_ERR_ALIGN, "mc_count: workspace must be 256-byte aligned");
    const McDims d = mc_dims(nx, ny, nz);
    const McWorkspace w = mc_layout(workspace, d);
    const int64_t n = d.plane * nx;
    mc_classify_kernel<<<grid_for(n, 256, 16), 256, 0, S(stream)>>>(u, d, iso, w.cntE, w.cntT);
    const int nbE = (int)cdiv(d.nE, kScanBlock), nbT = (int)cdiv(d.nT, kScanBlock);
    scan_sums_kernel<<<nbE, kScanThreads, 0, S(stream)>>>(w.cntE, d.nE, w.sumsE);
    scan_top_kernel<<<1, 1024, 0, S(stream)>>>(w.sumsE, nbE, w.totals + 0);
    scan_write_kernel<<<nbE, kScanThreads, 0, S(stream)>>>(w.cntE, d.nE, w.sumsE, w.offE);
    scan_sums_kernel<<<nbT, kScanThreads, 0, S(stream)>>>(w.cntT, d.nT, w.sumsT);
    scan_top_kernel<<<1, 1024, 0, S(stream)>>>(w.sumsT, nbT, w.totals + 2);
    scan_write_kernel<<<nbT, kScanThreads, 0, S(stream)>>>(w.cntT, d.nT, w.sumsT, w.offT);
    mc_totals_kernel<<<1, 1, 0, S(stream)>>>(w.offE, w.cntE, d, w.totals);
    SNB_LAUNCH_CHECK("mc_count");
    return SNB_OK;
}

extern "C" int32_t snb_mc_emit(const float *u, int32_t nx, int32_t ny, int32_t nz, float iso, float x_offset, const void *workspace,
                               int64_t v_cap, int64_t t_cap, float *vertices, int32_t *triangles, snb_stream_t stream) {
    SNB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2 && v_cap >= 0 && t_cap >= 0, SNB_ERR_ARG, "mc_emit: bad sizes");
    SNB_REQUIRE(u && workspace, SNB_ERR_NULL, "mc_emit: null buffer");
    SNB_REQUIRE((vertices || v_cap == 0) && (triangles || t_cap == 0), SNB_ERR_NULL, "mc_emit: null output");
    const McDims d = mc_dims(nx, ny, nz);
    const McWorkspace w = mc_layout(const_cast<void *>(workspace), d);
    mc_emit_kernel<<<grid_for(d.plane * nx, 256, 16), 256, 0, S(stream)>>>(u, d, iso, x_offset, w.offE, w.offT, v_cap, t_cap, vertices, triangles);
    SNB_LAUNCH_CHECK("mc_emit");
    return SNB_OK;
}
