// umma.cuh -- minimal hand-written tcgen05 (5th-gen tensor core) plumbing for sm_100a: shared-memory matrix descriptors for
// the canonical K-major no-swizzle layout, the kind::tf32 instruction descriptor, TMEM allocation, MMA issue, commit to an
// mbarrier and TMEM -> register loads.  Bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables
// (the same fields CUTLASS's cute/arch/mma_sm100_desc.hpp names); nothing here depends on CUTLASS.
//
// Operand layout used throughout (4-byte elements, "K-major, INTERLEAVE / no swizzle"):
//   core matrix = 8 rows x 16 bytes (4 elements), stored as 128 contiguous bytes (row r at +16 r);
//   a tile [R x K] puts core (r/8, k/4) at byte (r/8) * SBO + (k/4) * LBO with LBO = 128, SBO = (K/4) * 128;
//   one tcgen05.mma of kind::tf32 consumes K = 8 elements = 2 core matrices along K, so k-step s starts at +256 s bytes.
// Accumulator: M = 128 rows <-> TMEM lanes 0..127, N fp32 columns; warp w (w = warp id mod 4) reads lanes 32 w .. 32 w + 31
// with tcgen05.ld.32x32b, i.e. thread t of the CTA owns row t: exactly the thread-per-point layout of the SDF kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace snb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr uint32_t tile_bytes(int rows, int K) { return (uint32_t)rows * (uint32_t)K * 4u; }
// byte offset of element (r, k) of a K-major tile with K columns (K % 4 == 0, rows % 8 == 0)
__device__ __forceinline__ uint32_t kmajor_off(int r, int k, int K) {
    return (uint32_t)(((r >> 3) * (K >> 2) + (k >> 2)) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

// matrix descriptor: start address [0,14) | LBO [16,30) | SBO [32,46) (all >> 4) | version = 1 [46,48) | layout_type = 0 (no swizzle) [61,64)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t smem_byte_addr, int K) {
    uint64_t d = (uint64_t)((smem_byte_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)((((uint32_t)(K >> 2) * 128u) >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// The same layout with a padded leading-dimension byte offset LBO (> 128, multiple of 16): core (r/8, k/4) at (r/8) * (K/4) * LBO +
// (k/4) * LBO.  tcgen05 only sees a different stride in the descriptor; the threads gain conflict-free TRANSPOSED reads of the tile:
// the mma.sync fragments of the weight-gradient contraction read 4-byte words at (k/4 varies with lane bit 2) -- with LBO = 128 lanes
// g and g + 4 of a fragment load hit the same bank (2 wavefronts per LDS, 8.7 M conflicts per launch in
// profiles/r02_ncu_bwd_split_v2_it4800.txt); with LBO = 192 (48 words = 16 mod 32) the 32 lanes cover the 32 banks.
template <int LBO>
__host__ __device__ constexpr uint32_t tile_bytes_lbo(int rows, int K) { return (uint32_t)(rows / 8) * (uint32_t)(K / 4) * (uint32_t)LBO; }
template <int LBO>
__device__ __forceinline__ uint32_t kmajor_off_lbo(int r, int k, int K) {
    return (uint32_t)(((r >> 3) * (K >> 2) + (k >> 2)) * LBO + (r & 7) * 16 + (k & 3) * 4);
}
template <int LBO>
__device__ __forceinline__ uint64_t kmajor_desc_lbo(uint32_t smem_byte_addr, int K) {
    uint64_t d = (uint64_t)((smem_byte_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((uint32_t)LBO >> 4) << 16;
    d |= (uint64_t)((((uint32_t)(K >> 2) * (uint32_t)LBO) >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// MN-major TF32 operands are NOT available in this layout: tcgen05 accepts them only with the 128B_BASE32B swizzle (measured:
// wrong results with a no-swizzle MN-major descriptor), and for M = 128 the N extent must be a multiple of 16.  Contractions over
// the row index of these tiles therefore go through warp-level mma.sync (fused_sdf.cu, step (4) of the backward).
// instruction descriptor, kind::tf32, fp32 accumulate, both operands K-major:
// c_format = F32 (1) [4,6) | a_format = TF32 (2) [7,10) | b_format = TF32 (2) [10,13) | N >> 3 [17,23) | M >> 4 [24,29)
// a_major [15] / b_major [16]: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (one warp, all lanes) ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}

// ---- fences / barriers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_smem_to_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// spin on the phase parity; returns false if the barrier did not flip within ~2^26 polls (a fault upstream must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

// ---- MMA issue (ONE thread) and commit -------------------------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]^T   (A: M x 8, B: N x 8, both K-major)
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all tcgen05.mma issued so far by this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, 16 consecutive columns per call (warp-collective) -------------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float &v0, float &v1) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
    v0 = __uint_as_float(r0);
    v1 = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace umma
}  // namespace snb
