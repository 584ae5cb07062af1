// Microbenchmark: throughput of no-return global reductions (RED) into an L2-resident table on sm_100a -- the operation that bounds
// the hash-table gradient scatter of sdf_bwd_patch (8 corners x L levels per point).  Answers: how many reduction lanes per clock
// does the chip retire for .f32 / .v2.f32 / .v4.f32 / .f16x2 operands, for tables of 4 MB (one hashed level) and 47.5 MB (all 14
// levels), when lanes are scattered uniformly vs. when pairs / quads of lanes share a 32-byte sector?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate scripts/micro/red_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// KIND 0: red.f32   1: red.v2.f32   2: red.v4.f32   3: red.f16x2   4: red.v2.f32 with the 8-corner pattern of a hashed level
// (lane -> cell; 8 reds at px ^ py*P1 ^ pz*P2 over the corner offsets)
// share: 1 = every lane its own random slot; 2 / 4 = groups of 2 / 4 adjacent lanes hit adjacent slots of one 32-byte sector (v2: 4 slots
// per sector)
template <int KIND>
__global__ void __launch_bounds__(256) k(float *table, uint32_t n_slots_mask, int iters, int share, uint32_t seed) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid / share + seed * 0x9e3779b9u);
    const uint32_t sub = tid % share;
    for (int it = 0; it < iters; ++it) {
        s = mix(s + 0x632be5abu);
        if (KIND == 4) {
            const uint32_t px = s & 1023u, py = (s >> 10) & 1023u, pz = (s >> 20) & 1023u;
#pragma unroll
            for (uint32_t c = 0; c < 8; ++c) {
                const uint32_t idx = ((px + (c & 1u)) ^ ((py + ((c >> 1) & 1u)) * 2654435761u) ^ ((pz + (c >> 2)) * 805459861u)) & n_slots_mask;
                atomicAdd(reinterpret_cast<float2 *>(table) + idx, make_float2(1.f, 2.f));
            }
            continue;
        }
        const uint32_t slot = (((s & ~(uint32_t)(share - 1)) | sub)) & n_slots_mask;
        if (KIND == 0) atomicAdd(table + slot, 1.f);
        else if (KIND == 1) atomicAdd(reinterpret_cast<float2 *>(table) + slot, make_float2(1.f, 2.f));
        else if (KIND == 2) atomicAdd(reinterpret_cast<float4 *>(table) + slot, make_float4(1.f, 2.f, 3.f, 4.f));
        else if (KIND == 3) atomicAdd(reinterpret_cast<__half2 *>(table) + slot, __floats2half2_rn(1.f, 2.f));
    }
}

template <int KIND>
void run(const char *name, size_t table_bytes, int elem_bytes, int share, int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float *table; cudaMalloc(&table, table_bytes); cudaMemset(table, 0, table_bytes);
    uint32_t n_slots = (uint32_t)(table_bytes / elem_bytes);
    uint32_t p2 = 1; while (p2 * 2 <= n_slots) p2 *= 2;
    const int iters = KIND == 4 ? 64 : 512;
    const int blocks = sms * warps_per_sm / 8;
    k<KIND><<<blocks, 256>>>(table, p2 - 1, 8, share, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<KIND><<<blocks, 256>>>(table, p2 - 1, iters, share, 2 + rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double lanes = (double)blocks * 256 * iters * (KIND == 4 ? 8 : 1);
    printf("%-22s table %6.1f MB share %d warps/SM %2d : %8.3f ms  %7.2f Glanes/s  %6.1f lanes/clk (at %d MHz)  %7.1f GB/s payload\n", name,
           p2 * (double)elem_bytes / 1e6, share, warps_per_sm, best, lanes / (best * 1e-3) / 1e9, lanes / (best * 1e-3) / (clk_khz * 1e3), clk_khz / 1000,
           lanes * elem_bytes / (best * 1e-3) / 1e9);
    cudaFree(table);
}

int main() {
    const size_t MB = 1 << 20;
    for (size_t tb : {4 * MB, 48 * MB}) {
        for (int share : {1, 2, 4}) {
            run<0>("red.f32", tb, 4, share, 32);
            run<1>("red.v2.f32", tb, 8, share, 32);
            run<2>("red.v4.f32", tb, 16, share > 2 ? 2 : share, 32);
            run<3>("red.f16x2", tb, 4, share, 32);
        }
        run<4>("red.v2.f32 8-corner", tb, 8, 1, 32);
    }
    for (int w : {8, 16, 64}) run<1>("red.v2.f32", 48 * MB, 8, 1, w);
    return 0;
}
