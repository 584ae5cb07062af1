// Microbenchmark: issue rate of legacy warp-level mma.sync on sm_100a (TF32 m16n8k8, BF16/FP16 m16n8k16) vs FFMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate scripts/micro/mma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int KIND>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x + 4, threadIdx.x + 5};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fmaf(c[i][j], 1.0001f, 0.5f);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char *name, double flop_per_warp_instr, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * 256);
    const int iters = 20000;
    k<KIND><<<sms * blocks_per_sm, 256>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<sms * blocks_per_sm, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)sms * blocks_per_sm * 8 /*warps*/ * iters * 8.0 * (KIND == 3 ? 4 : 1);
    double per_sm_per_clk = instr / sms / (ms * 1e-3 * 1.965e9);
    printf("%-28s blocks/SM=%d  %8.3f ms  %6.3f warp-instr/clk/SM  %8.1f TFLOP/s\n", name, blocks_per_sm, ms, per_sm_per_clk, instr * flop_per_warp_instr / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    for (int b : {1, 2, 4}) {
        run<0>("mma.m16n8k8.tf32", 2.0 * 16 * 8 * 8, b);
        run<1>("mma.m16n8k16.bf16", 2.0 * 16 * 8 * 16, b);
        run<2>("mma.m16n8k16.f16", 2.0 * 16 * 8 * 16, b);
        run<3>("ffma", 2.0 * 32, b);
    }
    return 0;
}
