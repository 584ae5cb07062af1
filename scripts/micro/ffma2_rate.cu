// ffma2_rate.cu -- issue rate of scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100+) and bit-equality of their results.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate scripts/micro/ffma2_rate.cu ; run: ./ffma2_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ACC = 16;   // independent accumulators (scalar) / 8 packed pairs
__global__ void scalar_kernel(const float *__restrict__ w, float *out, int iters) {
    float acc[ACC], x = w[threadIdx.x & 31], y = w[32 + (threadIdx.x & 31)];
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = w[j];
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int j = 0; j < ACC; ++j) acc[j] = fmaf(acc[j], x, y);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packed_kernel(const float *__restrict__ w, float *out, int iters) {
    uint64_t acc[ACC / 2];
    const float x = w[threadIdx.x & 31], y = w[32 + (threadIdx.x & 31)];
    const uint64_t xx = pk(x, x), yy = pk(y, y);
#pragma unroll
    for (int j = 0; j < ACC / 2; ++j) acc[j] = pk(w[2 * j], w[2 * j + 1]);
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int j = 0; j < ACC / 2; ++j) acc[j] = fma2(acc[j], xx, yy);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ACC / 2; ++j) { float a, b; upk(acc[j], a, b); s += a; s += b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void equal_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ c, int n, int *bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 >= n) return;
    float r0, r1;
    upk(fma2(pk(a[2 * i], a[2 * i + 1]), pk(b[2 * i], b[2 * i + 1]), pk(c[2 * i], c[2 * i + 1])), r0, r1);
    if (__float_as_uint(r0) != __float_as_uint(fmaf(a[2 * i], b[2 * i], c[2 * i])) ||
        __float_as_uint(r1) != __float_as_uint(fmaf(a[2 * i + 1], b[2 * i + 1], c[2 * i + 1]))) atomicAdd(bad, 1);
}
int main() {
    float *w, *out; int *bad;
    cudaMalloc(&w, 4096); cudaMalloc(&out, sizeof(float) * 148 * 16 * 1024); cudaMalloc(&bad, 4);
    float hw[64]; for (int i = 0; i < 64; ++i) hw[i] = 0.5f + 0.001f * i;
    cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {4, 8, 16, 32}) {          // warps per SM
        for (int which = 0; which < 2; ++which) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) scalar_kernel<<<148, 32 * wps>>>(w, out, iters); else packed_kernel<<<148, 32 * wps>>>(w, out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double fmas = 148.0 * 32 * wps * (double)iters * ACC;
            printf("%s warps/SM=%2d  %.3f ms  %.1f GFMA/s  (%.2f FMA/clk/SM at 1.965 GHz)\n", which ? "FFMA2 " : "FFMA  ", wps, best,
                   fmas / best / 1e6, fmas / best / 1e6 / 148 / 1.965);
        }
    }
    // bit-equality on random operands (incl. denormal-range products)
    const int n = 1 << 22;
    float *a, *b, *c; cudaMalloc(&a, 4 * n); cudaMalloc(&b, 4 * n); cudaMalloc(&c, 4 * n);
    float *h = new float[3 * n]; uint32_t s = 12345u;
    for (int i = 0; i < 3 * n; ++i) { s = s * 1664525u + 1013904223u; uint32_t e = 90 + (s >> 8) % 70; h[i] = __builtin_bit_cast(float, (s & 0x807fffffu) | (e << 23)); }
    for (int i = 0; i < 4096; ++i) h[i] = __builtin_bit_cast(float, (uint32_t)(i * 2654435761u) & 0x807fffffu | ((i % 40) << 23));   // tiny values
    cudaMemcpy(a, h, 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(b, h + n, 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(c, h + 2 * n, 4 * n, cudaMemcpyHostToDevice);
    cudaMemset(bad, 0, 4);
    equal_kernel<<<n / 2 / 256, 256>>>(a, b, c, n, bad);
    int hb; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("fma.rn.f32x2 vs fmaf: %d mismatching pairs of %d\n", hb, n / 2);
    return 0;
}
