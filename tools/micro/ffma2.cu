// Microbenchmark: does fma.rn.f32x2 (FFMA2) raise FP32 FMA throughput on sm_100a?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a ffma2.cu -o ffma2 && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_ffma(float *out, float a, float b, int iters)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float *out, float a, float b, int iters)
{
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float lo = threadIdx.x + i, hi = threadIdx.x - i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi));
    }
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 16;
    const int grid = 148 * 8;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k_ffma<<<grid, 256>>>(out, 1.0001f, 0.5f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * double(grid) * 256;
        printf("FFMA   : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0);
        k_ffma2<<<grid, 256>>>(out, 1.0001f, 0.5f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2  : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
