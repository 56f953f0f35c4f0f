// Microbenchmark: issue rate of the packed complex multiply-add pattern of the chain kernel
// (acc.packed += a.scalar * b.packed) as a function of warps per scheduler, against scalar FFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long Pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

__constant__ float4 CB[1024];

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, const float4 *bsrc, int iters, long long *cyc)
{
    __shared__ float4 B[64];
    if (threadIdx.x < 64)
        B[threadIdx.x] = bsrc[threadIdx.x];
    __syncthreads();
    float2 E[16];
    unsigned long long acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        E[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
        acc[i] = 0ull;
    }
    float facc[32];
#pragma unroll
    for (int i = 0; i < 32; i++)
        facc[i] = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const float4 bb = MODE == 2 ? B[(k * 4 + n + it) & 63] : MODE == 4 ? CB[(k * 4 + n + it) & 1023] : bsrc[(k * 4 + n) & 63];
                if (MODE <= 2 || MODE == 4) {
                    const unsigned long long b0 = Pack2(bb.x, bb.y), b1 = Pack2(bb.z, bb.w);
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const float2 a = E[g * 4 + k];
                        const unsigned long long ax = Pack2(a.x, a.x), ay = Pack2(a.y, a.y);
                        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[g * 4 + n]) : "l"(ax), "l"(b0));
                        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[g * 4 + n]) : "l"(ay), "l"(b1));
                    }
                }
                else {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const float2 a = E[g * 4 + k];
                        float &x = facc[2 * (g * 4 + n)], &y = facc[2 * (g * 4 + n) + 1];
                        x = fmaf(a.x, bb.x, x);
                        x = fmaf(a.y, bb.z, x);
                        y = fmaf(a.x, bb.y, y);
                        y = fmaf(a.y, bb.w, y);
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi + facc[2 * i] + facc[2 * i + 1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int threads, float *out, float4 *b, long long *cyc)
{
    const int iters = 2000;
    k<MODE><<<148, threads>>>(out, b, iters, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per warp per iteration: 128 FFMA2 (or 256 FFMA) = 256 FMA-pipe cycles
    const double per_iter = double(h) / iters;
    const int warps_per_smsp = threads / 128;
    printf("%-28s warps/SMSP=%d  cycles/iter=%.1f  pipe util=%.2f\n", name, warps_per_smsp, per_iter,
           256.0 * warps_per_smsp / per_iter);
}

int main()
{
    float *out;
    float4 *b;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&b, 64 * 16);
    cudaMalloc(&cyc, 8);
    cudaMemset(b, 0, 64 * 16);
    for (int threads : {128, 256, 512, 1024}) {
        run<1>("FFMA2, B global/L1 hoisted", threads, out, b, cyc);
        run<2>("FFMA2, B from smem (LDS.128)", threads, out, b, cyc);
        run<3>("FFMA scalar", threads, out, b, cyc);
        run<4>("FFMA2, B from constant bank", threads, out, b, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
