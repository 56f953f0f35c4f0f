// Throughput of the legacy warp-level tensor path on sm_100a: mma.sync.m16n8k8 TF32 (and m16n8k4 / DMMA m8n8k4
// for reference), independent accumulators, no memory traffic.  Prints TFLOP/s for several warps-per-SM counts.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/mma_tf32.cu -o /tmp/mma_tf32 && /tmp/mma_tf32
#include <cstdio>
#include <cuda_runtime.h>

template <int ACC> __global__ void __launch_bounds__(1024) MmaTf32(float *out, int iters)
{
    float d[ACC][4];
    for (int a = 0; a < ACC; a++)
        for (int i = 0; i < 4; i++)
            d[a][i] = threadIdx.x * 1e-9f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < ACC; a++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[a][0]), "+f"(d[a][1]), "+f"(d[a][2]), "+f"(d[a][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int a = 0; a < ACC; a++)
        for (int i = 0; i < 4; i++)
            s += d[a][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC> __global__ void __launch_bounds__(1024) Ffma2(float *out, int iters)
{
    unsigned long long d[ACC];
    for (int a = 0; a < ACC; a++)
        d[a] = threadIdx.x + a;
    unsigned long long x = 0x3f8000013f800001ull + threadIdx.x, y = 0x3f0000013f000001ull;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < ACC; a++)
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[a]) : "l"(x), "l"(y));
    }
    unsigned long long s = 0;
    for (int a = 0; a < ACC; a++)
        s ^= d[a];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * 1024 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int kind = 0; kind < 2; kind++) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (kind == 0)
                    MmaTf32<8><<<sms, warps * 32>>>(out, iters);
                else
                    Ffma2<16><<<sms, warps * 32>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = ms < best ? ms : best;
            }
            const double flops = kind == 0 ? double(sms) * warps * iters * 8 * (16.0 * 8 * 8 * 2)
                                           : double(sms) * warps * 32 * iters * 16 * 4.0;
            printf("%s warps/SM %2d: %8.3f ms  %8.1f TFLOP/s  (%.0f flop/clk/SM at 1.965 GHz)\n",
                   kind == 0 ? "mma.sync m16n8k8 tf32" : "fma.rn.f32x2         ", warps, best, flops / best / 1e9,
                   flops / (best * 1e-3) / sms / 1.965e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
