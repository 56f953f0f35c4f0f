// Peak probes: the roofline denominators bench.py reports beside MEASURED_PEAKS.json (which holds only the HBM copy
// bandwidth and the bf16 tensor rate): FP32 / FP64 FMA issue rate, the FP64 tensor pipe (DMMA), the legacy
// warp-level TF32 tensor path (mma.sync), and a device copy.  Each probe runs dependent-free instruction streams
// from registers (no memory traffic) on every SM and is timed with CUDA events.
#include <algorithm>

#include "common.cuh"

namespace jb {
namespace {

template <int ACC> __global__ void __launch_bounds__(512) ProbeFfma2(float *out, int iters)
{
    unsigned long long d[ACC];
#pragma unroll
    for (int a = 0; a < ACC; a++)
        d[a] = threadIdx.x + a;
    const unsigned long long x = 0x3f8000013f800001ull + threadIdx.x, y = 0x3f0000013f000001ull;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < ACC; a++)
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[a]) : "l"(x), "l"(y));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int a = 0; a < ACC; a++)
        s ^= d[a];
    out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
}

template <int ACC> __global__ void __launch_bounds__(512) ProbeDfma(float *out, int iters)
{
    double d[ACC];
#pragma unroll
    for (int a = 0; a < ACC; a++)
        d[a] = threadIdx.x + a;
    const double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.5000001;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < ACC; a++)
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[a]) : "d"(x), "d"(y));
    }
    double s = 0;
#pragma unroll
    for (int a = 0; a < ACC; a++)
        s += d[a];
    out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
}

template <int ACC> __global__ void __launch_bounds__(512) ProbeDmma(float *out, int iters)
{
    double d[ACC][2];
#pragma unroll
    for (int a = 0; a < ACC; a++)
        d[a][0] = d[a][1] = threadIdx.x * 1e-9;
    const double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.5000001;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < ACC; a++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d[a][0]), "+d"(d[a][1])
                         : "d"(x), "d"(y));
    }
    double s = 0;
#pragma unroll
    for (int a = 0; a < ACC; a++)
        s += d[a][0] + d[a][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
}

template <int ACC> __global__ void __launch_bounds__(512) ProbeMmaTf32(float *out, int iters)
{
    float d[ACC][4];
#pragma unroll
    for (int a = 0; a < ACC; a++)
        for (int i = 0; i < 4; i++)
            d[a][i] = threadIdx.x * 1e-9f;
    const unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < ACC; a++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[a][0]), "+f"(d[a][1]), "+f"(d[a][2]), "+f"(d[a][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
#pragma unroll
    for (int a = 0; a < ACC; a++)
        for (int i = 0; i < 4; i++)
            s += d[a][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ProbeCopy(const uint4 *__restrict__ in, uint4 *__restrict__ out, long long n)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = in[i];
}

} // namespace
} // namespace jb

using namespace jb;

extern "C" int jb_probe_peak(int kind, double *value)
{
    JB_REQUIRE(value != nullptr, "probe: null argument");
    const int sms = NumSMs();
    constexpr int kThreads = 512, kIters = 8192;
    float *scratch = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    JB_CUDA(cudaEventCreate(&e0));
    JB_CUDA(cudaEventCreate(&e1));
    double best_ms = 1e30, work = 0.0; // work: flops (kinds 0-3) or bytes (kind 4) per launch
    uint4 *src = nullptr, *dst = nullptr;
    const long long copy_n = (1ll << 30) / 16; // 1 GiB each way
    if (kind == JB_PEAK_HBM_COPY) {
        JB_CUDA(cudaMalloc(reinterpret_cast<void **>(&src), size_t(1) << 30));
        JB_CUDA(cudaMalloc(reinterpret_cast<void **>(&dst), size_t(1) << 30));
        JB_CUDA(cudaMemset(src, 1, size_t(1) << 30));
    }
    else {
        JB_CUDA(cudaMalloc(reinterpret_cast<void **>(&scratch), sizeof(float) * sms * kThreads));
    }
    int rc = 0;
    for (int rep = 0; rep < 4 && rc == 0; rep++) {
        cudaEventRecord(e0);
        switch (kind) {
        case JB_PEAK_FP32_FMA:
            ProbeFfma2<16><<<sms, kThreads>>>(scratch, kIters);
            work = double(sms) * kThreads * kIters * 16 * 4.0;
            break;
        case JB_PEAK_FP64_FMA:
            ProbeDfma<16><<<sms, kThreads>>>(scratch, kIters / 4);
            work = double(sms) * kThreads * (kIters / 4) * 16 * 2.0;
            break;
        case JB_PEAK_FP64_DMMA:
            ProbeDmma<8><<<sms, kThreads>>>(scratch, kIters / 4);
            work = double(sms) * (kThreads / 32) * (kIters / 4) * 8 * (8.0 * 8 * 4 * 2);
            break;
        case JB_PEAK_TF32_MMA_SYNC:
            ProbeMmaTf32<8><<<sms, kThreads>>>(scratch, kIters);
            work = double(sms) * (kThreads / 32) * kIters * 8 * (16.0 * 8 * 8 * 2);
            break;
        case JB_PEAK_HBM_COPY:
            ProbeCopy<<<sms * 8, 256>>>(src, dst, copy_n);
            work = 2.0 * double(size_t(1) << 30);
            break;
        default:
            rc = Fail("probe: unknown kind");
        }
        cudaEventRecord(e1);
        if (rc == 0 && cudaEventSynchronize(e1) != cudaSuccess)
            rc = Fail("probe: kernel failed");
        float ms = 0.f;
        if (rc == 0) {
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0)
                best_ms = std::min<double>(best_ms, ms);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (scratch)
        cudaFree(scratch);
    if (src)
        cudaFree(src);
    if (dst)
        cudaFree(dst);
    if (rc != 0)
        return rc;
    JB_CUDA(cudaGetLastError());
    *value = work / (best_ms * 1e-3) / (kind == JB_PEAK_HBM_COPY ? 1e9 : 1e12); // GB/s or TFLOP/s
    return 0;
}
