// K2c — complex128 GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64).
//
// Replaces cblas_zgemm (reference include/jet/TensorHelpers.hpp:49-61) for compute-bound complex128
// shapes (the GBS fock-8 networks: M = 2^21, N = K = 64 per step; the square microbench).
//
//  * Complex -> real embedding, as in the tcgen05 kernel: interleaved complex A (M x K) *is* real
//    A' (M x 2K), C (M x N) *is* C' (M x 2N), and
//        B'[2k][2n] = Re B, B'[2k][2n+1] = Im B, B'[2k+1][2n] = -Im B, B'[2k+1][2n+1] = Re B.
//    B' is never materialised: a lane of the B fragment needs one real number of one complex entry
//    of the shared-memory B tile — which component, and with which sign, depends only on the lane.
//    8*M*N*K real flops, exactly the complex product; FP64 accumulation in the MMA.
//  * CTA tile 128 x 64 complex, K step 8 complex, 8 warps as 4 (M) x 2 (N): a warp owns 32 x 32
//    complex = 4 x 8 m8n8 accumulator blocks (64 doubles per thread).  Per k4 step a warp issues
//    4 + 8 shared-memory loads for 32 DMMAs; the DMMA pipe is the limit, not the LSU.
//  * Three-stage cp.async pipeline; shared-memory rows padded (A: 16 -> 20 doubles, B: 64 -> 66
//    complex) so that both fragment loads are bank-conflict-free.
//  * Ragged M and N: rows / columns beyond the matrix are zero-filled by cp.async (src-size 0) and
//    not stored.  K must be a multiple of 8.
#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace jb {
namespace {

constexpr int kDmBM = 128;
constexpr int kDmBN = 64; // complex columns
constexpr int kDmBK = 8;  // complex k per stage
constexpr int kDmThreads = 256;
constexpr int kDmStages = 3;
constexpr int kDmAPitch = 2 * kDmBK + 4; // doubles per A' row
constexpr int kDmBPitch = kDmBN + 2;     // complex per B row
constexpr int kDmAStage = kDmBM * kDmAPitch;     // doubles
constexpr int kDmBStage = kDmBK * kDmBPitch * 2; // doubles
constexpr size_t kDmSmemBytes = sizeof(double) * kDmStages * (kDmAStage + kDmBStage);

__device__ __forceinline__ void Dmma(double &c0, double &c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void CpAsync16(unsigned dst, const void *src, bool valid)
{
    const int bytes = valid ? 16 : 0; // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kDmThreads, 1)
    GemmDmmaKernel(const double2 *__restrict__ A, const double2 *__restrict__ B, double2 *__restrict__ C,
                   long long M, long long N, long long K, int tiles_n, int k_tiles_per_split)
{
    extern __shared__ __align__(16) double dm_smem[];
    double *As = dm_smem;
    double *Bs = dm_smem + kDmStages * kDmAStage;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const long long m0 = static_cast<long long>(blockIdx.x / tiles_n) * kDmBM;
    const long long n0 = static_cast<long long>(blockIdx.x % tiles_n) * kDmBN;
    // split-K: blockIdx.y owns k-tiles [kt0, kt0 + k_tiles) and writes its own partial result
    const int k_tiles_total = static_cast<int>(K / kDmBK);
    const int kt0 = blockIdx.y * k_tiles_per_split;
    const int k_tiles = min(k_tiles_per_split, k_tiles_total - kt0);
    C += static_cast<long long>(blockIdx.y) * M * N;
    const unsigned as_s = static_cast<unsigned>(__cvta_generic_to_shared(As));
    const unsigned bs_s = static_cast<unsigned>(__cvta_generic_to_shared(Bs));

    auto load_stage = [&](int kt, int s) {
        const long long k0 = static_cast<long long>(kt0 + kt) * kDmBK;
#pragma unroll
        for (int i = 0; i < (kDmBM * kDmBK) / kDmThreads; i++) { // A: 1024 complex, 4 per thread
            const int q = i * kDmThreads + tid;
            const int row = q / kDmBK, kc = q % kDmBK;
            const bool ok = m0 + row < M;
            const double2 *src = A + (ok ? (m0 + row) * K + k0 + kc : 0);
            CpAsync16(as_s + static_cast<unsigned>(sizeof(double)) * (s * kDmAStage + row * kDmAPitch + kc * 2), src, ok);
        }
#pragma unroll
        for (int i = 0; i < (kDmBK * kDmBN) / kDmThreads; i++) { // B: 512 complex, 2 per thread
            const int q = i * kDmThreads + tid;
            const int kr = q / kDmBN, n = q % kDmBN;
            const bool ok = n0 + n < N;
            const double2 *src = B + (ok ? (k0 + kr) * N + n0 + n : 0);
            CpAsync16(bs_s + static_cast<unsigned>(sizeof(double)) * (s * kDmBStage + (kr * kDmBPitch + n) * 2), src, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[4][8][2];
#pragma unroll
    for (int rb = 0; rb < 4; rb++)
#pragma unroll
        for (int cb = 0; cb < 8; cb++)
            acc[rb][cb][0] = acc[rb][cb][1] = 0.0;

    // lane constants of the fragments
    const int a_row = wm * 32 + (lane >> 2); // + rb * 8
    const int a_col = lane & 3;              // + ks * 4
    const int r = lane & 1, c = (lane >> 2) & 1;
    const int b_comp = r ^ c;                               // 0: Re, 1: Im
    const double b_sign = (r == 1 && c == 0) ? -1.0 : 1.0;  // B'[2k+1][2n] = -Im
    const int b_k = (lane & 3) >> 1;                        // + ks * 2
    const int b_n = wn * 32 + (lane >> 3);                  // + cb * 4

    for (int s = 0; s < kDmStages - 1; s++) {
        if (s < k_tiles)
            load_stage(s, s);
        else
            asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < k_tiles; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kDmStages - 2) : "memory");
        __syncthreads();
        // prefetch the tile kDmStages-1 ahead into the buffer that was consumed last iteration
        if (kt + kDmStages - 1 < k_tiles)
            load_stage(kt + kDmStages - 1, (kt + kDmStages - 1) % kDmStages);
        else
            asm volatile("cp.async.commit_group;" ::: "memory");
        const double *At = As + (kt % kDmStages) * kDmAStage;
        const double *Bt = Bs + (kt % kDmStages) * kDmBStage;
#pragma unroll
        for (int ks = 0; ks < (2 * kDmBK) / 4; ks++) {
            double a[4], b[8];
#pragma unroll
            for (int rb = 0; rb < 4; rb++)
                a[rb] = At[(a_row + rb * 8) * kDmAPitch + ks * 4 + a_col];
#pragma unroll
            for (int cb = 0; cb < 8; cb++)
                b[cb] = b_sign * Bt[((ks * 2 + b_k) * kDmBPitch + b_n + cb * 4) * 2 + b_comp];
#pragma unroll
            for (int rb = 0; rb < 4; rb++)
#pragma unroll
                for (int cb = 0; cb < 8; cb++)
                    Dmma(acc[rb][cb][0], acc[rb][cb][1], a[rb], b[cb]);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    // C fragment: row lane/4, real columns 2*(lane%4) + {0,1} = the complex column lane%4
#pragma unroll
    for (int rb = 0; rb < 4; rb++) {
        const long long m = m0 + wm * 32 + rb * 8 + (lane >> 2);
        if (m >= M)
            continue;
#pragma unroll
        for (int cb = 0; cb < 8; cb++) {
            const long long n = n0 + wn * 32 + cb * 4 + (lane & 3);
            if (n < N)
                C[m * N + n] = double2{acc[rb][cb][0], acc[rb][cb][1]};
        }
    }
}

} // namespace

namespace {

// sum of the split-K partials in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
    DmmaSplitReduceKernel(const double2 *__restrict__ partial, double2 *__restrict__ out, long long mn, int splits)
{
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < mn; i += step) {
        double re = 0.0, im = 0.0;
        for (int z = 0; z < splits; z++) {
            const double2 v = partial[static_cast<long long>(z) * mn + i];
            re += v.x;
            im += v.y;
        }
        out[i] = double2{re, im};
    }
}

struct DmmaShape {
    long long tiles;
    int tiles_n, splits, k_tiles_per_split;
};

DmmaShape DmmaChoose(int64_t m, int64_t n, int64_t k)
{
    DmmaShape t;
    t.tiles_n = static_cast<int>((n + kDmBN - 1) / kDmBN);
    t.tiles = ((m + kDmBM - 1) / kDmBM) * t.tiles_n;
    const long long k_tiles = k / kDmBK;
    const int sms = NumSMs();
    long long splits = 1;
    if (t.tiles < sms && k_tiles >= 64) // few output tiles and a long K: fill the machine along K
        splits = std::min<long long>({(2ll * sms + t.tiles - 1) / t.tiles, k_tiles / 16, 1024ll});
    splits = std::max<long long>(splits, 1);
    const long long per = (k_tiles + splits - 1) / splits;
    t.k_tiles_per_split = static_cast<int>(per);
    t.splits = static_cast<int>((k_tiles + per - 1) / per);
    return t;
}

} // namespace

bool GemmDmmaEligible(int dtype, int64_t m, int64_t n, int64_t k)
{
    if (dtype != JB_C128 || k % kDmBK != 0 || k < kDmBK || m < 32 || n < 16)
        return false;
    const long long tiles = ((m + kDmBM - 1) / kDmBM) * ((n + kDmBN - 1) / kDmBN);
    if (tiles >= (1ll << 31) || k / kDmBK >= (1ll << 31))
        return false;
    return static_cast<double>(m) * n * k >= double(1 << 18);
}

size_t GemmDmmaWorkspaceBytes(int64_t m, int64_t n, int64_t k)
{
    const DmmaShape t = DmmaChoose(m, n, k);
    return t.splits > 1 ? sizeof(double2) * static_cast<size_t>(t.splits) * m * n : 0;
}

int LaunchGemmDmma(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
                   cudaStream_t stream)
{
    JB_REQUIRE(GemmDmmaEligible(JB_C128, m, n, k), "gemm: shape not eligible for the FP64 tensor-core kernel");
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(GemmDmmaKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kDmSmemBytes));
    });
    JB_CUDA(attr_err);
    const DmmaShape t = DmmaChoose(m, n, k);
    double2 *dst = static_cast<double2 *>(c);
    if (t.splits > 1) {
        JB_REQUIRE(ws != nullptr && ws_bytes >= GemmDmmaWorkspaceBytes(m, n, k), "gemm: split-K workspace too small");
        dst = static_cast<double2 *>(ws);
    }
    dim3 grid(static_cast<unsigned>(t.tiles), static_cast<unsigned>(t.splits), 1);
    GemmDmmaKernel<<<grid, kDmThreads, kDmSmemBytes, stream>>>(static_cast<const double2 *>(a),
                                                               static_cast<const double2 *>(b), dst, m, n, k, t.tiles_n,
                                                               t.k_tiles_per_split);
    JB_CUDA(cudaGetLastError());
    if (t.splits > 1) {
        const long long mn = m * n;
        const int rgrid = static_cast<int>(std::min<long long>((mn + 255) / 256, NumSMs() * 8ll));
        DmmaSplitReduceKernel<<<rgrid, 256, 0, stream>>>(static_cast<const double2 *>(ws), static_cast<double2 *>(c), mn,
                                                        t.splits);
        JB_CUDA(cudaGetLastError());
    }
    return 0;
}

} // namespace jb
