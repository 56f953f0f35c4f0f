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
//  * CTA tile 128 x 32 complex (template BN; 128 x 64 kept for comparison), K step 8 complex, 8 warps as
//    4 (M) x 2 (N): a warp owns 32 x 16 complex = 4 x 4 m8n8 accumulator blocks (32 doubles per thread,
//    118 registers: two CTAs per SM).  Per k4 step a warp issues 4 + 4 shared-memory loads for 16 DMMAs;
//    the DMMA pipe is the limit, not the LSU.
//  * Three-stage cp.async pipeline; shared-memory rows padded (A: 16 -> 20 doubles, B: 64 -> 66
//    complex) so that both fragment loads are bank-conflict-free.
//  * Ragged M and N: rows / columns beyond the matrix are zero-filled by cp.async (src-size 0) and
//    not stored.  K must be a multiple of 8.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace jb {
namespace {

constexpr int kDmBM = 128;
constexpr int kDmBN = 64; // complex columns of the wide tile (the narrow one is 32: two CTAs per SM)
constexpr int kDmBK = 8;  // complex k per stage
constexpr int kDmThreads = 256;
constexpr int kDmStages = 3;
constexpr int kDmAPitch = 2 * kDmBK + 4; // doubles per A' row
constexpr int kDmAStage = kDmBM * kDmAPitch;     // doubles
template <int BN> struct DmCfg {
    static constexpr int kBPitch = BN + 2;                // complex per B row
    static constexpr int kBStage = kDmBK * kBPitch * 2;   // doubles
    static constexpr size_t kSmemBytes = sizeof(double) * kDmStages * (kDmAStage + kBStage);
    static constexpr int kMinCtas = BN == 64 ? 1 : 2;
};

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

// GATHER: A is read in its ORIGINAL tensor layout — the Transpose(A -> left ++ common) of
// Tensor::ContractTensors (reference include/jet/Tensor.hpp:743-746) is folded into the tile loads.
// All extents are powers of two: logical row bit j / k bit j of the GEMM is address bit m_bit[j] /
// k_bit[j] of the tensor (both ascending: dropping the contracted bits from A's address leaves the
// free indices in row-major order).  The 1024 elements of a tile are walked in ascending ADDRESS order
// (tile_pos / tile_kind / tile_idx: the 7 row bits and 3 k bits of the tile, merged by address), so the
// 16-byte cp.async of adjacent lanes touch adjacent memory wherever the tensor layout allows.
struct DmmaGather {
    int log_m, log_k;
    unsigned char m_bit[40], k_bit[32];
    unsigned char tile_pos[10], tile_kind[10], tile_idx[10]; // kind 0: row bit tile_idx, 1: k bit tile_idx
    int tile_bits;
};

// WM = warps along M: 4 (a warp owns 32 rows x BN/2 columns) or, for M <= 64, 2 (32 rows x BN/4 columns:
// no warp works on rows that do not exist)
template <bool GATHER, int BN, int WM>
__global__ void __launch_bounds__(kDmThreads, DmCfg<BN>::kMinCtas)
    GemmDmmaKernel(const double2 *__restrict__ A, const double2 *__restrict__ B, double2 *__restrict__ C,
                   long long M, long long N, long long K, int tiles_n, int k_tiles_per_split,
                   const __grid_constant__ DmmaGather ga)
{
    constexpr int kDmBPitch = DmCfg<BN>::kBPitch;
    constexpr int kDmBStage = DmCfg<BN>::kBStage;
    constexpr int WN = 8 / WM;             // warps along N
    constexpr int WTN = BN / WN;           // complex columns per warp
    constexpr int CBN = WTN / 4;           // m8n8 column blocks per warp
    static_assert(WM * WN == 8 && CBN >= 1, "warp layout");
    extern __shared__ __align__(16) double dm_smem[];
    double *As = dm_smem;
    double *Bs = dm_smem + kDmStages * kDmAStage;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const long long m0 = static_cast<long long>(blockIdx.x / tiles_n) * kDmBM;
    const long long n0 = static_cast<long long>(blockIdx.x % tiles_n) * BN;
    // split-K: blockIdx.y owns k-tiles [kt0, kt0 + k_tiles) and writes its own partial result
    const int k_tiles_total = static_cast<int>(K / kDmBK);
    const int kt0 = blockIdx.y * k_tiles_per_split;
    const int k_tiles = min(k_tiles_per_split, k_tiles_total - kt0);
    C += static_cast<long long>(blockIdx.y) * M * N;
    const unsigned as_s = static_cast<unsigned>(__cvta_generic_to_shared(As));
    const unsigned bs_s = static_cast<unsigned>(__cvta_generic_to_shared(Bs));

    // GATHER: per-thread constants of its four tile elements, and the CTA's row base
    long long g_off[4] = {0, 0, 0, 0};
    unsigned g_dst[4] = {0, 0, 0, 0};
    bool g_ok[4] = {false, false, false, false};
    long long g_rowbase = 0;
    if constexpr (GATHER) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = i * kDmThreads + tid;
            int row = 0, kc = 0;
            long long off = 0;
            for (int j = 0; j < ga.tile_bits; j++)
                if ((q >> j) & 1) {
                    off |= 1ll << ga.tile_pos[j];
                    if (ga.tile_kind[j] == 0)
                        row |= 1 << ga.tile_idx[j];
                    else
                        kc |= 1 << ga.tile_idx[j];
                }
            g_ok[i] = q < (1 << ga.tile_bits) && m0 + row < M;
            g_off[i] = off;
            g_dst[i] = static_cast<unsigned>(sizeof(double)) * (row * kDmAPitch + kc * 2);
        }
        const long long mt = m0 >> 7;
        for (int j = 7; j < ga.log_m; j++)
            if ((mt >> (j - 7)) & 1)
                g_rowbase |= 1ll << ga.m_bit[j];
    }

    auto load_stage = [&](int kt, int s) {
        const long long k0 = static_cast<long long>(kt0 + kt) * kDmBK;
        if constexpr (GATHER) {
            long long kbase = g_rowbase;
            const long long kq = k0 >> 3;
            for (int j = 3; j < ga.log_k; j++)
                if ((kq >> (j - 3)) & 1)
                    kbase |= 1ll << ga.k_bit[j];
            // (tile elements beyond 2^tile_bits do not exist when M < 128: those shared-memory rows stay
            // unwritten; they only feed output rows >= M, which are not stored)
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (g_ok[i])
                    CpAsync16(as_s + static_cast<unsigned>(sizeof(double)) * (s * kDmAStage) + g_dst[i],
                              A + kbase + g_off[i], true);
        }
        else {
#pragma unroll
            for (int i = 0; i < (kDmBM * kDmBK) / kDmThreads; i++) { // A: 1024 complex, 4 per thread
                const int q = i * kDmThreads + tid;
                const int row = q / kDmBK, kc = q % kDmBK;
                const bool ok = m0 + row < M;
                const double2 *src = A + (ok ? (m0 + row) * K + k0 + kc : 0);
                CpAsync16(as_s + static_cast<unsigned>(sizeof(double)) * (s * kDmAStage + row * kDmAPitch + kc * 2), src, ok);
            }
        }
#pragma unroll
        for (int i = 0; i < (kDmBK * BN) / kDmThreads; i++) { // B: 8 x BN complex, 2 or 1 per thread
            const int q = i * kDmThreads + tid;
            const int kr = q / BN, n = q % BN;
            const bool ok = n0 + n < N;
            const double2 *src = B + (ok ? (k0 + kr) * N + n0 + n : 0);
            CpAsync16(bs_s + static_cast<unsigned>(sizeof(double)) * (s * kDmBStage + (kr * kDmBPitch + n) * 2), src, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[4][CBN][2];
#pragma unroll
    for (int rb = 0; rb < 4; rb++)
#pragma unroll
        for (int cb = 0; cb < CBN; cb++)
            acc[rb][cb][0] = acc[rb][cb][1] = 0.0;

    // lane constants of the fragments
    const int a_row = wm * 32 + (lane >> 2); // + rb * 8
    const int a_col = lane & 3;              // + ks * 4
    const int r = lane & 1, c = (lane >> 2) & 1;
    const int b_comp = r ^ c;                               // 0: Re, 1: Im
    const double b_sign = (r == 1 && c == 0) ? -1.0 : 1.0;  // B'[2k+1][2n] = -Im
    const int b_k = (lane & 3) >> 1;                        // + ks * 2
    const int b_n = wn * WTN + (lane >> 3);                 // + cb * 4

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
            double a[4], b[CBN];
#pragma unroll
            for (int rb = 0; rb < 4; rb++)
                a[rb] = At[(a_row + rb * 8) * kDmAPitch + ks * 4 + a_col];
#pragma unroll
            for (int cb = 0; cb < CBN; cb++)
                b[cb] = b_sign * Bt[((ks * 2 + b_k) * kDmBPitch + b_n + cb * 4) * 2 + b_comp];
#pragma unroll
            for (int rb = 0; rb < 4; rb++)
#pragma unroll
                for (int cb = 0; cb < CBN; cb++)
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
        for (int cb = 0; cb < CBN; cb++) {
            const long long n = n0 + wn * WTN + cb * 4 + (lane & 3);
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
    int tiles_n, splits, k_tiles_per_split, bn;
};

// 128 x 32 tiles: 118 registers -> two CTAs per SM, so one CTA's pipeline fill, barriers and epilogue overlap the
// other's MMAs.  Measured faster than 128 x 64 (190 registers, one CTA per SM) on every shape tried: 2^21 x 64 x 64
// 29.2 vs 25.4 TFLOP/s, 4096^3 30.8 vs 29.0, 8192 x 64 x 1024 23.6 vs 19.4.  JB_DMMA_BN=64 selects the wide tile.
int DmmaTileN(int64_t /*m*/, int64_t /*n*/, int64_t /*k*/)
{
    static const int forced = [] {
        const char *e = getenv("JB_DMMA_BN");
        return e ? atoi(e) : 0;
    }();
    if (forced == 32 || forced == 64)
        return forced;
    return 32;
}

DmmaShape DmmaChoose(int64_t m, int64_t n, int64_t k)
{
    DmmaShape t;
    t.bn = DmmaTileN(m, n, k);
    t.tiles_n = static_cast<int>((n + t.bn - 1) / t.bn);
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

namespace {
int LaunchDmma(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
               const DmmaGather *ga, cudaStream_t stream);
}

int LaunchGemmDmma(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
                   cudaStream_t stream)
{
    return LaunchDmma(m, n, k, a, b, c, ws, ws_bytes, nullptr, stream);
}

// A in its original tensor layout: free_bits / common_bits = address bits (element units, ascending) of
// the free and of the contracted index bits of A
int LaunchGemmDmmaGatherA(int64_t m, int64_t n, int64_t k, const void *a_tensor, const int *free_bits, int n_free,
                          const int *common_bits, int n_common, const void *b, void *c, void *ws, size_t ws_bytes,
                          cudaStream_t stream)
{
    JB_REQUIRE((int64_t(1) << n_free) == m && (int64_t(1) << n_common) == k, "gemm: gather layout does not match M, K");
    JB_REQUIRE(n_free <= 40 && n_common <= 32 && n_common >= 3, "gemm: gather layout out of range");
    DmmaGather ga;
    std::memset(&ga, 0, sizeof(ga));
    ga.log_m = n_free;
    ga.log_k = n_common;
    for (int j = 0; j < n_free; j++)
        ga.m_bit[j] = static_cast<unsigned char>(free_bits[j]);
    for (int j = 0; j < n_common; j++)
        ga.k_bit[j] = static_cast<unsigned char>(common_bits[j]);
    // tile bits (<= 7 row bits, 3 k bits) merged in ascending address order
    struct TB {
        int pos, kind, idx;
    };
    std::vector<TB> tb;
    for (int j = 0; j < std::min(n_free, 7); j++)
        tb.push_back({free_bits[j], 0, j});
    for (int j = 0; j < 3; j++)
        tb.push_back({common_bits[j], 1, j});
    std::sort(tb.begin(), tb.end(), [](const TB &x, const TB &y) { return x.pos < y.pos; });
    ga.tile_bits = static_cast<int>(tb.size());
    for (size_t j = 0; j < tb.size(); j++) {
        ga.tile_pos[j] = static_cast<unsigned char>(tb[j].pos);
        ga.tile_kind[j] = static_cast<unsigned char>(tb[j].kind);
        ga.tile_idx[j] = static_cast<unsigned char>(tb[j].idx);
    }
    return LaunchDmma(m, n, k, a_tensor, b, c, ws, ws_bytes, &ga, stream);
}

namespace {
int LaunchDmma(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
               const DmmaGather *gather, cudaStream_t stream)
{
    JB_REQUIRE(GemmDmmaEligible(JB_C128, m, n, k), "gemm: shape not eligible for the FP64 tensor-core kernel");
    {
        const std::pair<const void *, size_t> kernels[] = {
            {reinterpret_cast<const void *>(GemmDmmaKernel<false, 64, 4>), DmCfg<64>::kSmemBytes},
            {reinterpret_cast<const void *>(GemmDmmaKernel<true, 64, 4>), DmCfg<64>::kSmemBytes},
            {reinterpret_cast<const void *>(GemmDmmaKernel<false, 32, 4>), DmCfg<32>::kSmemBytes},
            {reinterpret_cast<const void *>(GemmDmmaKernel<true, 32, 4>), DmCfg<32>::kSmemBytes},
            {reinterpret_cast<const void *>(GemmDmmaKernel<false, 32, 2>), DmCfg<32>::kSmemBytes},
            {reinterpret_cast<const void *>(GemmDmmaKernel<true, 32, 2>), DmCfg<32>::kSmemBytes}};
        for (const auto &kb : kernels)
            JB_TRY(EnsureDynamicSmem(kb.first, kb.second));
    }
    const DmmaShape t = DmmaChoose(m, n, k);
    double2 *dst = static_cast<double2 *>(c);
    if (t.splits > 1) {
        JB_REQUIRE(ws != nullptr && ws_bytes >= GemmDmmaWorkspaceBytes(m, n, k), "gemm: split-K workspace too small");
        dst = static_cast<double2 *>(ws);
    }
    dim3 grid(static_cast<unsigned>(t.tiles), static_cast<unsigned>(t.splits), 1);
    DmmaGather none;
    std::memset(&none, 0, sizeof(none));
    const double2 *pa = static_cast<const double2 *>(a), *pb = static_cast<const double2 *>(b);
    if (t.bn == 32 && m <= 64) { // short M: two warps along M, four along N
        if (gather != nullptr)
            GemmDmmaKernel<true, 32, 2><<<grid, kDmThreads, DmCfg<32>::kSmemBytes, stream>>>(pa, pb, dst, m, n, k, t.tiles_n,
                                                                                             t.k_tiles_per_split, *gather);
        else
            GemmDmmaKernel<false, 32, 2><<<grid, kDmThreads, DmCfg<32>::kSmemBytes, stream>>>(pa, pb, dst, m, n, k, t.tiles_n,
                                                                                              t.k_tiles_per_split, none);
    }
    else if (t.bn == 64) {
        if (gather != nullptr)
            GemmDmmaKernel<true, 64, 4><<<grid, kDmThreads, DmCfg<64>::kSmemBytes, stream>>>(pa, pb, dst, m, n, k, t.tiles_n,
                                                                                          t.k_tiles_per_split, *gather);
        else
            GemmDmmaKernel<false, 64, 4><<<grid, kDmThreads, DmCfg<64>::kSmemBytes, stream>>>(pa, pb, dst, m, n, k, t.tiles_n,
                                                                                           t.k_tiles_per_split, none);
    }
    else {
        if (gather != nullptr)
            GemmDmmaKernel<true, 32, 4><<<grid, kDmThreads, DmCfg<32>::kSmemBytes, stream>>>(pa, pb, dst, m, n, k, t.tiles_n,
                                                                                          t.k_tiles_per_split, *gather);
        else
            GemmDmmaKernel<false, 32, 4><<<grid, kDmThreads, DmCfg<32>::kSmemBytes, stream>>>(pa, pb, dst, m, n, k, t.tiles_n,
                                                                                           t.k_tiles_per_split, none);
    }
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
} // namespace

} // namespace jb
