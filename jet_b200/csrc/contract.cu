// K2 — pairwise contraction kernels and their planner.
//
// Replaces Tensor::ContractTensors (reference: include/jet/Tensor.hpp:709-752) =
// Transpose(A) + Transpose(B) + cblas_{c,z}gemm/gemv/dotu (include/jet/TensorHelpers.hpp:131-168).
//
// Two kernel families:
//
//  (0) StreamContractKernel — the workhorse.  In tensor-network paths one operand is almost always
//      tiny (<= 4096 elements; 99 % of the traffic of the Sycamore m10/m12 paths) and the GEMM is
//      extremely skinny (K, N <= 16, M up to 2^27).  The small ("resident") operand is gathered
//      once per CTA into shared memory as a K x Y matrix; the big ("streamed") operand is read
//      exactly once straight from its ORIGINAL layout — the index permutation of the reference's
//      Transpose() is folded into the load addresses (bit insertion for the contracted indices) —
//      and the output is written exactly once, already in the reference's (left ++ right) order.
//      HBM traffic = the algorithmic minimum sizeof(T)*(MK + KN + MN).  FP32 (FP64 for c128)
//      FMA with full-precision accumulation: bandwidth-bound, tensor cores would not help.
//
//  (1) TTGT — permute A and B with K1 into a workspace, then a dense row-major complex GEMM
//      (GemmKernel, FP32/FP64 FMA, shared-memory tiled, deterministic split-K).  Used when both
//      operands are large, for non-power-of-two extents, and for the GEMV / DOTU corners.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cooperative_groups.h>

#include "common.cuh"

namespace jb {

// =================================================================================================
// Stream kernel
// =================================================================================================
struct StreamParams {
    int log_x, log_k, log_y;
    int out_x_shift, out_y_shift;
    long long x_count;
    uint8_t cs[16]; // streamed-operand address bit of k bit q (ascending)
    uint8_t rk[16]; // resident-operand address bit of k bit q
    uint8_t ry[16]; // resident-operand address bit of y bit q (ascending)
};

namespace {

constexpr int kStreamThreads = 256;
constexpr int kStreamMaxResident = 4096;

template <typename R> struct Cx;
template <> struct Cx<float> {
    using type = float2;
};
template <> struct Cx<double> {
    using type = double2;
};

template <typename C> __device__ __forceinline__ void CFma(C &acc, const C a, const C b)
{
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ unsigned long long ScatterBits(unsigned long long x, const uint8_t *dst,
                                                          int nbits)
{
    unsigned long long r = 0;
    for (int q = 0; q < nbits; q++)
        r |= ((x >> q) & 1ull) << dst[q];
    return r;
}

// insert a zero bit at each position cs[0] < cs[1] < ... of x
__device__ __forceinline__ unsigned long long InsertZeros(unsigned long long x, const uint8_t *cs,
                                                          int c)
{
    for (int q = 0; q < c; q++) {
        const unsigned long long low = x & ((1ull << cs[q]) - 1ull);
        x = ((x ^ low) << 1) | low;
    }
    return x;
}

// KC: k values held in registers at once; NR: outputs per thread per pass; XT: x values per thread
template <typename R, int KC, int NR, int XT>
__global__ void __launch_bounds__(kStreamThreads, (KC * XT * sizeof(R) <= 32 && NR <= 4) ? 4 : ((KC * NR * XT * sizeof(R) <= 256) ? 3 : 2))
    StreamContractKernel(const typename Cx<R>::type *__restrict__ S,
                         const typename Cx<R>::type *__restrict__ Rsd,
                         typename Cx<R>::type *__restrict__ out,
                         const __grid_constant__ StreamParams p, const long long stride_s, const long long stride_r,
                         const long long stride_o)
{
    using C = typename Cx<R>::type;
    // slice batching: blockIdx.y = slice within the batch; its tensors lie stride_* BYTES further on
    S = reinterpret_cast<const C *>(reinterpret_cast<const unsigned char *>(S) + blockIdx.y * stride_s);
    Rsd = reinterpret_cast<const C *>(reinterpret_cast<const unsigned char *>(Rsd) + blockIdx.y * stride_r);
    out = reinterpret_cast<C *>(reinterpret_cast<unsigned char *>(out) + blockIdx.y * stride_o);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *Rm = reinterpret_cast<C *>(smem_raw); // [K][Y]

    const int tid = threadIdx.x;
    const int K = 1 << p.log_k;
    const int Y = 1 << p.log_y;

    // resident operand -> shared memory, as the K x Y matrix the streamed operand expects
    for (int e = tid; e < K * Y; e += kStreamThreads) {
        const unsigned k = e >> p.log_y;
        const unsigned y = e & (Y - 1);
        const unsigned long long addr = ScatterBits(k, p.rk, p.log_k) | ScatterBits(y, p.ry, p.log_y);
        Rm[e] = __ldg(Rsd + addr);
    }

    // address offset of the low log2(KC) bits of k (loop invariant)
    unsigned long long koff_lo[KC];
#pragma unroll
    for (int kk = 0; kk < KC; kk++) {
        unsigned long long r = 0;
#pragma unroll
        for (int q = 0; (1 << q) < KC; q++)
            if (kk & (1 << q))
                r |= 1ull << p.cs[q];
        koff_lo[kk] = r;
    }
    constexpr int kLogKC = (KC == 1) ? 0 : (KC == 2) ? 1 : (KC == 4) ? 2 : (KC == 8) ? 3 : 4;
    const int n_kchunks = K / KC;
    __syncthreads();

    const long long tile = static_cast<long long>(kStreamThreads) * XT;
    for (long long x0 = static_cast<long long>(blockIdx.x) * tile; x0 < p.x_count;
         x0 += static_cast<long long>(gridDim.x) * tile) {
        long long x[XT];
        unsigned long long sbase[XT];
        bool ok[XT];
#pragma unroll
        for (int j = 0; j < XT; j++) {
            x[j] = x0 + j * kStreamThreads + tid;
            ok[j] = x[j] < p.x_count;
            sbase[j] = InsertZeros(static_cast<unsigned long long>(ok[j] ? x[j] : 0), p.cs, p.log_k);
        }
        C a[XT][KC];
        for (int y0 = 0; y0 < Y; y0 += NR) {
            C acc[XT][NR];
#pragma unroll
            for (int j = 0; j < XT; j++)
#pragma unroll
                for (int yy = 0; yy < NR; yy++)
                    acc[j][yy] = C{R(0), R(0)};
            for (int kc = 0; kc < n_kchunks; kc++) {
                if (n_kchunks > 1 || y0 == 0) {
                    const unsigned long long khi =
                        ScatterBits(static_cast<unsigned long long>(kc), p.cs + kLogKC,
                                    p.log_k - kLogKC);
#pragma unroll
                    for (int j = 0; j < XT; j++)
#pragma unroll
                        for (int kk = 0; kk < KC; kk++)
                            a[j][kk] = ok[j] ? __ldg(S + (sbase[j] | khi | koff_lo[kk]))
                                             : C{R(0), R(0)};
                }
                const C *rrow = Rm + (kc * KC) * Y + y0;
#pragma unroll
                for (int kk = 0; kk < KC; kk++) {
#pragma unroll
                    for (int yy = 0; yy < NR; yy++) {
                        const C r = rrow[kk * Y + yy];
#pragma unroll
                        for (int j = 0; j < XT; j++)
                            CFma(acc[j][yy], a[j][kk], r);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < XT; j++) {
                if (!ok[j])
                    continue;
                if (p.out_y_shift == 0) {
                    C *dst = out + ((static_cast<unsigned long long>(x[j]) << p.out_x_shift) + y0);
                    if constexpr (sizeof(C) == 8 && NR >= 2) {
#pragma unroll
                        for (int yy = 0; yy < NR; yy += 2) {
                            float4 v = make_float4(acc[j][yy].x, acc[j][yy].y, acc[j][yy + 1].x,
                                                   acc[j][yy + 1].y);
                            *reinterpret_cast<float4 *>(dst + yy) = v;
                        }
                    }
                    else {
#pragma unroll
                        for (int yy = 0; yy < NR; yy++)
                            dst[yy] = acc[j][yy];
                    }
                }
                else {
#pragma unroll
                    for (int yy = 0; yy < NR; yy++)
                        out[(static_cast<unsigned long long>(y0 + yy) << p.out_y_shift) +
                            static_cast<unsigned long long>(x[j])] = acc[j][yy];
                }
            }
        }
    }
}


template <typename R, int KC, int NR>
int LaunchStreamT(const StreamParams &p, const void *s, const void *r, void *out,
                  cudaStream_t stream, int batch, long long stride_s, long long stride_r, long long stride_o)
{
    using C = typename Cx<R>::type;
    constexpr int XT = sizeof(R) == 4 ? 2 : 1;
    const long long tile = static_cast<long long>(kStreamThreads) * XT;
    const long long tiles = (p.x_count + tile - 1) / tile;
    const size_t smem = sizeof(C) << (p.log_k + p.log_y);
    auto kernel = StreamContractKernel<R, KC, NR, XT>;
    if (smem > 48 * 1024) {
        JB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    }
    const long long resident =
        static_cast<long long>(NumSMs()) * PersistentBlocksPerSM(kernel, kStreamThreads, smem);
    // a batch of slices shares the resident CTAs: each slice gets its share, at least one CTA
    const int grid = static_cast<int>(std::min<long long>(tiles, std::max<long long>(1, resident / batch)));
    kernel<<<dim3(grid, batch), kStreamThreads, smem, stream>>>(static_cast<const C *>(s), static_cast<const C *>(r),
                                                                static_cast<C *>(out), p, stride_s, stride_r, stride_o);
    JB_CUDA(cudaGetLastError());
    return 0;
}

template <typename R, int KC>
int LaunchStreamK(const StreamParams &p, const void *s, const void *r, void *out,
                  cudaStream_t stream, int batch, long long ss, long long sr, long long so)
{
    const int y = 1 << p.log_y;
    if (y >= 8)
        return LaunchStreamT<R, KC, 8>(p, s, r, out, stream, batch, ss, sr, so);
    if (y == 4)
        return LaunchStreamT<R, KC, 4>(p, s, r, out, stream, batch, ss, sr, so);
    if (y == 2)
        return LaunchStreamT<R, KC, 2>(p, s, r, out, stream, batch, ss, sr, so);
    return LaunchStreamT<R, KC, 1>(p, s, r, out, stream, batch, ss, sr, so);
}

template <typename R>
int LaunchStream(const StreamParams &p, const void *s, const void *r, void *out,
                 cudaStream_t stream, int batch = 1, long long ss = 0, long long sr = 0, long long so = 0)
{
    JB_REQUIRE(batch >= 1 && batch <= 65535, "contract: batch out of range");
    const int k = 1 << p.log_k;
    if (k >= 8)
        return LaunchStreamK<R, 8>(p, s, r, out, stream, batch, ss, sr, so);
    if (k == 4)
        return LaunchStreamK<R, 4>(p, s, r, out, stream, batch, ss, sr, so);
    if (k == 2)
        return LaunchStreamK<R, 2>(p, s, r, out, stream, batch, ss, sr, so);
    return LaunchStreamK<R, 1>(p, s, r, out, stream, batch, ss, sr, so);
}

// =================================================================================================
// Dense row-major complex GEMM (alpha = 1, beta = 0), optional deterministic split-K
// =================================================================================================
template <typename R, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
    GemmKernel(const typename Cx<R>::type *__restrict__ A, const typename Cx<R>::type *__restrict__ B,
               typename Cx<R>::type *__restrict__ Cout, long long M, long long N, long long K,
               long long k_per_split)
{
    using C = typename Cx<R>::type;
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ C As[BK][BM + 1];
    __shared__ C Bs[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);
    const long long tiles_n = (N + BN - 1) / BN;
    const long long m0 = (static_cast<long long>(blockIdx.x) / tiles_n) * BM;
    const long long n0 = (static_cast<long long>(blockIdx.x) % tiles_n) * BN;
    const long long kb = static_cast<long long>(blockIdx.y) * k_per_split;
    const long long ke = min(K, kb + k_per_split);

    C acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++)
            acc[i][j] = C{R(0), R(0)};

    for (long long k0 = kb; k0 < ke; k0 += BK) {
        // A tile: BM x BK, contiguous along k
        for (int e = tid; e < BM * BK; e += NT) {
            const int mm = e / BK, kk = e % BK;
            const long long m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < ke) ? __ldg(A + m * K + k) : C{R(0), R(0)};
        }
        // B tile: BK x BN, contiguous along n
        for (int e = tid; e < BK * BN; e += NT) {
            const int kk = e / BN, nn = e % BN;
            const long long n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < ke) ? __ldg(B + k * N + n) : C{R(0), R(0)};
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            C a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++)
                a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; j++)
                b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++)
                    CFma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }
    C *dst = Cout + static_cast<long long>(blockIdx.y) * M * N;
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const long long m = m0 + ty * TM + i;
        if (m >= M)
            continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const long long n = n0 + tx * TN + j;
            if (n < N)
                dst[m * N + n] = acc[i][j];
        }
    }
}

// sum the split-K partials in a fixed order (deterministic), accumulating in double
template <typename R>
__global__ void __launch_bounds__(256)
    SplitKReduceKernel(const typename Cx<R>::type *__restrict__ partial,
                       typename Cx<R>::type *__restrict__ out, long long mn, int splits)
{
    using C = typename Cx<R>::type;
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < mn;
         i += step) {
        double re = 0.0, im = 0.0;
        for (int z = 0; z < splits; z++) {
            const C v = partial[static_cast<long long>(z) * mn + i];
            re += static_cast<double>(v.x);
            im += static_cast<double>(v.y);
        }
        out[i] = C{static_cast<R>(re), static_cast<R>(im)};
    }
}

// C(M x N) = A(M x K) B(K x N) with M, N in {1, 2, 4} and a long K: the DOTU / GEMV corner of
// MultiplyTensorData (reference include/jet/TensorHelpers.hpp:79-111; the last step of every closed
// network is a DOTU over the whole remaining tensor).  HBM-bound: both operands are read once,
// coalesced; partial sums are kept in double and combined in a fixed order (deterministic).
template <typename R, int M, int N>
__global__ void __launch_bounds__(256)
    SmallMnKernel(const typename Cx<R>::type *__restrict__ A, const typename Cx<R>::type *__restrict__ B,
                  double2 *__restrict__ partial, long long K)
{
    using C = typename Cx<R>::type;
    double2 acc[M][N];
#pragma unroll
    for (int m = 0; m < M; m++)
#pragma unroll
        for (int n = 0; n < N; n++)
            acc[m][n] = double2{0.0, 0.0};
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
#pragma unroll 4
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < K; k += step) {
        C a[M], b[N];
#pragma unroll
        for (int m = 0; m < M; m++)
            a[m] = A[m * K + k];
#pragma unroll
        for (int n = 0; n < N; n++)
            b[n] = B[k * N + n];
#pragma unroll
        for (int m = 0; m < M; m++)
#pragma unroll
            for (int n = 0; n < N; n++) {
                const double ar = a[m].x, ai = a[m].y, br = b[n].x, bi = b[n].y;
                acc[m][n].x += ar * br - ai * bi;
                acc[m][n].y += ar * bi + ai * br;
            }
    }
    __shared__ double2 red[8][M * N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < M; m++)
#pragma unroll
        for (int n = 0; n < N; n++) {
            double x = acc[m][n].x, y = acc[m][n].y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x += __shfl_down_sync(0xffffffffu, x, o);
                y += __shfl_down_sync(0xffffffffu, y, o);
            }
            if (lane == 0)
                red[warp][m * N + n] = double2{x, y};
        }
    __syncthreads();
    if (threadIdx.x < M * N) {
        double x = 0.0, y = 0.0;
        for (int w = 0; w < 8; w++) {
            x += red[w][threadIdx.x].x;
            y += red[w][threadIdx.x].y;
        }
        partial[static_cast<long long>(blockIdx.x) * (M * N) + threadIdx.x] = double2{x, y};
    }
}

template <typename R>
__global__ void __launch_bounds__(32)
    SmallMnFinishKernel(const double2 *__restrict__ partial, typename Cx<R>::type *__restrict__ out, int mn, int blocks)
{
    using C = typename Cx<R>::type;
    if (static_cast<int>(threadIdx.x) < mn) {
        double x = 0.0, y = 0.0;
        for (int b = 0; b < blocks; b++) {
            x += partial[static_cast<long long>(b) * mn + threadIdx.x].x;
            y += partial[static_cast<long long>(b) * mn + threadIdx.x].y;
        }
        out[threadIdx.x] = C{static_cast<R>(x), static_cast<R>(y)};
    }
}


// ---- the same corner with both operands in their ORIGINAL tensor layouts -------------------------------------------
// The last step of a closed network contracts two large tensors over (almost) all their indices (m=20: two rank-31
// tensors, K = 2^29, M = N = 4).  Through TTGT that is two full permutations before the dot: three passes over
// memory instead of one.  Here nothing is permuted: the contracted index bits are split into TILE bits — the lowest
// contracted address bits of A and the lowest ones of B — and OUTER bits.  A CTA walks one tile at a time in A's
// address order (coalesced); the matching B elements of the tile are a bit-permutation of the same tile, spread
// over a few KB that the walk touches completely, so B's sectors are fetched from DRAM once and served from L1
// until the tile is done.  Partial sums in double, combined in a fixed order (deterministic), as above.
constexpr int kDotMaxTileBits = 11;
constexpr int kDotMaxOuterBits = 56;
struct DotGatherParams {
    int log_tile, log_outer, log_m, log_n;
    uint8_t tile_a[kDotMaxTileBits], tile_b[kDotMaxTileBits];   // address bit in A / B of tile bit q (bits 0..4 = lanes)
    uint8_t outer_a[kDotMaxOuterBits], outer_b[kDotMaxOuterBits]; // ... of outer bit q
    uint8_t m_a[2], n_b[2];                                       // address bit in A of m bit q / in B of n bit q
};

template <typename R, int M, int N>
__global__ void __launch_bounds__(256)
    DotGatherKernel(const typename Cx<R>::type *__restrict__ A, const typename Cx<R>::type *__restrict__ B,
                    double2 *__restrict__ partial, const __grid_constant__ DotGatherParams p)
{
    using C = typename Cx<R>::type;
    double2 acc[M][N];
#pragma unroll
    for (int m = 0; m < M; m++)
#pragma unroll
        for (int n = 0; n < N; n++)
            acc[m][n] = double2{0.0, 0.0};
    unsigned long long am[M], bn[N];
#pragma unroll
    for (int m = 0; m < M; m++)
        am[m] = ScatterBits(static_cast<unsigned long long>(m), p.m_a, p.log_m);
#pragma unroll
    for (int n = 0; n < N; n++)
        bn[n] = ScatterBits(static_cast<unsigned long long>(n), p.n_b, p.log_n);
    // tile index = thread (low 8 bits) | pass (the rest): both address contributions are additive
    const int tb = p.log_tile < 8 ? p.log_tile : 8;
    const unsigned long long ta = ScatterBits(threadIdx.x, p.tile_a, tb), tbb = ScatterBits(threadIdx.x, p.tile_b, tb);
    const int passes = p.log_tile > 8 ? 1 << (p.log_tile - 8) : 1;
    const bool active = static_cast<int>(threadIdx.x) < (1 << tb);
    const long long tiles = 1ll << p.log_outer;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const unsigned long long oa = ScatterBits(static_cast<unsigned long long>(t), p.outer_a, p.log_outer);
        const unsigned long long ob = ScatterBits(static_cast<unsigned long long>(t), p.outer_b, p.log_outer);
        for (int ps = 0; ps < passes; ps++) {
            const unsigned long long pa = oa | ta | ScatterBits(static_cast<unsigned long long>(ps), p.tile_a + 8, p.log_tile - tb);
            const unsigned long long pb = ob | tbb | ScatterBits(static_cast<unsigned long long>(ps), p.tile_b + 8, p.log_tile - tb);
            if (!active)
                continue;
            C a[M], b[N];
#pragma unroll
            for (int m = 0; m < M; m++)
                a[m] = __ldg(A + (pa | am[m]));
#pragma unroll
            for (int n = 0; n < N; n++)
                b[n] = __ldg(B + (pb | bn[n]));
#pragma unroll
            for (int m = 0; m < M; m++)
#pragma unroll
                for (int n = 0; n < N; n++) {
                    const double ar = a[m].x, ai = a[m].y, br = b[n].x, bi = b[n].y;
                    acc[m][n].x += ar * br - ai * bi;
                    acc[m][n].y += ar * bi + ai * br;
                }
        }
    }
    __shared__ double2 red[8][M * N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < M; m++)
#pragma unroll
        for (int n = 0; n < N; n++) {
            double x = acc[m][n].x, y = acc[m][n].y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x += __shfl_down_sync(0xffffffffu, x, o);
                y += __shfl_down_sync(0xffffffffu, y, o);
            }
            if (lane == 0)
                red[warp][m * N + n] = double2{x, y};
        }
    __syncthreads();
    if (threadIdx.x < M * N) {
        double x = 0.0, y = 0.0;
        for (int w = 0; w < 8; w++) {
            x += red[w][threadIdx.x].x;
            y += red[w][threadIdx.x].y;
        }
        partial[static_cast<long long>(blockIdx.x) * (M * N) + threadIdx.x] = double2{x, y};
    }
}

int DotGatherBlocks(const DotGatherParams &p)
{
    return static_cast<int>(std::max<long long>(1, std::min<long long>(1ll << p.log_outer, NumSMs() * 8ll)));
}

template <typename R, int M>
int LaunchDotGatherN(const DotGatherParams &p, const void *a, const void *b, void *ws, int blocks, cudaStream_t stream)
{
    using C = typename Cx<R>::type;
    const C *A = static_cast<const C *>(a);
    const C *B = static_cast<const C *>(b);
    double2 *P = static_cast<double2 *>(ws);
    if (p.log_n == 0)
        DotGatherKernel<R, M, 1><<<blocks, 256, 0, stream>>>(A, B, P, p);
    else if (p.log_n == 1)
        DotGatherKernel<R, M, 2><<<blocks, 256, 0, stream>>>(A, B, P, p);
    else
        DotGatherKernel<R, M, 4><<<blocks, 256, 0, stream>>>(A, B, P, p);
    JB_CUDA(cudaGetLastError());
    return 0;
}

template <typename R>
int LaunchDotGather(const DotGatherParams &p, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
                    cudaStream_t stream)
{
    const int blocks = DotGatherBlocks(p);
    const int mn = 1 << (p.log_m + p.log_n);
    JB_REQUIRE(ws != nullptr && ws_bytes >= sizeof(double2) * static_cast<size_t>(blocks) * mn, "gemm: workspace too small");
    if (p.log_m == 0)
        JB_TRY((LaunchDotGatherN<R, 1>(p, a, b, ws, blocks, stream)));
    else if (p.log_m == 1)
        JB_TRY((LaunchDotGatherN<R, 2>(p, a, b, ws, blocks, stream)));
    else
        JB_TRY((LaunchDotGatherN<R, 4>(p, a, b, ws, blocks, stream)));
    SmallMnFinishKernel<R><<<1, 32, 0, stream>>>(static_cast<const double2 *>(ws),
                                                static_cast<typename Cx<R>::type *>(c), mn, blocks);
    JB_CUDA(cudaGetLastError());
    return 0;
}

bool DotGatherEnabled()
{
    static const bool enabled = [] {
        const char *e = getenv("JB_DISABLE_DOT_GATHER");
        return !(e && e[0] == '1');
    }();
    return enabled;
}


// ---- small output, moderate K, operands in their original layouts, batched over slices --------------------------------
// The per-slice part of a small sliced network ends in contractions like 16 x 16 x 4096: through TTGT that is five
// launches (two permutations, GEMM, split-K reduce) — and with slice batching five launches PER SLICE of the batch,
// 84 % of a batched m10 replay.  Here one thread-block CLUSTER computes the whole M x N output of one slice
// (blockIdx.y = slice).  K is cut into chunks of 64, dealt round-robin to the CTAs of the cluster; a chunk is gathered
// into shared memory straight from the operands' own layouts (per-thread address offsets precomputed, the next chunk
// in flight in registers while the current one is multiplied); thread (k sub-split, m, n) sums its products of a chunk
// in the operand precision and adds the chunk to a double accumulator; sub-splits, then CTAs (through distributed
// shared memory, in rank order) are added in double.  No workspace, fixed order.
constexpr int kSmallGemmKT = 64;
constexpr int kSmallGemmMaxCluster = 8;
struct SmallGemmParams {
    int log_m, log_n, log_k;
    uint8_t m_a[4], n_b[4];   // address bit in A of m bit q / in B of n bit q (least significant first)
    uint8_t k_a[24], k_b[24]; // address bit in A / B of k bit q
};

template <typename R>
__global__ void __launch_bounds__(256)
    SmallGemmGatherKernel(const typename Cx<R>::type *__restrict__ A, const typename Cx<R>::type *__restrict__ B,
                          typename Cx<R>::type *__restrict__ Cout, const __grid_constant__ SmallGemmParams p,
                          const long long stride_a, const long long stride_b, const long long stride_c)
{
    using C = typename Cx<R>::type;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank()), csize = static_cast<int>(cluster.num_blocks());
    A = reinterpret_cast<const C *>(reinterpret_cast<const unsigned char *>(A) + blockIdx.y * stride_a);
    B = reinterpret_cast<const C *>(reinterpret_cast<const unsigned char *>(B) + blockIdx.y * stride_b);
    Cout = reinterpret_cast<C *>(reinterpret_cast<unsigned char *>(Cout) + blockIdx.y * stride_c);
    __shared__ C As[kSmallGemmKT][16 + 1];
    __shared__ C Bs[kSmallGemmKT][16];
    __shared__ double2 red[256];
    const int M = 1 << p.log_m, N = 1 << p.log_n;
    const int chunks = 1 << (p.log_k - 6);
    const int tid = threadIdx.x;
    // thread = (k sub-split, m, n): with fewer than 256 outputs the spare threads split each chunk's k range
    const int log_mn = p.log_m + p.log_n;
    const int mn = tid & ((1 << log_mn) - 1), ks = tid >> log_mn, S = 256 >> log_mn;
    const int m = mn >> p.log_n, n = mn & (N - 1);
    // gather slots of this thread: element e = tid + 256 j of a chunk is (kk, mm) = (e >> log_m, e & (M - 1))
    unsigned long long off_a[4], off_b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int e = tid + 256 * j;
        off_a[j] = ScatterBits(static_cast<unsigned long long>(e >> p.log_m), p.k_a, 6) |
                   ScatterBits(static_cast<unsigned long long>(e & (M - 1)), p.m_a, p.log_m);
        off_b[j] = ScatterBits(static_cast<unsigned long long>(e >> p.log_n), p.k_b, 6) |
                   ScatterBits(static_cast<unsigned long long>(e & (N - 1)), p.n_b, p.log_n);
    }
    C ra[4], rb[4];
    auto fetch = [&](int chunk) {
        const unsigned long long hi_a = ScatterBits(static_cast<unsigned long long>(chunk), p.k_a + 6, p.log_k - 6);
        const unsigned long long hi_b = ScatterBits(static_cast<unsigned long long>(chunk), p.k_b + 6, p.log_k - 6);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (tid + 256 * j < kSmallGemmKT * M)
                ra[j] = __ldg(A + (hi_a | off_a[j]));
            if (tid + 256 * j < kSmallGemmKT * N)
                rb[j] = __ldg(B + (hi_b | off_b[j]));
        }
    };
    double re = 0.0, im = 0.0;
    if (rank < chunks)
        fetch(rank);
    for (int chunk = rank; chunk < chunks; chunk += csize) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int e = tid + 256 * j;
            if (e < kSmallGemmKT * M)
                As[e >> p.log_m][e & (M - 1)] = ra[j];
            if (e < kSmallGemmKT * N)
                Bs[e >> p.log_n][e & (N - 1)] = rb[j];
        }
        __syncthreads();
        if (chunk + csize < chunks)
            fetch(chunk + csize);
        R cr = R(0), ci = R(0);
#pragma unroll 4
        for (int kk = ks; kk < kSmallGemmKT; kk += S) {
            const C a = As[kk][m], b = Bs[kk][n];
            cr += a.x * b.x - a.y * b.y;
            ci += a.x * b.y + a.y * b.x;
        }
        re += static_cast<double>(cr);
        im += static_cast<double>(ci);
        __syncthreads();
    }
    // sub-splits of this CTA, then the CTAs of the cluster, each in a fixed order
    red[tid] = make_double2(re, im);
    __syncthreads();
    if (ks == 0) {
        for (int s2 = 1; s2 < S; s2++) {
            re += red[(s2 << log_mn) | mn].x;
            im += red[(s2 << log_mn) | mn].y;
        }
        red[mn] = make_double2(re, im);
    }
    cluster.sync();
    if (rank == 0 && ks == 0) {
        for (int r = 1; r < csize; r++) {
            const double2 *peer = cluster.map_shared_rank(red, r);
            re += peer[mn].x;
            im += peer[mn].y;
        }
        Cout[m * N + n] = C{static_cast<R>(re), static_cast<R>(im)};
    }
    cluster.sync(); // peers' shared memory stays alive until rank 0 has read it
}

bool SmallGemmEnabled()
{
    static const bool enabled = [] {
        const char *e = getenv("JB_DISABLE_SMALL_GEMM");
        return !(e && e[0] == '1');
    }();
    return enabled;
}

template <typename R>
int LaunchSmallGemmT(const SmallGemmParams &p, const void *a, const void *b, void *c, cudaStream_t stream, int nb, long long sa,
                     long long sb, long long sc)
{
    using C = typename Cx<R>::type;
    const int chunks = 1 << (p.log_k - 6);
    const int csize = std::min(kSmallGemmMaxCluster, chunks);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, nb);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    JB_CUDA(cudaLaunchKernelEx(&cfg, SmallGemmGatherKernel<R>, static_cast<const C *>(a), static_cast<const C *>(b),
                               static_cast<C *>(c), p, sa, sb, sc));
    return 0;
}

int LaunchSmallGemm(int dtype, const SmallGemmParams &p, const void *a, const void *b, void *c, cudaStream_t stream,
                    const BatchArgs *batch)
{
    const int nb = batch ? batch->count : 1;
    JB_REQUIRE(nb >= 1 && nb <= 65535, "contract: batch out of range");
    const long long sa = batch ? batch->stride_a : 0, sb = batch ? batch->stride_b : 0, sc = batch ? batch->stride_c : 0;
    if (dtype == JB_C64)
        return LaunchSmallGemmT<float>(p, a, b, c, stream, nb, sa, sb, sc);
    return LaunchSmallGemmT<double>(p, a, b, c, stream, nb, sa, sb, sc);
}

bool SmallMnEligible(int64_t m, int64_t n, int64_t k)
{
    auto ok = [](int64_t v) { return v == 1 || v == 2 || v == 4; };
    return ok(m) && ok(n) && k >= 4096;
}

int SmallMnBlocks(int64_t k)
{
    return static_cast<int>(std::max<long long>(1, std::min<long long>((k + 1023) / 1024, NumSMs() * 8ll)));
}

size_t SmallMnWorkspaceBytes(int64_t m, int64_t n, int64_t k)
{
    return sizeof(double2) * static_cast<size_t>(SmallMnBlocks(k)) * m * n;
}

template <typename R, int M>
int LaunchSmallMnN(int64_t n, int64_t k, const void *a, const void *b, void *ws, int blocks, cudaStream_t stream)
{
    using C = typename Cx<R>::type;
    const C *A = static_cast<const C *>(a);
    const C *B = static_cast<const C *>(b);
    double2 *P = static_cast<double2 *>(ws);
    if (n == 1)
        SmallMnKernel<R, M, 1><<<blocks, 256, 0, stream>>>(A, B, P, k);
    else if (n == 2)
        SmallMnKernel<R, M, 2><<<blocks, 256, 0, stream>>>(A, B, P, k);
    else
        SmallMnKernel<R, M, 4><<<blocks, 256, 0, stream>>>(A, B, P, k);
    JB_CUDA(cudaGetLastError());
    return 0;
}

template <typename R>
int LaunchSmallMn(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
                  cudaStream_t stream)
{
    JB_REQUIRE(ws != nullptr && ws_bytes >= SmallMnWorkspaceBytes(m, n, k), "gemm: workspace too small");
    const int blocks = SmallMnBlocks(k);
    if (m == 1)
        JB_TRY((LaunchSmallMnN<R, 1>(n, k, a, b, ws, blocks, stream)));
    else if (m == 2)
        JB_TRY((LaunchSmallMnN<R, 2>(n, k, a, b, ws, blocks, stream)));
    else
        JB_TRY((LaunchSmallMnN<R, 4>(n, k, a, b, ws, blocks, stream)));
    SmallMnFinishKernel<R><<<1, 32, 0, stream>>>(static_cast<const double2 *>(ws),
                                                static_cast<typename Cx<R>::type *>(c), static_cast<int>(m * n), blocks);
    JB_CUDA(cudaGetLastError());
    return 0;
}

struct GemmConfig {
    bool skinny;
    int bm, bn;
    int splits;
    long long k_per_split;
};

GemmConfig ChooseGemm(int64_t m, int64_t n, int64_t k)
{
    GemmConfig c;
    c.skinny = (m <= 16 || n <= 32) && (m * n <= 64 * 64 * 4);
    c.bm = c.skinny ? 16 : 64;
    c.bn = c.skinny ? 32 : 64;
    const long long tiles = ((m + c.bm - 1) / c.bm) * ((n + c.bn - 1) / c.bn);
    const long long target = static_cast<long long>(NumSMs()) * 4;
    long long splits = 1;
    if (tiles < target && k >= 2048) {
        splits = std::min<long long>(target / std::max<long long>(tiles, 1), k / 512);
        splits = std::max<long long>(splits, 1);
    }
    long long kps = (k + splits - 1) / splits;
    kps = ((kps + 31) / 32) * 32;
    splits = (k + kps - 1) / kps;
    c.splits = static_cast<int>(std::max<long long>(splits, 1));
    c.k_per_split = kps;
    return c;
}

template <typename R>
int LaunchGemmT(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws,
                size_t ws_bytes, cudaStream_t stream)
{
    using C = typename Cx<R>::type;
    const GemmConfig cfg = ChooseGemm(m, n, k);
    C *dst = static_cast<C *>(c);
    if (cfg.splits > 1) {
        const size_t need = sizeof(C) * static_cast<size_t>(cfg.splits) * m * n;
        JB_REQUIRE(ws != nullptr && ws_bytes >= need, "gemm: split-K workspace too small");
        dst = static_cast<C *>(ws);
    }
    const long long tiles = ((n + cfg.bn - 1) / cfg.bn) * ((m + cfg.bm - 1) / cfg.bm);
    JB_REQUIRE(tiles < (1ll << 31) && cfg.splits <= 65535, "gemm: problem too large for the dense kernel");
    dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(cfg.splits), 1);
    if (cfg.skinny) {
        GemmKernel<R, 16, 32, 16, 2, 2><<<grid, 128, 0, stream>>>(
            static_cast<const C *>(a), static_cast<const C *>(b), dst, m, n, k, cfg.k_per_split);
    }
    else {
        GemmKernel<R, 64, 64, 8, 4, 4><<<grid, 256, 0, stream>>>(
            static_cast<const C *>(a), static_cast<const C *>(b), dst, m, n, k, cfg.k_per_split);
    }
    JB_CUDA(cudaGetLastError());
    if (cfg.splits > 1) {
        const long long mn = m * n;
        const int rgrid = static_cast<int>(std::min<long long>((mn + 255) / 256, NumSMs() * 8ll));
        SplitKReduceKernel<R><<<rgrid, 256, 0, stream>>>(static_cast<const C *>(ws),
                                                         static_cast<C *>(c), mn, cfg.splits);
        JB_CUDA(cudaGetLastError());
    }
    return 0;
}

// =================================================================================================
// Elementwise kernels
// =================================================================================================
template <typename C>
__global__ void __launch_bounds__(256)
    AddKernel(const C *__restrict__ a, const C *__restrict__ b, C *__restrict__ c, long long n)
{
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += step) {
        const C x = a[i], y = b[i];
        c[i] = C{x.x + y.x, x.y + y.y};
    }
}

template <typename C>
__global__ void __launch_bounds__(256)
    ConjKernel(const C *__restrict__ a, C *__restrict__ c, long long n)
{
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += step) {
        const C x = a[i];
        c[i] = C{x.x, -x.y};
    }
}

// out[o][i] = in[o][value][i]
template <typename V>
__global__ void __launch_bounds__(256)
    SliceKernel(const V *__restrict__ in, V *__restrict__ out, long long outer, long long extent,
                long long inner, long long value)
{
    const long long total = outer * inner;
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += step) {
        const long long o = i / inner, r = i % inner;
        out[i] = in[(o * extent + value) * inner + r];
    }
}

int GridFor(long long n)
{
    return static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, NumSMs() * 16ll)));
}

} // namespace

// =================================================================================================
// Host API
// =================================================================================================
static bool TcEnabled()
{
    static const bool enabled = [] {
        const char *e = getenv("JB_DISABLE_TC");
        return !(e && e[0] == '1');
    }();
    return enabled;
}

static bool DmmaEnabled()
{
    static const bool enabled = [] {
        const char *e = getenv("JB_DISABLE_DMMA");
        return !(e && e[0] == '1');
    }();
    return enabled;
}

size_t GemmWorkspaceBytes(int dtype, int64_t m, int64_t n, int64_t k)
{
    if (SmallMnEligible(m, n, k))
        return SmallMnWorkspaceBytes(m, n, k);
    if (DmmaEnabled() && GemmDmmaEligible(dtype, m, n, k))
        return GemmDmmaWorkspaceBytes(m, n, k);
    if (TcEnabled() && GemmTcEligible(dtype, m, n, k))
        return GemmTcWorkspaceBytes(m, n, k);
    const GemmConfig cfg = ChooseGemm(m, n, k);
    if (cfg.splits <= 1)
        return 0;
    return ElemBytes(dtype) * static_cast<size_t>(cfg.splits) * m * n;
}

int GemmKind(int dtype, int64_t m, int64_t n, int64_t k)
{
    // the same order as LaunchGemm below
    if (SmallMnEligible(m, n, k))
        return JB_GEMM_SMALL_MN;
    if (TcEnabled() && GemmTcEligible(dtype, m, n, k))
        return JB_GEMM_TCGEN05;
    if (DmmaEnabled() && GemmDmmaEligible(dtype, m, n, k))
        return JB_GEMM_DMMA;
    return JB_GEMM_FMA;
}

int LaunchGemm(int dtype, int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c,
               void *ws, size_t ws_bytes, cudaStream_t stream)
{
    JB_REQUIRE(m >= 1 && n >= 1 && k >= 1, "gemm: dimensions must be positive");
    if (SmallMnEligible(m, n, k) && ws != nullptr && ws_bytes >= SmallMnWorkspaceBytes(m, n, k)) {
        if (dtype == JB_C64)
            return LaunchSmallMn<float>(m, n, k, a, b, c, ws, ws_bytes, stream);
        if (dtype == JB_C128)
            return LaunchSmallMn<double>(m, n, k, a, b, c, ws, ws_bytes, stream);
    }
    if (TcEnabled() && GemmTcEligible(dtype, m, n, k) && ws != nullptr && ws_bytes >= GemmTcWorkspaceBytes(m, n, k))
        return LaunchGemmTc(m, n, k, a, b, c, ws, ws_bytes, stream);
    if (DmmaEnabled() && GemmDmmaEligible(dtype, m, n, k) &&
        (GemmDmmaWorkspaceBytes(m, n, k) == 0 || (ws != nullptr && ws_bytes >= GemmDmmaWorkspaceBytes(m, n, k))))
        return LaunchGemmDmma(m, n, k, a, b, c, ws, ws_bytes, stream);
    if (dtype == JB_C64)
        return LaunchGemmT<float>(m, n, k, a, b, c, ws, ws_bytes, stream);
    if (dtype == JB_C128)
        return LaunchGemmT<double>(m, n, k, a, b, c, ws, ws_bytes, stream);
    return Fail("gemm: unknown dtype");
}

int LaunchAdd(int dtype, int64_t n, const void *a, const void *b, void *c, cudaStream_t stream)
{
    if (n <= 0)
        return 0;
    if (dtype == JB_C64)
        AddKernel<float2><<<GridFor(n), 256, 0, stream>>>(static_cast<const float2 *>(a),
                                                         static_cast<const float2 *>(b),
                                                         static_cast<float2 *>(c), n);
    else
        AddKernel<double2><<<GridFor(n), 256, 0, stream>>>(static_cast<const double2 *>(a),
                                                          static_cast<const double2 *>(b),
                                                          static_cast<double2 *>(c), n);
    JB_CUDA(cudaGetLastError());
    return 0;
}

int LaunchConj(int dtype, int64_t n, const void *in, void *out, cudaStream_t stream)
{
    if (n <= 0)
        return 0;
    if (dtype == JB_C64)
        ConjKernel<float2><<<GridFor(n), 256, 0, stream>>>(static_cast<const float2 *>(in),
                                                          static_cast<float2 *>(out), n);
    else
        ConjKernel<double2><<<GridFor(n), 256, 0, stream>>>(static_cast<const double2 *>(in),
                                                           static_cast<double2 *>(out), n);
    JB_CUDA(cudaGetLastError());
    return 0;
}

int LaunchSlice(int dtype, const void *in, void *out, int rank, const int64_t *extent, int axis,
                int64_t value, cudaStream_t stream)
{
    JB_REQUIRE(axis >= 0 && axis < rank, "slice: axis out of range");
    JB_REQUIRE(value >= 0 && value < extent[axis], "slice: value out of range");
    long long outer = 1, inner = 1;
    for (int i = 0; i < axis; i++)
        outer *= extent[i];
    for (int i = axis + 1; i < rank; i++)
        inner *= extent[i];
    const long long total = outer * inner;
    if (dtype == JB_C64)
        SliceKernel<uint2><<<GridFor(total), 256, 0, stream>>>(static_cast<const uint2 *>(in),
                                                              static_cast<uint2 *>(out), outer,
                                                              extent[axis], inner, value);
    else
        SliceKernel<uint4><<<GridFor(total), 256, 0, stream>>>(static_cast<const uint4 *>(in),
                                                              static_cast<uint4 *>(out), outer,
                                                              extent[axis], inner, value);
    JB_CUDA(cudaGetLastError());
    return 0;
}

// -------------------------------------------------------------------------------------------------
// Contraction planning: index algebra of Tensor::ContractTensors (include/jet/Tensor.hpp:714-741)
// -------------------------------------------------------------------------------------------------
int MakeContractPlan(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                     int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                     ContractPlan *plan)
{
    JB_REQUIRE(dtype == JB_C64 || dtype == JB_C128, "contract: unknown dtype");
    JB_REQUIRE(rank_a >= 0 && rank_a <= JB_MAX_RANK && rank_b >= 0 && rank_b <= JB_MAX_RANK,
               "contract: rank out of range");
    ContractPlan &P = *plan;
    P = ContractPlan();
    P.dtype = dtype;
    P.rank_a = rank_a;
    P.rank_b = rank_b;
    P.extent_a.assign(extent_a, extent_a + rank_a);
    P.extent_b.assign(extent_b, extent_b + rank_b);
    P.modes_a.assign(modes_a, modes_a + rank_a);
    P.modes_b.assign(modes_b, modes_b + rank_b);
    for (int i = 0; i < rank_a; i++)
        for (int j = i + 1; j < rank_a; j++)
            JB_REQUIRE(modes_a[i] != modes_a[j], "contract: repeated mode in A");
    for (int i = 0; i < rank_b; i++)
        for (int j = i + 1; j < rank_b; j++)
            JB_REQUIRE(modes_b[i] != modes_b[j], "contract: repeated mode in B");

    auto find = [](const int32_t *modes, int rank, int32_t m) {
        for (int i = 0; i < rank; i++)
            if (modes[i] == m)
                return i;
        return -1;
    };
    std::vector<int> left, right, common_a, common_b; // axis positions
    for (int i = 0; i < rank_a; i++) {
        const int j = find(modes_b, rank_b, modes_a[i]);
        if (j < 0)
            left.push_back(i);
        else {
            common_a.push_back(i);
            common_b.push_back(j);
        }
    }
    for (int j = 0; j < rank_b; j++)
        if (find(modes_a, rank_a, modes_b[j]) < 0)
            right.push_back(j);

    bool pow2 = true;
    P.m = P.n = P.k = 1;
    for (int i : left) {
        P.m *= extent_a[i];
        P.modes_c.push_back(modes_a[i]);
        P.extent_c.push_back(extent_a[i]);
    }
    for (int j : right) {
        P.n *= extent_b[j];
        P.modes_c.push_back(modes_b[j]);
        P.extent_c.push_back(extent_b[j]);
    }
    for (size_t q = 0; q < common_a.size(); q++) {
        // the reference takes common extents from A without cross-checking B
        // (Tensor.hpp:727-730); a mismatch is an error here
        JB_REQUIRE(extent_a[common_a[q]] == extent_b[common_b[q]],
                   "contract: contracted extents differ between A and B");
        P.k *= extent_a[common_a[q]];
    }
    P.rank_c = static_cast<int>(P.modes_c.size());
    JB_REQUIRE(P.rank_c <= JB_MAX_RANK, "contract: output rank out of range");
    for (int i = 0; i < rank_a; i++)
        pow2 = pow2 && IsPow2(extent_a[i]);
    for (int j = 0; j < rank_b; j++)
        pow2 = pow2 && IsPow2(extent_b[j]);

    const int64_t size_a = P.m * P.k, size_b = P.k * P.n;

    // ---- stream kernel: one operand small, all extents powers of two ----------------------------
    // complex128 steps whose resident operand is a 32 x 32 or larger matrix are bound by the FP64 pipe,
    // not by HBM (>= 8 flop/B): they go to TTGT + the DMMA GEMM instead of the FP64-FMA stream kernel
    const bool dmma_step = dtype == JB_C128 && P.k >= 32 && P.n >= 32 && size_a >= size_b &&
                           GemmDmmaEligible(dtype, P.m, P.n, P.k) && DmmaEnabled();
    if (pow2 && std::min(size_a, size_b) <= kStreamMaxResident && !dmma_step) {
        const bool stream_a = size_a >= size_b; // A streamed, B resident
        const int rs = stream_a ? rank_a : rank_b;
        const int rr = stream_a ? rank_b : rank_a;
        const int64_t *es = stream_a ? extent_a : extent_b;
        const int64_t *er = stream_a ? extent_b : extent_a;
        const std::vector<int> &cs_axes = stream_a ? common_a : common_b;
        const std::vector<int> &cr_axes = stream_a ? common_b : common_a;
        const std::vector<int> &free_r = stream_a ? right : left;
        // bit layout of both operands (last axis lowest)
        std::vector<int> lo_s(rs), lo_r(rr);
        int ns = 0, nr = 0;
        for (int i = rs - 1; i >= 0; i--) {
            lo_s[i] = ns;
            ns += Log2(es[i]);
        }
        for (int i = rr - 1; i >= 0; i--) {
            lo_r[i] = nr;
            nr += Log2(er[i]);
        }
        // k bits, ascending in the streamed operand's address
        struct KB {
            int s_bit, r_bit;
        };
        std::vector<KB> kb;
        for (size_t q = 0; q < cs_axes.size(); q++) {
            const int bits = Log2(es[cs_axes[q]]);
            for (int bbit = 0; bbit < bits; bbit++)
                kb.push_back({lo_s[cs_axes[q]] + bbit, lo_r[cr_axes[q]] + bbit});
        }
        std::sort(kb.begin(), kb.end(), [](const KB &x, const KB &y) { return x.s_bit < y.s_bit; });
        std::vector<int> yb; // resident free bits ascending
        for (int ax : free_r) {
            const int bits = Log2(er[ax]);
            for (int bbit = 0; bbit < bits; bbit++)
                yb.push_back(lo_r[ax] + bbit);
        }
        std::sort(yb.begin(), yb.end());
        if (kb.size() <= 12 && yb.size() <= 12 && ns <= 62) {
            StreamParams sp;
            std::memset(&sp, 0, sizeof(sp));
            sp.log_k = static_cast<int>(kb.size());
            sp.log_y = static_cast<int>(yb.size());
            sp.log_x = ns - sp.log_k;
            sp.x_count = 1ll << sp.log_x;
            for (int q = 0; q < sp.log_k; q++) {
                sp.cs[q] = static_cast<uint8_t>(kb[q].s_bit);
                sp.rk[q] = static_cast<uint8_t>(kb[q].r_bit);
            }
            for (int q = 0; q < sp.log_y; q++)
                sp.ry[q] = static_cast<uint8_t>(yb[q]);
            if (stream_a) { // C = (x << log_y) | y
                sp.out_x_shift = sp.log_y;
                sp.out_y_shift = 0;
            }
            else { // C = (y << log_x) | x
                sp.out_x_shift = 0;
                sp.out_y_shift = sp.log_x;
            }
            P.kernel = 0;
            P.ws_bytes = 0;
            P.launches = 1;
            P.stream_blob.resize(sizeof(sp) + 1);
            std::memcpy(P.stream_blob.data(), &sp, sizeof(sp));
            P.stream_blob[sizeof(sp)] = stream_a ? 1 : 0;
            return 0;
        }
    }

    // ---- TTGT -----------------------------------------------------------------------------------
    P.kernel = 1;
    P.perm_a.clear();
    P.perm_b.clear();
    // small output, moderate K: one CTA per contraction, operands read in place, batched over slices
    if (pow2 && SmallGemmEnabled() && P.m <= 16 && P.n <= 16 && P.k >= 64 && P.k <= (1 << 14)) {
        std::vector<int> lo_a(rank_a), lo_b(rank_b);
        int na = 0, nb = 0;
        for (int i = rank_a - 1; i >= 0; i--) {
            lo_a[i] = na;
            na += Log2(extent_a[i]);
        }
        for (int j = rank_b - 1; j >= 0; j--) {
            lo_b[j] = nb;
            nb += Log2(extent_b[j]);
        }
        SmallGemmParams sp;
        std::memset(&sp, 0, sizeof(sp));
        for (auto it = left.rbegin(); it != left.rend(); ++it)
            for (int bbit = 0; bbit < Log2(extent_a[*it]); bbit++)
                sp.m_a[sp.log_m++] = static_cast<uint8_t>(lo_a[*it] + bbit);
        for (auto it = right.rbegin(); it != right.rend(); ++it)
            for (int bbit = 0; bbit < Log2(extent_b[*it]); bbit++)
                sp.n_b[sp.log_n++] = static_cast<uint8_t>(lo_b[*it] + bbit);
        // k bits in A's address order (the gather of A is then as contiguous as its layout allows)
        struct KBit {
            int a, b;
        };
        std::vector<KBit> kb;
        for (size_t q = 0; q < common_a.size(); q++)
            for (int bbit = 0; bbit < Log2(extent_a[common_a[q]]); bbit++)
                kb.push_back({lo_a[common_a[q]] + bbit, lo_b[common_b[q]] + bbit});
        std::sort(kb.begin(), kb.end(), [](const KBit &x, const KBit &y) { return x.a < y.a; });
        if (kb.size() <= 24 && na <= 62 && nb <= 62) {
            sp.log_k = static_cast<int>(kb.size());
            for (size_t q = 0; q < kb.size(); q++) {
                sp.k_a[q] = static_cast<uint8_t>(kb[q].a);
                sp.k_b[q] = static_cast<uint8_t>(kb[q].b);
            }
            P.small_gemm = true;
            P.small_blob.resize(sizeof(sp));
            std::memcpy(P.small_blob.data(), &sp, sizeof(sp));
            P.permute_a = P.permute_b = false;
            P.gemm_kind = JB_GEMM_SMALL_GATHER;
            P.launches = 1;
            P.ws_gemm_off = 0;
            P.ws_gemm_bytes = 0;
            P.ws_bytes = 0;
            return 0;
        }
    }
    // DOTU / GEMV corner with a long K: read both operands once, where they lie (DotGatherKernel)
    if (pow2 && DotGatherEnabled() && SmallMnEligible(P.m, P.n, P.k) && P.k >= (1 << 16)) {
        std::vector<int> lo_a(rank_a), lo_b(rank_b);
        int na = 0, nb = 0;
        for (int i = rank_a - 1; i >= 0; i--) {
            lo_a[i] = na;
            na += Log2(extent_a[i]);
        }
        for (int j = rank_b - 1; j >= 0; j--) {
            lo_b[j] = nb;
            nb += Log2(extent_b[j]);
        }
        struct KBit {
            int a, b;
        };
        std::vector<KBit> kb;
        for (size_t q = 0; q < common_a.size(); q++)
            for (int bbit = 0; bbit < Log2(extent_a[common_a[q]]); bbit++)
                kb.push_back({lo_a[common_a[q]] + bbit, lo_b[common_b[q]] + bbit});
        DotGatherParams dp;
        std::memset(&dp, 0, sizeof(dp));
        // free index bits, least significant first (row-major over the free axes, last axis fastest)
        for (auto it = left.rbegin(); it != left.rend(); ++it)
            for (int bbit = 0; bbit < Log2(extent_a[*it]); bbit++)
                dp.m_a[dp.log_m++] = static_cast<uint8_t>(lo_a[*it] + bbit);
        for (auto it = right.rbegin(); it != right.rend(); ++it)
            for (int bbit = 0; bbit < Log2(extent_b[*it]); bbit++)
                dp.n_b[dp.log_n++] = static_cast<uint8_t>(lo_b[*it] + bbit);
        // tile bits: the 6 lowest contracted address bits of A and the 5 lowest of B (a bit may be both)
        std::vector<KBit> by_a = kb, by_b = kb;
        std::sort(by_a.begin(), by_a.end(), [](const KBit &x, const KBit &y) { return x.a < y.a; });
        std::sort(by_b.begin(), by_b.end(), [](const KBit &x, const KBit &y) { return x.b < y.b; });
        std::vector<KBit> tile;
        auto in_tile = [&](const KBit &v) {
            for (const KBit &t : tile)
                if (t.a == v.a)
                    return true;
            return false;
        };
        for (size_t q = 0; q < by_a.size() && tile.size() < 6; q++)
            tile.push_back(by_a[q]);
        for (size_t q = 0; q < by_b.size() && static_cast<int>(tile.size()) < kDotMaxTileBits; q++) {
            if (q >= 5)
                break;
            if (!in_tile(by_b[q]))
                tile.push_back(by_b[q]);
        }
        std::sort(tile.begin(), tile.end(), [](const KBit &x, const KBit &y) { return x.a < y.a; });
        // which tile bits the LANES of a warp walk (tile bits 0..4): the two lowest A bits and the three lowest B bits
        // that are not among them — a warp's request then touches ~8 sectors of A and ~8 of B, instead of 8
        // (fully coalesced) of A and up to 32 of B; the remaining bits follow in A's order.  Measured on the last step
        // of an m=20 slice (tools/gpu/dot_ab.sh), A bits among the lanes 0..5: 6.03 / 5.12 / 4.93 / 5.15 / 5.52 / 5.59 ms.
        static const int lane_a_bits = [] {
            const char *e = getenv("JB_DOT_LANE_A_BITS"); // lane bits taken from A's lowest tile bits (5 = A order)
            return e ? std::max(0, std::min(5, atoi(e))) : 2;
        }();
        if (lane_a_bits < 5 && tile.size() > 5) {
            std::vector<KBit> lanes(tile.begin(), tile.begin() + lane_a_bits), rest;
            std::vector<KBit> tile_by_b = tile;
            std::sort(tile_by_b.begin(), tile_by_b.end(), [](const KBit &x, const KBit &y) { return x.b < y.b; });
            auto among = [](const std::vector<KBit> &v, const KBit &x) {
                for (const KBit &t : v)
                    if (t.a == x.a)
                        return true;
                return false;
            };
            for (const KBit &v : tile_by_b)
                if (lanes.size() < 5 && !among(lanes, v))
                    lanes.push_back(v);
            for (const KBit &v : tile)
                if (!among(lanes, v))
                    rest.push_back(v);
            tile = lanes;
            tile.insert(tile.end(), rest.begin(), rest.end());
        }
        std::vector<KBit> outer;
        for (const KBit &v : by_a)
            if (!in_tile(v))
                outer.push_back(v);
        if (static_cast<int>(outer.size()) <= kDotMaxOuterBits && na <= 62 && nb <= 62) {
            dp.log_tile = static_cast<int>(tile.size());
            dp.log_outer = static_cast<int>(outer.size());
            for (size_t q = 0; q < tile.size(); q++) {
                dp.tile_a[q] = static_cast<uint8_t>(tile[q].a);
                dp.tile_b[q] = static_cast<uint8_t>(tile[q].b);
            }
            for (size_t q = 0; q < outer.size(); q++) {
                dp.outer_a[q] = static_cast<uint8_t>(outer[q].a);
                dp.outer_b[q] = static_cast<uint8_t>(outer[q].b);
            }
            P.gather_dot = true;
            P.dot_blob.resize(sizeof(dp));
            std::memcpy(P.dot_blob.data(), &dp, sizeof(dp));
            P.permute_a = P.permute_b = false;
            P.gemm_kind = JB_GEMM_DOT_GATHER;
            P.launches = 2;
            P.ws_gemm_off = 0;
            P.ws_gemm_bytes = sizeof(double2) * static_cast<size_t>(DotGatherBlocks(dp)) * static_cast<size_t>(P.m * P.n);
            P.ws_bytes = (P.ws_gemm_bytes + 255) & ~size_t(255);
            return 0;
        }
    }
    // complex64, short M and a long N: the tensor-core GEMM wastes most of its 128-row tile on M, and its
    // B' expansion quadruples the traffic of the LARGE operand.  Compute C^T = B^T A^T instead: B (as
    // right ++ common) is the GEMM's row operand, A (as common ++ left) the expanded one, and the small
    // C^T (N x M) is transposed into C at the end.
    P.swap_roles = dtype == JB_C64 && TcEnabled() && P.m < 128 && P.n >= 1024 && size_b >= 8 * size_a &&
                   GemmTcEligible(dtype, P.n, P.m, P.k) && !SmallMnEligible(P.m, P.n, P.k);
    if (P.swap_roles) {
        for (int i : common_a)
            P.perm_a.push_back(i);
        for (int i : left)
            P.perm_a.push_back(i);
        for (int j : right)
            P.perm_b.push_back(j);
        for (int j : common_b)
            P.perm_b.push_back(j);
    }
    else {
        for (int i : left)
            P.perm_a.push_back(i);
        for (int i : common_a)
            P.perm_a.push_back(i);
        for (int j : common_b)
            P.perm_b.push_back(j);
        for (int j : right)
            P.perm_b.push_back(j);
    }
    P.permute_a = P.permute_b = false;
    for (int i = 0; i < rank_a; i++)
        P.permute_a = P.permute_a || P.perm_a[i] != i;
    for (int j = 0; j < rank_b; j++)
        P.permute_b = P.permute_b || P.perm_b[j] != j;
    const size_t eb = ElemBytes(dtype);
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t off = 0;
    P.launches = 1;
    // complex128 on the DMMA kernel: fold Transpose(A) into the GEMM's tile loads
    if (P.permute_a && pow2 && dtype == JB_C128 && DmmaEnabled() && GemmDmmaEligible(dtype, P.m, P.n, P.k) &&
        !SmallMnEligible(P.m, P.n, P.k)) {
        std::vector<int> lo(rank_a);
        int nb = 0;
        for (int i = rank_a - 1; i >= 0; i--) {
            lo[i] = nb;
            nb += Log2(extent_a[i]);
        }
        auto bits_of = [&](const std::vector<int> &axes, std::vector<int> *out) {
            out->clear();
            for (int ax : axes)
                for (int bbit = 0; bbit < Log2(extent_a[ax]); bbit++)
                    out->push_back(lo[ax] + bbit);
            std::sort(out->begin(), out->end());
        };
        bits_of(left, &P.a_free_bits);
        bits_of(common_a, &P.a_common_bits);
        if (P.a_free_bits.size() <= 40 && P.a_common_bits.size() <= 32 && P.a_common_bits.size() >= 3 && nb <= 62) {
            P.gather_a = true;
            P.permute_a = false;
        }
    }
    if (P.permute_a) {
        P.ws_a_off = off;
        off += align(eb * static_cast<size_t>(size_a));
        P.launches++;
    }
    if (P.permute_b) {
        P.ws_b_off = off;
        off += align(eb * static_cast<size_t>(size_b));
        P.launches++;
    }
    P.gemm_kind = P.gather_a ? JB_GEMM_DMMA : (P.swap_roles ? GemmKind(dtype, P.n, P.m, P.k) : GemmKind(dtype, P.m, P.n, P.k));
    P.ws_gemm_off = off;
    P.ws_gemm_bytes = P.swap_roles ? GemmWorkspaceBytes(dtype, P.n, P.m, P.k) : GemmWorkspaceBytes(dtype, P.m, P.n, P.k);
    if (P.ws_gemm_bytes > 0)
        P.launches++; // split-K reduce, or the B expansion of the tensor-core path
    off += align(P.ws_gemm_bytes);
    if (P.swap_roles) {
        P.ws_ct_off = off;
        off += align(eb * static_cast<size_t>(P.m * P.n));
        P.launches++;
    }
    P.ws_bytes = off;
    return 0;
}

int LaunchContract(const ContractPlan &P, const void *a, const void *b, void *c, void *ws,
                   cudaStream_t stream, const BatchArgs *batch)
{
    const int nb = batch ? batch->count : 1;
    if (P.kernel == 0) {
        StreamParams sp;
        std::memcpy(&sp, P.stream_blob.data(), sizeof(sp));
        const bool stream_a = P.stream_blob[sizeof(sp)] != 0;
        const void *s = stream_a ? a : b;
        const void *r = stream_a ? b : a;
        const long long ss = batch ? (stream_a ? batch->stride_a : batch->stride_b) : 0;
        const long long sr = batch ? (stream_a ? batch->stride_b : batch->stride_a) : 0;
        const long long so = batch ? batch->stride_c : 0;
        if (P.dtype == JB_C64)
            return LaunchStream<float>(sp, s, r, c, stream, nb, ss, sr, so);
        return LaunchStream<double>(sp, s, r, c, stream, nb, ss, sr, so);
    }
    if (P.small_gemm) {
        SmallGemmParams sp;
        std::memcpy(&sp, P.small_blob.data(), sizeof(sp));
        return LaunchSmallGemm(P.dtype, sp, a, b, c, stream, batch);
    }
    if (nb > 1) {
        // the other TTGT units are not batched: the slices of a batch run one after the other through the one workspace
        for (int z = 0; z < nb; z++) {
            const unsigned char *az = static_cast<const unsigned char *>(a) + z * batch->stride_a;
            const unsigned char *bz = static_cast<const unsigned char *>(b) + z * batch->stride_b;
            unsigned char *cz = static_cast<unsigned char *>(c) + z * batch->stride_c;
            JB_TRY(LaunchContract(P, az, bz, cz, ws, stream, nullptr));
        }
        return 0;
    }
    JB_REQUIRE(P.ws_bytes == 0 || ws != nullptr, "contract: workspace required");
    if (P.gather_dot) {
        DotGatherParams dp;
        std::memcpy(&dp, P.dot_blob.data(), sizeof(dp));
        if (P.dtype == JB_C64)
            return LaunchDotGather<float>(dp, a, b, c, ws, P.ws_bytes, stream);
        return LaunchDotGather<double>(dp, a, b, c, ws, P.ws_bytes, stream);
    }
    unsigned char *w = static_cast<unsigned char *>(ws);
    const void *at = a, *bt = b;
    if (P.permute_a) {
        JB_TRY(LaunchPermute(P.dtype, a, w + P.ws_a_off, P.rank_a, P.extent_a.data(),
                             P.perm_a.data(), stream));
        at = w + P.ws_a_off;
    }
    if (P.permute_b) {
        JB_TRY(LaunchPermute(P.dtype, b, w + P.ws_b_off, P.rank_b, P.extent_b.data(),
                             P.perm_b.data(), stream));
        bt = w + P.ws_b_off;
    }
    if (P.swap_roles) {
        void *ct = w + P.ws_ct_off;
        JB_TRY(LaunchGemm(P.dtype, P.n, P.m, P.k, bt, at, ct, w + P.ws_gemm_off, P.ws_gemm_bytes, stream));
        const int64_t ext[2] = {P.n, P.m};
        const int32_t perm[2] = {1, 0};
        return LaunchPermute(P.dtype, ct, c, 2, ext, perm, stream);
    }
    if (P.gather_a)
        return LaunchGemmDmmaGatherA(P.m, P.n, P.k, a, P.a_free_bits.data(), static_cast<int>(P.a_free_bits.size()),
                                     P.a_common_bits.data(), static_cast<int>(P.a_common_bits.size()), bt, c,
                                     w ? w + P.ws_gemm_off : nullptr, P.ws_gemm_bytes, stream);
    return LaunchGemm(P.dtype, P.m, P.n, P.k, at, bt, c, w ? w + P.ws_gemm_off : nullptr,
                      P.ws_gemm_bytes, stream);
}

} // namespace jb
