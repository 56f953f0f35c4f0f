// K2a — complex64 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), 3xTF32.
//
// Replaces cblas_cgemm (reference include/jet/TensorHelpers.hpp:49-61) for compute-bound shapes
// (both operands large: the Sycamore m=20 paths, the square microbench).  Design:
//
//  * Complex -> real embedding (the "4M" decomposition as ONE real GEMM): with interleaved storage
//    a complex row-major A (M x K) *is* a real row-major A' (M x 2K) and C (M x N) *is* C' (M x 2N):
//        C'[m][2n+c] = sum_k' A'[m][k'] * B'[k'][2n+c],
//        B'[2k][2n] = Re B, B'[2k][2n+1] = Im B, B'[2k+1][2n] = -Im B, B'[2k+1][2n+1] = Re B.
//    A and C need no conversion at all; B' is built (transposed, K-major) by a small pre-pass.
//    8*M*N*K real flops, exactly the complex product.
//  * 3xTF32: every fp32 operand is split into hi = fp32 with the low 13 mantissa bits cleared
//    (exactly representable in TF32) and lo = x - hi (exact in fp32); the accumulator receives
//    lo*hi + hi*lo + hi*hi (the lo*lo term, ~2^-22 relative, is dropped).  FP32 accumulation in TMEM.
//    Two variants of where the split happens (template parameter PRE):
//      PRE = true  (both M and N >= 1024): a pre-pass in global memory (A: SplitHiLoKernel, B: inside
//        the B' expansion); TMA brings four tiles per stage.  The kernel is bound by shared-memory
//        bandwidth (three MMAs read both operand tiles from shared memory per K step) and splitting
//        inside shared memory adds a read and two writes of every tile to that same bottleneck
//        (measured 180 -> 225 TFLOP/s on 2^13 x 2^14 x 2^12).
//      PRE = false (skinny or K-heavy shapes, where operand preparation is not negligible against
//        the GEMM): TMA brings the raw fp32 tiles, four "splitter" warps rewrite each as hi in place
//        and produce the lo tile beside it; global memory is read once.
//  * The tensor core adds into its FP32 accumulator with truncation, so a long K loop in TMEM
//    drifts (measured 1e-5 relative at K = 1024).  The accumulation is therefore chunked: every
//    kTcChunk k-blocks the TMEM partial sum is promoted (tcgen05.ld) into FP32 registers with
//    round-to-nearest adds, double-buffered in TMEM so the next chunk's MMAs overlap the drain.
//  * Warp roles per CTA: warp 0 = TMA producer, warp 1 = TMEM allocator + single thread
//    tcgen05.mma issuer, then (PRE = false only) 4 splitter warps, then 4 register-accumulator +
//    epilogue warps: 192 / 320 threads.  mbarriers: full (TMA landed) [-> split (hi/lo ready)] ->
//    empty (tcgen05.commit: the MMAs reading the stage retired); tmem_full / tmem_empty per
//    accumulator buffer.
//  * Tile 128 x 128 (real columns) x 32 (one 128-byte swizzle atom of K per stage), UMMA
//    M128 N128 K8 kind::tf32, operands K-major with the 128B swizzle TMA writes.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace jb {
namespace {

constexpr int kTcBM = 128;        // rows of C' per CTA
constexpr int kTcBN = 128;        // real columns of C' per CTA (64 complex)
constexpr int kTcBK = 32;         // floats of K' per stage (128 bytes)
constexpr int kTcStages = 3;
template <bool PRE> constexpr int TcThreads() { return PRE ? 192 : 320; } // TMA, MMA, [4 splitters,] 4 accumulators
constexpr int kTcChunk = 4;       // k-blocks accumulated in TMEM before promotion to registers
constexpr int kTcBand = 16;       // tile rows per rasterisation band
constexpr uint32_t kTileBytes = kTcBM * kTcBK * 4; // 16 KB (A and B tiles have the same size)
constexpr uint32_t kStageBytes = 4 * kTileBytes;   // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t kTcSmemBytes = kTcStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr uint32_t kTmemCols = 256; // two accumulator buffers of 128 columns

__device__ __forceinline__ uint32_t SmemAddr(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void MbarInit(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void MbarArriveExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void MbarWait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "WAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\t"
                 "bra WAIT_LOOP;\n\t"
                 "DONE:\n\t"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void TmaLoad2D(uint32_t smem_dst, const CUtensorMap *map, uint32_t bar, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
                 "l"(map), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t UmmaDesc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address
    d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (unused here)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset
    d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
    return d;
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, cta_group::1
__device__ __forceinline__ void UmmaTf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                 "}" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void UmmaCommit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <bool PRE>
__global__ void __launch_bounds__(TcThreads<PRE>(), 1)
    GemmTf32x3Kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_b_lo,
                     float *__restrict__ C, int ldc /*floats*/, int k_blocks_total, int k_blocks_per_split, int tiles_n,
                     int rows_total /*M*/, uint32_t tx_bytes /*bytes one stage's two TMA boxes deliver*/,
                     long long split_stride /*floats between the partial results of two splits*/)
{
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned stage buffers (SWIZZLE_128B requirement)
    const uint32_t smem_base = (SmemAddr(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem_base - SmemAddr(smem_raw));
    const uint32_t bars = smem_base + kTcStages * kStageBytes;
    // barrier layout (8 bytes each): full[3], split[3], empty[3], tmem_full[2], tmem_empty[2], then the TMEM address
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto split_bar = [&](int s) { return bars + 8u * (kTcStages + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * kTcStages + s); };
    auto tmem_full_bar = [&](int b) { return bars + 8u * (3 * kTcStages + b); };
    auto tmem_empty_bar = [&](int b) { return bars + 8u * (3 * kTcStages + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (3 * kTcStages + 4);
    volatile uint32_t *tmem_slot_gen =
        reinterpret_cast<volatile uint32_t *>(smem_gen + kTcStages * kStageBytes + 8u * (3 * kTcStages + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Rasterisation: CTAs that run at the same time should share operand tiles in L2.  Row-major
    // order over (tile_m, tile_n) makes one wave of 148 CTAs read 1 A tile and 148 different B tiles
    // (B is then re-read from DRAM once per tile row: 64 x 2 GB on the m=20 shapes).  Bands of
    // kTcBand tile rows, tile_m fastest inside a band: a wave covers ~16 x 9 tiles.
    const int tiles_m = (rows_total + kTcBM - 1) / kTcBM;
    const int band = blockIdx.x / (kTcBand * tiles_n);
    const int first_m = band * kTcBand;
    const int band_rows = min(kTcBand, tiles_m - first_m);
    const int in_band = blockIdx.x - band * (kTcBand * tiles_n);
    const int tile_m = first_m + in_band % band_rows;
    const int tile_n = in_band / band_rows;
    // split-K: blockIdx.y owns k-blocks [kb0, kb0 + k_blocks) and writes its own partial result
    const int kb0 = blockIdx.y * k_blocks_per_split;
    const int k_blocks = min(k_blocks_per_split, k_blocks_total - kb0);
    C += static_cast<long long>(blockIdx.y) * split_stride;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kTcStages; s++) {
            MbarInit(full_bar(s), 1);
            MbarInit(split_bar(s), 128);
            MbarInit(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; b++) {
            MbarInit(tmem_full_bar(b), 1);
            MbarInit(tmem_empty_bar(b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < k_blocks; kb++) {
                const int s = kb % kTcStages;
                const uint32_t phase = (kb / kTcStages) & 1;
                MbarWait(empty_bar(s), phase ^ 1);
                const uint32_t stage = smem_base + s * kStageBytes;
                MbarArriveExpectTx(full_bar(s), tx_bytes);
                TmaLoad2D(stage, &map_a, full_bar(s), (kb0 + kb) * kTcBK, tile_m * kTcBM);
                TmaLoad2D(stage + 2 * kTileBytes, &map_b, full_bar(s), (kb0 + kb) * kTcBK, tile_n * kTcBN);
                if constexpr (PRE) {
                    TmaLoad2D(stage + kTileBytes, &map_a_lo, full_bar(s), (kb0 + kb) * kTcBK, tile_m * kTcBM);
                    TmaLoad2D(stage + 3 * kTileBytes, &map_b_lo, full_bar(s), (kb0 + kb) * kTcBK, tile_n * kTcBN);
                }
            }
        }
    }
    else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D=F32, A=B=TF32, both K-major, N=128, M=128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTcBN >> 3) << 17) | ((kTcBM >> 4) << 24);
        for (int kb = 0; kb < k_blocks; kb++) {
            const int s = kb % kTcStages;
            const uint32_t phase = (kb / kTcStages) & 1;
            const int chunk = kb / kTcChunk;
            const int buf = chunk & 1;
            const bool chunk_first = (kb % kTcChunk) == 0;
            const bool chunk_last = (kb % kTcChunk) == kTcChunk - 1 || kb == k_blocks - 1;
            if (chunk_first) // the accumulator warps have drained this TMEM buffer
                MbarWait(tmem_empty_bar(buf), ((chunk >> 1) & 1) ^ 1);
            MbarWait(PRE ? full_bar(s) : split_bar(s), phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t stage = smem_base + s * kStageBytes;
                const uint32_t tmem_d = tmem_base + buf * kTcBN;
                const uint64_t a_hi = UmmaDesc(stage), a_lo = UmmaDesc(stage + kTileBytes);
                const uint64_t b_hi = UmmaDesc(stage + 2 * kTileBytes), b_lo = UmmaDesc(stage + 3 * kTileBytes);
#pragma unroll
                for (int k = 0; k < kTcBK / 8; k++) {
                    const uint64_t adv = static_cast<uint64_t>((k * 8 * 4) >> 4); // 32 bytes per K=8 step
                    UmmaTf32(tmem_d, a_lo + adv, b_hi + adv, idesc, !(chunk_first && k == 0));
                    UmmaTf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
                    UmmaTf32(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
                }
                UmmaCommit(empty_bar(s)); // stage reusable once these MMAs have read it
                if (chunk_last)
                    UmmaCommit(tmem_full_bar(buf));
            }
            __syncwarp();
        }
    }
    else if (!PRE && warp < 6) {
        // ===== splitters (128 threads, PRE = false) =====
        const int t = threadIdx.x - 64;
        for (int kb = 0; kb < k_blocks; kb++) {
            const int s = kb % kTcStages;
            const uint32_t phase = (kb / kTcStages) & 1;
            MbarWait(full_bar(s), phase);
            unsigned char *stage = smem_gen + s * kStageBytes;
#pragma unroll
            for (int half = 0; half < 2; half++) { // A tile, then B tile
                uint4 *hi = reinterpret_cast<uint4 *>(stage + half * 2 * kTileBytes);
                uint4 *lo = reinterpret_cast<uint4 *>(stage + half * 2 * kTileBytes + kTileBytes);
#pragma unroll
                for (int i = 0; i < static_cast<int>(kTileBytes / 16 / 128); i++) {
                    const int idx = i * 128 + t;
                    uint4 v = hi[idx];
                    uint4 h, l;
                    h.x = v.x & 0xFFFFE000u;
                    h.y = v.y & 0xFFFFE000u;
                    h.z = v.z & 0xFFFFE000u;
                    h.w = v.w & 0xFFFFE000u;
                    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
                    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
                    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
                    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
                    hi[idx] = h;
                    lo[idx] = l;
                }
            }
            // make the generic-proxy writes visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            MbarArrive(split_bar(s));
        }
    }
    else {
        // ===== accumulators: promote each TMEM chunk into FP32 registers (round-to-nearest) =====
        const int quad = warp & 3; // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        float acc[kTcBN];
#pragma unroll
        for (int j = 0; j < kTcBN; j++)
            acc[j] = 0.f;
        const int n_chunks = (k_blocks + kTcChunk - 1) / kTcChunk;
        for (int chunk = 0; chunk < n_chunks; chunk++) {
            const int buf = chunk & 1;
            MbarWait(tmem_full_bar(buf), (chunk >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < kTcBN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kTcBN + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j++)
                    acc[c0 + j] += __uint_as_float(r[j]);
            }
            // this buffer may be overwritten by the chunk after next
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            MbarArrive(tmem_empty_bar(buf));
        }
        // ragged edges: rows >= M and columns >= 2N of the tile were computed from whatever the
        // shared-memory tile held beyond the TMA box (each output depends only on its own row of A'
        // and of B'^T, so they cannot contaminate valid outputs) and are simply not stored
        const int cols_valid = min(kTcBN, ldc - tile_n * kTcBN);
        if (tile_m * kTcBM + row < rows_total) {
            float *crow = C + static_cast<size_t>(tile_m * kTcBM + row) * ldc + static_cast<size_t>(tile_n) * kTcBN;
#pragma unroll
            for (int j = 0; j < kTcBN; j += 4)
                if (j < cols_valid)
                    *reinterpret_cast<float4 *>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

__device__ __forceinline__ float TfHi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// B (K x N complex, row-major) -> B'^T (2N x 2K floats, K-major), split into hi and lo:
//   row 2n   = (Re B[k][n], -Im B[k][n]) over k;  row 2n+1 = (Im B[k][n], Re B[k][n]) over k
template <bool PRE>
__global__ void __launch_bounds__(256)
    ExpandBKernel(const float2 *__restrict__ B, float *__restrict__ Bt_hi, float *__restrict__ Bt_lo, long long K,
                  long long N)
{
    __shared__ float2 tile[32][33];
    const long long k0 = static_cast<long long>(blockIdx.y) * 32, n0 = static_cast<long long>(blockIdx.x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const long long k = k0 + r, n = n0 + tx;
        tile[r][tx] = (k < K && n < N) ? B[k * N + n] : float2{0.f, 0.f};
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long n = n0 + r, k = k0 + tx;
        if (n < N && k < K) {
            const float2 v = tile[tx][r];
            const long long o0 = (2 * n) * (2 * K) + 2 * k, o1 = (2 * n + 1) * (2 * K) + 2 * k;
            if constexpr (PRE) {
                const float hx = TfHi(v.x), hy = TfHi(v.y);
                const float lx = v.x - hx, ly = v.y - hy;
                *reinterpret_cast<float2 *>(Bt_hi + o0) = float2{hx, -hy};
                *reinterpret_cast<float2 *>(Bt_hi + o1) = float2{hy, hx};
                *reinterpret_cast<float2 *>(Bt_lo + o0) = float2{lx, -ly};
                *reinterpret_cast<float2 *>(Bt_lo + o1) = float2{ly, lx};
            }
            else { // raw fp32: the kernel splits in shared memory
                *reinterpret_cast<float2 *>(Bt_hi + o0) = float2{v.x, -v.y};
                *reinterpret_cast<float2 *>(Bt_hi + o1) = float2{v.y, v.x};
            }
        }
    }
}

// A' -> (hi, lo), elementwise
__global__ void __launch_bounds__(256)
    SplitHiLoKernel(const float4 *__restrict__ in, float4 *__restrict__ hi, float4 *__restrict__ lo, long long n4)
{
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += step) {
        const float4 v = in[i];
        const float4 h = make_float4(TfHi(v.x), TfHi(v.y), TfHi(v.z), TfHi(v.w));
        hi[i] = h;
        lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn GetEncodeTiled()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int MakeMap(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols_floats, uint32_t box_rows)
{
    EncodeTiledFn enc = GetEncodeTiled();
    JB_REQUIRE(enc != nullptr, "gemm: cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t dims[2] = {cols_floats, rows};
    const cuuint64_t strides[1] = {cols_floats * 4};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTcBK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    JB_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled failed");
    return 0;
}

} // namespace

namespace {

// sum of the split-K partials in a fixed order (deterministic), accumulated in double
__global__ void __launch_bounds__(256)
    TcSplitReduceKernel(const float4 *__restrict__ partial, float4 *__restrict__ out, long long n4, int splits)
{
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += step) {
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        for (int z = 0; z < splits; z++) {
            const float4 v = partial[static_cast<long long>(z) * n4 + i];
            a += v.x;
            b += v.y;
            c += v.z;
            d += v.w;
        }
        out[i] = make_float4(static_cast<float>(a), static_cast<float>(b), static_cast<float>(c), static_cast<float>(d));
    }
}

struct TcShape {
    long long tiles_m, tiles_n, k_blocks;
    int splits;
    long long k_blocks_per_split;
};

TcShape TcChoose(int64_t m, int64_t n, int64_t k)
{
    TcShape t;
    t.tiles_m = (m + kTcBM - 1) / kTcBM;
    t.tiles_n = (2 * n + kTcBN - 1) / kTcBN;
    t.k_blocks = (2 * k) / kTcBK;
    const long long tiles = t.tiles_m * t.tiles_n;
    const int sms = NumSMs();
    // few output tiles and a long K: split K so that the grid fills whole waves of CTAs (1 CTA/SM)
    int best = 1;
    if (tiles < sms) {
        const long long max_splits = std::min<long long>(64, t.k_blocks / (4 * kTcChunk));
        double best_eff = 0.0;
        for (long long sp = 1; sp <= std::max<long long>(1, max_splits); sp++) {
            const long long ctas = tiles * sp;
            const double eff = double(ctas) / double(((ctas + sms - 1) / sms) * sms);
            if (eff > best_eff + 0.04) { // prefer fewer splits unless clearly better
                best_eff = eff;
                best = static_cast<int>(sp);
            }
        }
    }
    long long kps = (t.k_blocks + best - 1) / best;
    kps = ((kps + kTcChunk - 1) / kTcChunk) * kTcChunk;
    t.splits = static_cast<int>((t.k_blocks + kps - 1) / kps);
    t.k_blocks_per_split = kps;
    return t;
}

} // namespace

bool GemmTcEligible(int dtype, int64_t m, int64_t n, int64_t k)
{
    if (dtype != JB_C64)
        return false;
    // ragged M and N are served by TMA boxes smaller than the tile and a masked epilogue; K must be
    // whole 128-byte swizzle atoms
    static const int64_t min_k = [] {
        // the shortest K (complex) the tensor-core kernel takes: 32 = two k-blocks.  Measured on the one TTGT step of an
        // m=20 slice (M = 2^21, N = 256, K = 32; tools/gpu/tc_short_k.py): 3.01 ms against 4.81 ms on the FMA GemmKernel,
        // error 7.8e-7 against 1.5e-7 (normwise, vs float64).  JB_TC_MIN_K overrides.
        const char *e = getenv("JB_TC_MIN_K");
        return e ? static_cast<int64_t>(atoll(e)) : int64_t(32);
    }();
    if ((2 * k) % kTcBK != 0 || k < min_k || m < 32 || n < 16 || n % 2 != 0)
        return false;
    if (m > (1ll << 30) || n > (1ll << 29) || k > (1ll << 29))
        return false;
    const long long tiles = ((m + kTcBM - 1) / kTcBM) * ((2 * n + kTcBN - 1) / kTcBN);
    return tiles < (1ll << 31) && static_cast<double>(m) * n * k >= double(1ll << 24);
}

namespace {
struct TcWs {
    size_t a_hi, a_lo, b_hi, b_lo, partial, total;
};
// operands split in global memory (PRE) when preparing them is cheap against the GEMM itself
bool TcPreSplit(int64_t m, int64_t n, int64_t /*k*/) { return m >= 1024 && n >= 1024; }
TcWs TcWorkspace(int64_t m, int64_t n, int64_t k)
{
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    const TcShape t = TcChoose(m, n, k);
    const bool pre = TcPreSplit(m, n, k);
    TcWs w;
    const size_t a_bytes = pre ? align(static_cast<size_t>(8) * m * k) : 0;
    const size_t b_bytes = align(static_cast<size_t>(16) * n * k); // one B'^T array (2N x 2K floats)
    w.a_hi = 0;
    w.a_lo = w.a_hi + a_bytes;
    w.b_hi = w.a_lo + a_bytes;
    w.b_lo = w.b_hi + b_bytes;
    w.partial = w.b_lo + (pre ? b_bytes : 0);
    w.total = w.partial + (t.splits > 1 ? static_cast<size_t>(8) * t.splits * m * n : 0);
    return w;
}
} // namespace

size_t GemmTcWorkspaceBytes(int64_t m, int64_t n, int64_t k) { return TcWorkspace(m, n, k).total; }

// C(MxN) = A(MxK) * B(KxN), complex64 row-major; ws holds A'_hi, A'_lo, B'^T_hi, B'^T_lo and the split-K partials
int LaunchGemmTc(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
                 cudaStream_t stream)
{
    JB_REQUIRE(GemmTcEligible(JB_C64, m, n, k), "gemm: shape not eligible for the tensor-core kernel");
    const TcWs w = TcWorkspace(m, n, k);
    JB_REQUIRE(ws != nullptr && ws_bytes >= w.total, "gemm: tensor-core workspace too small");
    const bool pre = TcPreSplit(m, n, k);
    JB_TRY(EnsureDynamicSmem(reinterpret_cast<const void *>(GemmTf32x3Kernel<true>), kTcSmemBytes));
    JB_TRY(EnsureDynamicSmem(reinterpret_cast<const void *>(GemmTf32x3Kernel<false>), kTcSmemBytes));
    unsigned char *wb = static_cast<unsigned char *>(ws);
    float *a_hi = reinterpret_cast<float *>(wb + w.a_hi), *a_lo = reinterpret_cast<float *>(wb + w.a_lo);
    float *b_hi = reinterpret_cast<float *>(wb + w.b_hi), *b_lo = reinterpret_cast<float *>(wb + w.b_lo);
    dim3 eg(static_cast<unsigned>((n + 31) / 32), static_cast<unsigned>((k + 31) / 32));
    if (pre) {
        ExpandBKernel<true><<<eg, 256, 0, stream>>>(static_cast<const float2 *>(b), b_hi, b_lo, k, n);
        JB_CUDA(cudaGetLastError());
        const long long n4 = m * k / 2; // float4 = two complex64 (k % 16 == 0)
        const int sgrid = static_cast<int>(std::min<long long>((n4 + 255) / 256, NumSMs() * 16ll));
        SplitHiLoKernel<<<sgrid, 256, 0, stream>>>(static_cast<const float4 *>(a), reinterpret_cast<float4 *>(a_hi),
                                                   reinterpret_cast<float4 *>(a_lo), n4);
        JB_CUDA(cudaGetLastError());
    }
    else {
        ExpandBKernel<false><<<eg, 256, 0, stream>>>(static_cast<const float2 *>(b), b_hi, nullptr, k, n);
        JB_CUDA(cudaGetLastError());
    }
    const TcShape t = TcChoose(m, n, k);
    const uint32_t box_a = static_cast<uint32_t>(std::min<int64_t>(kTcBM, m));
    const uint32_t box_b = static_cast<uint32_t>(std::min<int64_t>(kTcBN, 2 * n));
    CUtensorMap map_a, map_a_lo, map_b, map_b_lo;
    JB_TRY(MakeMap(&map_a, pre ? a_hi : a, static_cast<uint64_t>(m), static_cast<uint64_t>(2 * k), box_a));
    JB_TRY(MakeMap(&map_b, b_hi, static_cast<uint64_t>(2 * n), static_cast<uint64_t>(2 * k), box_b));
    if (pre) {
        JB_TRY(MakeMap(&map_a_lo, a_lo, static_cast<uint64_t>(m), static_cast<uint64_t>(2 * k), box_a));
        JB_TRY(MakeMap(&map_b_lo, b_lo, static_cast<uint64_t>(2 * n), static_cast<uint64_t>(2 * k), box_b));
    }
    else {
        map_a_lo = map_a;
        map_b_lo = map_b;
    }
    const uint32_t tx_bytes = (pre ? 2u : 1u) * (box_a + box_b) * kTcBK * 4;
    float *dst = static_cast<float *>(c);
    float *partial = nullptr;
    if (t.splits > 1) {
        partial = reinterpret_cast<float *>(wb + w.partial);
        dst = partial;
    }
    const long long tiles = t.tiles_m * t.tiles_n;
    dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(t.splits), 1);
    if (pre)
        GemmTf32x3Kernel<true><<<grid, TcThreads<true>(), kTcSmemBytes, stream>>>(
            map_a, map_a_lo, map_b, map_b_lo, dst, static_cast<int>(2 * n), static_cast<int>(t.k_blocks),
            static_cast<int>(t.k_blocks_per_split), static_cast<int>(t.tiles_n), static_cast<int>(m), tx_bytes,
            static_cast<long long>(2) * m * n);
    else
        GemmTf32x3Kernel<false><<<grid, TcThreads<false>(), kTcSmemBytes, stream>>>(
            map_a, map_a_lo, map_b, map_b_lo, dst, static_cast<int>(2 * n), static_cast<int>(t.k_blocks),
            static_cast<int>(t.k_blocks_per_split), static_cast<int>(t.tiles_n), static_cast<int>(m), tx_bytes,
            static_cast<long long>(2) * m * n);
    JB_CUDA(cudaGetLastError());
    if (t.splits > 1) {
        const long long n4 = m * n / 2; // float4 = two complex64
        const int rgrid = static_cast<int>(std::min<long long>((n4 + 255) / 256, NumSMs() * 8ll));
        TcSplitReduceKernel<<<rgrid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(partial),
                                                       static_cast<float4 *>(c), n4, t.splits);
        JB_CUDA(cudaGetLastError());
    }
    return 0;
}

} // namespace jb
