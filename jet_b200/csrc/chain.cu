// K3 — fused contraction chain.
//
// Replaces a RUN of Tensor::ContractTensors calls (reference: include/jet/Tensor.hpp:709-752) issued
// by TensorNetwork::Contract / TaskBasedContractor on consecutive path steps
// (include/jet/TensorNetwork.hpp:301-328, include/jet/TaskBasedContractor.hpp:386-390) in which each
// result is contracted next with a small tensor.  The reference (and one StreamContractKernel launch
// per step) moves the large intermediate through memory once per step; this kernel moves it once
// per CHAIN: a tile that contains every address bit the chain contracts or creates is loaded into
// shared memory, all steps are applied in place (FP32 / FP64 FMA, same association as the
// step-by-step contraction: sum over k per step), and only the final tensor is stored.
// See chain_plan.h for the tile construction; the kernel below is an interpreter of its GF(2)-linear
// index maps.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "chain_plan.h"
#include "common.cuh"

namespace jb {

struct ChainPtrs {
    const void *r[kChainMaxSteps];
    // slice batching: blockIdx.y = slice within the batch; byte distance between consecutive slices' tensors
    long long stride_r[kChainMaxSteps];
    long long stride_x0, stride_xk;
};

static_assert(sizeof(ChainParams) + sizeof(ChainPtrs) <= 4000, "kernel parameter space is 4 KB");

namespace {

template <typename R> struct Cplx;
template <> struct Cplx<float> {
    using type = float2;
};
template <> struct Cplx<double> {
    using type = double2;
};

template <typename C> __device__ __forceinline__ void CMulAdd(C &acc, const C a, const C b)
{
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

// Packed FP32 pair arithmetic (sm_100a FFMA2): one instruction issues two IEEE FMAs, so a complex
// multiply-add is 2 issue slots instead of 4 — the FMA pipe, not the scheduler, becomes the limit.
__device__ __forceinline__ unsigned long long Pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 Unpack2(unsigned long long v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// acc += a * b for complex a, b with b given as (b.x, b.y, -b.y, b.x): same two FMAs per component,
// in the same order, as CMulAdd
__device__ __forceinline__ void CMulAdd2(unsigned long long &acc, const float2 a, const float4 bb)
{
    const unsigned long long ax = Pack2(a.x, a.x), ay = Pack2(a.y, a.y);
    const unsigned long long b0 = Pack2(bb.x, bb.y), b1 = Pack2(bb.z, bb.w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ax), "l"(b0));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ay), "l"(b1));
}

__device__ __forceinline__ unsigned Lin(unsigned idx, const uint16_t *col, int nbits)
{
    unsigned r = 0;
    for (int q = 0; q < nbits; q++)
        if ((idx >> q) & 1u)
            r ^= col[q];
    return r;
}

__device__ __forceinline__ unsigned long long Deposit(unsigned long long idx, const uint8_t *bit,
                                                      int nbits)
{
    unsigned long long r = 0;
    for (int q = 0; q < nbits; q++)
        r |= ((idx >> q) & 1ull) << bit[q];
    return r;
}

// One contraction step applied in place to the tile.  A "group" is one assignment of the tile bits
// the step does not contract; its K inputs are read into registers, the N outputs are written to
// the positions of the new bits.  KC = K (all k values in registers), G = groups in flight.
template <typename R, int KC, int G, int LOGT>
__device__ __forceinline__ void ChainStep(typename Cplx<R>::type *__restrict__ tile,
                                          const typename Cplx<R>::type *__restrict__ Bm,
                                          const ChainStepParams &q,
                                          const uint16_t *__restrict__ gtab, const unsigned a_tid,
                                          const int tid)
{
    using C = typename Cplx<R>::type;
    const int log_g = q.log_g;
    const int log_n = q.log_n;
    const int N = 1 << log_n;
    const int np = q.np;
    unsigned koff[KC];
#pragma unroll
    for (int kk = 0; kk < KC; kk++) {
        unsigned r = 0;
#pragma unroll
        for (int b = 0; (1 << b) < KC; b++)
            if (kk & (1 << b))
                r ^= q.kcol[b];
        koff[kk] = r;
    }
    unsigned noff_lo[4];
    noff_lo[0] = 0;
    noff_lo[1] = log_n >= 1 ? q.ncol[0] : 0u;
    noff_lo[2] = log_n >= 2 ? q.ncol[1] : 0u;
    noff_lo[3] = noff_lo[1] ^ noff_lo[2];
    const unsigned ncol2 = log_n >= 3 ? q.ncol[2] : 0u;
    const unsigned ncol3 = log_n >= 4 ? q.ncol[3] : 0u;

    const bool t_ok = tid < (1 << log_g);
    const int per_thread = log_g > LOGT ? (1 << (log_g - LOGT)) : 1;
    const C *bbase = Bm + q.b_off;

    for (int j0 = 0; j0 < per_thread; j0 += G) {
        unsigned base[G];
        bool ok[G];
        C a[G][KC];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int j = j0 + g;
            ok[g] = t_ok && j < per_thread;
            base[g] = a_tid ^ gtab[j & 31];
#pragma unroll
            for (int kk = 0; kk < KC; kk++)
                a[g][kk] = ok[g] ? tile[base[g] ^ koff[kk]] : C{R(0), R(0)};
        }
        for (int y0 = 0; y0 < N; y0 += 4) {
            const unsigned nhi = ((y0 & 4) ? ncol2 : 0u) ^ ((y0 & 8) ? ncol3 : 0u);
            C acc[G][4];
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int yy = 0; yy < 4; yy++)
                    acc[g][yy] = C{R(0), R(0)};
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const C *brow = bbase + kk * np + y0;
                C r[4];
                if constexpr (sizeof(C) == 8) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(brow);
                    const float4 v1 = *reinterpret_cast<const float4 *>(brow + 2);
                    r[0] = C{v0.x, v0.y};
                    r[1] = C{v0.z, v0.w};
                    r[2] = C{v1.x, v1.y};
                    r[3] = C{v1.z, v1.w};
                }
                else {
#pragma unroll
                    for (int yy = 0; yy < 4; yy++)
                        r[yy] = brow[yy];
                }
#pragma unroll
                for (int yy = 0; yy < 4; yy++)
#pragma unroll
                    for (int g = 0; g < G; g++)
                        CMulAdd(acc[g][yy], a[g][kk], r[yy]);
                // complex128 with K >= 8: without a fence the compiler hoists the matrix loads of all K rows to the top
                // (K x 4 x 4 registers) and spills (692 bytes of spill stores in the K = 16 step of the GBS chains)
                if constexpr (sizeof(C) == 16 && KC >= 8) {
                    if ((kk & 7) == 7)
                        asm volatile("" ::: "memory");
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (!ok[g])
                    continue;
#pragma unroll
                for (int yy = 0; yy < 4; yy++)
                    if (y0 + yy < N)
                        tile[base[g] ^ nhi ^ noff_lo[yy]] = acc[g][yy];
            }
        }
    }
}

// ---- register stage -------------------------------------------------------------------------------
// E holds the 16 elements spanned by the stage's 4 local tile positions.  A step contracts the local
// bits in MASK (k bit q <-> q-th lowest set bit of MASK) and re-creates as many new bits in the same
// places: out[g | spread(n)] = sum_k in[g | spread(k)] * B[k][n] for every assignment g of the other
// local bits.  All register indices are compile-time constants.
template <int MASK> __device__ __forceinline__ constexpr int Spread(int v)
{
    int r = 0, q = 0;
    for (int b = 0; b < kChainLocalBits; b++)
        if (MASK & (1 << b)) {
            if (v & (1 << q))
                r |= 1 << b;
            q++;
        }
    return r;
}

// The matrix of a register-stage step: complex64 entries are stored as (b.x, b.y, -b.y, b.x) so that
// both packed FMAs of a complex multiply-add read their operand from one 16-byte load.
template <typename R> struct StageB;
template <> struct StageB<float> {
    using type = float4;
};
template <> struct StageB<double> {
    using type = double2;
};

// The matrices of the register-stage steps live in the CONSTANT bank: every thread of a warp reads
// the same entry, so the loads go through the uniform datapath (LDCU into a uniform register that the
// packed FMA takes directly as an operand) and never touch the LSU, the shared-memory pipe or the
// per-lane register file.  Measured (tools/micro/ffma2_mix.cu): 97 % of the FMA pipe with two warps
// per scheduler, against 53 % when the same entries come from shared memory.
// One 16 KB region per slot; a plan (or the operator-level entry point) owns a slot while it lives.
constexpr int kChainConstSlots = 6;      // 5 plans per device (LanePlans) + the operator-level slot
constexpr int kChainConstEntries = 512;  // 16-byte entries per slot (8 KB; 48 KB of the 64 KB bank in all)
__constant__ uint4 g_chain_const[kChainConstSlots * kChainConstEntries];

template <typename R> __device__ __forceinline__ typename StageB<R>::type ConstB(int idx);
template <> __device__ __forceinline__ float4 ConstB<float>(int idx)
{
    const uint4 v = g_chain_const[idx];
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
template <> __device__ __forceinline__ double2 ConstB<double>(int idx)
{
    const uint4 v = g_chain_const[idx];
    return make_double2(__hiloint2double(v.y, v.x), __hiloint2double(v.w, v.z));
}

// NL local bits: 4 for complex64 (16 elements = 32 registers), 3 for complex128.
// B = index of the step's matrix in g_chain_const.
template <typename R, int NL, int MASK>
__device__ __forceinline__ void ApplyLocal(typename Cplx<R>::type (&E)[1 << NL], const int B)
{
    using C = typename Cplx<R>::type;
    constexpr int LK = ((MASK >> 0) & 1) + ((MASK >> 1) & 1) + ((MASK >> 2) & 1) + ((MASK >> 3) & 1);
    constexpr int K = 1 << LK;
    constexpr int NE = 1 << NL;
    constexpr int NP = K < 4 ? 4 : K; // row stride of the matrix: ChainStepParams::np for K == N
    if constexpr (sizeof(R) == 4) {
        // All groups at once: out[] accumulates every output while the matrix streams through a
        // small register ring, one chunk of CH entries (one k, CH consecutive n) per step of the
        // software pipeline; chunk c + D is loaded before chunk c is consumed, so the shared-memory
        // latency of the matrix loads hides behind 32+ packed FMAs instead of stalling each of them.
        constexpr int CH = K < 4 ? K : 4;
        constexpr int CPR = K / CH;       // chunks per matrix row
        constexpr int NCH = K * CPR;      // chunks per step
        constexpr int D = K <= 4 ? 1 : 2; // prefetch distance
        unsigned long long out[NE];
#pragma unroll
        for (int e = 0; e < NE; e++)
            out[e] = 0ull;
        float4 ring[D + 1][CH];
#pragma unroll
        for (int c = 0; c < D && c < NCH; c++)
#pragma unroll
            for (int nn = 0; nn < CH; nn++)
                ring[c][nn] = ConstB<R>(B + (c / CPR) * NP + (c % CPR) * CH + nn);
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (c + D < NCH) {
#pragma unroll
                for (int nn = 0; nn < CH; nn++)
                    ring[(c + D) % (D + 1)][nn] = ConstB<R>(B + ((c + D) / CPR) * NP + ((c + D) % CPR) * CH + nn);
            }
            const int k = c / CPR, n0 = (c % CPR) * CH;
#pragma unroll
            for (int nn = 0; nn < CH; nn++) {
#pragma unroll
                for (int g = 0; g < NE; g++) {
                    if (g & MASK)
                        continue;
                    CMulAdd2(out[g | Spread<MASK>(n0 + nn)], E[g | Spread<MASK>(k)], ring[c % (D + 1)][nn]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < NE; e++)
            E[e] = Unpack2(out[e]);
    }
    else {
        // group by group, in place: the K inputs of a group are replaced by its K outputs
#pragma unroll
        for (int g = 0; g < NE; g++) {
            if (g & MASK)
                continue;
            C acc[K];
#pragma unroll
            for (int n = 0; n < K; n++)
                acc[n] = C{R(0), R(0)};
#pragma unroll
            for (int k = 0; k < K; k++) {
                const C a = E[g | Spread<MASK>(k)];
#pragma unroll
                for (int n = 0; n < K; n++)
                    CMulAdd(acc[n], a, ConstB<R>(B + k * NP + n));
            }
#pragma unroll
            for (int n = 0; n < K; n++)
                E[g | Spread<MASK>(n)] = acc[n];
        }
    }
}

template <typename R, int NL>
__device__ __forceinline__ void ApplyLocalDispatch(typename Cplx<R>::type (&E)[1 << NL], const int B,
                                                   const int mask)
{
    if constexpr (NL == 4) {
        switch (mask) {
        case 1: ApplyLocal<R, NL, 1>(E, B); break;
        case 2: ApplyLocal<R, NL, 2>(E, B); break;
        case 3: ApplyLocal<R, NL, 3>(E, B); break;
        case 4: ApplyLocal<R, NL, 4>(E, B); break;
        case 5: ApplyLocal<R, NL, 5>(E, B); break;
        case 6: ApplyLocal<R, NL, 6>(E, B); break;
        case 7: ApplyLocal<R, NL, 7>(E, B); break;
        case 8: ApplyLocal<R, NL, 8>(E, B); break;
        case 9: ApplyLocal<R, NL, 9>(E, B); break;
        case 10: ApplyLocal<R, NL, 10>(E, B); break;
        case 11: ApplyLocal<R, NL, 11>(E, B); break;
        case 12: ApplyLocal<R, NL, 12>(E, B); break;
        case 13: ApplyLocal<R, NL, 13>(E, B); break;
        default: ApplyLocal<R, NL, 14>(E, B); break; // K = 16 steps never enter a register stage
        }
    }
    else {
        switch (mask) {
        case 1: ApplyLocal<R, NL, 1>(E, B); break;
        case 2: ApplyLocal<R, NL, 2>(E, B); break;
        case 3: ApplyLocal<R, NL, 3>(E, B); break;
        case 4: ApplyLocal<R, NL, 4>(E, B); break;
        case 5: ApplyLocal<R, NL, 5>(E, B); break;
        case 6: ApplyLocal<R, NL, 6>(E, B); break;
        default: ApplyLocal<R, NL, 7>(E, B); break;
        }
    }
}

// Shared-memory addresses are BYTE offsets into the tile: the address of register-tile element e is
// tile + ((thread part ^ table part) ^ local part(e)) — one three-input XOR per element, the tile base
// rides in the load's uniform-register operand.
template <typename R, int LOGT>
__device__ __forceinline__ void ChainRegisterStage(unsigned char *__restrict__ tile, const int const_base,
                                                   const ChainStageParams &g,
                                                   const uint16_t *__restrict__ gtab,
                                                   const unsigned a_tid, const int tid)
{
    using C = typename Cplx<R>::type;
    constexpr int NL = sizeof(R) == 4 ? 4 : 3;
    constexpr int NE = 1 << NL;
    constexpr unsigned SH = sizeof(C) == 8 ? 3 : 4;
    const int log_g = g.log_g;
    if (tid >= (1 << log_g))
        return;
    const int per_thread = log_g > LOGT ? (1 << (log_g - LOGT)) : 1;
    const int count = g.count;
    unsigned l01[4], l23[4];
    l01[0] = 0;
    l01[1] = static_cast<unsigned>(g.lcol[0]) << SH;
    l01[2] = static_cast<unsigned>(g.lcol[1]) << SH;
    l01[3] = l01[1] ^ l01[2];
    l23[0] = 0;
    l23[1] = static_cast<unsigned>(g.lcol[2]) << SH;
    l23[2] = NL == 4 ? static_cast<unsigned>(g.lcol[3]) << SH : 0u;
    l23[3] = l23[1] ^ l23[2];
    for (int j = 0; j < per_thread; j++) {
        const unsigned base = (a_tid ^ gtab[j]) << SH;
        C E[NE];
        unsigned off[NE];
#pragma unroll
        for (int e = 0; e < NE; e++) {
            off[e] = base ^ l01[e & 3] ^ l23[e >> 2];
            E[e] = *reinterpret_cast<const C *>(tile + off[e]);
        }
        // the descriptor of step t + 1 is fetched while step t computes
        unsigned desc = g.desc[0];
        for (int t = 0; t < count; t++) {
            const unsigned cur = desc;
            desc = g.desc[(t + 1) & (kChainMaxStageSteps - 1)];
            ApplyLocalDispatch<R, NL>(E, const_base + static_cast<int>(cur >> 8), static_cast<int>(cur & 0xffu));
        }
#pragma unroll
        for (int e = 0; e < NE; e++)
            *reinterpret_cast<C *>(tile + off[e]) = E[e];
    }
}

// ---- the kernel -----------------------------------------------------------------------------------
// One persistent CTA per SM, warp-specialised over a ring of tile buffers:
//   compute warps: groups of 8 warps, each group applying the chain to ITS tile in place, stage by stage
//               (FMA pipe only — these warps never touch global memory).  complex64 runs two groups on
//               two different tiles, so one group's FMA phases overlap the other's shared-memory phases
//               and stage barriers; complex128 one group;
//   2 load warps:  cp.async tile i+1 / i+2 from X_0 into a free buffer;
//   2 store warps: write tile i-1 from its buffer to X_k, two adjacent complex64 per 16-byte store.
// One memory warp sits on each of the four schedulers.  Their per-element address work is one XOR
// (shared-memory side) and one 64-bit add (global side): the thread-independent halves of both maps
// are tabulated in shared memory once per CTA.
// Hand-off: named barriers (bar.arrive by the producer, bar.sync by the consumer): full[b] load -> compute,
// done[b] compute -> store, free[b] store -> load ("buffer b may be overwritten").
// Template parameters NG = compute groups, NB = tile buffers: complex64 NG 2, NB 3 (tiles of up to 2^13 elements, 64 KB
// each); complex128 one group over three buffers.  (NG 3 / NB 5 on 2^12 tiles was measured in round 1 and dropped.)
constexpr int kChainLoadThreads = 64;
constexpr int kChainStoreThreads = 64;
constexpr int kChainMemThreads = kChainLoadThreads + kChainStoreThreads;
constexpr int kChainMemTabLen = 1 << (kChainMaxTileBits - kChainMemLogLanes);
static_assert(kChainLoadThreads == (1 << kChainMemLogLanes) && kChainStoreThreads == (1 << kChainMemLogLanes),
              "memory warps: one thread per lane of the tile walk");
// named barriers: full[b] = 1 + b, done[b] = 1 + NB + b, stage barrier of group g = 1 + 2 NB + g, free[b] = 1 + 2 NB + NG + b
// (<= 15 in all)

struct __align__(16) ChainMemEntry {
    unsigned long long g; // byte offset in X_0 / X_k
    unsigned s;           // byte offset in the tile (XORed into the thread's)
    unsigned pad;
};

// A compute GROUP works on one tile: 2^LOGT threads (== ChainLogThreads in the planner).  Kernel shapes:
//   complex64, default: two groups of 8 warps on two different 2^13-element tiles, three tile buffers;
//   complex64, JB_CHAIN_LAYOUT=4x128: four groups of 4 warps (one warp per scheduler each) on 2^12-element tiles, six
//     buffers — more independent phase streams per scheduler, but shorter chains (fewer steps fit a tile);
//   complex128: one group of 8 warps (register pressure), three buffers of 2^12 elements.
__host__ __device__ constexpr int ChainCtaThreads(int groups, int log_threads) { return (groups << log_threads) + kChainMemThreads; }

template <size_t BYTES> __device__ __forceinline__ void CpAsyncElem(unsigned dst, const unsigned char *g)
{
    if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g) : "memory");
    else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g) : "memory");
}

__device__ __forceinline__ void BarSync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// mbarriers (shared-memory address): the tile hand-off between load warps, compute groups and store warps
__device__ __forceinline__ void MbarInit(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarArrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrives once every cp.async this thread has issued so far has landed (counts as one of the expected arrivals)
__device__ __forceinline__ void MbarArriveAfterCpAsync(unsigned bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void MbarWait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "CHAIN_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra CHAIN_DONE;\n\t"
                 "bra CHAIN_WAIT;\n\t"
                 "CHAIN_DONE:\n\t"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

template <typename R> size_t ChainSmemBytes(int log_tile, int resident_elems, int n_stages, int buffers, int log_threads)
{
    using C = typename Cplx<R>::type;
    const int ct = 1 << log_threads;
    size_t b = sizeof(C) * ((static_cast<size_t>(buffers) * (size_t(1) << log_tile) + 1) & ~size_t(1));
    b += 256; // mbarriers: full / done / free per tile buffer
    b += sizeof(C) * static_cast<size_t>((resident_elems + 1) & ~1);
    b += sizeof(ChainMemEntry) * 2 * kChainMemTabLen;           // load / store tables
    b += sizeof(uint16_t) * static_cast<size_t>(n_stages) * ct; // per-thread stage offsets
    return b;
}

template <typename R, int LOGT, int NG, int NB>
__global__ void __launch_bounds__(ChainCtaThreads(NG, LOGT), 1)
    ChainKernel(const typename Cplx<R>::type *__restrict__ X0,
                typename Cplx<R>::type *__restrict__ Xk, const __grid_constant__ ChainParams p,
                const __grid_constant__ ChainPtrs rp)
{
    using C = typename Cplx<R>::type;
    X0 = reinterpret_cast<const C *>(reinterpret_cast<const unsigned char *>(X0) + blockIdx.y * rp.stride_x0);
    Xk = reinterpret_cast<C *>(reinterpret_cast<unsigned char *>(Xk) + blockIdx.y * rp.stride_xk);
    constexpr int GT = 1 << LOGT;                      // threads of one compute group
    constexpr int CT = NG * GT;                        // all compute threads (NG groups, NB tile buffers)
    constexpr int kChainBuffers = NB;
    constexpr int kBarCompute = 1; // named barriers 1 .. NG: the stage barrier of each compute group
    static_assert(kBarCompute + NG <= 16, "named barriers");
    constexpr int ML = kChainMemLogLanes;
    // complex64: a store thread owns two X_k-adjacent elements (store-index bit 0) -> 16-byte stores
    constexpr int PAIR = sizeof(C) == 8 ? 1 : 0;
    extern __shared__ __align__(16) unsigned char chain_smem[];
    C *tiles = reinterpret_cast<C *>(chain_smem);
    const int tile_elems = 1 << p.log_tile;
    // (16-byte aligned also when the tile is a single complex64 element: the matrices are read with float4 loads)
    // mbarriers full[NB], done[NB], free[NB] (256 bytes reserved)
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tiles + ((kChainBuffers * tile_elems + 1) & ~1));
    static_assert(3 * NB * 8 <= 256, "mbarrier area");
    C *Bm = reinterpret_cast<C *>(reinterpret_cast<unsigned char *>(mbar) + 256);
    const unsigned mbar_s = static_cast<unsigned>(__cvta_generic_to_shared(mbar));
    // full[b]: the load threads' copies of a tile have landed -> compute group; done[b]: the group has applied the
    // chain -> store warps; free[b]: the store threads have read the tile out -> load warps.  Tile i lives in
    // buffer b = i % NB and is the (i / NB)-th use of that buffer's three barriers: waiters pass the phase parity.
    auto bar_full = [&](int b) { return mbar_s + 8u * static_cast<unsigned>(b); };
    auto bar_done = [&](int b) { return mbar_s + 8u * static_cast<unsigned>(NB + b); };
    auto bar_free = [&](int b) { return mbar_s + 8u * static_cast<unsigned>(2 * NB + b); };
    ChainMemEntry *tab_in = reinterpret_cast<ChainMemEntry *>(Bm + ((p.resident_elems + 1) & ~1));
    ChainMemEntry *tab_out = tab_in + kChainMemTabLen;
    uint16_t *atid = reinterpret_cast<uint16_t *>(tab_out + kChainMemTabLen);
    const int tid = threadIdx.x;
    const bool pair = PAIR && p.log_tile_out >= 1;
    const int out_bits = p.log_tile_out - (pair ? 1 : 0); // store-index bits walked by lanes and table

    if (tid < GT) {
        // per-thread part of every stage's tile address (tile independent): one table lookup per stage
        // instead of a bit loop over kernel parameters in the hot path
        for (int sg = 0; sg < p.n_stages; sg++) {
            const ChainStageParams &g = p.stage[sg];
            const uint16_t *gc = g.kind == 1 ? g.gcol : p.step[g.first].gcol;
            const int lg = g.kind == 1 ? g.log_g : p.step[g.first].log_g;
            atid[sg * GT + tid] = static_cast<uint16_t>(Lin(tid, gc, lg < LOGT ? lg : LOGT));
        }
    }

    else if (tid >= CT) {
        // thread-independent halves of the load / store index maps (index bits above the lane)
        for (int e = tid - CT; e < kChainMemTabLen; e += kChainMemThreads) {
            const int ib = max(0, p.log_tile_in - ML - 3), ob = max(0, out_bits - ML);
            ChainMemEntry a{0ull, 0u, 0u}, b{0ull, 0u, 0u};
            if (e < (1 << ib)) {
                a.g = Deposit(e, p.in_gbit + ML + 3, ib) * sizeof(C);
                a.s = Lin(e, p.in_scol + ML + 3, ib) * static_cast<unsigned>(sizeof(C));
            }
            if (e < (1 << ob)) {
                b.g = Deposit(e, p.out_gbit + ML + (pair ? 1 : 0), ob) * sizeof(C);
                b.s = Lin(e, p.out_scol + ML + (pair ? 1 : 0), ob) * static_cast<unsigned>(sizeof(C));
            }
            tab_in[e] = a;
            tab_out[e] = b;
        }
    }

    if (tid == 0) {
        for (int b = 0; b < NB; b++) {
            MbarInit(bar_full(b), kChainLoadThreads);
            MbarInit(bar_done(b), GT);
            MbarInit(bar_free(b), kChainStoreThreads);
        }
    }
    // resident operands -> shared memory as K x np matrices (columns n >= N are zero): the matrices
    // of the steps that run through the generic shared-memory path
    for (int s = 0; s < p.n_steps; s++) {
        const ChainStepParams &q = p.step[s];
        const C *Rs = reinterpret_cast<const C *>(static_cast<const unsigned char *>(rp.r[s]) + blockIdx.y * rp.stride_r[s]);
        const int np = q.np;
        const int N = 1 << q.log_n;
        const int total = np << q.log_k;
        for (int e = tid; e < total; e += ChainCtaThreads(NG, LOGT)) {
            const unsigned k = e / np, n = e % np;
            C v = C{R(0), R(0)};
            if (static_cast<int>(n) < N)
                v = __ldg(Rs + (Deposit(k, q.rk, q.log_k) | Deposit(n, q.rn, q.log_n)));
            Bm[q.b_off + e] = v;
        }
    }
    __syncthreads();

    const int n_my = static_cast<int>((p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x); // tiles of this CTA
    const unsigned tiles_s = static_cast<unsigned>(__cvta_generic_to_shared(tiles));
    const unsigned tile_bytes = static_cast<unsigned>(tile_elems * sizeof(C));

    if (tid < CT) {
        // ================================ compute warps =========================================
        // group gi takes tiles gi, gi + NG, ... of this CTA; tile i lives in buffer i % kChainBuffers
        // (the shuffle tells the compiler the group index is warp-uniform: tile base, buffer index and
        // loop counters then live in uniform registers and ride in the LDS/STS address operand)
        const int gi = __shfl_sync(0xffffffffu, tid >> LOGT, 0);
        const int gt = tid & (GT - 1);
        for (int i = gi; i < n_my; i += NG) {
            const int b = i % kChainBuffers;
            C *tile = tiles + b * tile_elems;
            const unsigned use = static_cast<unsigned>(i / kChainBuffers);
            MbarWait(bar_full(b), use & 1u);
            for (int sg = 0; sg < p.n_stages; sg++) {
                const ChainStageParams &g = p.stage[sg];
                const unsigned a_tid = atid[sg * GT + gt];
                if (g.kind == 1) {
                    ChainRegisterStage<R, LOGT>(reinterpret_cast<unsigned char *>(tile), p.const_base, g,
                                                p.stage_tab[sg], a_tid, gt);
                }
                else {
                    const ChainStepParams &q = p.step[g.first];
                    constexpr bool kF = sizeof(R) == 4;
                    switch (q.log_k) {
                    case 0:
                        ChainStep<R, 1, kF ? 4 : 2, LOGT>(tile, Bm, q, p.stage_tab[sg], a_tid, gt);
                        break;
                    case 1:
                        ChainStep<R, 2, kF ? 4 : 2, LOGT>(tile, Bm, q, p.stage_tab[sg], a_tid, gt);
                        break;
                    case 2:
                        ChainStep<R, 4, kF ? 2 : 1, LOGT>(tile, Bm, q, p.stage_tab[sg], a_tid, gt);
                        break;
                    case 3:
                        ChainStep<R, 8, kF ? 2 : 1, LOGT>(tile, Bm, q, p.stage_tab[sg], a_tid, gt);
                        break;
                    default:
                        ChainStep<R, 16, 1, LOGT>(tile, Bm, q, p.stage_tab[sg], a_tid, gt);
                        break;
                    }
                }
                if (sg + 1 < p.n_stages)
                    BarSync(kBarCompute + gi, GT);
            }
            MbarArrive(bar_done(b)); // release: this thread's tile stores are visible to whoever sees the phase complete
        }
    }
    else if (tid < CT + kChainLoadThreads) {
        // ================================== load warps ===========================================
        // X_0 tile -> shared memory, coalesced along the low X_0 address bits.  Load index = lane (ML
        // bits) | jl (3 bits, offsets in registers) | jh (offsets in the shared-memory table): the
        // inner loop is XOR + 64-bit add + cp.async per element.
        const int lt = tid - CT;
        const int lane_bits = min(p.log_tile_in, ML);
        const int in_bits = max(0, min(p.log_tile_in - ML, 3));
        const int inner_n = 1 << in_bits;
        const int outer_n = p.log_tile_in > ML + 3 ? 1 << (p.log_tile_in - ML - 3) : 1;
        const unsigned s_lane = Lin(lt, p.in_scol, lane_bits) * static_cast<unsigned>(sizeof(C));
        const unsigned long long g_lane = Deposit(lt, p.in_gbit, lane_bits) * sizeof(C);
        unsigned sl[8];
        unsigned long long gl[8];
#pragma unroll
        for (int jl = 0; jl < 8; jl++) {
            sl[jl] = Lin(jl, p.in_scol + ML, in_bits) * static_cast<unsigned>(sizeof(C));
            gl[jl] = Deposit(jl, p.in_gbit + ML, in_bits) * sizeof(C);
        }
        const bool ok = lt < (1 << p.log_tile_in);
        for (int i = 0; i < n_my; i++) {
            const int b = i % kChainBuffers;
            const unsigned long long t = blockIdx.x + static_cast<unsigned long long>(i) * gridDim.x;
            const unsigned long long base = Deposit(t, p.outer_in, p.log_outer) * sizeof(C);
            if (i >= kChainBuffers) {
                // buffer b is free again once both store warps have read tile i - NB out of it (they arrive on
                // free[b] after their last shared-memory read)
                MbarWait(bar_free(b), (static_cast<unsigned>(i / kChainBuffers) - 1u) & 1u);
            }
            const unsigned buf_s = tiles_s + static_cast<unsigned>(b) * tile_bytes;
            const unsigned char *src = reinterpret_cast<const unsigned char *>(X0) + (base + g_lane);
            if (ok) {
                if (p.log_tile_in >= ML + 3) {
                    for (int jh = 0; jh < outer_n; jh++) {
                        const ChainMemEntry e = tab_in[jh];
                        const unsigned so = s_lane ^ e.s;
                        const unsigned char *gp = src + e.g;
#pragma unroll
                        for (int jl = 0; jl < 8; jl++)
                            CpAsyncElem<sizeof(C)>(buf_s + (so ^ sl[jl]), gp + gl[jl]);
                    }
                }
                else { // tiny tiles
                    for (int jl = 0; jl < inner_n; jl++)
                        CpAsyncElem<sizeof(C)>(
                            buf_s + (s_lane ^ (Lin(jl, p.in_scol + ML, in_bits) * static_cast<unsigned>(sizeof(C)))),
                            src + Deposit(jl, p.in_gbit + ML, in_bits) * sizeof(C));
                }
            }
            // no wait here: the arrival fires when this thread's copies have landed, and the thread moves on to
            // the next tile (up to NB tiles ahead of the compute groups)
            MbarArriveAfterCpAsync(bar_full(b));
        }
    }
    else {
        // ================================== store warps ==========================================
        // shared memory -> X_k tile, coalesced along the low X_k address bits
        const int lt = tid - CT - kChainLoadThreads;
        const int lane_bits = min(out_bits, ML);
        const int iters = out_bits > ML ? 1 << (out_bits - ML) : 1;
        const int sh = pair ? 1 : 0;
        const unsigned s_lane = Lin(lt, p.out_scol + sh, lane_bits) * static_cast<unsigned>(sizeof(C));
        const unsigned s_pair = pair ? p.out_scol[0] * static_cast<unsigned>(sizeof(C)) : 0u;
        const unsigned long long g_lane = Deposit(lt, p.out_gbit + sh, lane_bits) * sizeof(C);
        const bool ok = lt < (1 << out_bits);
        for (int i = 0; i < n_my; i++) {
            const int b = i % kChainBuffers;
            const unsigned long long t = blockIdx.x + static_cast<unsigned long long>(i) * gridDim.x;
            const unsigned long long base = Deposit(t, p.outer_out, p.log_outer) * sizeof(C);
            const unsigned char *buf = reinterpret_cast<const unsigned char *>(tiles) + static_cast<size_t>(b) * tile_bytes;
            unsigned char *dst = reinterpret_cast<unsigned char *>(Xk) + (base + g_lane);
            MbarWait(bar_done(b), static_cast<unsigned>(i / kChainBuffers) & 1u);
            if (ok) {
                if (pair) {
                    if constexpr (PAIR) {
#pragma unroll 4
                        for (int j = 0; j < iters; j++) {
                            const ChainMemEntry e = tab_out[j];
                            const unsigned so = s_lane ^ e.s;
                            const float2 v0 = *reinterpret_cast<const float2 *>(buf + so);
                            const float2 v1 = *reinterpret_cast<const float2 *>(buf + (so ^ s_pair));
                            *reinterpret_cast<float4 *>(dst + e.g) = make_float4(v0.x, v0.y, v1.x, v1.y);
                        }
                    }
                }
                else {
#pragma unroll 4
                    for (int j = 0; j < iters; j++) {
                        const ChainMemEntry e = tab_out[j];
                        *reinterpret_cast<C *>(dst + e.g) = *reinterpret_cast<const C *>(buf + (s_lane ^ e.s));
                    }
                }
            }
            // every global store of this thread has read its shared-memory source: hand the buffer back to the
            // load warps — unless no later tile of this CTA will use it (nobody would wait on that arrival)
            if (i + kChainBuffers < n_my)
                MbarArrive(bar_free(b));
        }
    }
}

// Small operands -> this launch's constant-bank slot, as K x np matrices in the form the register stages read from the
// constant bank: complex64 (b.x, b.y, -b.y, b.x), complex128 (b.x, b.y); columns n >= N are zero.
template <typename R>
__global__ void __launch_bounds__(256)
    ChainGatherKernel(const __grid_constant__ ChainParams p, const __grid_constant__ ChainPtrs rp,
                      uint4 *__restrict__ staging)
{
    using C = typename Cplx<R>::type;
    for (int s = blockIdx.x; s < p.n_steps; s += gridDim.x) {
        const ChainStepParams &q = p.step[s];
        const C *Rs = static_cast<const C *>(rp.r[s]);
        const int np = q.np;
        const int N = 1 << q.log_n;
        const int total = np << q.log_k;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const unsigned k = e / np, n = e % np;
            C v = C{R(0), R(0)};
            if (static_cast<int>(n) < N)
                v = __ldg(Rs + (Deposit(k, q.rk, q.log_k) | Deposit(n, q.rn, q.log_n)));
            uint4 w;
            if constexpr (sizeof(R) == 4) {
                w = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(-v.y),
                               __float_as_uint(v.x));
            }
            else {
                w = make_uint4(static_cast<unsigned>(__double2loint(v.x)), static_cast<unsigned>(__double2hiint(v.x)),
                               static_cast<unsigned>(__double2loint(v.y)), static_cast<unsigned>(__double2hiint(v.y)));
            }
            staging[q.b_off + e] = w;
        }
    }
}

template <typename R>
int LaunchChainT(ChainParams p, const ChainPtrs &ptrs, const void *x0, void *xk, int slot,
                 cudaStream_t stream, int batch)
{
    using C = typename Cplx<R>::type;
    JB_REQUIRE(slot >= 0 && slot < kChainConstSlots, "chain: no constant-bank slot");
    p.const_base = slot * kChainConstEntries;
    // The matrices of the register stages go straight into this launch's constant-bank slot: the
    // bank is ordinary device memory behind the symbol's address, and a kernel boundary separates
    // the writer from the readers (the constant cache does not outlive a launch).  Chains without a
    // register stage (small tensors) read their matrices from shared memory and skip this.
    bool uses_const = false;
    for (int sg = 0; sg < p.n_stages; sg++)
        uses_const = uses_const || p.stage[sg].kind == 1;
    if (uses_const) {
        // one set of matrices per launch: the slices of a batch must share every small operand
        for (int st = 0; st < p.n_steps; st++)
            JB_REQUIRE(batch == 1 || ptrs.stride_r[st] == 0,
                       "chain: register stages cannot be batched over slice-dependent operands");
        JB_REQUIRE(p.resident_elems <= kChainConstEntries, "chain: too many matrix entries");
        void *sym = nullptr;
        JB_CUDA(cudaGetSymbolAddress(&sym, g_chain_const));
        ChainGatherKernel<R><<<std::min(p.n_steps, 8), 256, 0, stream>>>(p, ptrs, static_cast<uint4 *>(sym) + p.const_base);
        JB_CUDA(cudaGetLastError());
    }
    JB_REQUIRE(p.log_threads == ChainLogThreads(static_cast<int>(sizeof(C))), "chain: plan / kernel thread-count mismatch");
    // a batch of slices shares the SMs: each slice gets its share of the persistent CTAs, at least one
    const int grid = static_cast<int>(
        std::max<long long>(1, std::min<long long>(p.n_tiles, std::max(1, NumSMs() / batch))));
    auto launch = [&](auto kernel, int log_threads, int groups, int buffers) -> int {
        const size_t smem = ChainSmemBytes<R>(p.log_tile, p.resident_elems, p.n_stages, buffers, log_threads);
        JB_REQUIRE(smem <= 227 * 1024, "chain: shared memory");
        if (smem > 48 * 1024)
            JB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kernel<<<dim3(grid, batch), ChainCtaThreads(groups, log_threads), smem, stream>>>(static_cast<const C *>(x0),
                                                                                        static_cast<C *>(xk), p, ptrs);
        return 0;
    };
    if constexpr (sizeof(C) == 8) {
        if (p.log_threads == 7)
            JB_TRY(launch(ChainKernel<R, 7, 4, 6>, 7, 4, 6));
        else
            JB_TRY(launch(ChainKernel<R, 8, 2, 3>, 8, 2, 3));
    }
    else {
        JB_TRY(launch(ChainKernel<R, 8, 1, 3>, 8, 1, 3));
    }
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace

// -------------------------------------------------------------------------------------------------
// Host API
// -------------------------------------------------------------------------------------------------
int ChainMaxTileBits(int dtype)
{
    // 64 KiB of shared memory per tile buffer either way; JB_CHAIN_TILE_BITS lowers it (experiments)
    static const int cap = [] {
        const char *e = getenv("JB_CHAIN_TILE_BITS");
        return e ? atoi(e) : 99;
    }();
    // 64 KiB per tile buffer (three buffers), or 32 KiB with the four-group layout (six buffers)
    return std::min(cap, dtype == JB_C64 && !ChainFourGroups() ? 13 : 12);
}

bool ChainFusionEnabled()
{
    static const bool enabled = [] {
        const char *e = getenv("JB_DISABLE_CHAIN");
        return !(e && e[0] == '1');
    }();
    return enabled;
}

bool ChainTilePaddingEnabled()
{
    static const bool enabled = [] {
        const char *e = getenv("JB_CHAIN_NO_PAD");
        return !(e && e[0] == '1');
    }();
    return enabled;
}

bool ChainStepEligible(const ContractPlan &cp, bool *x_is_left)
{
    // the chained tensor is the larger operand; the other one must be gate-sized
    if (cp.kernel != 0)
        return false;
    const int64_t size_a = cp.m * cp.k, size_b = cp.k * cp.n;
    const bool a_big = size_a >= size_b;
    const int64_t small = a_big ? size_b : size_a;
    const int64_t free_small = a_big ? cp.n : cp.m;
    if (small > 256 || cp.k > 16 || free_small > 16)
        return false;
    if (x_is_left)
        *x_is_left = a_big;
    return true;
}

int MakeChainOp(int dtype, const std::vector<int32_t> &modes_x, const std::vector<int64_t> &extent_x,
                const std::vector<ChainOperand> &ops, int max_tile_bits, ChainOp *out,
                std::string *why, bool allow_register_stages)
{
    std::string dummy;
    if (why == nullptr)
        why = &dummy;
    std::map<std::pair<int32_t, int>, int> ids;
    auto bits_of = [&](const std::vector<int32_t> &modes, const std::vector<int64_t> &ext,
                       std::vector<int> *bits) {
        bits->clear();
        for (int i = static_cast<int>(modes.size()) - 1; i >= 0; i--) {
            if (!IsPow2(ext[i]))
                return false;
            const int nb = Log2(ext[i]);
            for (int b = 0; b < nb; b++) {
                const auto key = std::make_pair(modes[i], b);
                auto it = ids.find(key);
                if (it == ids.end())
                    it = ids.emplace(key, static_cast<int>(ids.size())).first;
                bits->push_back(it->second);
            }
        }
        return true;
    };
    ChainSpec spec;
    spec.elem_bytes = static_cast<int>(ElemBytes(dtype));
    if (!bits_of(modes_x, extent_x, &spec.x0_bits)) {
        *why = "non power-of-two extent";
        return 1;
    }
    // index-level replay for the output labels (Tensor.hpp:714-741)
    std::vector<int32_t> cur_m = modes_x;
    std::vector<int64_t> cur_e = extent_x;
    double flops = 0.0, step_bytes = 0.0, r_elems = 0.0;
    auto elems_of = [](const std::vector<int64_t> &e) {
        double s = 1.0;
        for (int64_t v : e)
            s *= double(v);
        return s;
    };
    const double x0_elems = elems_of(extent_x);
    for (const ChainOperand &op : ops) {
        ChainStepSpec st;
        st.x_is_left = op.x_is_left;
        if (!bits_of(op.modes, op.extent, &st.r_bits)) {
            *why = "non power-of-two extent";
            return 1;
        }
        spec.steps.push_back(st);
        std::vector<int32_t> xm, rm;
        std::vector<int64_t> xe, re;
        double k = 1.0;
        for (size_t i = 0; i < cur_m.size(); i++) {
            const auto it = std::find(op.modes.begin(), op.modes.end(), cur_m[i]);
            if (it == op.modes.end()) {
                xm.push_back(cur_m[i]);
                xe.push_back(cur_e[i]);
            }
            else {
                if (op.extent[it - op.modes.begin()] != cur_e[i]) {
                    *why = "contract: contracted extents differ between A and B";
                    return 1;
                }
                k *= double(cur_e[i]);
            }
        }
        for (size_t j = 0; j < op.modes.size(); j++)
            if (std::find(cur_m.begin(), cur_m.end(), op.modes[j]) == cur_m.end()) {
                rm.push_back(op.modes[j]);
                re.push_back(op.extent[j]);
            }
        const double xin = elems_of(cur_e), rin = elems_of(op.extent);
        if (op.x_is_left) {
            cur_m = xm;
            cur_e = xe;
            cur_m.insert(cur_m.end(), rm.begin(), rm.end());
            cur_e.insert(cur_e.end(), re.begin(), re.end());
        }
        else {
            cur_m = rm;
            cur_e = re;
            cur_m.insert(cur_m.end(), xm.begin(), xm.end());
            cur_e.insert(cur_e.end(), xe.begin(), xe.end());
        }
        const double xout = elems_of(cur_e);
        flops += 8.0 * xout * k;
        step_bytes += double(spec.elem_bytes) * (xin + rin + xout);
        r_elems += rin;
    }
    if (static_cast<int>(cur_m.size()) > JB_MAX_RANK) {
        *why = "contract: output rank out of range";
        return 1;
    }
    ChainLayout lay;
    const int wide = spec.elem_bytes == 8 ? 5 : 4;
    // small tensors are launch-latency bound: no register stages -> no constant-bank upload, one launch
    bool reg_stages = allow_register_stages && x0_elems > double(1 << 15);
    if (!PlanChain(spec, max_tile_bits, wide, &lay, why, 0, reg_stages) &&
        !PlanChain(spec, max_tile_bits, wide - 1, &lay, why, 0, reg_stages))
        return 1;
    if (reg_stages && lay.params.resident_elems > kChainConstEntries) {
        // the matrices do not fit a constant-bank slot: all steps through the shared-memory path
        reg_stages = false;
        if (!PlanChain(spec, max_tile_bits, wide, &lay, why, 0, false) &&
            !PlanChain(spec, max_tile_bits, wide - 1, &lay, why, 0, false))
            return 1;
    }
    // Short chains touch few bits and would get tiny tiles (2^8 elements: per-tile hand-off latency dominates,
    // measured 0.3 TB/s on a rank-28 tensor).  Pad the tile with untouched bits up to the largest tile that
    // still leaves every SM a few tiles.
    if (ChainTilePaddingEnabled()) {
        const int log_x = static_cast<int>(spec.x0_bits.size());
        const int want = std::min(max_tile_bits, std::max(0, log_x - 9));
        for (int pad = want - lay.params.log_tile; pad > 0; pad--) {
            ChainLayout bigger;
            std::string w2;
            if (PlanChain(spec, max_tile_bits, wide, &bigger, &w2, pad, reg_stages) && bigger.params.log_tile <= want &&
                bigger.params.log_tile > lay.params.log_tile) {
                lay = bigger;
                break;
            }
        }
    }
    {
        const size_t smem =
            spec.elem_bytes == 8
                ? ChainSmemBytes<float>(lay.params.log_tile, lay.params.resident_elems, lay.params.n_stages,
                                        ChainFourGroups() ? 6 : 3, ChainLogThreads(8))
                : ChainSmemBytes<double>(lay.params.log_tile, lay.params.resident_elems, lay.params.n_stages, 3, ChainLogThreads(16));
        if (smem > 227 * 1024) {
            *why = "shared memory";
            return 1;
        }
        if (reg_stages && lay.params.resident_elems > kChainConstEntries) {
            *why = "too many matrix entries";
            return 1;
        }
    }
    // the bit-level replay must agree with the index-level one
    {
        std::vector<int> chk;
        bits_of(cur_m, cur_e, &chk);
        if (chk != lay.xk_bits) {
            *why = "internal: chain output layout mismatch";
            return 1;
        }
    }
    out->dtype = dtype;
    out->n_steps = static_cast<int>(ops.size());
    out->blob.resize(sizeof(ChainParams));
    std::memcpy(out->blob.data(), &lay.params, sizeof(ChainParams));
    out->log_tile = lay.params.log_tile;
    out->conflict_free = lay.conflict_free;
    out->n_stages = lay.params.n_stages;
    out->launches = 1;
    out->register_steps = 0;
    for (int sg = 0; sg < lay.params.n_stages; sg++)
        if (lay.params.stage[sg].kind == 1) {
            out->launches = 2;
            out->register_steps += lay.params.stage[sg].count;
        }
    if (const char *dump = getenv("JB_CHAIN_DUMP_STAGES"); dump && dump[0] == '1') {
        // one line per chain: tile / tensor size and, per stage, kind and the local-bit masks of its steps
        std::fprintf(stderr, "chain log_x=%d log_tile=%d stages:", static_cast<int>(spec.x0_bits.size()), lay.params.log_tile);
        for (int sg = 0; sg < lay.params.n_stages; sg++) {
            const ChainStageParams &G = lay.params.stage[sg];
            if (G.kind == 1) {
                std::fprintf(stderr, " R(");
                for (int t = 0; t < G.count; t++)
                    std::fprintf(stderr, "%s%u", t ? "," : "", G.desc[t] & 0xffu);
                std::fprintf(stderr, ")");
            }
            else
                std::fprintf(stderr, " S(k%d,n%d)", lay.params.step[G.first].log_k, lay.params.step[G.first].log_n);
        }
        std::fprintf(stderr, "\n");
    }
    out->modes_c = cur_m;
    out->extent_c = cur_e;
    out->flops = flops;
    out->step_bytes = step_bytes;
    out->bytes = double(spec.elem_bytes) * (x0_elems + r_elems + elems_of(cur_e));
    return 0;
}


// constant-bank slots: [0, kChainConstSlots - 1) for plans, the last one for operator-level calls
namespace {
std::mutex g_slot_mutex;
bool g_slot_used[64][kChainConstSlots];
} // namespace

int ChainAcquireSlot(int device)
{
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    if (device < 0 || device >= 64)
        return -1;
    for (int s = 0; s < kChainConstSlots - 1; s++)
        if (!g_slot_used[device][s]) {
            g_slot_used[device][s] = true;
            return s;
        }
    return -1;
}

void ChainReleaseSlot(int device, int slot)
{
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    if (device >= 0 && device < 64 && slot >= 0 && slot < kChainConstSlots - 1)
        g_slot_used[device][slot] = false;
}

int ChainOperatorSlot() { return kChainConstSlots - 1; }

int LaunchChain(const ChainOp &op, const void *x0, const void *const *r, void *xk, int slot,
                cudaStream_t stream, const ChainBatchArgs *batch)
{
    ChainParams p;
    JB_REQUIRE(op.blob.size() == sizeof(p), "chain: not planned");
    std::memcpy(&p, op.blob.data(), sizeof(p));
    ChainPtrs ptrs;
    std::memset(&ptrs, 0, sizeof(ptrs));
    for (int s = 0; s < op.n_steps; s++)
        ptrs.r[s] = r[s];
    int count = 1;
    if (batch != nullptr && batch->count > 1) {
        JB_REQUIRE(batch->count <= 65535, "chain: batch out of range");
        count = batch->count;
        ptrs.stride_x0 = batch->stride_x0;
        ptrs.stride_xk = batch->stride_xk;
        for (int s = 0; s < op.n_steps; s++)
            ptrs.stride_r[s] = batch->stride_r[s];
    }
    if (op.dtype == JB_C64)
        return LaunchChainT<float>(p, ptrs, x0, xk, slot, stream, count);
    return LaunchChainT<double>(p, ptrs, x0, xk, slot, stream, count);
}

} // namespace jb
