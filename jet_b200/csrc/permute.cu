// K1 — index permutation kernels (bit-exact data movement).
//
// Replaces Permuter<QFlexPermuter<1024,32>> / Permuter<DefaultPermuter<1024>>
// (reference: include/jet/permute/Permuter.hpp:50-79, permute/QFlex.hpp:37-50,
// permute/Default.hpp:21-133) as reached from Tensor::Transpose (include/jet/Tensor.hpp:579-612).
//
// Design (B200-first, not the reference's three cache-blocked host passes):
//   * every power-of-two tensor is a 2^n array and the permutation is a permutation of the n
//     ADDRESS BITS (an extent-2^b axis is b adjacent bits).  Low bits that stay in place are folded
//     into the element (8 B -> 16 B vectors).
//   * one pass.  A tile is the union of the lowest input bits (coalesced loads) and the input
//     bits that become the lowest output bits (coalesced stores), grown to 2^10..2^11 elements
//     (16 KB of shared memory).  Loads walk the tile in input order, stores in output order; the
//     tile is staged in shared memory with an XOR swizzle that makes both phases bank-conflict
//     free for ANY bit permutation.
//   * persistent CTAs (8 per SM) loop over tiles with the next tile's loads issued before the
//     current tile's stores (register double buffering), so ~128 KB per SM is in flight.
//   * non-power-of-two tensors take a mixed-radix gather kernel (correctness path; no BASELINE
//     workload reaches it).
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace jb {
namespace {

constexpr int kPermThreads = 256;
constexpr int kPermSmemBytes = 16384;

struct BitPermParams {
    int t;       // tile bits
    int n_outer; // bits enumerated by the tile index
    long long n_tiles;
    int n_swz;
    uint8_t swz_src[4], swz_dst[4];
    uint8_t tin[12];  // input address bit of tile-in bit q
    uint8_t tout[12]; // output address bit of tile-out bit q
    uint8_t pos[12];  // tile-in bit that holds tile-out bit q
    uint8_t oin[64];  // input address bit of outer bit q
    uint8_t oout[64]; // output address bit of outer bit q
};

__device__ __forceinline__ unsigned long long Scatter64(unsigned long long x, const uint8_t *dst,
                                                        int nbits)
{
    unsigned long long r = 0;
    for (int q = 0; q < nbits; q++)
        r |= ((x >> q) & 1ull) << dst[q];
    return r;
}

__device__ __forceinline__ uint32_t Swizzle(uint32_t i, const BitPermParams &p)
{
    for (int s = 0; s < p.n_swz; s++)
        i ^= ((i >> p.swz_src[s]) & 1u) << p.swz_dst[s];
    return i;
}

template <typename V, int EPT, typename OffT>
__global__ void __launch_bounds__(kPermThreads, EPT >= 8 ? 3 : 4)
    PermuteBitsKernel(const V *__restrict__ in, V *__restrict__ out,
                      const __grid_constant__ BitPermParams p)
{
    __shared__ V tile[kPermSmemBytes / sizeof(V)];
    const uint32_t tid = threadIdx.x;
    const uint32_t tile_elems = 1u << p.t;
    const bool active = (EPT > 1) || (tid < tile_elems);

    // All index maps are linear over GF(2) in the bits of idx = tid | (e << 8), so each splits into
    // a per-thread part (from tid) and a warp-uniform part (from e) that the compiler keeps in
    // uniform registers.
    const OffT in_lo = static_cast<OffT>(Scatter64(tid, p.tin, min(p.t, 8)));
    const OffT out_lo = static_cast<OffT>(Scatter64(tid, p.tout, min(p.t, 8)));
    const uint32_t sww_lo = Swizzle(tid, p);
    const uint32_t swr_lo = Swizzle(static_cast<uint32_t>(Scatter64(tid, p.pos, min(p.t, 8))), p);
    OffT in_off[EPT], out_off[EPT];
    uint32_t sw_w[EPT], sw_r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        const uint32_t hi = e * kPermThreads; // bits >= 8
        in_off[e] = in_lo | static_cast<OffT>(Scatter64(hi, p.tin, p.t));
        out_off[e] = out_lo | static_cast<OffT>(Scatter64(hi, p.tout, p.t));
        sw_w[e] = sww_lo ^ Swizzle(hi, p);
        sw_r[e] = swr_lo ^ Swizzle(static_cast<uint32_t>(Scatter64(hi, p.pos, p.t)), p);
    }

    long long o = blockIdx.x;
    V v[EPT];
    if (o < p.n_tiles && active) {
        const V *src = in + static_cast<OffT>(Scatter64(o, p.oin, p.n_outer));
#pragma unroll
        for (int e = 0; e < EPT; e++)
            v[e] = __ldg(src + in_off[e]);
    }
    while (o < p.n_tiles) {
        V *dst = out + static_cast<OffT>(Scatter64(o, p.oout, p.n_outer));
        if (active) {
#pragma unroll
            for (int e = 0; e < EPT; e++)
                tile[sw_w[e]] = v[e];
        }
        __syncthreads();
        V w[EPT];
        if (active) {
#pragma unroll
            for (int e = 0; e < EPT; e++)
                w[e] = tile[sw_r[e]];
        }
        const long long o_next = o + gridDim.x;
        if (o_next < p.n_tiles && active) {
            const V *src = in + static_cast<OffT>(Scatter64(o_next, p.oin, p.n_outer));
#pragma unroll
            for (int e = 0; e < EPT; e++)
                v[e] = __ldg(src + in_off[e]);
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < EPT; e++)
                dst[out_off[e]] = w[e];
        }
        __syncthreads();
        o = o_next;
    }
}

// ---- mixed-radix gather for non-power-of-two shapes -------------------------------------------
struct GenericPermParams {
    int rank;
    long long total;
    long long ext_out[32];
    long long stride_in[32]; // input stride (elements) of output axis j
};

template <typename V>
__global__ void __launch_bounds__(256)
    PermuteGenericKernel(const V *__restrict__ in, V *__restrict__ out, const GenericPermParams p)
{
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
         idx < p.total; idx += step) {
        long long rem = idx, off = 0;
        for (int j = p.rank - 1; j >= 0; j--) {
            const long long c = rem % p.ext_out[j];
            rem /= p.ext_out[j];
            off += c * p.stride_in[j];
        }
        out[idx] = in[off];
    }
}

// ---- host-side planning -------------------------------------------------------------------------
int PlanBitPerm(int n, const uint8_t *src, int elem_bytes, BitPermParams *p)
{
    std::memset(p, 0, sizeof(*p));
    const int t_target = std::min(n, elem_bytes == 8 ? 11 : 10);
    uint8_t inv[64];
    for (int j = 0; j < n; j++)
        inv[src[j]] = static_cast<uint8_t>(j);
    bool in_tile[64] = {false};
    int count = 0;
    auto add = [&](int bit) {
        if (!in_tile[bit]) {
            in_tile[bit] = true;
            count++;
        }
    };
    int a = 0, b = 0;
    // seed with up to 5 low input bits and 5 low output bits, then grow both runs alternately
    for (int s = 0; s < 5 && count < t_target; s++) {
        if (a < n)
            add(a++);
        if (count < t_target && b < n)
            add(src[b++]);
    }
    bool turn = false;
    while (count < t_target) {
        if (!turn && a < n)
            add(a++);
        else if (b < n)
            add(src[b++]);
        else if (a < n)
            add(a++);
        turn = !turn;
    }
    p->t = count;
    p->n_outer = n - count;
    p->n_tiles = 1ll << p->n_outer;
    int rank_in[64];
    int q = 0, qo = 0;
    for (int x = 0; x < n; x++) {
        if (in_tile[x]) {
            p->tin[q] = static_cast<uint8_t>(x);
            rank_in[x] = q++;
        }
        else {
            p->oin[qo] = static_cast<uint8_t>(x);
            p->oout[qo] = inv[x];
            qo++;
        }
    }
    q = 0;
    for (int j = 0; j < n; j++) {
        if (in_tile[src[j]]) {
            p->tout[q] = static_cast<uint8_t>(j);
            p->pos[q] = static_cast<uint8_t>(rank_in[src[j]]);
            q++;
        }
    }
    // XOR swizzle: the lanes of one shared-memory wavefront vary tile-in bits [0, sb) while
    // writing and tile-out bits [0, sb) while reading; fold the latter's high tile-in bits into
    // the unused low address bits so both phases touch 2^sb distinct bank groups.
    const int sb = elem_bytes == 8 ? 4 : 3;
    if (p->t > sb) {
        bool taken[4] = {false, false, false, false};
        for (int j = 0; j < sb; j++)
            if (p->pos[j] < sb)
                taken[p->pos[j]] = true;
        int slot = 0;
        for (int j = 0; j < sb; j++) {
            if (p->pos[j] >= sb) {
                while (taken[slot])
                    slot++;
                taken[slot] = true;
                p->swz_src[p->n_swz] = p->pos[j];
                p->swz_dst[p->n_swz] = static_cast<uint8_t>(slot);
                p->n_swz++;
            }
        }
    }
    return 0;
}

template <typename V, typename OffT>
int LaunchBitsT(const void *in, void *out, const BitPermParams &p, cudaStream_t stream)
{
    const int ept = std::max(1, (1 << p.t) / kPermThreads);
    const V *i = static_cast<const V *>(in);
    V *o = static_cast<V *>(out);
    auto grid_for = [&](auto kernel) {
        const long long resident =
            static_cast<long long>(NumSMs()) * PersistentBlocksPerSM(kernel, kPermThreads, 0);
        return static_cast<int>(std::min<long long>(p.n_tiles, resident));
    };
    switch (ept) {
    case 1:
        PermuteBitsKernel<V, 1, OffT><<<grid_for(PermuteBitsKernel<V, 1, OffT>), kPermThreads, 0, stream>>>(i, o, p);
        break;
    case 2:
        PermuteBitsKernel<V, 2, OffT><<<grid_for(PermuteBitsKernel<V, 2, OffT>), kPermThreads, 0, stream>>>(i, o, p);
        break;
    case 4:
        PermuteBitsKernel<V, 4, OffT><<<grid_for(PermuteBitsKernel<V, 4, OffT>), kPermThreads, 0, stream>>>(i, o, p);
        break;
    case 8:
        if constexpr (sizeof(V) == 8) {
            PermuteBitsKernel<V, 8, OffT><<<grid_for(PermuteBitsKernel<V, 8, OffT>), kPermThreads, 0, stream>>>(i, o, p);
            break;
        }
        [[fallthrough]];
    default:
        return Fail("permute: unsupported tile size");
    }
    JB_CUDA(cudaGetLastError());
    return 0;
}

template <typename V>
int LaunchBits(const void *in, void *out, const BitPermParams &p, cudaStream_t stream)
{
    // element offsets fit 32 bits for tensors of up to 2^32 elements
    if (p.t + p.n_outer <= 32)
        return LaunchBitsT<V, uint32_t>(in, out, p, stream);
    return LaunchBitsT<V, unsigned long long>(in, out, p, stream);
}

} // namespace

int LaunchPermute(int dtype, const void *in, void *out, int rank, const int64_t *extent,
                  const int32_t *perm, cudaStream_t stream)
{
    JB_REQUIRE(dtype == JB_C64 || dtype == JB_C128, "permute: unknown dtype");
    JB_REQUIRE(rank >= 0 && rank <= JB_MAX_RANK, "permute: rank out of range");
    // validate the permutation
    {
        bool seen[JB_MAX_RANK] = {false};
        for (int j = 0; j < rank; j++) {
            JB_REQUIRE(perm[j] >= 0 && perm[j] < rank && !seen[perm[j]],
                       "permute: perm is not a permutation of the axes");
            seen[perm[j]] = true;
            JB_REQUIRE(extent[j] >= 1, "permute: extents must be positive");
        }
    }
    // drop extent-1 axes
    std::vector<int64_t> ext;
    std::vector<int> old_to_new(rank, -1);
    for (int i = 0; i < rank; i++) {
        if (extent[i] > 1) {
            old_to_new[i] = static_cast<int>(ext.size());
            ext.push_back(extent[i]);
        }
    }
    std::vector<int> pm;
    for (int j = 0; j < rank; j++)
        if (old_to_new[perm[j]] >= 0)
            pm.push_back(old_to_new[perm[j]]);
    const int r = static_cast<int>(ext.size());
    int64_t total = 1;
    bool pow2 = true, identity = true;
    for (int i = 0; i < r; i++) {
        total *= ext[i];
        pow2 = pow2 && IsPow2(ext[i]);
        identity = identity && pm[i] == i;
    }
    const size_t eb = ElemBytes(dtype);
    if (in == out)
        return Fail("permute: in-place permutation is not supported");
    if (identity) {
        JB_CUDA(cudaMemcpyAsync(out, in, static_cast<size_t>(total) * eb, cudaMemcpyDeviceToDevice,
                                stream));
        return 0;
    }

    if (pow2) {
        // input bit layout: the last axis owns the lowest bits
        std::vector<int> lo(r), nb(r);
        int n = 0;
        for (int i = r - 1; i >= 0; i--) {
            lo[i] = n;
            nb[i] = Log2(ext[i]);
            n += nb[i];
        }
        JB_REQUIRE(n <= 62, "permute: tensor too large");
        uint8_t src[64];
        int ob = 0;
        for (int j = r - 1; j >= 0; j--) {
            const int ax = pm[j];
            for (int bbit = 0; bbit < nb[ax]; bbit++)
                src[ob++] = static_cast<uint8_t>(lo[ax] + bbit);
        }
        // fold stationary low bits into the element (8 B -> 16 B)
        int elem = static_cast<int>(eb);
        while (elem < 16 && n > 0 && src[0] == 0) {
            elem *= 2;
            n--;
            for (int j = 0; j < n; j++)
                src[j] = static_cast<uint8_t>(src[j + 1] - 1);
        }
        bool ident_bits = true;
        for (int j = 0; j < n; j++)
            ident_bits = ident_bits && src[j] == j;
        if (n == 0 || ident_bits) {
            JB_CUDA(cudaMemcpyAsync(out, in, static_cast<size_t>(total) * eb,
                                    cudaMemcpyDeviceToDevice, stream));
            return 0;
        }
        BitPermParams p;
        JB_TRY(PlanBitPerm(n, src, elem, &p));
        if (elem == 8)
            return LaunchBits<uint2>(in, out, p, stream);
        return LaunchBits<uint4>(in, out, p, stream);
    }

    // generic mixed-radix path: merge axes that stay adjacent, then gather
    std::vector<int64_t> in_stride(r);
    {
        int64_t s = 1;
        for (int i = r - 1; i >= 0; i--) {
            in_stride[i] = s;
            s *= ext[i];
        }
    }
    std::vector<int64_t> oe, os;
    for (int j = 0; j < r; j++) {
        const int ax = pm[j];
        if (!oe.empty() && j > 0 && pm[j - 1] + 1 == ax) {
            // axis ax directly follows the previous output axis in the input too: merge
            oe.back() *= ext[ax];
            os.back() = in_stride[ax];
        }
        else {
            oe.push_back(ext[ax]);
            os.push_back(in_stride[ax]);
        }
    }
    JB_REQUIRE(oe.size() <= 32, "permute: more than 32 non-mergeable non-power-of-two axes");
    GenericPermParams g;
    std::memset(&g, 0, sizeof(g));
    g.rank = static_cast<int>(oe.size());
    g.total = total;
    for (int j = 0; j < g.rank; j++) {
        g.ext_out[j] = oe[j];
        g.stride_in[j] = os[j];
    }
    const int grid = static_cast<int>(
        std::min<long long>((total + 255) / 256, static_cast<long long>(NumSMs()) * 16));
    if (dtype == JB_C64)
        PermuteGenericKernel<uint2><<<grid, 256, 0, stream>>>(static_cast<const uint2 *>(in),
                                                              static_cast<uint2 *>(out), g);
    else
        PermuteGenericKernel<uint4><<<grid, 256, 0, stream>>>(static_cast<const uint4 *>(in),
                                                              static_cast<uint4 *>(out), g);
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace jb
