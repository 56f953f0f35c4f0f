// Host-side planner of the fused contraction chain (pure C++, no CUDA: chain.cu includes it for
// the product, tests/cpp/chain_emu.cpp includes it to replay the same parameters on the CPU).
//
// A chain is a run of consecutive Tensor::ContractTensors calls (reference:
// include/jet/Tensor.hpp:709-752) in which every result is consumed by the next call together with a
// small tensor — the shape of TensorNetwork::Contract (include/jet/TensorNetwork.hpp:301-328) on
// circuit networks, where one large intermediate absorbs one gate-sized tensor per path step.  The
// reference writes every intermediate to memory; here the large tensor is cut into tiles that hold
// ALL address bits any step of the chain contracts or creates, a tile is loaded into shared memory
// once, every step is applied to it in place, and only the final tensor is written: HBM traffic is
// |X_0| + |X_k| + sum |R_i| instead of sum (|X_{i-1}| + |R_i| + |X_i|).
//
// Everything is a power of two, so a tensor is an array indexed by address bits and every index map
// below is GF(2)-linear: "column" c[q] is the address contribution of bit q of a work index, and
// an address is the XOR of the columns of the set bits.  The shared-memory XOR swizzle that keeps
// the load and store phases bank-conflict-free is folded into the columns.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <cstdlib>
#include <string>
#include <vector>

namespace jb {

constexpr int kChainMaxSteps = 12;
constexpr int kChainMaxTileBits = 13;
constexpr int kChainMaxOuterBits = 56;
constexpr int kChainMaxLogK = 4;
constexpr int kChainMaxLogN = 4;
// Threads of one compute group (the threads that share a tile); the kernel may run several groups
// on different tiles, and the memory warps come on top.
// 256 threads; complex64 with JB_CHAIN_LAYOUT=4x128: 128 threads (four groups per CTA on tiles of up to 2^12 elements
// instead of two groups on 2^13-element tiles; measured both ways).
inline bool ChainFourGroups()
{
    static const bool four = [] {
        const char *e = std::getenv("JB_CHAIN_LAYOUT");
        return e != nullptr && e[0] == '4';
    }();
    return four;
}
inline int ChainLogThreads(int elem_bytes) { return elem_bytes == 8 && ChainFourGroups() ? 7 : 8; }
// work-index bits above the thread id (j = index >> log_threads): at most 5 (2^12-element tiles on 2^7 threads)
constexpr int kChainTabLen = 32;
// The memory warps walk a tile with 2^kChainMemLogLanes lanes (bits above are a per-CTA table).
constexpr int kChainMemLogLanes = 6;

struct ChainStepParams {
    uint8_t log_k, log_n, log_g, pad0;
    uint16_t b_off; // element offset of this step's K x np matrix in the resident area
    uint16_t np;    // row stride of that matrix (N rounded up to 4)
    uint16_t gcol[kChainMaxTileBits]; // tile address column of group-index bit q
    uint16_t kcol[kChainMaxLogK];     // tile address column of k bit q
    uint16_t ncol[kChainMaxLogN];     // tile address column of n bit q
    uint8_t rk[kChainMaxLogK];        // resident-operand address bit of k bit q
    uint8_t rn[kChainMaxLogN];        // resident-operand address bit of n bit q
};

// A stage is what runs between two block-wide barriers: either one step through the generic
// shared-memory path (kind 0), or a run of K == N steps applied to 16-element register tiles
// (kind 1): a thread loads the 16 elements spanned by 4 "local" tile positions, applies every step
// of the stage to them in registers, and writes them back — one shared-memory round trip for the
// whole run.
constexpr int kChainLocalBits = 4; // complex64: 4 local bits (16 elements); complex128: 3
constexpr int kChainMaxStageSteps = 8;

struct ChainStageParams {
    uint8_t kind, first, count, log_g;
    uint16_t gcol[kChainMaxTileBits];      // tile address column of register-tile index bit q
    uint16_t lcol[kChainLocalBits];        // tile address column of local bit q
    uint32_t desc[kChainMaxStageSteps];    // per step: local-bit mask | matrix offset (b_off) << 8
};

struct ChainParams {
    int32_t n_steps;
    int32_t n_stages;
    int32_t log_tile;     // shared-memory tile holds 2^log_tile elements
    int32_t log_tile_in;  // tile bits alive when the tile is loaded
    int32_t log_tile_out; // tile bits alive when the tile is stored
    int32_t log_outer;
    int32_t resident_elems; // total elements of the resident area
    int32_t const_base;     // first entry of this launch's matrices in the constant bank (set at launch)
    int32_t log_threads;    // compute threads per CTA = 2^log_threads (ChainLogThreads)
    long long n_tiles;
    uint16_t in_scol[kChainMaxTileBits];  // tile address column of load-index bit q
    uint16_t out_scol[kChainMaxTileBits]; // tile address column of store-index bit q
    uint8_t in_gbit[kChainMaxTileBits];   // X_0 address bit of load-index bit q
    uint8_t out_gbit[kChainMaxTileBits];  // X_k address bit of store-index bit q
    uint8_t outer_in[kChainMaxOuterBits]; // X_0 address bit of tile-number bit q
    uint8_t outer_out[kChainMaxOuterBits]; // X_k address bit of tile-number bit q
    ChainStepParams step[kChainMaxSteps];
    ChainStageParams stage[kChainMaxSteps];
    // thread-independent part of every stage's index map, tabulated for the CTA size: entry j is the
    // contribution of the work-index bits above the thread id (j = index >> log_threads)
    uint16_t stage_tab[kChainMaxSteps][kChainTabLen];
};

struct ChainStepSpec {
    std::vector<int> r_bits; // bit ids of the small operand, address-ascending (bit 0 first)
    bool x_is_left = true;   // the chained tensor is operand A of ContractTensors(A, B)
};

struct ChainSpec {
    int elem_bytes = 8;
    std::vector<int> x0_bits; // bit ids of the chained tensor before the first step, ascending
    std::vector<ChainStepSpec> steps;
};

struct ChainLayout {
    ChainParams params;
    std::vector<int> xk_bits; // bit ids of the final tensor, address-ascending
    int conflict_free = 1;    // 0 if no swizzle was found that clears the load/store phases
};

namespace chain_detail {

inline int Rank(std::vector<unsigned> v)
{
    int r = 0;
    for (int bit = 15; bit >= 0; bit--) {
        size_t piv = r;
        while (piv < v.size() && !((v[piv] >> bit) & 1u))
            piv++;
        if (piv == v.size())
            continue;
        std::swap(v[r], v[piv]);
        for (size_t i = 0; i < v.size(); i++)
            if (i != static_cast<size_t>(r) && ((v[i] >> bit) & 1u))
                v[i] ^= v[r];
        r++;
    }
    return r;
}

} // namespace chain_detail

// Symbolic replay of the chain: bit list of every X_i.  Returns false (with a reason) if a step
// does not fit the kernel (K or N above 16, repeated bits).
inline bool ChainReplay(const ChainSpec &spec, std::vector<std::vector<int>> *x_bits,
                        std::vector<std::vector<int>> *s_bits, std::vector<std::vector<int>> *f_bits,
                        std::string *why)
{
    x_bits->assign(1, spec.x0_bits);
    s_bits->clear();
    f_bits->clear();
    for (const ChainStepSpec &st : spec.steps) {
        const std::vector<int> &cur = x_bits->back();
        std::set<int> in_r(st.r_bits.begin(), st.r_bits.end());
        std::set<int> in_x(cur.begin(), cur.end());
        if (in_r.size() != st.r_bits.size() || in_x.size() != cur.size()) {
            *why = "repeated bit";
            return false;
        }
        std::vector<int> S, F, rest;
        for (int b : cur)
            (in_r.count(b) ? S : rest).push_back(b);
        for (int b : st.r_bits)
            if (!in_x.count(b))
                F.push_back(b);
        if (static_cast<int>(S.size()) > kChainMaxLogK || static_cast<int>(F.size()) > kChainMaxLogN) {
            *why = "K or N above 16";
            return false;
        }
        std::vector<int> next;
        if (st.x_is_left) { // C = left ++ right: the new (right) bits are the lowest address bits
            next = F;
            next.insert(next.end(), rest.begin(), rest.end());
        }
        else {
            next = rest;
            next.insert(next.end(), F.begin(), F.end());
        }
        s_bits->push_back(S);
        f_bits->push_back(F);
        x_bits->push_back(next);
    }
    return true;
}

// Plans the tile and every index map.  lane_bits = number of low address bits of X_0 / X_k that a
// tile must contain (coalescing run = 2^lane_bits elements).  Fails if the tile would exceed
// max_tile_bits.
inline bool PlanChain(const ChainSpec &spec, int max_tile_bits, int lane_bits, ChainLayout *out,
                      std::string *why, int extra_quiet = 0, bool register_stages = true)
{
    using namespace chain_detail;
    std::string dummy;
    if (why == nullptr)
        why = &dummy;
    const int n_steps = static_cast<int>(spec.steps.size());
    if (n_steps < 1 || n_steps > kChainMaxSteps) {
        *why = "step count";
        return false;
    }
    if (max_tile_bits > kChainMaxTileBits)
        max_tile_bits = kChainMaxTileBits;
    std::vector<std::vector<int>> X, S, F;
    if (!ChainReplay(spec, &X, &S, &F, why))
        return false;
    const std::vector<int> &x0 = X.front();
    const std::vector<int> &xk = X.back();
    if (x0.size() > 62 || xk.size() > 62) {
        *why = "tensor too large";
        return false;
    }
    const int bank_bits = spec.elem_bytes == 8 ? 4 : 3;

    std::set<int> touched;
    for (int i = 0; i < n_steps; i++) {
        touched.insert(S[i].begin(), S[i].end());
        touched.insert(F[i].begin(), F[i].end());
    }
    std::map<int, int> addr0, addrk;
    for (size_t q = 0; q < x0.size(); q++)
        addr0[x0[q]] = static_cast<int>(q);
    for (size_t q = 0; q < xk.size(); q++)
        addrk[xk[q]] = static_cast<int>(q);

    // quiet bits: untouched bits kept in the tile for coalescing and conflict-free compute phases
    std::set<int> quiet;
    for (int q = 0; q < lane_bits && q < static_cast<int>(x0.size()); q++)
        if (!touched.count(x0[q]))
            quiet.insert(x0[q]);
    for (int q = 0; q < lane_bits && q < static_cast<int>(xk.size()); q++)
        if (!touched.count(xk[q]))
            quiet.insert(xk[q]);
    for (size_t q = 0; q < x0.size() && static_cast<int>(quiet.size()) < bank_bits; q++)
        if (!touched.count(x0[q]))
            quiet.insert(x0[q]);
    // padding: more untouched bits make the tile larger — short chains touch few bits, and a tile of 2^8
    // elements is all hand-off latency (load -> compute -> store barriers per 2 KB).  The lowest untouched
    // address bits of X_0 and of X_k are taken alternately, so that both the load and the store walk
    // longer contiguous runs.
    {
        size_t q0 = 0, qk = 0;
        bool from_x0 = true;
        while (extra_quiet > 0 && (q0 < x0.size() || qk < xk.size())) {
            const std::vector<int> &src = from_x0 ? x0 : xk;
            size_t &q = from_x0 ? q0 : qk;
            while (q < src.size() && (touched.count(src[q]) || quiet.count(src[q])))
                q++;
            if (q < src.size()) {
                quiet.insert(src[q]);
                extra_quiet--;
            }
            from_x0 = !from_x0;
        }
    }

    // physical tile positions
    std::map<int, int> pos;
    std::set<int> free_pos;
    for (int p = 0; p < 64; p++)
        free_pos.insert(p);
    auto take = [&](int bit) {
        const int p = *free_pos.begin();
        free_pos.erase(free_pos.begin());
        pos[bit] = p;
        return p;
    };
    std::vector<int> tile_in; // ids alive at load, X_0 address-ascending
    for (int b : x0)
        if (quiet.count(b))
            take(b);
    for (int b : x0)
        if (touched.count(b))
            take(b);
    for (int b : x0)
        if (quiet.count(b) || touched.count(b))
            tile_in.push_back(b);
    int log_tile = static_cast<int>(tile_in.size());

    ChainParams &P = out->params;
    std::memset(&P, 0, sizeof(P));
    P.n_steps = n_steps;

    // per-step position bookkeeping (swizzle applied afterwards)
    struct StepPos {
        std::vector<int> g, k, n, alive_after;
        bool in_place = false;
    };
    std::vector<StepPos> sp(n_steps);
    std::set<int> alive(tile_in.begin(), tile_in.end());
    int resident = 0;
    for (int i = 0; i < n_steps; i++) {
        const ChainStepSpec &st = spec.steps[i];
        std::set<int> sset(S[i].begin(), S[i].end());
        std::vector<int> gpos;
        for (int b : alive)
            if (!sset.count(b))
                gpos.push_back(pos[b]);
        std::sort(gpos.begin(), gpos.end());
        sp[i].g = gpos;
        // A step that creates no more bits than it contracts reuses the contracted positions: k bit q <->
        // the contracted bit at the q-th lowest tile position, the q-th new bit takes over exactly that
        // position, and positions left over (K > N: the tensor shrinks) die.  Such steps can run in a
        // register stage: the matrix is padded to K x K with zero columns, the dead positions hold zeros.
        const bool in_place = !S[i].empty() && F[i].size() <= S[i].size();
        if (in_place)
            std::sort(S[i].begin(), S[i].end(), [&](int x, int y) { return pos[x] < pos[y]; });
        for (int b : S[i]) {
            sp[i].k.push_back(pos[b]);
            alive.erase(b);
        }
        for (size_t q = 0; q < S[i].size(); q++) {
            const int b = S[i][q];
            if (!in_place || q >= F[i].size())
                free_pos.insert(pos[b]);
            pos.erase(b);
        }
        for (size_t q = 0; q < F[i].size(); q++) {
            const int b = F[i][q];
            if (in_place) {
                pos[b] = sp[i].k[q];
                sp[i].n.push_back(sp[i].k[q]);
            }
            else {
                sp[i].n.push_back(take(b));
            }
            alive.insert(b);
        }
        sp[i].in_place = in_place;
        for (int b : alive)
            sp[i].alive_after.push_back(pos[b]);
        log_tile = std::max(log_tile, static_cast<int>(alive.size()));
        for (int p : sp[i].n)
            log_tile = std::max(log_tile, p + 1);
        ChainStepParams &Q = P.step[i];
        Q.log_k = static_cast<uint8_t>(S[i].size());
        Q.log_n = static_cast<uint8_t>(F[i].size());
        Q.log_g = static_cast<uint8_t>(gpos.size());
        // row stride of the K x np matrix: N rounded up to 4; K for position-reusing steps (K >= N), whose
        // register-stage form is a square matrix with zero columns beyond N
        Q.np = static_cast<uint16_t>(std::max(4, 1 << (in_place ? Q.log_k : Q.log_n)));
        Q.b_off = static_cast<uint16_t>(resident);
        resident += (1 << Q.log_k) * Q.np;
        for (size_t q = 0; q < S[i].size(); q++) {
            const auto it = std::find(st.r_bits.begin(), st.r_bits.end(), S[i][q]);
            Q.rk[q] = static_cast<uint8_t>(it - st.r_bits.begin());
        }
        for (size_t q = 0; q < F[i].size(); q++) {
            const auto it = std::find(st.r_bits.begin(), st.r_bits.end(), F[i][q]);
            Q.rn[q] = static_cast<uint8_t>(it - st.r_bits.begin());
        }
    }
    if (log_tile > max_tile_bits) {
        *why = "tile too large";
        return false;
    }
    std::vector<int> tile_out; // ids alive at store, X_k address-ascending
    for (int b : xk)
        if (alive.count(b))
            tile_out.push_back(b);
    if (tile_out.size() != alive.size()) {
        *why = "internal: tile/output mismatch";
        return false;
    }
    std::vector<int> outer; // untouched bits outside the tile, X_0 address-ascending
    for (int b : x0)
        if (!touched.count(b) && !quiet.count(b))
            outer.push_back(b);
    if (static_cast<int>(outer.size()) > kChainMaxOuterBits) {
        *why = "too many outer bits";
        return false;
    }

    // ---- swizzle: bank bits of physical position p >= bank_bits get XORed with sw[p] ---------------
    std::vector<unsigned> sw(64, 0);
    // The store warps write X_k-adjacent PAIRS of complex64 elements (one 16-byte store): the lanes of
    // one shared-memory read run over store-index bits 1.. instead of 0..
    const int store_skip = spec.elem_bytes == 8 && static_cast<int>(alive.size()) > bank_bits ? 1 : 0;
    auto lane_positions = [&](const std::vector<int> &ids) {
        std::vector<int> r;
        for (int q = store_skip; q < store_skip + bank_bits && q < static_cast<int>(ids.size()); q++)
            r.push_back(pos.count(ids[q]) ? pos[ids[q]] : -1);
        return r;
    };
    // positions at load time differ from the final `pos` map for bits that died: recompute
    std::map<int, int> pos0;
    {
        int p = 0;
        for (int b : x0)
            if (quiet.count(b))
                pos0[b] = p++;
        for (int b : x0)
            if (touched.count(b))
                pos0[b] = p++;
    }
    std::vector<int> lin, lout;
    for (int q = 0; q < bank_bits && q < static_cast<int>(tile_in.size()); q++)
        lin.push_back(pos0[tile_in[q]]);
    lout = lane_positions(tile_out);
    auto bank_vec = [&](int p) { return p < bank_bits ? (1u << p) : sw[p]; };
    auto phase_ok = [&](const std::vector<int> &lp) {
        std::vector<unsigned> v;
        for (int p : lp)
            v.push_back(bank_vec(p));
        return Rank(v) == static_cast<int>(lp.size());
    };
    out->conflict_free = 0;
    {
        std::vector<int> high;
        for (int p : lin)
            if (p >= bank_bits)
                high.push_back(p);
        for (int p : lout)
            if (p >= bank_bits && std::find(high.begin(), high.end(), p) == high.end())
                high.push_back(p);
        unsigned long long rng = 0x9E3779B97F4A7C15ull;
        const unsigned nvec = 1u << bank_bits;
        for (int attempt = 0; attempt < 4096 && !out->conflict_free; attempt++) {
            for (int p : high) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                sw[p] = attempt == 0 ? 0u : static_cast<unsigned>((rng >> 33) % nvec);
            }
            if (phase_ok(lin) && phase_ok(lout))
                out->conflict_free = 1;
        }
        if (!out->conflict_free)
            for (int p : high)
                sw[p] = 0;
    }
    auto col = [&](int p) { return static_cast<uint16_t>((1u << p) ^ (p >= bank_bits ? sw[p] : 0u)); };

    for (int i = 0; i < n_steps; i++) {
        ChainStepParams &Q = P.step[i];
        for (size_t q = 0; q < sp[i].g.size(); q++)
            Q.gcol[q] = col(sp[i].g[q]);
        for (size_t q = 0; q < sp[i].k.size(); q++)
            Q.kcol[q] = col(sp[i].k[q]);
        for (size_t q = 0; q < sp[i].n.size(); q++)
            Q.ncol[q] = col(sp[i].n[q]);
    }
    // ---- stages ---------------------------------------------------------------------------------
    {
        const int local_bits = spec.elem_bytes == 8 ? kChainLocalBits : kChainLocalBits - 1;
        int n_stages = 0;
        int i = 0;
        while (i < n_steps) {
            ChainStageParams &G = P.stage[n_stages++];
            G.first = static_cast<uint8_t>(i);
            G.count = 1;
            G.kind = 0;
            std::vector<int> occupied = sp[i].alive_after; // invariant over a run of in-place steps
            std::sort(occupied.begin(), occupied.end());
            if (!register_stages || !sp[i].in_place || static_cast<int>(occupied.size()) < local_bits ||
                static_cast<int>(sp[i].k.size()) > 3) {
                i++;
                continue;
            }
            std::set<int> local(sp[i].k.begin(), sp[i].k.end());
            int j = i + 1;
            while (j < n_steps && sp[j].in_place && sp[j].k.size() <= 3 && j - i < kChainMaxStageSteps) {
                std::set<int> u = local;
                u.insert(sp[j].k.begin(), sp[j].k.end());
                if (static_cast<int>(u.size()) > local_bits)
                    break;
                local = u;
                j++;
            }
            // pad the local set with the highest other occupied positions
            for (auto it = occupied.rbegin(); it != occupied.rend() && static_cast<int>(local.size()) < local_bits; ++it)
                local.insert(*it);
            std::vector<int> lvec(local.begin(), local.end());
            G.kind = 1;
            G.count = static_cast<uint8_t>(j - i);
            int ng = 0;
            for (int pp : occupied)
                if (!local.count(pp))
                    G.gcol[ng++] = col(pp);
            G.log_g = static_cast<uint8_t>(ng);
            for (int q = 0; q < local_bits; q++)
                G.lcol[q] = col(lvec[q]);
            for (int t = i; t < j; t++) {
                unsigned m = 0;
                for (int pp : sp[t].k)
                    m |= 1u << (std::find(lvec.begin(), lvec.end(), pp) - lvec.begin());
                G.desc[t - i] = m | (static_cast<uint32_t>(P.step[t].b_off) << 8);
            }
            i = j;
        }
        P.n_stages = n_stages;
    }
    P.log_tile = log_tile;
    P.log_tile_in = static_cast<int>(tile_in.size());
    P.log_tile_out = static_cast<int>(tile_out.size());
    P.log_outer = static_cast<int>(outer.size());
    P.n_tiles = 1ll << P.log_outer;
    P.resident_elems = resident;
    for (size_t q = 0; q < tile_in.size(); q++) {
        P.in_gbit[q] = static_cast<uint8_t>(addr0[tile_in[q]]);
        P.in_scol[q] = col(pos0[tile_in[q]]);
    }
    for (size_t q = 0; q < tile_out.size(); q++) {
        P.out_gbit[q] = static_cast<uint8_t>(addrk[tile_out[q]]);
        P.out_scol[q] = col(pos[tile_out[q]]);
    }
    for (size_t q = 0; q < outer.size(); q++) {
        P.outer_in[q] = static_cast<uint8_t>(addr0[outer[q]]);
        P.outer_out[q] = static_cast<uint8_t>(addrk[outer[q]]);
    }
    // ---- tables ---------------------------------------------------------------------------------
    {
        auto lin = [](unsigned idx, const uint16_t *c, int nbits) {
            unsigned r = 0;
            for (int q = 0; q < nbits; q++)
                if ((idx >> q) & 1u)
                    r ^= c[q];
            return r;
        };
        const int L = ChainLogThreads(spec.elem_bytes);
        P.log_threads = L;
        for (int j = 0; j < kChainTabLen; j++) {
            for (int sg = 0; sg < P.n_stages; sg++) {
                const ChainStageParams &G = P.stage[sg];
                const uint16_t *gc = G.kind == 1 ? G.gcol : P.step[G.first].gcol;
                const int lg = G.kind == 1 ? G.log_g : P.step[G.first].log_g;
                P.stage_tab[sg][j] = static_cast<uint16_t>(lin(j, gc + L, std::max(0, lg - L)));
            }
        }
    }
    out->xk_bits = xk;
    return true;
}

// Index helpers shared by the kernel's CPU replay (tests) — the kernel has its own device versions.
inline unsigned ChainLin(unsigned long long idx, const uint16_t *col, int nbits)
{
    unsigned r = 0;
    for (int q = 0; q < nbits; q++)
        if ((idx >> q) & 1ull)
            r ^= col[q];
    return r;
}
inline unsigned long long ChainDeposit(unsigned long long idx, const uint8_t *bit, int nbits)
{
    unsigned long long r = 0;
    for (int q = 0; q < nbits; q++)
        r |= ((idx >> q) & 1ull) << bit[q];
    return r;
}

} // namespace jb
