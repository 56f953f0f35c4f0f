// pathopt — contraction-path and slice search for sliced tensor-network contraction (host-side bookkeeping
// only: index sets and extents, no tensor data).
//
// The reference ships its benchmark paths pre-computed with cotengra (hyper-optimised KaHyPar bisection +
// SliceFinder, /root/reference/examples/paper_benchmarks/GPU/cot_gpu_m12/run_sliced.py:37-53) and carries a
// random-sampling finder in python/jet/interpreter.py:533-618.  Neither library is available offline, and a
// greedy path puts the m=20 Sycamore amplitude out of reach (1e24 flops, 2^58 slices), so this tool restates the
// published recipe from scratch:
//   1. initial contraction tree by recursive graph bisection (region growing + Fiduccia-Mattheyses refinement,
//      random imbalance per level), exact dynamic programming below a size threshold;
//   2. subtree reconfiguration: a connected piece of the tree with <= N frontier nodes is re-contracted
//      optimally (DP over subsets), sweeping over the tree until nothing improves;
//   3. slicing interleaved with reconfiguration: while the largest intermediate exceeds the target, slice the
//      index that minimises the total cost, then reconfigure with the sliced extents;
//   4. final sweeps under a hard size cap, and removal of slices that are no longer needed.
// Costs follow PathInfo (include/jet/PathInfo.hpp:157-183): flops of a step = 2 * M*N*K.
// Paths are emitted in the reference's format: pairs of node ids, step i creating node num_leaves + i
// (include/jet/TensorNetwork.hpp:301-328), in depth-first order, larger operand first.
//
// Input (stdin or file): "n_leaves n_indices", then n_indices lines "name dim", then one line per leaf
// "k i1 .. ik" (index numbers).  Output: one JSON object on stdout.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int kMaxWords = 40; // up to 2560 index bits
int g_words = 1;

struct BS {
    uint64_t w[kMaxWords];
    BS() { std::memset(w, 0, sizeof(w)); }
    void set(int i) { w[i >> 6] |= uint64_t(1) << (i & 63); }
    bool get(int i) const { return (w[i >> 6] >> (i & 63)) & 1; }
};
inline int PopXor(const BS &a, const BS &b, const BS &keep)
{
    int c = 0;
    for (int i = 0; i < g_words; i++)
        c += __builtin_popcountll((a.w[i] ^ b.w[i]) & keep.w[i]);
    return c;
}
inline int PopOr(const BS &a, const BS &b, const BS &keep)
{
    int c = 0;
    for (int i = 0; i < g_words; i++)
        c += __builtin_popcountll((a.w[i] | b.w[i]) & keep.w[i]);
    return c;
}
inline int Pop(const BS &a, const BS &keep)
{
    int c = 0;
    for (int i = 0; i < g_words; i++)
        c += __builtin_popcountll(a.w[i] & keep.w[i]);
    return c;
}
inline BS Xor(const BS &a, const BS &b)
{
    BS r;
    for (int i = 0; i < g_words; i++)
        r.w[i] = a.w[i] ^ b.w[i];
    return r;
}

// Every index of extent 2^k is represented by k bits, so sizes are 2^popcount.
struct Net {
    int n = 0;                        // leaves
    int bits = 0;                     // index bits
    std::vector<BS> leaf;             // bits of each leaf
    std::vector<int> bit_index;       // bit -> index number
    std::vector<std::string> names;   // index names
    std::vector<int> dim_bits;        // index -> bits
    std::vector<std::vector<int>> adj; // leaf graph: neighbours
    std::vector<std::vector<int>> adj_w;
};

struct Tree {
    int n = 0;
    std::vector<int> l, r, p;
    std::vector<BS> legs;
    int root = -1;
};

double Exp2(int e) { return std::ldexp(1.0, e); }

struct Cost {
    double flops = 0; // sum over steps of M*N*K (multiply by 2 for the Jet convention)
    int width = 0;    // log2 of the largest intermediate
    double write = 0; // sum of intermediate sizes
};

Cost TreeCost(const Tree &t, const BS &keep)
{
    Cost c;
    for (int v = t.n; v < 2 * t.n - 1; v++) {
        c.flops += Exp2(PopOr(t.legs[t.l[v]], t.legs[t.r[v]], keep));
        const int s = Pop(t.legs[v], keep);
        c.width = std::max(c.width, s);
        c.write += Exp2(s);
    }
    return c;
}

// ---- exact DP over subsets of k frontier tensors ----------------------------------------------------------------
struct DpResult {
    double cost = 0;
    std::vector<std::pair<int, int>> merges; // in terms of slot ids: 0..k-1 inputs, k.. created
};

// minimises sum of 2^|a u b| over the binary tree; `cap` (log2 elements) bounds every intermediate except the
// full set (whose size is fixed); returns false if no tree satisfies the cap
bool DpOptimal(const std::vector<BS> &in, const BS &keep, int cap, double size_weight, DpResult *out)
{
    const int k = static_cast<int>(in.size());
    const int full = (1 << k) - 1;
    static thread_local std::vector<BS> lg;
    static thread_local std::vector<double> best;
    static thread_local std::vector<int> split;
    static thread_local std::vector<int> pc;
    lg.resize(full + 1);
    best.assign(full + 1, 1e300);
    split.assign(full + 1, 0);
    pc.resize(full + 1);
    for (int m = 1; m <= full; m++) {
        const int low = __builtin_ctz(m);
        const int rest = m & (m - 1);
        if (rest == 0) {
            for (int i = 0; i < g_words; i++)
                lg[m].w[i] = in[low].w[i] & keep.w[i];
            best[m] = 0;
        }
        else {
            for (int i = 0; i < g_words; i++)
                lg[m].w[i] = lg[rest].w[i] ^ lg[1 << low].w[i];
        }
        int c = 0;
        for (int i = 0; i < g_words; i++)
            c += __builtin_popcountll(lg[m].w[i]);
        pc[m] = c;
    }
    // masks in increasing numeric order: every proper submask is smaller
    for (int m = 3; m <= full; m++) {
        if ((m & (m - 1)) == 0)
            continue;
        if (m != full && pc[m] > cap)
            continue; // this intermediate is not allowed
        const int top = 31 - __builtin_clz(m);
        const int topbit = 1 << top;
        double b = 1e300;
        int bs = 0;
        // enumerate submasks containing the top bit (each unordered split once)
        const int rest = m ^ topbit;
        for (int s = rest;; s = (s - 1) & rest) {
            const int a = s | topbit, o = m ^ a;
            if (o != 0 && best[a] < 1e299 && best[o] < 1e299) {
                int u = 0;
                for (int i = 0; i < g_words; i++)
                    u += __builtin_popcountll(lg[a].w[i] | lg[o].w[i]);
                const double c = best[a] + best[o] + Exp2(u);
                if (c < b) {
                    b = c;
                    bs = a;
                }
            }
            if (s == 0)
                break;
        }
        if (b < 1e299)
            b += size_weight * Exp2(pc[m]);
        best[m] = b;
        split[m] = bs;
    }
    if (best[full] >= 1e299)
        return false;
    out->cost = best[full];
    out->merges.clear();
    // emit merges bottom-up
    std::vector<int> slot_of(full + 1, -1);
    for (int i = 0; i < k; i++)
        slot_of[1 << i] = i;
    int next = k;
    // recursive emit
    struct Frame {
        int m;
        int stage;
    };
    std::vector<Frame> st;
    st.push_back({full, 0});
    while (!st.empty()) {
        Frame &f = st.back();
        if (slot_of[f.m] >= 0) {
            st.pop_back();
            continue;
        }
        const int a = split[f.m], o = f.m ^ a;
        if (f.stage == 0) {
            f.stage = 1;
            st.push_back({a, 0});
        }
        else if (f.stage == 1) {
            f.stage = 2;
            st.push_back({o, 0});
        }
        else {
            out->merges.emplace_back(slot_of[a], slot_of[o]);
            slot_of[f.m] = next++;
            st.pop_back();
        }
    }
    return true;
}

// ---- initial tree: recursive bisection -------------------------------------------------------------------------
struct Builder {
    const Net &net;
    std::mt19937_64 rng;
    Tree tree;
    int next_internal;
    int dp_leaves;
    double imbalance;

    Builder(const Net &n, uint64_t seed, int dp, double imb) : net(n), rng(seed), dp_leaves(dp), imbalance(imb)
    {
        tree.n = net.n;
        const int total = 2 * net.n - 1;
        tree.l.assign(total, -1);
        tree.r.assign(total, -1);
        tree.p.assign(total, -1);
        tree.legs.resize(total);
        for (int i = 0; i < net.n; i++)
            tree.legs[i] = net.leaf[i];
        next_internal = net.n;
    }

    int Join(int a, int b)
    {
        const int c = next_internal++;
        tree.l[c] = a;
        tree.r[c] = b;
        tree.p[a] = c;
        tree.p[b] = c;
        tree.legs[c] = Xor(tree.legs[a], tree.legs[b]);
        return c;
    }

    int BuildDp(const std::vector<int> &nodes, const BS &keep)
    {
        if (nodes.size() == 1)
            return nodes[0];
        std::vector<BS> in;
        for (int v : nodes)
            in.push_back(tree.legs[v]);
        DpResult res;
        DpOptimal(in, keep, 1 << 20, 0.0, &res);
        std::vector<int> slot(nodes);
        for (auto [a, b] : res.merges)
            slot.push_back(Join(slot[a], slot[b]));
        return slot.back();
    }

    // bisect `verts` (leaf ids) minimising the weight of cut edges; returns side flags
    std::vector<char> Bisect(const std::vector<int> &verts)
    {
        const int m = static_cast<int>(verts.size());
        std::vector<int> local(net.n, -1);
        for (int i = 0; i < m; i++)
            local[verts[i]] = i;
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::vector<char> best_side;
        long long best_cut = -1;
        const int tries = m > 200 ? 3 : 5;
        for (int t = 0; t < tries; t++) {
            const double frac = 0.5 + (U(rng) * 2 - 1) * imbalance;
            const int target = std::max(1, std::min(m - 1, static_cast<int>(std::lround(frac * m))));
            // region growing from a random vertex (BFS with random tie order)
            std::vector<char> side(m, 0);
            std::vector<int> order;
            std::vector<char> seen(m, 0);
            int grown = 0;
            std::vector<int> queue;
            while (grown < target) {
                if (queue.empty()) {
                    int s;
                    do {
                        s = static_cast<int>(rng() % m);
                    } while (seen[s]);
                    seen[s] = 1;
                    queue.push_back(s);
                }
                const size_t qi = rng() % queue.size() < queue.size() / 2 + 1 ? 0 : rng() % queue.size();
                const int v = queue[qi];
                queue.erase(queue.begin() + static_cast<long>(qi));
                side[v] = 1;
                grown++;
                for (int nb : net.adj[verts[v]]) {
                    const int u = local[nb];
                    if (u >= 0 && !seen[u]) {
                        seen[u] = 1;
                        queue.push_back(u);
                    }
                }
            }
            // FM refinement
            const int lo = std::max(1, static_cast<int>(std::floor((0.5 - imbalance) * m)));
            const int hi = std::min(m - 1, static_cast<int>(std::ceil((0.5 + imbalance) * m)));
            auto cut_of = [&]() {
                long long c = 0;
                for (int v = 0; v < m; v++)
                    for (size_t e = 0; e < net.adj[verts[v]].size(); e++) {
                        const int u = local[net.adj[verts[v]][e]];
                        if (u > v && side[u] != side[v])
                            c += net.adj_w[verts[v]][e];
                    }
                return c;
            };
            long long cut = cut_of();
            for (int pass = 0; pass < 12; pass++) {
                std::vector<int> gain(m, 0);
                for (int v = 0; v < m; v++)
                    for (size_t e = 0; e < net.adj[verts[v]].size(); e++) {
                        const int u = local[net.adj[verts[v]][e]];
                        if (u >= 0)
                            gain[v] += side[u] != side[v] ? net.adj_w[verts[v]][e] : -net.adj_w[verts[v]][e];
                    }
                std::vector<char> locked(m, 0);
                std::vector<int> moves;
                long long cur = cut, best_here = cut;
                int best_prefix = 0, ones = 0;
                for (int v = 0; v < m; v++)
                    ones += side[v];
                for (int step = 0; step < m; step++) {
                    int pick = -1, pg = -(1 << 30);
                    for (int v = 0; v < m; v++) {
                        if (locked[v])
                            continue;
                        const int new_ones = ones + (side[v] ? -1 : 1);
                        if (new_ones < lo || new_ones > hi)
                            continue;
                        if (gain[v] > pg || (gain[v] == pg && (rng() & 1))) {
                            pg = gain[v];
                            pick = v;
                        }
                    }
                    if (pick < 0)
                        break;
                    locked[pick] = 1;
                    ones += side[pick] ? -1 : 1;
                    side[pick] ^= 1;
                    cur -= pg;
                    moves.push_back(pick);
                    for (size_t e = 0; e < net.adj[verts[pick]].size(); e++) {
                        const int u = local[net.adj[verts[pick]][e]];
                        if (u < 0)
                            continue;
                        const int w = net.adj_w[verts[pick]][e];
                        // after the move: if u is on the other side now, the edge is cut
                        gain[u] += side[u] != side[pick] ? 2 * w : -2 * w;
                    }
                    gain[pick] = -pg;
                    if (cur < best_here) {
                        best_here = cur;
                        best_prefix = static_cast<int>(moves.size());
                    }
                    if (static_cast<int>(moves.size()) - best_prefix > 60)
                        break; // no improvement for a while
                }
                for (int i = static_cast<int>(moves.size()) - 1; i >= best_prefix; i--)
                    side[moves[i]] ^= 1;
                if (best_here >= cut)
                    break;
                cut = best_here;
            }
            if (best_cut < 0 || cut < best_cut) {
                best_cut = cut;
                best_side = side;
            }
        }
        return best_side;
    }

    int Build(const std::vector<int> &verts, const BS &keep)
    {
        if (static_cast<int>(verts.size()) <= dp_leaves)
            return BuildDp(verts, keep);
        const std::vector<char> side = Bisect(verts);
        std::vector<int> a, b;
        for (size_t i = 0; i < verts.size(); i++)
            (side[i] ? a : b).push_back(verts[i]);
        if (a.empty() || b.empty()) { // degenerate (disconnected singletons): split evenly
            a.assign(verts.begin(), verts.begin() + static_cast<long>(verts.size() / 2));
            b.assign(verts.begin() + static_cast<long>(verts.size() / 2), verts.end());
        }
        const int ra = Build(a, keep), rb = Build(b, keep);
        return Join(ra, rb);
    }
};

// ---- subtree reconfiguration ---------------------------------------------------------------------------------
// Re-contracts the piece of the tree hanging below `top` whose frontier has up to `k` nodes.  Returns the change
// in (flops + size_weight * write) (<= 0).
double Reconfigure(Tree &t, int top, int k, const BS &keep, int cap, double size_weight, std::mt19937_64 &rng,
                   bool by_size)
{
    if (top < t.n)
        return 0;
    std::vector<int> frontier = {t.l[top], t.r[top]};
    std::vector<int> internal = {top};
    while (static_cast<int>(frontier.size()) < k) {
        // expand the largest (or a random) internal frontier node
        int pick = -1, ps = -1;
        for (size_t i = 0; i < frontier.size(); i++) {
            if (frontier[i] < t.n)
                continue;
            const int s = by_size ? Pop(t.legs[frontier[i]], keep) * 4 + static_cast<int>(rng() & 3)
                                  : static_cast<int>(rng() & 0xffff);
            if (s > ps) {
                ps = s;
                pick = static_cast<int>(i);
            }
        }
        if (pick < 0)
            break;
        const int v = frontier[pick];
        frontier[pick] = t.l[v];
        frontier.push_back(t.r[v]);
        internal.push_back(v);
    }
    if (frontier.size() < 3)
        return 0;
    double old_cost = 0;
    for (int v : internal) {
        old_cost += Exp2(PopOr(t.legs[t.l[v]], t.legs[t.r[v]], keep));
        if (v != top)
            old_cost += size_weight * Exp2(Pop(t.legs[v], keep));
    }
    old_cost += size_weight * Exp2(Pop(t.legs[top], keep));
    std::vector<BS> in;
    for (int v : frontier)
        in.push_back(t.legs[v]);
    DpResult res;
    if (!DpOptimal(in, keep, cap, size_weight, &res))
        return 0;
    if (res.cost >= old_cost * (1 - 1e-12))
        return 0;
    // rebuild using the same internal ids; `top` stays the root of the piece
    std::vector<int> ids(internal.begin() + 1, internal.end());
    ids.push_back(top);
    std::vector<int> slot(frontier);
    size_t used = 0;
    for (auto [a, b] : res.merges) {
        const int c = ids[used++];
        const int x = slot[a], y = slot[b];
        t.l[c] = x;
        t.r[c] = y;
        t.p[x] = c;
        t.p[y] = c;
        t.legs[c] = Xor(t.legs[x], t.legs[y]);
        slot.push_back(c);
    }
    return res.cost - old_cost;
}

struct Options {
    int target = 30;        // log2 of the largest allowed intermediate (elements)
    int max_slices = 40;    // at most 2^max_slices slices
    int trials = 16;
    int threads = 8;
    double seconds = 120;
    uint64_t seed = 1;
    int reconf_k = 10;
    std::vector<std::string> fixed_slices; // always sliced (first in the output)
};

struct Solution {
    Tree tree;
    BS sliced;
    std::vector<int> slice_order; // bit ids in the order they were chosen
    double total = 1e300;         // 2 * flops per slice * slices
    Cost per_slice;
};

void Sweep(Tree &t, const BS &keep, int cap, int k, double size_weight, std::mt19937_64 &rng, int rounds,
           double deadline_s, const std::chrono::steady_clock::time_point &t0)
{
    std::vector<int> order;
    for (int v = t.n; v < 2 * t.n - 1; v++)
        order.push_back(v);
    for (int r = 0; r < rounds; r++) {
        std::shuffle(order.begin(), order.end(), rng);
        double gain = 0;
        for (int v : order) {
            gain += Reconfigure(t, v, k, keep, cap, size_weight, rng, (rng() & 3) != 0);
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > deadline_s)
                return;
        }
        if (gain > -1e-9)
            break;
    }
}

BS KeepMask(const Net &net, const BS &sliced)
{
    BS keep;
    for (int b = 0; b < net.bits; b++)
        if (!sliced.get(b))
            keep.set(b);
    return keep;
}

Solution Solve(const Net &net, const Options &opt, uint64_t seed, const std::vector<int> &fixed_bits, double budget_s)
{
    const auto t0 = std::chrono::steady_clock::now();
    std::mt19937_64 rng(seed * 0x9E3779B97F4A7C15ull + 12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    Solution sol;
    for (int b : fixed_bits) {
        sol.sliced.set(b);
        sol.slice_order.push_back(b);
    }
    BS keep = KeepMask(net, sol.sliced);
    const double imbalance = 0.02 + 0.25 * U(rng);
    Builder builder(net, rng(), 8, imbalance);
    std::vector<int> all(net.n);
    for (int i = 0; i < net.n; i++)
        all[i] = i;
    builder.tree.root = builder.Build(all, keep);
    Tree &t = builder.tree;
    const int k = opt.reconf_k;
    // phase A: flops only, no cap
    Sweep(t, keep, 1 << 20, std::min(k, 8), 0.0, rng, 3, budget_s * 0.25, t0);
    Sweep(t, keep, 1 << 20, k, 0.0, rng, 4, budget_s * 0.45, t0);
    // phase B: slice until the width fits, reconfiguring as we go
    Cost c = TreeCost(t, keep);
    while (c.width > opt.target && static_cast<int>(sol.slice_order.size()) < opt.max_slices) {
        // candidates: bits on the largest intermediates
        BS cand;
        for (int v = t.n; v < 2 * t.n - 1; v++)
            if (Pop(t.legs[v], keep) >= c.width - 1)
                for (int i = 0; i < g_words; i++)
                    cand.w[i] |= t.legs[v].w[i] & keep.w[i];
        int best_bit = -1;
        double best_total = 1e300;
        int best_width = 1 << 20;
        for (int b = 0; b < net.bits; b++) {
            if (!cand.get(b))
                continue;
            BS k2 = keep;
            k2.w[b >> 6] &= ~(uint64_t(1) << (b & 63));
            const Cost c2 = TreeCost(t, k2);
            // prefer lower flops; ties by width
            if (c2.flops < best_total * (1 - 1e-12) || (c2.flops <= best_total * (1 + 1e-12) && c2.width < best_width)) {
                best_total = c2.flops;
                best_width = c2.width;
                best_bit = b;
            }
        }
        if (best_bit < 0)
            break;
        sol.sliced.set(best_bit);
        sol.slice_order.push_back(best_bit);
        keep = KeepMask(net, sol.sliced);
        // soft pressure on size while the cap is not met: weight the written elements
        Sweep(t, keep, 1 << 20, std::min(k, 8), 0.0, rng, 1, budget_s * 0.75, t0);
        c = TreeCost(t, keep);
    }
    // phase C: hard cap, full-size subtrees
    if (c.width <= opt.target) {
        Sweep(t, keep, opt.target, k, 0.0, rng, 6, budget_s * 0.95, t0);
        // phase D: drop slices that are not needed any more (most recent first)
        for (int i = static_cast<int>(sol.slice_order.size()) - 1; i >= static_cast<int>(fixed_bits.size()); i--) {
            BS s2 = sol.sliced;
            const int b = sol.slice_order[i];
            s2.w[b >> 6] &= ~(uint64_t(1) << (b & 63));
            const BS k2 = KeepMask(net, s2);
            if (TreeCost(t, k2).width <= opt.target) {
                sol.sliced = s2;
                sol.slice_order.erase(sol.slice_order.begin() + i);
                keep = k2;
            }
        }
        Sweep(t, keep, opt.target, k, 0.0, rng, 2, budget_s, t0);
    }
    sol.per_slice = TreeCost(t, keep);
    sol.total = sol.per_slice.width <= opt.target
                    ? 2.0 * sol.per_slice.flops * Exp2(static_cast<int>(sol.slice_order.size()))
                    : 1e300;
    sol.tree = t;
    return sol;
}

// depth-first emission, larger subtree first (keeps few large tensors alive)
void EmitPath(const Tree &t, const BS &keep, std::vector<std::pair<int, int>> *path)
{
    std::vector<int> new_id(2 * t.n - 1, -1);
    for (int i = 0; i < t.n; i++)
        new_id[i] = i;
    int next = t.n;
    struct Frame {
        int v;
        int stage;
    };
    std::vector<Frame> st = {{t.root, 0}};
    while (!st.empty()) {
        Frame &f = st.back();
        if (f.v < t.n) {
            st.pop_back();
            continue;
        }
        int a = t.l[f.v], b = t.r[f.v];
        if (Pop(t.legs[a], keep) < Pop(t.legs[b], keep))
            std::swap(a, b);
        if (f.stage == 0) {
            f.stage = 1;
            st.push_back({a, 0});
        }
        else if (f.stage == 1) {
            f.stage = 2;
            st.push_back({b, 0});
        }
        else {
            path->emplace_back(new_id[a], new_id[b]);
            new_id[f.v] = next++;
            st.pop_back();
        }
    }
}

} // namespace

int main(int argc, char **argv)
{
    Options opt;
    std::string input;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
        if (a == "--target")
            opt.target = std::stoi(next());
        else if (a == "--max-slices")
            opt.max_slices = std::stoi(next());
        else if (a == "--trials")
            opt.trials = std::stoi(next());
        else if (a == "--threads")
            opt.threads = std::stoi(next());
        else if (a == "--seconds")
            opt.seconds = std::stod(next());
        else if (a == "--seed")
            opt.seed = std::stoull(next());
        else if (a == "--k")
            opt.reconf_k = std::stoi(next());
        else if (a == "--slice")
            opt.fixed_slices.push_back(next());
        else
            input = a;
    }
    std::ifstream fin;
    if (!input.empty())
        fin.open(input);
    std::istream &in = input.empty() ? std::cin : fin;
    int n_leaves = 0, n_indices = 0;
    if (!(in >> n_leaves >> n_indices) || n_leaves < 1) {
        std::cerr << "pathopt: bad input header" << std::endl;
        return 2;
    }
    Net net;
    net.n = n_leaves;
    std::vector<int> first_bit(n_indices);
    for (int i = 0; i < n_indices; i++) {
        std::string name;
        long long dim;
        in >> name >> dim;
        int b = 0;
        while ((1ll << b) < dim)
            b++;
        if ((1ll << b) != dim) {
            std::cerr << "pathopt: extent " << dim << " of index " << name << " is not a power of two" << std::endl;
            return 3;
        }
        net.names.push_back(name);
        net.dim_bits.push_back(b);
        first_bit[i] = net.bits;
        for (int j = 0; j < b; j++)
            net.bit_index.push_back(i);
        net.bits += b;
    }
    g_words = (net.bits + 63) / 64;
    if (g_words > kMaxWords) {
        std::cerr << "pathopt: too many indices" << std::endl;
        return 3;
    }
    net.leaf.resize(n_leaves);
    std::vector<std::vector<int>> owners(n_indices);
    for (int v = 0; v < n_leaves; v++) {
        int k;
        in >> k;
        for (int j = 0; j < k; j++) {
            int idx;
            in >> idx;
            for (int b = 0; b < net.dim_bits[idx]; b++)
                net.leaf[v].set(first_bit[idx] + b);
            owners[idx].push_back(v);
        }
    }
    net.adj.resize(n_leaves);
    net.adj_w.resize(n_leaves);
    for (int i = 0; i < n_indices; i++) {
        if (owners[i].size() > 2) {
            std::cerr << "pathopt: index " << net.names[i] << " appears on more than two tensors" << std::endl;
            return 3;
        }
        if (owners[i].size() != 2 || owners[i][0] == owners[i][1])
            continue;
        for (int s = 0; s < 2; s++) {
            const int a = owners[i][s], b = owners[i][1 - s];
            auto it = std::find(net.adj[a].begin(), net.adj[a].end(), b);
            if (it == net.adj[a].end()) {
                net.adj[a].push_back(b);
                net.adj_w[a].push_back(net.dim_bits[i]);
            }
            else {
                net.adj_w[a][it - net.adj[a].begin()] += net.dim_bits[i];
            }
        }
    }
    std::vector<int> fixed_bits;
    for (const auto &name : opt.fixed_slices) {
        const auto it = std::find(net.names.begin(), net.names.end(), name);
        if (it == net.names.end()) {
            std::cerr << "pathopt: unknown index " << name << std::endl;
            return 3;
        }
        const int idx = static_cast<int>(it - net.names.begin());
        for (int b = 0; b < net.dim_bits[idx]; b++)
            fixed_bits.push_back(first_bit[idx] + b);
    }

    const auto t_start = std::chrono::steady_clock::now();
    std::mutex mu;
    Solution best;
    std::atomic<int> next_trial{0};
    std::vector<std::pair<double, double>> history;
    const int threads = std::max(1, std::min(opt.threads, opt.trials));
    const double per_trial = opt.seconds * threads / std::max(1, opt.trials);
    std::vector<std::thread> pool;
    for (int th = 0; th < threads; th++)
        pool.emplace_back([&]() {
            for (;;) {
                const int trial = next_trial++;
                if (trial >= opt.trials)
                    return;
                Solution s = Solve(net, opt, opt.seed * 1000 + trial, fixed_bits, per_trial);
                std::lock_guard<std::mutex> lk(mu);
                history.emplace_back(s.total, static_cast<double>(s.slice_order.size()));
                if (s.total < best.total)
                    best = std::move(s);
            }
        });
    for (auto &th : pool)
        th.join();
    if (best.total >= 1e299) {
        std::cerr << "pathopt: no trial met the target width " << opt.target << " within 2^" << opt.max_slices
                  << " slices" << std::endl;
        return 4;
    }
    const BS keep = KeepMask(net, best.sliced);
    std::vector<std::pair<int, int>> path;
    EmitPath(best.tree, keep, &path);
    // sliced indices by name (an index of extent 2^k is sliced as a whole when any of its bits is)
    std::vector<std::string> sliced_names;
    std::vector<char> seen(n_indices, 0);
    for (int b : best.slice_order) {
        const int idx = net.bit_index[b];
        if (!seen[idx]) {
            seen[idx] = 1;
            sliced_names.push_back(net.names[idx]);
        }
    }
    BS none;
    const BS keep_all = KeepMask(net, none);
    const Cost unsliced = TreeCost(best.tree, keep_all);
    std::ostringstream os;
    os.precision(17);
    os << "{\"path\": [";
    for (size_t i = 0; i < path.size(); i++)
        os << (i ? "," : "") << "[" << path[i].first << "," << path[i].second << "]";
    os << "], \"sliced\": [";
    for (size_t i = 0; i < sliced_names.size(); i++)
        os << (i ? "," : "") << "\"" << sliced_names[i] << "\"";
    os << "], \"log2_slices\": " << best.slice_order.size() << ", \"log2_peak_per_slice\": " << best.per_slice.width
       << ", \"jet_flops_per_slice\": " << 2.0 * best.per_slice.flops << ", \"jet_flops_total\": " << best.total
       << ", \"log2_peak_unsliced\": " << unsliced.width << ", \"jet_flops_unsliced\": " << 2.0 * unsliced.flops
       << ", \"trials\": " << history.size() << ", \"seconds\": "
       << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() << ", \"trial_totals\": [";
    for (size_t i = 0; i < history.size(); i++)
        os << (i ? "," : "") << history[i].first;
    os << "]}";
    std::cout << os.str() << std::endl;
    return 0;
}
