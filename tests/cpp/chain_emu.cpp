// TEST INFRASTRUCTURE (not product code): CPU replay of the fused-chain kernel's parameter block.
//
// Plans a chain with the product's own planner (jet_b200/csrc/chain_plan.h) and then interprets
// the resulting ChainParams exactly the way ChainKernel (jet_b200/csrc/chain.cu) does — same tile
// loop, same thread/group enumeration, same in-place updates — but one "thread" at a time on the
// host, in complex<double>.  It also checks what a GPU cannot tell us cheaply: (i) no thread reads
// a shared-memory word another thread writes in the same phase (the in-place update is race
// free), (ii) the worst shared-memory bank-conflict degree of every phase.
// Built on the fly by tests/test_chain_plan.py with g++; never linked into libjetb200.so.
#include <complex>
#include <cstdio>
#include <map>
#include <vector>

#include "chain_plan.h"

using namespace jb;
using cd = std::complex<double>;

namespace {
constexpr int kMemLog = kChainMemLogLanes;
constexpr int kMemThreads = 1 << kMemLog;

// worst number of distinct addresses that fall into one bank group within a half/quarter warp
int ConflictDegree(const std::vector<unsigned> &addr_by_lane, int elem_bytes)
{
    const int lanes_per_phase = elem_bytes == 8 ? 16 : 8;
    const int banks = elem_bytes == 8 ? 16 : 8; // in units of one element
    int worst = 1;
    for (size_t l0 = 0; l0 < addr_by_lane.size(); l0 += lanes_per_phase) {
        std::map<unsigned, std::vector<unsigned>> per_bank;
        for (size_t l = l0; l < std::min(addr_by_lane.size(), l0 + lanes_per_phase); l++) {
            if (addr_by_lane[l] == 0xffffffffu)
                continue;
            auto &v = per_bank[addr_by_lane[l] % banks];
            if (std::find(v.begin(), v.end(), addr_by_lane[l]) == v.end())
                v.push_back(addr_by_lane[l]);
        }
        for (auto &kv : per_bank)
            worst = std::max<int>(worst, static_cast<int>(kv.second.size()));
    }
    return worst;
}
} // namespace

static int g_extra_quiet = 0;
// tile padding with untouched bits (what MakeChainOp asks for on short chains): applies to later runs
extern "C" void chain_emu_set_extra_quiet(int n) { g_extra_quiet = n; }

extern "C" int chain_emu_run(int elem_bytes, int n_x0_bits, const int *x0_bits, int n_steps,
                             const int *r_nbits, const int *r_bits, const int *x_is_left,
                             int max_tile_bits, int lane_bits, const double *x0,
                             const double *const *r, double *out, int *n_out_bits, int *out_bits,
                             int *stats /* [8] */)
{
    ChainSpec spec;
    spec.elem_bytes = elem_bytes;
    spec.x0_bits.assign(x0_bits, x0_bits + n_x0_bits);
    int off = 0;
    for (int s = 0; s < n_steps; s++) {
        ChainStepSpec st;
        st.r_bits.assign(r_bits + off, r_bits + off + r_nbits[s]);
        off += r_nbits[s];
        st.x_is_left = x_is_left[s] != 0;
        spec.steps.push_back(st);
    }
    ChainLayout lay;
    std::string why;
    if (!PlanChain(spec, max_tile_bits, lane_bits, &lay, &why, g_extra_quiet)) {
        std::fprintf(stderr, "chain_emu: %s\n", why.c_str());
        return 1;
    }
    const ChainParams &p = lay.params;
    const int kLogThreads = p.log_threads;
    const int kThreads = 1 << kLogThreads;
    if (kLogThreads != ChainLogThreads(elem_bytes))
        return 2;
    *n_out_bits = static_cast<int>(lay.xk_bits.size());
    for (size_t q = 0; q < lay.xk_bits.size(); q++)
        out_bits[q] = lay.xk_bits[q];
    int hazards = 0, load_conf = 1, step_conf = 1, store_conf = 1;

    std::vector<cd> tile(size_t(1) << p.log_tile), Bm(p.resident_elems);
    for (int s = 0; s < n_steps; s++) {
        const ChainStepParams &q = p.step[s];
        const int np = q.np, N = 1 << q.log_n, total = np << q.log_k;
        for (int e = 0; e < total; e++) {
            const unsigned k = e / np, n = e % np;
            cd v = 0;
            if (static_cast<int>(n) < N) {
                const unsigned long long a = ChainDeposit(k, q.rk, q.log_k) | ChainDeposit(n, q.rn, q.log_n);
                v = cd(r[s][2 * a], r[s][2 * a + 1]);
            }
            Bm[q.b_off + e] = v;
        }
    }
    // memory warps: 2^kMemLog lanes, the index bits above come from a per-CTA table; complex64 store
    // threads own two X_k-adjacent elements (store-index bit 0)
    const int pair = (elem_bytes == 8 && p.log_tile_out >= 1) ? 1 : 0;
    const int st_bits = p.log_tile_out - pair;
    const int in_lane_bits = std::min(p.log_tile_in, kMemLog);
    const int out_lane_bits = std::min(st_bits, kMemLog);
    const int in_iters = p.log_tile_in > kMemLog ? 1 << (p.log_tile_in - kMemLog) : 1;
    const int out_iters = st_bits > kMemLog ? 1 << (st_bits - kMemLog) : 1;
    const cd poison(1e300, 1e300);

    for (long long t = 0; t < p.n_tiles; t++) {
        const unsigned long long in_base = ChainDeposit(t, p.outer_in, p.log_outer);
        const unsigned long long out_base = ChainDeposit(t, p.outer_out, p.log_outer);
        std::fill(tile.begin(), tile.end(), poison);
        // load
        for (int j = 0; j < in_iters; j++) {
            for (int w = 0; w < kMemThreads / 32; w++) {
                std::vector<unsigned> lanes(32, 0xffffffffu);
                for (int l = 0; l < 32; l++) {
                    const int lt = w * 32 + l;
                    if (lt >= (1 << p.log_tile_in))
                        continue;
                    const int ib = std::max(0, p.log_tile_in - kMemLog);
                    const unsigned long long g = in_base | ChainDeposit(lt, p.in_gbit, in_lane_bits) |
                                                 ChainDeposit(j, p.in_gbit + kMemLog, ib);
                    const unsigned sa = ChainLin(lt, p.in_scol, in_lane_bits) ^ ChainLin(j, p.in_scol + kMemLog, ib);
                    if (tile[sa] != poison)
                        hazards++; // two lanes load the same tile element
                    tile[sa] = cd(x0[2 * g], x0[2 * g + 1]);
                    lanes[l] = sa;
                }
                if (t == 0)
                    load_conf = std::max(load_conf, ConflictDegree(lanes, elem_bytes));
            }
        }
        // stages
        for (int sg = 0; sg < p.n_stages; sg++) {
            const ChainStageParams &G = p.stage[sg];
            std::vector<int> reader(tile.size(), -1), writer(tile.size(), -1);
            std::vector<cd> next = tile; // writes land here; reads come from `tile` (= barrier semantics)
            if (G.kind == 1) {
                const int NL = elem_bytes == 8 ? kChainLocalBits : kChainLocalBits - 1;
                const int NE = 1 << NL;
                const int log_g = G.log_g;
                const int tid_bits = std::min(log_g, kLogThreads);
                const int per_thread = log_g > kLogThreads ? 1 << (log_g - kLogThreads) : 1;
                for (int j = 0; j < per_thread; j++) {
                    for (int w = 0; w < kThreads / 32; w++) {
                        std::vector<std::vector<unsigned>> acc_addr(NE, std::vector<unsigned>(32, 0xffffffffu));
                        for (int l = 0; l < 32; l++) {
                            const int tid = w * 32 + l;
                            if (tid >= (1 << log_g))
                                continue;
                            const unsigned base = ChainLin(tid, G.gcol, tid_bits) ^
                                                  p.stage_tab[sg][j];
                            std::vector<cd> E(NE);
                            std::vector<unsigned> ad(NE);
                            for (int e = 0; e < NE; e++) {
                                ad[e] = base ^ ChainLin(e, G.lcol, NL);
                                E[e] = tile[ad[e]];
                                acc_addr[e][l] = ad[e];
                                if (reader[ad[e]] >= 0)
                                    hazards++; // two threads own the same element
                                reader[ad[e]] = tid;
                            }
                            for (int t = 0; t < G.count; t++) {
                                const ChainStepParams &q = p.step[G.first + t];
                                const int mask = G.desc[t] & 0xff;
                                if ((G.desc[t] >> 8) != q.b_off)
                                    hazards += 1000;
                                std::vector<int> mpos;
                                for (int b = 0; b < NL; b++)
                                    if (mask & (1 << b))
                                        mpos.push_back(b);
                                const int K = 1 << mpos.size();
                                if (static_cast<int>(mpos.size()) != q.log_k || q.log_n > q.log_k ||
                                    q.np != std::max(4, K))
                                    hazards += 1000; // planner inconsistency
                                const int N = 1 << q.log_n; // N < K: the matrix has zero columns beyond N
                                auto spread = [&](int v) {
                                    int r = 0;
                                    for (size_t b = 0; b < mpos.size(); b++)
                                        if (v & (1 << b))
                                            r |= 1 << mpos[b];
                                    return r;
                                };
                                std::vector<cd> out(NE, cd(0));
                                for (int g = 0; g < NE; g++) {
                                    if (g & mask)
                                        continue;
                                    for (int n = 0; n < N; n++)
                                        for (int k = 0; k < K; k++)
                                            out[g | spread(n)] += E[g | spread(k)] * Bm[q.b_off + k * q.np + n];
                                }
                                E = out;
                            }
                            for (int e = 0; e < NE; e++) {
                                next[ad[e]] = E[e];
                                writer[ad[e]] = tid;
                            }
                        }
                        if (t == 0)
                            for (auto &v : acc_addr)
                                step_conf = std::max(step_conf, ConflictDegree(v, elem_bytes));
                    }
                }
                tile.swap(next);
                continue;
            }
            const int s = G.first;
            const ChainStepParams &q = p.step[s];
            const int K = 1 << q.log_k, N = 1 << q.log_n, np = q.np;
            const int log_g = q.log_g;
            const int tid_bits = std::min(log_g, kLogThreads);
            const int per_thread = log_g > kLogThreads ? 1 << (log_g - kLogThreads) : 1;
            for (int j = 0; j < per_thread; j++) {
                for (int w = 0; w < kThreads / 32; w++) {
                    std::vector<std::vector<unsigned>> rd(K, std::vector<unsigned>(32, 0xffffffffu));
                    std::vector<std::vector<unsigned>> wr(N, std::vector<unsigned>(32, 0xffffffffu));
                    for (int l = 0; l < 32; l++) {
                        const int tid = w * 32 + l;
                        if (tid >= (1 << log_g))
                            continue;
                        const unsigned base = ChainLin(tid, q.gcol, tid_bits) ^
                                              p.stage_tab[sg][j];
                        std::vector<cd> a(K);
                        for (int kk = 0; kk < K; kk++) {
                            const unsigned ad = base ^ ChainLin(kk, q.kcol, q.log_k);
                            a[kk] = tile[ad];
                            rd[kk][l] = ad;
                            if (writer[ad] >= 0 && writer[ad] != tid)
                                hazards++;
                            reader[ad] = tid;
                        }
                        for (int y = 0; y < N; y++) {
                            cd acc = 0;
                            for (int kk = 0; kk < K; kk++)
                                acc += a[kk] * Bm[q.b_off + kk * np + y];
                            const unsigned ad = base ^ ChainLin(y, q.ncol, q.log_n);
                            if ((reader[ad] >= 0 && reader[ad] != tid) || (writer[ad] >= 0))
                                hazards++;
                            writer[ad] = tid;
                            next[ad] = acc;
                            wr[y][l] = ad;
                        }
                    }
                    if (t == 0) {
                        for (auto &v : rd)
                            step_conf = std::max(step_conf, ConflictDegree(v, elem_bytes));
                        for (auto &v : wr)
                            step_conf = std::max(step_conf, ConflictDegree(v, elem_bytes));
                    }
                }
            }
            // a second pass for reads that came after a foreign write in emulation order
            for (size_t ad = 0; ad < tile.size(); ad++)
                if (reader[ad] >= 0 && writer[ad] >= 0 && reader[ad] != writer[ad])
                    hazards++;
            tile.swap(next);
        }
        // store
        for (int j = 0; j < out_iters; j++) {
            for (int w = 0; w < kMemThreads / 32; w++) {
              for (int half = 0; half <= pair; half++) {
                std::vector<unsigned> lanes(32, 0xffffffffu);
                for (int l = 0; l < 32; l++) {
                    const int lt = w * 32 + l;
                    if (lt >= (1 << st_bits))
                        continue;
                    const int ob = std::max(0, st_bits - kMemLog);
                    unsigned long long g = out_base | ChainDeposit(lt, p.out_gbit + pair, out_lane_bits) |
                                           ChainDeposit(j, p.out_gbit + kMemLog + pair, ob);
                    unsigned sa = ChainLin(lt, p.out_scol + pair, out_lane_bits) ^
                                  ChainLin(j, p.out_scol + kMemLog + pair, ob);
                    if (half) {
                        if (p.out_gbit[0] != 0)
                            hazards += 1000; // the pair must be X_k-adjacent
                        g |= 1ull;
                        sa ^= p.out_scol[0];
                    }
                    out[2 * g] = tile[sa].real();
                    out[2 * g + 1] = tile[sa].imag();
                    lanes[l] = sa;
                }
                if (t == 0)
                    store_conf = std::max(store_conf, ConflictDegree(lanes, elem_bytes));
              }
            }
        }
    }
    stats[0] = p.log_tile;
    stats[1] = lay.conflict_free;
    stats[2] = hazards;
    stats[3] = load_conf;
    stats[4] = step_conf;
    stats[5] = store_conf;
    stats[6] = p.log_outer;
    stats[7] = p.n_stages;
    return 0;
}
