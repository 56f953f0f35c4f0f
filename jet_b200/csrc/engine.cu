// Contraction plan: a whole (sliced) tensor network + path resident on one GPU.
//
// Replaces the execution half of TaskBasedContractor (reference:
// include/jet/TaskBasedContractor.hpp:162-322): instead of one host task per contraction that
// allocates, zero-fills and page-faults a fresh std::vector (include/jet/Tensor.hpp:58-60,743),
//   * leaves are uploaded once into a device arena;
//   * TensorNetwork::SliceIndices (include/jet/TensorNetwork.hpp:210-284) becomes a device-side
//     view: a tiny kernel gathers the sliced leaves for the current slice id, so 2^s slices never
//     exist as 2^s host copies of the network (examples/paper_benchmarks/CPU/jet_cpu_m10/
//     jet_sliced.cpp:69-75);
//   * slice-independent steps (the reference's de-duplication by task name,
//     TaskBasedContractor.hpp:216-222) run once; the per-slice steps are captured into ONE CUDA
//     graph that is re-launched per slice with no host synchronisation in between;
//   * intermediates get arena offsets from their lifetimes at plan time (AddDeletionTasks,
//     TaskBasedContractor.hpp:290-315, becomes a plan-time analysis);
//   * the reduction over slices (AddReductionTask, TaskBasedContractor.hpp:258-280) is a
//     double-precision accumulator on the device, summed in slice order (deterministic).
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <set>

#include "chain_plan.h"
#include "common.cuh"

namespace jb {
namespace {

// ---- offset allocator (plan-time simulation of the arena) ---------------------------------------
class OffsetAllocator {
  public:
    size_t Alloc(size_t bytes)
    {
        bytes = Align(bytes);
        auto best = free_.end();
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second >= bytes && (best == free_.end() || it->second < best->second))
                best = it;
        }
        if (best != free_.end()) {
            const size_t off = best->first, sz = best->second;
            free_.erase(best);
            if (sz > bytes)
                free_[off + bytes] = sz - bytes;
            return off;
        }
        // extend the top (merging with a trailing free block)
        if (!free_.empty()) {
            auto last = std::prev(free_.end());
            if (last->first + last->second == top_) {
                const size_t off = last->first;
                free_.erase(last);
                top_ = off + bytes;
                peak_ = std::max(peak_, top_);
                return off;
            }
        }
        const size_t off = top_;
        top_ += bytes;
        peak_ = std::max(peak_, top_);
        return off;
    }
    void Free(size_t off, size_t bytes)
    {
        bytes = Align(bytes);
        auto it = free_.emplace(off, bytes).first;
        auto next = std::next(it);
        if (next != free_.end() && it->first + it->second == next->first) {
            it->second += next->second;
            free_.erase(next);
        }
        if (it != free_.begin()) {
            auto prev = std::prev(it);
            if (prev->first + prev->second == it->first) {
                prev->second += it->second;
                free_.erase(it);
            }
        }
    }
    size_t Peak() const { return peak_; }
    static size_t Align(size_t b) { return (std::max<size_t>(b, 1) + 511) & ~size_t(511); }

  private:
    std::map<size_t, size_t> free_;
    size_t top_ = 0, peak_ = 0;
};

// ---- device-side slice selection -----------------------------------------------------------------
constexpr int kSliceMaxRem = 16;
constexpr int kSliceMaxSl = 8;

struct SliceLeafDesc {
    long long src_off; // element offset of the unsliced leaf in the arena
    long long dst_off; // element offset of the sliced view
    long long out_elems;
    int n_rem, n_sl;
    long long rem_ext[kSliceMaxRem], rem_stride[kSliceMaxRem];
    long long sl_div[kSliceMaxSl], sl_dim[kSliceMaxSl], sl_stride[kSliceMaxSl];
    int pow2;                    // every remaining extent is a power of two: digits by shift and mask
    int rem_shift[kSliceMaxRem]; // bit position of digit j in the view's element index (pow2 only)
};
constexpr int kSliceChunk = 1024; // view elements per CTA of SliceLeavesKernel

struct DeviceState {
    long long next_id;  // slice id of the next graph launch (sequential mode)
    long long ordinal;  // slices accumulated since reset
    long long list_pos; // >= 0: take ids from the list
    unsigned int ctas_done; // AccumulateKernel: CTAs that have finished their share of the current slice
    unsigned int pad;
};

// grid (chunks of kSliceChunk view elements, descriptors): the chunk index is in grid.x (limit 2^31 - 1) so a
// view of any size fits; the descriptor count (sliced leaves + deferred roots) stays far below 65535
template <typename V>
__global__ void __launch_bounds__(128)
    SliceLeavesKernel(V *__restrict__ arena, const SliceLeafDesc *__restrict__ descs,
                      const DeviceState *__restrict__ st, const long long *__restrict__ list,
                      const long long view_base, const long long view_stride)
{
    const SliceLeafDesc &d = descs[blockIdx.y];
    const long long e0 = static_cast<long long>(blockIdx.x) * kSliceChunk;
    if (e0 >= d.out_elems)
        return;
    const long long e1 = min(d.out_elems, e0 + kSliceChunk);
    // slice batching: blockIdx.z = slice within the batch; its views live view_stride elements further on
    const long long sid = st->list_pos >= 0 ? list[st->list_pos + blockIdx.z] : st->next_id + blockIdx.z;
    V *__restrict__ views = arena + view_base + blockIdx.z * view_stride;
    long long base = d.src_off;
    for (int s = 0; s < d.n_sl; s++)
        base += ((sid / d.sl_div[s]) % d.sl_dim[s]) * d.sl_stride[s];
    if (d.pow2) {
        for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            long long off = base;
            for (int j = 0; j < d.n_rem; j++)
                off += ((e >> d.rem_shift[j]) & (d.rem_ext[j] - 1)) * d.rem_stride[j];
            views[d.dst_off + e] = arena[off];
        }
        return;
    }
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        long long rem = e, off = base;
        for (int j = d.n_rem - 1; j >= 0; j--) {
            off += (rem % d.rem_ext[j]) * d.rem_stride[j];
            rem /= d.rem_ext[j];
        }
        views[d.dst_off + e] = arena[off];
    }
}

// Grid-stride over the result elements; the results of the `batch` slices of one launch (result_stride elements
// apart) are added in slice order; the CTA that finishes last advances the slice cursor (every CTA has read
// `ordinal` before it counts itself done, so the update cannot race with a reader).
template <typename C>
__global__ void __launch_bounds__(256)
    AccumulateKernel(const C *__restrict__ result, double2 *__restrict__ acc, C *__restrict__ store,
                     long long elems, long long store_cap, DeviceState *st, int batch, long long result_stride)
{
    const long long ordinal = st->ordinal;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < elems; i += stride) {
        double2 a = acc[i];
        for (int z = 0; z < batch; z++) {
            const C v = result[z * result_stride + i];
            a.x += static_cast<double>(v.x);
            a.y += static_cast<double>(v.y);
            if (store != nullptr && ordinal + z < store_cap)
                store[(ordinal + z) * elems + i] = v;
        }
        acc[i] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool last = true;
        if (gridDim.x > 1) {
            __threadfence();
            last = atomicAdd(&st->ctas_done, 1u) == gridDim.x - 1;
        }
        if (last) {
            st->ctas_done = 0;
            st->ordinal = ordinal + batch;
            if (st->list_pos >= 0)
                st->list_pos = st->list_pos + batch;
            else
                st->next_id = st->next_id + batch;
        }
    }
}

__global__ void SetStateKernel(DeviceState *st, long long next_id, long long list_pos,
                               int reset_ordinal)
{
    st->next_id = next_id;
    st->list_pos = list_pos;
    if (reset_ordinal)
        st->ordinal = 0;
}

struct Node {
    std::vector<int32_t> modes;
    std::vector<int64_t> extent;
    int64_t elems = 1;
    bool slice_dep = false; // depends on the slice id
    bool is_leaf = false;
    size_t offset = 0; // byte offset in the arena of the tensor the steps read
    size_t raw_offset = 0; // leaves / deferred roots: byte offset of the unsliced data
    bool is_view = false;  // deferred root: `offset` is the per-slice view of the shared tensor at raw_offset
    int64_t raw_elems = 0; // deferred root: elements of the unsliced tensor
    int desc = -1;         // index of this node's slice descriptor (sliced leaves and deferred roots)
    int last_use = -1;     // last step (in execution order) reading this node
    bool used_by_slice_step = false;
    bool per_slice = false; // `offset` is relative to the per-slice region (replicated `batch` times)
};

struct Step {
    int a, b, c;
    bool shared;
    ContractPlan cp;
    int op = -1; // per-slice launch unit executing this step
};

// A per-slice launch unit: one path step, or a fused chain of path steps (chain.cu)
struct Op {
    int kernel = 0; // 0 stream, 1 ttgt, 2 fused chain
    std::vector<int> steps;
    int x0 = -1;              // chain: node of the chained tensor before the first step
    std::vector<int> r_nodes; // chain: small operand of every step
    int out = -1;             // node written
    ChainOp chain;
};

} // namespace
} // namespace jb

using namespace jb;

struct jb_plan {
    int dtype = JB_C64;
    int device = 0;
    int flags = 0;
    size_t eb = 8;
    int num_leaves = 0;
    std::vector<Node> nodes;
    std::vector<Step> steps;
    std::vector<int> shared_order, slice_order;
    std::vector<Op> ops;        // per-slice launch units in execution order
    std::vector<Op> shared_ops; // launch units of the slice-independent steps
    std::vector<int32_t> sliced_modes;
    std::vector<int64_t> sliced_dims;
    int64_t num_slices = 1;
    std::vector<SliceLeafDesc> slice_descs;

    unsigned char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t ws_off = 0, ws_bytes = 0;
    size_t acc_off = 0, store_off = 0, state_off = 0, list_off = 0, descs_off = 0;
    int chain_slot = -1; // constant-bank slot of the chains' register stages (-1: none, chains keep their matrices in shared memory)
    // slice batching: the per-slice tensors (views, per-slice intermediates) live in a region of slice_bytes that
    // is replicated `batch` times from slice_base on; one graph replay then contracts `batch` slices
    int batch = 1;
    size_t slice_base = 0, slice_bytes = 0;
    cudaGraph_t graph1 = nullptr; // the single-slice variant (remainders, profiling)
    cudaGraphExec_t graph_exec1 = nullptr;
    int64_t store_cap = 0, list_cap = 0;
    int64_t result_elems = 1;
    int result_node = -1;

    cudaStream_t stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t shared_graph = nullptr; // the slice-independent steps
    cudaGraphExec_t shared_exec = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool shared_done = false;
    bool have_run = false;
    long long *h_list = nullptr;   // two halves of list_cap ids, used alternately
    cudaEvent_t ev_list[2] = {nullptr, nullptr};
    int list_half = 0;
    // upload staging: the raw leaves are contiguous at the start of the arena; small networks are
    // uploaded as ONE copy from a pinned mirror instead of one copy per leaf
    unsigned char *h_stage = nullptr;
    size_t stage_bytes = 0;
    cudaEvent_t ev_stage = nullptr;
    jb_plan_stats_t stats;

    template <typename T> T *At(size_t off) const { return reinterpret_cast<T *>(arena + off); }
};

namespace {

unsigned char *NodePtr(const jb_plan *p, int node)
{
    const Node &n = p->nodes[node];
    return p->arena + (n.per_slice ? p->slice_base : 0) + n.offset;
}
long long NodeStride(const jb_plan *p, int node)
{
    return p->nodes[node].per_slice ? static_cast<long long>(p->slice_bytes) : 0;
}

// `batch` slices per launch: per-slice operands lie slice_bytes apart, shared ones have stride 0
int LaunchOp(jb_plan *p, const Op &op, int batch)
{
    // a deferred root is written unsliced (raw_offset, shared region); everything else where its consumers read it
    const Node &out = p->nodes[op.out];
    unsigned char *dst = out.is_view ? p->arena + out.raw_offset : NodePtr(p, op.out);
    const long long dst_stride = out.is_view ? 0 : NodeStride(p, op.out);
    if (op.kernel == 2) {
        const void *r[kChainMaxSteps];
        ChainBatchArgs ba;
        ba.count = batch;
        ba.stride_x0 = NodeStride(p, op.x0);
        ba.stride_xk = dst_stride;
        for (size_t i = 0; i < op.r_nodes.size(); i++) {
            r[i] = NodePtr(p, op.r_nodes[i]);
            ba.stride_r[i] = NodeStride(p, op.r_nodes[i]);
        }
        // (a plan without a slot has no register stage: the slot argument is then unused)
        return LaunchChain(op.chain, NodePtr(p, op.x0), r, dst, p->chain_slot >= 0 ? p->chain_slot : ChainOperatorSlot(),
                           p->stream, batch > 1 ? &ba : nullptr);
    }
    const Step &st = p->steps[op.steps[0]];
    BatchArgs ba;
    ba.count = batch;
    ba.stride_a = NodeStride(p, st.a);
    ba.stride_b = NodeStride(p, st.b);
    ba.stride_c = dst_stride;
    return LaunchContract(st.cp, NodePtr(p, st.a), NodePtr(p, st.b), dst, p->arena + p->ws_off, p->stream,
                          batch > 1 ? &ba : nullptr);
}

int LaunchSliceViews(jb_plan *p, int batch)
{
    if (p->slice_descs.empty())
        return 0;
    long long max_out = 1;
    for (const SliceLeafDesc &sd : p->slice_descs)
        max_out = std::max(max_out, sd.out_elems);
    JB_REQUIRE(p->slice_descs.size() <= 65535, "plan: more than 65535 sliced leaves");
    const dim3 n(static_cast<unsigned>((max_out + kSliceChunk - 1) / kSliceChunk),
                 static_cast<unsigned>(p->slice_descs.size()), static_cast<unsigned>(batch));
    const long long base = static_cast<long long>(p->slice_base / p->eb), stride = static_cast<long long>(p->slice_bytes / p->eb);
    if (p->dtype == JB_C64)
        SliceLeavesKernel<uint2><<<n, 128, 0, p->stream>>>(p->At<uint2>(0), p->At<SliceLeafDesc>(p->descs_off),
                                                          p->At<DeviceState>(p->state_off),
                                                          p->At<long long>(p->list_off), base, stride);
    else
        SliceLeavesKernel<uint4><<<n, 128, 0, p->stream>>>(p->At<uint4>(0), p->At<SliceLeafDesc>(p->descs_off),
                                                          p->At<DeviceState>(p->state_off),
                                                          p->At<long long>(p->list_off), base, stride);
    JB_CUDA(cudaGetLastError());
    return 0;
}

int EnqueueSliceBody(jb_plan *p, int batch)
{
    JB_TRY(LaunchSliceViews(p, batch));
    for (const Op &op : p->ops)
        JB_TRY(LaunchOp(p, op, batch));
    const bool store = (p->flags & JB_PLAN_STORE_RESULTS) != 0;
    const void *res = NodePtr(p, p->result_node);
    const long long res_stride = NodeStride(p, p->result_node) / static_cast<long long>(p->eb);
    const unsigned acc_grid = static_cast<unsigned>(
        std::max<long long>(1, std::min<long long>((p->result_elems + 1023) / 1024, 4ll * NumSMs())));
    if (p->dtype == JB_C64)
        AccumulateKernel<float2><<<acc_grid, 256, 0, p->stream>>>(
            static_cast<const float2 *>(res), p->At<double2>(p->acc_off),
            store ? p->At<float2>(p->store_off) : nullptr, p->result_elems, p->store_cap,
            p->At<DeviceState>(p->state_off), batch, res_stride);
    else
        AccumulateKernel<double2><<<acc_grid, 256, 0, p->stream>>>(
            static_cast<const double2 *>(res), p->At<double2>(p->acc_off),
            store ? p->At<double2>(p->store_off) : nullptr, p->result_elems, p->store_cap,
            p->At<DeviceState>(p->state_off), batch, res_stride);
    JB_CUDA(cudaGetLastError());
    return 0;
}

int EnqueueShared(jb_plan *p)
{
    for (const Op &op : p->shared_ops)
        JB_TRY(LaunchOp(p, op, 1));
    return 0;
}

// The slice-independent steps (and the deferred subtrees) run once per reset / upload; they are mostly
// tiny, so they too are replayed from a CUDA graph (hundreds of launches otherwise).
int RunShared(jb_plan *p)
{
    if (p->shared_done)
        return 0;
    if (p->shared_ops.empty()) {
        p->shared_done = true;
        return 0;
    }
    if (p->flags & JB_PLAN_NO_GRAPH) {
        JB_TRY(EnqueueShared(p));
        p->shared_done = true;
        return 0;
    }
    if (p->shared_exec == nullptr) {
        JB_CUDA(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = EnqueueShared(p);
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(p->stream, &g);
        if (rc != 0) {
            if (g)
                cudaGraphDestroy(g);
            return rc;
        }
        JB_CUDA(ce);
        p->shared_graph = g;
        JB_CUDA(cudaGraphInstantiate(&p->shared_exec, p->shared_graph, 0));
    }
    JB_CUDA(cudaGraphLaunch(p->shared_exec, p->stream));
    p->shared_done = true;
    return 0;
}

int CaptureBody(jb_plan *p, int batch, cudaGraph_t *graph, cudaGraphExec_t *exec)
{
    if (*exec != nullptr)
        return 0;
    JB_CUDA(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = EnqueueSliceBody(p, batch);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(p->stream, &g);
    if (rc != 0) {
        if (g)
            cudaGraphDestroy(g);
        return rc;
    }
    JB_CUDA(ce);
    *graph = g;
    JB_CUDA(cudaGraphInstantiate(exec, g, 0));
    return 0;
}

int RunSlices(jb_plan *p, long long first, long long list_pos, long long count)
{
    JB_REQUIRE(p->arena != nullptr, "plan: created with JB_PLAN_DRY_RUN (no device resources)");
    JB_CUDA(cudaSetDevice(p->device));
    JB_TRY(RunShared(p));
    const bool use_graph = !(p->flags & JB_PLAN_NO_GRAPH);
    const long long full = p->batch > 1 ? count / p->batch : 0, rest = count - full * p->batch;
    if (use_graph && full > 0)
        JB_TRY(CaptureBody(p, p->batch, &p->graph, &p->graph_exec));
    if (use_graph && rest > 0)
        JB_TRY(CaptureBody(p, 1, &p->graph1, &p->graph_exec1));
    SetStateKernel<<<1, 1, 0, p->stream>>>(p->At<DeviceState>(p->state_off), first, list_pos, 0);
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaEventRecord(p->ev0, p->stream));
    // whole batches first (one replay contracts `batch` slices), then the remainder slice by slice; the device-side
    // cursor advances by what each replay consumed
    for (long long i = 0; i < full; i++) {
        if (use_graph)
            JB_CUDA(cudaGraphLaunch(p->graph_exec, p->stream));
        else
            JB_TRY(EnqueueSliceBody(p, p->batch));
    }
    for (long long i = 0; i < rest; i++) {
        if (use_graph)
            JB_CUDA(cudaGraphLaunch(p->graph_exec1, p->stream));
        else
            JB_TRY(EnqueueSliceBody(p, 1));
    }
    JB_CUDA(cudaEventRecord(p->ev1, p->stream));
    p->have_run = true;
    return 0;
}

// Everything a plan owns on its device: arena, stream, events, pinned staging.  Shared by jb_plan_create and
// jb_plan_clone (the host-side plan is computed once and copied).
int AllocDeviceResources(jb_plan *p)
{
    size_t free_b = 0, total_b = 0;
    JB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (p->arena_bytes > free_b) {
        return Fail("plan: the contraction needs " + std::to_string(p->arena_bytes >> 20) +
                    " MiB of device memory but only " + std::to_string(free_b >> 20) +
                    " MiB are free; slice more indices");
    }
    JB_CUDA(cudaMalloc(reinterpret_cast<void **>(&p->arena), p->arena_bytes));
    JB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    JB_CUDA(cudaEventCreate(&p->ev0));
    JB_CUDA(cudaEventCreate(&p->ev1));
    JB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&p->h_list), sizeof(long long) * p->list_cap * 2));
    for (auto &e : p->ev_list)
        JB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        size_t end = 0;
        for (int i = 0; i < p->num_leaves; i++) {
            const Node &n = p->nodes[i];
            end = std::max(end, n.raw_offset + static_cast<size_t>(n.desc >= 0 ? n.raw_elems : n.elems) * p->eb);
        }
        constexpr size_t kStageMax = size_t(32) << 20;
        if (end > 0 && end <= kStageMax && p->num_leaves > 1) {
            p->stage_bytes = end;
            JB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&p->h_stage), end));
            std::memset(p->h_stage, 0, end);
            JB_CUDA(cudaEventCreateWithFlags(&p->ev_stage, cudaEventDisableTiming));
        }
    }
    if (!p->slice_descs.empty())
        JB_CUDA(cudaMemcpyAsync(p->arena + p->descs_off, p->slice_descs.data(),
                                sizeof(SliceLeafDesc) * p->slice_descs.size(), cudaMemcpyHostToDevice,
                                p->stream));
    JB_CUDA(cudaMemsetAsync(p->arena + p->acc_off, 0, sizeof(double2) * p->result_elems, p->stream));
    JB_CUDA(cudaMemsetAsync(p->arena + p->state_off, 0, sizeof(DeviceState), p->stream));

    return 0;
}

} // namespace

extern "C" {

int jb_plan_create(const jb_network_desc_t *d, jb_plan **out)
{
    JB_REQUIRE(d != nullptr && out != nullptr, "plan: null argument");
    JB_REQUIRE(d->dtype == JB_C64 || d->dtype == JB_C128, "plan: unknown dtype");
    JB_REQUIRE(d->num_leaves >= 1, "An empty tensor network cannot be contracted.");
    const bool dry = (d->flags & JB_PLAN_DRY_RUN) != 0; // host-side planning only: no device is touched
    if (!dry)
        JB_CUDA(cudaSetDevice(d->device));
    // every early return below releases what the plan holds by then (constant-bank slot, stream, events,
    // pinned buffers, arena): jb_plan_destroy tolerates partially initialised plans
    struct PlanDeleter {
        void operator()(jb_plan *q) const { jb_plan_destroy(q); }
    };
    std::unique_ptr<jb_plan, PlanDeleter> up(new jb_plan());
    jb_plan *p = up.get();
    p->dtype = d->dtype;
    p->device = d->device;
    p->flags = d->flags;
    p->eb = ElemBytes(d->dtype);
    p->num_leaves = d->num_leaves;
    p->sliced_modes.assign(d->sliced_modes, d->sliced_modes + d->num_sliced);
    const bool keep = (d->flags & JB_PLAN_KEEP_INTERMEDIATES) != 0;

    // ---- nodes ----------------------------------------------------------------------------------
    std::map<int32_t, int64_t> mode_dim;
    size_t off = 0;
    for (int i = 0; i < d->num_leaves; i++) {
        Node n;
        n.is_leaf = true;
        const int r = d->rank[i];
        JB_REQUIRE(r >= 0 && r <= JB_MAX_RANK, "plan: leaf rank out of range");
        n.modes.assign(d->mode + off, d->mode + off + r);
        n.extent.assign(d->extent + off, d->extent + off + r);
        for (int j = 0; j < r; j++) {
            JB_REQUIRE(n.extent[j] >= 1, "plan: extents must be positive");
            n.elems *= n.extent[j];
            mode_dim[n.modes[j]] = n.extent[j];
        }
        off += r;
        p->nodes.push_back(n);
    }
    p->sliced_dims.clear();
    p->num_slices = 1;
    for (int32_t m : p->sliced_modes) {
        auto it = mode_dim.find(m);
        JB_REQUIRE(it != mode_dim.end(), "Sliced index does not exist.");
        p->sliced_dims.push_back(it->second);
        JB_REQUIRE(p->num_slices <= (int64_t(1) << 62) / it->second,
                   "plan: more than 2^62 slices (slice ids are 64-bit)");
        p->num_slices *= it->second;
    }
    auto is_sliced = [&](int32_t m) {
        return std::find(p->sliced_modes.begin(), p->sliced_modes.end(), m) != p->sliced_modes.end();
    };

    // ---- deferred slicing ------------------------------------------------------------------------
    // Small subtrees that depend on the slice only through OPEN sliced indices are contracted once,
    // unsliced (the sliced indices stay ordinary indices inside), in the shared phase; the slice is
    // selected on the subtree's root, exactly as it is on a leaf.  The reference de-duplicates tasks
    // that do not depend on the slice at all (TaskBasedContractor.hpp:216-222); this extends it to
    // tasks whose dependence can be postponed — per slice, only the steps that touch large tensors
    // remain.  Conditions: no sliced index is contracted inside the subtree (it would become a sum),
    // every tensor of the unsliced subtree stays small, and the root fits the slice descriptor.
    const int n_total = d->num_leaves + std::max(d->num_steps, 0);
    std::vector<char> defer_inside(n_total, 0), defer_root(n_total, 0);
    const bool defer_enabled = !keep && !(d->flags & JB_PLAN_NO_FUSE) && d->num_sliced > 0 && d->num_steps > 1 && [] {
        const char *e = getenv("JB_DISABLE_DEFER");
        return !(e && e[0] == '1');
    }();
    if (defer_enabled) {
        // every tensor of the unsliced subtree stays below this (JB_DEFER_MAX_LOG2 overrides, for experiments)
        const int64_t kDeferMaxElems = [] {
            const char *e = getenv("JB_DEFER_MAX_LOG2");
            const int l = e ? atoi(e) : 14;
            return int64_t(1) << std::max(0, std::min(l, 30));
        }();
        struct Sym {
            std::vector<int32_t> modes;
            int64_t elems = 1, max_elems = 1;
            bool dep = false, closed = false, ok = false;
            int a = -1, b = -1;
        };
        std::vector<Sym> sym(n_total);
        bool valid = true;
        for (int i = 0; i < d->num_leaves; i++) {
            sym[i].modes = p->nodes[i].modes;
            sym[i].elems = sym[i].max_elems = p->nodes[i].elems;
            for (int32_t m : sym[i].modes)
                sym[i].dep = sym[i].dep || is_sliced(m);
        }
        for (int s = 0; s < d->num_steps && valid; s++) {
            const int a = d->path[2 * s], b = d->path[2 * s + 1], c = d->num_leaves + s;
            if (a < 0 || b < 0 || a >= c || b >= c || a == b) {
                valid = false; // reported by the replay below
                break;
            }
            Sym &C = sym[c];
            C.a = a;
            C.b = b;
            C.dep = sym[a].dep || sym[b].dep;
            C.closed = sym[a].closed || sym[b].closed;
            for (int32_t m : sym[a].modes) {
                if (std::find(sym[b].modes.begin(), sym[b].modes.end(), m) == sym[b].modes.end())
                    C.modes.push_back(m);
                else if (is_sliced(m))
                    C.closed = true;
            }
            for (int32_t m : sym[b].modes)
                if (std::find(sym[a].modes.begin(), sym[a].modes.end(), m) == sym[a].modes.end())
                    C.modes.push_back(m);
            for (int32_t m : C.modes)
                C.elems *= mode_dim[m];
            C.max_elems = std::max({C.elems, sym[a].max_elems, sym[b].max_elems});
            int n_sl = 0, n_rem = 0;
            for (int32_t m : C.modes)
                (is_sliced(m) ? n_sl : n_rem)++;
            C.ok = C.dep && !C.closed && C.max_elems <= kDeferMaxElems && n_sl >= 1 && n_sl <= kSliceMaxSl &&
                   n_rem <= kSliceMaxRem && static_cast<int>(C.modes.size()) <= JB_MAX_RANK;
        }
        if (valid) {
            std::vector<int> consumer(n_total, -1);
            for (int s = 0; s < d->num_steps; s++) {
                consumer[d->path[2 * s]] = d->num_leaves + s;
                consumer[d->path[2 * s + 1]] = d->num_leaves + s;
            }
            const int last = n_total - 1;
            for (int c = d->num_leaves; c < n_total; c++) {
                if (!sym[c].ok || c == last)
                    continue;
                const int up_node = consumer[c];
                if (up_node >= 0 && up_node != last && sym[up_node].ok)
                    continue; // not maximal
                defer_root[c] = 1;
                std::vector<int> stack = {sym[c].a, sym[c].b};
                while (!stack.empty()) {
                    const int v = stack.back();
                    stack.pop_back();
                    defer_inside[v] = 1;
                    if (v >= d->num_leaves) {
                        stack.push_back(sym[v].a);
                        stack.push_back(sym[v].b);
                    }
                }
            }
        }
    }

    // ---- arena layout: raw leaves, then sliced views ----------------------------------------------
    OffsetAllocator alloc;  // shared region: raw leaves, slice-independent tensors, fixed regions
    OffsetAllocator salloc; // per-slice region (replicated `batch` times): views and per-slice intermediates
    for (int i = 0; i < d->num_leaves; i++) {
        Node &n = p->nodes[i];
        n.raw_offset = alloc.Alloc(n.elems * p->eb);
        n.offset = n.raw_offset;
    }
    // turns node n (unsliced layout) into its per-slice view: descriptor + view buffer; the node's
    // modes / extents / elems become those of the view.  src_off is filled in by the caller.
    auto make_view = [&](Node &n) -> int {
        std::vector<int> sl_axes;
        for (size_t j = 0; j < n.modes.size(); j++)
            if (is_sliced(n.modes[j]))
                sl_axes.push_back(static_cast<int>(j));
        if (sl_axes.empty())
            return 0;
        n.slice_dep = true;
        SliceLeafDesc sd;
        std::memset(&sd, 0, sizeof(sd));
        JB_REQUIRE(sl_axes.size() <= kSliceMaxSl, "plan: too many sliced indices on one leaf");
        std::vector<int64_t> stride(n.modes.size());
        int64_t s = 1;
        for (int j = static_cast<int>(n.modes.size()) - 1; j >= 0; j--) {
            stride[j] = s;
            s *= n.extent[j];
        }
        std::vector<int32_t> new_modes;
        std::vector<int64_t> new_ext;
        int64_t out_elems = 1;
        for (size_t j = 0; j < n.modes.size(); j++) {
            if (is_sliced(n.modes[j])) {
                // digit of this mode inside the slice id: first listed mode is slowest
                int64_t div = 1;
                const size_t pos = std::find(p->sliced_modes.begin(), p->sliced_modes.end(), n.modes[j]) -
                                   p->sliced_modes.begin();
                for (size_t q = pos + 1; q < p->sliced_modes.size(); q++)
                    div *= p->sliced_dims[q];
                sd.sl_div[sd.n_sl] = div;
                sd.sl_dim[sd.n_sl] = n.extent[j];
                sd.sl_stride[sd.n_sl] = stride[j];
                sd.n_sl++;
            }
            else {
                JB_REQUIRE(sd.n_rem < kSliceMaxRem, "plan: sliced leaf rank too large");
                sd.rem_ext[sd.n_rem] = n.extent[j];
                sd.rem_stride[sd.n_rem] = stride[j];
                sd.n_rem++;
                new_modes.push_back(n.modes[j]);
                new_ext.push_back(n.extent[j]);
                out_elems *= n.extent[j];
            }
        }
        sd.out_elems = out_elems;
        sd.pow2 = 1;
        {
            int shift = 0;
            for (int j = sd.n_rem - 1; j >= 0; j--) {
                sd.rem_shift[j] = shift;
                if (!IsPow2(sd.rem_ext[j]))
                    sd.pow2 = 0;
                else
                    shift += Log2(sd.rem_ext[j]);
            }
        }
        sd.src_off = static_cast<long long>(n.raw_offset / p->eb);
        n.raw_elems = n.elems;
        n.offset = salloc.Alloc(out_elems * p->eb);
        n.per_slice = true;
        sd.dst_off = static_cast<long long>(n.offset / p->eb); // relative to the slice's region
        n.modes = new_modes;
        n.extent = new_ext;
        n.elems = out_elems;
        n.desc = static_cast<int>(p->slice_descs.size());
        p->slice_descs.push_back(sd);
        return 0;
    };
    for (int i = 0; i < d->num_leaves; i++)
        if (!defer_inside[i])
            JB_TRY(make_view(p->nodes[i]));

    // ---- steps: symbolic replay of the path (include/jet/PathInfo.hpp:262-297) --------------------
    JB_REQUIRE(d->num_steps >= 0, "plan: negative step count");
    for (int s = 0; s < d->num_steps; s++) {
        const int a = d->path[2 * s], b = d->path[2 * s + 1];
        const int cur = static_cast<int>(p->nodes.size());
        JB_REQUIRE(a >= 0 && a < cur, "Node ID 1 in contraction pair is invalid.");
        JB_REQUIRE(b >= 0 && b < cur, "Node ID 2 in contraction pair is invalid.");
        JB_REQUIRE(a != b, "plan: a node cannot be contracted with itself");
        Step st;
        st.a = a;
        st.b = b;
        st.c = cur;
        const Node &A = p->nodes[a];
        const Node &B = p->nodes[b];
        JB_TRY(MakeContractPlan(p->dtype, static_cast<int>(A.modes.size()), A.extent.data(),
                                A.modes.data(), static_cast<int>(B.modes.size()), B.extent.data(),
                                B.modes.data(), &st.cp));
        st.shared = !(A.slice_dep || B.slice_dep);
        Node C;
        C.modes = st.cp.modes_c;
        C.extent = st.cp.extent_c;
        C.elems = st.cp.m * st.cp.n;
        C.slice_dep = !st.shared;
        if (defer_root[cur] && st.shared) {
            // the step writes the unsliced tensor (shared phase); consumers read its per-slice view
            C.is_view = true;
            JB_TRY(make_view(C)); // src_off is patched once the unsliced tensor has an arena offset
            JB_REQUIRE(C.desc >= 0, "plan: internal: deferred root without sliced index");
        }
        p->nodes.push_back(C);
        p->steps.push_back(st);
    }
    p->result_node = static_cast<int>(p->nodes.size()) - 1;
    p->result_elems = p->nodes[p->result_node].elems;
    if (d->num_steps > 0)
        p->nodes[p->result_node].slice_dep = true; // the accumulate step runs per slice
    // the final step is always executed per slice so the accumulator sees one result per slice
    if (!p->steps.empty())
        p->steps.back().shared = false;

    for (size_t s = 0; s < p->steps.size(); s++)
        (p->steps[s].shared ? p->shared_order : p->slice_order).push_back(static_cast<int>(s));

    // ---- slice batching (decided in two steps: candidate here, batch size once the region sizes are known) ----
    constexpr int64_t kBatchMaxElems = 1 << 20; // per-slice tensors above this fill the GPU on their own
    int want_batch = d->batch;
    if (const char *e = getenv("JB_PLAN_BATCH"))
        want_batch = atoi(e);
    bool batch_candidate = false;

    // ---- per-slice launch units: fuse runs of "large tensor absorbs a small tensor" steps ----------
    {
        bool fuse = !keep && !(d->flags & JB_PLAN_NO_FUSE) && ChainFusionEnabled();
        // the matrices of the chains' register stages go through a constant-bank slot owned by the plan; when every
        // slot of the device is taken (more than five live plans) the chains are still fused, with all their matrices in
        // shared memory
        if (fuse)
            p->chain_slot = ChainAcquireSlot(p->device);
        const bool have_slot = p->chain_slot >= 0;
        std::vector<int> consumer(p->nodes.size(), -1);
        for (size_t s = 0; s < p->steps.size(); s++) {
            consumer[p->steps[s].a] = static_cast<int>(s);
            consumer[p->steps[s].b] = static_cast<int>(s);
        }
        // a step can join a chain whose running tensor is node x when its other operand is small; chains
        // never mix slice-independent and per-slice steps
        bool want_shared = false;
        auto as_operand = [&](int s, int x, ChainOperand *o, int *r_node) {
            const Step &st = p->steps[s];
            if (st.shared != want_shared || st.cp.kernel != 0)
                return false;
            const bool x_left = st.a == x;
            const int r = x_left ? st.b : st.a;
            const int64_t free_r = x_left ? st.cp.n : st.cp.m;
            if (st.cp.k > 16 || free_r > 16)
                return false;
            o->modes = p->nodes[r].modes;
            o->extent = p->nodes[r].extent;
            o->x_is_left = x_left;
            *r_node = r;
            return true;
        };
        std::vector<char> taken(p->steps.size(), 0);
        const int max_tile = ChainMaxTileBits(p->dtype);
        // slice batching needs per-slice step matrices: chains whose small operands depend on the slice then
        // cannot use the (one set per launch) constant-bank register stages
        int64_t max_slice_elems = 1;
        for (const Node &n : p->nodes)
            if (n.slice_dep)
                max_slice_elems = std::max(max_slice_elems, n.elems);
        batch_candidate = want_batch != 1 && !keep && p->num_slices > 1 && max_slice_elems <= kBatchMaxElems;
        auto build = [&](const std::vector<int> &order, bool shared_steps) {
        want_shared = shared_steps;
        std::vector<Op> ops;
        for (int s : order) {
            if (taken[s])
                continue;
            const Step &st = p->steps[s];
            Op single;
            single.kernel = st.cp.kernel;
            single.steps = {s};
            single.out = st.c;
            if (!fuse) {
                ops.push_back(single);
                continue;
            }
            const int x0 = (p->nodes[st.a].elems >= p->nodes[st.b].elems) ? st.a : st.b;
            Op chain;
            chain.kernel = 2;
            chain.x0 = x0;
            std::vector<ChainOperand> operands;
            bool r_dep = false; // a small operand of the chain depends on the slice
            int x = x0, cur = s;
            while (cur >= 0 && !taken[cur] && static_cast<int>(operands.size()) < kChainMaxSteps) {
                ChainOperand o;
                int r_node = -1;
                if (!as_operand(cur, x, &o, &r_node))
                    break;
                operands.push_back(o);
                const bool dep_now = r_dep || p->nodes[r_node].slice_dep;
                ChainOp trial;
                if (MakeChainOp(p->dtype, p->nodes[x0].modes, p->nodes[x0].extent, operands, max_tile,
                                &trial, nullptr, have_slot && !(batch_candidate && dep_now)) != 0) {
                    operands.pop_back();
                    break;
                }
                r_dep = dep_now;
                chain.chain = trial;
                chain.steps.push_back(cur);
                chain.r_nodes.push_back(r_node);
                chain.out = p->steps[cur].c;
                x = p->steps[cur].c;
                cur = consumer[x];
            }
            if (chain.steps.size() >= 2) {
                for (int cs : chain.steps)
                    taken[cs] = 1;
                ops.push_back(chain);
            }
            else {
                taken[s] = 1;
                ops.push_back(single);
            }
        }
        // a unit runs where its LAST step stood in the path: everything it reads exists by then
        std::stable_sort(ops.begin(), ops.end(),
                         [](const Op &x, const Op &y) { return x.steps.back() < y.steps.back(); });
        return ops;
        };
        p->ops = build(p->slice_order, false);
        p->shared_ops = build(p->shared_order, true); // the slice-independent steps fuse the same way
        for (size_t o = 0; o < p->ops.size(); o++)
            for (int cs : p->ops[o].steps)
                p->steps[cs].op = static_cast<int>(o);
    }

    // ---- lifetimes in execution order (shared steps first, then per-slice launch units) -----------
    struct Exec {
        std::vector<int> reads;
        int out;
        bool shared;
    };
    std::vector<Exec> exec;
    auto add_exec = [&](const Op &op, bool shared) {
        Exec e;
        e.shared = shared;
        e.out = op.out;
        if (op.kernel == 2) {
            e.reads = op.r_nodes;
            e.reads.push_back(op.x0);
        }
        else {
            e.reads = {p->steps[op.steps[0]].a, p->steps[op.steps[0]].b};
        }
        exec.push_back(e);
    };
    for (const Op &op : p->shared_ops)
        add_exec(op, true);
    for (const Op &op : p->ops)
        add_exec(op, false);
    for (size_t e = 0; e < exec.size(); e++) {
        for (int in : exec[e].reads) {
            p->nodes[in].last_use = static_cast<int>(e);
            if (!exec[e].shared)
                p->nodes[in].used_by_slice_step = true;
        }
    }
    // fixed regions first (never recycled): workspace, accumulator, per-slice result store, state,
    // slice-id list, slice descriptors
    size_t ws_max = 0;
    for (const Step &st : p->steps)
        ws_max = std::max(ws_max, st.cp.ws_bytes);
    p->ws_bytes = ws_max;
    p->ws_off = alloc.Alloc(std::max<size_t>(ws_max, 512));
    p->acc_off = alloc.Alloc(sizeof(double2) * p->result_elems);
    // per-slice result store / explicit slice-id lists hold at most 2^16 entries per run
    constexpr int64_t kMaxListed = 1 << 16;
    p->store_cap = (d->flags & JB_PLAN_STORE_RESULTS) ? std::min<int64_t>(p->num_slices, kMaxListed) : 0;
    if (p->store_cap > 0)
        p->store_off = alloc.Alloc(p->eb * p->result_elems * p->store_cap);
    p->state_off = alloc.Alloc(sizeof(DeviceState));
    p->list_cap = kMaxListed; // ids may repeat: the list length does not depend on the number of slices
    p->list_off = alloc.Alloc(sizeof(long long) * p->list_cap);
    p->descs_off = alloc.Alloc(sizeof(SliceLeafDesc) * std::max<size_t>(p->slice_descs.size(), 1));
    for (size_t e = 0; e < exec.size(); e++) {
        Node &C = p->nodes[exec[e].out];
        if (C.is_view) { // the view buffer exists already; the unsliced tensor lives for the whole run
            C.raw_offset = alloc.Alloc(C.raw_elems * p->eb);
            p->slice_descs[C.desc].src_off = static_cast<long long>(C.raw_offset / p->eb);
        }
        else {
            C.per_slice = C.slice_dep;
            C.offset = (C.per_slice ? salloc : alloc).Alloc(C.elems * p->eb);
        }
        if (keep)
            continue;
        std::set<int> seen;
        for (int in : exec[e].reads) {
            Node &I = p->nodes[in];
            if (I.is_leaf || I.is_view || I.last_use != static_cast<int>(e) || !seen.insert(in).second)
                continue;
            // a shared tensor read by per-slice steps must survive every slice
            const bool producer_shared = !I.slice_dep;
            if (producer_shared && I.used_by_slice_step)
                continue;
            (I.per_slice ? salloc : alloc).Free(I.offset, I.elems * p->eb);
        }
    }
    p->slice_base = OffsetAllocator::Align(alloc.Peak());
    p->slice_bytes = OffsetAllocator::Align(salloc.Peak());
    p->batch = 1;
    if (batch_candidate) {
        // as many slices per launch as keep the replicated region small (L2-sized working sets are the point)
        int64_t b = want_batch > 1 ? want_batch : 64;
        b = std::min<int64_t>(b, p->num_slices);
        b = std::min<int64_t>(b, std::max<int64_t>(1, static_cast<int64_t>((size_t(1) << 30) / p->slice_bytes)));
        b = std::min<int64_t>(b, 1024);
        int pow2 = 1;
        while (2 * pow2 <= b)
            pow2 *= 2;
        p->batch = pow2;
    }
    p->arena_bytes = p->slice_base + static_cast<size_t>(p->batch) * p->slice_bytes;

    // ---- statistics -----------------------------------------------------------------------------------
    jb_plan_stats_t &S = p->stats;
    std::memset(&S, 0, sizeof(S));
    S.num_slices = p->num_slices;
    S.result_elems = p->result_elems;
    const Node &RN = p->nodes[p->result_node];
    S.result_rank = static_cast<int32_t>(RN.modes.size());
    for (size_t j = 0; j < RN.modes.size(); j++) {
        S.result_modes[j] = RN.modes[j];
        S.result_extent[j] = RN.extent[j];
    }
    S.steps_total = static_cast<int32_t>(p->steps.size());
    S.steps_shared = static_cast<int32_t>(p->shared_order.size());
    S.launches_per_slice = (p->slice_descs.empty() ? 0 : 1) + 1;
    for (const Step &st : p->steps) {
        // Jet-convention flops / memory are those of the SLICED network (PathInfo.hpp:168-182 after
        // SliceIndices): a step of a deferred subtree carries its open sliced indices in M or N
        const Node &C = p->nodes[st.c];
        double unsliced = 1.0;
        if (C.is_view)
            unsliced = double(C.raw_elems) / double(C.elems);
        else
            for (size_t j = 0; j < C.modes.size(); j++)
                if (is_sliced(C.modes[j]))
                    unsliced *= double(C.extent[j]);
        S.jet_flops_per_slice += 2.0 * double(st.cp.m) * double(st.cp.n) * double(st.cp.k) / unsliced;
        S.max_step_elems = std::max<int64_t>(S.max_step_elems, static_cast<int64_t>(double(st.cp.m * st.cp.n) / unsliced));
        if (st.shared) {
            S.flops_shared += st.cp.flops();
            S.bytes_shared += st.cp.bytes();
        }
        else {
            S.flops_per_slice += st.cp.flops();
            S.bytes_per_slice += st.cp.bytes();
        }
    }
    for (const Op &op : p->ops) {
        if (op.kernel == 2) {
            S.chains++;
            S.steps_chained += static_cast<int32_t>(op.steps.size());
            S.launches_per_slice += op.chain.launches; // chain kernel (+ matrix gather)
            S.fused_bytes_per_slice += op.chain.bytes;
        }
        else {
            const Step &st = p->steps[op.steps[0]];
            S.launches_per_slice += st.cp.launches;
            S.fused_bytes_per_slice += st.cp.bytes();
            if (st.cp.kernel == 0)
                S.steps_stream++;
            else
                S.steps_ttgt++;
        }
    }
    S.arena_bytes = p->arena_bytes;
    S.batch = p->batch;

    if (!dry) {
        JB_TRY(AllocDeviceResources(p));
        if (d->h_data != nullptr)
            JB_TRY(jb_plan_upload(p, d->h_data));
    }
    *out = up.release();
    return 0;
}


int jb_plan_clone(const jb_plan *src, int device, jb_plan **out)
{
    JB_REQUIRE(src != nullptr && out != nullptr, "plan: null argument");
    JB_CUDA(cudaSetDevice(device));
    struct PlanDeleter {
        void operator()(jb_plan *q) const { jb_plan_destroy(q); }
    };
    std::unique_ptr<jb_plan, PlanDeleter> up(new jb_plan());
    jb_plan *p = up.get();
    // host-side plan: copied; device-side resources: fresh
    p->dtype = src->dtype;
    p->device = device;
    p->flags = src->flags;
    p->eb = src->eb;
    p->num_leaves = src->num_leaves;
    p->nodes = src->nodes;
    p->steps = src->steps;
    p->shared_order = src->shared_order;
    p->slice_order = src->slice_order;
    p->ops = src->ops;
    p->shared_ops = src->shared_ops;
    p->sliced_modes = src->sliced_modes;
    p->sliced_dims = src->sliced_dims;
    p->num_slices = src->num_slices;
    p->slice_descs = src->slice_descs;
    p->arena_bytes = src->arena_bytes;
    p->batch = src->batch;
    p->slice_base = src->slice_base;
    p->slice_bytes = src->slice_bytes;
    p->ws_off = src->ws_off;
    p->ws_bytes = src->ws_bytes;
    p->acc_off = src->acc_off;
    p->store_off = src->store_off;
    p->state_off = src->state_off;
    p->list_off = src->list_off;
    p->descs_off = src->descs_off;
    p->store_cap = src->store_cap;
    p->list_cap = src->list_cap;
    p->result_elems = src->result_elems;
    p->result_node = src->result_node;
    p->stats = src->stats;
    if (src->chain_slot >= 0) {
        p->chain_slot = ChainAcquireSlot(device);
        JB_REQUIRE(p->chain_slot >= 0, "plan: no free constant-bank slot for another plan on this device");
    }
    JB_TRY(AllocDeviceResources(p));
    *out = up.release();
    return 0;
}

int jb_plan_destroy(jb_plan *p)
{
    if (p == nullptr)
        return 0;
    if (p->arena == nullptr && p->stream == nullptr) { // dry run / failed before any device resource
        if (p->chain_slot >= 0)
            ChainReleaseSlot(p->device, p->chain_slot);
        delete p;
        return 0;
    }
    cudaSetDevice(p->device);
    if (p->stream)
        cudaStreamSynchronize(p->stream);
    if (p->graph_exec)
        cudaGraphExecDestroy(p->graph_exec);
    if (p->graph)
        cudaGraphDestroy(p->graph);
    if (p->graph_exec1)
        cudaGraphExecDestroy(p->graph_exec1);
    if (p->graph1)
        cudaGraphDestroy(p->graph1);
    if (p->shared_exec)
        cudaGraphExecDestroy(p->shared_exec);
    if (p->shared_graph)
        cudaGraphDestroy(p->shared_graph);
    if (p->ev0)
        cudaEventDestroy(p->ev0);
    if (p->ev1)
        cudaEventDestroy(p->ev1);
    if (p->stream)
        cudaStreamDestroy(p->stream);
    if (p->h_list)
        cudaFreeHost(p->h_list);
    for (auto &e : p->ev_list)
        if (e)
            cudaEventDestroy(e);
    if (p->h_stage)
        cudaFreeHost(p->h_stage);
    if (p->ev_stage)
        cudaEventDestroy(p->ev_stage);
    if (p->arena)
        cudaFree(p->arena);
    if (p->chain_slot >= 0)
        ChainReleaseSlot(p->device, p->chain_slot);
    delete p;
    return 0;
}

int jb_plan_stats(const jb_plan *p, jb_plan_stats_t *stats)
{
    JB_REQUIRE(p && stats, "plan: null argument");
    *stats = p->stats;
    return 0;
}

int jb_plan_upload(jb_plan *p, const void *const *h_data)
{
    JB_REQUIRE(p && h_data, "plan: null argument");
    JB_REQUIRE(p->arena != nullptr, "plan: created with JB_PLAN_DRY_RUN (no device resources)");
    JB_CUDA(cudaSetDevice(p->device));
    if (p->h_stage != nullptr)
        JB_CUDA(cudaEventSynchronize(p->ev_stage)); // the previous upload has left the pinned mirror
    for (int i = 0; i < p->num_leaves; i++) {
        const Node &n = p->nodes[i];
        const size_t raw_elems = static_cast<size_t>(n.desc >= 0 ? n.raw_elems : n.elems); // unsliced
        JB_REQUIRE(h_data[i] != nullptr, "plan: null leaf data");
        if (p->h_stage != nullptr)
            std::memcpy(p->h_stage + n.raw_offset, h_data[i], raw_elems * p->eb);
        else
            JB_CUDA(cudaMemcpyAsync(p->arena + n.raw_offset, h_data[i], raw_elems * p->eb,
                                    cudaMemcpyHostToDevice, p->stream));
    }
    if (p->h_stage != nullptr) {
        JB_CUDA(cudaMemcpyAsync(p->arena, p->h_stage, p->stage_bytes, cudaMemcpyHostToDevice, p->stream));
        JB_CUDA(cudaEventRecord(p->ev_stage, p->stream));
    }
    p->shared_done = false;
    return 0;
}

int jb_plan_reset(jb_plan *p)
{
    JB_REQUIRE(p, "plan: null argument");
    JB_REQUIRE(p->arena != nullptr, "plan: created with JB_PLAN_DRY_RUN (no device resources)");
    JB_CUDA(cudaSetDevice(p->device));
    JB_CUDA(cudaMemsetAsync(p->arena + p->acc_off, 0, sizeof(double2) * p->result_elems, p->stream));
    SetStateKernel<<<1, 1, 0, p->stream>>>(p->At<DeviceState>(p->state_off), 0, -1, 1);
    JB_CUDA(cudaGetLastError());
    return RunShared(p);
}

int jb_plan_run(jb_plan *p, int64_t first_slice, int64_t count)
{
    JB_REQUIRE(p, "plan: null argument");
    JB_REQUIRE(first_slice >= 0 && count >= 0 && first_slice + count <= p->num_slices,
               "plan: slice range out of bounds");
    return RunSlices(p, first_slice, -1, count);
}

int jb_plan_run_list(jb_plan *p, const int64_t *ids, int64_t count)
{
    JB_REQUIRE(p && (ids || count == 0), "plan: null argument");
    JB_REQUIRE(count <= p->list_cap, "plan: slice list too long (at most 65536 ids per call)");
    for (int64_t i = 0; i < count; i++)
        JB_REQUIRE(ids[i] >= 0 && ids[i] < p->num_slices, "plan: slice id out of bounds");
    JB_CUDA(cudaSetDevice(p->device));
    // Two pinned halves used alternately: the copy out of a half may still be in flight from the call
    // before last (waited on through its event); the device-side list itself is rewritten in stream
    // order, after the slices of the previous call have consumed it.
    const int half = p->list_half;
    p->list_half ^= 1;
    JB_CUDA(cudaEventSynchronize(p->ev_list[half]));
    long long *h = p->h_list + static_cast<size_t>(half) * p->list_cap;
    for (int64_t i = 0; i < count; i++)
        h[i] = ids[i];
    if (count > 0)
        JB_CUDA(cudaMemcpyAsync(p->arena + p->list_off, h, sizeof(long long) * count, cudaMemcpyHostToDevice,
                                p->stream));
    JB_CUDA(cudaEventRecord(p->ev_list[half], p->stream));
    return RunSlices(p, 0, 0, count);
}

int jb_plan_sync(jb_plan *p)
{
    JB_REQUIRE(p, "plan: null argument");
    JB_REQUIRE(p->arena != nullptr, "plan: created with JB_PLAN_DRY_RUN (no device resources)");
    JB_CUDA(cudaSetDevice(p->device));
    JB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int jb_plan_result(jb_plan *p, double *h_out)
{
    JB_REQUIRE(p && h_out, "plan: null argument");
    JB_REQUIRE(p->arena != nullptr, "plan: created with JB_PLAN_DRY_RUN (no device resources)");
    JB_CUDA(cudaSetDevice(p->device));
    JB_CUDA(cudaMemcpyAsync(h_out, p->arena + p->acc_off, sizeof(double2) * p->result_elems,
                            cudaMemcpyDeviceToHost, p->stream));
    JB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int jb_plan_slice_result(jb_plan *p, int64_t ordinal, void *h_out)
{
    JB_REQUIRE(p && h_out, "plan: null argument");
    JB_REQUIRE(p->store_cap > 0, "plan: created without JB_PLAN_STORE_RESULTS");
    JB_REQUIRE(ordinal >= 0 && ordinal < p->store_cap, "plan: slice ordinal out of range");
    JB_CUDA(cudaSetDevice(p->device));
    JB_CUDA(cudaMemcpyAsync(h_out, p->arena + p->store_off + p->eb * p->result_elems * ordinal,
                            p->eb * p->result_elems, cudaMemcpyDeviceToHost, p->stream));
    JB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int jb_plan_slice_results(jb_plan *p, int64_t first_ordinal, int64_t count, void *h_out)
{
    JB_REQUIRE(p && (h_out || count == 0), "plan: null argument");
    JB_REQUIRE(p->store_cap > 0, "plan: created without JB_PLAN_STORE_RESULTS");
    JB_REQUIRE(first_ordinal >= 0 && count >= 0 && first_ordinal + count <= p->store_cap,
               "plan: slice ordinal out of range");
    if (count == 0)
        return 0;
    JB_CUDA(cudaSetDevice(p->device));
    JB_CUDA(cudaMemcpyAsync(h_out, p->arena + p->store_off + p->eb * p->result_elems * first_ordinal,
                            p->eb * p->result_elems * count, cudaMemcpyDeviceToHost, p->stream));
    JB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int jb_plan_node(jb_plan *p, int32_t node, void *h_out, int64_t *elems)
{
    JB_REQUIRE(p, "plan: null argument");
    JB_REQUIRE(node >= 0 && node < static_cast<int>(p->nodes.size()), "plan: node out of range");
    JB_REQUIRE((p->flags & JB_PLAN_KEEP_INTERMEDIATES) || node == p->result_node ||
                   p->nodes[node].is_leaf,
               "plan: created without JB_PLAN_KEEP_INTERMEDIATES");
    const Node &n = p->nodes[node];
    if (elems)
        *elems = n.elems;
    if (h_out) {
        JB_CUDA(cudaSetDevice(p->device));
        JB_CUDA(cudaMemcpyAsync(h_out, NodePtr(p, node), p->eb * n.elems, cudaMemcpyDeviceToHost, p->stream));
        JB_CUDA(cudaStreamSynchronize(p->stream));
    }
    return 0;
}

int jb_plan_node_info(const jb_plan *p, int32_t node, int32_t *rank, int32_t *modes, int64_t *extent)
{
    JB_REQUIRE(p && rank, "plan: null argument");
    JB_REQUIRE(node >= 0 && node < static_cast<int>(p->nodes.size()), "plan: node out of range");
    const Node &n = p->nodes[node];
    *rank = static_cast<int32_t>(n.modes.size());
    for (size_t j = 0; j < n.modes.size(); j++) {
        if (modes)
            modes[j] = n.modes[j];
        if (extent)
            extent[j] = n.extent[j];
    }
    return 0;
}

int jb_plan_last_ms(jb_plan *p, float *ms)
{
    JB_REQUIRE(p && ms, "plan: null argument");
    JB_REQUIRE(p->have_run, "plan: nothing has been run yet");
    JB_CUDA(cudaSetDevice(p->device));
    JB_CUDA(cudaEventSynchronize(p->ev1));
    JB_CUDA(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return 0;
}

int jb_plan_stream(jb_plan *p, void **stream)
{
    JB_REQUIRE(p && stream, "plan: null argument");
    *stream = p->stream;
    return 0;
}

int jb_plan_accumulator(jb_plan *p, void **d_acc, int64_t *elems)
{
    JB_REQUIRE(p && d_acc, "plan: null argument");
    *d_acc = p->arena + p->acc_off;
    if (elems)
        *elems = p->result_elems;
    return 0;
}

int jb_plan_device(const jb_plan *p, int *device)
{
    JB_REQUIRE(p && device, "plan: null argument");
    *device = p->device;
    return 0;
}

int jb_plan_steps(const jb_plan *p, jb_step_info_t *steps, int32_t cap, int32_t *count)
{
    JB_REQUIRE(p && count, "plan: null argument");
    *count = static_cast<int32_t>(p->steps.size());
    for (int32_t i = 0; i < std::min<int32_t>(cap, *count); i++) {
        const Step &st = p->steps[i];
        jb_step_info_t &o = steps[i];
        o.node_a = st.a;
        o.node_b = st.b;
        o.node_c = st.c;
        o.shared = st.shared ? 1 : 0;
        o.kernel = st.cp.kernel;
        o.m = st.cp.m;
        o.n = st.cp.n;
        o.k = st.cp.k;
        o.flops = st.cp.flops();
        o.bytes = st.cp.bytes();
        o.op = st.op;
        o.pad = 0;
    }
    return 0;
}

int jb_plan_ops(const jb_plan *p, jb_op_info_t *ops, int32_t cap, int32_t *count)
{
    JB_REQUIRE(p && count, "plan: null argument");
    *count = static_cast<int32_t>(p->ops.size());
    for (int32_t i = 0; i < std::min<int32_t>(cap, *count); i++) {
        const Op &op = p->ops[i];
        jb_op_info_t &o = ops[i];
        o.kernel = op.kernel;
        o.n_steps = static_cast<int32_t>(op.steps.size());
        o.first_step = op.steps.front();
        o.last_step = op.steps.back();
        o.log_tile = op.kernel == 2 ? op.chain.log_tile : 0;
        o.n_stages = op.kernel == 2 ? op.chain.n_stages : 0;
        o.gemm_kind = op.kernel == 1 ? p->steps[op.steps[0]].cp.gemm_kind : 0;
        o.flops = o.bytes = o.step_bytes = 0.0;
        o.register_steps = op.kernel == 2 ? op.chain.register_steps : 0;
        o.pad = 0;
        if (op.kernel == 2) {
            o.launches = op.chain.launches;
            o.flops = op.chain.flops;
            o.bytes = op.chain.bytes;
            o.step_bytes = op.chain.step_bytes;
        }
        else {
            const Step &st = p->steps[op.steps[0]];
            o.launches = st.cp.launches;
            o.flops = st.cp.flops();
            o.bytes = o.step_bytes = st.cp.bytes();
        }
    }
    return 0;
}

int jb_plan_profile_ops(jb_plan *p, int64_t slice, int reps, float *ms, int32_t cap)
{
    JB_REQUIRE(p && ms, "plan: null argument");
    JB_REQUIRE(slice >= 0 && slice < p->num_slices, "plan: slice id out of bounds");
    JB_REQUIRE(p->arena != nullptr, "plan: created with JB_PLAN_DRY_RUN (no device resources)");
    JB_CUDA(cudaSetDevice(p->device));
    JB_TRY(RunShared(p));
    for (int32_t i = 0; i < cap; i++)
        ms[i] = 0.f;
    // one eager pass leaves every per-slice input in place (lifetimes are respected because the
    // units are replayed in plan order)
    // JB_PROFILE_BATCH=1: time the launch units as the batched graph runs them (`batch` slices per launch, starting
    // at `slice`); the default times one slice per launch
    const char *pb = getenv("JB_PROFILE_BATCH");
    const int prof_batch = (pb && pb[0] == '1' && slice + p->batch <= p->num_slices) ? p->batch : 1;
    std::vector<cudaEvent_t> ev(p->ops.size() + 1);
    for (auto &e : ev)
        JB_CUDA(cudaEventCreate(&e));
    std::vector<double> total(p->ops.size(), 0.0);
    for (int r = 0; r < reps + 1; r++) {
        SetStateKernel<<<1, 1, 0, p->stream>>>(p->At<DeviceState>(p->state_off), slice, -1, 0);
        JB_TRY(LaunchSliceViews(p, prof_batch));
        for (size_t i = 0; i < p->ops.size(); i++) {
            JB_CUDA(cudaEventRecord(ev[i], p->stream));
            JB_TRY(LaunchOp(p, p->ops[i], prof_batch));
        }
        JB_CUDA(cudaEventRecord(ev.back(), p->stream));
        JB_CUDA(cudaStreamSynchronize(p->stream));
        if (r == 0)
            continue; // warm-up
        for (size_t i = 0; i < p->ops.size(); i++) {
            float t = 0.f;
            JB_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            total[i] += t;
        }
    }
    for (size_t i = 0; i < p->ops.size(); i++)
        if (static_cast<int32_t>(i) < cap)
            ms[i] = static_cast<float>(total[i] / std::max(reps, 1));
    for (auto &e : ev)
        cudaEventDestroy(e);
    return 0;
}

int jb_plan_profile(jb_plan *p, int64_t slice, int reps, float *ms, int32_t cap)
{
    JB_REQUIRE(p && ms, "plan: null argument");
    std::vector<float> per_op(p->ops.size() + 1, 0.f);
    JB_TRY(jb_plan_profile_ops(p, slice, reps, per_op.data(), static_cast<int32_t>(p->ops.size())));
    for (int32_t i = 0; i < cap; i++)
        ms[i] = 0.f;
    // a fused chain's time is attributed to its last step
    for (size_t i = 0; i < p->ops.size(); i++) {
        const int s = p->ops[i].steps.back();
        if (s < cap)
            ms[s] = per_op[i];
    }
    return 0;
}

} // extern "C"
