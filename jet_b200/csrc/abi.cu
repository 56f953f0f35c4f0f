// extern "C" entry points of libjetb200.so (see include/jetb200.h for the contract and the
// reference interfaces each one replaces).
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace jb {

std::string &LastError()
{
    thread_local std::string err;
    return err;
}

int Fail(const std::string &msg)
{
    LastError() = msg;
    return 1;
}

int NumSMs()
{
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

// Occupancy and the >48 KB dynamic shared-memory opt-in are properties of (device, kernel): a process that
// drives several GPUs (jb_network_desc_t.device, jb_set_device) must query / set them on each one.
int BlocksPerSM(const void *kernel, int threads, size_t dyn_smem)
{
    static std::mutex mu;
    static std::map<std::tuple<int, const void *, size_t>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_tuple(dev, kernel, dyn_smem);
    auto it = cache.find(key);
    if (it != cache.end())
        return it->second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess || n < 1)
        n = 1;
    cache[key] = n;
    return n;
}

int EnsureDynamicSmem(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> done;
    int dev = 0;
    JB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    size_t &have = done[std::make_pair(dev, kernel)];
    if (have >= bytes)
        return 0;
    JB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    have = bytes;
    return 0;
}

namespace {

// Grow-only, per-thread device scratch for the host-buffer entry points: Jet::Tensor operators call
// these many times on tiny tensors, so a cudaMalloc/cudaFree per call would dominate.
struct ScratchSlot {
    void *p = nullptr;
    size_t cap = 0;
    int device = -1;
};
struct ScratchPool {
    ScratchSlot slot[4];
    ~ScratchPool()
    {
        for (auto &s : slot)
            if (s.p)
                cudaFree(s.p); // best effort (the context may already be gone at exit)
    }
};
struct DevBuf {
    void *p = nullptr;
    int index;
    explicit DevBuf(int i) : index(i) {}
    int Alloc(size_t bytes)
    {
        thread_local ScratchPool pool;
        ScratchSlot &s = pool.slot[index];
        int dev = 0;
        JB_CUDA(cudaGetDevice(&dev));
        bytes = bytes == 0 ? 16 : bytes;
        if (s.p == nullptr || s.cap < bytes || s.device != dev) {
            if (s.p) {
                cudaFree(s.p);
                s.p = nullptr;
                s.cap = 0;
            }
            const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
            JB_CUDA(cudaMalloc(&s.p, want));
            s.cap = want;
            s.device = dev;
        }
        p = s.p;
        return 0;
    }
};

int64_t Product(int rank, const int64_t *e)
{
    int64_t n = 1;
    for (int i = 0; i < rank; i++)
        n *= e[i];
    return n;
}

} // namespace
} // namespace jb

using namespace jb;

extern "C" {

const char *jb_last_error(void) { return LastError().c_str(); }
const char *jb_version(void) { return "jetb200 0.1 (sm_100a; drop-in for Jet 0.2.3-dev contraction path)"; }

int jb_device_count(int *count)
{
    JB_REQUIRE(count, "null argument");
    JB_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int jb_set_device(int device)
{
    JB_CUDA(cudaSetDevice(device));
    return 0;
}

int jb_device_info(int device, int *sm_count, size_t *total_bytes, size_t *l2_bytes, int *cc_major,
                   int *cc_minor)
{
    cudaDeviceProp prop;
    JB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count)
        *sm_count = prop.multiProcessorCount;
    if (total_bytes)
        *total_bytes = prop.totalGlobalMem;
    if (l2_bytes)
        *l2_bytes = static_cast<size_t>(prop.l2CacheSize);
    if (cc_major)
        *cc_major = prop.major;
    if (cc_minor)
        *cc_minor = prop.minor;
    return 0;
}

int jb_malloc(void **d_ptr, size_t bytes)
{
    JB_REQUIRE(d_ptr, "null argument");
    JB_CUDA(cudaMalloc(d_ptr, bytes == 0 ? 1 : bytes));
    return 0;
}
int jb_free(void *d_ptr)
{
    JB_CUDA(cudaFree(d_ptr));
    return 0;
}
int jb_host_alloc(void **h_ptr, size_t bytes)
{
    JB_REQUIRE(h_ptr, "null argument");
    JB_CUDA(cudaMallocHost(h_ptr, bytes == 0 ? 1 : bytes));
    return 0;
}
int jb_host_free(void *h_ptr)
{
    JB_CUDA(cudaFreeHost(h_ptr));
    return 0;
}
int jb_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    JB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice,
                            static_cast<cudaStream_t>(stream)));
    return 0;
}
int jb_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream)
{
    JB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost,
                            static_cast<cudaStream_t>(stream)));
    return 0;
}
int jb_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream)
{
    JB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice,
                            static_cast<cudaStream_t>(stream)));
    return 0;
}
int jb_memset_zero(void *d_dst, size_t bytes, void *stream)
{
    JB_CUDA(cudaMemsetAsync(d_dst, 0, bytes, static_cast<cudaStream_t>(stream)));
    return 0;
}
int jb_stream_create(void **stream)
{
    JB_REQUIRE(stream, "null argument");
    cudaStream_t s;
    JB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return 0;
}
int jb_stream_destroy(void *stream)
{
    JB_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
    return 0;
}
int jb_stream_sync(void *stream)
{
    JB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return 0;
}

int jb_permute(int dtype, const void *d_in, void *d_out, int rank, const int64_t *extent_in,
               const int32_t *perm, void *stream)
{
    JB_REQUIRE(d_in && d_out, "permute: null buffer");
    JB_REQUIRE(rank == 0 || (extent_in && perm), "permute: null shape");
    return LaunchPermute(dtype, d_in, d_out, rank, extent_in, perm, static_cast<cudaStream_t>(stream));
}

size_t jb_gemm_ws_bytes(int dtype, int64_t m, int64_t n, int64_t k)
{
    return GemmWorkspaceBytes(dtype, m, n, k);
}

int jb_gemm(int dtype, int64_t m, int64_t n, int64_t k, const void *d_a, const void *d_b,
            void *d_c, void *d_ws, size_t ws_bytes, void *stream)
{
    JB_REQUIRE(d_a && d_b && d_c, "gemm: null buffer");
    return LaunchGemm(dtype, m, n, k, d_a, d_b, d_c, d_ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int jb_contract_info(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                     int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                     jb_contract_info_t *info)
{
    JB_REQUIRE(info, "contract: null argument");
    ContractPlan P;
    JB_TRY(MakeContractPlan(dtype, rank_a, extent_a, modes_a, rank_b, extent_b, modes_b, &P));
    info->rank_c = P.rank_c;
    for (int i = 0; i < P.rank_c; i++) {
        info->modes_c[i] = P.modes_c[i];
        info->extent_c[i] = P.extent_c[i];
    }
    info->m = P.m;
    info->n = P.n;
    info->k = P.k;
    info->ws_bytes = P.ws_bytes;
    info->kernel = P.kernel;
    return 0;
}

int jb_contract(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                const void *d_a, int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                const void *d_b, void *d_c, void *d_ws, size_t ws_bytes, void *stream)
{
    JB_REQUIRE(d_a && d_b && d_c, "contract: null buffer");
    ContractPlan P;
    JB_TRY(MakeContractPlan(dtype, rank_a, extent_a, modes_a, rank_b, extent_b, modes_b, &P));
    JB_REQUIRE(ws_bytes >= P.ws_bytes, "contract: workspace too small (see jb_contract_info)");
    return LaunchContract(P, d_a, d_b, d_c, d_ws, static_cast<cudaStream_t>(stream));
}

// ---- fused contraction chain ---------------------------------------------------------------------
namespace {
int BuildChain(const jb_chain_desc_t *d, ChainOp *op)
{
    JB_REQUIRE(d != nullptr, "chain: null argument");
    JB_REQUIRE(d->dtype == JB_C64 || d->dtype == JB_C128, "chain: unknown dtype");
    JB_REQUIRE(d->n_steps >= 1, "chain: at least one step is required");
    JB_REQUIRE(d->rank_x >= 0 && d->rank_x <= JB_MAX_RANK, "chain: rank out of range");
    std::vector<int32_t> mx(d->modes_x, d->modes_x + d->rank_x);
    std::vector<int64_t> ex(d->extent_x, d->extent_x + d->rank_x);
    std::vector<ChainOperand> ops;
    size_t off = 0;
    for (int s = 0; s < d->n_steps; s++) {
        const int r = d->rank_r[s];
        JB_REQUIRE(r >= 0 && r <= JB_MAX_RANK, "chain: rank out of range");
        ChainOperand o;
        o.modes.assign(d->modes_r + off, d->modes_r + off + r);
        o.extent.assign(d->extent_r + off, d->extent_r + off + r);
        o.x_is_left = d->x_is_left[s] != 0;
        off += r;
        ops.push_back(o);
    }
    std::string why;
    if (MakeChainOp(d->dtype, mx, ex, ops, ChainMaxTileBits(d->dtype), op, &why) != 0)
        return Fail("chain: " + why);
    return 0;
}
} // namespace

int jb_chain_info(const jb_chain_desc_t *desc, jb_chain_info_t *info)
{
    JB_REQUIRE(info, "chain: null argument");
    ChainOp op;
    JB_TRY(BuildChain(desc, &op));
    info->rank_c = static_cast<int32_t>(op.modes_c.size());
    for (size_t i = 0; i < op.modes_c.size(); i++) {
        info->modes_c[i] = op.modes_c[i];
        info->extent_c[i] = op.extent_c[i];
    }
    info->log_tile = op.log_tile;
    info->conflict_free = op.conflict_free;
    info->n_stages = op.n_stages;
    info->pad = 0;
    info->flops = op.flops;
    info->bytes = op.bytes;
    info->step_bytes = op.step_bytes;
    return 0;
}

namespace {
// Operator-level chains share one constant-bank slot per device: calls are serialised with a mutex
// and, across streams, with an event recorded after each launch.
struct OperatorChainState {
    cudaEvent_t done = nullptr;
};
std::mutex g_op_chain_mutex;
OperatorChainState g_op_chain[64];

int LaunchOperatorChain(const ChainOp &op, const void *d_x, const void *const *d_r, void *d_out,
                        cudaStream_t stream)
{
    int dev = 0;
    JB_CUDA(cudaGetDevice(&dev));
    JB_REQUIRE(dev >= 0 && dev < 64, "chain: device index out of range");
    std::lock_guard<std::mutex> lock(g_op_chain_mutex);
    OperatorChainState &st = g_op_chain[dev];
    if (st.done == nullptr) {
        JB_CUDA(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    }
    else {
        JB_CUDA(cudaStreamWaitEvent(stream, st.done, 0)); // the previous call may be on another stream
    }
    JB_TRY(LaunchChain(op, d_x, d_r, d_out, ChainOperatorSlot(), stream));
    JB_CUDA(cudaEventRecord(st.done, stream));
    return 0;
}
} // namespace

int jb_contract_chain(const jb_chain_desc_t *desc, const void *d_x, const void *const *d_r,
                      void *d_out, void *stream)
{
    JB_REQUIRE(d_x && d_r && d_out, "chain: null buffer");
    ChainOp op;
    JB_TRY(BuildChain(desc, &op));
    return LaunchOperatorChain(op, d_x, d_r, d_out, static_cast<cudaStream_t>(stream));
}

int jb_contract_chain_host(const jb_chain_desc_t *desc, const void *h_x, const void *const *h_r,
                           void *h_out)
{
    JB_REQUIRE(h_x && h_r && h_out, "chain: null buffer");
    ChainOp op;
    JB_TRY(BuildChain(desc, &op));
    const size_t eb = ElemBytes(desc->dtype);
    const size_t bx = eb * static_cast<size_t>(Product(desc->rank_x, desc->extent_x));
    size_t bout = eb;
    for (int64_t e : op.extent_c)
        bout *= static_cast<size_t>(e);
    DevBuf x(0), out(1), rs(2);
    JB_TRY(x.Alloc(bx));
    JB_TRY(out.Alloc(bout));
    std::vector<size_t> roff(desc->n_steps);
    size_t rtotal = 0, off = 0;
    for (int s = 0; s < desc->n_steps; s++) {
        roff[s] = rtotal;
        const size_t b = eb * static_cast<size_t>(Product(desc->rank_r[s], desc->extent_r + off));
        rtotal += (b + 255) & ~size_t(255);
        off += desc->rank_r[s];
    }
    JB_TRY(rs.Alloc(rtotal));
    JB_CUDA(cudaMemcpy(x.p, h_x, bx, cudaMemcpyHostToDevice));
    std::vector<const void *> rp(desc->n_steps);
    off = 0;
    for (int s = 0; s < desc->n_steps; s++) {
        const size_t b = eb * static_cast<size_t>(Product(desc->rank_r[s], desc->extent_r + off));
        off += desc->rank_r[s];
        JB_REQUIRE(h_r[s] != nullptr, "chain: null buffer");
        JB_CUDA(cudaMemcpy(static_cast<unsigned char *>(rs.p) + roff[s], h_r[s], b, cudaMemcpyHostToDevice));
        rp[s] = static_cast<unsigned char *>(rs.p) + roff[s];
    }
    JB_TRY(LaunchOperatorChain(op, x.p, rp.data(), out.p, nullptr));
    JB_CUDA(cudaMemcpy(h_out, out.p, bout, cudaMemcpyDeviceToHost));
    return 0;
}

int jb_add(int dtype, int64_t n, const void *d_a, const void *d_b, void *d_c, void *stream)
{
    return LaunchAdd(dtype, n, d_a, d_b, d_c, static_cast<cudaStream_t>(stream));
}

int jb_conj(int dtype, int64_t n, const void *d_in, void *d_out, void *stream)
{
    return LaunchConj(dtype, n, d_in, d_out, static_cast<cudaStream_t>(stream));
}

int jb_slice(int dtype, const void *d_in, void *d_out, int rank, const int64_t *extent_in,
             int axis, int64_t value, void *stream)
{
    return LaunchSlice(dtype, d_in, d_out, rank, extent_in, axis, value,
                       static_cast<cudaStream_t>(stream));
}

// ---- host-buffer forms ---------------------------------------------------------------------------
int jb_permute_host(int dtype, const void *h_in, void *h_out, int rank, const int64_t *extent_in,
                    const int32_t *perm)
{
    JB_REQUIRE(h_in && h_out, "permute: null buffer");
    const size_t bytes = ElemBytes(dtype) * static_cast<size_t>(Product(rank, extent_in));
    DevBuf in(0), out(1);
    JB_TRY(in.Alloc(bytes));
    JB_TRY(out.Alloc(bytes));
    JB_CUDA(cudaMemcpy(in.p, h_in, bytes, cudaMemcpyHostToDevice));
    JB_TRY(LaunchPermute(dtype, in.p, out.p, rank, extent_in, perm, nullptr));
    JB_CUDA(cudaMemcpy(h_out, out.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int jb_contract_host(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                     const void *h_a, int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                     const void *h_b, void *h_c)
{
    JB_REQUIRE(h_a && h_b && h_c, "contract: null buffer");
    ContractPlan P;
    JB_TRY(MakeContractPlan(dtype, rank_a, extent_a, modes_a, rank_b, extent_b, modes_b, &P));
    const size_t eb = ElemBytes(dtype);
    const size_t ba = eb * static_cast<size_t>(P.m * P.k), bb = eb * static_cast<size_t>(P.k * P.n),
                 bc = eb * static_cast<size_t>(P.m * P.n);
    DevBuf a(0), b(1), c(2), ws(3);
    JB_TRY(a.Alloc(ba));
    JB_TRY(b.Alloc(bb));
    JB_TRY(c.Alloc(bc));
    JB_TRY(ws.Alloc(P.ws_bytes));
    JB_CUDA(cudaMemcpy(a.p, h_a, ba, cudaMemcpyHostToDevice));
    JB_CUDA(cudaMemcpy(b.p, h_b, bb, cudaMemcpyHostToDevice));
    JB_TRY(LaunchContract(P, a.p, b.p, c.p, ws.p, nullptr));
    JB_CUDA(cudaMemcpy(h_c, c.p, bc, cudaMemcpyDeviceToHost));
    return 0;
}

int jb_gemm_host(int dtype, int64_t m, int64_t n, int64_t k, const void *h_a, const void *h_b,
                 void *h_c)
{
    JB_REQUIRE(h_a && h_b && h_c, "gemm: null buffer");
    JB_REQUIRE(m >= 1 && n >= 1 && k >= 1, "gemm: dimensions must be positive");
    const size_t eb = ElemBytes(dtype);
    DevBuf a(0), b(1), c(2), ws(3);
    const size_t wsb = GemmWorkspaceBytes(dtype, m, n, k);
    JB_TRY(a.Alloc(eb * m * k));
    JB_TRY(b.Alloc(eb * k * n));
    JB_TRY(c.Alloc(eb * m * n));
    JB_TRY(ws.Alloc(wsb));
    JB_CUDA(cudaMemcpy(a.p, h_a, eb * m * k, cudaMemcpyHostToDevice));
    JB_CUDA(cudaMemcpy(b.p, h_b, eb * k * n, cudaMemcpyHostToDevice));
    JB_TRY(LaunchGemm(dtype, m, n, k, a.p, b.p, c.p, ws.p, wsb, nullptr));
    JB_CUDA(cudaMemcpy(h_c, c.p, eb * m * n, cudaMemcpyDeviceToHost));
    return 0;
}

int jb_add_host(int dtype, int64_t n, const void *h_a, const void *h_b, void *h_c)
{
    JB_REQUIRE(h_a && h_b && h_c, "add: null buffer");
    const size_t bytes = ElemBytes(dtype) * static_cast<size_t>(n);
    DevBuf a(0), b(1), c(2);
    JB_TRY(a.Alloc(bytes));
    JB_TRY(b.Alloc(bytes));
    JB_TRY(c.Alloc(bytes));
    JB_CUDA(cudaMemcpy(a.p, h_a, bytes, cudaMemcpyHostToDevice));
    JB_CUDA(cudaMemcpy(b.p, h_b, bytes, cudaMemcpyHostToDevice));
    JB_TRY(LaunchAdd(dtype, n, a.p, b.p, c.p, nullptr));
    JB_CUDA(cudaMemcpy(h_c, c.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int jb_conj_host(int dtype, int64_t n, const void *h_in, void *h_out)
{
    JB_REQUIRE(h_in && h_out, "conj: null buffer");
    const size_t bytes = ElemBytes(dtype) * static_cast<size_t>(n);
    DevBuf in(0), out(1);
    JB_TRY(in.Alloc(bytes));
    JB_TRY(out.Alloc(bytes));
    JB_CUDA(cudaMemcpy(in.p, h_in, bytes, cudaMemcpyHostToDevice));
    JB_TRY(LaunchConj(dtype, n, in.p, out.p, nullptr));
    JB_CUDA(cudaMemcpy(h_out, out.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int jb_slice_host(int dtype, const void *h_in, void *h_out, int rank, const int64_t *extent_in,
                  int axis, int64_t value)
{
    JB_REQUIRE(h_in && h_out, "slice: null buffer");
    JB_REQUIRE(axis >= 0 && axis < rank, "slice: axis out of range");
    const size_t eb = ElemBytes(dtype);
    const size_t n_in = static_cast<size_t>(Product(rank, extent_in));
    const size_t n_out = n_in / static_cast<size_t>(extent_in[axis]);
    DevBuf in(0), out(1);
    JB_TRY(in.Alloc(eb * n_in));
    JB_TRY(out.Alloc(eb * n_out));
    JB_CUDA(cudaMemcpy(in.p, h_in, eb * n_in, cudaMemcpyHostToDevice));
    JB_TRY(LaunchSlice(dtype, in.p, out.p, rank, extent_in, axis, value, nullptr));
    JB_CUDA(cudaMemcpy(h_out, out.p, eb * n_out, cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"
