// Slices of one network spread over several plans: `lanes` plans per device (several slices in flight on one
// GPU, the stream analogue of the reference running tasks of different slices on Taskflow workers,
// include/jet/TaskBasedContractor.hpp:322) on any number of devices of ONE process, and NCCL communicators for
// one-process-per-GPU jobs.  The sum over slices (AddReductionTask, TaskBasedContractor.hpp:258-280) never
// leaves the devices: every plan accumulates its share in FP64, the partial sums are added on the device in a
// fixed order (lane order, then device order; peer copies over NVLink), and across processes one ncclReduce on
// the plan's stream finishes the job.
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only: the library is bound at run time (dlopen), never linked

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>

#include "common.cuh"

namespace jb {
namespace {

__global__ void __launch_bounds__(256) AddDoublesKernel(double *__restrict__ dst, const double *__restrict__ src, long long n)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] += src[i];
}

int AddDoubles(double *dst, const double *src, long long n, cudaStream_t stream)
{
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 4ll * NumSMs())));
    AddDoublesKernel<<<grid, 256, 0, stream>>>(dst, src, n);
    JB_CUDA(cudaGetLastError());
    return 0;
}

// ---- NCCL, bound at run time -----------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi &Nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a process that already holds libnccl.so.2 (torch.distributed) gets that copy back; JB_NCCL_LIB overrides
        const char *names[] = {getenv("JB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *name : names) {
            if (name == nullptr || name[0] == 0)
                continue;
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle)
                break;
        }
        if (!api.handle) {
            api.error = "NCCL is not available: dlopen(libnccl.so.2) failed; set JB_NCCL_LIB to its path";
            return;
        }
        bool ok = true;
        auto bind = [&](auto &fn, const char *sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, sym));
            if (!fn) {
                ok = false;
                api.error = std::string("NCCL symbol missing: ") + sym;
            }
        };
        bind(api.GetUniqueId, "ncclGetUniqueId");
        bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommInitAll, "ncclCommInitAll");
        bind(api.CommDestroy, "ncclCommDestroy");
        bind(api.Reduce, "ncclReduce");
        bind(api.AllReduce, "ncclAllReduce");
        bind(api.GroupStart, "ncclGroupStart");
        bind(api.GroupEnd, "ncclGroupEnd");
        bind(api.GetErrorString, "ncclGetErrorString");
        bind(api.GetVersion, "ncclGetVersion");
        if (!ok) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api;
}

#define JB_NCCL(expr)                                                                             \
    do {                                                                                          \
        ncclResult_t jb_nr__ = (expr);                                                            \
        if (jb_nr__ != ncclSuccess)                                                               \
            return ::jb::Fail(std::string(#expr) + ": " + Nccl().GetErrorString(jb_nr__));        \
    } while (0)

} // namespace
} // namespace jb

using namespace jb;

struct jb_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
};

struct jb_multi {
    std::vector<int> devices;
    int lanes = 1;
    std::vector<jb_plan *> plans; // device-major: plans[d * lanes + l]
    jb_plan_stats_t stats;
    int64_t result_elems = 1;
    size_t elem_bytes = 8;
    double *d_total = nullptr;   // on devices[0]: the sum over all plans (2 * result_elems doubles)
    double *d_staging = nullptr; // on devices[0]: one partial sum per other device
    std::vector<cudaEvent_t> ev;  // one per plan: "this plan's slices (and its device-local sum) are done"
    std::vector<int64_t> first_ordinal; // ordinal (position in the run since reset) of each plan's first slice
    std::vector<int64_t> counts;
    bool total_valid = false;
    bool peer_checked = false;
};

namespace {

int MultiStream(jb_multi *m, size_t i, cudaStream_t *s)
{
    void *v = nullptr;
    JB_TRY(jb_plan_stream(m->plans[i], &v));
    *s = static_cast<cudaStream_t>(v);
    return 0;
}

// total = sum over plans, computed on the devices in a fixed order: the lanes of devices[0] are added on its
// lane-0 stream; every accumulator of the other devices is copied to devices[0] (NVLink peer copy; result-sized,
// 16 bytes for an amplitude) and added in (device, lane) order.  The accumulators themselves are left untouched,
// so more slices can be run afterwards without a reset.
int ComputeTotal(jb_multi *m)
{
    if (m->total_valid)
        return 0;
    const long long n = 2 * m->result_elems;
    const size_t bytes = sizeof(double) * static_cast<size_t>(n);
    const int nd = static_cast<int>(m->devices.size());
    cudaStream_t s0 = nullptr;
    JB_TRY(MultiStream(m, 0, &s0));
    for (int d = 0; d < nd; d++) {
        JB_CUDA(cudaSetDevice(m->devices[d]));
        cudaStream_t sd = nullptr;
        JB_TRY(MultiStream(m, static_cast<size_t>(d) * m->lanes, &sd));
        void *acc0 = nullptr;
        JB_TRY(jb_plan_accumulator(m->plans[static_cast<size_t>(d) * m->lanes], &acc0, nullptr));
        if (d == 0) {
            JB_CUDA(cudaMemcpyAsync(m->d_total, acc0, bytes, cudaMemcpyDeviceToDevice, sd));
            for (int l = 1; l < m->lanes; l++) {
                const size_t i = static_cast<size_t>(l);
                cudaStream_t sl = nullptr;
                JB_TRY(MultiStream(m, i, &sl));
                JB_CUDA(cudaEventRecord(m->ev[i], sl));
                JB_CUDA(cudaStreamWaitEvent(sd, m->ev[i], 0));
                void *accl = nullptr;
                JB_TRY(jb_plan_accumulator(m->plans[i], &accl, nullptr));
                JB_TRY(AddDoubles(m->d_total, static_cast<const double *>(accl), n, sd));
            }
            continue;
        }
        // devices 1..: lanes are copied to devices[0] one by one (each is result-sized: 16 bytes for an amplitude)
        for (int l = 0; l < m->lanes; l++) {
            const size_t i = static_cast<size_t>(d) * m->lanes + l;
            cudaStream_t sl = nullptr;
            JB_TRY(MultiStream(m, i, &sl));
            void *accl = nullptr;
            JB_TRY(jb_plan_accumulator(m->plans[i], &accl, nullptr));
            double *slot = m->d_staging + static_cast<size_t>(n) * (static_cast<size_t>(d - 1) * m->lanes + l);
            JB_CUDA(cudaMemcpyPeerAsync(slot, m->devices[0], accl, m->devices[d], bytes, sl));
            JB_CUDA(cudaEventRecord(m->ev[i], sl));
        }
    }
    JB_CUDA(cudaSetDevice(m->devices[0]));
    for (int d = 1; d < nd; d++)
        for (int l = 0; l < m->lanes; l++) {
            const size_t i = static_cast<size_t>(d) * m->lanes + l;
            JB_CUDA(cudaStreamWaitEvent(s0, m->ev[i], 0));
            JB_TRY(AddDoubles(m->d_total, m->d_staging + static_cast<size_t>(n) * (static_cast<size_t>(d - 1) * m->lanes + l), n, s0));
        }
    m->total_valid = true;
    return 0;
}

} // namespace

extern "C" {

int jb_multi_destroy(jb_multi *m)
{
    if (m == nullptr)
        return 0;
    for (jb_plan *p : m->plans)
        jb_plan_destroy(p);
    if (!m->devices.empty())
        cudaSetDevice(m->devices[0]);
    if (m->d_total)
        cudaFree(m->d_total);
    if (m->d_staging)
        cudaFree(m->d_staging);
    for (size_t i = 0; i < m->ev.size(); i++)
        if (m->ev[i]) {
            cudaSetDevice(m->devices[i / m->lanes]);
            cudaEventDestroy(m->ev[i]);
        }
    delete m;
    return 0;
}

int jb_multi_create(const jb_network_desc_t *desc, int num_devices, const int *devices, int lanes, jb_multi **out)
{
    JB_REQUIRE(desc && out, "multi: null argument");
    JB_REQUIRE(lanes >= 0 && lanes <= 5, "multi: lanes must be in 0..5 (0 = automatic)");
    JB_REQUIRE(num_devices >= 1 && num_devices <= 64, "multi: device count out of range");
    struct Deleter {
        void operator()(jb_multi *q) const { jb_multi_destroy(q); }
    };
    std::unique_ptr<jb_multi, Deleter> up(new jb_multi());
    jb_multi *m = up.get();
    for (int d = 0; d < num_devices; d++)
        m->devices.push_back(devices ? devices[d] : desc->device + d);
    for (size_t a = 0; a < m->devices.size(); a++)
        for (size_t b = a + 1; b < m->devices.size(); b++)
            JB_REQUIRE(m->devices[a] != m->devices[b], "multi: a device is listed twice");
    // the network is planned ONCE; every other (device, lane) clones the host-side plan and gets its own
    // arena, stream and CUDA graphs
    jb_network_desc_t d0 = *desc;
    d0.device = m->devices[0];
    jb_plan *first = nullptr;
    JB_TRY(jb_plan_create(&d0, &first));
    m->plans.push_back(first);
    JB_TRY(jb_plan_stats(first, &m->stats));
    if (lanes == 0) {
        // several slices in flight pay off while one slice cannot fill the GPU (launch-latency bound): small arenas
        const size_t arena = m->stats.arena_bytes;
        lanes = arena <= (size_t(512) << 20) ? 4 : arena <= (size_t(8) << 30) ? 2 : 1;
        size_t free_b = 0, total_b = 0;
        JB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        while (lanes > 1 && arena * static_cast<size_t>(lanes - 1) > free_b / 2)
            lanes--;
        const int64_t per_device = (m->stats.num_slices + num_devices - 1) / num_devices;
        lanes = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(lanes, per_device)));
    }
    m->lanes = lanes;
    for (size_t i = 1; i < m->devices.size() * static_cast<size_t>(lanes); i++) {
        jb_plan *p = nullptr;
        JB_TRY(jb_plan_clone(first, m->devices[i / lanes], &p));
        m->plans.push_back(p);
        if (desc->h_data != nullptr)
            JB_TRY(jb_plan_upload(p, desc->h_data));
    }
    m->result_elems = m->stats.result_elems;
    m->elem_bytes = ElemBytes(desc->dtype);
    m->ev.assign(m->plans.size(), nullptr);
    for (size_t i = 0; i < m->plans.size(); i++) {
        JB_CUDA(cudaSetDevice(m->devices[i / lanes]));
        JB_CUDA(cudaEventCreateWithFlags(&m->ev[i], cudaEventDisableTiming));
    }
    JB_CUDA(cudaSetDevice(m->devices[0]));
    const size_t bytes = sizeof(double) * 2 * static_cast<size_t>(m->result_elems);
    JB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->d_total), bytes));
    if (m->devices.size() > 1) {
        JB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->d_staging), bytes * (m->devices.size() - 1) * lanes));
        // NVLink peer access where the topology offers it (cudaMemcpyPeerAsync stages through the host otherwise)
        for (size_t d = 1; d < m->devices.size(); d++) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->devices[d], m->devices[0]) == cudaSuccess && can) {
                cudaSetDevice(m->devices[d]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[0], 0);
                if (e != cudaSuccess)
                    cudaGetLastError(); // already enabled is fine
            }
        }
        JB_CUDA(cudaSetDevice(m->devices[0]));
    }
    m->first_ordinal.assign(m->plans.size(), 0);
    m->counts.assign(m->plans.size(), 0);
    *out = up.release();
    return 0;
}

int jb_multi_stats(const jb_multi *m, jb_plan_stats_t *stats)
{
    JB_REQUIRE(m && stats, "multi: null argument");
    *stats = m->stats; // arena_bytes is per plan: a device holds `lanes` of them
    return 0;
}

int jb_multi_num_plans(const jb_multi *m, int *num_devices, int *lanes)
{
    JB_REQUIRE(m, "multi: null argument");
    if (num_devices)
        *num_devices = static_cast<int>(m->devices.size());
    if (lanes)
        *lanes = m->lanes;
    return 0;
}

int jb_multi_plan(jb_multi *m, int index, jb_plan **plan)
{
    JB_REQUIRE(m && plan && index >= 0 && index < static_cast<int>(m->plans.size()), "multi: plan index out of range");
    *plan = m->plans[index];
    return 0;
}

int jb_multi_upload(jb_multi *m, const void *const *h_data)
{
    JB_REQUIRE(m && h_data, "multi: null argument");
    for (jb_plan *p : m->plans)
        JB_TRY(jb_plan_upload(p, h_data));
    return 0;
}

int jb_multi_reset(jb_multi *m)
{
    JB_REQUIRE(m, "multi: null argument");
    for (jb_plan *p : m->plans)
        JB_TRY(jb_plan_reset(p));
    std::fill(m->first_ordinal.begin(), m->first_ordinal.end(), 0);
    std::fill(m->counts.begin(), m->counts.end(), 0);
    m->total_valid = false;
    return 0;
}

// Plan i of P takes the i-th contiguous block of the run (devices first, lanes inside a device): with one
// device and one lane this is the slice order of a single plan.
static void BlockOf(int64_t count, size_t plans, size_t i, int64_t *lo, int64_t *hi)
{
    const int64_t block = (count + static_cast<int64_t>(plans) - 1) / static_cast<int64_t>(plans);
    *lo = std::min<int64_t>(count, static_cast<int64_t>(i) * block);
    *hi = std::min<int64_t>(count, static_cast<int64_t>(i + 1) * block);
}

int jb_multi_run(jb_multi *m, int64_t first_slice, int64_t count)
{
    JB_REQUIRE(m, "multi: null argument");
    JB_REQUIRE(first_slice >= 0 && count >= 0 && first_slice + count <= m->stats.num_slices,
               "plan: slice range out of bounds");
    JB_TRY(jb_multi_reset(m));
    int64_t ordinal = 0;
    for (size_t i = 0; i < m->plans.size(); i++) {
        int64_t lo, hi;
        BlockOf(count, m->plans.size(), i, &lo, &hi);
        m->first_ordinal[i] = ordinal;
        m->counts[i] = hi - lo;
        ordinal += hi - lo;
        if (hi > lo)
            JB_TRY(jb_plan_run(m->plans[i], first_slice + lo, hi - lo));
    }
    return 0;
}

int jb_multi_run_list(jb_multi *m, const int64_t *ids, int64_t count)
{
    JB_REQUIRE(m && (ids || count == 0), "multi: null argument");
    JB_TRY(jb_multi_reset(m));
    constexpr int64_t kChunk = 1 << 16; // jb_plan_run_list takes at most 2^16 ids per call
    int64_t ordinal = 0;
    for (size_t i = 0; i < m->plans.size(); i++) {
        int64_t lo, hi;
        BlockOf(count, m->plans.size(), i, &lo, &hi);
        m->first_ordinal[i] = ordinal;
        m->counts[i] = hi - lo;
        ordinal += hi - lo;
    }
    // chunks are dealt round-robin over the plans so that all of them are busy from the start
    bool more = true;
    for (int64_t c = 0; more; c++) {
        more = false;
        for (size_t i = 0; i < m->plans.size(); i++) {
            int64_t lo, hi;
            BlockOf(count, m->plans.size(), i, &lo, &hi);
            const int64_t a = lo + c * kChunk, b = std::min(hi, a + kChunk);
            if (a >= hi)
                continue;
            JB_TRY(jb_plan_run_list(m->plans[i], ids + a, b - a));
            more = more || b < hi;
        }
    }
    return 0;
}

int jb_multi_sync(jb_multi *m)
{
    JB_REQUIRE(m, "multi: null argument");
    for (jb_plan *p : m->plans)
        JB_TRY(jb_plan_sync(p));
    return 0;
}

int jb_multi_result(jb_multi *m, double *h_out)
{
    JB_REQUIRE(m && h_out, "multi: null argument");
    JB_TRY(ComputeTotal(m));
    cudaStream_t s0 = nullptr;
    JB_TRY(MultiStream(m, 0, &s0));
    JB_CUDA(cudaSetDevice(m->devices[0]));
    JB_CUDA(cudaMemcpyAsync(h_out, m->d_total, sizeof(double) * 2 * static_cast<size_t>(m->result_elems),
                            cudaMemcpyDeviceToHost, s0));
    JB_CUDA(cudaStreamSynchronize(s0));
    m->total_valid = false; // more slices may follow; the next read recomputes the sum
    return 0;
}

int jb_multi_slice_result(jb_multi *m, int64_t ordinal, void *h_out)
{
    JB_REQUIRE(m && h_out, "multi: null argument");
    for (size_t i = 0; i < m->plans.size(); i++)
        if (ordinal >= m->first_ordinal[i] && ordinal < m->first_ordinal[i] + m->counts[i])
            return jb_plan_slice_result(m->plans[i], ordinal - m->first_ordinal[i], h_out);
    return Fail("plan: slice ordinal out of range");
}

int jb_multi_slice_results(jb_multi *m, void *h_out)
{
    JB_REQUIRE(m && h_out, "multi: null argument");
    for (size_t i = 0; i < m->plans.size(); i++) {
        if (m->counts[i] == 0)
            continue;
        unsigned char *dst = static_cast<unsigned char *>(h_out) +
                             static_cast<size_t>(m->first_ordinal[i]) * static_cast<size_t>(m->result_elems) * m->elem_bytes;
        JB_TRY(jb_plan_slice_results(m->plans[i], 0, m->counts[i], dst));
    }
    return 0;
}

int jb_multi_last_ms(jb_multi *m, float *ms)
{
    JB_REQUIRE(m && ms, "multi: null argument");
    *ms = 0.f;
    for (size_t i = 0; i < m->plans.size(); i++) {
        if (m->counts[i] == 0)
            continue;
        float t = 0.f;
        JB_TRY(jb_plan_last_ms(m->plans[i], &t));
        *ms = std::max(*ms, t);
    }
    return 0;
}

// ---- communicators (one process per GPU) ---------------------------------------------------------------------
int jb_comm_unique_id(void *id128)
{
    JB_REQUIRE(id128, "comm: null argument");
    NcclApi &api = Nccl();
    JB_REQUIRE(api.handle, api.error);
    static_assert(sizeof(ncclUniqueId) == JB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    JB_NCCL(api.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}

int jb_comm_create(int world, int rank, const void *id128, int device, jb_comm **out)
{
    JB_REQUIRE(id128 && out && world >= 1 && rank >= 0 && rank < world, "comm: bad argument");
    NcclApi &api = Nccl();
    JB_REQUIRE(api.handle, api.error);
    JB_CUDA(cudaSetDevice(device));
    std::unique_ptr<jb_comm> c(new jb_comm());
    c->world = world;
    c->rank = rank;
    c->device = device;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    JB_NCCL(api.CommInitRank(&c->comm, world, id, rank));
    *out = c.release();
    return 0;
}

int jb_comm_destroy(jb_comm *c)
{
    if (c == nullptr)
        return 0;
    if (c->comm && Nccl().handle)
        Nccl().CommDestroy(c->comm);
    delete c;
    return 0;
}

int jb_comm_info(const jb_comm *c, int *world, int *rank, int *nccl_version)
{
    JB_REQUIRE(c, "comm: null argument");
    if (world)
        *world = c->world;
    if (rank)
        *rank = c->rank;
    if (nccl_version) {
        *nccl_version = 0;
        Nccl().GetVersion(nccl_version);
    }
    return 0;
}

int jb_reduce_sum(jb_comm *c, void *d_buf, int64_t n_doubles, int root, void *stream)
{
    JB_REQUIRE(c && d_buf && n_doubles >= 0, "comm: bad argument");
    NcclApi &api = Nccl();
    JB_REQUIRE(api.handle, api.error);
    JB_CUDA(cudaSetDevice(c->device));
    if (root < 0)
        JB_NCCL(api.AllReduce(d_buf, d_buf, static_cast<size_t>(n_doubles), ncclDouble, ncclSum, c->comm,
                              static_cast<cudaStream_t>(stream)));
    else
        JB_NCCL(api.Reduce(d_buf, d_buf, static_cast<size_t>(n_doubles), ncclDouble, ncclSum, root, c->comm,
                           static_cast<cudaStream_t>(stream)));
    return 0;
}

int jb_multi_reduce(jb_multi *m, jb_comm *c, int root)
{
    JB_REQUIRE(m && c, "multi: null argument");
    JB_REQUIRE(c->device == m->devices[0], "multi: the communicator must live on the first device of the set");
    JB_TRY(ComputeTotal(m));
    cudaStream_t s0 = nullptr;
    JB_TRY(MultiStream(m, 0, &s0));
    JB_TRY(jb_reduce_sum(c, m->d_total, 2 * m->result_elems, root, s0));
    return 0; // total_valid stays set: jb_multi_result now returns the reduced sum
}

} // extern "C"
