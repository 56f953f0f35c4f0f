// TEST INFRASTRUCTURE ONLY (oracle/): C entry points over the UNMODIFIED
// reference headers, compiled where they lie under /root/reference/include
// (recipe: oracle/Makefile).  The output (oracle/_ref/libjetref.so) is the
// checker and the CPU baseline; nothing on the product path loads it.
//
// Reference entry points exercised:
//   Jet::Tensor<T>::Transpose       include/jet/Tensor.hpp:579-612
//   Jet::Tensor<T>::ContractTensors include/jet/Tensor.hpp:709-752
//   Jet::Tensor<T>::SliceIndex      include/jet/Tensor.hpp:494-526
//   Jet::Tensor<T>::AddTensors      include/jet/Tensor.hpp:413-454
//   TensorNetwork::SliceIndices     include/jet/TensorNetwork.hpp:210-284
//   TensorNetwork::Contract(path)   include/jet/TensorNetwork.hpp:301-328
//   PathInfo                        include/jet/PathInfo.hpp:81-113
//   TaskBasedContractor             include/jet/TaskBasedContractor.hpp:162-322
//                                   (through the Taskflow stand-in in ref_shim/)
//   TensorNetworkSerializer         include/jet/TensorNetworkIO.hpp:155-187
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "jet/PathInfo.hpp"
#include "jet/TaskBasedContractor.hpp"
#include "jet/Tensor.hpp"
#include "jet/TensorNetwork.hpp"
#include "jet/TensorNetworkIO.hpp"

namespace {

thread_local std::string g_error;

template <class F> int Guard(F &&f)
{
    try {
        f();
        return 0;
    }
    catch (const std::exception &e) {
        g_error = e.what();
        return 1;
    }
    catch (...) {
        g_error = "unknown exception";
        return 2;
    }
}

std::vector<std::string> Labels(int rank, const int32_t *ids)
{
    std::vector<std::string> out(rank);
    for (int i = 0; i < rank; i++)
        out[i] = "i" + std::to_string(ids[i]);
    return out;
}

std::vector<std::string> SplitWs(const char *s)
{
    std::vector<std::string> out;
    std::istringstream is(s ? s : "");
    std::string tok;
    while (is >> tok)
        out.push_back(tok);
    return out;
}

template <class T>
Jet::Tensor<T> MakeTensor(int rank, const int64_t *shape, const int32_t *ids, const void *data)
{
    std::vector<size_t> shp(shape, shape + rank);
    size_t n = 1;
    for (auto s : shp)
        n *= s;
    const T *p = static_cast<const T *>(data);
    std::vector<T> v(p, p + n);
    return Jet::Tensor<T>(Labels(rank, ids), shp, v);
}

template <class T>
void Transpose(int rank, const int64_t *shape, const int32_t *perm, const void *in, void *out)
{
    std::vector<int32_t> ids(rank);
    for (int i = 0; i < rank; i++)
        ids[i] = i;
    auto t = MakeTensor<T>(rank, shape, ids.data(), in);
    std::vector<std::string> new_idx(rank);
    for (int j = 0; j < rank; j++)
        new_idx[j] = "i" + std::to_string(perm[j]);
    auto r = Jet::Tensor<T>::template Transpose<>(t, new_idx);
    std::memcpy(out, r.GetData().data(), r.GetSize() * sizeof(T));
}

template <class T>
void Contract(int ra, const int64_t *sa, const int32_t *ia, const void *a, int rb,
              const int64_t *sb, const int32_t *ib, const void *b, void *c, int64_t *c_size)
{
    auto A = MakeTensor<T>(ra, sa, ia, a);
    auto B = MakeTensor<T>(rb, sb, ib, b);
    auto C = Jet::Tensor<T>::template ContractTensors<>(A, B);
    *c_size = static_cast<int64_t>(C.GetSize());
    std::memcpy(c, C.GetData().data(), C.GetSize() * sizeof(T));
}

template <class T>
void Network(const char *json, const char *sliced, uint64_t value, int mode, int threads,
             uint64_t num_slices, double *out, int64_t out_cap, int64_t *out_size,
             double *seconds, double *flops)
{
    using Tensor = Jet::Tensor<T>;
    using namespace std::chrono;
    Jet::TensorNetworkSerializer<Tensor> ser;
    auto file = ser(std::string(json));
    auto path = file.path.value().GetPath();
    auto idx = SplitWs(sliced);
    std::vector<T> result;
    if (mode == 0) {
        // serial: TensorNetwork::Contract(path) on one slice
        auto tn = file.tensors;
        if (!idx.empty())
            tn.SliceIndices(idx, value);
        Jet::PathInfo pi(tn, path);
        *flops = pi.GetTotalFlops();
        auto t1 = high_resolution_clock::now();
        const auto &r = tn.Contract(path);
        auto t2 = high_resolution_clock::now();
        *seconds = duration_cast<duration<double>>(t2 - t1).count();
        result = r.GetData();
    }
    else {
        // TaskBasedContractor over slices value .. value+num_slices-1 with a reduce,
        // mirroring examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp:69-93
        std::vector<Jet::TensorNetwork<Tensor>> slices(num_slices);
        for (uint64_t i = 0; i < num_slices; i++) {
            slices[i] = file.tensors;
            if (!idx.empty())
                slices[i].SliceIndices(idx, value + i);
        }
        Jet::TaskBasedContractor<Tensor> tbc(static_cast<size_t>(threads));
        for (uint64_t i = 0; i < num_slices; i++) {
            Jet::PathInfo pi(slices[i], path);
            tbc.AddContractionTasks(slices[i], pi);
        }
        tbc.AddReductionTask();
        if (mode == 2)
            tbc.AddDeletionTasks();
        *flops = tbc.GetFlops();
        auto t1 = high_resolution_clock::now();
        tbc.Contract().wait();
        auto t2 = high_resolution_clock::now();
        *seconds = duration_cast<duration<double>>(t2 - t1).count();
        result = tbc.GetReductionResult().GetData();
    }
    *out_size = static_cast<int64_t>(result.size());
    for (int64_t i = 0; i < std::min<int64_t>(out_cap, result.size()); i++) {
        out[2 * i] = static_cast<double>(result[i].real());
        out[2 * i + 1] = static_cast<double>(result[i].imag());
    }
}

} // namespace

#pragma GCC visibility push(default)
extern "C" {

const char *ref_last_error() { return g_error.c_str(); }

void ref_set_blas_threads(int n) { scipy_openblas_set_num_threads(n); }
int ref_get_blas_threads() { return scipy_openblas_get_num_threads(); }
const char *ref_blas_config() { return scipy_openblas_get_config(); }

// dtype: 0 = complex<float>, 1 = complex<double>.
// out axis j = in axis perm[j]  (new_indices[j] = indices[perm[j]]).
int ref_transpose(int dtype, int rank, const int64_t *shape, const int32_t *perm, const void *in,
                  void *out)
{
    return Guard([&] {
        if (dtype == 0)
            Transpose<std::complex<float>>(rank, shape, perm, in, out);
        else
            Transpose<std::complex<double>>(rank, shape, perm, in, out);
    });
}

// Index labels are integer ids; equal ids are contracted.
int ref_contract(int dtype, int ra, const int64_t *sa, const int32_t *ia, const void *a, int rb,
                 const int64_t *sb, const int32_t *ib, const void *b, void *c, int64_t *c_size)
{
    return Guard([&] {
        if (dtype == 0)
            Contract<std::complex<float>>(ra, sa, ia, a, rb, sb, ib, b, c, c_size);
        else
            Contract<std::complex<double>>(ra, sa, ia, a, rb, sb, ib, b, c, c_size);
    });
}

// mode 0: TensorNetwork::Contract(path) on slice `value` (no slicing when `sliced` is empty)
// mode 1: TaskBasedContractor(threads) over `num_slices` slices + reduction
// mode 2: mode 1 + AddDeletionTasks()
// out receives (re, im) pairs as doubles.
int ref_network(int dtype, const char *json, const char *sliced, uint64_t value, int mode,
                int threads, uint64_t num_slices, double *out, int64_t out_cap, int64_t *out_size,
                double *seconds, double *flops)
{
    return Guard([&] {
        if (dtype == 0)
            Network<std::complex<float>>(json, sliced, value, mode, threads, num_slices, out,
                                         out_cap, out_size, seconds, flops);
        else
            Network<std::complex<double>>(json, sliced, value, mode, threads, num_slices, out,
                                          out_cap, out_size, seconds, flops);
    });
}

// SliceIndex / AddTensors for fixture generation.
int ref_slice_index(int dtype, int rank, const int64_t *shape, int axis, int64_t value,
                    const void *in, void *out)
{
    return Guard([&] {
        std::vector<int32_t> ids(rank);
        for (int i = 0; i < rank; i++)
            ids[i] = i;
        if (dtype == 0) {
            auto t = MakeTensor<std::complex<float>>(rank, shape, ids.data(), in);
            auto r = Jet::Tensor<std::complex<float>>::SliceIndex(t, "i" + std::to_string(axis),
                                                                  static_cast<size_t>(value));
            std::memcpy(out, r.GetData().data(), r.GetSize() * sizeof(std::complex<float>));
        }
        else {
            auto t = MakeTensor<std::complex<double>>(rank, shape, ids.data(), in);
            auto r = Jet::Tensor<std::complex<double>>::SliceIndex(t, "i" + std::to_string(axis),
                                                                   static_cast<size_t>(value));
            std::memcpy(out, r.GetData().data(), r.GetSize() * sizeof(std::complex<double>));
        }
    });
}

} // extern "C"
#pragma GCC visibility pop
