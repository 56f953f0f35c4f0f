// Jet::Permuter<Backend> — drop-in for /root/reference/include/jet/permute/Permuter.hpp:50-79 (and the back ends
// permute/Default.hpp:21-133, permute/QFlex.hpp:37-50): the same two Transpose overloads with the same argument
// checks and messages; the data movement runs on the GPU (jb_permute_host -> PermuteBitsKernel for power-of-two
// extents, PermuteGenericKernel otherwise).  The back-end template parameters (block sizes) are accepted for source
// compatibility and ignored: there is one B200 kernel family behind every back end.
#pragma once

#include <algorithm>
#include <complex>
#include <cstdint>
#include <set>
#include <string>
#include <type_traits>
#include <vector>

#include "../Abort.hpp"
#include "../Utilities.hpp"
#include "jetb200.h"

namespace Jet {

namespace PermuterDetail {
template <class DataType>
inline void DeviceTranspose(const std::vector<DataType> &data_in, const std::vector<size_t> &shape,
                            std::vector<DataType> &data_out, const std::vector<std::string> &current_order,
                            const std::vector<std::string> &new_order)
{
    static_assert(std::is_same_v<DataType, std::complex<float>> || std::is_same_v<DataType, std::complex<double>>,
                  "the B200 permuter moves complex<float> / complex<double> elements");
    std::vector<int64_t> extent(shape.begin(), shape.end());
    std::vector<int32_t> perm(new_order.size());
    for (size_t j = 0; j < new_order.size(); j++)
        perm[j] = static_cast<int32_t>(std::find(current_order.begin(), current_order.end(), new_order[j]) -
                                       current_order.begin());
    constexpr int dtype = std::is_same_v<DataType, std::complex<float>> ? JB_C64 : JB_C128;
    JET_JB_CHECK(jb_permute_host(dtype, data_in.data(), data_out.data(), static_cast<int>(shape.size()), extent.data(),
                                 perm.data()));
}
} // namespace PermuterDetail

template <size_t blocksize = 1024> class DefaultPermuter {
  public:
    template <class DataType>
    void Transpose(const std::vector<DataType> &data_in, const std::vector<size_t> &shape, std::vector<DataType> &data_out,
                   const std::vector<std::string> &current_order, const std::vector<std::string> &new_order)
    {
        PermuterDetail::DeviceTranspose(data_in, shape, data_out, current_order, new_order);
    }
    template <class DataType>
    std::vector<DataType> Transpose(const std::vector<DataType> &data_in, const std::vector<size_t> &shape,
                                    const std::vector<std::string> &current_order, const std::vector<std::string> &new_order)
    {
        std::vector<DataType> data_out(data_in.size());
        PermuterDetail::DeviceTranspose(data_in, shape, data_out, current_order, new_order);
        return data_out;
    }
};

template <size_t blocksize = 1024, size_t min_dims = 32> class QFlexPermuter : public DefaultPermuter<blocksize> {
};

template <class PermuterBackend> class Permuter {
  public:
    template <class DataType>
    void Transpose(const std::vector<DataType> &data_in, const std::vector<size_t> &shape, std::vector<DataType> &data_out,
                   const std::vector<std::string> &current_order, const std::vector<std::string> &new_order)
    {
        Check_(shape, data_in.size(), current_order, new_order);
        JET_ABORT_IF_NOT(Jet::Utilities::ShapeToSize(shape) == data_out.size(),
                         "Tensor shape does not match given output tensor data.");
        permuter_b_.Transpose(data_in, shape, data_out, current_order, new_order);
    }

    template <class DataType>
    std::vector<DataType> Transpose(const std::vector<DataType> &data_in, const std::vector<size_t> &shape,
                                    const std::vector<std::string> &current_order, const std::vector<std::string> &new_order)
    {
        Check_(shape, data_in.size(), current_order, new_order);
        return permuter_b_.Transpose(data_in, shape, current_order, new_order);
    }

  private:
    PermuterBackend permuter_b_;

    static void Check_(const std::vector<size_t> &shape, size_t size_in, const std::vector<std::string> &current_order,
                       const std::vector<std::string> &new_order)
    {
        const std::set<std::string> idx_old(current_order.begin(), current_order.end());
        const std::set<std::string> idx_new(new_order.begin(), new_order.end());
        JET_ABORT_IF_NOT(idx_old.size() == current_order.size(),
                         "Duplicate existing indices found. Please ensure indices are unique.");
        JET_ABORT_IF_NOT(idx_new.size() == new_order.size(),
                         "Duplicate transpose indices found. Please ensure indices are unique.");
        JET_ABORT_IF_NOT(shape.size() == new_order.size(), "Tensor shape does not match number of indices.");
        JET_ABORT_IF_NOT(Jet::Utilities::ShapeToSize(shape) == size_in,
                         "Tensor shape does not match given input tensor data.");
        JET_ABORT_IF_NOT(idx_old == idx_new, "New indices are an invalid permutation of the existing indices");
    }
};

} // namespace Jet
