// Jet::TensorHelpers::MultiplyTensorData — drop-in for /root/reference/include/jet/TensorHelpers.hpp:131-168: the
// row-major complex product C(M x N) = A(M x K) B(K x N) behind Tensor::ContractTensors, with the reference's
// four-way dispatch (GEMM / GEMV / GEMV-transposed / DOTU: an empty index list means extent 1 on that side)
// folded into one call: jb_gemm_host treats M == 1 / N == 1 as the GEMV / DOTU corners of the same product
// (GemmTf32x3Kernel / GemmDmmaKernel / SmallMnKernel / GemmKernel by shape and dtype).
#pragma once

#include <complex>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

#include "Abort.hpp"
#include "Tensor.hpp"
#include "jetb200.h"

namespace Jet {
namespace TensorHelpers {

template <typename ComplexPrecision, std::enable_if_t<is_supported_data_type<ComplexPrecision>, bool> = true>
inline void MultiplyTensorData(const std::vector<ComplexPrecision> &A, const std::vector<ComplexPrecision> &B,
                               std::vector<ComplexPrecision> &C, const std::vector<std::string> &left_indices,
                               const std::vector<std::string> &right_indices, size_t left_dim, size_t right_dim,
                               size_t common_dim)
{
    const int64_t m = left_indices.empty() ? 1 : static_cast<int64_t>(left_dim);
    const int64_t n = right_indices.empty() ? 1 : static_cast<int64_t>(right_dim);
    JET_JB_CHECK(jb_gemm_host(DtypeCode<ComplexPrecision>(), m, n, static_cast<int64_t>(common_dim), A.data(), B.data(),
                              C.data()));
}

} // namespace TensorHelpers
} // namespace Jet
