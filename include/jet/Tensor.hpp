// Jet::Tensor<T> — drop-in for /root/reference/include/jet/Tensor.hpp on top of libjetb200.so.
//
// Same public surface (constructors, accessors, static + member forms of AddTensors, SliceIndex,
// Reshape, Transpose, Conj, ContractTensors) and the same value semantics: a host-side
// std::vector<T> in row-major order with string index labels.  What changed is where the work
// happens: every operator that moves or multiplies data calls the C ABI (include/jetb200.h) and
// runs on the GPU —
//   Transpose        -> jb_permute_host   (replaces Permuter<QFlex/Default>, Tensor.hpp:594-611)
//   ContractTensors  -> jb_contract_host  (replaces 2x Transpose + cblas gemm/gemv/dotu,
//                                          Tensor.hpp:709-752, TensorHelpers.hpp:131-168)
//   AddTensors       -> jb_permute_host + jb_add_host   (Tensor.hpp:413-454)
//   SliceIndex       -> jb_slice_host     (Tensor.hpp:494-526)
//   Conj             -> jb_conj_host
// There is no CPU implementation behind these: if the library or a GPU is missing they throw
// Jet::Exception.  Whole-network contractions should go through TaskBasedContractor /
// SlicedContractor, which keep every intermediate on the device.
#pragma once

#include <algorithm>
#include <complex>
#include <cstdint>
#include <iostream>
#include <random>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "Abort.hpp"
#include "Utilities.hpp"
#include "jetb200.h"

namespace Jet {

namespace TensorHelpers {
template <class T>
constexpr bool is_supported_data_type =
    std::is_same_v<T, std::complex<float>> || std::is_same_v<T, std::complex<double>>;

template <class T> constexpr int DtypeCode()
{
    return std::is_same_v<T, std::complex<float>> ? JB_C64 : JB_C128;
}
} // namespace TensorHelpers

template <class T = std::complex<float>> class Tensor {
    static_assert(TensorHelpers::is_supported_data_type<T>,
                  "Tensor supports only complex<float> and complex<double>.");

  public:
    using scalar_type_t = T;

    /// Default tensor: one zero element, no indices (a scalar; the additive identity of
    /// AddTensors).
    Tensor() : data_(1) {}

    /// Zero tensor of the given shape with default labels "?a", "?b", ...
    Tensor(const std::vector<size_t> &shape) : data_(Utilities::ShapeToSize(shape))
    {
        std::vector<std::string> labels(shape.size());
        for (size_t i = 0; i < labels.size(); i++)
            labels[i] = "?" + Utilities::GenerateStringIndex(i);
        InitIndicesAndShape(labels, shape);
    }

    Tensor(const std::vector<std::string> &indices, const std::vector<size_t> &shape)
        : data_(Utilities::ShapeToSize(shape))
    {
        InitIndicesAndShape(indices, shape);
    }

    Tensor(const std::vector<std::string> &indices, const std::vector<size_t> &shape,
           const std::vector<T> &data)
        : Tensor(indices, shape)
    {
        const size_t n = std::min(data.size(), data_.size());
        std::copy(data.begin(), data.begin() + n, data_.begin());
    }

    Tensor(const Tensor &other)
        : indices_(other.indices_), shape_(other.shape_),
          index_to_dimension_(other.index_to_dimension_), data_(other.data_)
    {
    }

    Tensor(Tensor &&other)
        : indices_(std::move(other.indices_)), shape_(std::move(other.shape_)),
          index_to_dimension_(std::move(other.index_to_dimension_)), data_(std::move(other.data_))
    {
    }

    virtual ~Tensor() {}

    void InitIndicesAndShape(const std::vector<std::string> &indices,
                             const std::vector<size_t> &shape) noexcept
    {
        indices_ = indices;
        shape_ = shape;
        index_to_dimension_.clear();
        for (size_t i = 0; i < shape_.size(); i++)
            index_to_dimension_[indices_[i]] = shape_[i];
    }

    void SetShape(const std::vector<size_t> &shape) noexcept { shape_ = shape; }
    const std::vector<size_t> &GetShape() const noexcept { return shape_; }

    T &operator[](size_t pos) { return data_[pos]; }
    const T &operator[](size_t pos) const { return data_[pos]; }

    void RenameIndex(size_t pos, std::string new_label) noexcept
    {
        const std::string old_label = indices_[pos];
        const size_t dim = index_to_dimension_[old_label];
        index_to_dimension_.erase(old_label);
        indices_[pos] = new_label;
        index_to_dimension_.emplace(new_label, dim);
    }

    bool operator==(const Tensor<T> &other) const noexcept
    {
        return shape_ == other.shape_ && indices_ == other.indices_ &&
               index_to_dimension_ == other.index_to_dimension_ && data_ == other.data_;
    }
    bool operator!=(const Tensor<T> &other) const { return !(*this == other); }

    const Tensor<T> &operator=(const Tensor<T> &other)
    {
        if (this != &other) {
            indices_ = other.indices_;
            shape_ = other.shape_;
            index_to_dimension_ = other.index_to_dimension_;
            data_ = other.data_;
        }
        return *this;
    }

    const Tensor<T> &operator=(Tensor<T> &&other)
    {
        if (this != &other) {
            indices_ = std::move(other.indices_);
            shape_ = std::move(other.shape_);
            index_to_dimension_ = std::move(other.index_to_dimension_);
            data_ = std::move(other.data_);
        }
        return *this;
    }

    const std::unordered_map<std::string, size_t> &GetIndexToDimension() const
    {
        return index_to_dimension_;
    }

    void SetValue(const std::vector<size_t> &indices, const T &value)
    {
        data_[Utilities::RavelIndex(indices, shape_)] = value;
    }
    T GetValue(const std::vector<size_t> &indices) const
    {
        return data_[Utilities::RavelIndex(indices, shape_)];
    }

    void SetData(const std::vector<T> &data)
    {
        JET_ABORT_IF_NOT(data.size() == GetSize(), "Size of data and tensor do not match.");
        data_ = data;
    }
    const std::vector<T> &GetData() const noexcept { return data_; }
    std::vector<T> &GetData() { return data_; }

    const std::vector<std::string> &GetIndices() const noexcept { return indices_; }
    size_t GetSize() const { return data_.size(); }
    const T &GetScalar() const { return data_[0]; }
    bool IsScalar() const noexcept { return GetSize() == 1; }

    /// Deterministic fill: real then imaginary part per element from mt19937(seed),
    /// uniform in [-1, 1)  (host-side, like the reference; inputs are not on the hot path).
    void FillRandom(size_t seed)
    {
        std::mt19937 engine(seed);
        std::uniform_real_distribution<typename T::value_type> dist(-1, 1);
        for (auto &z : data_) {
            const auto re = dist(engine);
            const auto im = dist(engine);
            z = T{re, im};
        }
    }

    void FillRandom()
    {
        static std::mt19937 engine(std::random_device{}());
        static std::uniform_real_distribution<typename T::value_type> dist(-1, 1);
        for (auto &z : data_) {
            const auto re = dist(engine);
            const auto im = dist(engine);
            z = T{re, im};
        }
    }

    // ---- AddTensors -----------------------------------------------------------------------------
    template <class U = T> static Tensor<U> AddTensors(const Tensor<U> &A, const Tensor<U> &B)
    {
        static const Tensor<U> zero;
        if (A == zero)
            return B;
        if (B == zero)
            return A;
        JET_ABORT_IF_NOT(Utilities::VectorDisjunctiveUnion(A.GetIndices(), B.GetIndices()).empty(),
                         "Tensor addition with disjoint indices is not supported.");
        JET_ABORT_IF_NOT(A.GetSize() == B.GetSize(), "Size is inconsistent between tensors.");
        const Tensor<U> Bt = A.GetIndices() == B.GetIndices() ? B : Transpose<U>(B, A.GetIndices());
        Tensor<U> C(A.GetIndices(), A.GetShape());
        JET_JB_CHECK(jb_add_host(TensorHelpers::DtypeCode<U>(), static_cast<int64_t>(C.GetSize()),
                                 A.GetData().data(), Bt.GetData().data(), C.GetData().data()));
        return C;
    }
    Tensor<T> AddTensor(const Tensor<T> &other) const { return AddTensors<T>(*this, other); }

    // ---- SliceIndex -----------------------------------------------------------------------------
    template <class U = T>
    static Tensor<U> SliceIndex(const Tensor<U> &tensor, const std::string &index, size_t value)
    {
        const auto &idx = tensor.GetIndices();
        const auto it = std::find(idx.begin(), idx.end(), index);
        JET_ABORT_IF(it == idx.end(), "Sliced index does not exist.");
        const size_t axis = static_cast<size_t>(it - idx.begin());
        JET_ABORT_IF_NOT(value < tensor.GetShape()[axis], "Sliced value is out of range.");
        std::vector<std::string> new_indices = idx;
        std::vector<size_t> new_shape = tensor.GetShape();
        new_indices.erase(new_indices.begin() + axis);
        new_shape.erase(new_shape.begin() + axis);
        Tensor<U> out(new_indices, new_shape);
        std::vector<int64_t> extent(tensor.GetShape().begin(), tensor.GetShape().end());
        JET_JB_CHECK(jb_slice_host(TensorHelpers::DtypeCode<U>(), tensor.GetData().data(),
                                   out.GetData().data(), static_cast<int>(extent.size()),
                                   extent.data(), static_cast<int>(axis),
                                   static_cast<int64_t>(value)));
        return out;
    }
    Tensor<T> SliceIndex(const std::string &index, size_t value) const
    {
        return SliceIndex<T>(*this, index, value);
    }

    // ---- Reshape --------------------------------------------------------------------------------
    template <class U = T>
    static Tensor<U> Reshape(const Tensor<U> &old_tensor, const std::vector<size_t> &new_shape)
    {
        JET_ABORT_IF_NOT(old_tensor.GetSize() == Utilities::ShapeToSize(new_shape),
                         "Size is inconsistent between tensors.");
        Tensor<U> out(new_shape);
        out.GetData() = old_tensor.GetData();
        return out;
    }
    Tensor<T> Reshape(const std::vector<size_t> &new_shape) const
    {
        return Reshape<T>(*this, new_shape);
    }

    // ---- Transpose ------------------------------------------------------------------------------
    // BLOCKSIZE / MINSIZE are accepted for source compatibility (they tuned the reference's
    // cache-blocked QFlex passes); the GPU kernel plans its own tiles.
    template <class U = T, size_t BLOCKSIZE = 1024, size_t MINSIZE = 32>
    static Tensor<U> Transpose(const Tensor<U> &A, const std::vector<std::string> &new_indices)
    {
        const auto &old_indices = A.GetIndices();
        if (new_indices == old_indices)
            return A;
        JET_ABORT_IF(old_indices.empty(), "Number of indices cannot be zero.");
        JET_ABORT_IF_NOT(new_indices.size() == old_indices.size(),
                         "Tensor shape does not match number of new indices.");
        const size_t rank = old_indices.size();
        std::vector<int32_t> perm(rank);
        std::vector<size_t> new_shape(rank);
        std::vector<bool> used(rank, false);
        for (size_t j = 0; j < rank; j++) {
            const auto it = std::find(old_indices.begin(), old_indices.end(), new_indices[j]);
            JET_ABORT_IF(it == old_indices.end(),
                         "New indices are an invalid permutation of the existing indices.");
            const size_t p = static_cast<size_t>(it - old_indices.begin());
            JET_ABORT_IF(used[p], "Duplicate new indices found.");
            used[p] = true;
            perm[j] = static_cast<int32_t>(p);
            new_shape[j] = A.GetShape()[p];
        }
        Tensor<U> out(new_indices, new_shape);
        std::vector<int64_t> extent(A.GetShape().begin(), A.GetShape().end());
        JET_JB_CHECK(jb_permute_host(TensorHelpers::DtypeCode<U>(), A.GetData().data(),
                                     out.GetData().data(), static_cast<int>(rank), extent.data(),
                                     perm.data()));
        return out;
    }

    template <class U = T, size_t BLOCKSIZE = 1024, size_t MINSIZE = 32>
    static Tensor<U> Transpose(const Tensor<U> &A, const std::vector<size_t> &new_ordering)
    {
        const auto &old_indices = A.GetIndices();
        JET_ABORT_IF_NOT(old_indices.size() == new_ordering.size(),
                         "Size of ordering must match number of tensor indices.");
        std::vector<std::string> new_indices(new_ordering.size());
        for (size_t i = 0; i < new_ordering.size(); i++) {
            JET_ABORT_IF_NOT(new_ordering[i] < old_indices.size(), "Ordering entry is out of range.");
            new_indices[i] = old_indices[new_ordering[i]];
        }
        return Transpose<U, BLOCKSIZE, MINSIZE>(A, new_indices);
    }
    Tensor<T> Transpose(const std::vector<size_t> &new_ordering) const
    {
        return Transpose<T>(*this, new_ordering);
    }
    Tensor<T> Transpose(const std::vector<std::string> &new_indices) const
    {
        return Transpose<T>(*this, new_indices);
    }

    // ---- Conj -----------------------------------------------------------------------------------
    template <class U = T> static Tensor<U> Conj(const Tensor<U> &A)
    {
        Tensor<U> out(A.GetIndices(), A.GetShape());
        JET_JB_CHECK(jb_conj_host(TensorHelpers::DtypeCode<U>(), static_cast<int64_t>(A.GetSize()),
                                  A.GetData().data(), out.GetData().data()));
        return out;
    }
    Tensor<T> Conj() const { return Conj<T>(*this); }

    // ---- ContractTensors ------------------------------------------------------------------------
    // C[left ++ right] = sum over common indices of A * B (no conjugation); left = A\B in A's
    // order, right = B\A in B's order.  One fused GPU call: the transposes of the reference are
    // folded into the kernel's load addresses.
    template <class U = T> static Tensor<U> ContractTensors(const Tensor<U> &A, const Tensor<U> &B)
    {
        std::unordered_map<std::string, int32_t> label;
        auto modes_of = [&label](const std::vector<std::string> &idx) {
            std::vector<int32_t> m(idx.size());
            for (size_t i = 0; i < idx.size(); i++)
                m[i] = label.emplace(idx[i], static_cast<int32_t>(label.size())).first->second;
            return m;
        };
        const auto modes_a = modes_of(A.GetIndices());
        const auto modes_b = modes_of(B.GetIndices());
        std::vector<int64_t> ext_a(A.GetShape().begin(), A.GetShape().end());
        std::vector<int64_t> ext_b(B.GetShape().begin(), B.GetShape().end());
        const auto left = Utilities::VectorSubtraction(A.GetIndices(), B.GetIndices());
        const auto right = Utilities::VectorSubtraction(B.GetIndices(), A.GetIndices());
        std::vector<std::string> c_indices = Utilities::VectorConcatenation(left, right);
        std::vector<size_t> c_shape;
        for (const auto &i : left)
            c_shape.push_back(A.GetIndexToDimension().at(i));
        for (const auto &i : right)
            c_shape.push_back(B.GetIndexToDimension().at(i));
        Tensor<U> C(c_indices, c_shape);
        JET_JB_CHECK(jb_contract_host(TensorHelpers::DtypeCode<U>(), static_cast<int>(ext_a.size()),
                                      ext_a.data(), modes_a.data(), A.GetData().data(),
                                      static_cast<int>(ext_b.size()), ext_b.data(), modes_b.data(),
                                      B.GetData().data(), C.GetData().data()));
        return C;
    }
    Tensor<T> ContractWithTensor(const Tensor<T> &other) const
    {
        return ContractTensors<T>(*this, other);
    }

  private:
    std::vector<std::string> indices_;
    std::vector<size_t> shape_;
    std::unordered_map<std::string, size_t> index_to_dimension_;
    std::vector<T> data_;
};

template <class T> inline std::ostream &operator<<(std::ostream &out, const Tensor<T> &tensor)
{
    using namespace Jet::Utilities;
    out << "Size = " << tensor.GetSize() << std::endl;
    out << "Indices = " << tensor.GetIndices() << std::endl;
    out << "Data = " << tensor.GetData();
    return out;
}

} // namespace Jet
