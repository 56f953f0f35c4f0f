// Host-side helpers the drop-in headers (and users of the Jet API) rely on: ordered set algebra on
// vectors, row-major ravel/unravel, default index labels.  Same names and results as
// /root/reference/include/jet/Utilities.hpp (:86-97 labels, :317-384 set algebra, :438-501
// shape helpers); implementations are independent.
#pragma once

#include <algorithm>
#include <cstddef>
#include <ostream>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#include "Abort.hpp"

namespace Jet {
namespace Utilities {

inline constexpr bool is_pow_2(size_t value) { return value != 0 && (value & (value - 1)) == 0; }

inline constexpr size_t fast_log2(size_t value)
{
    size_t l = 0;
    while (value >>= 1)
        ++l;
    return l;
}

/// Label for the id-th default index: a..z, A..Z, then a0..Z0, a1.. (52 letters per block).
inline std::string GenerateStringIndex(size_t id)
{
    static const std::string alphabet = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ";
    const size_t n = alphabet.size();
    std::string label(1, alphabet[id % n]);
    if (id >= n)
        label += std::to_string(id / n - 1);
    return label;
}

template <class T> inline std::ostream &operator<<(std::ostream &os, const std::vector<T> &v)
{
    os << '{';
    for (size_t i = 0; i < v.size(); i++) {
        if (i)
            os << "  ";
        os << v[i];
    }
    os << '}';
    return os;
}

template <class T> inline bool InVector(const T &e, const std::vector<T> &v)
{
    return std::find(v.begin(), v.end(), e) != v.end();
}

/// Elements of a followed by the elements of b that are not in a.
template <class T> inline std::vector<T> VectorUnion(const std::vector<T> &a, const std::vector<T> &b)
{
    std::vector<T> out = a;
    for (const auto &e : b)
        if (!InVector(e, a))
            out.push_back(e);
    return out;
}

/// Elements of a that are also in b, in a's order.
template <class T>
inline std::vector<T> VectorIntersection(const std::vector<T> &a, const std::vector<T> &b)
{
    std::vector<T> out;
    for (const auto &e : a)
        if (InVector(e, b))
            out.push_back(e);
    return out;
}

/// Elements of a that are not in b, in a's order.
template <class T>
inline std::vector<T> VectorSubtraction(const std::vector<T> &a, const std::vector<T> &b)
{
    std::vector<T> out;
    for (const auto &e : a)
        if (!InVector(e, b))
            out.push_back(e);
    return out;
}

/// (a \ b) followed by (b \ a).
template <class T>
inline std::vector<T> VectorDisjunctiveUnion(const std::vector<T> &a, const std::vector<T> &b)
{
    std::vector<T> out = VectorSubtraction(a, b);
    for (const auto &e : b)
        if (!InVector(e, a))
            out.push_back(e);
    return out;
}

template <class T>
inline std::vector<T> VectorConcatenation(const std::vector<T> &a, const std::vector<T> &b)
{
    std::vector<T> out = a;
    out.insert(out.end(), b.begin(), b.end());
    return out;
}

inline std::string JoinStringVector(const std::vector<std::string> &v)
{
    std::string out;
    for (const auto &s : v)
        out += s;
    return out;
}

inline size_t ShapeToSize(const std::vector<size_t> &shape)
{
    size_t n = 1;
    for (size_t s : shape)
        n *= s;
    return n;
}

/// Row-major position -> multi-index over `shape` (first axis slowest).
inline std::vector<size_t> UnravelIndex(unsigned long long index, const std::vector<size_t> &shape)
{
    const size_t size = ShapeToSize(shape);
    JET_ABORT_IF(size <= index && !(shape.empty() && index == 0),
                 "Linear index does not fit in the shape.");
    std::vector<size_t> out(shape.size());
    for (size_t j = shape.size(); j-- > 0;) {
        out[j] = index % shape[j];
        index /= shape[j];
    }
    return out;
}

/// Multi-index -> row-major position.
inline unsigned long long RavelIndex(const std::vector<size_t> &index, const std::vector<size_t> &shape)
{
    JET_ABORT_IF_NOT(index.size() == shape.size(),
                     "Number of index and shape dimensions must match.");
    unsigned long long pos = 0;
    for (size_t j = 0; j < shape.size(); j++) {
        JET_ABORT_IF(index[j] >= shape[j], "Multi-dimensional index does not fit in the shape.");
        pos = pos * shape[j] + index[j];
    }
    return pos;
}

inline void SplitStringOnMultipleDelimiters(std::string s, const std::vector<std::string> &delims,
                                            std::vector<std::string> &out)
{
    for (const auto &d : delims) {
        size_t pos;
        while (!d.empty() && (pos = s.find(d)) != std::string::npos)
            s.replace(pos, d.size(), " ");
    }
    std::istringstream is(s);
    std::string tok;
    while (is >> tok)
        out.push_back(tok);
}

} // namespace Utilities
} // namespace Jet
