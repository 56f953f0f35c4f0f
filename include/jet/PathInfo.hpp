// Jet::PathInfo / PathStepInfo — drop-in for /root/reference/include/jet/PathInfo.hpp: a symbolic
// replay of a contraction path (names, node/tensor indices, contracted indices, parent/children)
// with the reference's cost conventions: flops of a step = (elements of its tensor) * 2 *
// (product of contracted extents) — i.e. 2*M*N*K — and memory = elements of its tensor.
#pragma once

#include <limits>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "Abort.hpp"
#include "TensorNetwork.hpp"
#include "Utilities.hpp"

namespace Jet {

struct PathStepInfo {
    size_t id;
    size_t parent;
    std::pair<size_t, size_t> children;
    std::string name;
    std::vector<std::string> node_indices;
    std::vector<std::string> tensor_indices;
    std::vector<std::string> tags;
    std::vector<std::string> contracted_indices;
    static constexpr size_t MISSING_ID = std::numeric_limits<size_t>::max();
};

class PathInfo {
  public:
    using NodeID_t = size_t;
    using Path = std::vector<std::pair<NodeID_t, NodeID_t>>;
    using IndexToSizeMap = std::unordered_map<std::string, size_t>;
    using Steps = std::vector<PathStepInfo>;

    PathInfo() : num_leaves_(0) {}

    template <typename Tensor>
    PathInfo(const TensorNetwork<Tensor> &tn, const Path &path) : path_(path)
    {
        constexpr size_t none = PathStepInfo::MISSING_ID;
        num_leaves_ = tn.GetNodes().size();
        for (const auto &node : tn.GetNodes())
            steps_.push_back(PathStepInfo{node.id, none, {none, none}, node.name, node.indices,
                                          node.tensor.GetIndices(), node.tags, {}});
        for (const auto &[index, edge] : tn.GetIndexToEdgeMap())
            index_to_size_map_.emplace(index, edge.dim);
        for (const auto &[a, b] : path) {
            JET_ABORT_IF_NOT(a < steps_.size(), "Node ID 1 in contraction path pair is invalid.");
            JET_ABORT_IF_NOT(b < steps_.size(), "Node ID 2 in contraction path pair is invalid.");
            Replay_(a, b);
        }
    }

    const IndexToSizeMap &GetIndexSizes() const noexcept { return index_to_size_map_; }
    size_t GetNumLeaves() const noexcept { return num_leaves_; }
    const Path &GetPath() const noexcept { return path_; }
    const Steps &GetSteps() const noexcept { return steps_; }

    double GetPathStepFlops(size_t id) const
    {
        JET_ABORT_IF_NOT(id < steps_.size(), "Step ID is invalid.");
        if (id < num_leaves_)
            return 0;
        const double k = Product_(steps_[id].contracted_indices);
        return Product_(steps_[id].tensor_indices) * (k + k);
    }

    double GetTotalFlops() const noexcept
    {
        double total = 0;
        for (size_t id = num_leaves_; id < steps_.size(); id++)
            total += GetPathStepFlops(id);
        return total;
    }

    double GetPathStepMemory(size_t id) const
    {
        JET_ABORT_IF_NOT(id < steps_.size(), "Step ID is invalid.");
        return Product_(steps_[id].tensor_indices);
    }

    double GetTotalMemory() const noexcept
    {
        double total = 0;
        for (size_t id = 0; id < steps_.size(); id++)
            total += GetPathStepMemory(id);
        return total;
    }

  private:
    Path path_;
    Steps steps_;
    size_t num_leaves_;
    IndexToSizeMap index_to_size_map_;

    // unknown (e.g. sliced) indices count as extent 1
    double Product_(const std::vector<std::string> &indices) const
    {
        double p = 1;
        for (const auto &index : indices) {
            const auto it = index_to_size_map_.find(index);
            if (it != index_to_size_map_.end())
                p *= static_cast<double>(it->second);
        }
        return p;
    }

    void Replay_(size_t a, size_t b)
    {
        using namespace Utilities;
        const size_t c = steps_.size();
        const auto contracted = VectorIntersection(steps_[a].tensor_indices, steps_[b].tensor_indices);
        const auto node_indices = VectorSubtraction(
            VectorConcatenation(steps_[a].node_indices, steps_[b].node_indices), contracted);
        const auto tensor_indices =
            VectorDisjunctiveUnion(steps_[a].tensor_indices, steps_[b].tensor_indices);
        const auto tags = VectorUnion(steps_[a].tags, steps_[b].tags);
        steps_[a].parent = c;
        steps_[b].parent = c;
        steps_.push_back(PathStepInfo{c, PathStepInfo::MISSING_ID, {a, b},
                                      node_indices.empty() ? "_" : JoinStringVector(node_indices),
                                      node_indices, tensor_indices, tags, contracted});
    }
};

} // namespace Jet
