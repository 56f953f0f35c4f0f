// Jet::TensorNetwork<Tensor> — drop-in for /root/reference/include/jet/TensorNetwork.hpp.
// Host-side graph bookkeeping (nodes, index->edge map, tags, slicing, serial contraction in path
// order); every tensor operation it triggers (ContractTensors, SliceIndex, Transpose) runs on the
// GPU through Jet::Tensor.  For whole-network GPU-resident contraction use TaskBasedContractor or
// SlicedContractor.
#pragma once

#include <algorithm>
#include <iostream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "Abort.hpp"
#include "Utilities.hpp"

namespace Jet {

template <class Tensor> class TensorNetwork {
  public:
    using NodeID_t = size_t;

    struct Node {
        NodeID_t id;
        std::string name;
        std::vector<std::string> indices; // keep "(v)" annotations of sliced indices
        std::vector<std::string> tags;
        bool contracted;
        Tensor tensor;
    };

    struct Edge {
        size_t dim;
        std::vector<NodeID_t> node_ids;
        bool operator==(const Edge &other) const noexcept
        {
            const std::unordered_set<size_t> a(node_ids.begin(), node_ids.end());
            const std::unordered_set<size_t> b(other.node_ids.begin(), other.node_ids.end());
            return dim == other.dim && a == b;
        }
    };

    using Nodes = std::vector<Node>;
    using IndexToEdgeMap = std::unordered_map<std::string, Edge>;
    using TagToNodeIdsMap = std::unordered_multimap<std::string, NodeID_t>;
    using Path = std::vector<std::pair<NodeID_t, NodeID_t>>;

    const Nodes &GetNodes() const noexcept { return nodes_; }
    const IndexToEdgeMap &GetIndexToEdgeMap() const noexcept { return index_to_edge_map_; }
    const TagToNodeIdsMap &GetTagToNodesMap() const noexcept { return tag_to_nodes_map_; }
    const Path &GetPath() noexcept { return path_; }
    size_t NumIndices() const noexcept { return index_to_edge_map_.size(); }
    size_t NumTensors() const noexcept { return nodes_.size(); }

    NodeID_t AddTensor(const Tensor &tensor, const std::vector<std::string> &tags) noexcept
    {
        const NodeID_t id = nodes_.size();
        nodes_.push_back(Node{id, NameOf_(tensor.GetIndices()), tensor.GetIndices(), tags, false, tensor});
        RegisterIndices_(nodes_.back());
        for (const auto &tag : tags)
            tag_to_nodes_map_.emplace(tag, id);
        return id;
    }

    /// Fixes each listed index to the digit of `value` (row-major over the listed indices, first
    /// index slowest).  Tensors lose the axis; node names keep it, annotated "idx(v)".
    void SliceIndices(const std::vector<std::string> &indices, unsigned long long value)
    {
        std::vector<size_t> dims(indices.size());
        for (size_t i = 0; i < indices.size(); i++) {
            const auto it = index_to_edge_map_.find(indices[i]);
            JET_ABORT_IF(it == index_to_edge_map_.end(), "Sliced index does not exist.");
            dims[i] = it->second.dim;
        }
        const auto digits = Utilities::UnravelIndex(value, dims);
        for (size_t i = 0; i < indices.size(); i++) {
            const std::string &index = indices[i];
            const std::vector<NodeID_t> touched = index_to_edge_map_.at(index).node_ids;
            for (const NodeID_t id : touched) {
                Node &node = nodes_[id];
                node.tensor = Tensor::SliceIndex(node.tensor, index, digits[i]);
                for (auto &label : node.indices) {
                    if (label == index)
                        label += "(" + std::to_string(digits[i]) + ")";
                }
                node.name = NameOf_(node.indices);
            }
        }
        for (const auto &index : indices)
            index_to_edge_map_.erase(index);
    }

    /// Contracts along `path` (pairs of node ids; step i appends node NumTensors()), or, with an
    /// empty path, every 2-node edge and then all scalars.  Returns the last node's tensor.
    const Tensor &Contract(const Path &path = {})
    {
        JET_ABORT_IF(nodes_.empty(), "An empty tensor network cannot be contracted.");
        if (!path.empty()) {
            for (const auto &[a, b] : path) {
                JET_ABORT_IF_NOT(a < nodes_.size(), "Node ID 1 in contraction pair is invalid.");
                JET_ABORT_IF_NOT(b < nodes_.size(), "Node ID 2 in contraction pair is invalid.");
                ContractPair_(a, b);
            }
            path_ = path;
        }
        else {
            std::vector<std::string> shared;
            for (const auto &kv : index_to_edge_map_)
                shared.push_back(kv.first);
            for (const auto &index : shared) {
                const auto it = index_to_edge_map_.find(index);
                if (it == index_to_edge_map_.end() || it->second.node_ids.size() != 2)
                    continue;
                const NodeID_t a = it->second.node_ids[0], b = it->second.node_ids[1];
                ContractPair_(a, b);
                path_.emplace_back(a, b);
            }
            std::vector<NodeID_t> scalars;
            for (const auto &node : nodes_)
                if (node.tensor.GetIndices().empty())
                    scalars.push_back(node.id);
            if (scalars.size() >= 2) {
                NodeID_t acc = scalars[0];
                for (size_t i = 1; i < scalars.size(); i++) {
                    path_.emplace_back(acc, scalars[i]);
                    acc = ContractPair_(acc, scalars[i]);
                }
            }
        }
        return nodes_.back().tensor;
    }

  private:
    Nodes nodes_;
    IndexToEdgeMap index_to_edge_map_;
    TagToNodeIdsMap tag_to_nodes_map_;
    Path path_;

    static std::string NameOf_(const std::vector<std::string> &indices)
    {
        return indices.empty() ? "_" : Utilities::JoinStringVector(indices);
    }

    void RegisterIndices_(const Node &node)
    {
        const auto &shape = node.tensor.GetShape();
        for (size_t i = 0; i < node.indices.size(); i++) {
            if (shape[i] < 2)
                continue; // extent-1 axes do not form edges
            auto it = index_to_edge_map_.find(node.indices[i]);
            if (it == index_to_edge_map_.end())
                index_to_edge_map_.emplace(node.indices[i], Edge{shape[i], {node.id}});
            else
                it->second.node_ids.push_back(node.id);
        }
    }

    NodeID_t ContractPair_(NodeID_t a, NodeID_t b)
    {
        using namespace Utilities;
        Tensor result = Tensor::ContractTensors(nodes_[a].tensor, nodes_[b].tensor);
        nodes_[a].contracted = true;
        nodes_[b].contracted = true;
        const NodeID_t c = nodes_.size();
        const auto indices = VectorDisjunctiveUnion(nodes_[a].indices, nodes_[b].indices);
        const auto tags = VectorUnion(nodes_[a].tags, nodes_[b].tags);
        const auto contracted =
            VectorIntersection(nodes_[a].tensor.GetIndices(), nodes_[b].tensor.GetIndices());
        nodes_.push_back(Node{c, NameOf_(indices), indices, tags, false, std::move(result)});
        // the new node inherits the surviving edges of its children
        for (const auto &index : nodes_[c].indices) {
            const auto it = index_to_edge_map_.find(index);
            if (it == index_to_edge_map_.end())
                continue;
            for (auto &id : it->second.node_ids)
                if (id == a || id == b)
                    id = c;
        }
        for (const auto &index : contracted)
            index_to_edge_map_.erase(index);
        return c;
    }
};

template <class Tensor>
inline std::ostream &operator<<(std::ostream &out, const TensorNetwork<Tensor> &tn)
{
    using namespace Jet::Utilities;
    out << "Printing Nodes" << std::endl;
    for (const auto &node : tn.GetNodes())
        out << node.id << ' ' << node.name << ' ' << node.tags << std::endl;
    out << "Printing Edges" << std::endl;
    for (const auto &[index, edge] : tn.GetIndexToEdgeMap())
        out << index << ' ' << edge.dim << ' ' << edge.node_ids << std::endl;
    return out;
}

} // namespace Jet
