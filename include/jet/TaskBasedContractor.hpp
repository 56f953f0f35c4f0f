// Jet::TaskBasedContractor<Tensor> — drop-in for
// /root/reference/include/jet/TaskBasedContractor.hpp, executing on one B200.
//
// The public bookkeeping is the reference's (task names "<id>:<name>", ":results[r]" suffix on the
// final step, name->tensor / name->parents maps, de-duplication of tasks by name across calls,
// flops/memory counters, results + reduction result, Contract() -> std::future<void>).  What runs
// inside Contract() is new: instead of one Taskflow host task per contraction that allocates a
// fresh std::vector (TaskBasedContractor.hpp:386-390), the whole DAG is lowered once —
//   * every named tensor gets an offset in ONE device arena, assigned from its lifetime when
//     deletion tasks were requested (AddDeletionTasks becomes a plan-time analysis);
//   * leaves are uploaded once, each unique contraction runs once as a fused GPU contraction
//     (jb_contract: the operand transposes are folded into the kernel's loads), in task order on
//     one stream with no host synchronisation in between;
//   * results are reduced on the device in task order (deterministic);
//   * tensors are copied back to the host only at the end, and only those the API exposes
//     (everything without deletion tasks; just the results with them).
// `num_threads` is accepted for source compatibility; the GPU stream replaces the thread pool.
#pragma once

#include <algorithm>
#include <cstdint>
#include <future>
#include <map>
#include <memory>
#include <ostream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "Abort.hpp"
#include "PathInfo.hpp"
#include "TensorNetwork.hpp"
#include "jetb200.h"

namespace Jet {

/// Handle of a scheduled task (what tf::Task is to the reference's NameToTaskMap).
class Task {
  public:
    Task() = default;
    explicit Task(std::string name) : name_(std::move(name)) {}
    const std::string &name() const { return name_; }
    bool empty() const { return name_.empty(); }

  private:
    std::string name_;
};

/// The dependency graph as data (what tf::Taskflow is to GetTaskflow()): task names and edges.
class TaskGraph {
  public:
    void AddTask(const std::string &name) { tasks_.push_back(name); }
    void AddEdge(const std::string &from, const std::string &to) { edges_.emplace_back(from, to); }
    size_t num_tasks() const { return tasks_.size(); }
    bool empty() const { return tasks_.empty(); }
    const std::vector<std::string> &tasks() const { return tasks_; }
    const std::vector<std::pair<std::string, std::string>> &edges() const { return edges_; }
    void dump(std::ostream &os) const
    {
        os << "digraph Taskflow {\n";
        for (const auto &t : tasks_)
            os << "\"" << t << "\";\n";
        for (const auto &[a, b] : edges_)
            os << "\"" << a << "\" -> \"" << b << "\";\n";
        os << "}\n";
    }

  private:
    std::vector<std::string> tasks_;
    std::vector<std::pair<std::string, std::string>> edges_;
};

template <class TensorType> class TaskBasedContractor {
  public:
    using NameToTaskMap = std::unordered_map<std::string, Task>;
    using NameToTensorMap = std::unordered_map<std::string, std::unique_ptr<TensorType>>;
    using NameToParentsMap = std::unordered_map<std::string, std::unordered_set<std::string>>;
    using TaskFlow = TaskGraph;
    using scalar_t = typename TensorType::scalar_type_t;

    TaskBasedContractor(size_t num_threads = std::thread::hardware_concurrency())
        : num_threads_(num_threads), memory_(0), flops_(0), reduced_(false), delete_(false)
    {
    }

    const NameToTaskMap &GetNameToTaskMap() const noexcept { return name_to_task_map_; }
    const NameToTensorMap &GetNameToTensorMap() const noexcept { return name_to_tensor_map_; }
    const NameToParentsMap &GetNameToParentsMap() const noexcept { return name_to_parents_map_; }
    const std::vector<TensorType> &GetResults() const noexcept { return results_; }
    const TensorType &GetReductionResult() const noexcept { return reduction_result_; }
    const TaskFlow &GetTaskflow() const noexcept { return graph_; }
    double GetFlops() const noexcept { return flops_; }
    double GetMemory() const noexcept { return memory_; }

    /// Adds the contraction tasks of one network + path; returns how many of them were already
    /// scheduled by an earlier call (shared work).
    size_t AddContractionTasks(const TensorNetwork<TensorType> &tn, const PathInfo &path_info) noexcept
    {
        const auto &path = path_info.GetPath();
        const auto &steps = path_info.GetSteps();
        if (path.empty())
            return 0;
        const auto &nodes = tn.GetNodes();
        const size_t num_leaves = nodes.size();
        const size_t result_id = results_.size();
        results_.resize(result_id + 1);

        size_t shared = 0;
        for (size_t i = 0; i < path.size(); i++) {
            const auto [id_1, id_2] = path[i];
            const std::string name_1 = TaskName_(steps[id_1]);
            const std::string name_2 = TaskName_(steps[id_2]);
            std::string name_3 = TaskName_(steps[num_leaves + i]);
            const bool last = i + 1 == path.size();
            if (last)
                name_3 += ":results[" + std::to_string(result_id) + "]";

            name_to_parents_map_[name_1].emplace(name_3);
            name_to_parents_map_[name_2].emplace(name_3);
            if (id_1 < num_leaves)
                name_to_tensor_map_.try_emplace(name_1, std::make_unique<TensorType>(nodes[id_1].tensor));
            if (id_2 < num_leaves)
                name_to_tensor_map_.try_emplace(name_2, std::make_unique<TensorType>(nodes[id_2].tensor));
            name_to_tensor_map_.try_emplace(name_3, nullptr);

            if (name_to_task_map_.count(name_3)) {
                shared++;
                continue;
            }
            flops_ += path_info.GetPathStepFlops(num_leaves + i);
            memory_ += path_info.GetPathStepMemory(num_leaves + i);

            name_to_task_map_.emplace(name_3, Task(name_3));
            graph_.AddTask(name_3);
            contractions_.push_back({name_1, name_2, name_3});
            if (id_1 >= num_leaves)
                graph_.AddEdge(name_1, name_3);
            if (id_2 >= num_leaves)
                graph_.AddEdge(name_2, name_3);
            if (last) {
                const std::string storage = name_3 + ":storage[" + std::to_string(result_id) + "]";
                graph_.AddTask(storage);
                graph_.AddEdge(name_3, storage);
                storages_.push_back({name_3, result_id});
            }
        }
        return shared;
    }

    /// Schedules the sum of all results; only the first call has an effect (returns 1, then 0).
    size_t AddReductionTask() noexcept
    {
        if (reduced_)
            return 0;
        reduced_ = true;
        graph_.AddTask("reduce");
        for (const auto &s : storages_)
            graph_.AddEdge(s.name + ":storage[" + std::to_string(s.result_id) + "]", "reduce");
        reduce_count_ = storages_.size(); // like the reference, results added later are not reduced
        return 1;
    }

    /// Every tensor that feeds a contraction is released once its last consumer has run.
    size_t AddDeletionTasks() noexcept
    {
        size_t n = 0;
        for (const auto &[name, parents] : name_to_parents_map_) {
            if (parents.empty())
                continue;
            graph_.AddTask(name + ":delete");
            for (const auto &parent : parents)
                if (name_to_task_map_.count(parent))
                    graph_.AddEdge(parent, name + ":delete");
            deleted_.insert(name);
            n++;
        }
        delete_ = delete_ || n > 0;
        return n;
    }

    /// Runs everything on the GPU; the future becomes ready when results are back on the host.
    std::future<void> Contract()
    {
        return std::async(std::launch::async, [this]() { Run_(); });
    }

  private:
    struct Contraction {
        std::string name_1, name_2, name_3;
    };
    struct Storage {
        std::string name;
        size_t result_id;
    };
    struct DeviceTensor {
        std::vector<std::string> indices;
        std::vector<int64_t> extent;
        std::vector<int32_t> modes;
        int64_t elems = 1;
        size_t offset = 0;
        int last_use = -1;
        bool is_leaf = false;
    };

    size_t num_threads_;
    TaskGraph graph_;
    NameToTaskMap name_to_task_map_;
    NameToTensorMap name_to_tensor_map_;
    NameToParentsMap name_to_parents_map_;
    std::vector<TensorType> results_;
    TensorType reduction_result_;
    double memory_;
    double flops_;
    bool reduced_;
    bool delete_;
    size_t reduce_count_ = 0;
    std::vector<Contraction> contractions_;
    std::vector<Storage> storages_;
    std::unordered_set<std::string> deleted_;

    static std::string TaskName_(const PathStepInfo &step)
    {
        return std::to_string(step.id) + ":" + step.name;
    }

    // best-fit offset allocator over a growing arena (plan-time only)
    class Arena {
      public:
        size_t Alloc(size_t bytes)
        {
            bytes = Round_(bytes);
            auto best = free_.end();
            for (auto it = free_.begin(); it != free_.end(); ++it)
                if (it->second >= bytes && (best == free_.end() || it->second < best->second))
                    best = it;
            if (best != free_.end()) {
                const size_t off = best->first, size = best->second;
                free_.erase(best);
                if (size > bytes)
                    free_[off + bytes] = size - bytes;
                return off;
            }
            const size_t off = top_;
            top_ += bytes;
            return off;
        }
        void Free(size_t off, size_t bytes)
        {
            bytes = Round_(bytes);
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
        size_t Top() const { return top_; }

      private:
        static size_t Round_(size_t b) { return (std::max<size_t>(b, 1) + 511) & ~size_t(511); }
        std::map<size_t, size_t> free_;
        size_t top_ = 0;
    };

    void Run_()
    {
        constexpr int dtype = TensorHelpers::DtypeCode<scalar_t>();
        constexpr size_t eb = sizeof(scalar_t);
        if (contractions_.empty() && storages_.empty())
            return;

        // ---- describe every named tensor ------------------------------------------------------
        std::unordered_map<std::string, int32_t> label;
        std::unordered_map<std::string, DeviceTensor> dt;
        auto modes_of = [&label](const std::vector<std::string> &idx) {
            std::vector<int32_t> m(idx.size());
            for (size_t i = 0; i < idx.size(); i++)
                m[i] = label.emplace(idx[i], static_cast<int32_t>(label.size())).first->second;
            return m;
        };
        auto describe_leaf = [&](const std::string &name) {
            if (dt.count(name))
                return;
            const auto it = name_to_tensor_map_.find(name);
            JET_ABORT_IF(it == name_to_tensor_map_.end() || it->second == nullptr,
                         "Tensor '" + name + "' is not available for contraction.");
            DeviceTensor d;
            d.indices = it->second->GetIndices();
            d.extent.assign(it->second->GetShape().begin(), it->second->GetShape().end());
            d.modes = modes_of(d.indices);
            d.elems = static_cast<int64_t>(it->second->GetSize());
            d.is_leaf = true;
            dt.emplace(name, std::move(d));
        };
        std::unordered_set<std::string> produced;
        for (const auto &c : contractions_)
            produced.insert(c.name_3);
        size_t ws_bytes = 0;
        for (size_t t = 0; t < contractions_.size(); t++) {
            const auto &c = contractions_[t];
            if (!produced.count(c.name_1))
                describe_leaf(c.name_1);
            if (!produced.count(c.name_2))
                describe_leaf(c.name_2);
            const DeviceTensor &A = dt.at(c.name_1);
            const DeviceTensor &B = dt.at(c.name_2);
            jb_contract_info_t info;
            JET_JB_CHECK(jb_contract_info(dtype, static_cast<int>(A.extent.size()), A.extent.data(),
                                          A.modes.data(), static_cast<int>(B.extent.size()),
                                          B.extent.data(), B.modes.data(), &info));
            DeviceTensor C;
            std::unordered_map<int32_t, std::string> name_of;
            for (size_t i = 0; i < A.modes.size(); i++)
                name_of[A.modes[i]] = A.indices[i];
            for (size_t i = 0; i < B.modes.size(); i++)
                name_of[B.modes[i]] = B.indices[i];
            for (int i = 0; i < info.rank_c; i++) {
                C.modes.push_back(info.modes_c[i]);
                C.extent.push_back(info.extent_c[i]);
                C.indices.push_back(name_of.at(info.modes_c[i]));
            }
            C.elems = info.m * info.n;
            ws_bytes = std::max(ws_bytes, info.ws_bytes);
            dt[c.name_3] = std::move(C);
            dt[c.name_1].last_use = static_cast<int>(t);
            dt[c.name_2].last_use = static_cast<int>(t);
        }

        // ---- arena offsets from lifetimes ------------------------------------------------------
        Arena arena;
        const size_t ws_off = arena.Alloc(std::max<size_t>(ws_bytes, 512));
        for (auto &[name, d] : dt)
            if (d.is_leaf)
                d.offset = arena.Alloc(eb * d.elems);
        for (size_t t = 0; t < contractions_.size(); t++) {
            const auto &c = contractions_[t];
            dt[c.name_3].offset = arena.Alloc(eb * dt[c.name_3].elems);
            for (const std::string *in : {&c.name_1, &c.name_2}) {
                DeviceTensor &I = dt[*in];
                if (delete_ && deleted_.count(*in) && I.last_use == static_cast<int>(t))
                    arena.Free(I.offset, eb * I.elems);
            }
        }
        size_t acc_off = 0;
        int64_t acc_elems = 0;
        if (reduced_ && reduce_count_ > 0) {
            acc_elems = dt.at(storages_[0].name).elems;
            acc_off = arena.Alloc(eb * acc_elems);
        }

        // ---- run ---------------------------------------------------------------------------------
        struct Guard {
            void *arena = nullptr, *stream = nullptr;
            ~Guard()
            {
                if (stream)
                    jb_stream_destroy(stream);
                if (arena)
                    jb_free(arena);
            }
        } g;
        JET_JB_CHECK(jb_malloc(&g.arena, arena.Top()));
        JET_JB_CHECK(jb_stream_create(&g.stream));
        auto at = [&g](size_t off) { return static_cast<void *>(static_cast<char *>(g.arena) + off); };

        for (const auto &[name, d] : dt)
            if (d.is_leaf)
                JET_JB_CHECK(jb_memcpy_h2d(at(d.offset), name_to_tensor_map_.at(name)->GetData().data(),
                                           eb * d.elems, g.stream));
        for (const auto &c : contractions_) {
            const DeviceTensor &A = dt.at(c.name_1);
            const DeviceTensor &B = dt.at(c.name_2);
            JET_JB_CHECK(jb_contract(dtype, static_cast<int>(A.extent.size()), A.extent.data(),
                                     A.modes.data(), at(A.offset), static_cast<int>(B.extent.size()),
                                     B.extent.data(), B.modes.data(), at(B.offset),
                                     at(dt.at(c.name_3).offset), at(ws_off), ws_bytes, g.stream));
        }
        // reduction on the device, in result order
        if (acc_elems > 0) {
            const DeviceTensor &first = dt.at(storages_[0].name);
            JET_JB_CHECK(jb_memcpy_d2d(at(acc_off), at(first.offset), eb * acc_elems, g.stream));
            for (size_t r = 1; r < reduce_count_; r++) {
                const DeviceTensor &R = dt.at(storages_[r].name);
                JET_ABORT_IF_NOT(R.elems == acc_elems &&
                                     Utilities::VectorDisjunctiveUnion(R.indices, first.indices).empty(),
                                 "Tensor addition with disjoint indices is not supported.");
                const void *src = at(R.offset);
                if (R.indices != first.indices) {
                    // align to the first result's index order (AddTensors semantics)
                    std::vector<int32_t> perm(first.indices.size());
                    for (size_t j = 0; j < perm.size(); j++)
                        perm[j] = static_cast<int32_t>(
                            std::find(R.indices.begin(), R.indices.end(), first.indices[j]) -
                            R.indices.begin());
                    JET_ABORT_IF(ws_bytes < eb * static_cast<size_t>(acc_elems),
                                 "Results with permuted indices cannot be reduced on the device.");
                    JET_JB_CHECK(jb_permute(dtype, at(R.offset), at(ws_off), static_cast<int>(perm.size()),
                                            R.extent.data(), perm.data(), g.stream));
                    src = at(ws_off);
                }
                JET_JB_CHECK(jb_add(dtype, acc_elems, at(acc_off), src, at(acc_off), g.stream));
            }
        }

        // ---- copy back what the API exposes --------------------------------------------------------
        auto download = [&](const DeviceTensor &d, size_t off) {
            std::vector<size_t> shape(d.extent.begin(), d.extent.end());
            TensorType t(d.indices, shape);
            JET_JB_CHECK(jb_memcpy_d2h(t.GetData().data(), at(off), eb * d.elems, g.stream));
            return t;
        };
        for (const auto &s : storages_)
            results_[s.result_id] = download(dt.at(s.name), dt.at(s.name).offset);
        for (const auto &c : contractions_) {
            if (delete_ && deleted_.count(c.name_3))
                name_to_tensor_map_[c.name_3] = nullptr;
            else
                name_to_tensor_map_[c.name_3] =
                    std::make_unique<TensorType>(download(dt.at(c.name_3), dt.at(c.name_3).offset));
        }
        if (delete_)
            for (const auto &name : deleted_)
                name_to_tensor_map_[name] = nullptr;
        if (acc_elems > 0)
            reduction_result_ = download(dt.at(storages_[0].name), acc_off);
        JET_JB_CHECK(jb_stream_sync(g.stream));
    }
};

template <class TensorType>
inline std::ostream &operator<<(std::ostream &out, const TaskBasedContractor<TensorType> &tbc)
{
    tbc.GetTaskflow().dump(out);
    return out;
}

} // namespace Jet
