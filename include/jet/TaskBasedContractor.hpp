// Jet::TaskBasedContractor<Tensor> — drop-in for
// /root/reference/include/jet/TaskBasedContractor.hpp, executing on one B200.
//
// The public bookkeeping is the reference's (task names "<id>:<name>", ":results[r]" suffix on the
// final step, name->tensor / name->parents maps, de-duplication of tasks by name across calls,
// flops/memory counters, results + reduction result, Contract() -> std::future<void>).  What runs
// inside Contract() is new.  The reference executes one Taskflow host task per contraction, each
// allocating a fresh std::vector (TaskBasedContractor.hpp:386-390); its sliced benchmarks
// (examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp:69-93) add 2^s SliceIndices copies of one
// network and rely on the de-duplication of equal task names (TaskBasedContractor.hpp:216-222) to share
// the slice-independent work.  Here Contract()
//   * groups the added networks by structure: copies made by TensorNetwork::SliceIndices differ only in
//     the "idx(value)" annotations of their node names (TensorNetwork.hpp:266-274), so a group is one
//     network + path + sliced indices, and every member is one slice id of it;
//   * lowers each group to ONE device-resident plan set (jb_multi_*, include/jetb200.h): the unsliced
//     leaves are rebuilt from the members and uploaded once, the slice-independent subtrees run once
//     (the plan-time form of the name de-duplication), per-slice steps run as fused chains from a CUDA
//     graph with several slices in flight, and the per-slice results are summed in FP64 on the device;
//     unrelated networks each get their own plan;
//   * copies back only what the API must expose at once: GetResults() and GetReductionResult().
// GetNameToTensorMap() keeps the reference's contents (every intermediate of every network unless
// deletion tasks were added), but those tensors are produced on first access: asking for the map after
// Contract() replays the tasks one fused GPU contraction (jb_contract) per task and downloads them.
// Environment: JET_B200_DEVICES="0,1,.."|"all" spreads slices over GPUs of this process,
// JET_B200_LANES=n fixes the slices in flight per GPU (default: automatic), JET_B200_TBC=stepwise forces
// the task-per-kernel path for everything.
// `num_threads` is accepted for source compatibility; streams replace the thread pool.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <ostream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "Abort.hpp"
#include "PathInfo.hpp"
#include "PlanCache.hpp"
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
    /// Every named tensor (nullptr while not computed / after deletion).  Intermediates are materialised on
    /// the first call after Contract() (see the header comment); results and leaves are always in place.
    const NameToTensorMap &GetNameToTensorMap() const
    {
        if (intermediates_pending_)
            const_cast<TaskBasedContractor *>(this)->MaterialiseIntermediates_();
        return name_to_tensor_map_;
    }
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

        NetworkRecord rec;
        rec.result_id = result_id;
        rec.num_leaves = num_leaves;
        rec.path = path;
        rec.leaves.resize(num_leaves);

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
            if (id_1 < num_leaves) {
                const auto it = name_to_tensor_map_.try_emplace(name_1, std::make_unique<TensorType>(nodes[id_1].tensor)).first;
                rec.leaves[id_1] = LeafRecord{it->second.get(), steps[id_1].node_indices};
            }
            if (id_2 < num_leaves) {
                const auto it = name_to_tensor_map_.try_emplace(name_2, std::make_unique<TensorType>(nodes[id_2].tensor)).first;
                rec.leaves[id_2] = LeafRecord{it->second.get(), steps[id_2].node_indices};
            }
            name_to_tensor_map_.try_emplace(name_3, nullptr);
            if (last)
                rec.final_name = name_3;

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
        networks_.push_back(std::move(rec));
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
    // one used leaf of an added network: its tensor (owned by name_to_tensor_map_; stable address) and the
    // node's index labels, where sliced indices appear as "idx(value)" (TensorNetwork.hpp:266-274)
    struct LeafRecord {
        const TensorType *tensor = nullptr;
        std::vector<std::string> node_indices;
    };
    // what one AddContractionTasks call described: enough to lower it onto a device plan
    struct NetworkRecord {
        size_t result_id = 0;
        size_t num_leaves = 0;
        PathInfo::Path path;
        std::vector<LeafRecord> leaves; // indexed by node id; tensor == nullptr for leaves the path never touches
        std::string final_name;
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
    std::vector<NetworkRecord> networks_;
    bool intermediates_pending_ = false;

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

    // ---- Contract(): plan sets for the results, tasks replayed one by one only for what they cannot give ----
    void Run_()
    {
        if (contractions_.empty() && storages_.empty())
            return;
        const char *mode = std::getenv("JET_B200_TBC");
        const bool stepwise = mode != nullptr && std::string(mode) == "stepwise";
        // names that must hold a tensor after Contract() besides the results: every contraction output that no
        // deletion task removes
        std::unordered_set<std::string> final_names;
        for (const auto &st : storages_)
            final_names.insert(st.name);
        bool survivors = false;
        for (const auto &c : contractions_) {
            if (!deleted_.count(c.name_3) && !final_names.count(c.name_3)) {
                survivors = true;
                break;
            }
        }
        if (stepwise || (delete_ && survivors)) {
            // (deletion tasks cover only part of the graph: the leaves some surviving tensors would be
            // re-derived from are gone after Contract(), so everything is produced now)
            RunStepwise_(true);
            intermediates_pending_ = false;
            return;
        }
        RunPlans_();
        for (const auto &net : networks_)
            name_to_tensor_map_[net.final_name] = std::make_unique<TensorType>(results_[net.result_id]);
        if (delete_)
            for (const auto &name : deleted_)
                name_to_tensor_map_[name] = nullptr;
        intermediates_pending_ = survivors;
    }

    void MaterialiseIntermediates_()
    {
        intermediates_pending_ = false;
        RunStepwise_(false);
    }

    // ---- lowering onto plan sets -----------------------------------------------------------------------
    struct Group {
        std::vector<size_t> members; // indices into networks_, in the order they were added
    };

    // "idx(value)" -> (idx, value) for a node label that is not an index of the tensor
    static bool SplitSliced_(const std::string &label, size_t *prefix_len, size_t *value) noexcept
    {
        if (label.size() < 4 || label.back() != ')')
            return false;
        const size_t open = label.rfind('(');
        if (open == std::string::npos || open == 0 || open + 2 >= label.size())
            return false; // needs a name before '(' and at least one digit inside
        size_t v = 0;
        for (size_t i = open + 1; i + 1 < label.size(); i++) {
            if (label[i] < '0' || label[i] > '9')
                return false;
            v = v * 10 + static_cast<size_t>(label[i] - '0');
        }
        *prefix_len = open;
        *value = v;
        return true;
    }
    static bool ParseSliced_(const std::string &label, std::string *index, size_t *value)
    {
        size_t n = 0;
        if (!SplitSliced_(label, &n, value))
            return false;
        *index = label.substr(0, n);
        return true;
    }

    // the structure of a network with the slice values blanked out: equal keys <=> slices of one network.
    // Two independent 64-bit hashes over (path, per leaf: node labels with "(value)" blanked, tensor indices,
    // shape), mixed a word at a time and built without materialising the text (1024 copies x 400 leaves in ~3 ms).
    struct StructureKey {
        uint64_t a = 14695981039346656037ull, b = 0x9E3779B97F4A7C15ull;
        void Word(uint64_t v) noexcept // two independent multiply-xorshift mixers, one step per 64-bit word
        {
            a = (a ^ v) * 0x100000001B3ull;
            a ^= a >> 29;
            b = (b + v + 0x632BE59BD9B4E019ull) * 0xD6E8FEB86659FD93ull;
            b ^= b >> 32;
        }
        void Byte(unsigned char c) noexcept { Word(0xA500u | c); }
        void Text(const std::string &t, size_t n) noexcept
        {
            size_t i = 0;
            for (; i + 8 <= n; i += 8) {
                uint64_t w;
                std::memcpy(&w, t.data() + i, 8);
                Word(w);
            }
            uint64_t tail = 0;
            std::memcpy(&tail, t.data() + i, n - i);
            Word(tail ^ (static_cast<uint64_t>(n) << 56) ^ 0xFF00000000000000ull);
        }
        void Number(uint64_t v) noexcept { Word(v); }
        bool operator==(const StructureKey &o) const noexcept { return a == o.a && b == o.b; }
    };
    struct StructureKeyHash {
        size_t operator()(const StructureKey &k) const noexcept { return static_cast<size_t>(k.a ^ (k.b >> 1)); }
    };

    static StructureKey StructureKey_(const NetworkRecord &net)
    {
        StructureKey key;
        for (const auto &[a, b] : net.path) {
            key.Number(a);
            key.Number(b);
        }
        for (const auto &leaf : net.leaves) {
            key.Byte(0xFE);
            if (leaf.tensor == nullptr)
                continue;
            // A leaf no sliced index touches is the SAME tensor object in every copy (AddContractionTasks keeps one
            // tensor per task name): its address stands for labels + indices + shape.  Only the few sliced leaves are
            // hashed by content, with the slice values blanked.
            bool annotated = false;
            for (const auto &label : leaf.node_indices)
                annotated = annotated || (!label.empty() && label.back() == ')');
            if (!annotated) {
                key.Number(reinterpret_cast<uintptr_t>(leaf.tensor));
                continue;
            }
            const auto &tidx = leaf.tensor->GetIndices();
            for (const auto &label : leaf.node_indices) {
                size_t prefix = 0, value = 0;
                if (!label.empty() && label.back() == ')' && std::find(tidx.begin(), tidx.end(), label) == tidx.end() &&
                    SplitSliced_(label, &prefix, &value)) {
                    key.Text(label, prefix);
                    key.Byte(0xFD); // a sliced label: the value is not part of the structure
                }
                else {
                    key.Text(label, label.size());
                }
            }
            key.Byte(0xFC);
            for (size_t i = 0; i < tidx.size(); i++) {
                key.Text(tidx[i], tidx[i].size());
                key.Number(leaf.tensor->GetShape()[i]);
            }
        }
        return key;
    }

    static std::vector<int> DevicesFromEnv_()
    {
        std::vector<int> devices;
        const char *e = std::getenv("JET_B200_DEVICES");
        if (e == nullptr || e[0] == 0)
            return {0};
        if (std::string(e) == "all") {
            int n = 1;
            JET_JB_CHECK(jb_device_count(&n));
            for (int d = 0; d < n; d++)
                devices.push_back(d);
            return devices;
        }
        int cur = -1;
        for (const char *c = e;; c++) {
            if (*c >= '0' && *c <= '9') {
                cur = (cur < 0 ? 0 : cur * 10) + (*c - '0');
            }
            else {
                if (cur >= 0)
                    devices.push_back(cur);
                cur = -1;
                if (*c == 0)
                    break;
            }
        }
        return devices.empty() ? std::vector<int>{0} : devices;
    }

    /// Owns a plan set for the duration of one group: back to the cache when the group succeeded, destroyed otherwise.
    struct MultiGuard {
        jb_multi *m = nullptr;
        std::string key;
        size_t bytes = 0;
        bool reusable = false;
        ~MultiGuard()
        {
            if (m == nullptr)
                return;
            if (reusable)
                detail::PlanCache::Get().CheckIn(std::move(key), m, bytes);
            else
                jb_multi_destroy(m);
        }
    };

    void RunPlans_()
    {
        constexpr int dtype = TensorHelpers::DtypeCode<scalar_t>();
        using R = typename scalar_t::value_type;

        // ---- group the networks by structure (first-seen order) ----------------------------------------
        std::vector<Group> groups;
        {
            std::unordered_map<StructureKey, size_t, StructureKeyHash> group_of;
            for (size_t n = 0; n < networks_.size(); n++) {
                const auto it = group_of.emplace(StructureKey_(networks_[n]), groups.size()).first;
                if (it->second == groups.size())
                    groups.emplace_back();
                groups[it->second].members.push_back(n);
            }
        }
        const std::vector<int> devices = DevicesFromEnv_();
        int lanes = 0;
        if (const char *e = std::getenv("JET_B200_LANES"))
            lanes = std::max(0, std::min(5, std::atoi(e)));

        // the reduction over the first reduce_count_ results, in double, in the index order of result 0
        std::vector<double> total;
        std::vector<std::string> total_indices;
        std::vector<size_t> total_shape;

        for (const Group &group : groups) {
            const NetworkRecord &first = networks_[group.members[0]];
            // ---- sliced indices of the group: labels of the first member, digits of every member --------
            std::vector<std::string> sliced_names;
            std::unordered_map<std::string, size_t> sliced_pos;
            struct LeafSlicing {
                std::vector<size_t> label_pos;  // positions in node_indices that carry "idx(v)"
                std::vector<size_t> sliced_ids; // which sliced index each of them is
            };
            std::vector<LeafSlicing> leaf_slicing(first.leaves.size());
            for (size_t l = 0; l < first.leaves.size(); l++) {
                const LeafRecord &leaf = first.leaves[l];
                if (leaf.tensor == nullptr)
                    continue;
                const auto &tidx = leaf.tensor->GetIndices();
                for (size_t q = 0; q < leaf.node_indices.size(); q++) {
                    std::string index;
                    size_t value = 0;
                    if (std::find(tidx.begin(), tidx.end(), leaf.node_indices[q]) != tidx.end() ||
                        !ParseSliced_(leaf.node_indices[q], &index, &value))
                        continue;
                    const auto it = sliced_pos.emplace(index, sliced_names.size()).first;
                    if (it->second == sliced_names.size())
                        sliced_names.push_back(index);
                    leaf_slicing[l].label_pos.push_back(q);
                    leaf_slicing[l].sliced_ids.push_back(it->second);
                }
            }
            const size_t ns = sliced_names.size();
            std::vector<std::vector<size_t>> digits(group.members.size(), std::vector<size_t>(ns, 0));
            std::vector<int64_t> dims(ns, 1);
            for (size_t g = 0; g < group.members.size(); g++) {
                const NetworkRecord &net = networks_[group.members[g]];
                for (size_t l = 0; l < net.leaves.size(); l++)
                    for (size_t q = 0; q < leaf_slicing[l].label_pos.size(); q++) {
                        size_t prefix = 0, value = 0;
                        SplitSliced_(net.leaves[l].node_indices[leaf_slicing[l].label_pos[q]], &prefix, &value);
                        digits[g][leaf_slicing[l].sliced_ids[q]] = value;
                        dims[leaf_slicing[l].sliced_ids[q]] =
                            std::max<int64_t>(dims[leaf_slicing[l].sliced_ids[q]], static_cast<int64_t>(value) + 1);
                    }
            }

            // ---- the unsliced network: used leaves only, sliced axes restored in front --------------------
            std::unordered_map<std::string, int32_t> label;
            std::vector<std::string> names;
            auto mode_of = [&](const std::string &index) {
                const auto it = label.emplace(index, static_cast<int32_t>(label.size())).first;
                if (static_cast<size_t>(it->second) == names.size())
                    names.push_back(index);
                return it->second;
            };
            std::vector<int32_t> sliced_modes;
            for (const auto &name : sliced_names)
                sliced_modes.push_back(mode_of(name));
            std::vector<int32_t> new_id(first.num_leaves + first.path.size(), -1);
            std::vector<int32_t> rank, mode;
            std::vector<int64_t> extent;
            std::vector<const void *> data;
            std::vector<std::vector<scalar_t>> rebuilt; // storage of the leaves that had to be reassembled
            rebuilt.reserve(first.leaves.size());
            int32_t used = 0;
            for (size_t l = 0; l < first.leaves.size(); l++) {
                const LeafRecord &leaf = first.leaves[l];
                if (leaf.tensor == nullptr)
                    continue;
                new_id[l] = used++;
                const LeafSlicing &ls = leaf_slicing[l];
                rank.push_back(static_cast<int32_t>(ls.sliced_ids.size() + leaf.tensor->GetIndices().size()));
                size_t blocks = 1;
                for (const size_t sid : ls.sliced_ids) {
                    mode.push_back(sliced_modes[sid]);
                    extent.push_back(dims[sid]);
                    blocks *= static_cast<size_t>(dims[sid]);
                }
                for (size_t i = 0; i < leaf.tensor->GetIndices().size(); i++) {
                    mode.push_back(mode_of(leaf.tensor->GetIndices()[i]));
                    extent.push_back(static_cast<int64_t>(leaf.tensor->GetShape()[i]));
                }
                if (ls.sliced_ids.empty()) {
                    data.push_back(leaf.tensor->GetData().data());
                    continue;
                }
                // block b of the rebuilt leaf = the member whose digits ravel to b (row-major over the leaf's own
                // sliced axes); blocks no member selects stay zero and are never read
                const size_t elems = leaf.tensor->GetSize();
                rebuilt.emplace_back(blocks * elems);
                std::vector<char> filled(blocks, 0);
                for (size_t g = 0; g < group.members.size(); g++) {
                    size_t b = 0;
                    for (const size_t sid : ls.sliced_ids)
                        b = b * static_cast<size_t>(dims[sid]) + digits[g][sid];
                    if (filled[b])
                        continue;
                    filled[b] = 1;
                    const TensorType *t = networks_[group.members[g]].leaves[l].tensor;
                    std::memcpy(rebuilt.back().data() + b * elems, t->GetData().data(), sizeof(scalar_t) * elems);
                }
                data.push_back(rebuilt.back().data());
            }
            std::vector<int32_t> flat_path;
            for (size_t i = 0; i < first.path.size(); i++) {
                new_id[first.num_leaves + i] = used + static_cast<int32_t>(i);
                flat_path.push_back(new_id[first.path[i].first]);
                flat_path.push_back(new_id[first.path[i].second]);
            }
            std::vector<int64_t> slice_ids(group.members.size(), 0);
            for (size_t g = 0; g < group.members.size(); g++)
                for (size_t q = 0; q < ns; q++)
                    slice_ids[g] = slice_ids[g] * dims[q] + static_cast<int64_t>(digits[g][q]);

            jb_network_desc_t d{};
            d.dtype = dtype;
            d.device = devices[0];
            d.num_leaves = used;
            d.rank = rank.data();
            d.extent = extent.data();
            d.mode = mode.data();
            d.h_data = data.data();
            d.num_steps = static_cast<int32_t>(first.path.size());
            d.path = flat_path.data();
            d.num_sliced = static_cast<int32_t>(ns);
            d.sliced_modes = sliced_modes.data();
            d.flags = JB_PLAN_STORE_RESULTS;
            MultiGuard guard;
            const int nd = static_cast<int>(std::min<size_t>(devices.size(), group.members.size()));
            const int use_lanes = static_cast<int>(std::min<size_t>(lanes, group.members.size()));
            // everything the plans depend on except the leaf data
            {
                auto put = [&](const void *ptr, size_t bytes) {
                    guard.key.append(static_cast<const char *>(ptr), bytes);
                    guard.key.push_back('|');
                };
                const int32_t head[4] = {dtype, nd, use_lanes, static_cast<int32_t>(d.flags)};
                put(head, sizeof(head));
                put(devices.data(), sizeof(int) * static_cast<size_t>(nd));
                put(rank.data(), sizeof(int32_t) * rank.size());
                put(extent.data(), sizeof(int64_t) * extent.size());
                put(mode.data(), sizeof(int32_t) * mode.size());
                put(flat_path.data(), sizeof(int32_t) * flat_path.size());
                put(sliced_modes.data(), sizeof(int32_t) * sliced_modes.size());
            }
            guard.m = detail::PlanCache::Get().CheckOut(guard.key);
            if (guard.m != nullptr) {
                // same structure as an earlier Contract(): new leaves into the existing plans
                JET_JB_CHECK(jb_multi_upload(guard.m, data.data()));
                JET_JB_CHECK(jb_multi_sync(guard.m)); // the host buffers in `rebuilt` are released below
                JET_JB_CHECK(jb_multi_reset(guard.m));
            }
            else {
                // idle plan sets of OTHER structures hold device memory and constant-bank slots (a plan that finds no
                // free slot runs without fused chains): they go before a new plan set is built
                detail::PlanCache::Get().Flush();
                JET_JB_CHECK(jb_multi_create(&d, nd, devices.data(), use_lanes, &guard.m));
            }
            rebuilt.clear(); // uploaded
            jb_plan_stats_t stats;
            JET_JB_CHECK(jb_multi_stats(guard.m, &stats));
            {
                int plan_devices = 1, plan_lanes = 1;
                JET_JB_CHECK(jb_multi_num_plans(guard.m, &plan_devices, &plan_lanes));
                guard.bytes = stats.arena_bytes * static_cast<size_t>(plan_devices) * static_cast<size_t>(plan_lanes);
            }
            // a network added twice is one slice id: it runs once and its result is handed to every copy
            std::vector<int64_t> run_ids;
            std::vector<size_t> ordinal_of(group.members.size());
            {
                std::unordered_map<int64_t, size_t> seen;
                for (size_t g = 0; g < group.members.size(); g++) {
                    const auto it = seen.emplace(slice_ids[g], run_ids.size()).first;
                    if (it->second == run_ids.size())
                        run_ids.push_back(slice_ids[g]);
                    ordinal_of[g] = it->second;
                }
            }
            const bool duplicates = run_ids.size() != group.members.size();
            JET_JB_CHECK(jb_multi_run_list(guard.m, run_ids.data(), static_cast<int64_t>(run_ids.size())));

            // ---- results ------------------------------------------------------------------------------------
            std::vector<std::string> indices;
            std::vector<size_t> shape;
            for (int i = 0; i < stats.result_rank; i++) {
                indices.push_back(names[static_cast<size_t>(stats.result_modes[i])]);
                shape.push_back(static_cast<size_t>(stats.result_extent[i]));
            }
            const size_t elems = static_cast<size_t>(stats.result_elems);
            std::vector<scalar_t> all(elems * run_ids.size());
            JET_JB_CHECK(jb_multi_slice_results(guard.m, all.data()));
            bool all_reduced = reduced_ && !duplicates, none_reduced = true;
            for (size_t g = 0; g < group.members.size(); g++) {
                const size_t rid = networks_[group.members[g]].result_id;
                TensorType t(indices, shape);
                std::memcpy(t.GetData().data(), all.data() + ordinal_of[g] * elems, sizeof(scalar_t) * elems);
                results_[rid] = std::move(t);
                all_reduced = all_reduced && rid < reduce_count_;
                none_reduced = none_reduced && !(reduced_ && rid < reduce_count_);
            }
            if (none_reduced) {
                guard.reusable = true;
                continue;
            }
            // ---- this group's share of the reduction: the device's FP64 sum when every member takes part --------
            std::vector<double> part(2 * elems, 0.0);
            if (all_reduced) {
                JET_JB_CHECK(jb_multi_result(guard.m, part.data()));
            }
            else {
                for (size_t g = 0; g < group.members.size(); g++) {
                    if (networks_[group.members[g]].result_id >= reduce_count_)
                        continue;
                    for (size_t i = 0; i < elems; i++) {
                        part[2 * i] += static_cast<double>(all[ordinal_of[g] * elems + i].real());
                        part[2 * i + 1] += static_cast<double>(all[ordinal_of[g] * elems + i].imag());
                    }
                }
            }
            guard.reusable = true; // every call into the plan set succeeded
            if (total.empty()) {
                total = std::move(part);
                total_indices = indices;
                total_shape = shape;
                continue;
            }
            JET_ABORT_IF_NOT(total.size() == part.size() &&
                                 Utilities::VectorDisjunctiveUnion(indices, total_indices).empty(),
                             "Tensor addition with disjoint indices is not supported.");
            if (indices == total_indices) {
                for (size_t i = 0; i < total.size(); i++)
                    total[i] += part[i];
                continue;
            }
            // align to the first result's index order (AddTensors semantics, Tensor.hpp:200-215)
            const size_t rk = indices.size();
            std::vector<size_t> src_stride(rk, 1), perm(rk);
            for (size_t j = rk; j-- > 1;)
                src_stride[j - 1] = src_stride[j] * shape[j];
            for (size_t j = 0; j < rk; j++)
                perm[j] = static_cast<size_t>(std::find(indices.begin(), indices.end(), total_indices[j]) - indices.begin());
            std::vector<size_t> counter(rk, 0);
            for (size_t i = 0; i < elems; i++) {
                size_t src = 0;
                for (size_t j = 0; j < rk; j++)
                    src += counter[j] * src_stride[perm[j]];
                total[2 * i] += part[2 * src];
                total[2 * i + 1] += part[2 * src + 1];
                for (size_t j = rk; j-- > 0;) {
                    if (++counter[j] < total_shape[j])
                        break;
                    counter[j] = 0;
                }
            }
        }
        if (reduced_ && !total.empty()) {
            TensorType out(total_indices, total_shape);
            for (size_t i = 0; i < out.GetSize(); i++)
                out[i] = scalar_t{static_cast<R>(total[2 * i]), static_cast<R>(total[2 * i + 1])};
            reduction_result_ = std::move(out);
        }
    }

    // ---- one fused GPU contraction per task (jb_contract), every named tensor downloaded ------------------
    void RunStepwise_(bool write_results)
    {
        constexpr int dtype = TensorHelpers::DtypeCode<scalar_t>();
        constexpr size_t eb = sizeof(scalar_t);

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
        if (write_results)
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
        if (acc_elems > 0 && write_results)
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
