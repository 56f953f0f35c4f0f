// Jet::SlicedContractor<Tensor> — the direct form of the reference's sliced benchmarks
// (/root/reference/examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp:53-93): instead of
// making 2^s sliced copies of the network on the host and adding each to a TaskBasedContractor,
// hand the UNSLICED network, the path and the list of sliced indices to device-resident plans
// (jb_multi_*, include/jetb200.h).  Slices are selected on the device, slice-independent steps run
// once, the per-slice steps replay as a CUDA graph, and the sum over slices is accumulated on the
// devices in double precision.  `lanes` > 1 keeps that many slices in flight per GPU (one arena, stream
// and CUDA graph per lane; pays off when one slice cannot fill the GPU) — the stream analogue of the
// reference running tasks of different slices on Taskflow workers (TaskBasedContractor.hpp:322) — and
// `devices` spreads the slice range over several GPUs of this process; the partial sums are added on the
// devices (NVLink peer copies) in (device, lane) order.
// This class is an addition to the Jet API (the reference has no equivalent); TaskBasedContractor keeps the
// reference's interface and lowers the reference's 2^s-copies flow onto the same engine.
#pragma once

#include <algorithm>
#include <complex>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "Abort.hpp"
#include "PathInfo.hpp"
#include "PlanCache.hpp"
#include "Tensor.hpp"
#include "TensorNetwork.hpp"
#include "jetb200.h"

namespace Jet {

template <class TensorType> class SlicedContractor {
  public:
    using scalar_t = typename TensorType::scalar_type_t;

    /// One device (`device`), `lanes` slices in flight (0 = chosen from the plan's memory footprint).
    SlicedContractor(const TensorNetwork<TensorType> &tn, const PathInfo::Path &path,
                     const std::vector<std::string> &sliced_indices, int device = 0, int flags = 0, int lanes = 1)
    {
        Init_(tn, path, sliced_indices, {device}, flags, lanes);
    }

    /// Several devices of this process: the slice range is dealt to them in contiguous blocks.
    SlicedContractor(const TensorNetwork<TensorType> &tn, const PathInfo::Path &path,
                     const std::vector<std::string> &sliced_indices, const std::vector<int> &devices, int flags = 0,
                     int lanes = 1)
    {
        JET_ABORT_IF(devices.empty(), "SlicedContractor: no device given.");
        Init_(tn, path, sliced_indices, devices, flags, lanes);
    }

    SlicedContractor(const SlicedContractor &) = delete;
    SlicedContractor &operator=(const SlicedContractor &) = delete;
    ~SlicedContractor() { jb_multi_destroy(multi_); }

    size_t NumSlices() const noexcept { return static_cast<size_t>(stats_.num_slices); }
    double GetFlops() const noexcept { return stats_.jet_flops_per_slice; } // PathInfo convention
    const jb_plan_stats_t &GetStats() const noexcept { return stats_; }
    int NumDevices() const noexcept { return num_devices_; }
    int NumLanes() const noexcept { return lanes_; }

    /// Contracts slices [first, first + count) and returns their sum (all slices by default).
    TensorType Contract(size_t first = 0, size_t count = static_cast<size_t>(-1))
    {
        if (count == static_cast<size_t>(-1))
            count = NumSlices() - first;
        // plan (device d, lane l) takes the (d * lanes + l)-th contiguous block of the range; everything is
        // enqueued before anything is read back, and the partial sums meet on the first device
        JET_JB_CHECK(jb_multi_run(multi_, static_cast<int64_t>(first), static_cast<int64_t>(count)));
        std::vector<double> acc(2 * static_cast<size_t>(stats_.result_elems), 0.0);
        JET_JB_CHECK(jb_multi_result(multi_, acc.data()));
        std::vector<std::string> indices;
        std::vector<size_t> shape;
        for (int i = 0; i < stats_.result_rank; i++) {
            indices.push_back(names_[static_cast<size_t>(stats_.result_modes[i])]);
            shape.push_back(static_cast<size_t>(stats_.result_extent[i]));
        }
        TensorType out(indices, shape);
        using R = typename scalar_t::value_type;
        for (size_t i = 0; i < out.GetSize(); i++)
            out[i] = scalar_t{static_cast<R>(acc[2 * i]), static_cast<R>(acc[2 * i + 1])};
        return out;
    }

    /// Device time of the last Contract() in milliseconds (CUDA events on the plans' streams; the plans run
    /// concurrently: the longest one).
    float LastMilliseconds()
    {
        float ms = 0;
        JET_JB_CHECK(jb_multi_last_ms(multi_, &ms));
        return ms;
    }

  private:
    jb_multi *multi_ = nullptr;
    jb_plan_stats_t stats_{};
    std::vector<std::string> names_;
    int num_devices_ = 1, lanes_ = 1;

    void Init_(const TensorNetwork<TensorType> &tn, const PathInfo::Path &path,
               const std::vector<std::string> &sliced_indices, const std::vector<int> &devices, int flags, int lanes)
    {
        JET_ABORT_IF(lanes < 0 || lanes > 5, "SlicedContractor: lanes must be in 0..5.");
        std::unordered_map<std::string, int32_t> label;
        std::vector<int32_t> rank, mode, flat_path, sliced;
        std::vector<int64_t> extent;
        std::vector<const void *> data;
        for (const auto &node : tn.GetNodes()) {
            const auto &t = node.tensor;
            rank.push_back(static_cast<int32_t>(t.GetIndices().size()));
            for (size_t i = 0; i < t.GetIndices().size(); i++) {
                const auto it = label.emplace(t.GetIndices()[i], static_cast<int32_t>(label.size())).first;
                if (static_cast<size_t>(it->second) == names_.size())
                    names_.push_back(t.GetIndices()[i]);
                mode.push_back(it->second);
                extent.push_back(static_cast<int64_t>(t.GetShape()[i]));
            }
            data.push_back(t.GetData().data());
        }
        for (const auto &[a, b] : path) {
            flat_path.push_back(static_cast<int32_t>(a));
            flat_path.push_back(static_cast<int32_t>(b));
        }
        for (const auto &s : sliced_indices) {
            const auto it = label.find(s);
            JET_ABORT_IF(it == label.end(), "Sliced index does not exist.");
            sliced.push_back(it->second);
        }
        jb_network_desc_t d{};
        d.dtype = TensorHelpers::DtypeCode<scalar_t>();
        d.device = devices[0];
        d.num_leaves = static_cast<int32_t>(rank.size());
        d.rank = rank.data();
        d.extent = extent.data();
        d.mode = mode.data();
        d.h_data = data.data();
        d.num_steps = static_cast<int32_t>(path.size());
        d.path = flat_path.data();
        d.num_sliced = static_cast<int32_t>(sliced.size());
        d.sliced_modes = sliced.data();
        d.flags = flags;
        // jb_multi_create releases whatever it had built when a later plan fails (e.g. out of memory on lane 3)
        // plan sets a TaskBasedContractor left in the cache hold device memory and constant-bank slots (a plan that
        // finds no free slot runs without fused chains): they go before a new plan set is built
        detail::PlanCache::Get().Flush();
        JET_JB_CHECK(jb_multi_create(&d, static_cast<int>(devices.size()), devices.data(), lanes, &multi_));
        JET_JB_CHECK(jb_multi_stats(multi_, &stats_));
        JET_JB_CHECK(jb_multi_num_plans(multi_, &num_devices_, &lanes_));
    }
};

} // namespace Jet
