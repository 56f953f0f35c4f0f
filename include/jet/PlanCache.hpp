#pragma once
/**
 * @file PlanCache.hpp
 * jet-b200 addition (no counterpart in the reference): the process-wide cache of lowered plan sets shared by
 * `TaskBasedContractor` (include/jet/TaskBasedContractor.hpp) and released on demand by `SlicedContractor`.
 */
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "jetb200.h"

namespace Jet {
namespace detail {
/**
 * Process-wide cache of lowered plan sets (jb_multi), keyed by everything a plan is built from except the leaf
 * data: dtype, devices, lanes, leaf ranks / extents / modes, path, sliced modes.  A program that contracts the same
 * circuit again and again (other output bitstrings, other parameters: the same structure with new tensors) pays
 * for planning, arena allocation and graph capture once; later `Contract()` calls upload the new leaves into the
 * cached plans and run.  An entry is checked OUT while a contractor uses it, so two threads never share one;
 * at most `JET_B200_PLAN_CACHE` entries (default 2; 0 disables), least recently used first out, and only plan
 * sets whose arenas total at most `JET_B200_PLAN_CACHE_MIB` (default 4096) — larger ones are cheap to rebuild
 * relative to their run time and would pin device memory.  Idle entries are dropped whenever a plan set of another
 * structure is created (`Flush`): they would otherwise keep constant-bank slots and device memory from it.  The cache
 * is never destroyed at exit (the CUDA context may already be gone by then).
 */
struct PlanCache {
    struct Entry {
        std::string key;
        jb_multi *m;
    };
    std::mutex mutex;
    std::vector<Entry> idle; // most recently used last
    size_t capacity = 2;
    size_t max_bytes = size_t(4096) << 20;

    PlanCache()
    {
        if (const char *e = std::getenv("JET_B200_PLAN_CACHE"))
            capacity = static_cast<size_t>(std::max(0, std::atoi(e)));
        if (const char *e = std::getenv("JET_B200_PLAN_CACHE_MIB"))
            max_bytes = static_cast<size_t>(std::max(0, std::atoi(e))) << 20;
    }
    jb_multi *CheckOut(const std::string &key)
    {
        std::lock_guard<std::mutex> lock(mutex);
        for (size_t i = idle.size(); i-- > 0;)
            if (idle[i].key == key) {
                jb_multi *m = idle[i].m;
                idle.erase(idle.begin() + static_cast<std::ptrdiff_t>(i));
                return m;
            }
        return nullptr;
    }
    /// Takes ownership of `m`: kept for the next contractor, or destroyed.
    void CheckIn(std::string key, jb_multi *m, size_t bytes)
    {
        std::vector<jb_multi *> drop;
        {
            std::lock_guard<std::mutex> lock(mutex);
            if (capacity == 0 || bytes > max_bytes)
                drop.push_back(m);
            else {
                idle.push_back({std::move(key), m});
                while (idle.size() > capacity) {
                    drop.push_back(idle.front().m);
                    idle.erase(idle.begin());
                }
            }
        }
        for (jb_multi *d : drop)
            jb_multi_destroy(d);
    }
    /// Destroys every idle entry (device memory or constant-bank slots are needed for a new plan set).
    bool Flush()
    {
        std::vector<Entry> drop;
        {
            std::lock_guard<std::mutex> lock(mutex);
            drop.swap(idle);
        }
        for (Entry &e : drop)
            jb_multi_destroy(e.m);
        return !drop.empty();
    }
    static PlanCache &Get()
    {
        static PlanCache *cache = new PlanCache; // intentionally leaked, see above
        return *cache;
    }
};
} // namespace detail

} // namespace Jet
