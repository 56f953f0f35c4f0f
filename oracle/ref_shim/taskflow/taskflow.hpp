// TEST INFRASTRUCTURE ONLY (oracle/): a stand-in for the slice of the Taskflow
// v3.1.0 API that the reference scheduler uses.
//
// Taskflow is fetched from the network by the reference's build
// (/root/reference/CMakeLists.txt:100-110) and is not available offline, so
// /root/reference/include/jet/TaskBasedContractor.hpp cannot be compiled as
// shipped.  This header supplies exactly the surface that file touches
// (TaskBasedContractor.hpp:33,44,131-134,270-277,303-310,322,392-393,420-425,
// 449-450): tf::Task{name,precede}, tf::Taskflow{emplace,reduce,composed_of,
// dump}, tf::Executor{run}.  It is an independent implementation (a mutex +
// condition-variable work queue over a dependency-counted DAG), NOT Taskflow's
// work-stealing runtime; every report that times the reference through it says
// "Taskflow stand-in".
#pragma once

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <ostream>
#include <string>
#include <thread>
#include <vector>

namespace tf {

namespace detail {
struct Node {
    std::string name;
    std::function<void()> work;
    std::vector<Node *> successors;
    size_t num_predecessors = 0;
    size_t pending = 0; // guarded by the run's mutex
};
} // namespace detail

class Task {
  public:
    Task() = default;
    explicit Task(detail::Node *node) : node_(node) {}

    Task &name(const std::string &n)
    {
        node_->name = n;
        return *this;
    }
    const std::string &name() const { return node_->name; }

    template <class... Ts> Task &precede(Ts &&...tasks)
    {
        (AddEdge_(tasks), ...);
        return *this;
    }

    bool empty() const { return node_ == nullptr; }

  private:
    void AddEdge_(const Task &t)
    {
        node_->successors.push_back(t.node_);
        t.node_->num_predecessors++;
    }
    detail::Node *node_ = nullptr;
};

class Taskflow {
  public:
    template <class Callable> Task emplace(Callable &&c)
    {
        nodes_.emplace_back(std::make_unique<detail::Node>());
        nodes_.back()->work = std::forward<Callable>(c);
        return Task(nodes_.back().get());
    }

    // Reduction of [first, last) into `init` (the real Taskflow splits the
    // range over workers; the order of summation is unspecified there too).
    template <class It, class T, class BinOp> Task reduce(It first, It last, T &init, BinOp bop)
    {
        return emplace([first, last, &init, bop]() {
            for (It it = first; it != last; ++it) {
                init = bop(init, *it);
            }
        });
    }

    // Module task: runs another taskflow to completion, serially, in a valid
    // topological order.
    Task composed_of(Taskflow &other)
    {
        return emplace([&other]() { other.RunSerial_(); });
    }

    void dump(std::ostream &os) const
    {
        os << "digraph Taskflow {\n";
        for (const auto &n : nodes_) {
            os << "\"" << n->name << "\";\n";
            for (const auto *s : n->successors) {
                os << "\"" << n->name << "\" -> \"" << s->name << "\";\n";
            }
        }
        os << "}\n";
    }

    size_t num_tasks() const { return nodes_.size(); }
    bool empty() const { return nodes_.empty(); }

  private:
    friend class Executor;
    void RunSerial_()
    {
        std::deque<detail::Node *> ready;
        for (auto &n : nodes_) {
            n->pending = n->num_predecessors;
            if (n->pending == 0)
                ready.push_back(n.get());
        }
        while (!ready.empty()) {
            auto *n = ready.front();
            ready.pop_front();
            if (n->work)
                n->work();
            for (auto *s : n->successors) {
                if (--s->pending == 0)
                    ready.push_back(s);
            }
        }
    }
    std::deque<std::unique_ptr<detail::Node>> nodes_;
};

class Executor {
  public:
    explicit Executor(size_t num_threads = std::thread::hardware_concurrency())
        : num_threads_(num_threads == 0 ? 1 : num_threads)
    {
    }
    ~Executor()
    {
        for (auto &t : managers_) {
            if (t.joinable())
                t.join();
        }
    }

    std::future<void> run(Taskflow &flow)
    {
        auto promise = std::make_shared<std::promise<void>>();
        auto future = promise->get_future();
        managers_.emplace_back([this, &flow, promise]() {
            Run_(flow);
            promise->set_value();
        });
        return future;
    }

  private:
    struct RunState {
        std::mutex m;
        std::condition_variable cv;
        std::deque<detail::Node *> ready;
        size_t remaining = 0;
    };

    void Run_(Taskflow &flow)
    {
        RunState st;
        st.remaining = flow.nodes_.size();
        for (auto &n : flow.nodes_) {
            n->pending = n->num_predecessors;
            if (n->pending == 0)
                st.ready.push_back(n.get());
        }
        if (st.remaining == 0)
            return;
        auto worker = [&st]() {
            for (;;) {
                detail::Node *n = nullptr;
                {
                    std::unique_lock<std::mutex> lk(st.m);
                    st.cv.wait(lk, [&] { return !st.ready.empty() || st.remaining == 0; });
                    if (st.remaining == 0)
                        return;
                    n = st.ready.front();
                    st.ready.pop_front();
                }
                if (n->work)
                    n->work();
                {
                    std::lock_guard<std::mutex> lk(st.m);
                    for (auto *s : n->successors) {
                        if (--s->pending == 0)
                            st.ready.push_back(s);
                    }
                    --st.remaining;
                }
                st.cv.notify_all();
            }
        };
        std::vector<std::thread> pool;
        for (size_t i = 1; i < num_threads_; ++i)
            pool.emplace_back(worker);
        worker();
        for (auto &t : pool)
            t.join();
    }

    size_t num_threads_;
    std::vector<std::thread> managers_;
};

} // namespace tf
