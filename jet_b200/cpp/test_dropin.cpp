// C++ tests of the drop-in headers (include/jet/*.hpp over libjetb200.so), restating the
// reference's known-answer tests; each group cites the reference test it mirrors.  Needs a GPU:
// run by tests/test_cpp_dropin_gpu.py on the B200 box.  Exit code = number of failed checks.
#include <cmath>
#include <complex>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "Jet.hpp"

using namespace Jet;
using c64 = std::complex<float>;
using c128 = std::complex<double>;

static int g_failed = 0, g_checked = 0;
#define CHECK(cond)                                                                               \
    do {                                                                                          \
        g_checked++;                                                                              \
        if (!(cond)) {                                                                            \
            g_failed++;                                                                           \
            std::cerr << __FILE__ << ":" << __LINE__ << ": CHECK failed: " #cond << std::endl;    \
        }                                                                                         \
    } while (0)
#define CHECK_THROWS_WITH(expr, text)                                                             \
    do {                                                                                          \
        g_checked++;                                                                              \
        bool ok__ = false;                                                                        \
        try {                                                                                     \
            (void)(expr);                                                                         \
        }                                                                                         \
        catch (const std::exception &e) {                                                         \
            ok__ = std::string(e.what()).find(text) != std::string::npos;                         \
            if (!ok__)                                                                            \
                std::cerr << "  got: " << e.what() << std::endl;                                  \
        }                                                                                         \
        if (!ok__) {                                                                              \
            g_failed++;                                                                           \
            std::cerr << __FILE__ << ":" << __LINE__ << ": expected exception containing '"       \
                      << text << "'" << std::endl;                                                \
        }                                                                                         \
    } while (0)

template <class T> Tensor<T> MakeTensor(const std::vector<std::string> &indices, const std::vector<size_t> &shape, bool twice_imag = false)
{
    Tensor<T> t(indices, shape);
    if (!shape.empty())
        for (size_t i = 0; i < t.GetSize(); i++)
            t[i] = T(static_cast<typename T::value_type>(i), twice_imag ? static_cast<typename T::value_type>(2 * i) : 0);
    return t;
}

template <class T> bool Near(T a, T b, double tol)
{
    return std::abs(std::complex<double>(a) - std::complex<double>(b)) <= tol * std::max(1.0, std::abs(std::complex<double>(b)));
}

// ---- Tensor (reference test/Test_Tensor.cpp) ----------------------------------------------------
template <class T> void TestTensor()
{
    const double tol = std::is_same_v<T, c64> ? 1e-5 : 1e-12;
    { // default + shape constructors (:57-61,119-123,150-154,192-198)
        Tensor<T> t;
        CHECK(t.GetSize() == 1 && t.GetIndices().empty() && t.GetShape().empty() && t.IsScalar());
        CHECK(t.GetData()[0] == T(0, 0));
        CHECK(t == Tensor<T>());
        Tensor<T> s({2, 3});
        CHECK((s.GetIndices() == std::vector<std::string>{"?a", "?b"}));
        CHECK(s.GetSize() == 6);
        Tensor<T> one({"x"}, {1});
        CHECK(one.IsScalar());
    }
    { // SetData size check (:175-187), SetShape only touches the shape
        Tensor<T> t({"i", "j"}, {2, 3});
        CHECK_THROWS_WITH(t.SetData({T(1, 0)}), "Size of data and tensor do not match.");
        t.SetShape({3, 2});
        CHECK((t.GetShape() == std::vector<size_t>{3, 2}) && t.GetIndexToDimension().at("i") == 2);
    }
    { // move leaves the source unequal (:75-87); copy is deep
        Tensor<T> a({"i"}, {2}, {T(1, 2), T(3, 4)});
        Tensor<T> b(a);
        CHECK(a == b);
        Tensor<T> c(std::move(a));
        CHECK(c == b && a != c);
    }
    { // FillRandom determinism (:269-302)
        Tensor<T> a({"i", "j"}, {3, 2}), b({"i", "j"}, {3, 2});
        a.FillRandom(7);
        b.FillRandom(7);
        CHECK(a == b);
        b.FillRandom(8);
        CHECK(a != b);
    }
    { // RenameIndex / Set/GetValue
        Tensor<T> t({"i", "j"}, {2, 3});
        t.RenameIndex(1, "k");
        CHECK(t.GetIndices()[1] == "k" && t.GetIndexToDimension().count("k") && !t.GetIndexToDimension().count("j"));
        t.SetValue({1, 2}, T(5, -1));
        CHECK(t.GetValue({1, 2}) == T(5, -1) && t[5] == T(5, -1));
    }
    { // Transpose (:579-655): errors, identity copy, by labels and by ordering
        Tensor<T> d;
        CHECK_THROWS_WITH(d.Transpose(std::vector<std::string>{"x"}), "Number of indices cannot be zero.");
        CHECK_THROWS_WITH(d.Transpose(std::vector<size_t>{0}), "Size of ordering must match number of tensor indices.");
        auto t = MakeTensor<T>({"a", "b", "c"}, {2, 3, 2}, true);
        CHECK(t.Transpose(std::vector<std::string>{"a", "b", "c"}) == t);
        auto u = t.Transpose(std::vector<std::string>{"c", "a", "b"});
        CHECK((u.GetShape() == std::vector<size_t>{2, 2, 3}));
        bool ok = true;
        for (size_t a = 0; a < 2; a++)
            for (size_t b = 0; b < 3; b++)
                for (size_t c = 0; c < 2; c++)
                    ok = ok && u.GetValue({c, a, b}) == t.GetValue({a, b, c});
        CHECK(ok);
        CHECK(t.Transpose(std::vector<size_t>{2, 0, 1}) == u);
        // 2x2x2x2 literal from test/Test_Permuter.cpp:182-230 ({a,b,c,d} -> {a,b,d,c})
        auto p = MakeTensor<T>({"a", "b", "c", "d"}, {2, 2, 2, 2});
        auto q = p.Transpose(std::vector<std::string>{"a", "b", "d", "c"});
        const int want[16] = {0, 2, 1, 3, 4, 6, 5, 7, 8, 10, 9, 11, 12, 14, 13, 15};
        ok = true;
        for (int i = 0; i < 16; i++)
            ok = ok && q[i] == T(want[i], 0);
        CHECK(ok);
    }
    { // Reshape (:675-699)
        auto t = MakeTensor<T>({"a", "b"}, {2, 3});
        auto r = t.Reshape({3, 2});
        CHECK((r.GetIndices() == std::vector<std::string>{"?a", "?b"}) && r.GetData() == t.GetData());
        CHECK_THROWS_WITH(t.Reshape({4, 2}), "Size is inconsistent between tensors.");
    }
    { // AddTensors (:659-668 and Tensor.hpp:415-431)
        auto a = MakeTensor<T>({"i", "j"}, {2, 3});
        auto b = MakeTensor<T>({"j", "i"}, {3, 2}, true);
        auto c = a.AddTensor(b);
        CHECK((c.GetIndices() == std::vector<std::string>{"i", "j"}));
        bool ok = true;
        for (size_t i = 0; i < 2; i++)
            for (size_t j = 0; j < 3; j++)
                ok = ok && c.GetValue({i, j}) == a.GetValue({i, j}) + b.GetValue({j, i});
        CHECK(ok);
        CHECK(Tensor<T>().AddTensor(a) == a && a.AddTensor(Tensor<T>()) == a);
        auto z = MakeTensor<T>({"i", "k"}, {2, 3});
        CHECK_THROWS_WITH(a.AddTensor(z), "Tensor addition with disjoint indices is not supported.");
    }
    { // SliceIndex (test/Test_Tensor.cpp:579-612)
        auto t = MakeTensor<T>({"a", "b", "c"}, {2, 3, 4});
        auto s = t.SliceIndex("b", 1);
        CHECK((s.GetIndices() == std::vector<std::string>{"a", "c"}) && (s.GetShape() == std::vector<size_t>{2, 4}));
        bool ok = true;
        for (size_t a = 0; a < 2; a++)
            for (size_t c = 0; c < 4; c++)
                ok = ok && s.GetValue({a, c}) == t.GetValue({a, 1, c});
        CHECK(ok);
    }
    { // Conj
        Tensor<T> t({"i"}, {2}, {T(1, 2), T(-3, -4)});
        auto c = t.Conj();
        CHECK(c[0] == T(1, -2) && c[1] == T(-3, 4));
    }
    { // ContractTensors (:399-564): GEMM (2.25,3.0), M.v, v.M, (a,b,c).(b,c,d), dot -> scalar
        Tensor<T> a({"i", "j"}, {2, 12}), b({"j", "k"}, {12, 2});
        for (size_t i = 0; i < 24; i++)
            a[i] = b[i] = T(0.5, 0.25);
        auto c = Tensor<T>::ContractTensors(a, b);
        CHECK((c.GetIndices() == std::vector<std::string>{"i", "k"}));
        bool ok = true;
        for (size_t i = 0; i < 4; i++)
            ok = ok && Near(c[i], T(2.25, 3.0), tol);
        CHECK(ok);
        auto m = MakeTensor<T>({"i", "j"}, {2, 3}, true);
        auto v = MakeTensor<T>({"j"}, {3}, true);
        auto mv = m.ContractWithTensor(v);
        CHECK((mv.GetIndices() == std::vector<std::string>{"i"}));
        T want0(0), want1(0);
        for (size_t j = 0; j < 3; j++) {
            want0 += m.GetValue({0, j}) * v[j];
            want1 += m.GetValue({1, j}) * v[j];
        }
        CHECK(Near(mv[0], want0, tol) && Near(mv[1], want1, tol));
        auto w = MakeTensor<T>({"i"}, {2}, true);
        auto wm = w.ContractWithTensor(m);
        CHECK((wm.GetIndices() == std::vector<std::string>{"j"}) && wm.GetSize() == 3);
        T wj(0);
        for (size_t i = 0; i < 2; i++)
            wj += w[i] * m.GetValue({i, 2});
        CHECK(Near(wm[2], wj, tol));
        auto t3 = MakeTensor<T>({"a", "b", "c"}, {2, 3, 5}, true);
        auto u3 = MakeTensor<T>({"c", "b", "d"}, {5, 3, 4}, true);
        auto r = t3.ContractWithTensor(u3);
        CHECK((r.GetIndices() == std::vector<std::string>{"a", "d"}) && (r.GetShape() == std::vector<size_t>{2, 4}));
        T want(0);
        for (size_t bb = 0; bb < 3; bb++)
            for (size_t cc = 0; cc < 5; cc++)
                want += t3.GetValue({1, bb, cc}) * u3.GetValue({cc, bb, 2});
        CHECK(Near(r.GetValue({1, 2}), want, tol));
        auto dot = v.ContractWithTensor(v); // unconjugated
        CHECK(dot.GetIndices().empty() && dot.GetShape().empty() && dot.IsScalar());
        T dwant(0);
        for (size_t j = 0; j < 3; j++)
            dwant += v[j] * v[j];
        CHECK(Near(dot.GetScalar(), dwant, tol));
    }
}

// ---- TensorNetwork (reference test/Test_TensorNetwork.cpp) ------------------------------------------
template <class T> void TestTensorNetwork()
{
    using TN = TensorNetwork<Tensor<T>>;
    { // names, edges (:118-181): extent-1 axes form no edge
        TN tn;
        tn.AddTensor(MakeTensor<T>({"A0", "B1"}, {2, 3}), {"x"});
        tn.AddTensor(MakeTensor<T>({"B1", "C2"}, {3, 1}), {"y"});
        tn.AddTensor(Tensor<T>(), {});
        CHECK(tn.GetNodes()[0].name == "A0B1" && tn.GetNodes()[2].name == "_");
        CHECK(tn.NumTensors() == 3 && tn.NumIndices() == 2);
        CHECK(tn.GetIndexToEdgeMap().count("C2") == 0);
        CHECK((tn.GetIndexToEdgeMap().at("B1").node_ids == std::vector<size_t>{0, 1}));
        CHECK(tn.GetTagToNodesMap().count("x") == 1);
    }
    { // Contract(path) KAT (:529-579)
        TN tn;
        tn.AddTensor(MakeTensor<T>({"A0", "B1"}, {2, 3}, true), {});
        tn.AddTensor(MakeTensor<T>({"C2", "B1"}, {2, 3}, true), {});
        tn.AddTensor(MakeTensor<T>({"C2", "D3"}, {2, 2}, true), {});
        const auto &r = tn.Contract({{1, 2}, {0, 3}});
        const std::vector<T> want = {T(-308, -56), T(-517, -94), T(-1100, -200), T(-1804, -328)};
        CHECK(r.GetData() == want);
        CHECK((r.GetIndices() == std::vector<std::string>{"A0", "D3"}));
        CHECK(tn.GetNodes().size() == 5 && tn.GetNodes()[0].contracted && !tn.GetNodes()[4].contracted);
        CHECK(tn.GetNodes()[3].name == "B1D3" && tn.GetPath().size() == 2);
    }
    { // invalid ids / empty network (:587-614)
        TN tn;
        CHECK_THROWS_WITH(tn.Contract(), "An empty tensor network cannot be contracted.");
        tn.AddTensor(MakeTensor<T>({"A0"}, {2}), {});
        tn.AddTensor(MakeTensor<T>({"A0"}, {2}), {});
        CHECK_THROWS_WITH(tn.Contract({{2, 0}}), "Node ID 1 in contraction pair is invalid.");
        CHECK_THROWS_WITH(tn.Contract({{0, 2}}), "Node ID 2 in contraction pair is invalid.");
        const auto &r = tn.Contract(); // automatic path: one shared edge
        CHECK(r.GetScalar() == T(1, 0));
    }
    { // SliceIndices (:201-342)
        TN tn;
        tn.AddTensor(MakeTensor<T>({"A0", "B1", "C2"}, {2, 3, 4}), {});
        tn.AddTensor(MakeTensor<T>({"A0", "C2"}, {2, 4}), {});
        tn.AddTensor(MakeTensor<T>({"B1", "D3"}, {3, 2}), {});
        tn.SliceIndices({"A0", "C2"}, 1 * 4 + 2);
        const auto &n0 = tn.GetNodes()[0];
        CHECK(n0.name == "A0(1)B1C2(2)");
        CHECK((n0.tensor.GetIndices() == std::vector<std::string>{"B1"}));
        const std::vector<T> want = {T(14, 0), T(18, 0), T(22, 0)};
        CHECK(n0.tensor.GetData() == want);
        CHECK(tn.GetNodes()[1].tensor.GetIndices().empty() && tn.GetNodes()[1].tensor.GetScalar() == T(6, 0));
        CHECK(tn.GetNodes()[2].name == "B1D3" && tn.GetNodes()[2].tensor.GetSize() == 6);
        CHECK(tn.GetIndexToEdgeMap().count("A0") == 0 && tn.GetIndexToEdgeMap().count("B1") == 1);
        CHECK_THROWS_WITH(tn.SliceIndices({"Z9"}, 0), "Sliced index does not exist.");
    }
}

// ---- PathInfo (reference test/Test_PathInfo.cpp) ---------------------------------------------------
void TestPathInfo()
{
    using TN = TensorNetwork<Tensor<c64>>;
    TN tn;
    tn.AddTensor(MakeTensor<c64>({"A0", "B1"}, {2, 3}), {"t0"});
    tn.AddTensor(MakeTensor<c64>({"B1", "C2"}, {3, 4}), {"t1"});
    tn.AddTensor(MakeTensor<c64>({"C2", "D3"}, {4, 5}), {"t2"});
    PathInfo pi(tn, {{0, 1}, {3, 2}});
    CHECK(pi.GetNumLeaves() == 3 && pi.GetSteps().size() == 5);
    const auto &s3 = pi.GetSteps()[3];
    CHECK((s3.id == 3 && s3.name == "A0C2" && s3.parent == 4 && s3.children == std::pair<size_t, size_t>(0, 1)));
    CHECK((s3.contracted_indices == std::vector<std::string>{"B1"}));
    CHECK((s3.tags == std::vector<std::string>{"t0", "t1"}));
    CHECK(pi.GetSteps()[0].parent == 3 && pi.GetSteps()[4].parent == PathStepInfo::MISSING_ID);
    CHECK(pi.GetSteps()[4].name == "A0D3");
    CHECK(pi.GetPathStepFlops(0) == 0 && pi.GetPathStepFlops(3) == 2 * 4 * 2 * 3 && pi.GetPathStepFlops(4) == 2 * 5 * 2 * 4);
    CHECK(pi.GetTotalFlops() == 48 + 80);
    CHECK(pi.GetPathStepMemory(3) == 8 && pi.GetTotalMemory() == 6 + 12 + 20 + 8 + 10);
    CHECK_THROWS_WITH(pi.GetPathStepFlops(9), "Step ID is invalid.");
    CHECK_THROWS_WITH(PathInfo(tn, {{0, 7}}), "Node ID 2 in contraction path pair is invalid.");
    // sliced network (:195-252): node names keep "(v)", tensor indices drop the sliced index
    TN sl = tn;
    sl.SliceIndices({"B1"}, 2);
    PathInfo ps(sl, {{0, 1}, {3, 2}});
    CHECK(ps.GetSteps()[0].name == "A0B1(2)" && ps.GetSteps()[3].name == "A0B1(2)B1(2)C2");
    CHECK((ps.GetSteps()[3].tensor_indices == std::vector<std::string>{"A0", "C2"}));
    CHECK(ps.GetSteps()[3].contracted_indices.empty());
    CHECK(ps.GetPathStepFlops(3) == 2 * 4 * 2);
}

// ---- TaskBasedContractor (reference test/Test_TaskBasedContractor.cpp) ------------------------------
void TestTaskBasedContractor()
{
    using tensor_t = Tensor<c64>;
    using TN = TensorNetwork<tensor_t>;
    auto data_of = [](const TaskBasedContractor<tensor_t> &tbc) {
        std::map<std::string, std::vector<c64>> m;
        for (const auto &[name, ptr] : tbc.GetNameToTensorMap())
            m[name] = ptr ? ptr->GetData() : std::vector<c64>{};
        return m;
    };
    { // empty network / empty path (:94-166)
        TaskBasedContractor<tensor_t> tbc;
        CHECK(tbc.AddContractionTasks(TN(), PathInfo()) == 0);
        CHECK(tbc.GetFlops() == 0 && tbc.GetMemory() == 0 && tbc.GetNameToTaskMap().empty() && tbc.GetNameToTensorMap().empty());
        tbc.Contract().wait(); // empty taskflow completes (:288-295)
        CHECK(tbc.GetResults().empty());
        CHECK(tbc.GetReductionResult() == tensor_t());
    }
    { // names, parents, counters (:168-222)
        TaskBasedContractor<tensor_t> tbc;
        TN tn;
        tn.AddTensor(MakeTensor<c64>({"A0", "C2"}, {2, 4}), {});
        tn.AddTensor(MakeTensor<c64>({"A0", "B1"}, {2, 3}), {});
        tn.AddTensor(MakeTensor<c64>({"B1", "C2"}, {3, 4}), {});
        PathInfo pi(tn, {{0, 1}, {1, 2}, {3, 4}});
        CHECK(tbc.AddContractionTasks(tn, pi) == 0);
        CHECK(tbc.GetFlops() == (3 * 4) * 4 + (2 * 4) * 6 + (2 * 3) * 8);
        CHECK(tbc.GetMemory() == (3 * 4) + (2 * 4) + (2 * 3));
        std::map<std::string, std::string> tasks;
        for (const auto &[n, t] : tbc.GetNameToTaskMap())
            tasks[n] = t.name();
        CHECK((tasks == std::map<std::string, std::string>{{"3:C2B1", "3:C2B1"}, {"4:A0C2", "4:A0C2"}, {"5:B1A0:results[0]", "5:B1A0:results[0]"}}));
        const auto m = data_of(tbc);
        CHECK(m.size() == 6 && m.at("0:A0C2").size() == 8 && m.at("3:C2B1").empty() && m.at("5:B1A0:results[0]").empty());
        const auto &parents = tbc.GetNameToParentsMap();
        CHECK((parents.at("1:A0B1") == std::unordered_set<std::string>{"3:C2B1", "4:A0C2"}));
        CHECK(parents.count("5:B1A0:results[0]") == 0);
        tbc.Contract().wait();
        // check against the serial network contraction
        TN serial = tn;
        const auto &want = serial.Contract({{0, 1}, {1, 2}, {3, 4}});
        CHECK(tbc.GetResults().size() == 1 && tbc.GetResults()[0] == want);
        CHECK(*tbc.GetNameToTensorMap().at("3:C2B1") == serial.GetNodes()[3].tensor);
    }
    { // shared contractions (:224-281)
        TaskBasedContractor<tensor_t> tbc;
        TN tn;
        tn.AddTensor(MakeTensor<c64>({"A0", "B1"}, {2, 3}), {});
        tn.AddTensor(MakeTensor<c64>({"A0"}, {2}), {});
        tn.AddTensor(MakeTensor<c64>({"B1"}, {3}), {});
        PathInfo pi(tn, {{0, 1}, {2, 3}});
        CHECK(tbc.AddContractionTasks(tn, pi) == 0);
        CHECK(tbc.AddContractionTasks(tn, pi) == 1);
        CHECK(tbc.GetFlops() == 3 * 4 + 2 * 6 && tbc.GetMemory() == 3 + 2 * 1);
        CHECK(tbc.GetNameToTaskMap().count("4:_:results[0]") && tbc.GetNameToTaskMap().count("4:_:results[1]") && tbc.GetNameToTaskMap().size() == 3);
        CHECK((tbc.GetNameToParentsMap().at("3:B1") == std::unordered_set<std::string>{"4:_:results[0]", "4:_:results[1]"}));
        CHECK(tbc.AddReductionTask() == 1);
        tbc.Contract().wait();
        CHECK(tbc.GetResults().size() == 2 && tbc.GetResults()[0] == tensor_t({}, {}, {c64(14, 0)}) && tbc.GetResults()[1] == tbc.GetResults()[0]);
        CHECK(tbc.GetReductionResult() == tensor_t({}, {}, {c64(28, 0)}));
    }
    { // Contract() results (:297-362)
        TaskBasedContractor<tensor_t> tbc;
        TN tn;
        tn.AddTensor(MakeTensor<c64>({"A0"}, {2}), {});
        tn.AddTensor(MakeTensor<c64>({"A0", "B1"}, {2, 3}), {});
        tbc.AddContractionTasks(tn, PathInfo(tn, {{0, 1}}));
        tbc.Contract().wait();
        CHECK(tbc.GetResults().size() == 1 && tbc.GetResults()[0] == tensor_t({"B1"}, {3}, {c64(3, 0), c64(4, 0), c64(5, 0)}));
    }
    { // several results + reduction (:336-362, 453-543)
        TaskBasedContractor<tensor_t> tbc;
        TN tn;
        tn.AddTensor(MakeTensor<c64>({"A0"}, {3}), {});
        tn.AddTensor(MakeTensor<c64>({"A0"}, {3}), {});
        tn.AddTensor(MakeTensor<c64>({"B1"}, {4}), {});
        tn.AddTensor(MakeTensor<c64>({"B1"}, {4}), {});
        tbc.AddContractionTasks(tn, PathInfo(tn, {{0, 1}}));
        tbc.AddContractionTasks(tn, PathInfo(tn, {{2, 3}}));
        CHECK(tbc.AddReductionTask() == 1 && tbc.AddReductionTask() == 0 && tbc.AddReductionTask() == 0);
        tbc.Contract().wait();
        CHECK(tbc.GetResults().size() == 2 && tbc.GetResults()[0] == tensor_t({}, {}, {c64(5, 0)}) && tbc.GetResults()[1] == tensor_t({}, {}, {c64(14, 0)}));
        CHECK(tbc.GetReductionResult() == tensor_t({}, {}, {c64(19, 0)}));
    }
    { // non-scalar reduction (:473-491)
        TaskBasedContractor<tensor_t> tbc;
        TN tn;
        tn.AddTensor(MakeTensor<c64>({"A0", "B1"}, {2, 3}), {});
        tn.AddTensor(MakeTensor<c64>({"A0"}, {2}), {});
        tbc.AddContractionTasks(tn, PathInfo(tn, {{0, 1}}));
        tbc.AddReductionTask();
        tbc.Contract().wait();
        CHECK(tbc.GetReductionResult() == tensor_t({"B1"}, {3}, {c64(3, 0), c64(4, 0), c64(5, 0)}));
    }
    { // deletion tasks (:365-440)
        TaskBasedContractor<tensor_t> tbc;
        CHECK(tbc.AddDeletionTasks() == 0);
        TN tn;
        tn.AddTensor(MakeTensor<c64>({"A0", "B1"}, {2, 3}), {});
        tn.AddTensor(MakeTensor<c64>({"A0"}, {2}), {});
        tn.AddTensor(MakeTensor<c64>({"B1"}, {3}), {});
        tbc.AddContractionTasks(tn, PathInfo(tn, {{0, 1}, {2, 3}}));
        CHECK(tbc.AddDeletionTasks() == 4);
        tbc.Contract().wait();
        const auto m = data_of(tbc);
        CHECK(m.at("0:A0B1").empty() && m.at("1:A0").empty() && m.at("2:B1").empty() && m.at("3:B1").empty());
        CHECK((m.at("4:_:results[0]") == std::vector<c64>{c64(14, 0)}));
        std::ostringstream os;
        os << tbc;
        CHECK(os.str().find("4:_:results[0]") != std::string::npos && os.str().find(":delete") != std::string::npos);
    }
}


// ---- TaskBasedContractor lowered onto plan sets: sliced copies, several groups, partial reductions -------------
// (the flow of examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp:69-93 on small random networks; every
// number is checked against the same tasks replayed one kernel per task, JET_B200_TBC=stepwise)
template <class T> void TestTaskBasedContractorLowering()
{
    using tensor_t = Tensor<T>;
    using TN = TensorNetwork<tensor_t>;
    const double tol = std::is_same_v<T, c64> ? 1e-5 : 1e-12;
    unsigned seed = 12345;
    auto rnd = [&seed]() {
        seed = seed * 1664525u + 1013904223u;
        return static_cast<typename T::value_type>((seed >> 8) & 0xffff) / 32768 - 1;
    };
    auto random_tensor = [&](const std::vector<std::string> &indices, const std::vector<size_t> &shape) {
        tensor_t t(indices, shape);
        for (size_t i = 0; i < t.GetSize(); i++)
            t[i] = T(rnd(), rnd());
        return t;
    };
    auto near_tensor = [&](const tensor_t &a, const tensor_t &b) {
        if (a.GetIndices() != b.GetIndices() || a.GetShape() != b.GetShape())
            return false;
        double num = 0, den = 0;
        for (size_t i = 0; i < a.GetSize(); i++) {
            num += std::norm(std::complex<double>(a.GetData()[i]) - std::complex<double>(b.GetData()[i]));
            den += std::norm(std::complex<double>(b.GetData()[i]));
        }
        return std::sqrt(num) <= tol * std::max(std::sqrt(den), 1e-30);
    };
    // a ladder: two rails of 5 rank-3/4 tensors, rungs r*, rail bonds a*/b*, open legs o0 (dim 3) and o1
    TN tn;
    const size_t L = 5;
    for (size_t i = 0; i < L; i++) {
        std::vector<std::string> ia = {"r" + std::to_string(i)}, ib = {"r" + std::to_string(i)};
        std::vector<size_t> sa = {i == 2 ? size_t(3) : size_t(2)}, sb = sa;
        if (i > 0) {
            ia.push_back("a" + std::to_string(i - 1));
            sa.push_back(2);
            ib.push_back("b" + std::to_string(i - 1));
            sb.push_back(i == 3 ? 4 : 2);
        }
        if (i + 1 < L) {
            ia.push_back("a" + std::to_string(i));
            sa.push_back(2);
            ib.push_back("b" + std::to_string(i));
            sb.push_back(i == 2 ? 4 : 2);
        }
        if (i == 0) {
            ia.push_back("o0");
            sa.push_back(3);
        }
        if (i == L - 1) {
            ib.push_back("o1");
            sb.push_back(2);
        }
        tn.AddTensor(random_tensor(ia, sa), {});
        tn.AddTensor(random_tensor(ib, sb), {});
    }
    // path: contract rung pairs, then sweep
    typename TN::Path path;
    for (size_t i = 0; i < L; i++)
        path.emplace_back(2 * i, 2 * i + 1);
    size_t acc = 2 * L;
    for (size_t i = 1; i < L; i++) {
        path.emplace_back(acc, 2 * L + i);
        acc = 3 * L + i - 1;
    }
    const std::vector<std::string> sliced = {"a1", "r2", "b2"}; // dims 2, 3, 4 -> 24 slices
    const size_t num_slices = 24;
    tensor_t full;
    {
        TN copy = tn;
        full = copy.Contract(path);
    }
    auto build = [&](TaskBasedContractor<tensor_t> &tbc, size_t count, bool reduce_first) {
        if (reduce_first)
            tbc.AddReductionTask();
        size_t shared = 0;
        for (size_t v = 0; v < count; v++) {
            TN slice = tn;
            slice.SliceIndices(sliced, v);
            shared += tbc.AddContractionTasks(slice, PathInfo(slice, path));
        }
        return shared;
    };
    { // all 24 slices: reduction == unsliced contraction; per-slice results == stepwise replay
        TaskBasedContractor<tensor_t> tbc, ref;
        const size_t shared = build(tbc, num_slices, false);
        build(ref, num_slices, false);
        CHECK(shared > 0);
        tbc.AddReductionTask();
        ref.AddReductionTask();
        tbc.Contract().get();
        setenv("JET_B200_TBC", "stepwise", 1);
        ref.Contract().get();
        unsetenv("JET_B200_TBC");
        CHECK(tbc.GetResults().size() == num_slices);
        bool all_near = true;
        for (size_t v = 0; v < num_slices; v++)
            all_near = all_near && near_tensor(tbc.GetResults()[v], ref.GetResults()[v]);
        CHECK(all_near);
        CHECK(near_tensor(tbc.GetReductionResult(), full));
        CHECK(near_tensor(ref.GetReductionResult(), full));
        // the name -> tensor map holds every intermediate of every slice once it is asked for
        const auto &lazy = tbc.GetNameToTensorMap();
        const auto &eager = ref.GetNameToTensorMap();
        CHECK(lazy.size() == eager.size());
        bool same = true;
        for (const auto &[name, ptr] : eager) {
            const auto it = lazy.find(name);
            same = same && it != lazy.end() && (ptr == nullptr) == (it->second == nullptr) &&
                   (ptr == nullptr || near_tensor(*it->second, *ptr));
        }
        CHECK(same);
    }
    { // a subset of the slices, deletion tasks, results added after the reduction task are not reduced
        TaskBasedContractor<tensor_t> tbc;
        build(tbc, 5, false);
        tbc.AddReductionTask();
        TN extra = tn;
        extra.SliceIndices(sliced, 7);
        tbc.AddContractionTasks(extra, PathInfo(extra, path));
        tbc.AddDeletionTasks();
        tbc.Contract().get();
        CHECK(tbc.GetResults().size() == 6);
        tensor_t want = tbc.GetResults()[0];
        for (size_t v = 1; v < 5; v++)
            want = want.AddTensor(tbc.GetResults()[v]);
        const double keep = tol;
        (void)keep;
        CHECK(near_tensor(tbc.GetReductionResult(), want));
        size_t alive = 0;
        for (const auto &[name, ptr] : tbc.GetNameToTensorMap())
            alive += ptr != nullptr;
        CHECK(alive == 6); // only the six results survive the deletion tasks
    }
    { // the same structure again with NEW leaf data: the second Contract() reuses the cached plan set (upload + run)
      // and must give the new network's results, not the old one's; a third contractor with the cache disabled agrees
        TN other;
        for (size_t id = 0; id < tn.NumTensors(); id++) {
            const auto &node = tn.GetNodes()[id];
            other.AddTensor(random_tensor(node.tensor.GetIndices(), node.tensor.GetShape()), {});
        }
        tensor_t other_full;
        {
            TN copy = other;
            other_full = copy.Contract(path);
        }
        auto run = [&](TN &net) {
            TaskBasedContractor<tensor_t> tbc;
            for (size_t v = 0; v < num_slices; v++) {
                TN slice = net;
                slice.SliceIndices(sliced, v);
                tbc.AddContractionTasks(slice, PathInfo(slice, path));
            }
            tbc.AddReductionTask();
            tbc.Contract().get();
            return tbc.GetReductionResult();
        };
        const tensor_t first = run(tn);      // creates (or reuses) the plan set of this structure
        const tensor_t second = run(other);  // cache hit: new leaves uploaded into the same plans
        const tensor_t third = run(tn);      // and back
        CHECK(near_tensor(first, full));
        CHECK(near_tensor(second, other_full));
        CHECK(!near_tensor(second, full));
        CHECK(near_tensor(third, full));
        bool identical = first.GetData().size() == third.GetData().size();
        for (size_t i = 0; identical && i < first.GetData().size(); i++)
            identical = first.GetData()[i] == third.GetData()[i];
        CHECK(identical); // a reused plan set is deterministic: bit-identical to the first run
        // something was cached (unless JET_B200_PLAN_CACHE=0); released here
        CHECK(Jet::detail::PlanCache::Get().Flush() == (Jet::detail::PlanCache::Get().capacity > 0));
        CHECK(near_tensor(run(other), other_full));   // rebuilt from scratch after the flush
    }
    { // two unrelated networks (two groups) whose results carry the same indices in different orders
        TN n1, n2;
        n1.AddTensor(random_tensor({"x", "k"}, {3, 4}), {});
        n1.AddTensor(random_tensor({"k", "y"}, {4, 2}), {});
        n2.AddTensor(random_tensor({"y", "q", "p"}, {2, 5, 2}), {});
        n2.AddTensor(random_tensor({"q", "x", "p"}, {5, 3, 2}), {});
        TaskBasedContractor<tensor_t> tbc;
        tbc.AddContractionTasks(n1, PathInfo(n1, {{0, 1}}));
        tbc.AddContractionTasks(n2, PathInfo(n2, {{0, 1}}));
        tbc.AddReductionTask();
        tbc.Contract().get();
        TN c1 = n1, c2 = n2;
        const tensor_t want = c1.Contract({{0, 1}}).AddTensor(c2.Contract({{0, 1}}));
        CHECK(near_tensor(tbc.GetReductionResult(), want));
        CHECK((tbc.GetResults()[1].GetIndices() == std::vector<std::string>{"y", "x"}));
    }
    { // SlicedContractor over every device of the process == over one device
        int ndev = 1;
        jb_device_count(&ndev);
        std::vector<int> devices;
        for (int d = 0; d < std::min(ndev, 4); d++)
            devices.push_back(d);
        SlicedContractor<tensor_t> one(tn, path, sliced), many(tn, path, sliced, devices, 0, 2);
        CHECK(one.NumSlices() == num_slices && many.NumDevices() == static_cast<int>(devices.size()));
        CHECK(near_tensor(one.Contract(), full));
        CHECK(near_tensor(many.Contract(), full));
        CHECK(near_tensor(many.Contract(3, 11), one.Contract(3, 11)));
    }
}

// ---- Permuter<Backend> and TensorHelpers::MultiplyTensorData (reference test/Test_Permuter.cpp, Test_TensorHelpers.cpp) ----
template <class T> void TestPermuterAndHelpers()
{
    { // Test_Permuter.cpp: 2-D and 3-D transposes through both back ends, argument checks with the reference's messages
        const std::vector<T> in = {T(0, 0), T(1, 0), T(2, 0), T(3, 0), T(4, 0), T(5, 0)};
        Permuter<DefaultPermuter<>> pd;
        Permuter<QFlexPermuter<>> pq;
        const std::vector<T> want = {T(0, 0), T(3, 0), T(1, 0), T(4, 0), T(2, 0), T(5, 0)};
        CHECK(pd.Transpose(in, {2, 3}, {"a", "b"}, {"b", "a"}) == want);
        CHECK(pq.Transpose(in, {2, 3}, {"a", "b"}, {"b", "a"}) == want);
        std::vector<T> out(6);
        pq.Transpose(in, {2, 3}, out, {"a", "b"}, {"b", "a"});
        CHECK(out == want);
        std::vector<T> big(2 * 3 * 4 * 5);
        for (size_t i = 0; i < big.size(); i++)
            big[i] = T(static_cast<typename T::value_type>(i), -static_cast<typename T::value_type>(i));
        const auto r = pd.Transpose(big, {2, 3, 4, 5}, {"a", "b", "c", "d"}, {"d", "b", "a", "c"});
        bool ok = true;
        for (size_t a = 0; a < 2; a++)
            for (size_t b = 0; b < 3; b++)
                for (size_t c = 0; c < 4; c++)
                    for (size_t d = 0; d < 5; d++)
                        ok = ok && r[((d * 3 + b) * 2 + a) * 4 + c] == big[((a * 3 + b) * 4 + c) * 5 + d];
        CHECK(ok);
        CHECK_THROWS_WITH(pd.Transpose(in, {2, 3}, {"a", "a"}, {"b", "a"}), "Duplicate existing indices found.");
        CHECK_THROWS_WITH(pd.Transpose(in, {2, 3}, {"a", "b"}, {"b", "b"}), "Duplicate transpose indices found.");
        CHECK_THROWS_WITH(pd.Transpose(in, {2, 3}, {"a", "b"}, {"b", "c"}), "New indices are an invalid permutation");
        CHECK_THROWS_WITH(pd.Transpose(in, {2, 2}, {"a", "b"}, {"b", "a"}), "Tensor shape does not match given input tensor data.");
        CHECK_THROWS_WITH(pd.Transpose(in, {2, 3}, {"a", "b"}, {"b", "a", "c"}), "Tensor shape does not match number of indices.");
    }
    { // TensorHelpers.hpp:131-168: GEMM, GEMV, transposed GEMV and DOTU through the one product
        const std::vector<T> A = {T(1, 0), T(2, 0), T(3, 0), T(4, 0), T(0, 1), T(1, 1)}; // 2 x 3
        const std::vector<T> B = {T(1, 0), T(0, 1), T(2, 0), T(1, 0), T(0, 0), T(1, -1)}; // 3 x 2
        std::vector<T> C(4);
        TensorHelpers::MultiplyTensorData<T>(A, B, C, {"i"}, {"j"}, 2, 2, 3);
        CHECK(Near(C[0], T(5, 0), 1e-6) && Near(C[1], T(5, -2), 1e-6) && Near(C[2], T(4, 2), 1e-6) && Near(C[3], T(2, 5), 1e-6));
        const std::vector<T> v = {T(1, 0), T(0, 1), T(2, 0)};
        std::vector<T> y(2);
        TensorHelpers::MultiplyTensorData<T>(A, v, y, {"i"}, {}, 2, 1, 3); // A v
        CHECK(Near(y[0], T(7, 2), 1e-6) && Near(y[1], T(5, 2), 1e-6));
        std::vector<T> z(2);
        TensorHelpers::MultiplyTensorData<T>(v, B, z, {}, {"j"}, 1, 2, 3); // v^T B
        CHECK(Near(z[0], T(1, 2), 1e-6) && Near(z[1], T(2, 0), 1e-6));
        std::vector<T> s(1);
        TensorHelpers::MultiplyTensorData<T>(v, v, s, {}, {}, 1, 1, 3); // unconjugated dot
        CHECK(Near(s[0], T(4, 0), 1e-6));
    }
}

// ---- TensorNetworkSerializer (reference test/Test_TensorNetworkIO.cpp) ------------------------------
template <class T> void TestIO()
{
    using tensor_t = Tensor<T>;
    TensorNetworkSerializer<tensor_t> ser;
    { // error classes (:12-59)
        bool threw = false;
        try {
            ser("");
        }
        catch (const JsonException &) {
            threw = true;
        }
        CHECK(threw);
        CHECK_THROWS_WITH(ser("[]"), "Error parsing tensor network file");
        CHECK_THROWS_WITH(ser("{}"), "Error parsing tensor network file");
        CHECK_THROWS_WITH(ser(R"({"path": [[0,1]]})"), "tensors");
        CHECK_THROWS_WITH(ser(R"({"tensors": [[["I0"], ["a"], [2], [[1.0], [0.0,0.0]]]]})"), "[1.0]");
    }
    { // round trip byte for byte with indent -1 (:61-127)
        const std::string text =
            R"({"path":[[0,2],[1,3]],"tensors":[[["A","hermitian"],["a","b"],[2,2],[[1.0,0.0],[0.0,1.0],[0.0,-1.0],[1.0,0.0]]],[["B"],["b","c"],[2,1],[[0.5,0.25],[1e-09,2.0]]],[[],["a"],[2],[[3.0,0.0],[0.0,0.0]]]]})";
        auto file = ser(text);
        CHECK(file.tensors.NumTensors() == 3 && file.path.has_value() && file.path->GetPath().size() == 2);
        CHECK((file.tensors.GetNodes()[0].tags == std::vector<std::string>{"A", "hermitian"}));
        CHECK(file.tensors.GetNodes()[1].tensor.GetValue({1, 0}) == T(static_cast<typename T::value_type>(1e-09), 2.0));
        if (std::is_same_v<T, c128>)
            CHECK(ser(file.tensors, *file.path) == text);
        auto again = ser(ser(file.tensors, *file.path));
        CHECK(again.tensors.GetNodes()[1].tensor == file.tensors.GetNodes()[1].tensor);
        CHECK(ser(file.tensors).find("\"path\"") == std::string::npos);
        // column-major load reverses indices and shapes
        auto cm = ser(text, true);
        CHECK((cm.tensors.GetNodes()[1].tensor.GetIndices() == std::vector<std::string>{"c", "b"}));
        CHECK((cm.tensors.GetNodes()[1].tensor.GetShape() == std::vector<size_t>{1, 2}));
        const std::string pretty = TensorNetworkSerializer<tensor_t>(2)(file.tensors);
        CHECK(pretty.find("\n  \"tensors\": [") != std::string::npos);
    }
}

// ---- SlicedContractor vs TaskBasedContractor vs serial contraction on a real file -------------------
void TestSlicedFile(const std::string &file_name)
{
    using tensor_t = Tensor<c64>;
    std::ifstream in(file_name);
    if (!in) {
        std::cerr << "skipping file test: cannot open " << file_name << std::endl;
        return;
    }
    std::string text{std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()};
    auto file = TensorNetworkSerializer<tensor_t>()(text);
    const auto path = file.path.value().GetPath();
    const std::vector<std::string> sliced = {"p7", "s7", "h4", "m1", "m2", "I2"};
    // the reference's sliced driver (examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp:69-93)
    // on the first 4 of the 64 slices, through the drop-in TaskBasedContractor
    TaskBasedContractor<tensor_t> tbc;
    size_t shared_total = 0;
    for (size_t v = 0; v < 4; v++) {
        auto slice = file.tensors;
        slice.SliceIndices(sliced, v);
        PathInfo pi(slice, path);
        shared_total += tbc.AddContractionTasks(slice, pi);
    }
    tbc.AddReductionTask();
    tbc.AddDeletionTasks();
    tbc.Contract().wait();
    const c64 via_tbc = tbc.GetReductionResult().GetScalar();
    CHECK(shared_total >= 3 * 244); // >= 244 slice-independent steps per extra slice
    SlicedContractor<tensor_t> sc(file.tensors, path, sliced);
    CHECK(sc.NumSlices() == 64);
    const c64 via_plan = sc.Contract(0, 4).GetScalar();
    // golden: reference slices 0..3 summed (tests/golden/amplitudes.json)
    const std::complex<double> want(1.6379545497713366e-09 - 3.4047192842834306e-10 - 1.5378681661459837e-09 + 1.4890139121703783e-09,
                                    -1.3602353965413982e-10 + 2.735360549177557e-10 + 3.457236996684543e-10 - 1.6441056294169698e-09);
    CHECK(std::abs(std::complex<double>(via_tbc) - want) / std::abs(want) < 1e-5);
    CHECK(std::abs(std::complex<double>(via_plan) - want) / std::abs(want) < 1e-5);
    std::cout << "m10 slices 0..3: tbc " << via_tbc << " plan " << via_plan << " want " << want << std::endl;
    // three slices in flight: same sum (FP64 accumulators per lane, added in lane order)
    SlicedContractor<tensor_t> sc3(file.tensors, path, sliced, 0, 0, 3);
    const c64 via_lanes = sc3.Contract(0, 4).GetScalar();
    CHECK(std::abs(std::complex<double>(via_lanes) - want) / std::abs(want) < 1e-5);
    const std::complex<double> full3(sc3.Contract().GetScalar()), full(sc.Contract().GetScalar());
    std::cout << "m10 all 64 slices: " << sc3.NumLanes() << " lanes " << full3 << ", " << sc.NumLanes() << " lanes " << full << std::endl;
    CHECK(std::abs(full3 - full) < 1e-6 * std::abs(full));
}

int main(int argc, char **argv)
{
    try {
        TestTensor<c64>();
        TestTensor<c128>();
        TestTensorNetwork<c64>();
        TestTensorNetwork<c128>();
        TestPathInfo();
        TestPermuterAndHelpers<c64>();
        TestPermuterAndHelpers<c128>();
        TestTaskBasedContractor();
        TestTaskBasedContractorLowering<c64>();
        TestTaskBasedContractorLowering<c128>();
        TestIO<c64>();
        TestIO<c128>();
        if (argc > 1)
            TestSlicedFile(argv[1]);
    }
    catch (const std::exception &e) {
        std::cerr << "unexpected exception: " << e.what() << std::endl;
        g_failed++;
    }
    std::cout << "version " << Jet::Version() << ": " << g_checked << " checks, " << g_failed << " failed" << std::endl;
    return g_failed > 100 ? 100 : g_failed;
}
