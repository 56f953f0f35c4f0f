// pybind11 module `jet_b200.bindings`: the same Python names the reference's `jet.bindings`
// exports (/root/reference/python/src/Python.cpp:12-34 and python/src/{Tensor,TensorNetwork,
// PathInfo,TaskBasedContractor,TensorNetworkIO}.hpp), bound to the B200 drop-in headers.
// Additions: SlicedContractorC64/C128 (device-resident sliced contraction) and NumPy-array access
// to tensor data (`Tensor.array`), neither of which exists in the reference.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "Jet.hpp"

namespace py = pybind11;

namespace {

template <class T> void BindTensor(py::module_ &m, const char *name)
{
    using tensor_t = Jet::Tensor<T>;
    py::class_<tensor_t>(m, name, "Tensor of complex values with labelled indices (GPU-backed operators).")
        .def_property_readonly_static("dtype", [](const py::object &) { return py::dtype::of<T>(); })
        .def(py::init<>())
        .def(py::init<const std::vector<size_t> &>(), py::arg("shape"))
        .def(py::init<const std::vector<std::string> &, const std::vector<size_t> &>(), py::arg("indices"),
             py::arg("shape"))
        .def(py::init<const std::vector<std::string> &, const std::vector<size_t> &, const std::vector<T> &>(),
             py::arg("indices"), py::arg("shape"), py::arg("data"))
        .def(py::init<const tensor_t &>(), py::arg("other"))
        .def_property("data", [](const tensor_t &t) { return t.GetData(); }, &tensor_t::SetData)
        .def_property("shape", &tensor_t::GetShape, &tensor_t::SetShape)
        .def_property_readonly("index_to_dimension_map", &tensor_t::GetIndexToDimension)
        .def_property_readonly("indices", &tensor_t::GetIndices)
        .def_property_readonly("scalar", &tensor_t::GetScalar)
        .def_property_readonly(
            "array",
            [](const tensor_t &t) {
                std::vector<py::ssize_t> shape(t.GetShape().begin(), t.GetShape().end());
                py::array_t<T> arr(shape);
                std::copy(t.GetData().begin(), t.GetData().end(), arr.mutable_data());
                return arr;
            },
            "Copy of the data as a NumPy array shaped like the tensor.")
        .def("__getitem__",
             [](const tensor_t &t, size_t pos) {
                 if (pos >= t.GetSize())
                     throw py::index_error("Tensor index out of range.");
                 return t[pos];
             })
        .def("__len__", &tensor_t::GetSize)
        .def("__repr__",
             [](const tensor_t &t) {
                 std::ostringstream os;
                 os << t;
                 return os.str();
             })
        .def(py::self == py::self, py::arg("other"))
        .def(py::self != py::self, py::arg("other"))
        .def("fill_random", py::overload_cast<>(&tensor_t::FillRandom))
        .def("fill_random", py::overload_cast<size_t>(&tensor_t::FillRandom), py::arg("seed"))
        .def("init_indices_and_shape", &tensor_t::InitIndicesAndShape, py::arg("indices"), py::arg("shape"))
        .def("get_value", &tensor_t::GetValue, py::arg("indices"))
        .def("is_scalar", &tensor_t::IsScalar)
        .def("rename_index", &tensor_t::RenameIndex, py::arg("pos"), py::arg("new_label"))
        .def("set_value", &tensor_t::SetValue, py::arg("indices"), py::arg("value"))
        .def("add_tensor", &tensor_t::AddTensor, py::arg("other"))
        .def("conj", [](const tensor_t &t) { return t.Conj(); })
        .def("contract_with_tensor", &tensor_t::ContractWithTensor, py::arg("other"))
        .def("reshape", [](const tensor_t &t, const std::vector<size_t> &shape) { return t.Reshape(shape); },
             py::arg("shape"))
        .def("slice_index", [](const tensor_t &t, const std::string &index, size_t value) {
            return t.SliceIndex(index, value);
        }, py::arg("index"), py::arg("value"))
        .def("transpose", [](const tensor_t &t, const std::vector<std::string> &indices) {
            return t.Transpose(indices);
        }, py::arg("indices"))
        .def("transpose", [](const tensor_t &t, const std::vector<size_t> &ordering) {
            return t.Transpose(ordering);
        }, py::arg("ordering"));

    m.def("add_tensors", [](const tensor_t &a, const tensor_t &b) { return tensor_t::AddTensors(a, b); },
          py::arg("A"), py::arg("B"));
    m.def("conj", [](const tensor_t &a) { return tensor_t::Conj(a); }, py::arg("A"));
    m.def("contract_tensors", [](const tensor_t &a, const tensor_t &b) { return tensor_t::ContractTensors(a, b); },
          py::arg("A"), py::arg("B"));
    m.def("reshape", [](const tensor_t &a, const std::vector<size_t> &shape) { return tensor_t::Reshape(a, shape); },
          py::arg("tensor"), py::arg("shape"));
    m.def("slice_index", [](const tensor_t &a, const std::string &index, size_t value) {
        return tensor_t::SliceIndex(a, index, value);
    }, py::arg("tensor"), py::arg("index"), py::arg("value"));
    m.def("transpose", [](const tensor_t &a, const std::vector<std::string> &indices) {
        return tensor_t::Transpose(a, indices);
    }, py::arg("tensor"), py::arg("indices"));
    m.def("transpose", [](const tensor_t &a, const std::vector<size_t> &ordering) {
        return tensor_t::Transpose(a, ordering);
    }, py::arg("tensor"), py::arg("ordering"));
}

template <class T> void BindTensorNetwork(py::module_ &m, const char *name)
{
    using tensor_t = Jet::Tensor<T>;
    using TN = Jet::TensorNetwork<tensor_t>;
    using Node = typename TN::Node;
    using Edge = typename TN::Edge;
    auto cls = py::class_<TN>(m, name, "Tensor network: nodes (tensors) joined by shared indices.")
        .def_property_readonly_static("dtype", [](const py::object &) { return py::dtype::of<T>(); })
        .def(py::init<>())
        .def("__str__", [](const TN &tn) {
            std::ostringstream os;
            os << tn;
            return os.str();
        })
        .def_property_readonly("index_to_edge_map", &TN::GetIndexToEdgeMap)
        .def_property_readonly("tag_to_node_id_map", [](const TN &tn) {
            std::unordered_map<std::string, std::vector<size_t>> out;
            for (const auto &[tag, id] : tn.GetTagToNodesMap())
                out[tag].push_back(id);
            return out;
        })
        .def_property_readonly("path", [](TN &tn) { return tn.GetPath(); })
        .def_property_readonly("nodes", &TN::GetNodes)
        .def_property_readonly("num_tensors", &TN::NumTensors)
        .def_property_readonly("num_indices", &TN::NumIndices)
        .def("add_tensor", &TN::AddTensor, py::arg("tensor"), py::arg("tags") = std::vector<std::string>())
        .def("slice_indices", &TN::SliceIndices, py::arg("indices"), py::arg("value"))
        .def("contract", [](TN &tn, const typename TN::Path &path) { return tn.Contract(path); },
             py::arg("path") = typename TN::Path());
    py::class_<Node>(cls, (std::string(name) + "Node").c_str())
        .def_readonly("id", &Node::id)
        .def_readonly("name", &Node::name)
        .def_readonly("indices", &Node::indices)
        .def_readonly("tags", &Node::tags)
        .def_readonly("contracted", &Node::contracted)
        .def_readonly("tensor", &Node::tensor);
    py::class_<Edge>(cls, (std::string(name) + "Edge").c_str())
        .def_readonly("dim", &Edge::dim)
        .def_readonly("node_ids", &Edge::node_ids)
        .def("__eq__", &Edge::operator==);
}

void BindPathInfo(py::module_ &m)
{
    py::class_<Jet::PathStepInfo>(m, "PathStepInfo")
        .def_property_readonly_static("MISSING_ID", [](const py::object &) { return Jet::PathStepInfo::MISSING_ID; })
        .def_readonly("id", &Jet::PathStepInfo::id)
        .def_readonly("parent", &Jet::PathStepInfo::parent)
        .def_readonly("children", &Jet::PathStepInfo::children)
        .def_readonly("name", &Jet::PathStepInfo::name)
        .def_readonly("node_indices", &Jet::PathStepInfo::node_indices)
        .def_readonly("tensor_indices", &Jet::PathStepInfo::tensor_indices)
        .def_readonly("tags", &Jet::PathStepInfo::tags)
        .def_readonly("contracted_indices", &Jet::PathStepInfo::contracted_indices);
    py::class_<Jet::PathInfo>(m, "PathInfo", "Symbolic replay of a contraction path.")
        .def(py::init<>())
        .def(py::init<const Jet::TensorNetwork<Jet::Tensor<std::complex<float>>> &, const Jet::PathInfo::Path &>(),
             py::arg("tn"), py::arg("path"))
        .def(py::init<const Jet::TensorNetwork<Jet::Tensor<std::complex<double>>> &, const Jet::PathInfo::Path &>(),
             py::arg("tn"), py::arg("path"))
        .def_property_readonly("index_to_size_map", &Jet::PathInfo::GetIndexSizes)
        .def_property_readonly("num_leaves", &Jet::PathInfo::GetNumLeaves)
        .def_property_readonly("path", &Jet::PathInfo::GetPath)
        .def_property_readonly("steps", &Jet::PathInfo::GetSteps)
        .def("total_flops", &Jet::PathInfo::GetTotalFlops)
        .def("total_memory", &Jet::PathInfo::GetTotalMemory);
}

template <class T> void BindContractors(py::module_ &m, const char *tbc_name, const char *sliced_name)
{
    using tensor_t = Jet::Tensor<T>;
    using TBC = Jet::TaskBasedContractor<tensor_t>;
    py::class_<TBC>(m, tbc_name, "Task-based contractor executing on the GPU.")
        .def_property_readonly_static("dtype", [](const py::object &) { return py::dtype::of<T>(); })
        .def(py::init<>())
        .def_property_readonly("name_to_tensor_map", [](const TBC &tbc) {
            std::unordered_map<std::string, tensor_t *> out;
            for (const auto &[name, ptr] : tbc.GetNameToTensorMap())
                out.emplace(name, ptr.get());
            return out;
        }, py::return_value_policy::reference_internal)
        .def_property_readonly("name_to_parents_map", &TBC::GetNameToParentsMap)
        .def_property_readonly("results", &TBC::GetResults)
        .def_property_readonly("reduction_result", &TBC::GetReductionResult)
        .def_property_readonly("flops", &TBC::GetFlops)
        .def_property_readonly("memory", &TBC::GetMemory)
        .def("add_contraction_tasks", &TBC::AddContractionTasks, py::arg("tn"), py::arg("path_info"))
        .def("add_reduction_task", &TBC::AddReductionTask)
        .def("add_deletion_tasks", &TBC::AddDeletionTasks)
        .def("contract", [](TBC &tbc) {
            py::gil_scoped_release release;
            tbc.Contract().get(); // get(): rethrow GPU errors instead of swallowing them
        });

    using SC = Jet::SlicedContractor<tensor_t>;
    py::class_<SC>(m, sliced_name, "Device-resident sliced contraction (B200 extension).")
        .def(py::init<const Jet::TensorNetwork<tensor_t> &, const Jet::PathInfo::Path &,
                      const std::vector<std::string> &, int, int, int>(),
             py::arg("tn"), py::arg("path"), py::arg("sliced_indices"), py::arg("device") = 0, py::arg("flags") = 0,
             py::arg("lanes") = 1)
        .def(py::init<const Jet::TensorNetwork<tensor_t> &, const Jet::PathInfo::Path &,
                      const std::vector<std::string> &, const std::vector<int> &, int, int>(),
             py::arg("tn"), py::arg("path"), py::arg("sliced_indices"), py::arg("devices"), py::arg("flags") = 0,
             py::arg("lanes") = 1)
        .def_property_readonly("num_slices", &SC::NumSlices)
        .def_property_readonly("num_devices", &SC::NumDevices)
        .def_property_readonly("num_lanes", &SC::NumLanes)
        .def_property_readonly("flops", &SC::GetFlops)
        .def("contract", [](SC &sc, size_t first, py::object count) {
            const size_t n = count.is_none() ? sc.NumSlices() - first : count.cast<size_t>();
            py::gil_scoped_release release;
            return sc.Contract(first, n);
        }, py::arg("first") = 0, py::arg("count") = py::none())
        .def("last_milliseconds", &SC::LastMilliseconds);
}

template <class T> void BindIO(py::module_ &m, const char *file_name, const char *ser_name)
{
    using tensor_t = Jet::Tensor<T>;
    using File = Jet::TensorNetworkFile<tensor_t>;
    using Ser = Jet::TensorNetworkSerializer<tensor_t>;
    py::class_<File>(m, file_name)
        .def_property_readonly_static("dtype", [](const py::object &) { return py::dtype::of<T>(); })
        .def(py::init<const std::optional<Jet::PathInfo> &, const Jet::TensorNetwork<tensor_t> &>(),
             py::arg("path") = std::nullopt, py::arg("tensors") = Jet::TensorNetwork<tensor_t>())
        .def_readwrite("path", &File::path)
        .def_readwrite("tensors", &File::tensors);
    py::class_<Ser>(m, ser_name)
        .def_property_readonly_static("dtype", [](const py::object &) { return py::dtype::of<T>(); })
        .def(py::init<int>(), py::arg("indent") = -1)
        .def("__call__", [](Ser &s, const Jet::TensorNetwork<tensor_t> &tn) { return s(tn); }, py::arg("tn"))
        .def("__call__", [](Ser &s, const Jet::TensorNetwork<tensor_t> &tn, const Jet::PathInfo &p) { return s(tn, p); },
             py::arg("tn"), py::arg("path_info"))
        .def("__call__", [](Ser &s, const std::string &text, bool col_major) { return s(text, col_major); },
             py::arg("js_str"), py::arg("col_major") = false);
}

} // namespace

PYBIND11_MODULE(bindings, m)
{
    m.doc() = "B200-native bindings with the names of the reference's jet.bindings";
    BindTensor<std::complex<float>>(m, "TensorC64");
    BindTensor<std::complex<double>>(m, "TensorC128");
    BindTensorNetwork<std::complex<float>>(m, "TensorNetworkC64");
    BindTensorNetwork<std::complex<double>>(m, "TensorNetworkC128");
    BindPathInfo(m);
    BindContractors<std::complex<float>>(m, "TaskBasedContractorC64", "SlicedContractorC64");
    BindContractors<std::complex<double>>(m, "TaskBasedContractorC128", "SlicedContractorC128");
    BindIO<std::complex<float>>(m, "TensorNetworkFileC64", "TensorNetworkSerializerC64");
    BindIO<std::complex<double>>(m, "TensorNetworkFileC128", "TensorNetworkSerializerC128");
    m.def("version", &Jet::Version);
    py::register_exception<Jet::Exception>(m, "JetException", PyExc_RuntimeError);
    py::register_exception<Jet::JsonException>(m, "JsonException", PyExc_ValueError);
}
