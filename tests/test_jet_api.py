"""CPU tests of the host-side mirror of the reference interface (jet_b200.jet = the names of the
reference's `jet` package): everything that is bookkeeping and needs no GPU — constructors,
PathInfo, TaskBasedContractor task/parents/flops/memory accounting, serializer round trips and the
error behaviour.  Modelled on the reference's python/tests/*.py (cited per test)."""
import numpy as np
import pytest

from jet_b200 import jet


def make_network(dtype):
    # python/tests/conftest.py:6-21
    A = jet.Tensor(shape=[2, 2], indices=["i", "j"], data=[1, 1j, -1j, 1], dtype=dtype)
    B = jet.Tensor(shape=[2, 2], indices=["j", "k"], data=[1, 0, 0, 1], dtype=dtype)
    C = jet.Tensor(shape=[2], indices=["k"], data=[1, 0], dtype=dtype)
    tn = jet.TensorNetwork(dtype=dtype)
    tn.add_tensor(A, ["A", "Hermitian"])
    tn.add_tensor(B, ["B", "Identity", "Real"])
    tn.add_tensor(C, ["C", "Vector", "Real"])
    return tn


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
class TestBookkeeping:
    def test_factory_dtypes(self, dtype):
        # python/jet/factory.py:40-151: dtype dispatch, default complex128
        assert jet.Tensor(dtype=dtype).dtype == np.dtype(dtype)
        assert jet.Tensor().dtype == np.dtype("complex128")
        assert jet.TaskBasedContractor().dtype == np.dtype("complex128")
        with pytest.raises(TypeError, match="is not supported"):
            jet.Tensor(dtype="float32")

    def test_tensor_constructors_and_accessors(self, dtype):
        # python/tests/test_tensor.py (constructors, properties, set_shape quirk)
        t = jet.Tensor(dtype=dtype)
        assert len(t) == 1 and t.indices == [] and t.shape == [] and t.is_scalar() and t.scalar == 0
        t = jet.Tensor(shape=[2, 3], dtype=dtype)
        assert t.indices == ["?a", "?b"] and t.data == [0] * 6
        t = jet.Tensor(["i", "j"], [2, 2], [1, 2j, 3, 4], dtype=dtype)
        assert t.get_value([0, 1]) == 2j and t[3] == 4
        assert t.index_to_dimension_map == {"i": 2, "j": 2}
        t.set_value([1, 0], 7)
        assert t.data == [1, 2j, 7, 4]
        t.rename_index(0, "z")
        assert t.indices == ["z", "j"]
        t.shape = [4]
        assert t.shape == [4] and len(t) == 4
        with pytest.raises(RuntimeError, match="Size of data and tensor do not match."):
            t.data = [1]
        with pytest.raises(IndexError):
            t[9]
        assert np.array_equal(jet.Tensor(["i"], [2], [1, 2], dtype=dtype).array, np.array([1, 2], dtype=dtype))
        a = jet.Tensor(["i"], [2], dtype=dtype)
        b = jet.Tensor(a, dtype=dtype)
        a.fill_random(3)
        b.fill_random(3)
        assert a == b and a != jet.Tensor(["i"], [2], dtype=dtype)

    def test_tensor_network_nodes_edges(self, dtype):
        # python/tests/test_tensor_network.py
        tn = make_network(dtype)
        assert tn.num_tensors == 3 and tn.num_indices == 3
        assert [n.name for n in tn.nodes] == ["ij", "jk", "k"]
        assert tn.nodes[0].tags == ["A", "Hermitian"] and not tn.nodes[0].contracted
        assert tn.index_to_edge_map["j"].dim == 2 and tn.index_to_edge_map["j"].node_ids == [0, 1]
        assert tn.tag_to_node_id_map["Real"] == [1, 2] or sorted(tn.tag_to_node_id_map["Real"]) == [1, 2]
        assert tn.path == []
        with pytest.raises(IndexError):
            tn.nodes[3]

    def test_path_info(self, dtype):
        # python/tests/test_path_info.py
        tn = make_network(dtype)
        pi = jet.PathInfo(tn=tn, path=[[0, 1], [2, 3]])
        assert pi.num_leaves == 3 and pi.path == [(0, 1), (2, 3)]
        assert pi.index_to_size_map == {"i": 2, "j": 2, "k": 2}
        s = pi.steps
        assert [x.name for x in s] == ["ij", "jk", "k", "ik", "i"]
        assert s[3].children == (0, 1) and s[0].parent == 3 and s[4].parent == jet.PathStepInfo.MISSING_ID
        assert s[3].contracted_indices == ["j"] and s[3].tensor_indices == ["i", "k"]
        assert s[3].tags == ["A", "Hermitian", "B", "Identity", "Real"]
        assert pi.total_flops() == 2 * 2 * 4 + 2 * 4 and pi.total_memory() == 4 + 4 + 2 + 4 + 2
        with pytest.raises(RuntimeError, match="Node ID 2 in contraction path pair is invalid."):
            jet.PathInfo(tn=tn, path=[[0, 9]])

    def test_task_based_contractor_accounting(self, dtype):
        # python/tests/test_task_based_contractor.py:8-50 (everything before contract())
        tbc = jet.TaskBasedContractor(dtype=dtype)
        assert tbc.name_to_tensor_map == {} and tbc.name_to_parents_map == {} and tbc.results == []
        assert tbc.reduction_result == jet.Tensor(dtype=dtype) and tbc.flops == 0 and tbc.memory == 0
        tn = make_network(dtype)
        path = jet.PathInfo(tn=tn, path=[[0, 1], [2, 3]])
        assert tbc.add_contraction_tasks(tn, path) == 0
        assert tbc.add_deletion_tasks() == 4
        assert tbc.add_reduction_task() == 1 and tbc.add_reduction_task() == 0
        m = tbc.name_to_tensor_map
        assert set(m) == {"0:ij", "1:jk", "2:k", "3:ik", "4:i:results[0]"}
        assert m["0:ij"] == tn.nodes[0].tensor and m["3:ik"] is None and m["4:i:results[0]"] is None
        assert tbc.name_to_parents_map == {"0:ij": {"3:ik"}, "1:jk": {"3:ik"}, "2:k": {"4:i:results[0]"},
                                           "3:ik": {"4:i:results[0]"}}
        assert tbc.flops == 2 * 2 * 4 + 2 * 4 and tbc.memory == 6
        # a second, identical network shares everything but the final task
        assert tbc.add_contraction_tasks(tn, path) == 1
        assert "4:i:results[1]" in tbc.name_to_tensor_map

    def test_serializer_round_trip_and_errors(self, dtype):
        # python/tests/test_tensor_network_io.py + test/Test_TensorNetworkIO.cpp:12-127
        tn = make_network(dtype)
        pi = jet.PathInfo(tn=tn, path=[[0, 1], [2, 3]])
        ser = jet.TensorNetworkSerializer(dtype=dtype)
        text = ser(tn, pi)
        assert text.startswith('{"path":[[0,1],[2,3]],"tensors":[[["A","Hermitian"],["i","j"],[2,2],[[1.0,0.0],[0.0,1.0],')
        f = ser(text)
        assert f.path.path == [(0, 1), (2, 3)] and f.tensors.num_tensors == 3
        assert f.tensors.nodes[0].tensor == tn.nodes[0].tensor and f.tensors.nodes[2].tags == ["C", "Vector", "Real"]
        assert ser(f.tensors, f.path) == text
        assert '"path"' not in ser(tn)
        assert ser(ser(tn)).path is None
        with pytest.raises(ValueError):
            ser("")
        for bad in ("[]", "{}", '{"path": [[0,1]]}'):
            with pytest.raises(RuntimeError, match="Error parsing tensor network file"):
                ser(bad)
        with pytest.raises(RuntimeError, match=r"\[1.0\]"):
            ser('{"tensors": [[["I0"], ["a"], [2], [[1.0], [0.0,0.0]]]]}')
        file = jet.TensorNetworkFile(dtype=dtype)
        assert file.path is None and file.tensors.num_tensors == 0


def test_reference_data_file_loads_identically(data_dir):
    """The shipped m10.json loads through both host layers (C++ serializer and the Python
    NetworkFile) to the same leaves and path."""
    import os
    from jet_b200 import NetworkFile
    text = open(os.path.join(data_dir, "m10.json")).read()
    f = jet.TensorNetworkSerializer(dtype="complex64")(text)
    n = NetworkFile.loads(text, np.complex64)
    assert f.tensors.num_tensors == len(n.tensors) == 322
    assert f.path.path == n.path and len(n.path) == 321
    for node, (idx, arr) in list(zip(f.tensors.nodes, n.tensors))[::40]:
        assert node.indices == idx and np.array_equal(node.tensor.array, arr)
    assert f.path.total_flops() == 12791615632.0  # tests/golden/amplitudes.json m10_full jet_flops
