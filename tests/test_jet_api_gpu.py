"""GPU tests of the `jet`-compatible API (pybind11 module over the drop-in C++ headers) and of the
C++ test binary.  Modelled on the reference's python/tests/test_task_based_contractor.py,
test_tensor.py, test_tensor_network.py."""
import json
import os
import subprocess

import numpy as np
import pytest

from jet_b200 import jet
from tests.test_jet_api import make_network

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
class TestJetApiOnGpu:
    def test_tbc_contract(self, dtype):
        # python/tests/test_task_based_contractor.py:18-50
        tn = make_network(dtype)
        path = jet.PathInfo(tn=tn, path=[[0, 1], [2, 3]])
        tbc = jet.TaskBasedContractor(dtype=dtype)
        assert tbc.add_contraction_tasks(tn, path) == 0
        assert tbc.add_deletion_tasks() == 4
        assert tbc.add_reduction_task() == 1
        tbc.contract()
        want = jet.Tensor(shape=[2], indices=["i"], data=[1, -1j], dtype=dtype)
        assert tbc.name_to_tensor_map == {"0:ij": None, "1:jk": None, "2:k": None, "3:ik": None, "4:i:results[0]": want}
        assert tbc.results == [want] and tbc.reduction_result == want
        assert tbc.flops == 2 * 2 * 4 + 2 * 4 and tbc.memory == 6

    def test_tbc_keeps_intermediates_without_deletion(self, dtype):
        tn = make_network(dtype)
        tbc = jet.TaskBasedContractor(dtype=dtype)
        tbc.add_contraction_tasks(tn, jet.PathInfo(tn=tn, path=[[0, 1], [2, 3]]))
        tbc.contract()
        m = tbc.name_to_tensor_map
        assert m["3:ik"] == jet.Tensor(["i", "k"], [2, 2], [1, 1j, -1j, 1], dtype=dtype)
        assert m["0:ij"] == tn.nodes[0].tensor

    def test_network_contract_and_slicing(self, dtype):
        # python/tests/test_tensor_network.py (contract, slice_indices)
        tn = make_network(dtype)
        r = tn.contract([[0, 1], [2, 3]])
        assert r == jet.Tensor(["i"], [2], [1, -1j], dtype=dtype)
        assert tn.num_tensors == 5 and tn.nodes[0].contracted and tn.path == [(0, 1), (2, 3)]
        total = None
        for v in range(2):
            s = make_network(dtype)
            s.slice_indices(["j"], v)
            assert s.nodes[0].name == f"ij({v})" and s.nodes[0].tensor.indices == ["i"]
            part = s.contract([[0, 1], [2, 3]])
            total = part if total is None else jet.add_tensors(total, part)
        assert total == r
        with pytest.raises(RuntimeError, match="Sliced index does not exist."):
            tn.slice_indices(["nope"], 0)

    def test_tensor_free_functions(self, dtype):
        # python/tests/test_tensor.py (contract/transpose/slice/add/reshape/conj)
        a = jet.Tensor(["i", "j"], [2, 3], [1, 2, 3, 4, 5, 6], dtype=dtype)
        b = jet.Tensor(["j"], [3], [1, 1j, -1], dtype=dtype)
        c = jet.contract_tensors(a, b)
        assert c.indices == ["i"] and c.data == [1 + 2j - 3, 4 + 5j - 6]
        t = jet.transpose(a, ["j", "i"])
        assert t.shape == [3, 2] and t.data == [1, 4, 2, 5, 3, 6]
        assert jet.transpose(a, [1, 0]) == t and a.transpose(["j", "i"]) == t
        assert jet.slice_index(a, "i", 1).data == [4, 5, 6] and jet.slice_index(a, "j", 2).data == [3, 6]
        assert jet.add_tensors(a, t).data == [2, 4, 6, 8, 10, 12]
        assert jet.reshape(a, [3, 2]).indices == ["?a", "?b"]
        assert jet.conj(b).data == [1, -1j, -1]
        with pytest.raises(RuntimeError, match="Size is inconsistent between tensors."):
            jet.reshape(a, [4, 2])
        with pytest.raises(RuntimeError, match="Tensor addition with disjoint indices is not supported."):
            jet.add_tensors(a, jet.Tensor(["i", "q"], [2, 3], dtype=dtype))
        dot = jet.contract_tensors(b, b)
        assert dot.indices == [] and dot.is_scalar() and dot.scalar == 1 - 1 + 1


def test_sliced_contractor_on_m10(data_dir):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))
    text = open(os.path.join(data_dir, "m10.json")).read()
    f = jet.TensorNetworkSerializer(dtype="complex64")(text)
    sc = jet.SlicedContractor(f.tensors, f.path.path, "p7 s7 h4 m1 m2 I2".split(), dtype="complex64")
    assert sc.num_slices == 64 and sc.flops == gold["m10_s6_slice0_complex64"]["jet_flops"]
    r = sc.contract()
    want = complex(gold["m10_s6_sum64_complex128"]["re"], gold["m10_s6_sum64_complex128"]["im"])
    assert abs(r.scalar - want) / abs(want) < 1e-5
    r0 = sc.contract(0, 1)
    want0 = complex(gold["m10_s6_slice0_complex64"]["re"], gold["m10_s6_slice0_complex64"]["im"])
    assert abs(r0.scalar - want0) / abs(want0) < 1e-5
    assert sc.last_milliseconds() > 0


def test_cpp_dropin_binary(data_dir):
    """The C++ restatement of the reference's Catch2 tests over include/jet/*.hpp."""
    exe = os.path.join(ROOT, "jet_b200", "cpp", "test_dropin")
    assert os.path.exists(exe), "build with __graft_entry__.build()"
    p = subprocess.run([exe, os.path.join(data_dir, "m10.json")], capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:], p.stderr[-4000:])
    assert p.returncode == 0, p.stderr[-4000:]
    assert " 0 failed" in p.stdout


def _run(cmd, timeout=900, env=None):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert p.returncode == 0, (cmd, p.stdout[-2000:], p.stderr[-4000:])
    return p.stdout


def _parse_result(stdout):
    # "result=Size=1\nIndices={}\nData={(re,im)}" — the reference's Tensor operator<<
    import re

    m = re.search(r"Data\s*=\s*\{\(([-+0-9.eE]+),([-+0-9.eE]+)\)", stdout)
    assert m, stdout[-500:]
    return complex(float(m.group(1)), float(m.group(2)))


def test_reference_jet_sliced_driver_on_the_plan_engine(data_dir):
    """The reference's OWN benchmark driver (examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp, compiled
    from where it lies against include/ by `make -C jet_b200/cpp examples`): 64 SliceIndices copies ->
    AddContractionTasks -> AddReductionTask -> Contract().wait() must give the reference's 64-slice sum."""
    exe = os.path.join(ROOT, "jet_b200", "cpp", "examples", "jet_sliced_m10")
    if not os.path.exists(exe):
        pytest.skip("examples not built (needs /root/reference at build time)")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))
    out = _run([exe, os.path.join(data_dir, "m10.json"), "1", "6"])
    got = _parse_result(out)
    want = complex(gold["m10_s6_sum64_complex128"]["re"], gold["m10_s6_sum64_complex128"]["im"])
    assert abs(got - want) / abs(want) < 2e-5, (got, want)  # the driver prints 6 significant digits
    assert "number_of_slices = 64" in out


def test_reference_jet_full_driver(data_dir):
    exe = os.path.join(ROOT, "jet_b200", "cpp", "examples", "jet_full_m10")
    if not os.path.exists(exe):
        pytest.skip("examples not built (needs /root/reference at build time)")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))
    got = _parse_result(_run([exe, os.path.join(data_dir, "m10.json"), "1"]))
    want = complex(gold["m10_full_complex64"]["re"], gold["m10_full_complex64"]["im"])
    assert abs(got - want) / abs(want) < 2e-5, (got, want)


@pytest.mark.parametrize("indices,n,key", [("p7,s7,h4,m1,m2,I2", 64, "m10_s6_sum64_complex128"),
                                           ("p7,s7,h4,m1,m2,I2,V4,z2,t4,C1", 16, "m10_s10_sum_first16_complex128")])
def test_tbc_flow_matches_sliced_contractor(data_dir, indices, n, key):
    """TaskBasedContractor fed with SliceIndices copies (lowered onto plan sets) == SlicedContractor == golden."""
    exe = os.path.join(ROOT, "jet_b200", "cpp", "tbc_bench")
    assert os.path.exists(exe), "build with __graft_entry__.build()"
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))
    want = complex(gold[key]["re"], gold[key]["im"])
    res = {}
    for api in ("tbc", "sliced"):
        line = _run([exe, os.path.join(data_dir, "m10.json"), indices, "--api", api, "--slices", str(n), "--reps", "2"])
        res[api] = json.loads(line.strip().splitlines()[-1])
        got = complex(*res[api]["result"])
        assert abs(got - want) / abs(want) < 1e-5, (api, got, want)
    assert res["tbc"]["shared_tasks"] > 0
    print(res)
