"""CPU tests: pin oracle/jet_oracle.py (the numpy restatement) against
  (i) fixtures produced by the unmodified reference (tests/golden/*.npz, amplitudes.json) and
 (ii) the reference's own known-answer tests, restated (file:line cited per test).
"""
import json
import os

import numpy as np
import pytest

from oracle import jet_oracle as jo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cases(name):
    z = np.load(os.path.join(GOLDEN, name))
    return z, int(z["count"])


def test_permute_matches_reference_fixtures():
    z, n = _cases("permute_cases.npz")
    assert n >= 100
    for i in range(n):
        out = jo.transpose(z[f"c{i}_in"], z[f"c{i}_shape"].tolist(), z[f"c{i}_perm"].tolist())
        assert out.dtype == z[f"c{i}_out"].dtype
        assert np.array_equal(out, z[f"c{i}_out"]), f"case {i}"


def test_contract_matches_reference_fixtures():
    z, n = _cases("contract_cases.npz")
    assert n >= 50
    for i in range(n):
        sa, sb = z[f"c{i}_sa"].tolist(), z[f"c{i}_sb"].tolist()
        ia = [str(v) for v in z[f"c{i}_ia"]]
        ib = [str(v) for v in z[f"c{i}_ib"]]
        a = z[f"c{i}_a"].reshape(sa)
        b = z[f"c{i}_b"].reshape(sb)
        _, c = jo.contract((ia, a), (ib, b))
        ref = z[f"c{i}_c"]
        tol = 1e-5 if a.dtype == np.complex64 else 1e-12
        err = np.linalg.norm(c.reshape(-1) - ref) / max(np.linalg.norm(ref), 1e-300)
        assert err < tol, f"case {i}: {err}"


def test_permuter_kat_2x2x2x2():
    # test/Test_Permuter.cpp:182-230 — element i = i, indices {a,b,c,d} -> {d,c,b,a} etc.
    data = np.arange(16, dtype=np.complex64)
    out = jo.transpose(data, [2, 2, 2, 2], [3, 2, 1, 0])
    expect = np.arange(16).reshape(2, 2, 2, 2).transpose(3, 2, 1, 0).reshape(-1)
    assert np.array_equal(out, expect.astype(np.complex64))
    out = jo.transpose(data, [2, 2, 2, 2], [0, 1, 3, 2])
    assert np.array_equal(out.real, [0, 2, 1, 3, 4, 6, 5, 7, 8, 10, 9, 11, 12, 14, 13, 15])


def test_permuter_kat_2x3x5():
    # test/Test_Permuter.cpp:319-342 (DefaultPermuter, non power of two)
    data = np.arange(30, dtype=np.complex128)
    out = jo.transpose(data, [2, 3, 5], [2, 0, 1])
    assert out[1] == 5 and out[6] == 1 and out[29] == 29


def test_contract_kat_matrix_product():
    # test/Test_Tensor.cpp:372-397 — 2x12 . 12x2 with all elements (0.5, 0.25): each output is
    # 12 * (0.5+0.25i)^2 = (2.25, 3.0)
    a = np.full((2, 12), 0.5 + 0.25j, dtype=np.complex64)
    b = np.full((12, 2), 0.5 + 0.25j, dtype=np.complex64)
    idx, c = jo.contract((["i", "j"], a), (["j", "k"], b))
    assert idx == ["i", "k"]
    assert np.allclose(c, 2.25 + 3.0j)


def test_contract_kat_index_order_and_scalar():
    # test/Test_Tensor.cpp:543-563 — full contraction gives a scalar; no conjugation (DOTU)
    a = np.array([1j, 1, 2, 3], dtype=np.complex64)
    idx, c = jo.contract((["i"], a), (["i"], a))
    assert idx == [] and c.shape == ()
    assert c == (-1 + 1 + 4 + 9)


def test_network_contract_kat():
    # test/Test_TensorNetwork.cpp:529-552 — (A0,B1)[2,3].(C2,B1)[2,3].(C2,D3)[2,2], element i=(i,2i),
    # path {{1,2},{0,3}} -> {(-308,-56),(-517,-94),(-1100,-200),(-1804,-328)}
    def mk(shape):
        n = int(np.prod(shape))
        return (np.arange(n) + 2j * np.arange(n)).astype(np.complex64).reshape(shape)

    net = jo.Network([(["A0", "B1"], mk([2, 3])), (["C2", "B1"], mk([2, 3])), (["C2", "D3"], mk([2, 2]))],
                     [(1, 2), (0, 3)])
    idx, r = net.contract()
    assert np.array_equal(r.reshape(-1), np.array([-308 - 56j, -517 - 94j, -1100 - 200j, -1804 - 328j]))


def test_slice_indices_kat():
    # test/Test_TensorNetwork.cpp:201-337 — slicing [1,:,2] of element-i tensors
    t = (["A0", "B1", "C2"], np.arange(2 * 3 * 4).reshape(2, 3, 4).astype(np.complex64))
    net = jo.Network([t], [])
    s = net.slice_indices(["A0", "C2"], 1 * 4 + 2)
    assert s.tensors[0][0] == ["B1"]
    assert np.array_equal(s.tensors[0][1].real, [14, 18, 22])
    with pytest.raises(ValueError, match="Sliced index does not exist."):
        net.slice_indices(["Z9"], 0)


def test_add_tensors_kat():
    # test/Test_Tensor.cpp:659-668 and include/jet/Tensor.hpp:415-424
    a = (["i", "j"], np.arange(6).reshape(2, 3).astype(np.complex64))
    b = (["j", "i"], np.arange(6).reshape(3, 2).astype(np.complex64))
    idx, c = jo.add_tensors(a, b)
    assert idx == ["i", "j"]
    assert np.array_equal(c, a[1] + b[1].T)
    zero = ([], np.zeros((), np.complex64))
    assert np.array_equal(jo.add_tensors(zero, a)[1], a[1])


@pytest.mark.parametrize("dt", ["complex64", "complex128"])
def test_m10_slice_amplitudes_match_reference(data_dir, dt):
    gold = json.load(open(os.path.join(GOLDEN, "amplitudes.json")))
    net = jo.Network.from_file(os.path.join(data_dir, "m10.json"), dt)
    sliced = "p7 s7 h4 m1 m2 I2".split()
    assert abs(net.slice_indices(sliced, 0).jet_flops() - gold[f"m10_s6_slice0_{dt}"]["jet_flops"]) < 1
    tol = 1e-5 if dt == "complex64" else 1e-12
    for v in (0, 3):
        g = gold[f"m10_s6_slice{v}_{dt}"]
        r = jo.amplitude(net, sliced, [v]).reshape(-1)[0]
        ref = complex(g["re"], g["im"])
        assert abs(r - ref) / abs(ref) < tol


def test_gbs_amplitude_matches_reference(data_dir):
    gold = json.load(open(os.path.join(GOLDEN, "amplitudes.json")))
    net = jo.Network.from_file(os.path.join(data_dir, "gbs_dim2_nc1_lw8_rp5_fock4_total10_0.kraken.json"),
                               "complex128")
    g = gold["gbs_fock4_total10_complex128"]
    r = jo.amplitude(net).reshape(-1)[0]
    ref = complex(g["re"], g["im"])
    assert abs(r - ref) / abs(ref) < 1e-12
