"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference headers
(oracle/_ref/libjetref.so, built by `make -C oracle` from /root/reference/include).

Run here (the container that has /root/reference):  python tests/golden/make_golden.py [--heavy]
Outputs (committed):
  permute_cases.npz   inputs + Jet::Tensor::Transpose outputs (QFlexPermuter / DefaultPermuter)
  contract_cases.npz  inputs + Jet::Tensor::ContractTensors outputs
  amplitudes.json     network amplitudes from TensorNetwork::Contract / TaskBasedContractor
The fixtures pin oracle/jet_oracle.py (tests/test_oracle.py) and the CUDA path (-m gpu tests).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def rand_c(rng, n, dtype):
    real = np.float32 if dtype == np.complex64 else np.float64
    return (rng.uniform(-1, 1, n).astype(real) + 1j * rng.uniform(-1, 1, n).astype(real)).astype(dtype)


def permute_cases():
    rng = np.random.default_rng(20240601)
    cases = {}
    specs = []
    # shapes from the reference's own permuter tests (test/Test_Permuter.cpp:182-342,458-490)
    for shape in ([2, 2, 2, 2], [4, 4], [2, 4, 2], [4, 2, 2], [2, 2, 4], [2] * 9, [2, 3, 5], [3, 3], [5, 2, 3, 2]):
        for _ in range(3):
            specs.append((shape, list(rng.permutation(len(shape)))))
    specs.append(([2, 2, 2, 2], [3, 2, 1, 0]))
    specs.append(([2] * 9, [8, 7, 6, 5, 4, 3, 2, 1, 0]))
    # random pow2 ranks incl. mixed extents, and patterns P1..P5 of SURVEY §8(d) at small rank
    for r in (1, 2, 3, 5, 7, 10, 11, 12):
        for _ in range(2):
            specs.append(([2] * r, list(rng.permutation(r))))
    for shape in ([4, 8, 2, 16], [16, 2, 4, 4, 2], [8, 8, 8], [2, 64, 4], [32, 32]):
        specs.append((shape, list(rng.permutation(len(shape)))))
    r = 12
    pull = sorted(rng.choice(r, 3, replace=False).tolist())
    rest = [i for i in range(r) if i not in pull]
    specs.append(([2] * r, rest + pull))  # P1 pull to back
    specs.append(([2] * r, pull + rest))  # P2 pull to front
    specs.append(([2] * r, list(range(r))[::-1]))  # P4 reversal
    specs.append(([2] * r, list(rng.permutation(r - 5)) + list(range(r - 5, r))))  # P5 last 5 fixed
    specs.append(([2] * r, list(range(r))))  # identity
    n = 0
    for shape, perm in specs:
        for dtype in (np.complex64, np.complex128):
            size = int(np.prod(shape))
            x = rand_c(rng, size, dtype)
            y = ref.transpose(x, shape, perm)
            cases[f"c{n}_shape"] = np.array(shape, dtype=np.int64)
            cases[f"c{n}_perm"] = np.array(perm, dtype=np.int32)
            cases[f"c{n}_in"] = x
            cases[f"c{n}_out"] = y
            n += 1
    cases["count"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "permute_cases.npz"), **cases)
    print("permute cases:", n)


def contract_cases():
    rng = np.random.default_rng(777)
    cases = {}
    n = 0
    specs = []

    def spec(shape_a, ids_a, shape_b, ids_b):
        specs.append((shape_a, ids_a, shape_b, ids_b))

    # corners of MultiplyTensorData (TensorHelpers.hpp:147-167): GEMM / GEMV / GEMV^T / DOTU / outer
    spec([2, 3], [0, 1], [3, 2], [1, 2])
    spec([2, 12], [0, 1], [12, 2], [1, 2])
    spec([4, 3], [0, 1], [3], [1])
    spec([3], [0], [3, 4], [0, 1])
    spec([5], [0], [5], [0])
    spec([2, 2], [0, 1], [2, 2], [1, 0])
    spec([2, 3], [0, 1], [4], [2])
    spec([2, 3, 5], [0, 1, 2], [5, 3, 4], [2, 1, 3])  # the reference's CPU-vs-GPU case
    spec([3, 2, 4], [0, 1, 2], [4, 5, 2], [2, 3, 1])
    # tensor-network-like pow2 cases: big x small with scattered common indices
    for ra, rb, c in ((8, 3, 1), (10, 4, 2), (11, 4, 2), (11, 6, 3), (11, 6, 4), (6, 6, 3), (3, 10, 2), (4, 12, 2),
                      (9, 9, 9), (6, 6, 0), (11, 3, 3), (12, 2, 1), (2, 12, 1), (7, 7, 4), (12, 12, 10), (13, 13, 12)):
        ids_a = list(range(ra))
        common = sorted(rng.choice(ra, c, replace=False).tolist())
        new = list(range(100, 100 + rb - c))
        ids_b = common + new
        ids_b = [ids_b[i] for i in rng.permutation(rb)]
        spec([2] * ra, ids_a, [2] * rb, ids_b)
    # mixed pow2 extents (dim 4 like GBS)
    spec([4, 4, 4, 4, 4], [0, 1, 2, 3, 4], [4, 4, 4], [3, 9, 1])
    spec([4, 2, 8, 2], [0, 1, 2, 3], [8, 4, 2], [2, 0, 7])
    for sa, ia, sb, ib in specs:
        for dtype in (np.complex64, np.complex128):
            a = rand_c(rng, int(np.prod(sa)), dtype).reshape(sa)
            b = rand_c(rng, int(np.prod(sb)), dtype).reshape(sb)
            c_ = ref.contract(ia, a, ib, b)
            cases[f"c{n}_sa"] = np.array(sa, dtype=np.int64)
            cases[f"c{n}_ia"] = np.array(ia, dtype=np.int32)
            cases[f"c{n}_a"] = a.reshape(-1)
            cases[f"c{n}_sb"] = np.array(sb, dtype=np.int64)
            cases[f"c{n}_ib"] = np.array(ib, dtype=np.int32)
            cases[f"c{n}_b"] = b.reshape(-1)
            cases[f"c{n}_c"] = c_
            n += 1
    cases["count"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "contract_cases.npz"), **cases)
    print("contract cases:", n)


M10_SLICED = "p7 s7 h4 m1 m2 I2 V4 z2 t4 C1".split()
M12_SLICED = "h5 m H10 w y J S G10 P0".split()


def amplitudes(heavy):
    path = os.path.join(OUT, "amplitudes.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    ref.set_blas_threads(1)

    def put(key, r, extra=None):
        out[key] = {"re": float(r[0].real), "im": float(r[0].imag)}
        if extra:
            out[key].update(extra)
        print(key, out[key])

    m10 = open(os.path.join(ref.DATA_DIR, "m10.json")).read()
    for dt in ("complex64", "complex128"):
        for v in range(4):
            r, _, fl = ref.network(m10, dt, M10_SLICED[:6], v, 0)
            put(f"m10_s6_slice{v}_{dt}", r, {"jet_flops": fl})
        r, _, _ = ref.network(m10, dt, M10_SLICED[:6], 0, 1, 8, 64)
        put(f"m10_s6_sum64_{dt}", r)
        r, _, _ = ref.network(m10, dt, M10_SLICED[:10], 0, 1, 8, 16)
        put(f"m10_s10_sum_first16_{dt}", r)
    for tot in (0, 10, 20, 30, 40, 50, 60):
        g = open(os.path.join(ref.DATA_DIR, f"gbs_dim2_nc1_lw8_rp5_fock4_total{tot}_0.kraken.json")).read()
        for dt in ("complex64", "complex128"):
            r, _, fl = ref.network(g, dt)
            put(f"gbs_fock4_total{tot}_{dt}", r, {"jet_flops": fl})
    if heavy:
        r, _, fl = ref.network(m10, "complex64")
        put("m10_full_complex64", r, {"jet_flops": fl})
        m12 = open(os.path.join(ref.DATA_DIR, "m12.json")).read()
        ref.set_blas_threads(8)
        for v in (0, 1):
            r, sec, fl = ref.network(m12, "complex64", M12_SLICED, v, 0)
            put(f"m12_s9_slice{v}_complex64", r, {"jet_flops": fl, "ref_seconds_here": sec})
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if not ref.available():
        sys.exit("oracle/_ref/libjetref.so missing: run `make -C oracle`")
    heavy = "--heavy" in sys.argv
    if "--amplitudes-only" not in sys.argv:
        permute_cases()
        contract_cases()
    amplitudes(heavy)
