"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product (``jet_b200``) never does.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here against
(i) the reference's own known-answer tests restated in ``tests/golden/kat.py`` and
(ii) ``tests/golden/*.npz`` fixtures produced by running the unmodified reference headers
(``oracle/_ref/libjetref.so`` built by ``oracle/Makefile``; generator: ``tests/golden/make_golden.py``).

Each function cites the reference file:line it restates (paths relative to /root/reference).
Tensors are ``(indices: list[str], array: np.ndarray)`` pairs, row-major (C order), exactly like
``Jet::Tensor`` (include/jet/Tensor.hpp:766-775).
"""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Tuple

import numpy as np

Tensor = Tuple[List[str], np.ndarray]

DTYPES = {"complex64": np.complex64, "complex128": np.complex128, 0: np.complex64, 1: np.complex128}


# --------------------------------------------------------------------------------------
# Permuter  (include/jet/permute/Permuter.hpp:50-79; semantics in QFlex.hpp / Default.hpp)
# --------------------------------------------------------------------------------------
def transpose(data: np.ndarray, shape: Sequence[int], perm: Sequence[int]) -> np.ndarray:
    """out axis j = in axis perm[j]; data row-major over ``shape``.

    Restates Permuter::Transpose (permute/Permuter.hpp:50-79): for every multi-index x over the new
    shape, out[ravel(x)] = in[ravel(y)] with y[perm[j]] = x[j].  No arithmetic: bit-exact.
    """
    a = np.asarray(data).reshape(tuple(shape))
    return np.ascontiguousarray(np.transpose(a, tuple(perm))).reshape(-1)


def transpose_tensor(t: Tensor, new_indices: Sequence[str]) -> Tensor:
    """Tensor::Transpose(tensor, new_indices) (include/jet/Tensor.hpp:579-612)."""
    idx, arr = t
    perm = [idx.index(i) for i in new_indices]
    return list(new_indices), np.ascontiguousarray(np.transpose(arr, perm))


# --------------------------------------------------------------------------------------
# ContractTensors (include/jet/Tensor.hpp:709-752) + MultiplyTensorData
# (include/jet/TensorHelpers.hpp:131-168)
# --------------------------------------------------------------------------------------
def contraction_indices(ia: Sequence[str], ib: Sequence[str]):
    """left = A\\B in A's order, right = B\\A in B's order, common = A∩B in A's order
    (Tensor.hpp:714-718; Utilities.hpp:317-369)."""
    sb, sa = set(ib), set(ia)
    left = [i for i in ia if i not in sb]
    right = [i for i in ib if i not in sa]
    common = [i for i in ia if i in sb]
    return left, right, common


def contract(a: Tensor, b: Tensor) -> Tensor:
    """C[left ++ right] = sum_common A * B, no conjugation (Tensor.hpp:709-752).

    A is permuted to (left ++ common) and viewed M x K, B to (common ++ right) viewed K x N
    (Tensor.hpp:744-745) and multiplied row-major with alpha=1, beta=0
    (TensorHelpers.hpp:147-167: GEMM / GEMV / DOTU are the same product with N=1 / M=1 corners).
    All extents are taken from A for left/common and from B for right (Tensor.hpp:720-730).
    """
    ia, A = a
    ib, B = b
    # a scalar tensor (no indices, one element: Tensor.hpp:47, and what SliceIndex leaves of a rank-1 tensor)
    # may arrive with shape (1,); give every operand exactly one axis per index
    A = np.reshape(A, np.shape(A)[len(np.shape(A)) - len(ia):] if len(ia) else ())
    B = np.reshape(B, np.shape(B)[len(np.shape(B)) - len(ib):] if len(ib) else ())
    left, right, common = contraction_indices(ia, ib)
    dim_a = dict(zip(ia, A.shape))
    dim_b = dict(zip(ib, B.shape))
    m = int(np.prod([dim_a[i] for i in left], dtype=np.int64)) if left else 1
    k = int(np.prod([dim_a[i] for i in common], dtype=np.int64)) if common else 1
    n = int(np.prod([dim_b[i] for i in right], dtype=np.int64)) if right else 1
    At = np.transpose(A, [ia.index(i) for i in left + common]).reshape(m, k)
    Bt = np.transpose(B, [ib.index(i) for i in common + right]).reshape(k, n)
    C = At @ Bt
    shape = [dim_a[i] for i in left] + [dim_b[i] for i in right]
    return left + right, np.ascontiguousarray(C).reshape(shape)


# --------------------------------------------------------------------------------------
# SliceIndex / Reshape / AddTensors
# --------------------------------------------------------------------------------------
def slice_index(t: Tensor, index: str, value: int) -> Tensor:
    """Tensor::SliceIndex (include/jet/Tensor.hpp:494-526): fix `index` to `value`, drop the axis."""
    idx, arr = t
    ax = idx.index(index)
    out = np.take(arr, value, axis=ax)
    return [i for i in idx if i != index], np.ascontiguousarray(out)


def add_tensors(a: Tensor, b: Tensor) -> Tensor:
    """Tensor::AddTensors (include/jet/Tensor.hpp:413-454): result in A's index order; B may be
    permuted; a default tensor (no indices, one zero) is the additive identity (:415-424)."""
    ia, A = a
    ib, B = b
    if not ia and A.size == 1 and A.reshape(-1)[0] == 0:
        return list(ib), B.copy()
    if not ib and B.size == 1 and B.reshape(-1)[0] == 0:
        return list(ia), A.copy()
    if set(ia) != set(ib):
        raise ValueError("Tensor addition with disjoint indices is not supported.")
    Bt = np.transpose(B, [ib.index(i) for i in ia]) if ia else B
    return list(ia), A + Bt


# --------------------------------------------------------------------------------------
# Tensor network: file format, slicing, path replay
# --------------------------------------------------------------------------------------
class Network:
    """Leaves + path as stored by TensorNetworkSerializer (include/jet/TensorNetworkIO.hpp:94-187):
    ``{"path": [[i, j], ...], "tensors": [[tags, indices, shape, [[re, im], ...]], ...]}``."""

    def __init__(self, tensors: List[Tensor], path: List[Tuple[int, int]], tags=None):
        self.tensors = tensors
        self.path = [tuple(p) for p in path]
        self.tags = tags or [[] for _ in tensors]

    @staticmethod
    def from_json(text: str, dtype="complex64") -> "Network":
        js = json.loads(text)
        dt = DTYPES[dtype]
        tensors, tags = [], []
        for tg, idx, shape, data in js["tensors"]:
            arr = np.array([complex(re, im) for re, im in data], dtype=np.complex128)
            tensors.append((list(idx), arr.astype(dt).reshape(tuple(shape))))
            tags.append(list(tg))
        return Network(tensors, js.get("path", []), tags)

    @staticmethod
    def from_file(path: str, dtype="complex64") -> "Network":
        with open(path) as f:
            return Network.from_json(f.read(), dtype)

    def to_json(self) -> str:
        out = {"path": [list(p) for p in self.path], "tensors": []}
        for (idx, arr), tg in zip(self.tensors, self.tags):
            flat = arr.reshape(-1)
            out["tensors"].append(
                [list(tg), list(idx), [int(s) for s in arr.shape], [[float(z.real), float(z.imag)] for z in flat]]
            )
        return json.dumps(out, separators=(",", ":"))

    def index_dims(self) -> Dict[str, int]:
        d = {}
        for idx, arr in self.tensors:
            for i, s in zip(idx, arr.shape):
                d[i] = int(s)
        return d

    def slice_indices(self, indices: Sequence[str], value: int) -> "Network":
        """TensorNetwork::SliceIndices (include/jet/TensorNetwork.hpp:210-284): `value` is the
        row-major ravel over the listed indices (first listed index slowest, Utilities.hpp:463-475)."""
        dims = self.index_dims()
        for i in indices:
            if i not in dims:
                raise ValueError("Sliced index does not exist.")
        shape = [dims[i] for i in indices]
        vals = list(np.unravel_index(value, shape)) if indices else []
        out = []
        for t in self.tensors:
            for i, v in zip(indices, vals):
                if i in t[0]:
                    t = slice_index(t, i, int(v))
            out.append(t)
        return Network(out, self.path, self.tags)

    def num_slices(self, indices: Sequence[str]) -> int:
        dims = self.index_dims()
        return int(np.prod([dims[i] for i in indices], dtype=np.int64)) if indices else 1

    def contract(self, keep_steps: bool = False):
        """TensorNetwork::Contract(path) (include/jet/TensorNetwork.hpp:301-328,394-421): step i
        contracts nodes path[i] and appends the result as node num_leaves + i."""
        nodes: List[Tensor] = list(self.tensors)
        alive = [True] * len(nodes)
        for a, b in self.path:
            nodes.append(contract(nodes[a], nodes[b]))
            alive.append(True)
            if not keep_steps:
                nodes[a] = None
                nodes[b] = None
        return nodes if keep_steps else nodes[-1]

    def path_steps(self):
        """Symbolic replay (include/jet/PathInfo.hpp:262-297): per step (rankA, rankB, M, N, K)."""
        dims = self.index_dims()
        nodes = [list(t[0]) for t in self.tensors]
        steps = []
        for a, b in self.path:
            left, right, common = contraction_indices(nodes[a], nodes[b])
            m = int(np.prod([dims[i] for i in left], dtype=object)) if left else 1
            n = int(np.prod([dims[i] for i in right], dtype=object)) if right else 1
            k = int(np.prod([dims[i] for i in common], dtype=object)) if common else 1
            steps.append(dict(a=a, b=b, ia=nodes[a], ib=nodes[b], left=left, right=right, common=common, m=m, n=n, k=k))
            nodes.append(left + right)
        return steps

    def jet_flops(self) -> float:
        """PathInfo::GetTotalFlops (include/jet/PathInfo.hpp:157-199): sum of 2*M*N*K."""
        return float(sum(2 * s["m"] * s["n"] * s["k"] for s in self.path_steps()))


def amplitude(net: Network, sliced: Sequence[str] = (), slice_ids: Sequence[int] | None = None) -> np.ndarray:
    """Sum over slices of the contracted network, accumulated in complex128 (the reference reduces
    with AddTensors in an unspecified order, include/jet/TaskBasedContractor.hpp:258-280)."""
    if not sliced:
        return np.asarray(net.contract()[1], dtype=np.complex128)
    ids = range(net.num_slices(sliced)) if slice_ids is None else slice_ids
    acc = None
    for v in ids:
        r = np.asarray(net.slice_indices(sliced, v).contract()[1], dtype=np.complex128)
        acc = r if acc is None else acc + r
    return acc
