"""Network files and contraction plans (host side, Python).

``NetworkFile`` reads/writes the reference's JSON tensor-network format
(``TensorNetworkSerializer``, reference include/jet/TensorNetworkIO.hpp:94-187).
``ContractionPlan`` owns a ``jb_plan`` (include/jetb200.h): the whole sliced network resident on
one GPU, replacing per-slice ``TensorNetwork::SliceIndices`` copies + ``TaskBasedContractor`` tasks
(reference include/jet/TensorNetwork.hpp:210-284, include/jet/TaskBasedContractor.hpp:162-322).
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import (JB_PLAN_DRY_RUN, JB_PLAN_KEEP_INTERMEDIATES, JB_PLAN_NO_FUSE, JB_PLAN_NO_GRAPH, JB_PLAN_STORE_RESULTS, NetworkDesc,
                   OpInfo, PlanStats, StepInfo, check, lib)
from .ops import dtype_code


class NetworkFile:
    """Leaves (tags, indices, array) + optional path, as stored in the reference's JSON files."""

    def __init__(self, tensors: List[Tuple[List[str], np.ndarray]], path: Sequence[Sequence[int]] = (),
                 tags: Optional[List[List[str]]] = None):
        self.tensors = [(list(idx), np.ascontiguousarray(arr)) for idx, arr in tensors]
        self.path = [(int(a), int(b)) for a, b in path]
        self.tags = tags if tags is not None else [[] for _ in tensors]

    @classmethod
    def loads(cls, text: str, dtype=np.complex64) -> "NetworkFile":
        try:
            js = json.loads(text)
        except json.JSONDecodeError as e:
            raise ValueError(f"Error parsing tensor network file: {e}") from e
        if not isinstance(js, dict):
            raise ValueError("Error parsing tensor network file: root element must be an object.")
        if "tensors" not in js:
            raise ValueError("Error parsing tensor network file: root object must contain 'tensors' key.")
        tensors, tags = [], []
        for i, entry in enumerate(js["tensors"]):
            if len(entry) != 4:
                raise ValueError(f"Error parsing tensor network file: tensor {i} must have 4 fields.")
            tg, idx, shape, data = entry
            flat = np.asarray(data, dtype=np.float64)
            if flat.size and (flat.ndim != 2 or flat.shape[1] != 2):
                raise ValueError(f"Error parsing tensor network file: invalid complex data in tensor {i}.")
            arr = (flat[:, 0] + 1j * flat[:, 1]).astype(dtype) if flat.size else np.zeros(0, dtype)
            if int(np.prod(shape, dtype=np.int64)) != arr.size:
                raise ValueError(f"Error parsing tensor network file: tensor {i} has inconsistent shape.")
            tensors.append((list(idx), arr.reshape([int(s) for s in shape])))
            tags.append(list(tg))
        return cls(tensors, js.get("path", []), tags)

    @classmethod
    def load(cls, filename: str, dtype=np.complex64) -> "NetworkFile":
        with open(filename) as f:
            return cls.loads(f.read(), dtype)

    def dumps(self, indent=None) -> str:
        out: Dict[str, list] = {}
        if self.path:
            out["path"] = [list(p) for p in self.path]
        out["tensors"] = [
            [list(tg), list(idx), [int(s) for s in arr.shape],
             [[float(z.real), float(z.imag)] for z in arr.reshape(-1)]]
            for (idx, arr), tg in zip(self.tensors, self.tags)
        ]
        return json.dumps(out, indent=indent, separators=(",", ":") if indent is None else None)

    @property
    def dtype(self):
        return self.tensors[0][1].dtype

    def index_dims(self) -> Dict[str, int]:
        dims: Dict[str, int] = {}
        for idx, arr in self.tensors:
            for i, s in zip(idx, arr.shape):
                dims[i] = int(s)
        return dims


class ContractionPlan:
    """A sliced network + path resident on one GPU (wraps jb_plan)."""

    def __init__(self, net: NetworkFile, sliced: Sequence[str] = (), device: int = 0, keep_intermediates=False,
                 use_graph=True, store_results=False, path: Optional[Sequence[Sequence[int]]] = None,
                 fuse=True, dry_run=False, batch=0):
        self.net = net
        self.sliced = list(sliced)
        self.device = device
        self.dtype = np.dtype(net.dtype)
        labels: Dict[str, int] = {}
        for idx, _ in net.tensors:
            for i in idx:
                labels.setdefault(i, len(labels))
        self.labels = labels
        self.label_names = {v: k for k, v in labels.items()}
        for s in self.sliced:
            if s not in labels:
                raise ValueError("Sliced index does not exist.")
        ranks = [arr.ndim for _, arr in net.tensors]
        extents = [int(s) for _, arr in net.tensors for s in arr.shape]
        modes = [labels[i] for idx, _ in net.tensors for i in idx]
        self._leaves = [np.ascontiguousarray(arr, dtype=self.dtype) for _, arr in net.tensors]
        steps = [list(p) for p in (net.path if path is None else path)]
        flat_path = [v for p in steps for v in p]
        n = len(ranks)
        self._rank = (C.c_int32 * max(n, 1))(*ranks)
        self._extent = (C.c_int64 * max(len(extents), 1))(*extents)
        self._mode = (C.c_int32 * max(len(modes), 1))(*modes)
        self._data = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in self._leaves])
        self._path = (C.c_int32 * max(len(flat_path), 1))(*flat_path)
        self._sliced = (C.c_int32 * max(len(self.sliced), 1))(*[labels[s] for s in self.sliced])
        flags = (JB_PLAN_KEEP_INTERMEDIATES if keep_intermediates else 0) | (0 if use_graph else JB_PLAN_NO_GRAPH) | (
            JB_PLAN_STORE_RESULTS if store_results else 0) | (0 if fuse else JB_PLAN_NO_FUSE) | (
            JB_PLAN_DRY_RUN if dry_run else 0)
        desc = NetworkDesc(dtype_code(self.dtype), device, n, self._rank, self._extent, self._mode, self._data,
                           len(steps), self._path, len(self.sliced), self._sliced, flags, batch)
        self._h = C.c_void_p()
        check(lib().jb_plan_create(C.byref(desc), C.byref(self._h)))
        st = PlanStats()
        check(lib().jb_plan_stats(self._h, C.byref(st)))
        self.stats = st
        self.num_slices = int(st.num_slices)
        self.result_elems = int(st.result_elems)
        self.result_shape = [int(st.result_extent[i]) for i in range(st.result_rank)]
        self.result_indices = [self.label_names[st.result_modes[i]] for i in range(st.result_rank)]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().jb_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- execution ------------------------------------------------------------------------
    def upload(self, leaves: Optional[Sequence[np.ndarray]] = None):
        """(Re)copy leaf data host -> device; `leaves` may be pinned host buffers."""
        if leaves is not None:
            self._upload_ptrs = (C.c_void_p * len(leaves))(*[a.ctypes.data for a in leaves])
            self._upload_keep = leaves
            check(lib().jb_plan_upload(self._h, self._upload_ptrs))
        else:
            check(lib().jb_plan_upload(self._h, self._data))

    def upload_ptrs(self, ptrs: Sequence[int]):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        check(lib().jb_plan_upload(self._h, arr))

    def reset(self):
        check(lib().jb_plan_reset(self._h))

    def run(self, first: int = 0, count: Optional[int] = None):
        count = self.num_slices - first if count is None else count
        check(lib().jb_plan_run(self._h, first, count))

    def run_list(self, ids: Sequence[int]):
        arr = (C.c_int64 * max(len(ids), 1))(*[int(i) for i in ids])
        check(lib().jb_plan_run_list(self._h, arr, len(ids)))

    def sync(self):
        check(lib().jb_plan_sync(self._h))

    def result(self) -> np.ndarray:
        """Sum over the slices run since reset(), complex128, shaped like the final tensor."""
        out = np.empty(self.result_elems, dtype=np.complex128)
        check(lib().jb_plan_result(self._h, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(self.result_shape)

    def slice_result(self, ordinal: int) -> np.ndarray:
        out = np.empty(self.result_elems, dtype=self.dtype)
        check(lib().jb_plan_slice_result(self._h, ordinal, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(self.result_shape)

    def node(self, node: int) -> np.ndarray:
        n = C.c_int64()
        check(lib().jb_plan_node(self._h, node, None, C.byref(n)))
        out = np.empty(n.value, dtype=self.dtype)
        check(lib().jb_plan_node(self._h, node, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out

    def node_modes(self, node: int) -> List[int]:
        """Integer mode labels of node `node` as the steps see it (row-major axes, first slowest)."""
        rank = C.c_int32()
        modes = (C.c_int32 * 64)()
        check(lib().jb_plan_node_info(self._h, node, C.byref(rank), modes, None))
        return [int(modes[i]) for i in range(rank.value)]

    def node_shape(self, node: int) -> List[int]:
        rank = C.c_int32()
        ext = (C.c_int64 * 64)()
        check(lib().jb_plan_node_info(self._h, node, C.byref(rank), None, ext))
        return [int(ext[i]) for i in range(rank.value)]

    def last_ms(self) -> float:
        ms = C.c_float()
        check(lib().jb_plan_last_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().jb_plan_stream(self._h, C.byref(s)))
        return int(s.value or 0)

    def steps(self) -> List[StepInfo]:
        n = C.c_int32()
        check(lib().jb_plan_steps(self._h, None, 0, C.byref(n)))
        arr = (StepInfo * max(n.value, 1))()
        check(lib().jb_plan_steps(self._h, arr, n.value, C.byref(n)))
        return [arr[i] for i in range(n.value)]

    def profile(self, slice_id: int = 0, reps: int = 3) -> np.ndarray:
        n = len(self.steps())
        ms = np.zeros(max(n, 1), dtype=np.float32)
        check(lib().jb_plan_profile(self._h, slice_id, reps, ms.ctypes.data_as(C.c_void_p), n))
        return ms[:n]

    def ops(self) -> List[OpInfo]:
        """Launch units of one slice in execution order (single steps and fused chains)."""
        n = C.c_int32()
        check(lib().jb_plan_ops(self._h, None, 0, C.byref(n)))
        arr = (OpInfo * max(n.value, 1))()
        check(lib().jb_plan_ops(self._h, arr, n.value, C.byref(n)))
        return [arr[i] for i in range(n.value)]

    def profile_ops(self, slice_id: int = 0, reps: int = 3) -> np.ndarray:
        n = len(self.ops())
        ms = np.zeros(max(n, 1), dtype=np.float32)
        check(lib().jb_plan_profile_ops(self._h, slice_id, reps, ms.ctypes.data_as(C.c_void_p), n))
        return ms[:n]

    def amplitude(self, slice_ids: Optional[Sequence[int]] = None) -> np.ndarray:
        """reset + run (all slices, or the listed ones) + result."""
        self.reset()
        if slice_ids is None:
            self.run(0, self.num_slices)
        else:
            self.run_list(list(slice_ids))
        return self.result()


class _PlanView(ContractionPlan):
    """One plan of a MultiPlan, borrowed (profiling, streams): owned and destroyed by the set."""

    def __init__(self, owner: "MultiPlan", handle):  # noqa: D401 - not calling ContractionPlan.__init__
        self._h = handle
        self._owner = owner
        self.net, self.sliced, self.dtype = owner.net, owner.sliced, owner.dtype
        self.stats, self.num_slices = owner.stats, owner.num_slices
        self.result_elems, self.result_shape, self.result_indices = (owner.result_elems, owner.result_shape,
                                                                     owner.result_indices)

    def close(self):
        self._h = None


class MultiPlan:
    """One sliced network on `lanes` plans per device over any number of devices of this process (wraps
    jb_multi, include/jetb200.h).  Each plan has its own arena, stream and CUDA graphs, so `lanes` independent
    slice contractions are in flight per GPU — the reference runs the tasks of different slices concurrently on
    Taskflow worker threads (include/jet/TaskBasedContractor.hpp:322, examples/paper_benchmarks/CPU/
    jet_cpu_m10/jet_sliced.cpp:69-93); this is the same idea with streams.  Lanes pay off when one slice is too
    small to fill the GPU (m10, GBS fock4: launch-latency bound); large slices (m12, m=20) gain little.
    A run deals CONTIGUOUS blocks of its slice list to the plans in (device, lane) order — the same rule as the
    C++ Jet::SlicedContractor — and the FP64 partial sums are added on the devices in that order: the result is
    deterministic for a given (devices, lanes).  `reduce(comm)` finishes a one-process-per-GPU job with one NCCL
    reduce of the on-device total."""

    MAX_LANES = 5  # constant-bank slots available to plans (jet_b200/csrc/chain.cu)

    def __init__(self, net: NetworkFile, sliced: Sequence[str] = (), lanes: int = 2, device: int = 0,
                 devices: Optional[Sequence[int]] = None, keep_intermediates=False, use_graph=True,
                 store_results=False, path: Optional[Sequence[Sequence[int]]] = None, fuse=True, batch=0):
        if not 0 <= lanes <= self.MAX_LANES:
            raise ValueError(f"lanes must be in 0..{self.MAX_LANES} (0 = automatic)")
        self.net = net
        self.sliced = list(sliced)
        self.devices = [int(d) for d in (devices if devices is not None else [device])]
        self.dtype = np.dtype(net.dtype)
        labels: Dict[str, int] = {}
        for idx, _ in net.tensors:
            for i in idx:
                labels.setdefault(i, len(labels))
        self.label_names = {v: k for k, v in labels.items()}
        for s in self.sliced:
            if s not in labels:
                raise ValueError("Sliced index does not exist.")
        ranks = [arr.ndim for _, arr in net.tensors]
        extents = [int(s) for _, arr in net.tensors for s in arr.shape]
        modes = [labels[i] for idx, _ in net.tensors for i in idx]
        self._leaves = [np.ascontiguousarray(arr, dtype=self.dtype) for _, arr in net.tensors]
        steps = [list(p) for p in (net.path if path is None else path)]
        flat_path = [v for p in steps for v in p]
        n = len(ranks)
        self._rank = (C.c_int32 * max(n, 1))(*ranks)
        self._extent = (C.c_int64 * max(len(extents), 1))(*extents)
        self._mode = (C.c_int32 * max(len(modes), 1))(*modes)
        self._data = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in self._leaves])
        self._path = (C.c_int32 * max(len(flat_path), 1))(*flat_path)
        self._sliced = (C.c_int32 * max(len(self.sliced), 1))(*[labels[s] for s in self.sliced])
        flags = (JB_PLAN_KEEP_INTERMEDIATES if keep_intermediates else 0) | (0 if use_graph else JB_PLAN_NO_GRAPH) | (
            JB_PLAN_STORE_RESULTS if store_results else 0) | (0 if fuse else JB_PLAN_NO_FUSE)
        desc = NetworkDesc(dtype_code(self.dtype), self.devices[0], n, self._rank, self._extent, self._mode, self._data,
                           len(steps), self._path, len(self.sliced), self._sliced, flags, batch)
        devs = (C.c_int * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        check(lib().jb_multi_create(C.byref(desc), len(self.devices), devs, lanes, C.byref(self._h)))
        st = PlanStats()
        check(lib().jb_multi_stats(self._h, C.byref(st)))
        nd, nl = C.c_int(), C.c_int()
        check(lib().jb_multi_num_plans(self._h, C.byref(nd), C.byref(nl)))
        self.lanes = int(nl.value)
        self.stats = st
        self.num_slices = int(st.num_slices)
        self.result_elems = int(st.result_elems)
        self.result_shape = [int(st.result_extent[i]) for i in range(st.result_rank)]
        self.result_indices = [self.label_names[st.result_modes[i]] for i in range(st.result_rank)]
        self.plans: List[_PlanView] = []
        for i in range(nd.value * nl.value):
            h = C.c_void_p()
            check(lib().jb_multi_plan(self._h, i, C.byref(h)))
            self.plans.append(_PlanView(self, h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            for p in self.plans:
                p.close()
            lib().jb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def streams(self) -> List[int]:
        return [p.stream() for p in self.plans]

    def upload_ptrs(self, ptrs: Sequence[int]):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        check(lib().jb_multi_upload(self._h, arr))

    def reset(self):
        check(lib().jb_multi_reset(self._h))

    def run_list(self, ids: Sequence[int]):
        """reset + enqueue the listed slice ids (asynchronous)."""
        arr = (C.c_int64 * max(len(ids), 1))(*[int(i) for i in ids])
        check(lib().jb_multi_run_list(self._h, arr, len(ids)))

    def run(self, first: int = 0, count: Optional[int] = None):
        """reset + enqueue slices first .. first + count - 1 (asynchronous)."""
        count = self.num_slices - first if count is None else count
        check(lib().jb_multi_run(self._h, first, count))

    def sync(self):
        check(lib().jb_multi_sync(self._h))

    def reduce(self, comm: "Communicator", root: int = 0):
        """Sum the set's on-device total over the ranks of `comm` (NCCL, on the first plan's stream)."""
        check(lib().jb_multi_reduce(self._h, comm._h, root))

    def result(self) -> np.ndarray:
        out = np.empty(self.result_elems, dtype=np.complex128)
        check(lib().jb_multi_result(self._h, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(self.result_shape)

    def slice_results(self, count: int) -> np.ndarray:
        out = np.empty((count, self.result_elems), dtype=self.dtype)
        check(lib().jb_multi_slice_results(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def last_ms(self) -> float:
        ms = C.c_float()
        check(lib().jb_multi_last_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def amplitude(self, slice_ids: Optional[Sequence[int]] = None) -> np.ndarray:
        if slice_ids is None:
            self.run(0, self.num_slices)
        else:
            self.run_list(list(slice_ids))
        return self.result()


LanePlans = MultiPlan  # the round-1 name (one device, several lanes)


class Communicator:
    """NCCL communicator of a one-process-per-GPU job (wraps jb_comm).  `exchange(id_bytes_or_None) -> bytes`
    must hand rank 0's 128-byte id to every rank (e.g. a torch.distributed / MPI broadcast)."""

    def __init__(self, world: int, rank: int, device: int, exchange):
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(lib().jb_comm_unique_id(buf))
        ident = exchange(bytes(buf.raw) if rank == 0 else None)
        assert len(ident) == 128
        self._h = C.c_void_p()
        check(lib().jb_comm_create(world, rank, C.c_char_p(ident), device, C.byref(self._h)))
        self.world, self.rank, self.device = world, rank, device

    def nccl_version(self) -> int:
        v = C.c_int()
        check(lib().jb_comm_info(self._h, None, None, C.byref(v)))
        return int(v.value)

    def reduce_sum(self, d_ptr: int, n_doubles: int, root: int = 0, stream: int = 0):
        check(lib().jb_reduce_sum(self._h, C.c_void_p(d_ptr), n_doubles, root, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().jb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
