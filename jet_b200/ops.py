"""Host-side operator layer over the C ABI: numpy (host) buffers or raw device pointers in, results
out.  Mirrors the static operators of ``Jet::Tensor`` (reference include/jet/Tensor.hpp):
``Transpose`` (:579-612), ``ContractTensors`` (:709-752), ``AddTensors`` (:413-454),
``SliceIndex`` (:494-526) and ``MultiplyTensorData`` (include/jet/TensorHelpers.hpp:131-168).

Every function runs on the GPU through libjetb200.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import JB_C64, JB_C128, ChainDesc, ChainInfo, ContractInfo, check, lib


def dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.complex64:
        return JB_C64
    if dt == np.complex128:
        return JB_C128
    raise TypeError(f"unsupported dtype {dt}: Jet tensors are complex64 or complex128")


def _i64(seq):
    return (C.c_int64 * max(len(seq), 1))(*[int(v) for v in seq])


def _i32(seq):
    return (C.c_int32 * max(len(seq), 1))(*[int(v) for v in seq])


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
# host-buffer operators
# ---------------------------------------------------------------------------------------------
def permute(data: np.ndarray, shape: Sequence[int], perm: Sequence[int]) -> np.ndarray:
    """out axis j = in axis perm[j] (Permuter::Transpose, permute/Permuter.hpp:50-79)."""
    data = np.ascontiguousarray(data).reshape(-1)
    if int(np.prod(shape, dtype=np.int64)) != data.size:
        raise ValueError("Tensor shape does not match number of tensor elements.")
    out = np.empty_like(data)
    check(lib().jb_permute_host(dtype_code(data.dtype), _ptr(data), _ptr(out), len(shape), _i64(shape), _i32(perm)))
    return out


def contract_info(dtype, shape_a, modes_a, shape_b, modes_b) -> ContractInfo:
    info = ContractInfo()
    check(lib().jb_contract_info(dtype_code(dtype), len(shape_a), _i64(shape_a), _i32(modes_a), len(shape_b),
                                 _i64(shape_b), _i32(modes_b), C.byref(info)))
    return info


def contract(a: np.ndarray, modes_a: Sequence[int], b: np.ndarray, modes_b: Sequence[int]) -> Tuple[np.ndarray, list]:
    """ContractTensors(A, B): returns (C, modes_c) with C shaped left ++ right (Tensor.hpp:709-752)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.dtype != b.dtype:
        raise TypeError("operands must have the same dtype")
    info = contract_info(a.dtype, a.shape, modes_a, b.shape, modes_b)
    shape_c = [info.extent_c[i] for i in range(info.rank_c)]
    out = np.empty(int(info.m * info.n), dtype=a.dtype)
    check(lib().jb_contract_host(dtype_code(a.dtype), a.ndim, _i64(a.shape), _i32(modes_a), _ptr(a), b.ndim,
                                 _i64(b.shape), _i32(modes_b), _ptr(b), _ptr(out)))
    return out.reshape(shape_c), [info.modes_c[i] for i in range(info.rank_c)]


def _chain_desc(dtype, shape_x, modes_x, operands):
    """operands: list of (shape_r, modes_r, x_is_left)."""
    n = len(operands)
    keep = dict(
        ex=_i64(shape_x), mx=_i32(modes_x), rr=_i32([len(o[0]) for o in operands]),
        er=_i64([v for o in operands for v in o[0]]), mr=_i32([v for o in operands for v in o[1]]),
        xl=_i32([1 if o[2] else 0 for o in operands]))
    desc = ChainDesc(dtype_code(dtype), n, len(shape_x), keep["ex"], keep["mx"], keep["rr"], keep["er"], keep["mr"],
                     keep["xl"])
    return desc, keep


def chain_info(dtype, shape_x, modes_x, operands) -> ChainInfo:
    desc, _keep = _chain_desc(dtype, shape_x, modes_x, operands)
    info = ChainInfo()
    check(lib().jb_chain_info(C.byref(desc), C.byref(info)))
    return info


def contract_chain(x: np.ndarray, modes_x: Sequence[int], operands) -> Tuple[np.ndarray, list]:
    """A run of ContractTensors calls as ONE fused launch.  operands: list of (r, modes_r, x_is_left);
    step i computes ContractTensors(X, r) if x_is_left else ContractTensors(r, X) (Tensor.hpp:709-752)
    and feeds the result to step i+1.  Returns (X_k, modes of X_k)."""
    x = np.ascontiguousarray(x)
    rs = [np.ascontiguousarray(r, dtype=x.dtype) for r, _, _ in operands]
    spec = [(list(r.shape), list(m), bool(left)) for r, (_, m, left) in zip(rs, operands)]
    desc, _keep = _chain_desc(x.dtype, x.shape, modes_x, spec)
    info = ChainInfo()
    check(lib().jb_chain_info(C.byref(desc), C.byref(info)))
    shape_c = [info.extent_c[i] for i in range(info.rank_c)]
    out = np.empty(int(np.prod(shape_c, dtype=np.int64)) if shape_c else 1, dtype=x.dtype)
    rp = (C.c_void_p * len(rs))(*[r.ctypes.data for r in rs])
    check(lib().jb_contract_chain_host(C.byref(desc), _ptr(x), rp, _ptr(out)))
    return out.reshape(shape_c), [info.modes_c[i] for i in range(info.rank_c)]


def gemm(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Row-major C = A @ B, alpha=1, beta=0 (MultiplyTensorData, TensorHelpers.hpp:131-168)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    m, k = a.shape
    k2, n = b.shape
    if k != k2:
        raise ValueError("inner dimensions differ")
    out = np.empty((m, n), dtype=a.dtype)
    check(lib().jb_gemm_host(dtype_code(a.dtype), m, n, k, _ptr(a), _ptr(b), _ptr(out)))
    return out


def add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("operands must have the same shape and dtype")
    out = np.empty_like(a)
    check(lib().jb_add_host(dtype_code(a.dtype), a.size, _ptr(a), _ptr(b), _ptr(out)))
    return out


def slice_index(a: np.ndarray, axis: int, value: int) -> np.ndarray:
    a = np.ascontiguousarray(a)
    shape = list(a.shape)
    out = np.empty(shape[:axis] + shape[axis + 1:], dtype=a.dtype)
    check(lib().jb_slice_host(dtype_code(a.dtype), _ptr(a), _ptr(out), a.ndim, _i64(shape), axis, int(value)))
    return out


# ---------------------------------------------------------------------------------------------
# device-pointer operators (ints are raw CUDA device addresses, e.g. torch.Tensor.data_ptr())
# ---------------------------------------------------------------------------------------------
def permute_device(dtype, d_in: int, d_out: int, shape, perm, stream: int = 0):
    check(lib().jb_permute(dtype_code(dtype), d_in, d_out, len(shape), _i64(shape), _i32(perm), stream))


def gemm_ws_bytes(dtype, m, n, k) -> int:
    return int(lib().jb_gemm_ws_bytes(dtype_code(dtype), m, n, k))


def gemm_device(dtype, m, n, k, d_a: int, d_b: int, d_c: int, d_ws: int = 0, ws_bytes: int = 0, stream: int = 0):
    check(lib().jb_gemm(dtype_code(dtype), m, n, k, d_a, d_b, d_c, d_ws, ws_bytes, stream))


def contract_device(dtype, shape_a, modes_a, d_a: int, shape_b, modes_b, d_b: int, d_c: int, d_ws: int = 0,
                    ws_bytes: int = 0, stream: int = 0):
    check(lib().jb_contract(dtype_code(dtype), len(shape_a), _i64(shape_a), _i32(modes_a), d_a, len(shape_b),
                            _i64(shape_b), _i32(modes_b), d_b, d_c, d_ws, ws_bytes, stream))


def version() -> str:
    return lib().jb_version().decode()


def device_info(device: int = 0) -> dict:
    sm = C.c_int()
    tot = C.c_size_t()
    l2 = C.c_size_t()
    ma = C.c_int()
    mi = C.c_int()
    check(lib().jb_device_info(device, C.byref(sm), C.byref(tot), C.byref(l2), C.byref(ma), C.byref(mi)))
    return dict(sm_count=sm.value, total_bytes=tot.value, l2_bytes=l2.value, cc=(ma.value, mi.value))


PEAK_KINDS = {"fp32_fma": 0, "fp64_fma": 1, "fp64_dmma": 2, "tf32_mma_sync": 3, "hbm_copy": 4}


def probe_peak(kind: str) -> float:
    """Measured peak of the current device: TFLOP/s (fp32_fma, fp64_fma, fp64_dmma, tf32_mma_sync) or GB/s (hbm_copy)."""
    import ctypes as _C

    v = _C.c_double()
    check(lib().jb_probe_peak(PEAK_KINDS[kind], _C.byref(v)))
    return float(v.value)
