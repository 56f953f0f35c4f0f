"""TEST INFRASTRUCTURE ONLY — ctypes access to oracle/_ref/libjetref.so (the unmodified reference
headers compiled by oracle/Makefile).  Used to validate oracle/jet_oracle.py, to generate the golden
fixtures under tests/golden/ and as the `"kind": "reference"` CPU baseline in bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libjetref.so")
DATA_DIR = os.path.join(os.path.dirname(HERE), "data", "_ref")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_blas_config.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())


def _dt(dtype):
    dtype = np.dtype(dtype)
    return 0 if dtype == np.complex64 else 1


def set_blas_threads(n: int):
    lib().ref_set_blas_threads(int(n))


def blas_config() -> str:
    return lib().ref_blas_config().decode()


def transpose(data: np.ndarray, shape, perm) -> np.ndarray:
    data = np.ascontiguousarray(data).reshape(-1)
    out = np.empty_like(data)
    shp = (C.c_int64 * len(shape))(*shape)
    pm = (C.c_int32 * len(perm))(*perm)
    _check(lib().ref_transpose(_dt(data.dtype), len(shape), shp, pm, data.ctypes.data_as(C.c_void_p),
                               out.ctypes.data_as(C.c_void_p)))
    return out


def contract(ia, a: np.ndarray, ib, b: np.ndarray) -> np.ndarray:
    """ia / ib: integer index ids; equal ids are contracted. Returns flat C (left ++ right)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    sa = (C.c_int64 * a.ndim)(*a.shape)
    sb = (C.c_int64 * b.ndim)(*b.shape)
    ca = (C.c_int32 * a.ndim)(*ia)
    cb = (C.c_int32 * b.ndim)(*ib)
    out = np.empty(a.size * b.size, dtype=a.dtype)
    n = C.c_int64(0)
    _check(lib().ref_contract(_dt(a.dtype), a.ndim, sa, ca, a.ctypes.data_as(C.c_void_p), b.ndim, sb, cb,
                              b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(n)))
    return out[: n.value].copy()


def slice_index(a: np.ndarray, axis: int, value: int) -> np.ndarray:
    a = np.ascontiguousarray(a)
    sa = (C.c_int64 * a.ndim)(*a.shape)
    out = np.empty(a.size // a.shape[axis], dtype=a.dtype)
    _check(lib().ref_slice_index(_dt(a.dtype), a.ndim, sa, axis, C.c_int64(value), a.ctypes.data_as(C.c_void_p),
                                 out.ctypes.data_as(C.c_void_p)))
    return out


def network(json_text: str, dtype="complex64", sliced=(), value=0, mode=0, threads=1, num_slices=1, cap=1 << 16):
    """Returns (result complex128 array, seconds of the timed contraction, Jet-convention flops)."""
    out = np.zeros(2 * cap, dtype=np.float64)
    n = C.c_int64(0)
    sec = C.c_double(0)
    fl = C.c_double(0)
    _check(lib().ref_network(_dt(dtype), json_text.encode(), " ".join(sliced).encode(), C.c_uint64(value), mode,
                             threads, C.c_uint64(num_slices), out.ctypes.data_as(C.c_void_p), C.c_int64(cap),
                             C.byref(n), C.byref(sec), C.byref(fl)))
    r = out[: 2 * n.value].view(np.complex128).copy()
    return r, sec.value, fl.value
