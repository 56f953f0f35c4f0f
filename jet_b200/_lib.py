"""ctypes binding of libjetb200.so (the C ABI in include/jetb200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, the calls
raise.  Build the library with ``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C jet_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JETB200_LIB") or os.path.join(HERE, "lib", "libjetb200.so")

JB_C64, JB_C128 = 0, 1
JB_MAX_RANK = 64
JB_PLAN_KEEP_INTERMEDIATES = 1
JB_PLAN_NO_GRAPH = 2
JB_PLAN_STORE_RESULTS = 4
JB_PLAN_NO_FUSE = 8
JB_PLAN_DRY_RUN = 16


class JetB200Error(RuntimeError):
    """Raised for every non-zero status of the C ABI (message from jb_last_error())."""


class ContractInfo(C.Structure):
    _fields_ = [
        ("rank_c", C.c_int32),
        ("modes_c", C.c_int32 * JB_MAX_RANK),
        ("extent_c", C.c_int64 * JB_MAX_RANK),
        ("m", C.c_int64),
        ("n", C.c_int64),
        ("k", C.c_int64),
        ("ws_bytes", C.c_size_t),
        ("kernel", C.c_int32),
    ]


class NetworkDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("device", C.c_int32),
        ("num_leaves", C.c_int32),
        ("rank", C.POINTER(C.c_int32)),
        ("extent", C.POINTER(C.c_int64)),
        ("mode", C.POINTER(C.c_int32)),
        ("h_data", C.POINTER(C.c_void_p)),
        ("num_steps", C.c_int32),
        ("path", C.POINTER(C.c_int32)),
        ("num_sliced", C.c_int32),
        ("sliced_modes", C.POINTER(C.c_int32)),
        ("flags", C.c_int32),
        ("batch", C.c_int32),
    ]


class PlanStats(C.Structure):
    _fields_ = [
        ("num_slices", C.c_int64),
        ("result_elems", C.c_int64),
        ("result_rank", C.c_int32),
        ("result_modes", C.c_int32 * JB_MAX_RANK),
        ("result_extent", C.c_int64 * JB_MAX_RANK),
        ("steps_total", C.c_int32),
        ("steps_shared", C.c_int32),
        ("steps_stream", C.c_int32),
        ("steps_ttgt", C.c_int32),
        ("steps_chained", C.c_int32),
        ("chains", C.c_int32),
        ("launches_per_slice", C.c_int32),
        ("flops_per_slice", C.c_double),
        ("bytes_per_slice", C.c_double),
        ("fused_bytes_per_slice", C.c_double),
        ("flops_shared", C.c_double),
        ("bytes_shared", C.c_double),
        ("jet_flops_per_slice", C.c_double),
        ("arena_bytes", C.c_size_t),
        ("max_step_elems", C.c_int64),
        ("batch", C.c_int32),
        ("pad", C.c_int32),
    ]


class StepInfo(C.Structure):
    _fields_ = [
        ("node_a", C.c_int32),
        ("node_b", C.c_int32),
        ("node_c", C.c_int32),
        ("shared", C.c_int32),
        ("kernel", C.c_int32),
        ("m", C.c_int64),
        ("n", C.c_int64),
        ("k", C.c_int64),
        ("flops", C.c_double),
        ("bytes", C.c_double),
        ("op", C.c_int32),
        ("pad", C.c_int32),
    ]


class OpInfo(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32),
        ("n_steps", C.c_int32),
        ("first_step", C.c_int32),
        ("last_step", C.c_int32),
        ("log_tile", C.c_int32),
        ("launches", C.c_int32),
        ("n_stages", C.c_int32),
        ("gemm_kind", C.c_int32),
        ("flops", C.c_double),
        ("bytes", C.c_double),
        ("step_bytes", C.c_double),
        ("register_steps", C.c_int32),
        ("pad", C.c_int32),
    ]


GEMM_KIND_NAMES = {0: "GemmKernel", 1: "SmallMnKernel", 2: "GemmTf32x3Kernel", 3: "GemmDmmaKernel", 4: "DotGatherKernel",
                   5: "SmallGemmGatherKernel"}


class ChainDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("n_steps", C.c_int32),
        ("rank_x", C.c_int32),
        ("extent_x", C.POINTER(C.c_int64)),
        ("modes_x", C.POINTER(C.c_int32)),
        ("rank_r", C.POINTER(C.c_int32)),
        ("extent_r", C.POINTER(C.c_int64)),
        ("modes_r", C.POINTER(C.c_int32)),
        ("x_is_left", C.POINTER(C.c_int32)),
    ]


class ChainInfo(C.Structure):
    _fields_ = [
        ("rank_c", C.c_int32),
        ("modes_c", C.c_int32 * JB_MAX_RANK),
        ("extent_c", C.c_int64 * JB_MAX_RANK),
        ("log_tile", C.c_int32),
        ("conflict_free", C.c_int32),
        ("n_stages", C.c_int32),
        ("pad", C.c_int32),
        ("flops", C.c_double),
        ("bytes", C.c_double),
        ("step_bytes", C.c_double),
    ]


# every symbol include/jetb200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "jb_last_error", "jb_version", "jb_device_count", "jb_set_device", "jb_device_info",
    "jb_malloc", "jb_free", "jb_host_alloc", "jb_host_free", "jb_memcpy_h2d", "jb_memcpy_d2h",
    "jb_memcpy_d2d", "jb_memset_zero", "jb_stream_create", "jb_stream_destroy", "jb_stream_sync",
    "jb_permute", "jb_gemm_ws_bytes", "jb_gemm", "jb_contract_info", "jb_contract", "jb_add",
    "jb_slice", "jb_conj", "jb_permute_host", "jb_contract_host", "jb_gemm_host", "jb_add_host",
    "jb_slice_host", "jb_conj_host", "jb_plan_create", "jb_plan_destroy", "jb_plan_stats", "jb_plan_upload",
    "jb_plan_reset", "jb_plan_run", "jb_plan_run_list", "jb_plan_result", "jb_plan_slice_result",
    "jb_plan_node", "jb_plan_sync", "jb_plan_last_ms", "jb_plan_stream", "jb_plan_steps",
    "jb_plan_profile", "jb_plan_ops", "jb_plan_profile_ops", "jb_chain_info", "jb_contract_chain",
    "jb_contract_chain_host", "jb_plan_clone", "jb_plan_accumulator", "jb_plan_device", "jb_plan_slice_results",
    "jb_multi_create", "jb_multi_destroy", "jb_multi_stats", "jb_multi_num_plans", "jb_multi_plan", "jb_multi_upload",
    "jb_multi_reset", "jb_multi_run", "jb_multi_run_list", "jb_multi_sync", "jb_multi_result",
    "jb_multi_slice_result", "jb_multi_slice_results", "jb_multi_last_ms", "jb_comm_unique_id", "jb_comm_create",
    "jb_comm_destroy", "jb_comm_info", "jb_reduce_sum", "jb_multi_reduce", "jb_probe_peak", "jb_plan_node_info",
]

_lib = None


def lib():
    """Load libjetb200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JetB200Error(
                f"{LIB_PATH} is missing: build it with `make -C jet_b200/csrc` "
                "(there is no CPU fallback)"
            )
        L = C.CDLL(LIB_PATH)
        L.jb_last_error.restype = C.c_char_p
        L.jb_version.restype = C.c_char_p
        L.jb_gemm_ws_bytes.restype = C.c_size_t
        L.jb_gemm_ws_bytes.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64]
        L.jb_gemm.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.jb_gemm_host.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p]
        L.jb_permute.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.jb_permute_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.jb_contract.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]
        L.jb_contract_host.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.jb_contract_info.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.POINTER(ContractInfo)]
        L.jb_add.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.jb_add_host.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.jb_conj.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.jb_slice.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_int64, C.c_void_p]
        L.jb_slice_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int64]
        L.jb_conj_host.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.jb_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.jb_free.argtypes = [C.c_void_p]
        L.jb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.jb_host_free.argtypes = [C.c_void_p]
        for name in ("jb_memcpy_h2d", "jb_memcpy_d2h", "jb_memcpy_d2d"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.jb_memset_zero.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.jb_stream_create.argtypes = [C.POINTER(C.c_void_p)]
        L.jb_stream_destroy.argtypes = [C.c_void_p]
        L.jb_stream_sync.argtypes = [C.c_void_p]
        L.jb_plan_create.argtypes = [C.POINTER(NetworkDesc), C.POINTER(C.c_void_p)]
        L.jb_plan_destroy.argtypes = [C.c_void_p]
        L.jb_plan_stats.argtypes = [C.c_void_p, C.POINTER(PlanStats)]
        L.jb_plan_upload.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.jb_plan_reset.argtypes = [C.c_void_p]
        L.jb_plan_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.jb_plan_run_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.jb_plan_result.argtypes = [C.c_void_p, C.c_void_p]
        L.jb_plan_slice_result.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.jb_plan_node.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]
        L.jb_plan_sync.argtypes = [C.c_void_p]
        L.jb_plan_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.jb_plan_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.jb_plan_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.jb_plan_profile.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int32]
        L.jb_plan_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.jb_plan_profile_ops.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int32]
        L.jb_chain_info.argtypes = [C.POINTER(ChainDesc), C.POINTER(ChainInfo)]
        L.jb_contract_chain.argtypes = [C.POINTER(ChainDesc), C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p,
                                        C.c_void_p]
        L.jb_contract_chain_host.argtypes = [C.POINTER(ChainDesc), C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
        L.jb_plan_clone.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.jb_plan_accumulator.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.jb_plan_device.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.jb_plan_slice_results.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.jb_multi_create.argtypes = [C.POINTER(NetworkDesc), C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.jb_multi_destroy.argtypes = [C.c_void_p]
        L.jb_multi_stats.argtypes = [C.c_void_p, C.POINTER(PlanStats)]
        L.jb_multi_num_plans.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.jb_multi_plan.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.jb_multi_upload.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.jb_multi_reset.argtypes = [C.c_void_p]
        L.jb_multi_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.jb_multi_run_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.jb_multi_sync.argtypes = [C.c_void_p]
        L.jb_multi_result.argtypes = [C.c_void_p, C.c_void_p]
        L.jb_multi_slice_result.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.jb_multi_slice_results.argtypes = [C.c_void_p, C.c_void_p]
        L.jb_multi_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.jb_comm_unique_id.argtypes = [C.c_void_p]
        L.jb_comm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.jb_comm_destroy.argtypes = [C.c_void_p]
        L.jb_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.jb_reduce_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.jb_multi_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.jb_probe_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.jb_plan_node_info.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise JetB200Error(lib().jb_last_error().decode(errors="replace"))
