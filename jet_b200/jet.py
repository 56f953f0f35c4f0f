"""`jet`-compatible Python API: the names of the reference's ``jet`` package for the hot path
(reference python/jet/__init__.py:6-21 and python/jet/factory.py:40-151), backed by the B200
engine.  Usage::

    from jet_b200 import jet
    tn = jet.TensorNetwork(dtype=np.complex64)
    tbc = jet.TaskBasedContractor(dtype=np.complex64)

As in the reference the default ``dtype`` is complex128.  The gate / state / circuit classes of the reference package
(python/jet/{gate,state,circuit}.py) are re-exported from ``jet_b200.gate`` / ``state`` / ``circuit``; the XIR
interpreter (python/jet/interpreter.py: ``get_xir_manifest``, ``run_xir_program``) runs on programs read by
``jet_b200.xir_lite.parse_script`` — the ``xir`` package is absent here — and contracts through ``jet_b200.simulate``
(``compute_amplitude`` / ``compute_probabilities`` / ``compute_expected_value``).
"""
from typing import Union

import numpy as np

from .bindings import (  # noqa: F401
    PathInfo,
    PathStepInfo,
    SlicedContractorC64,
    SlicedContractorC128,
    TaskBasedContractorC64,
    TaskBasedContractorC128,
    TensorC64,
    TensorC128,
    TensorNetworkC64,
    TensorNetworkC128,
    TensorNetworkFileC64,
    TensorNetworkFileC128,
    TensorNetworkSerializerC64,
    TensorNetworkSerializerC128,
    add_tensors,
    conj,
    contract_tensors,
    reshape,
    slice_index,
    transpose,
    version,
)

from .circuit import *  # noqa: F401,F403,E402
from .circuit import __all__ as _circuit_all  # noqa: E402
from .gate import *  # noqa: F401,F403,E402
from .gate import __all__ as _gate_all  # noqa: E402
from .state import *  # noqa: F401,F403,E402
from .state import __all__ as _state_all  # noqa: E402
from .simulate import *  # noqa: F401,F403,E402
from .simulate import __all__ as _simulate_all  # noqa: E402
from .interpreter import *  # noqa: F401,F403,E402
from .interpreter import __all__ as _interpreter_all  # noqa: E402

__all__ = [
    "PathInfo", "PathStepInfo", "add_tensors", "conj", "contract_tensors", "reshape", "slice_index", "transpose",
    "version", "TaskBasedContractorType", "TensorType", "TensorNetworkType", "TensorNetworkFileType",
    "TensorNetworkSerializerType", "TaskBasedContractor", "Tensor", "TensorNetwork", "TensorNetworkFile",
    "TensorNetworkSerializer", "SlicedContractor",
] + _circuit_all + _gate_all + _state_all + _simulate_all + _interpreter_all

TaskBasedContractorType = Union[TaskBasedContractorC64, TaskBasedContractorC128]
TensorType = Union[TensorC64, TensorC128]
TensorNetworkType = Union[TensorNetworkC64, TensorNetworkC128]
TensorNetworkFileType = Union[TensorNetworkFileC64, TensorNetworkFileC128]
TensorNetworkSerializerType = Union[TensorNetworkSerializerC64, TensorNetworkSerializerC128]


def _pick(c64, c128, args, kwargs):
    dtype = np.dtype(kwargs.pop("dtype", np.complex128))
    if dtype == np.complex64:
        return c64(*args, **kwargs)
    if dtype == np.complex128:
        return c128(*args, **kwargs)
    raise TypeError(f"Data type '{dtype}' is not supported.")


def TaskBasedContractor(*args, **kwargs) -> TaskBasedContractorType:
    return _pick(TaskBasedContractorC64, TaskBasedContractorC128, args, kwargs)


def Tensor(*args, **kwargs) -> TensorType:
    return _pick(TensorC64, TensorC128, args, kwargs)


def TensorNetwork(*args, **kwargs) -> TensorNetworkType:
    return _pick(TensorNetworkC64, TensorNetworkC128, args, kwargs)


def TensorNetworkFile(*args, **kwargs) -> TensorNetworkFileType:
    return _pick(TensorNetworkFileC64, TensorNetworkFileC128, args, kwargs)


def TensorNetworkSerializer(*args, **kwargs) -> TensorNetworkSerializerType:
    return _pick(TensorNetworkSerializerC64, TensorNetworkSerializerC128, args, kwargs)


def SlicedContractor(*args, **kwargs):
    """B200 extension: device-resident contraction of a sliced network (jb_plan)."""
    return _pick(SlicedContractorC64, SlicedContractorC128, args, kwargs)


__version__ = version()
