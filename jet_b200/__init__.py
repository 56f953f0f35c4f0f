"""jet_b200 — B200-native (sm_100a) engine for the contraction hot path of XanaduAI/jet.

Layers:
  csrc/          CUDA kernels + the C ABI (include/jetb200.h) -> lib/libjetb200.so
  _lib.py        ctypes loader (fails loudly if the library is missing; no CPU fallback)
  ops.py         operator layer: permute / contract / gemm / add / slice_index
  plan.py        NetworkFile (reference JSON format) + ContractionPlan (whole sliced network on a GPU)
                 + LanePlans (several slices in flight on one GPU)
"""
from ._lib import JetB200Error  # noqa: F401
from .plan import Communicator, ContractionPlan, LanePlans, MultiPlan, NetworkFile  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"
