"""How fast does the error of a 3xTF32 tensor-core product grow when it is applied repeatedly (the fused chains apply
~40 small unitary-like matrices to a tensor per slice)?  X_{s+1} = X_s U_s with random unitary U_s (64 x 64 complex),
through ops.gemm (GemmTf32x3Kernel for complex64, FFMA GemmKernel when JB_DISABLE_TC=1), against complex128 numpy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ops  # noqa: E402

rng = np.random.default_rng(1)
m, n = 1 << 15, 64
x = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))).astype(np.complex64)
ref = x.astype(np.complex128)
cur = x.copy()
for s in range(1, 49):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    u = q.astype(np.complex64)
    ref = ref @ u.astype(np.complex128)
    cur = ops.gemm(cur, u)
    if s in (1, 2, 4, 8, 16, 24, 32, 40, 48):
        err = np.linalg.norm(cur - ref) / np.linalg.norm(ref)
        shrink = np.linalg.norm(cur) / np.linalg.norm(ref) - 1
        print("stages %2d: normwise error %.3e   norm drift %+.3e" % (s, err, shrink), flush=True)
