"""Launch the dominant kernels a few times on representative shapes (for ncu captures).
  python tools/prof_kernels.py [stream|stream16|permute|all]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
reps = 3


def stream(ra, rb, common):
    c = len(common)
    a = torch.empty(2 ** ra, dtype=torch.complex64, device=dev).normal_()
    b = torch.empty(2 ** rb, dtype=torch.complex64, device=dev).normal_()
    out = torch.empty(2 ** (ra + rb - 2 * c), dtype=torch.complex64, device=dev)
    ia = list(range(ra))
    ib = list(common) + list(range(100, 100 + rb - c))
    for _ in range(reps):
        ops.contract_device(np.complex64, [2] * ra, ia, a.data_ptr(), [2] * rb, ib, b.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()


if which in ("stream", "all"):
    stream(27, 4, [7, 22])      # the dominant m12 step shape: M=2^25, N=4, K=4
if which in ("stream16", "all"):
    stream(26, 8, [3, 9, 10, 23])  # K=N=16: the FMA-heavy m10 step
if which in ("permute", "all"):
    r = 27
    x = torch.empty(2 ** r, dtype=torch.complex64, device=dev).normal_()
    y = torch.empty_like(x)
    perm = np.random.default_rng(0).permutation(r).tolist()
    for _ in range(reps):
        ops.permute_device(np.complex64, x.data_ptr(), y.data_ptr(), [2] * r, perm)
    torch.cuda.synchronize()
if which in ("gemm_tc", "all_tc"):
    m = n = k = 4096
    a = torch.empty(m * k, dtype=torch.complex64, device=dev).normal_()
    b = torch.empty(k * n, dtype=torch.complex64, device=dev).normal_()
    c = torch.empty(m * n, dtype=torch.complex64, device=dev)
    wsb = ops.gemm_ws_bytes(np.complex64, m, n, k)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    for _ in range(reps):
        ops.gemm_device(np.complex64, m, n, k, a.data_ptr(), b.data_ptr(), c.data_ptr(), ws.data_ptr(), wsb)
    torch.cuda.synchronize()
if which in ("gemm_tc_skinny", "dot"):
    m, n, k = (512, 1024, 1 << 16) if which == "gemm_tc_skinny" else (1, 1, 1 << 26)
    a = torch.empty(m * k, dtype=torch.complex64, device=dev).normal_()
    b = torch.empty(k * n, dtype=torch.complex64, device=dev).normal_()
    c = torch.empty(m * n, dtype=torch.complex64, device=dev)
    wsb = ops.gemm_ws_bytes(np.complex64, m, n, k)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    for _ in range(reps):
        ops.gemm_device(np.complex64, m, n, k, a.data_ptr(), b.data_ptr(), c.data_ptr(), ws.data_ptr(), wsb)
    torch.cuda.synchronize()
