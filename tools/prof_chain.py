"""Launches the dominant fused-chain shape a few times (for ncu): a rank-R complex64 tensor absorbs `steps` rank-4 gate
tensors (K = N = 4) at random positions — the pattern of the m12 / m=20 chains.  python tools/prof_chain.py [R] [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ops  # noqa: E402
from jet_b200._lib import check, lib  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 27
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rng = np.random.default_rng(5)
dev = torch.device("cuda:0")
modes = list(range(R))
cur = list(modes)
nid = 1000
operands, gates = [], []
pool = [int(b) for b in rng.choice(modes[6:], size=8, replace=False)]  # the bits the gates act on (positions get reused)
for s in range(steps):
    sb = [int(b) for b in rng.choice(pool, size=2, replace=False)]
    f = [nid, nid + 1]
    nid += 2
    m = sb + f
    m = [m[i] for i in rng.permutation(4)]
    operands.append(([2] * 4, m, True))
    cur = [b for b in cur if b not in sb] + f
    pool = [b for b in pool if b not in sb] + f
desc, keep = ops._chain_desc(np.complex64, [2] * R, modes, operands)
info = ops.chain_info(np.complex64, [2] * R, modes, operands)
print("tile", info.log_tile, "stages", info.n_stages, "bytes %.3g flops %.3g" % (info.bytes, info.flops))
if "--info" in sys.argv:
    sys.exit(0)
x = torch.empty(2 ** R, dtype=torch.complex64, device=dev)
x.view(torch.float32).normal_()
out = torch.empty_like(x)
for s in range(steps):
    g = torch.empty(16, dtype=torch.complex64, device=dev)
    g.view(torch.float32).normal_()
    gates.append(g)
ptrs = (C.c_void_p * steps)(*[g.data_ptr() for g in gates])
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib().jb_contract_chain(C.byref(desc), C.c_void_p(x.data_ptr()), ptrs, C.c_void_p(out.data_ptr()), None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rep %d: %.3f ms  %.1f GB/s  %.1f TFLOP/s" % (rep, ms, info.bytes / ms / 1e6, info.flops / ms / 1e9))
