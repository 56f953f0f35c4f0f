"""Random pairwise contractions on the GPU against numpy (complex128 reference), sized to reach every
kernel family: stream, TTGT + FMA GEMM, tcgen05 (plain / ragged / split-K / role swap), DMMA (plain / gather),
SmallMn.   python tools/stress_contract_gpu.py [trials]"""
import os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ops  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
TOL = {np.complex64: 1e-5, np.complex128: 1e-12}
rng = np.random.default_rng(2024)
kinds = collections.Counter()
bad = 0
for trial in range(trials):
    dtype = np.complex64 if rng.integers(0, 2) else np.complex128
    dims = int(rng.choice([2, 2, 2, 4, 8]))
    lg = {2: 1, 4: 2, 8: 3}[dims]
    max_bits = 22 if dtype == np.complex64 else 21
    ra = int(rng.integers(1, max_bits // lg + 1))
    rb = int(rng.integers(1, max_bits // lg + 1))
    nc = int(rng.integers(0, min(ra, rb) + 1))
    if (ra + rb - 2 * nc) * lg > 24:
        continue
    ia = [int(v) for v in rng.permutation(ra)]
    common = [int(v) for v in rng.choice(ia, nc, replace=False)] if nc else []
    ib = common + list(range(100, 100 + rb - nc))
    ib = [ib[i] for i in rng.permutation(rb)]
    real = np.float32 if dtype == np.complex64 else np.float64
    a = (rng.uniform(-1, 1, dims ** ra).astype(real) + 1j * rng.uniform(-1, 1, dims ** ra).astype(real)).astype(dtype).reshape([dims] * ra)
    b = (rng.uniform(-1, 1, dims ** rb).astype(real) + 1j * rng.uniform(-1, 1, dims ** rb).astype(real)).astype(dtype).reshape([dims] * rb)
    info = ops.contract_info(dtype, a.shape, ia, b.shape, ib)
    out, modes = ops.contract(a, ia, b, ib)
    ax_a = [ia.index(c) for c in common]
    ax_b = [ib.index(c) for c in common]
    ref = np.tensordot(a.astype(np.complex128), b.astype(np.complex128), axes=(ax_a, ax_b))
    want_modes = [m for m in ia if m not in common] + [m for m in ib if m not in common]
    err = np.linalg.norm(out.reshape(-1) - ref.reshape(-1)) / max(np.linalg.norm(ref.reshape(-1)), 1e-300)
    kinds[(dtype.__name__, int(info.kernel), "big" if info.m * info.n * info.k >= 1 << 24 else "small")] += 1
    if list(modes) != want_modes or err >= TOL[dtype]:
        bad += 1
        print("FAIL", trial, dtype.__name__, dims, ia, ib, info.m, info.n, info.k, err)
print(dict(kinds))
print("bad", bad)
