"""Random GEMM shapes (jb_gemm) against numpy complex128: ragged tiles of the tcgen05 / DMMA kernels, split-K,
the SmallMn corner and the FMA fallback.   python tools/stress_gemm_gpu.py [trials] [seed]"""
import os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ops  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
vals = [1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 33, 48, 63, 64, 65, 96, 127, 128, 129, 192, 255, 256, 320, 511, 512, 1000, 1024,
        2048, 4096, 1 << 13, 1 << 14, 1 << 16, 1 << 18]
bad = 0
for t in range(trials):
    dtype = np.complex64 if rng.integers(0, 2) else np.complex128
    m, n, k = (int(rng.choice(vals)) for _ in range(3))
    if m * k > 1 << 24 or k * n > 1 << 24 or m * n > 1 << 24 or m * n * k > 1 << 34:
        continue
    real = np.float32 if dtype == np.complex64 else np.float64
    if rng.integers(0, 3) == 0:  # integer data: exact
        a = (rng.integers(-2, 3, (m, k)) + 1j * rng.integers(-2, 3, (m, k))).astype(dtype)
        b = (rng.integers(-2, 3, (k, n)) + 1j * rng.integers(-2, 3, (k, n))).astype(dtype)
        exact = k * 32 < (1 << 23 if dtype == np.complex64 else 1 << 50)
    else:
        a = (rng.uniform(-1, 1, (m, k)).astype(real) + 1j * rng.uniform(-1, 1, (m, k)).astype(real)).astype(dtype)
        b = (rng.uniform(-1, 1, (k, n)).astype(real) + 1j * rng.uniform(-1, 1, (k, n)).astype(real)).astype(dtype)
        exact = False
    c = ops.gemm(a, b)
    ref = a.astype(np.complex128) @ b.astype(np.complex128)
    err = np.linalg.norm(c - ref) / max(np.linalg.norm(ref), 1e-300)
    ok = np.array_equal(c, ref.astype(dtype)) if exact else err < (1e-5 if dtype == np.complex64 else 1e-12)
    if not ok:
        bad += 1
        print("FAIL", dtype.__name__, m, n, k, "exact" if exact else "", err)
print("bad", bad)
