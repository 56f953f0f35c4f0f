"""Random fused chains on the GPU against the step-by-step numpy oracle (more trials than the test suite).
  python tools/stress_chain_gpu.py [trials]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_chain_gpu as t  # noqa: E402
from jet_b200 import JetB200Error, ops  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ran = bad = 0
for dtype in (np.complex64, np.complex128):
    for dims in (2, 4):
        rng = np.random.default_rng(777 + dims + (0 if dtype == np.complex64 else 50))
        for trial in range(trials):
            rank = int(rng.integers(2, 21 if dims == 2 else 10))
            modes, spec = t.random_chain(rng, rank, int(rng.integers(1, 9)), dims, max_c=3 if dims == 2 else 1,
                                         max_f=3 if dims == 2 else 1)
            x = t.rand_c(rng, [dims] * rank, dtype)
            operands = [(t.rand_c(rng, [dims] * len(mr), dtype), mr, left) for mr, left in spec]
            try:
                got, modes_c = ops.contract_chain(x, modes, operands)
            except JetB200Error as e:
                assert "chain:" in str(e)
                continue
            want_modes, want = t.oracle_chain(x, modes, operands)
            ran += 1
            if [str(m) for m in modes_c] != want_modes or t.rel_err(got, want) >= t.TOL[np.dtype(dtype)]:
                bad += 1
                print("FAIL", dtype.__name__, dims, trial, rank, spec, t.rel_err(got, want))
print("ran", ran, "bad", bad)
