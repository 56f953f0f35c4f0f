"""GPU parity of the fused contraction chain (K3, jet_b200/csrc/chain.cu) through the C ABI.

A chain must equal the run of Tensor::ContractTensors calls it replaces (reference
include/jet/Tensor.hpp:709-752): same labels, same extents, values within the BASELINE tolerance
(1e-5 complex64 / 1e-12 complex128, normwise) of the oracle contracting step by step; at plan
level the fused engine must agree with the unfused one and with the reference goldens."""
import os

import numpy as np
import pytest

from oracle import jet_oracle as jo

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.complex64): 1e-5, np.dtype(np.complex128): 1e-12}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rand_c(rng, shape, dtype):
    n = int(np.prod(shape)) if len(shape) else 1
    real = np.float32 if dtype == np.complex64 else np.float64
    return (rng.uniform(-1, 1, n).astype(real) + 1j * rng.uniform(-1, 1, n).astype(real)).astype(dtype).reshape(shape)


def rel_err(x, ref):
    x = np.asarray(x, dtype=np.complex128).reshape(-1)
    ref = np.asarray(ref, dtype=np.complex128).reshape(-1)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300))


def oracle_chain(x, modes_x, operands):
    cur = ([str(m) for m in modes_x], x)
    for r, modes_r, left in operands:
        other = ([str(m) for m in modes_r], r)
        cur = jo.contract(cur, other) if left else jo.contract(other, cur)
    return cur


def random_chain(rng, rank, n_steps, dims=2, max_c=2, max_f=2):
    nxt = [1000]
    modes = list(rng.permutation(rank))
    modes = [int(m) for m in modes]
    cur = list(modes)
    ops = []
    for _ in range(n_steps):
        c = int(rng.integers(0, min(max_c, len(cur)) + 1))
        f = int(rng.integers(0 if c else 1, max_f + 1))
        common = [int(m) for m in rng.choice(cur, size=c, replace=False)] if c else []
        fresh = list(range(nxt[0], nxt[0] + f))
        nxt[0] += f
        mr = common + fresh
        rng.shuffle(mr)
        left = bool(rng.integers(0, 2))
        ops.append((mr, left))
        rest = [m for m in cur if m not in common]
        fr = [m for m in mr if m in fresh]
        cur = rest + fr if left else fr + rest
    return modes, ops


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("dims", [2, 4])
def test_random_chains_vs_stepwise_oracle(dtype, dims):
    from jet_b200 import JetB200Error, ops
    rng = np.random.default_rng(100 + dims)
    ran = 0
    for trial in range(40):
        rank = int(rng.integers(2, 19 if dims == 2 else 9))
        modes, spec = random_chain(rng, rank, int(rng.integers(1, 7)), dims, max_c=2 if dims == 2 else 1,
                                   max_f=2 if dims == 2 else 1)
        x = rand_c(rng, [dims] * rank, dtype)
        operands = [(rand_c(rng, [dims] * len(mr), dtype), mr, left) for mr, left in spec]
        try:
            got, modes_c = ops.contract_chain(x, modes, operands)
        except JetB200Error as e:
            assert "chain:" in str(e)
            continue
        want_modes, want = oracle_chain(x, modes, operands)
        assert [str(m) for m in modes_c] == want_modes
        assert list(got.shape) == list(np.shape(want))
        assert rel_err(got, want) < TOL[np.dtype(dtype)], (trial, rank, spec)
        ran += 1
    assert ran >= 20


def test_gate_chain_large_exact_integers():
    """Full-size property: a rank-24 tensor absorbs five rank-4 tensors; small-integer data makes the
    FP32 result exact, so the fused launch must equal the step-by-step launches bit for bit."""
    from jet_b200 import ops
    rng = np.random.default_rng(3)
    rank = 24
    modes = list(range(rank))
    x = (rng.integers(-2, 3, 2 ** rank) + 1j * rng.integers(-2, 3, 2 ** rank)).astype(np.complex64).reshape([2] * rank)
    cur = list(modes)
    operands = []
    nid = 100
    for s in range(5):
        common = [int(m) for m in rng.choice(cur, size=2, replace=False)]
        mr = common + [nid, nid + 1]
        rng.shuffle(mr)
        r = (rng.integers(-1, 2, 16) + 1j * rng.integers(-1, 2, 16)).astype(np.complex64).reshape([2] * 4)
        operands.append((r, mr, True))
        cur = [m for m in cur if m not in common] + [m for m in mr if m >= 100 and m in (nid, nid + 1)]
        nid += 2
    got, modes_c = ops.contract_chain(x, modes, operands)
    step, step_modes = x, modes
    for r, mr, _ in operands:
        step, step_modes = ops.contract(step, step_modes, r, mr)
    assert modes_c == step_modes
    assert np.array_equal(got, step)
    info = ops.chain_info(np.complex64, x.shape, modes, [(list(r.shape), mr, True) for r, mr, _ in operands])
    assert info.conflict_free == 1 and info.bytes < info.step_bytes / 3


@pytest.mark.parametrize("dt", ["complex64", "complex128"])
def test_m10_fused_equals_unfused_and_reference(data_dir, dt):
    from jet_b200 import ContractionPlan, NetworkFile
    dtype = np.dtype(dt)
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), dtype)
    sliced = ["p7", "s7", "h4", "m1", "m2", "I2"]
    with ContractionPlan(net, sliced, fuse=True, store_results=True) as fused, \
            ContractionPlan(net, sliced, fuse=False, store_results=True) as plain:
        assert fused.stats.chains > 0 and plain.stats.chains == 0
        # deferred slicing: small subtrees with open sliced indices moved to the slice-independent phase
        assert fused.stats.steps_shared > plain.stats.steps_shared
        assert fused.stats.jet_flops_per_slice == plain.stats.jet_flops_per_slice  # Jet-convention flops unchanged
        assert fused.stats.launches_per_slice < plain.stats.launches_per_slice
        assert fused.stats.fused_bytes_per_slice < 0.6 * fused.stats.bytes_per_slice
        ids = [0, 1, 17, 63]
        a = fused.amplitude(ids).reshape(-1)[0]
        b = plain.amplitude(ids).reshape(-1)[0]
        assert abs(a - b) / abs(b) < 10 * TOL[dtype]
        for o, sid in enumerate(ids):
            fa = fused.slice_result(o).reshape(-1)[0]
            pa = plain.slice_result(o).reshape(-1)[0]
            assert abs(fa - pa) / abs(pa) < 10 * TOL[dtype], (sid, fa, pa)


def test_four_group_layout_in_a_subprocess():
    """JB_CHAIN_LAYOUT=4x128 (four compute groups of four warps on 2^12-element tiles, six buffers: the measured
    alternative to the default two groups on 2^13-element tiles) is read once per process: the random-chain parity
    test above is repeated in a child process with the variable set."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JB_CHAIN_LAYOUT="4x128")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "random_chains or large_exact"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
