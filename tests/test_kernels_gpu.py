"""GPU parity tests for K1 (permute), K2 (gemm) and the fused contraction, through the C ABI.

Bar: permutation bit-exact; contraction within 1e-5 (complex64) / 1e-12 (complex128) relative of
the oracle (normwise), the tolerance BASELINE.json's north_star states.
"""
import os

import numpy as np
import pytest

from oracle import jet_oracle as jo

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = {np.dtype(np.complex64): 1e-5, np.dtype(np.complex128): 1e-12}


def rand_c(rng, n, dtype):
    real = np.float32 if dtype == np.complex64 else np.float64
    return (rng.uniform(-1, 1, n).astype(real) + 1j * rng.uniform(-1, 1, n).astype(real)).astype(dtype)


def rel_err(x, ref):
    x = np.asarray(x, dtype=np.complex128).reshape(-1)
    ref = np.asarray(ref, dtype=np.complex128).reshape(-1)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300))


@pytest.fixture(scope="module")
def ops():
    from jet_b200 import ops as o
    return o


# ---------------------------------------------------------------- K1
def test_permute_reference_fixtures_bit_exact(ops):
    z = np.load(os.path.join(GOLDEN, "permute_cases.npz"))
    bad = []
    for i in range(int(z["count"])):
        out = ops.permute(z[f"c{i}_in"], z[f"c{i}_shape"].tolist(), z[f"c{i}_perm"].tolist())
        if not np.array_equal(out, z[f"c{i}_out"]):
            bad.append((i, z[f"c{i}_shape"].tolist(), z[f"c{i}_perm"].tolist(), str(out.dtype)))
    assert not bad, bad[:10]


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_permute_random_bit_permutations_vs_oracle(ops, dtype):
    rng = np.random.default_rng(11)
    bad = []
    for r in list(range(1, 21)) + [22, 23]:
        for trial in range(3 if r <= 16 else 1):
            perm = rng.permutation(r).tolist()
            x = rand_c(rng, 2 ** r, dtype)
            out = ops.permute(x, [2] * r, perm)
            if not np.array_equal(out, jo.transpose(x, [2] * r, perm)):
                bad.append((r, perm))
    assert not bad, bad[:5]


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_permute_pattern_classes(ops, dtype):
    """P1..P5 of SURVEY §8(d): pull-to-back, pull-to-front, random, reversal, last-5-fixed."""
    rng = np.random.default_rng(5)
    r = 20
    x = rand_c(rng, 2 ** r, dtype)
    pull = sorted(rng.choice(r, 3, replace=False).tolist())
    rest = [i for i in range(r) if i not in pull]
    perms = [rest + pull, pull + rest, rng.permutation(r).tolist(), list(range(r))[::-1],
             rng.permutation(r - 5).tolist() + list(range(r - 5, r)), list(range(r)),
             [1, 0] + list(range(2, r)), list(range(r - 2)) + [r - 1, r - 2]]
    for perm in perms:
        assert np.array_equal(ops.permute(x, [2] * r, perm), jo.transpose(x, [2] * r, perm)), perm


def test_permute_mixed_and_non_pow2_shapes(ops):
    rng = np.random.default_rng(6)
    for shape in ([4, 16, 2, 8, 4], [64, 64], [30, 50, 70], [3, 5, 7, 2, 4], [1, 8, 1, 4], [7], [2, 3], [6, 1, 5]):
        for dtype in (np.complex64, np.complex128):
            perm = rng.permutation(len(shape)).tolist()
            x = rand_c(rng, int(np.prod(shape)), dtype)
            assert np.array_equal(ops.permute(x, shape, perm), jo.transpose(x, shape, perm)), (shape, perm)


def test_permute_large_roundtrip_property(ops):
    """Full-size property (2^26 complex64 = 512 MiB): permute then inverse permute is the identity,
    and a permutation preserves the multiset (checksum)."""
    r = 26
    rng = np.random.default_rng(9)
    x = (np.arange(2 ** r, dtype=np.float32) % 8191).astype(np.float32)
    x = (x + 1j * (x + 1)).astype(np.complex64)
    perm = rng.permutation(r).tolist()
    inv = np.argsort(perm).tolist()
    y = ops.permute(x, [2] * r, perm)
    assert y.view(np.float32).astype(np.float64).sum() == x.view(np.float32).astype(np.float64).sum()
    # spot-check against the definition on a sample of addresses
    idx = rng.integers(0, 2 ** r, 2000)
    bits = (idx[:, None] >> np.arange(r - 1, -1, -1)[None, :]) & 1  # out multi-index, axis 0 first
    src = np.zeros(len(idx), dtype=np.int64)
    for j in range(r):
        src |= bits[:, j].astype(np.int64) << (r - 1 - perm[j])
    assert np.array_equal(y[idx], x[src])
    z = ops.permute(y, [2] * r, inv)
    assert np.array_equal(z, x)


def test_permute_rejects_bad_arguments(ops):
    from jet_b200 import JetB200Error
    x = np.zeros(8, np.complex64)
    with pytest.raises(JetB200Error):
        ops.permute(x, [2, 2, 2], [0, 0, 1])
    with pytest.raises(ValueError):
        ops.permute(x, [2, 2], [0, 1])


# ---------------------------------------------------------------- K2
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_gemm_shapes_vs_oracle(ops, dtype):
    rng = np.random.default_rng(3)
    shapes = [(1, 1, 1), (1, 1, 1000), (1, 7, 33), (9, 1, 130), (2, 2, 12), (64, 64, 64), (65, 63, 17), (128, 256, 96),
              (16, 32, 1 << 15), (1, 1, 1 << 18), (300, 5, 4096), (5, 300, 2500), (1000, 3, 2), (3, 1000, 1)]
    for m, n, k in shapes:
        a = rand_c(rng, m * k, dtype).reshape(m, k)
        b = rand_c(rng, k * n, dtype).reshape(k, n)
        c = ops.gemm(a, b)
        ref = a.astype(np.complex128) @ b.astype(np.complex128)
        assert rel_err(c, ref) < TOL[np.dtype(dtype)], (m, n, k, rel_err(c, ref))


def test_gemm_kat_from_reference(ops):
    # test/Test_Tensor.cpp:372-397: 2x12 . 12x2 of (0.5, 0.25) -> (2.25, 3.0)
    a = np.full((2, 12), 0.5 + 0.25j, dtype=np.complex64)
    b = np.full((12, 2), 0.5 + 0.25j, dtype=np.complex64)
    assert np.array_equal(ops.gemm(a, b), np.full((2, 2), 2.25 + 3.0j, dtype=np.complex64))


# ---------------------------------------------------------------- fused contraction
def test_contract_reference_fixtures(ops):
    z = np.load(os.path.join(GOLDEN, "contract_cases.npz"))
    bad = []
    for i in range(int(z["count"])):
        sa, sb = z[f"c{i}_sa"].tolist(), z[f"c{i}_sb"].tolist()
        a = z[f"c{i}_a"].reshape(sa)
        b = z[f"c{i}_b"].reshape(sb)
        c, _ = ops.contract(a, z[f"c{i}_ia"].tolist(), b, z[f"c{i}_ib"].tolist())
        e = rel_err(c, z[f"c{i}_c"])
        if not e < TOL[a.dtype]:
            bad.append((i, sa, z[f"c{i}_ia"].tolist(), sb, z[f"c{i}_ib"].tolist(), str(a.dtype), e))
    assert not bad, bad[:6]


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_contract_random_tn_shapes_vs_oracle(ops, dtype):
    """S1/S3-style shapes of SURVEY §8(d): big operand x small operand with scattered contracted
    indices in both orders, both-large (TTGT), split-K."""
    rng = np.random.default_rng(17)
    bad = []
    specs = []
    for ra in (5, 9, 14, 18, 20):
        for rb in (2, 3, 4, 6, 8):
            for c in range(0, min(ra, rb) + 1):
                if rb - c > 6 or (rng.random() < 0.5 and ra > 9):
                    continue
                specs.append((ra, rb, c))
    specs += [(13, 13, 5), (14, 14, 10), (16, 15, 13), (18, 18, 16), (12, 12, 12), (20, 20, 20)]
    for ra, rb, c in specs:
        for swap in (False, True):
            ids_a = list(range(ra))
            common = sorted(rng.choice(ra, c, replace=False).tolist())
            ids_b = common + list(range(100, 100 + rb - c))
            ids_b = [ids_b[i] for i in rng.permutation(rb)]
            a = rand_c(rng, 2 ** ra, dtype).reshape([2] * ra)
            b = rand_c(rng, 2 ** rb, dtype).reshape([2] * rb)
            if swap:
                a, b, ids_a, ids_b = b, a, ids_b, ids_a
            out, modes = ops.contract(a, ids_a, b, ids_b)
            idx, ref = jo.contract(([str(i) for i in ids_a], a), ([str(i) for i in ids_b], b))
            if [str(m) for m in modes] != idx:
                bad.append(("modes", ra, rb, c, swap))
                continue
            e = rel_err(out, ref)
            if not e < TOL[np.dtype(dtype)]:
                bad.append((ra, rb, c, swap, e))
    assert not bad, bad[:8]


def test_contract_dim4_and_non_pow2(ops):
    rng = np.random.default_rng(23)
    cases = [([4] * 7, list(range(7)), [4] * 3, [2, 50, 5]), ([4] * 6, list(range(6)), [4] * 6, [9, 3, 8, 1, 7, 0]),
             ([3, 5, 7], [0, 1, 2], [7, 3, 2], [2, 0, 9]), ([6, 10], [0, 1], [10, 6], [1, 0]),
             ([2, 3, 5], [0, 1, 2], [5, 3, 4], [2, 1, 3])]
    for sa, ia, sb, ib in cases:
        for dtype in (np.complex64, np.complex128):
            a = rand_c(rng, int(np.prod(sa)), dtype).reshape(sa)
            b = rand_c(rng, int(np.prod(sb)), dtype).reshape(sb)
            out, _ = ops.contract(a, ia, b, ib)
            _, ref = jo.contract(([str(i) for i in ia], a), ([str(i) for i in ib], b))
            assert out.shape == ref.shape
            assert rel_err(out, ref) < TOL[np.dtype(dtype)], (sa, ia, sb, ib)


def test_contract_linearity_property_large(ops):
    """Size-independent property at 2^24 elements: contract(a1 + a2, b) == contract(a1, b) +
    contract(a2, b) up to rounding, and exact small-integer data gives exact results."""
    r = 24
    rng = np.random.default_rng(31)
    ids_a = list(range(r))
    ids_b = [3, 100, 17, 101, 23]
    a = (rng.integers(-2, 3, 2 ** r) + 1j * rng.integers(-2, 3, 2 ** r)).astype(np.complex64).reshape([2] * r)
    b = (rng.integers(-2, 3, 32) + 1j * rng.integers(-2, 3, 32)).astype(np.complex64).reshape([2] * 5)
    out, _ = ops.contract(a, ids_a, b, ids_b)
    _, ref = jo.contract(([str(i) for i in ids_a], a), ([str(i) for i in ids_b], b))
    assert np.array_equal(out, ref)  # integers: exact in fp32 irrespective of summation order


def test_add_and_slice(ops):
    rng = np.random.default_rng(2)
    for dtype in (np.complex64, np.complex128):
        a = rand_c(rng, 3000, dtype)
        b = rand_c(rng, 3000, dtype)
        assert np.array_equal(ops.add(a, b), a + b)
        t = rand_c(rng, 2 * 3 * 4 * 5, dtype).reshape(2, 3, 4, 5)
        for ax in range(4):
            assert np.array_equal(ops.slice_index(t, ax, 1), np.take(t, 1, axis=ax))


# ---------------------------------------------------------------- tensor-core GEMM (tcgen05, 3xTF32)
def test_gemm_tensor_core_shapes_vs_fp64(ops):
    """Shapes eligible for the tcgen05 3xTF32 kernel (M % 128 == 0, N % 64 == 0, K % 16 == 0,
    K >= 32, M N K >= 2^24): complex64 result within 1e-5 of the complex128 product (normwise and on the largest
    elements), i.e. the split keeps FP32-class accuracy.  (1 << 13, 256, 32) and (4096, 128, 48) are the short-K
    shapes — one and one and a half accumulation chunks — of the kind one m=20 slice has."""
    rng = np.random.default_rng(41)
    for m, n, k in [(1 << 13, 256, 32), (4096, 128, 48), (128, 64, 64), (256, 128, 512), (1024, 1024, 1024), (4096, 512, 2048), (128, 4096, 96), (384, 192, 4000 // 16 * 16)]:
        a = rand_c(rng, m * k, np.complex64).reshape(m, k)
        b = rand_c(rng, k * n, np.complex64).reshape(k, n)
        c = ops.gemm(a, b)
        ref = a.astype(np.complex128) @ b.astype(np.complex128)
        e = rel_err(c, ref)
        assert e < 2e-6, (m, n, k, e)
        big = np.abs(ref) > 0.5 * np.abs(ref).max()
        assert np.max(np.abs(c[big] - ref[big]) / np.abs(ref[big])) < 1e-5, (m, n, k)


def test_gemm_tensor_core_ragged_and_split_k(ops):
    """Shapes the tcgen05 kernel serves with TMA boxes smaller than its 128 x 128 tile (M < 128, 2N < 128,
    M not a multiple of 128) and with split-K (few output tiles, long K: the shapes of the last steps of
    the m=20 paths).  Integer data makes FP32 exact, so the result must be bit-identical; random data
    must stay within 2e-6 normwise of the complex128 product."""
    rng = np.random.default_rng(45)
    shapes = [(1 << 12, 128, 32), (64, 512, 256), (32, 1024, 256), (2048, 32, 1024), (4096, 16, 512), (192, 96, 128), (320, 40, 96),
              (512, 1024, 1 << 14), (256, 64, 1 << 15), (128, 128, 8192), (64, 32, 1 << 14)]
    for m, n, k in shapes:
        a = (rng.integers(-2, 3, (m, k)) + 1j * rng.integers(-2, 3, (m, k))).astype(np.complex64)
        b = (rng.integers(-2, 3, (k, n)) + 1j * rng.integers(-2, 3, (k, n))).astype(np.complex64)
        ref = (a.astype(np.complex128) @ b.astype(np.complex128)).astype(np.complex64)
        assert np.array_equal(ops.gemm(a, b), ref), (m, n, k)
        a = rand_c(rng, m * k, np.complex64).reshape(m, k)
        b = rand_c(rng, k * n, np.complex64).reshape(k, n)
        ref = a.astype(np.complex128) @ b.astype(np.complex128)
        e = rel_err(ops.gemm(a, b), ref)
        assert e < 2e-6, (m, n, k, e)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_gemm_dot_and_gemv_long_k(ops, dtype):
    """The DOTU / GEMV corner (M, N in {1, 2, 4}, long K: reference TensorHelpers.hpp:79-111 — the last
    step of every closed network) runs on the streaming SmallMn kernel with double accumulation."""
    rng = np.random.default_rng(46)
    for m, n, k in [(1, 1, 1 << 22), (1, 1, 4096), (1, 4, 1 << 18), (4, 1, 1 << 18), (2, 2, 100003), (4, 4, 1 << 16),
                    (2, 1, 5000), (1, 2, 1 << 20)]:
        a = rand_c(rng, m * k, dtype).reshape(m, k)
        b = rand_c(rng, k * n, dtype).reshape(k, n)
        ref = a.astype(np.complex128) @ b.astype(np.complex128)
        scale = np.sqrt(k)  # |sum| ~ sqrt(k): compare against the conditioning-free magnitude
        err = np.max(np.abs(ops.gemm(a, b) - ref)) / scale
        assert err < (1e-6 if dtype == np.complex64 else 1e-13), (m, n, k, err)
    k = 1 << 20
    a = (rng.integers(-2, 3, (1, k)) + 1j * rng.integers(-2, 3, (1, k))).astype(dtype)
    b = (rng.integers(-2, 3, (k, 1)) + 1j * rng.integers(-2, 3, (k, 1))).astype(dtype)
    assert np.array_equal(ops.gemm(a, b), (a.astype(np.complex128) @ b.astype(np.complex128)).astype(dtype))


def test_gemm_fp64_tensor_pipe_shapes(ops):
    """complex128 shapes served by the DMMA kernel (gemm_dmma.cu: K % 8 == 0, M >= 32, N >= 16, ragged
    M / N tiles): 1e-12 normwise against numpy's complex128 product (BASELINE tolerance), and
    bit-identical on small integers (FP64 exact)."""
    rng = np.random.default_rng(48)
    for m, n, k in [(128, 64, 64), (4096, 64, 64), (1 << 15, 64, 64), (200, 48, 24), (32, 16, 8), (1000, 100, 200),
                    (512, 512, 512), (96, 200, 1024), (64, 64, 1 << 16), (128, 32, 40000)]:
        a = rand_c(rng, m * k, np.complex128).reshape(m, k)
        b = rand_c(rng, k * n, np.complex128).reshape(k, n)
        ref = a @ b
        e = rel_err(ops.gemm(a, b), ref)
        assert e < 1e-12, (m, n, k, e)
        a = (rng.integers(-3, 4, (m, k)) + 1j * rng.integers(-3, 4, (m, k))).astype(np.complex128)
        b = (rng.integers(-3, 4, (k, n)) + 1j * rng.integers(-3, 4, (k, n))).astype(np.complex128)
        assert np.array_equal(ops.gemm(a, b), a @ b), (m, n, k)


def test_contract_c128_compute_bound_step_uses_ttgt_dmma(ops):
    """The GBS fock-8 step shape (dim-8 indices: M large, N = K = 64) in complex128 is routed to
    TTGT + DMMA (kernel 1) and matches the oracle's ContractTensors to 1e-12."""
    rng = np.random.default_rng(49)
    shape_a, ia = [8, 8, 8, 8, 8], [0, 1, 2, 3, 4]
    shape_b, ib = [8, 8, 8, 8], [3, 100, 1, 101]
    a = rand_c(rng, 8 ** 5, np.complex128).reshape(shape_a)
    b = rand_c(rng, 8 ** 4, np.complex128).reshape(shape_b)
    info = ops.contract_info(np.complex128, shape_a, ia, shape_b, ib)
    assert info.kernel == 1 and info.m == 512 and info.n == 64 and info.k == 64
    out, modes = ops.contract(a, ia, b, ib)
    _, ref = jo.contract(([str(i) for i in ia], a), ([str(i) for i in ib], b))
    assert rel_err(out, ref) < 1e-12


def test_contract_c128_gather_dmma_random_layouts(ops):
    """complex128 TTGT steps read A in its original layout (the DMMA kernel folds Transpose(A) into its
    tile loads): random positions of the contracted indices, dims 2 / 4 / 8, ragged M (< 128 rows)."""
    rng = np.random.default_rng(50)
    for dims, ra, nc, nf in [(2, 18, 6, 6), (2, 14, 5, 7), (4, 9, 3, 3), (8, 6, 2, 2), (2, 12, 6, 6), (2, 13, 5, 5)]:
        ia = list(range(ra))
        common = sorted(rng.choice(ra, nc, replace=False).tolist())
        ib = common + list(range(100, 100 + nf))
        ib = [ib[i] for i in rng.permutation(len(ib))]
        a = rand_c(rng, dims ** ra, np.complex128).reshape([dims] * ra)
        b = rand_c(rng, dims ** len(ib), np.complex128).reshape([dims] * len(ib))
        info = ops.contract_info(np.complex128, a.shape, ia, b.shape, ib)
        assert info.kernel == 1, (dims, ra, nc, nf)
        out, _ = ops.contract(a, ia, b, ib)
        _, ref = jo.contract(([str(i) for i in ia], a), ([str(i) for i in ib], b))
        assert rel_err(out, ref) < 1e-12, (dims, ra, nc, nf, rel_err(out, ref))


def test_contract_c64_short_m_long_n_swaps_gemm_roles(ops):
    """complex64 TTGT step with a short M and a long N (the m=20 shapes 64 x 2^18 x 1024, 32 x 2^18 x 256):
    the engine computes C^T = B^T A^T on the tensor cores and transposes the small result; labels, order
    (left ++ right) and values must be those of ContractTensors."""
    rng = np.random.default_rng(51)
    for ra, rb, nc in [(13, 20, 7), (13, 21, 8), (14, 22, 8)]:
        ia = [int(v) for v in rng.permutation(ra)]
        common = sorted(rng.choice(ia, nc, replace=False).tolist())
        ib = common + list(range(100, 100 + rb - nc))
        ib = [ib[i] for i in rng.permutation(rb)]
        a = rand_c(rng, 2 ** ra, np.complex64).reshape([2] * ra)
        b = rand_c(rng, 2 ** rb, np.complex64).reshape([2] * rb)
        info = ops.contract_info(np.complex64, a.shape, ia, b.shape, ib)
        assert info.kernel == 1 and info.m == 2 ** (ra - nc) and info.n == 2 ** (rb - nc)
        out, modes = ops.contract(a, ia, b, ib)
        want_modes, ref = jo.contract(([str(i) for i in ia], a.astype(np.complex128)), ([str(i) for i in ib], b.astype(np.complex128)))
        assert [str(m) for m in modes] == want_modes
        assert rel_err(out, ref) < 2e-6, (ra, rb, nc, rel_err(out, ref))
    # integers: exact
    a = (rng.integers(-2, 3, 2 ** 13) + 1j * rng.integers(-2, 3, 2 ** 13)).astype(np.complex64).reshape([2] * 13)
    b = (rng.integers(-2, 3, 2 ** 20) + 1j * rng.integers(-2, 3, 2 ** 20)).astype(np.complex64).reshape([2] * 20)
    ia, ib = list(range(13)), [0, 2, 4, 6, 8, 10, 12] + list(range(100, 113))
    out, _ = ops.contract(a, ia, b, ib)
    _, ref = jo.contract(([str(i) for i in ia], a.astype(np.complex128)), ([str(i) for i in ib], b.astype(np.complex128)))
    assert np.array_equal(out, ref.astype(np.complex64))


def test_gemm_tensor_core_exact_on_small_integers(ops):
    """Integer data: every partial product and sum is exact in FP32, so the tensor-core path must
    reproduce the integer result bit for bit (catches layout / swizzle / sign errors)."""
    rng = np.random.default_rng(43)
    m, n, k = 256, 192, 320
    a = (rng.integers(-3, 4, (m, k)) + 1j * rng.integers(-3, 4, (m, k))).astype(np.complex64)
    b = (rng.integers(-3, 4, (k, n)) + 1j * rng.integers(-3, 4, (k, n))).astype(np.complex64)
    ref = (a.astype(np.complex128) @ b.astype(np.complex128)).astype(np.complex64)
    assert np.array_equal(ops.gemm(a, b), ref)


def test_contract_both_operands_large_uses_tensor_cores(ops):
    rng = np.random.default_rng(47)
    ra = rb = 20
    c = 10
    ids_a = list(range(ra))
    common = sorted(rng.choice(ra, c, replace=False).tolist())
    ids_b = common + list(range(100, 100 + rb - c))
    ids_b = [ids_b[i] for i in rng.permutation(rb)]
    a = rand_c(rng, 2 ** ra, np.complex64).reshape([2] * ra)
    b = rand_c(rng, 2 ** rb, np.complex64).reshape([2] * rb)
    info = ops.contract_info(np.complex64, a.shape, ids_a, b.shape, ids_b)
    assert info.kernel == 1 and info.m == info.n == info.k == 1024
    out, _ = ops.contract(a, ids_a, b, ids_b)
    _, ref = jo.contract(([str(i) for i in ids_a], a.astype(np.complex128)), ([str(i) for i in ids_b], b.astype(np.complex128)))
    assert rel_err(out, ref) < 2e-6


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_contract_small_output_moderate_k_reads_operands_in_place(ops, dtype):
    """M, N <= 16 and 64 <= K <= 2^14 (the last per-slice steps of the small sliced networks, e.g. 16 x 16 x 4096), both
    operands larger than the stream kernel's resident limit.  SmallGemmGatherKernel: one thread-block cluster per contraction, both
    operands gathered from their original layouts, chunks of 64 products summed in the operand precision and
    accumulated in double.  Random index orders, dim-2 and dim-4 axes, against the oracle."""
    rng = np.random.default_rng(29)
    for trial, (n_common, fa, fb, with_dim4) in enumerate([(12, 4, 4, False), (13, 0, 3, False), (12, 2, 2, True),
                                                           (14, 1, 0, False), (13, 4, 0, False), (10, 3, 4, True),
                                                           (9, 4, 4, False)]):
        common = [f"k{i}" for i in range(n_common)]
        dims = {i: 2 for i in common}
        if with_dim4:
            dims[common[0]] = 4
        ia = common + [f"a{i}" for i in range(fa)]
        ib = common + [f"b{i}" for i in range(fb)]
        for i in ia + ib:
            dims.setdefault(i, 2)
        ia = [ia[i] for i in rng.permutation(len(ia))]
        ib = [ib[i] for i in rng.permutation(len(ib))]
        sa, sb = [dims[i] for i in ia], [dims[i] for i in ib]
        a = rand_c(rng, int(np.prod(sa)), dtype).reshape(sa)
        b = rand_c(rng, int(np.prod(sb)), dtype).reshape(sb)
        labels = {name: k for k, name in enumerate(sorted(dims))}
        info = ops.contract_info(dtype, sa, [labels[i] for i in ia], sb, [labels[i] for i in ib])
        assert info.kernel == 1 and info.ws_bytes == 0, (trial, info.kernel, info.ws_bytes)
        got, modes = ops.contract(a, [labels[i] for i in ia], b, [labels[i] for i in ib])
        idx, want = jo.contract((ia, a), (ib, b))
        assert [labels[i] for i in idx] == list(modes)
        scale = np.sqrt(float(info.k))
        err = np.abs(np.asarray(got, dtype=np.complex128).reshape(-1) - np.asarray(want, dtype=np.complex128).reshape(-1)).max() / scale
        assert err < TOL[np.dtype(dtype)], (trial, err)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_contract_final_dot_reads_operands_in_place(ops, dtype):
    """The last step of a closed network: two large tensors contracted over (almost) all indices, M, N in {1, 2, 4},
    K >= 2^16 (DotGatherKernel: both operands read once in their original layouts, no permuted copies; FP64
    partial sums in a fixed order).  Random index orders on both sides, dim-2 and dim-4 axes, against the oracle;
    the complex64 result must also agree with the TTGT path (JB_DISABLE_DOT_GATHER) to FP32 accuracy."""
    rng = np.random.default_rng(23)
    for trial, (n_common, fa, fb, with_dim4) in enumerate([(17, 1, 1, False), (18, 2, 2, False), (16, 0, 2, True),
                                                           (19, 2, 0, False), (16, 0, 0, True), (20, 1, 2, False)]):
        common = [f"k{i}" for i in range(n_common)]
        dims = {i: 2 for i in common}
        if with_dim4:
            for i in common[:3]:
                dims[i] = 4
        ia = common + [f"a{i}" for i in range(fa)]
        ib = common + [f"b{i}" for i in range(fb)]
        for i in ia + ib:
            dims.setdefault(i, 2)
        ia = [ia[i] for i in rng.permutation(len(ia))]
        ib = [ib[i] for i in rng.permutation(len(ib))]
        sa, sb = [dims[i] for i in ia], [dims[i] for i in ib]
        a = rand_c(rng, int(np.prod(sa)), dtype).reshape(sa)
        b = rand_c(rng, int(np.prod(sb)), dtype).reshape(sb)
        labels = {name: k for k, name in enumerate(sorted(dims))}
        info = ops.contract_info(dtype, sa, [labels[i] for i in ia], sb, [labels[i] for i in ib])
        assert info.kernel == 1 and info.m * info.n <= 16
        got, modes = ops.contract(a, [labels[i] for i in ia], b, [labels[i] for i in ib])
        idx, want = jo.contract((ia, a), (ib, b))
        assert [labels[i] for i in idx] == list(modes)
        # the sum has sqrt(K) growth and cancellation: compare against the scale of the terms
        scale = np.sqrt(float(info.k)) * 1.0
        err = np.abs(np.asarray(got, dtype=np.complex128).reshape(-1) - np.asarray(want, dtype=np.complex128).reshape(-1)).max() / scale
        assert err < TOL[np.dtype(dtype)], (trial, err)
