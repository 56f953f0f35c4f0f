"""Fused-chain planner (jet_b200/csrc/chain_plan.h) replayed on the CPU.

The CUDA kernel jet_b200/csrc/chain.cu is an interpreter of the planner's parameter block; these
tests run the same interpretation on the host (tests/cpp/chain_emu.cpp, built here with g++) and
compare it with a step-by-step numpy contraction that follows Tensor::ContractTensors
(reference include/jet/Tensor.hpp:709-752: C = left ++ right, sum over the common indices).  They
also assert that the in-place shared-memory update has no cross-thread read/write overlap and report
the bank-conflict degree of every phase.  No GPU needed."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(tempfile.mkdtemp(prefix="chain_emu_"), "chain_emu.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "jet_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "chain_emu.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.chain_emu_run.restype = C.c_int
    return lib


def contract_bits(x, xb, r, rb, x_is_left):
    """x, r: arrays of shape (2,)*n whose axes are the bit ids xb / rb listed address-ASCENDING
    (so numpy axis i is bit xb[n-1-i]).  Returns (result, result bits ascending)."""
    xd, rd = xb[::-1], rb[::-1]  # axis order
    common = [b for b in xd if b in rd]
    xa = [xd.index(b) for b in common]
    ra = [rd.index(b) for b in common]
    xfree = [b for b in xd if b not in common]
    rfree = [b for b in rd if b not in common]
    if x_is_left:
        c = np.tensordot(x, r, axes=(xa, ra))
        cd = xfree + rfree
    else:
        c = np.tensordot(r, x, axes=(ra, xa))
        cd = rfree + xfree
    return c, cd[::-1]


def random_chain(rng, n_bits, n_steps, max_k=3, max_n=3):
    next_id = [0]

    def fresh(n):
        ids = list(range(next_id[0], next_id[0] + n))
        next_id[0] += n
        return ids

    x0_bits = fresh(n_bits)
    rng.shuffle(x0_bits)
    cur = list(x0_bits)
    steps = []
    for _ in range(n_steps):
        k = int(rng.integers(0, min(max_k, len(cur)) + 1))
        n = int(rng.integers(0 if k > 0 else 1, max_n + 1))
        s = list(rng.choice(cur, size=k, replace=False)) if k else []
        f = fresh(n)
        rb = [int(b) for b in s] + f
        rng.shuffle(rb)
        left = bool(rng.integers(0, 2))
        steps.append((rb, left))
        rest = [b for b in cur if b not in s]
        fr = [b for b in rb if b in f]
        cur = (fr + rest) if left else (rest + fr)
    return x0_bits, steps


def run_emu(emu, x0_bits, steps, x0, rs, elem_bytes=8, max_tile=13, lane_bits=4):
    n_steps = len(steps)
    r_nbits = (C.c_int * n_steps)(*[len(rb) for rb, _ in steps])
    flat = [b for rb, _ in steps for b in rb]
    r_bits = (C.c_int * max(len(flat), 1))(*flat)
    left = (C.c_int * n_steps)(*[int(l) for _, l in steps])
    xb = (C.c_int * len(x0_bits))(*x0_bits)
    x0c = np.ascontiguousarray(x0.reshape(-1), dtype=np.complex128)
    rcs = [np.ascontiguousarray(r.reshape(-1), dtype=np.complex128) for r in rs]
    rp = (C.c_void_p * n_steps)(*[r.ctypes.data for r in rcs])
    n_out = len(x0_bits) + sum(len(rb) for rb, _ in steps)
    out = np.zeros(1 << min(n_out, 26), dtype=np.complex128)
    n_out_bits = C.c_int(0)
    out_bits = (C.c_int * 128)()
    stats = (C.c_int * 8)()
    rc = emu.chain_emu_run(elem_bytes, len(x0_bits), xb, n_steps, r_nbits, r_bits, left, max_tile, lane_bits,
                           C.c_void_p(x0c.ctypes.data), rp, C.c_void_p(out.ctypes.data), C.byref(n_out_bits),
                           out_bits, stats)
    if rc != 0:
        return None
    nb = n_out_bits.value
    return out[: 1 << nb], [out_bits[i] for i in range(nb)], list(stats)


@pytest.mark.parametrize("elem_bytes", [8, 16])
def test_random_chains_match_stepwise_contraction(emu, elem_bytes):
    rng = np.random.default_rng(1234 + elem_bytes)
    planned = 0
    worst_step_conflict = 1
    for trial in range(120):
        n_bits = int(rng.integers(3, 17))
        n_steps = int(rng.integers(1, 6))
        x0_bits, steps = random_chain(rng, n_bits, n_steps)
        x0 = rng.standard_normal((2,) * n_bits) + 1j * rng.standard_normal((2,) * n_bits)
        rs = [rng.standard_normal((2,) * len(rb)) + 1j * rng.standard_normal((2,) * len(rb)) if rb
              else np.array(rng.standard_normal() + 1j * rng.standard_normal()) for rb, _ in steps]
        want, wb = x0, list(x0_bits)
        for (rb, left), r in zip(steps, rs):
            want, wb = contract_bits(want, wb, r, list(rb), left)
        if len(wb) > 22:
            continue
        res = run_emu(emu, x0_bits, steps, x0, rs, elem_bytes=elem_bytes)
        if res is None:
            continue  # the chain does not fit one tile: the engine would split it
        got, gb, stats = res
        planned += 1
        assert gb == wb, "output bit order differs from ContractTensors' left ++ right"
        assert stats[2] == 0, f"cross-thread hazard in the in-place update (trial {trial})"
        np.testing.assert_allclose(got, np.asarray(want).reshape(-1), rtol=1e-10, atol=1e-10)
        if n_bits - sum(1 for rb, _ in steps for b in rb if b in x0_bits) >= 5:
            worst_step_conflict = max(worst_step_conflict, stats[4])
    assert planned >= 60
    # with >= bank_bits untouched bits in the tile the compute phases are conflict-free by construction
    assert worst_step_conflict == 1


def test_gate_like_chain_is_conflict_free(emu):
    """The dominant Sycamore pattern: a rank-20 tensor absorbs four rank-4 tensors (K = N = 4)."""
    rng = np.random.default_rng(7)
    n_bits = 20
    x0_bits = list(range(n_bits))
    cur = list(x0_bits)
    nid = n_bits
    steps = []
    for s in range(4):
        sb = [int(b) for b in rng.choice(cur, size=2, replace=False)]
        f = [nid, nid + 1]
        nid += 2
        rb = sb + f
        rng.shuffle(rb)
        steps.append((rb, True))
        cur = [b for b in rb if b in f] + [b for b in cur if b not in sb]
    x0 = rng.standard_normal((2,) * n_bits) + 1j * rng.standard_normal((2,) * n_bits)
    rs = [rng.standard_normal((2,) * 4) + 1j * rng.standard_normal((2,) * 4) for _ in steps]
    want, wb = x0, list(x0_bits)
    for (rb, left), r in zip(steps, rs):
        want, wb = contract_bits(want, wb, r, list(rb), left)
    got, gb, stats = run_emu(emu, x0_bits, steps, x0, rs, lane_bits=5)
    assert gb == wb
    np.testing.assert_allclose(got, want.reshape(-1), rtol=1e-10, atol=1e-10)
    assert stats[0] <= 13 and stats[2] == 0
    assert stats[1] == 1 and stats[3] == 1 and stats[4] == 1 and stats[5] == 1, stats


@pytest.mark.parametrize("elem_bytes", [8, 16])
def test_padded_tiles_match_stepwise_contraction(emu, elem_bytes):
    """Short chains get their tile padded with untouched bits (MakeChainOp, jet_b200/csrc/chain.cu: a 2^8 tile is
    all hand-off latency): same results, same hazard-freedom, for every amount of padding."""
    rng = np.random.default_rng(99 + elem_bytes)
    planned = 0
    try:
        for trial in range(80):
            n_bits = int(rng.integers(8, 17))
            n_steps = int(rng.integers(1, 4))
            x0_bits, steps = random_chain(rng, n_bits, n_steps, max_k=2, max_n=2)
            x0 = rng.standard_normal((2,) * n_bits) + 1j * rng.standard_normal((2,) * n_bits)
            rs = [rng.standard_normal((2,) * len(rb)) + 1j * rng.standard_normal((2,) * len(rb)) if rb
                  else np.array(rng.standard_normal() + 1j * rng.standard_normal()) for rb, _ in steps]
            want, wb = x0, list(x0_bits)
            for (rb, left), r in zip(steps, rs):
                want, wb = contract_bits(want, wb, r, list(rb), left)
            emu.chain_emu_set_extra_quiet(0)
            base = run_emu(emu, x0_bits, steps, x0, rs, elem_bytes=elem_bytes, lane_bits=5)
            if base is None:
                continue
            emu.chain_emu_set_extra_quiet(int(rng.integers(1, 7)))
            res = run_emu(emu, x0_bits, steps, x0, rs, elem_bytes=elem_bytes, lane_bits=5)
            if res is None:
                continue  # the padded tile exceeds the limit: MakeChainOp then tries less padding
            got, gb, stats = res
            planned += 1
            assert gb == wb and stats[2] == 0
            assert stats[0] >= base[2][0]  # the padded tile is at least as large
            np.testing.assert_allclose(got, np.asarray(want).reshape(-1), rtol=1e-10, atol=1e-10)
    finally:
        emu.chain_emu_set_extra_quiet(0)
    assert planned >= 40
