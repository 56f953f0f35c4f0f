"""Contraction-path search (host-side bookkeeping, no tensor data): `optimize` drives the native optimiser
(jet_b200/cpp/pathopt.cpp: recursive bisection + subtree reconfiguration + slicing-aware search, the recipe of
the cotengra runs the reference's benchmarks were prepared with); `greedy_path` / `search` are the seeded greedy
finder of round 1, kept for small networks and as a fallback for non-power-of-two extents.

Replaces, for the purpose of synthesising inputs, the reference's random-sampling path finder
(reference python/jet/interpreter.py:533-618) and the offline cotengra search its benchmarks used
(examples/paper_benchmarks/GPU/cot_gpu_m12/run_sliced.py:37-53).  Paths are emitted in the
reference's format: pairs of node ids, step i creating node num_leaves + i
(include/jet/TensorNetwork.hpp:301-328).  Path *quality* is not part of parity: both engines consume
the same emitted file.
"""
from __future__ import annotations

import heapq
import math
import random
import json
import os
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

PATHOPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "pathopt")


def optimize(leaf_indices: Sequence[Sequence[str]], dims: Dict[str, int], target_log2: int = 30, max_slices_log2: int = 40,
             trials: int = 16, seconds: float = 60.0, threads: int = 8, seed: int = 1, k: int = 10,
             fixed_slices: Sequence[str] = ()) -> dict:
    """Path + sliced indices for a network whose extents are powers of two.  Returns the optimiser's report:
    {"path": [[a, b], ...], "sliced": [names], "log2_slices", "log2_peak_per_slice", "jet_flops_per_slice",
    "jet_flops_total", "log2_peak_unsliced", "jet_flops_unsliced", ...} (Jet convention: 2*M*N*K per step,
    reference include/jet/PathInfo.hpp:157-183)."""
    if not os.path.exists(PATHOPT):
        raise RuntimeError(f"{PATHOPT} is missing: build it with `make -C jet_b200/cpp pathopt`")
    number: Dict[str, int] = {}
    for idx in leaf_indices:
        for i in idx:
            number.setdefault(i, len(number))
    order = sorted(number, key=number.get)
    lines = [f"{len(leaf_indices)} {len(order)}"] + [f"{name} {int(dims[name])}" for name in order]
    lines += [" ".join([str(len(idx))] + [str(number[i]) for i in idx]) for idx in leaf_indices]
    with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
        f.write("\n".join(lines) + "\n")
        name = f.name
    try:
        cmd = [PATHOPT, name, "--target", str(target_log2), "--max-slices", str(max_slices_log2), "--trials", str(trials),
               "--seconds", str(seconds), "--threads", str(threads), "--seed", str(seed), "--k", str(k)]
        for s in fixed_slices:
            cmd += ["--slice", s]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"pathopt failed: {p.stderr.strip()}")
        out = json.loads(p.stdout)
    finally:
        os.unlink(name)
    out["path"] = [(int(a), int(b)) for a, b in out["path"]]
    return out


def greedy_path(leaf_indices: Sequence[Sequence[str]], dims: Dict[str, int], seed: int = 0, temperature: float = 0.0,
                alpha: float = 1.0) -> List[Tuple[int, int]]:
    """Greedy pairwise contraction: repeatedly contract the connected pair minimising
    size(out) - alpha * (size(a) + size(b)), with optional Boltzmann noise (temperature > 0)."""
    rng = random.Random(seed)
    logd = {i: math.log2(d) for i, d in dims.items()}
    nodes: Dict[int, frozenset] = {n: frozenset(idx) for n, idx in enumerate(leaf_indices)}
    owners: Dict[str, set] = {}
    for n, idx in nodes.items():
        for i in idx:
            owners.setdefault(i, set()).add(n)
    next_id = len(nodes)
    path: List[Tuple[int, int]] = []

    def size(idx):
        return 2.0 ** sum(logd[i] for i in idx)

    def score(a, b):
        ia, ib = nodes[a], nodes[b]
        out = ia ^ ib
        s = size(out) - alpha * (size(ia) + size(ib))
        if temperature > 0:
            # Boltzmann noise on the (log-scaled) magnitude
            g = -math.log(-math.log(rng.random() + 1e-300) + 1e-300)
            s = s - temperature * g * max(abs(s), 1.0)
        return s

    heap = []

    def push_pairs(n):
        seen = set()
        for i in nodes[n]:
            for other in owners.get(i, ()):
                if other != n and other not in seen and other in nodes:
                    seen.add(other)
                    a, b = (n, other) if n < other else (other, n)
                    heapq.heappush(heap, (score(a, b), a, b))

    for n in list(nodes):
        push_pairs(n)

    def contract(a, b):
        nonlocal next_id
        ia, ib = nodes.pop(a), nodes.pop(b)
        out = ia ^ ib
        for i in ia | ib:
            s = owners.get(i)
            if s is not None:
                s.discard(a)
                s.discard(b)
                if i in out:
                    s.add(next_id)
                elif not s:
                    del owners[i]
        nodes[next_id] = out
        path.append((a, b))
        c = next_id
        next_id += 1
        return c

    while heap:
        _, a, b = heapq.heappop(heap)
        if a not in nodes or b not in nodes:
            continue
        c = contract(a, b)
        push_pairs(c)
    # disconnected remainders: outer products, smallest first
    rest = sorted(nodes, key=lambda n: size(nodes[n]))
    while len(rest) > 1:
        a, b = rest[0], rest[1]
        c = contract(a, b)
        rest = sorted([n for n in rest[2:]] + [c], key=lambda n: size(nodes[n]))
    return path


def path_cost(leaf_indices, dims, path, sliced: Sequence[str] = ()):
    """(log2 of largest intermediate, Jet-convention flops 2*M*N*K summed over the steps)."""
    from .slicing import replay
    flops, mx, _ = replay(leaf_indices, dims, path, sliced)
    return math.log2(max(mx, 1)), flops


def search(leaf_indices, dims, trials: int = 16, seed: int = 0, temperature: float = 0.3):
    """Best of `trials` seeded greedy runs by (peak size, flops)."""
    best = None
    for t in range(trials):
        p = greedy_path(leaf_indices, dims, seed=seed + t, temperature=0.0 if t == 0 else temperature)
        key = path_cost(leaf_indices, dims, p)
        if best is None or key < best[0]:
            best = (key, p)
    return best[1], best[0]
