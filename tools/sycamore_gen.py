#!/usr/bin/env python
"""Synthetic Sycamore-style random-circuit amplitude networks in the reference's JSON format
(BASELINE config 4: the reference ships only m10.json / m12.json).

Circuit: sites of a rotated square lattice (rows x cols, optionally one site removed: 9 x 6 minus one
= 53 qubits, 86-88 couplers in four classes A/B/C/D), `cycles` cycles of [random single-qubit gate
from {sqrt X, sqrt Y, sqrt W} on every qubit, never repeating on a qubit] + [fSim(pi/2, pi/6) on one
coupler class, sequence ABCDCDAB], a final single-qubit layer, input |0...0>, output bitstring `bits`.
The exact Sycamore coupler map is not recoverable from the reference; this layout is a documented
design choice (parity is judged against the reference engine on the SAME generated file).

Network: fSim(pi/2, phi) = SWAP . diag(1, -i, -i, e^{-i phi}) and the diagonal factor has
operator-Schmidt rank 2, so every two-qubit gate becomes two rank-3 (in, out, bond) tensors of
dimension 2 — the structure the tags of the reference's m10.json show (`FSIM` split in two) —
with the single-qubit gates, inputs and outputs absorbed into them.

  python tools/sycamore_gen.py --rows 9 --cols 6 --remove 0,0 --cycles 20 --seed 1 --target-log2 28 \
         --out data/syc53_m20.json
writes the network + path JSON and <out>.meta.json (sliced indices, costs).
`amplitude_statevector` is an independent brute-force check used by the tests on small lattices.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEQUENCE = "ABCDCDAB"
_s = 1 / math.sqrt(2)
GATES_1Q = {
    "X_1/2": _s * np.array([[1, -1j], [-1j, 1]], dtype=np.complex128),
    "Y_1/2": _s * np.array([[1, -1], [1, 1]], dtype=np.complex128),
    "W_1/2": _s * np.array([[1, -np.exp(0.25j * np.pi)], [np.exp(-0.25j * np.pi), 1]], dtype=np.complex128),
}
PHI = np.pi / 6
FSIM_DIAG = np.array([[1, -1j], [-1j, np.exp(-1j * PHI)]], dtype=np.complex128)  # d[a][b]
FSIM = np.array([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, np.exp(-1j * PHI)]], dtype=np.complex128)


def lattice(rows, cols, removed=None):
    """Sites (i, j) at x = 2j + (i mod 2); couplers join row i and i+1 at x +- 1.  Class A/B = even
    row gap left/right, C/D = odd row gap left/right."""
    sites = [(i, j) for i in range(rows) for j in range(cols) if (i, j) != removed]
    by_x = {(i, 2 * j + (i % 2)): (i, j) for (i, j) in sites}
    couplers = {c: [] for c in "ABCD"}
    for (i, j) in sites:
        x = 2 * j + (i % 2)
        for dx, side in ((-1, 0), (1, 1)):
            other = by_x.get((i + 1, x + dx))
            if other is not None:
                couplers["ABCD"[2 * (i % 2) + side]].append(((i, j), other))
    return sites, couplers


# The 54-qubit Sycamore grid as rows of a square lattice (a diamond-shaped patch; '-' = no qubit), written down
# from the published device layout; one qubit is removed for the 53-qubit processor (which one is a documented
# choice here: (3, 2), a boundary site).
SYCAMORE_GRID = """
-----AB---
----ABCD--
---ABCDEF-
--ABCDEFGH
-ABCDEFGHI
ABCDEFGHI-
-CDEFGHI--
--EFGHI---
---GHI----
----I-----
"""


def lattice_sycamore(removed=(3, 2)):
    """Square-lattice sites of the Sycamore patch and the four staggered coupler classes of the supremacy
    sequence: A / B = the two halves of the vertical couplers (r, c)-(r+1, c), C / D = the two halves of the
    horizontal couplers (r, c)-(r, c+1); a coupler belongs to the half fixed by the parity of its position in a
    2 x 2 unit cell, shifted by one on every other line ("staggered": the half-grid pattern whose cycles
    ABCDCDAB do not factor into independent stripes)."""
    rows = [ln for ln in SYCAMORE_GRID.strip("\n").split("\n")]
    sites = [(r, c) for r, ln in enumerate(rows) for c, ch in enumerate(ln) if ch != "-" and (r, c) != removed]
    have = set(sites)

    def in_layer(a, b, col_offset, vertical):
        if vertical:  # transpose so that the pair lies along a row
            a, b = (a[1], a[0]), (b[1], b[0])
        a, b = sorted((a, b))
        if a[0] != b[0] or b[1] != a[1] + 1:
            return False
        pos = (a[0] % 2, (a[1] - col_offset) % 2)
        return pos == (0, 0) or pos == (1, 1)

    layers = {"A": (0, True), "B": (1, True), "C": (1, False), "D": (0, False)}
    couplers = {c: [] for c in "ABCD"}
    for (r, c) in sites:
        for other in ((r + 1, c), (r, c + 1)):
            if other in have:
                for name, (off, vert) in layers.items():
                    if in_layer((r, c), other, off, vert):
                        couplers[name].append(((r, c), other))
    return sites, couplers


def circuit(rows, cols, removed, cycles, seed, layout="brick"):
    """List of ("1q", site, name) / ("fsim", site_u, site_v) in time order."""
    rng = np.random.default_rng(seed)
    sites, couplers = lattice_sycamore(removed if removed is not None else (3, 2)) if layout == "sycamore" else lattice(rows, cols, removed)
    names = sorted(GATES_1Q)
    last = {}
    ops = []

    def layer_1q():
        for s in sites:
            choices = [n for n in names if n != last.get(s)]
            g = choices[int(rng.integers(len(choices)))]
            last[s] = g
            ops.append(("1q", s, g))

    for t in range(cycles):
        layer_1q()
        for (u, v) in couplers[SEQUENCE[t % len(SEQUENCE)]]:
            ops.append(("fsim", u, v))
    layer_1q()
    return sites, ops


def amplitude_statevector(sites, ops, bits):
    """<bits| C |0...0> by brute-force state-vector simulation (small lattices only)."""
    n = len(sites)
    pos = {s: k for k, s in enumerate(sites)}
    psi = np.zeros([2] * n, dtype=np.complex128)
    psi[(0,) * n] = 1
    for op in ops:
        if op[0] == "1q":
            k = pos[op[1]]
            psi = np.moveaxis(np.tensordot(GATES_1Q[op[2]], psi, axes=([1], [k])), 0, k)
        else:
            a, b = pos[op[1]], pos[op[2]]
            g = FSIM.reshape(2, 2, 2, 2)  # [a_out, b_out, a_in, b_in]
            psi = np.moveaxis(np.tensordot(g, psi, axes=([2, 3], [a, b])), [0, 1], [a, b])
    return psi[tuple(bits[pos[s]] for s in sites)]


def to_network(sites, ops, bits):
    """Leaves [(tags, indices, array)] of the closed amplitude network."""
    pos = {s: k for k, s in enumerate(sites)}
    counter = {"w": 0, "b": 0}

    def new(kind):
        counter[kind] += 1
        return f"{kind}{counter[kind]}"

    # per site: the wire label leaving the last tensor on it (None while only the input state and
    # single-qubit gates were seen) and the pending product of single-qubit gates since then
    wire = {s: None for s in sites}
    pend = {s: np.array([1, 0], dtype=np.complex128) for s in sites}  # vector until the first fsim
    leaves = []
    last_leaf = {s: None for s in sites}  # (leaf number, axis of the open wire) for the final absorb
    scalar = 1.0 + 0j
    ngate = 0
    for op in ops:
        if op[0] == "1q":
            pend[op[1]] = GATES_1Q[op[2]] @ pend[op[1]]
            continue
        _, u, v = op
        bond = new("b")
        ngate += 1
        outs = {}
        for site, table in ((u, FSIM_DIAG), (v, np.eye(2, dtype=np.complex128))):
            out = new("w")
            p = pend[site]
            if p.ndim == 1:  # input state absorbed: T[out, s] = p[out] * table[out, s]
                arr = p[:, None] * table
                idx = [out, bond]
                axis = 0
            else:  # T[prev, out, s] = p[out, prev] * table[out, s]
                arr = np.einsum("op,os->pos", p, table)
                idx = [wire[site], out, bond]
                axis = 1
            leaves.append(([f"FSIM", f"GATE_{ngate}", f"Q{pos[site]}"], idx, arr))
            outs[site] = (out, len(leaves) - 1, axis)
        # SWAP: the state leaving u's tensor continues on site v and vice versa
        for src, dst in ((u, v), (v, u)):
            out, leaf_no, axis = outs[src]
            wire[dst] = out
            last_leaf[dst] = (leaf_no, axis)
            pend[dst] = np.eye(2, dtype=np.complex128)
    # outputs: <bit| (pending gates) absorbed into the last tensor on each site
    for s in sites:
        p = pend[s]
        bra = np.zeros(2, dtype=np.complex128)
        bra[bits[pos[s]]] = 1
        if p.ndim == 1:  # a qubit no two-qubit gate ever touched: plain scalar factor
            scalar *= bra @ p
            continue
        row = bra @ p  # row[w] multiplies the open wire
        leaf_no, axis = last_leaf[s]
        tags, idx, arr = leaves[leaf_no]
        arr = np.tensordot(arr, row, axes=([axis], [0]))
        idx = idx[:axis] + idx[axis + 1:]
        leaves[leaf_no] = (tags, idx, arr)
        for other in sites:  # axes after the removed one shift down
            if last_leaf[other] is not None and last_leaf[other][0] == leaf_no and last_leaf[other][1] > axis:
                last_leaf[other] = (leaf_no, last_leaf[other][1] - 1)
    if leaves:
        tags, idx, arr = leaves[0]
        leaves[0] = (tags, idx, arr * scalar)
    return leaves


def network_json(leaves, path, dtype=np.complex64):
    out = {"path": [list(p) for p in path], "tensors": []}
    for tags, idx, arr in leaves:
        a = np.asarray(arr, dtype=dtype).reshape(-1)
        out["tensors"].append([tags, idx, list(arr.shape), [[float(z.real), float(z.imag)] for z in a]])
    return json.dumps(out, separators=(",", ":"))


def build(rows, cols, removed, cycles, seed, trials=8, target_log2=28, bits=None, max_sliced=62, layout="brick",
          optimizer="greedy", seconds=60.0, k=10):
    """Circuit -> network -> path + sliced indices.  optimizer = "greedy" (round-1 seeded greedy finder + greedy
    slicer) or "pathopt" (jet_b200/cpp/pathopt.cpp: bisection + subtree reconfiguration + slicing-aware search)."""
    from jet_b200.pathfinder import optimize, path_cost, search
    from jet_b200.slicing import find_slices
    sites, ops = circuit(rows, cols, removed, cycles, seed, layout)
    if bits is None:
        bits = np.random.default_rng(seed + 12345).integers(0, 2, len(sites)).tolist()
    leaves = to_network(sites, ops, bits)
    leaf_idx = [idx for _, idx, _ in leaves]
    dims = {i: 2 for idx in leaf_idx for i in idx}
    extra = {}
    if optimizer == "pathopt":
        rep = optimize(leaf_idx, dims, target_log2=target_log2, max_slices_log2=max_sliced, trials=trials, seconds=seconds,
                       seed=seed, k=k)
        path, sliced = rep["path"], rep["sliced"]
        peak, flops = rep["log2_peak_unsliced"], rep["jet_flops_unsliced"]
        extra = dict(optimizer="pathopt", jet_flops_total=rep["jet_flops_total"], search_seconds=rep["seconds"],
                     search_trials=rep["trials"])
    else:
        path, (peak, flops) = search(leaf_idx, dims, trials=trials, seed=seed)
        sliced = []
        if peak > target_log2:
            sliced = find_slices(leaf_idx, dims, path, [], max_elems=2 ** target_log2)
            if len(sliced) > max_sliced:
                sliced = sliced[:max_sliced]
    peak_s, flops_s = path_cost(leaf_idx, dims, path, sliced)
    meta = dict(rows=rows, cols=cols, removed=removed, cycles=cycles, seed=seed, layout=layout, qubits=len(sites), bits=bits,
                leaves=len(leaves), steps=len(path), log2_peak_unsliced=peak, jet_flops_unsliced=flops,
                sliced_indices=sliced, log2_num_slices=len(sliced), log2_peak_per_slice=peak_s, jet_flops_per_slice=flops_s,
                **extra)
    return sites, ops, bits, leaves, path, meta


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=9)
    ap.add_argument("--cols", type=int, default=6)
    ap.add_argument("--remove", default="0,0")
    ap.add_argument("--cycles", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--target-log2", type=int, default=28)
    ap.add_argument("--layout", default="brick", choices=["brick", "sycamore"],
                    help="brick: rows x cols rotated lattice (round 1); sycamore: the 54-site Sycamore patch with the "
                         "staggered half-grid coupler classes (--remove picks the missing qubit, default 3,2)")
    ap.add_argument("--optimizer", default="greedy", choices=["greedy", "pathopt"])
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--max-sliced", type=int, default=62)
    ap.add_argument("--out", default=os.path.join(ROOT, "data", "syc53_m20.json"))
    args = ap.parse_args()
    removed = tuple(int(v) for v in args.remove.split(",")) if args.remove else None
    _, _, _, leaves, path, meta = build(args.rows, args.cols, removed, args.cycles, args.seed, args.trials, args.target_log2,
                                        max_sliced=args.max_sliced, layout=args.layout, optimizer=args.optimizer,
                                        seconds=args.seconds, k=args.k)
    with open(args.out, "w") as f:
        f.write(network_json(leaves, path))
    json.dump(meta, open(args.out + ".meta.json", "w"), indent=1)
    print(json.dumps({k: v for k, v in meta.items() if k != "bits"}))


if __name__ == "__main__":
    main()
