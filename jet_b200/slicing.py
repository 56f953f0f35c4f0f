"""Host-side path replay and greedy slice finding.

``replay`` restates the symbolic path replay of PathInfo (reference include/jet/PathInfo.hpp:262-297)
on plain Python lists; ``find_slices`` greedily picks indices to slice until the largest
intermediate fits a target (the offline cotengra SliceFinder step of the reference's benchmarks,
examples/paper_benchmarks/GPU/cot_gpu_m12/run_sliced.py:49-53).  Pure bookkeeping: no tensor data.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def replay(leaf_indices: Sequence[Sequence[str]], dims: Dict[str, int], path: Sequence[Tuple[int, int]],
           sliced: Sequence[str] = ()):
    """Returns (jet_flops, max_elems, candidates): 2*M*N*K summed over the steps
    (PathInfo::GetPathStepFlops convention, include/jet/PathInfo.hpp:157-183), the size of the
    largest intermediate and the indices of all intermediates of that size."""
    sl = set(sliced)
    nodes: List[List[str]] = [[i for i in idx if i not in sl] for idx in leaf_indices]
    flops = 0.0
    max_elems = 0
    biggest: List[str] = []
    sizes = []
    for a, b in path:
        A, B = nodes[a], nodes[b]
        sb = set(B)
        sa = set(A)
        out = [i for i in A if i not in sb] + [i for i in B if i not in sa]
        size = 1
        for i in out:
            size *= dims[i]
        k = 1
        for i in A:
            if i in sb:
                k *= dims[i]
        flops += 2.0 * size * k
        if size > max_elems:
            max_elems = size
        sizes.append(size)
        nodes.append(out)
    # candidate indices to slice next: every index of every largest intermediate
    seen = set()
    n_leaves = len(leaf_indices)
    for s, size in enumerate(sizes):
        if size == max_elems:
            for i in nodes[n_leaves + s]:
                if i not in seen:
                    seen.add(i)
                    biggest.append(i)
    return flops, max_elems, biggest


def find_slices(leaf_indices, dims, path, already: Sequence[str] = (), extra: int = 0, max_elems: int = 0):
    """Greedy: repeatedly slice the index of the current largest intermediate that gives the
    lowest total flops, until `extra` indices were added (extra > 0) or the largest intermediate
    has at most `max_elems` elements.  Returns the full list (already + new)."""
    chosen = list(already)
    while True:
        flops, mx, biggest = replay(leaf_indices, dims, path, chosen)
        if extra > 0 and len(chosen) - len(already) >= extra:
            break
        if extra <= 0 and mx <= max_elems:
            break
        best = None
        for cand in biggest:
            f, m, _ = replay(leaf_indices, dims, path, chosen + [cand])
            key = (m, f)
            if best is None or key < best[0]:
                best = (key, cand)
        if best is None:
            break
        chosen.append(best[1])
    return chosen
