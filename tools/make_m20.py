#!/usr/bin/env python
"""Writes data/sycamore53_m20.json (+ .meta.json): BASELINE config 4, the north-star workload.

The reference ships no m=20 network (only m10.json / m12.json), so the circuit is synthesised by
tools/sycamore_gen.py with the faithful device layout (--layout sycamore: the 54-site Sycamore patch minus one
qubit, staggered half-grid coupler classes, 20 cycles ABCDCDAB of sqrt-X/Y/W + fSim(pi/2, pi/6)), and the path
and the sliced indices come from the native optimiser (jet_b200/cpp/pathopt.cpp).  Sanity anchor: the same
generator at 10 / 12 cycles gives networks whose optimised cost (2.6e10 / 3.2e13 Jet-flops) matches the
reference's shipped m10.json / m12.json + their cotengra paths (1.4e10 / 2.0e13).

    python tools/make_m20.py [--target 30] [--seconds 240] [--trials 48] [--seed 7] [--k 11]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sycamore_gen as sg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--target", type=int, default=30)
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--trials", type=int, default=48)
ap.add_argument("--seed", type=int, default=7)
ap.add_argument("--k", type=int, default=11)
ap.add_argument("--cycles", type=int, default=20)
ap.add_argument("--circuit-seed", type=int, default=1)
ap.add_argument("--out", default=os.path.join(ROOT, "data", "sycamore53_m20.json"))
args = ap.parse_args()

from jet_b200.pathfinder import optimize, path_cost  # noqa: E402

sites, ops = sg.circuit(0, 0, (3, 2), args.cycles, args.circuit_seed, layout="sycamore")
import numpy as np  # noqa: E402

bits = np.random.default_rng(args.circuit_seed + 12345).integers(0, 2, len(sites)).tolist()
leaves = sg.to_network(sites, ops, bits)
leaf_idx = [idx for _, idx, _ in leaves]
dims = {i: 2 for idx in leaf_idx for i in idx}
rep = optimize(leaf_idx, dims, target_log2=args.target, max_slices_log2=60, trials=args.trials, seconds=args.seconds,
               seed=args.seed, k=args.k)
peak_s, flops_s = path_cost(leaf_idx, dims, rep["path"], rep["sliced"])
assert peak_s == rep["log2_peak_per_slice"] and flops_s == rep["jet_flops_per_slice"], (peak_s, flops_s, rep)
meta = dict(layout="sycamore", removed=[3, 2], cycles=args.cycles, seed=args.circuit_seed, qubits=len(sites), bits=bits,
            leaves=len(leaves), steps=len(rep["path"]), log2_peak_unsliced=rep["log2_peak_unsliced"],
            jet_flops_unsliced=rep["jet_flops_unsliced"], sliced_indices=rep["sliced"], log2_num_slices=rep["log2_slices"],
            log2_peak_per_slice=peak_s, jet_flops_per_slice=flops_s, jet_flops_total=rep["jet_flops_total"],
            optimizer=dict(tool="jet_b200/cpp/pathopt", target_log2=args.target, trials=args.trials, seed=args.seed, k=args.k,
                           seconds=rep["seconds"], trial_totals=sorted(rep["trial_totals"])[:8]))
with open(args.out, "w") as f:
    f.write(sg.network_json(leaves, rep["path"]))
json.dump(meta, open(args.out[:-5] + ".meta.json", "w"), indent=1)
print(json.dumps({k: v for k, v in meta.items() if k not in ("bits", "sliced_indices")}))
