#!/usr/bin/env python
"""Engine-aware selection of the m=20 contraction path.

For every (target width, seed): jet_b200/cpp/pathopt proposes a path + sliced indices; the engine plans it WITHOUT a GPU
(JB_PLAN_DRY_RUN) and the launch units are costed with a per-kernel model calibrated on B200 measurements
(fused chain: max(bytes / 3.2 TB/s, flops / 46 TFLOP/s); stream step: max(bytes / 3.0 TB/s, flops / 25 TFLOP/s); final dot
(DotGatherKernel): bytes / 3.0 TB/s; permute + GEMM step: max(3 x bytes / 4 TB/s, flops / 150 TFLOP/s); + launch latency).  Ranking = estimated GPU-seconds for the WHOLE
amplitude = ms per slice x number of slices.  One JSON line per candidate is appended to the log
(profiles/r2_m20_path_candidates.jsonl holds the 54 candidates of round 2; the model predicted 221 ms for the chosen
slice, 215 were measured).

    python tools/select_m20_path.py --targets 30,31 --seeds 1,2,3 [--log profiles/r2_m20_path_candidates.jsonl]
        [--write-best data/sycamore53_m20]      # network + path JSON and .meta.json of the best candidate in the log's run
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sycamore_gen as sg  # noqa: E402
from jet_b200 import ContractionPlan, NetworkFile  # noqa: E402
from jet_b200.pathfinder import optimize, path_cost  # noqa: E402


def estimate_ms(plan) -> float:
    t = 0.0
    for u in plan.ops():
        if u.kernel == 2:
            t += max(u.bytes / 3.2e12, u.flops / 46e12) + 4e-6
        elif u.kernel == 0:
            t += max(u.bytes / 3.0e12, u.flops / 25e12) + 4e-6
        elif u.gemm_kind in (1, 4):  # DOTU / GEMV corner: DotGatherKernel reads both operands once, in place
            t += u.bytes / 3.0e12 + 1e-5
        else:
            t += max(3 * u.bytes / 4e12, u.flops / 150e12) + 2e-5
    return t * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", default="30,31")
    ap.add_argument("--seeds", default="1,2,3")
    ap.add_argument("--trials", type=int, default=24)
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cycles", type=int, default=20)
    ap.add_argument("--circuit-seed", type=int, default=1)
    ap.add_argument("--max-arena-gib", type=float, default=150.0)
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "m20_path_candidates.jsonl"))
    ap.add_argument("--write-best", default="")
    args = ap.parse_args()

    sites, ops = sg.circuit(0, 0, (3, 2), args.cycles, args.circuit_seed, layout="sycamore")
    bits = np.random.default_rng(args.circuit_seed + 12345).integers(0, 2, len(sites)).tolist()
    leaves = sg.to_network(sites, ops, bits)
    leaf_idx = [idx for _, idx, _ in leaves]
    dims = {i: 2 for idx in leaf_idx for i in idx}
    tensors = [(idx, np.asarray(arr, dtype=np.complex64)) for _, idx, arr in leaves]
    best = None
    os.makedirs(os.path.dirname(args.log), exist_ok=True)
    for target in [int(a) for a in args.targets.split(",")]:
        for seed in [int(a) for a in args.seeds.split(",")]:
            rep = optimize(leaf_idx, dims, target_log2=target, max_slices_log2=60, trials=args.trials, seconds=args.seconds,
                           seed=seed, k=args.k, threads=args.threads)
            try:
                with ContractionPlan(NetworkFile(tensors, rep["path"]), rep["sliced"], dry_run=True) as plan:
                    ms, st = estimate_ms(plan), plan.stats
                    rec = dict(target=target, seed=seed, log2_slices=rep["log2_slices"], jet_total=rep["jet_flops_total"],
                               jet_slice=rep["jet_flops_per_slice"], est_ms=ms, est_total_gpu_s=ms * 1e-3 * 2 ** rep["log2_slices"],
                               chains=int(st.chains), units=len(plan.ops()), arena_GiB=st.arena_bytes / 2 ** 30,
                               fused_GB=st.fused_bytes_per_slice / 1e9)
            except Exception as e:  # a plan the engine refuses (rank limits ...) is not a candidate
                print("plan failed:", e, flush=True)
                continue
            print(json.dumps(rec), flush=True)
            with open(args.log, "a") as f:
                f.write(json.dumps(rec) + "\n")
            if rec["arena_GiB"] <= args.max_arena_gib and (best is None or rec["est_total_gpu_s"] < best[0]["est_total_gpu_s"]):
                best = (rec, rep)
    if args.write_best and best is not None:
        rec, rep = best
        path = [tuple(p) for p in rep["path"]]
        peak, flops = path_cost(leaf_idx, dims, path, rep["sliced"])
        with open(args.write_best + ".json", "w") as f:
            f.write(sg.network_json(leaves, path))
        meta = dict(layout="sycamore", removed=[3, 2], cycles=args.cycles, seed=args.circuit_seed, qubits=len(sites), bits=bits,
                    leaves=len(leaves), steps=len(path), log2_peak_unsliced=rep["log2_peak_unsliced"],
                    jet_flops_unsliced=rep["jet_flops_unsliced"], sliced_indices=rep["sliced"], log2_num_slices=rep["log2_slices"],
                    log2_peak_per_slice=peak, jet_flops_per_slice=flops, jet_flops_total=rep["jet_flops_total"],
                    optimizer=dict(tool="jet_b200/cpp/pathopt + tools/select_m20_path.py",
                                   args=f"--target {rec['target']} --max-slices 60 --trials {args.trials} --seconds {args.seconds} "
                                        f"--seed {rec['seed']} --k {args.k} --threads {args.threads}",
                                   est_ms_per_slice=rec["est_ms"], est_total_gpu_s=rec["est_total_gpu_s"], arena_GiB=rec["arena_GiB"]))
        json.dump(meta, open(args.write_best + ".meta.json", "w"), indent=1)
        print("wrote", args.write_best, json.dumps(rec))


if __name__ == "__main__":
    main()
