"""Generate data/<stem>.golden.json (default stem: sycamore53_m20): slices of the synthetic m=20 network contracted by
the UNMODIFIED reference (oracle/_ref: TaskBasedContractor + reduction + deletion tasks) in complex64 and complex128.

A slice of the GPU workload holds tensors of 2^30-2^31 elements, which the reference's CPU path cannot hold (one
std::vector per task, complex128) or takes hours on; so the reference evaluates the SAME number another way: with
the workload's sliced indices fixed to the slice's digits, jet_b200/cpp/pathopt finds a path and `extra` more
indices to slice for a small peak (the value of a slice does not depend on the contraction order), and the 2^extra
sub-slices go through the reference's own sliced flow (jet_sliced.cpp: SliceIndices copies -> AddContractionTasks
-> AddReductionTask -> Contract); the reduction result IS the slice amplitude — an identity of the contraction,
not an approximation.  Run in the container that has /root/reference:
    python tools/make_m20_golden.py [--stem sycamore53_m20] [--peak-log2 24] [--dtypes complex128] [--sub-count 8] [slice ids ...]
The complex128 value is the truth the tests compare against (1e-12 for the complex128 engine; the complex64 engine
is compared with it at the accuracy the conditioning of the sum allows, see tests/test_m20_synth.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200.pathfinder import optimize  # noqa: E402
from oracle import ref  # noqa: E402

DATA = os.path.join(ROOT, "data")
argv = sys.argv[1:]
stem, peak_log2, only, plan_only, sub_count = "sycamore53_m20", 24, [], False, 0
while argv and argv[0].startswith("--"):
    if argv[0] == "--stem":
        stem, argv = argv[1], argv[2:]
    elif argv[0] == "--peak-log2":
        peak_log2, argv = int(argv[1]), argv[2:]
    elif argv[0] == "--dtypes":
        only, argv = argv[1].split(","), argv[2:]
    elif argv[0] == "--plan-only":
        plan_only, argv = True, argv[1:]
    elif argv[0] == "--sub-count":
        sub_count, argv = int(argv[1]), argv[2:]
    else:
        raise SystemExit("unknown option " + argv[0])
ids = [int(a) for a in argv] or [0, 1234567]
meta = json.load(open(os.path.join(DATA, stem + ".meta.json")))
text = open(os.path.join(DATA, stem + ".json")).read()
js = json.loads(text)
leaf = [t[1] for t in js["tensors"]]
dims = {i: d for t in js["tensors"] for i, d in zip(t[1], t[2])}
path = [tuple(p) for p in js["path"]]
sliced = list(meta["sliced_indices"])
rep = optimize(leaf, dims, target_log2=peak_log2, max_slices_log2=60, trials=32, seconds=120, seed=3, k=10,
               fixed_slices=sliced)
full = list(rep["sliced"])
assert full[:len(sliced)] == sliced, "the workload's sliced indices must stay the slowest digits"
extra = full[len(sliced):]
n_sub = 1
for i in extra:
    n_sub *= dims[i]
js["path"] = [list(p) for p in rep["path"]]
text = json.dumps(js, separators=(",", ":"))
# the alternative evaluation order, committed so that tests can recompute a slice the same way on the GPU in complex128
json.dump({"path": js["path"], "sliced_indices": full, "extra_indices": extra, "sub_slices": n_sub,
           "log2_peak": rep["log2_peak_per_slice"], "jet_flops_per_sub_slice": rep["jet_flops_per_slice"]},
          open(os.path.join(DATA, stem + ".subslice.json"), "w"))
print(f"{stem}: {len(sliced)} sliced indices + {len(extra)} extra {extra} -> {n_sub} sub-slices of peak 2^{rep['log2_peak_per_slice']} "
      f"elements, {rep['jet_flops_per_slice']:.3g} Jet-flops each ({n_sub * rep['jet_flops_per_slice']:.3g} per slice; the workload's "
      f"own path: {meta['jet_flops_per_slice']:.3g})", flush=True)
out_path = os.path.join(DATA, stem + ".golden.json")
gold = json.load(open(out_path)) if os.path.exists(out_path) else {}
ref.set_blas_threads(1)


def _one(args):
    """One sub-slice through the reference's TaskBasedContractor (+ deletion tasks: bounded memory), 2 threads."""
    dt, sub_id = args
    r, sec, fl = ref.network(text, dt, full, sub_id, mode=2, threads=2, num_slices=1)
    return complex(r[0]), sec, fl


if __name__ == "__main__" and not plan_only:
    import multiprocessing as mp
    import time

    workers = max(1, min(4, (os.cpu_count() or 2) // 2))
    for v in ids:
        e = gold.setdefault(str(v), {})
        # --sub-count n: the reference evaluates only the first n sub-slices (one sub-slice of peak 2^26 costs minutes of
        # CPU); tests then compare that PARTIAL sum with the same partial sum on the GPU
        count = sub_count if sub_count > 0 else n_sub
        e["extra_indices"], e["sub_slices"], e["sub_first"], e["sub_count"] = extra, n_sub, 0, count
        for dt, key in (("complex64", ""), ("complex128", "_c128")):
            if ("re" + key) in e or (only and dt not in only):
                continue
            t0 = time.time()
            with mp.Pool(workers) as pool:
                parts = pool.map(_one, [(dt, v * n_sub + j) for j in range(count)], chunksize=1)
            total = sum(p[0] for p in parts)  # summed in complex128 (Python complex), sub-slice order
            e["re" + key], e["im" + key] = float(total.real), float(total.imag)
            e["jet_flops_reference"], e["ref_seconds_here" + key] = sum(p[2] for p in parts), time.time() - t0
            print(v, dt, total, f"{time.time() - t0:.1f} s", flush=True)
            json.dump(gold, open(out_path, "w"), indent=1)
