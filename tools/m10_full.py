"""BASELINE config 1 on the GPU: the reference's shipped Sycamore-53 m=10 network + path, UNSLICED, single
amplitude, complex64 — time from pinned host leaves to the amplitude on the host (upload + contraction +
read-back), compared with the reference golden (tests/golden/amplitudes.json).
  python tools/m10_full.py [reps]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jet_b200 import ContractionPlan, NetworkFile  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
net = NetworkFile.load(bench.data_path("m10.json"), np.complex64)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))
with ContractionPlan(net, []) as plan:
    st = plan.stats
    ts = []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        plan.upload()
        plan.reset()
        plan.run(0, 1)
        amp = plan.result().reshape(-1)[0]
        ts.append(time.perf_counter() - t0)
    ts = ts[1:]
    want = None
    for k, v in gold.items():
        if k.startswith("m10_full") and "complex64" in k:
            want = complex(v["re"], v["im"])
    err = abs(amp - want) / abs(want) if want is not None else None
    flops = st.flops_shared + st.flops_per_slice
    print(json.dumps({"workload": "sycamore53_m10_full", "amplitude": [amp.real, amp.imag], "rel_err_vs_reference": err,
                      "seconds_median": float(np.median(ts)), "seconds_min": float(min(ts)),
                      "real_TFLOPs": flops / float(np.median(ts)) / 1e12, "arena_GiB": st.arena_bytes / 2 ** 30,
                      "steps": int(st.steps_total)}))
