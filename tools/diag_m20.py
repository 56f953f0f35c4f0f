"""Diagnostic: m=20 slice amplitudes under c64 (fused / unfused / no tensor cores) and c128 vs the
reference golden (GPU)."""
import json, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ContractionPlan, NetworkFile
DATA = os.path.join(ROOT, "data")
gold = json.load(open(os.path.join(DATA, "syc53_m20_seed1.golden.json")))
meta = json.load(open(os.path.join(DATA, "syc53_m20_seed1.meta.json")))
ids = [int(k) for k in gold]
mode = sys.argv[1]
dt = np.complex128 if mode == "c128" else np.complex64
net = NetworkFile.load(os.path.join(DATA, "syc53_m20_seed1.json"), dt)
with ContractionPlan(net, meta["sliced_indices"], store_results=True, fuse=(mode != "nofuse")) as plan:
    plan.reset(); plan.run_list(ids)
    for n, k in enumerate(gold):
        got = complex(plan.slice_result(n).reshape(-1)[0])
        want = complex(gold[k]["re"], gold[k]["im"])
        print(mode, k, repr(got), "rel-to-golden %.3e" % (abs(got - want) / abs(want)), flush=True)
