"""Per-slice time of a workload broken down by kernel family and (K, N) class.
  python tools/plan_classes.py [workload]"""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jet_b200 import ContractionPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "sycamore53_m12_s9"
net, sliced, dt, _ = bench.load_network(wl)
plan = ContractionPlan(net, sliced)
plan.reset()
plan.run(0, 1)
plan.sync()
prof = plan.profile(0, 3)
steps = plan.steps()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, s in enumerate(steps):
    if s.shared:
        continue
    big = max(s.m * s.k, s.k * s.n, s.m * s.n)
    key = (s.kernel, min(s.k, 1 << 20), min(s.m, s.n) if s.kernel == 0 else 0, "big" if big >= (1 << 22) else "small")
    a = agg[key]
    a[0] += 1
    a[1] += float(prof[i])
    a[2] += s.bytes
    a[3] += s.flops
tot = sum(a[1] for a in agg.values())
print(f"{wl}: per-slice sum of step times {tot:.3f} ms over {sum(a[0] for a in agg.values())} steps")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  kernel={key[0]} K={key[1]:>8} minMN={key[2]:>4} {key[3]:5s} n={a[0]:3d} ms={a[1]:8.3f} share={a[1] / tot:6.3f} "
          f"GB/s={a[2] / max(a[1], 1e-9) / 1e6:8.0f} TFLOP/s={a[3] / max(a[1], 1e-9) / 1e9:7.2f}")
plan.close()
