"""Run one slice of a workload unit by unit (for ncu captures of the per-slice kernels).
  python tools/prof_plan.py [workload] [--no-graph] [--no-fuse] [--reps N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jet_b200 import ContractionPlan  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = args[0] if args else "sycamore53_m12_s9"
net, sliced, dt, _ = bench.load_network(wl)
plan = ContractionPlan(net, sliced, fuse="--no-fuse" not in sys.argv, use_graph="--no-graph" not in sys.argv)
plan.reset()
plan.run(0, 1)
plan.sync()
print("slice 0 done", plan.result().reshape(-1)[:1], flush=True)
ms = plan.profile_ops(0, 1)
print("profile_ops done, sum ms", float(ms.sum()), flush=True)
plan.close()
