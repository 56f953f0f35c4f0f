"""Per-slice launch units of a workload with their device times (CUDA events per unit).
  python tools/plan_ops.py [workload] [--no-fuse]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jet_b200 import ContractionPlan  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = args[0] if args else "sycamore53_m12_s9"
net, sliced, dt, _ = bench.load_network(wl)
plan = ContractionPlan(net, sliced, fuse="--no-fuse" not in sys.argv)
plan.reset()
plan.run(0, 1)
plan.sync()
ms = plan.profile_ops(0, 3)
ops = plan.ops()
tot = float(ms.sum())
st = plan.stats
print(f"{wl}: {len(ops)} launch units / {st.steps_total - st.steps_shared} per-slice steps, chains={st.chains} "
      f"steps_chained={st.steps_chained}; sum of unit times {tot:.3f} ms; step bytes {st.bytes_per_slice / 1e9:.2f} GB, "
      f"fused bytes {st.fused_bytes_per_slice / 1e9:.2f} GB, flops {st.flops_per_slice / 1e9:.1f} GF")
names = {0: "stream", 1: "ttgt", 2: "chain"}
for i, (o, t) in enumerate(zip(ops, ms)):
    if t < 0.02 * tot / max(len(ops), 1) and t < 0.02:
        continue
    print(f"  unit {i:3d} {names[o.kernel]:6s} steps {o.first_step:3d}..{o.last_step:3d} (n={o.n_steps}) tile 2^{o.log_tile:<2d} stages={o.n_stages:<2d} "
          f"ms={t:7.3f} fusedGB/s={o.bytes / max(t, 1e-9) / 1e6:7.0f} stepGB/s={o.step_bytes / max(t, 1e-9) / 1e6:7.0f} "
          f"TFLOP/s={o.flops / max(t, 1e-9) / 1e9:6.2f}")
plan.close()
