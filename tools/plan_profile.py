"""Per-launch-unit profile of one slice of a bench workload (CUDA events on the plan's stream):
    python tools/plan_profile.py <workload> [slice_id] [--top N] [--json FILE]
Prints plan statistics, the time per unit class (stream / chain / ttgt by GEMM kernel) and the slowest units."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jet_b200 import ContractionPlan  # noqa: E402
from jet_b200._lib import GEMM_KIND_NAMES  # noqa: E402

workload = sys.argv[1]
slice_id = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
net, sliced, dt, _ = bench.load_network(workload)
t0 = time.time()
plan = ContractionPlan(net, sliced)
t_create = time.time() - t0
st = plan.stats
print(json.dumps(dict(workload=workload, create_s=t_create, num_slices=int(st.num_slices), steps=int(st.steps_total),
                      shared=int(st.steps_shared), stream=int(st.steps_stream), ttgt=int(st.steps_ttgt), chained=int(st.steps_chained),
                      chains=int(st.chains), launches=int(st.launches_per_slice), flops=st.flops_per_slice, bytes=st.bytes_per_slice,
                      fused_bytes=st.fused_bytes_per_slice, arena_GiB=st.arena_bytes / 2**30, max_elems_log2=float(np.log2(max(st.max_step_elems, 1))))))
units = plan.ops()
steps = plan.steps()
ms = plan.profile_ops(slice_id, 3)
names = {0: "stream", 2: "chain"}
agg = {}
rows = []
for i, (u, t) in enumerate(zip(units, ms)):
    name = names.get(u.kernel) or ("ttgt:" + GEMM_KIND_NAMES[u.gemm_kind])
    a = agg.setdefault(name, dict(ms=0.0, flops=0.0, bytes=0.0, units=0, steps=0))
    a["ms"] += float(t); a["flops"] += u.flops; a["bytes"] += u.bytes; a["units"] += 1; a["steps"] += u.n_steps
    s0 = steps[u.first_step]
    rows.append((float(t), i, name, u.n_steps, (int(s0.m), int(s0.n), int(s0.k)), u.flops, u.bytes))
total = float(ms.sum())
print("total ms per slice (sum of units): %.3f -> %.2f slices/s; %.1f TFLOP/s useful" % (total, 1e3 / total, st.flops_per_slice / total / 1e9))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-24s %8.3f ms (%5.1f%%) units %3d steps %3d  %7.1f TFLOP/s  %7.1f GB/s" % (
        name, a["ms"], 100 * a["ms"] / total, a["units"], a["steps"], a["flops"] / a["ms"] / 1e9 if a["ms"] else 0,
        a["bytes"] / a["ms"] / 1e6 if a["ms"] else 0))
print("slowest units:")
for t, i, name, ns, mnk, fl, by in sorted(rows, reverse=True)[:top]:
    print("  #%3d %-22s %8.3f ms steps %2d first (m,n,k)=(2^%.0f,2^%.0f,2^%.0f) %7.1f TFLOP/s %7.1f GB/s" % (
        i, name, t, ns, np.log2(mnk[0]), np.log2(mnk[1]), np.log2(mnk[2]), fl / t / 1e9 if t else 0, by / t / 1e6 if t else 0))
if "--json" in sys.argv:
    # every unit with its steps, for offline fits of the kernel cost model (tools/chain_cost_fit.py)
    dump = []
    for (t, i, name, ns, mnk, fl, by) in sorted(rows, key=lambda r: r[1]):
        u = units[i]
        dump.append(dict(unit=i, name=name, ms=t, flops=fl, bytes=by, log_tile=int(u.log_tile), n_stages=int(u.n_stages),
                         register_steps=int(u.register_steps),
                         steps=[[int(steps[q].m), int(steps[q].n), int(steps[q].k)] for q in range(u.first_step, u.last_step + 1)
                                if steps[q].op == i]))
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
        json.dump(dict(workload=workload, units=dump), f)
# wall-clock through the graph
plan.reset(); plan.run(slice_id, 2); plan.sync()
plan.reset(); t0 = time.time(); plan.run(slice_id, 4); plan.sync(); dt_s = time.time() - t0
print("graph replay: %.3f ms per slice" % (dt_s * 1e3 / 4))
