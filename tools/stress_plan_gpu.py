"""Random tensor networks (random graph, random path, random sliced indices, optional open indices) through
jb_plan_* on the GPU against the numpy oracle — exercises deferred slicing, fused chains of shared and per-slice
steps, slice views and the FP64 accumulation on shapes no fixture covers.
  python tools/stress_plan_gpu.py [trials] [only_trial | -1] [seed] [big]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ContractionPlan, LanePlans, NetworkFile  # noqa: E402
from oracle import jet_oracle as jo  # noqa: E402  (checker)

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 150
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1  # run just this trial (same random sequence), verbosely
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 31337
big = len(sys.argv) > 4 and sys.argv[4] == "big"  # larger leaves / intermediates: reaches TTGT, tcgen05, DMMA
rng = np.random.default_rng(seed)
bad = ran = 0
for trial in range(trials):
    dtype = np.complex64 if rng.integers(0, 2) else np.complex128
    dims = int(rng.choice([2, 2, 4, 8] if big else [2, 2, 2, 4]))
    lgd = {2: 1, 4: 2, 8: 3}[dims]
    nt = int(rng.integers(3, 9 if big else 15))
    # random connected multigraph: a spanning chain plus extra edges; every edge is an index shared by 2 tensors
    edges = [(i, i + 1) for i in range(nt - 1)]
    for _ in range(int(rng.integers(0, (3 * nt if big else nt) + 3))):
        u, v = rng.choice(nt, 2, replace=False)
        edges.append((int(u), int(v)))
    idx_of = [[] for _ in range(nt)]
    for e, (u, v) in enumerate(edges):
        idx_of[u].append(f"e{e}")
        idx_of[v].append(f"e{e}")
    n_open = int(rng.integers(0, 3)) if rng.integers(0, 4) == 0 else 0
    for o in range(n_open):
        idx_of[int(rng.integers(0, nt))].append(f"o{o}")
    if max(len(x) for x in idx_of) * lgd > (18 if big else 12):
        continue
    real = np.float32 if dtype == np.complex64 else np.float64
    tensors = []
    for t in range(nt):
        idx = list(idx_of[t])
        rng.shuffle(idx)
        n = dims ** len(idx)
        arr = (rng.uniform(-1, 1, n).astype(real) + 1j * rng.uniform(-1, 1, n).astype(real)).astype(dtype)
        tensors.append((idx, arr.reshape([dims] * len(idx))))
    # random path: contract two random live nodes that share an index when possible
    live = list(range(nt))
    node_idx = [set(x) for x in idx_of]
    path = []
    too_big = False
    while len(live) > 1:
        pairs = [(a, b) for i, a in enumerate(live) for b in live[i + 1:] if node_idx[a] & node_idx[b]]
        a, b = pairs[int(rng.integers(0, len(pairs)))] if pairs else (live[0], live[1])
        if rng.integers(0, 2):
            a, b = b, a
        new = node_idx[a] ^ node_idx[b]
        if len(new) * lgd > (23 if big else 20):
            too_big = True
            break
        path.append((a, b))
        node_idx.append(new)
        live = [x for x in live if x not in (a, b)] + [len(node_idx) - 1]
    if too_big:
        continue
    closed = [f"e{e}" for e in range(len(edges))]
    ns = int(rng.integers(0, min(4, len(closed)) + 1))
    sliced = [str(s) for s in rng.choice(closed, ns, replace=False)] if ns else []
    use_lanes = bool(rng.integers(0, 3) == 0 and sliced)
    use_graph = bool(rng.integers(0, 2))
    if only >= 0:
        if trial != only:
            continue
        print("trial", trial, dtype.__name__, "dims", dims, "tensors", [(i, a.shape) for i, a in tensors], "path", path,
              "sliced", sliced, "lanes", use_lanes, "graph", use_graph, flush=True)
    onet = jo.Network(tensors, path)
    ref = np.asarray(jo.amplitude(onet, sliced)).reshape(-1)
    tol = 1e-5 if dtype == np.complex64 else 1e-12
    try:
        if use_lanes:
            ctx = LanePlans(NetworkFile(tensors, path), sliced, lanes=2)
        else:
            ctx = ContractionPlan(NetworkFile(tensors, path), sliced, use_graph=use_graph)
        with ctx as plan:
            got = np.asarray(plan.amplitude()).reshape(-1)
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("EXC", trial, dtype.__name__, dims, nt, sliced, e)
        continue
    ran += 1
    err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
    if got.shape != ref.shape or err >= 20 * tol:
        bad += 1
        print("FAIL", trial, dtype.__name__, dims, nt, len(edges), sliced, n_open, err)
print("ran", ran, "bad", bad)
