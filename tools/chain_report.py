"""Which path steps of a workload fuse into chains, and what that does to the traffic of a slice.
Planning only (jb_chain_info needs no GPU):  python tools/chain_report.py [workload]"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jet_b200 import JetB200Error, ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "sycamore53_m12_s9"
net, sliced, dt, _ = bench.load_network(wl)
eb = np.dtype(dt).itemsize
dims = net.index_dims()
labels = {}
for idx, _ in net.tensors:
    for i in idx:
        labels.setdefault(i, len(labels))
sl = set(sliced)
nodes = [[i for i in idx if i not in sl] for idx, _ in net.tensors]
dep = [any(i in sl for i in idx) for idx, _ in net.tensors]
steps = []
for a, b in net.path:
    A, B = nodes[a], nodes[b]
    common = [i for i in A if i in B]
    nodes.append([i for i in A if i not in B] + [i for i in B if i not in A])
    dep.append(dep[a] or dep[b])
    steps.append((a, b, len(nodes) - 1, math.prod(dims[i] for i in common)))
steps_dep = [dep[c] for _, _, c, _ in steps]
steps_dep[-1] = True
size = lambda n: math.prod(dims[i] for i in nodes[n])
consumer = {}
for s, (a, b, c, k) in enumerate(steps):
    consumer[a] = s
    consumer[b] = s
taken = set()
total_step = total_fused = 0.0
launches = 0
for s, (a, b, c, k) in enumerate(steps):
    if not steps_dep[s] or s in taken:
        continue
    x0 = a if size(a) >= size(b) else b
    operands, chain = [], []
    x, cur = x0, s
    info_ok = None
    while cur is not None and cur not in taken and len(operands) < 12:
        ca, cb, cc, ck = steps[cur]
        left = ca == x
        r = cb if left else ca
        free_r = size(r) // ck
        if ck > 16 or free_r > 16:
            break
        operands.append(([dims[i] for i in nodes[r]], [labels[i] for i in nodes[r]], left))
        try:
            info = ops.chain_info(dt, [dims[i] for i in nodes[x0]], [labels[i] for i in nodes[x0]], operands)
        except JetB200Error:
            operands.pop()
            break
        info_ok = info
        chain.append(cur)
        x = cc
        cur = consumer.get(x)
    if len(chain) >= 2:
        taken.update(chain)
        total_step += info_ok.step_bytes
        total_fused += info_ok.bytes
        launches += 1
        kn = " ".join(f"{int(math.log2(steps[c][3]))}:{int(math.log2(size(steps[c][1] if steps[c][0] in (x0,) or size(steps[c][0]) >= size(steps[c][1]) else steps[c][0]) // steps[c][3]))}" for c in chain)
        print(f"chain of {len(chain):2d} at step {chain[0]:3d}: log2|X0|={int(math.log2(size(x0))):2d} -> "
              f"log2|Xk|={int(math.log2(size(steps[chain[-1]][2]))):2d}  tile 2^{info_ok.log_tile:2d} "
              f"cf={info_ok.conflict_free} stages={info_ok.n_stages} logK:logN [{kn}]  bytes {info_ok.step_bytes / 1e9:7.3f} -> {info_ok.bytes / 1e9:7.3f} GB")
    else:
        taken.add(s)
        bts = eb * (size(a) + size(b) + size(c))
        total_step += bts
        total_fused += bts
        launches += 1
        if bts > 1e8:
            print(f"single step {s}: log2 sizes {math.log2(size(a)):.0f} {math.log2(size(b)):.0f} K={k} bytes {bts / 1e9:.3f} GB")
print(f"{wl}: per-slice step bytes {total_step / 1e9:.2f} GB -> fused {total_fused / 1e9:.2f} GB "
      f"({total_step / total_fused:.2f}x), {launches} launch units for {sum(steps_dep)} steps")
