"""GPU probe: bandwidth of K1 / the streaming contraction kernel on representative shapes, and plan
timings for m10 / m12.  Run on the B200 box:  python tools/gpu_probe.py [--m12]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jet_b200 import ContractionPlan, NetworkFile, ops  # noqa: E402

DATA = os.path.join(ROOT, "data", "_ref")
dev = torch.device("cuda:0")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def probe_permute(out):
    rng = np.random.default_rng(0)
    for dtype, tdt, eb in ((np.complex64, torch.complex64, 8), (np.complex128, torch.complex128, 16)):
        for r in (20, 24, 27):
            n = 2 ** r
            x = torch.empty(n, dtype=tdt, device=dev).normal_()
            y = torch.empty_like(x)
            pull = sorted(rng.choice(r, 3, replace=False).tolist())
            rest = [i for i in range(r) if i not in pull]
            perms = {"P1_pull_back": rest + pull, "P2_pull_front": pull + rest, "P3_random": rng.permutation(r).tolist(),
                     "P4_reverse": list(range(r))[::-1],
                     "P5_last5_fixed": rng.permutation(r - 5).tolist() + list(range(r - 5, r))}
            for name, perm in perms.items():
                ms = timeit(lambda: ops.permute_device(dtype, x.data_ptr(), y.data_ptr(), [2] * r, perm))
                gbs = 2 * n * eb / ms / 1e6
                out.append(dict(kind="permute", dtype=str(np.dtype(dtype)), rank=r, pattern=name, ms=ms, GBs=gbs))
                print(out[-1], flush=True)
            ms = timeit(lambda: y.copy_(x))
            print(dict(kind="torch_copy", rank=r, dtype=str(np.dtype(dtype)), ms=ms, GBs=2 * n * eb / ms / 1e6), flush=True)


def probe_contract(out):
    rng = np.random.default_rng(1)
    for dtype, tdt, eb in ((np.complex64, torch.complex64, 8), (np.complex128, torch.complex128, 16)):
        for ra in (20, 26):
            for rb, c in ((2, 1), (4, 2), (6, 3), (8, 4), (3, 2), (5, 3)):
                a = torch.empty(2 ** ra, dtype=tdt, device=dev).normal_()
                b = torch.empty(2 ** rb, dtype=tdt, device=dev).normal_()
                ia = list(range(ra))
                common = sorted(rng.choice(ra, c, replace=False).tolist())
                ib = common + list(range(100, 100 + rb - c))
                for swap in (False, True):
                    m = 2 ** (ra - c)
                    nn = 2 ** (rb - c)
                    cbuf = torch.empty(m * nn, dtype=tdt, device=dev)
                    if swap:
                        fn = lambda: ops.contract_device(dtype, [2] * rb, ib, b.data_ptr(), [2] * ra, ia, a.data_ptr(), cbuf.data_ptr())
                    else:
                        fn = lambda: ops.contract_device(dtype, [2] * ra, ia, a.data_ptr(), [2] * rb, ib, b.data_ptr(), cbuf.data_ptr())
                    ms = timeit(fn)
                    byt = eb * (2 ** ra + 2 ** rb + m * nn)
                    fl = 8.0 * m * nn * 2 ** c
                    out.append(dict(kind="contract", dtype=str(np.dtype(dtype)), ra=ra, rb=rb, c=c, swap=swap, common=common,
                                    ms=ms, GBs=byt / ms / 1e6, TFLOPs=fl / ms / 1e9))
                    print(out[-1], flush=True)
                del a, b, cbuf


def probe_plan(out, name, sliced, nrun, dtype=np.complex64):
    net = NetworkFile.load(os.path.join(DATA, name), dtype)
    t0 = time.time()
    plan = ContractionPlan(net, sliced)
    st = plan.stats
    print(name, "plan built in %.2fs" % (time.time() - t0), dict(slices=st.num_slices, shared=st.steps_shared, stream=st.steps_stream,
          ttgt=st.steps_ttgt, launches=st.launches_per_slice, flops=st.flops_per_slice, bytes=st.bytes_per_slice,
          arena_MiB=st.arena_bytes >> 20), flush=True)
    plan.reset()
    plan.run(0, min(2, plan.num_slices))
    plan.sync()
    plan.reset()
    plan.run(0, nrun)
    res = plan.result()
    ms = plan.last_ms()
    rec = dict(kind="plan", name=name, sliced=len(sliced), slices_run=nrun, ms=ms, ms_per_slice=ms / nrun,
               slices_per_s=nrun / ms * 1e3, GBs=st.bytes_per_slice * nrun / ms / 1e6, TFLOPs=st.flops_per_slice * nrun / ms / 1e9,
               result=[float(res.reshape(-1)[0].real), float(res.reshape(-1)[0].imag)])
    out.append(rec)
    print(rec, flush=True)
    prof = plan.profile(0, 2)
    steps = plan.steps()
    top = sorted(range(len(steps)), key=lambda i: -prof[i])[:12]
    tot = float(prof.sum())
    print("  sum of per-step ms: %.3f" % tot)
    for i in top:
        s = steps[i]
        print("   step %d kernel=%d m=%d n=%d k=%d ms=%.3f GB/s=%.0f" % (i, s.kernel, s.m, s.n, s.k, prof[i], s.bytes / max(prof[i], 1e-6) / 1e6))
    plan.close()


if __name__ == "__main__":
    out = []
    print(ops.version(), ops.device_info(0))
    if "--no-micro" not in sys.argv:
        probe_permute(out)
        probe_contract(out)
    m10 = "p7 s7 h4 m1 m2 I2 V4 z2 t4 C1".split()
    probe_plan(out, "m10.json", [], 1)
    probe_plan(out, "m10.json", m10[:6], 64)
    probe_plan(out, "gbs_dim2_nc1_lw8_rp5_fock4_total10_0.kraken.json", [], 1, np.complex128)
    if "--m12" in sys.argv:
        probe_plan(out, "m12.json", "h5 m H10 w y J S G10 P0".split(), 4)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
