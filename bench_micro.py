#!/usr/bin/env python
"""BASELINE config 2: Permuter transpose + ContractTensors microbench sweep on 1 GPU
(rank 10-30 tensors of dim 2, complex64 / complex128) — SURVEY.md §8(d).

  python bench_micro.py [--max-rank 30] [--quick] [--out profiles/r1_microbench.json]

Per case: median CUDA-event time over `reps` launches after warm-up (inputs larger than L2 from
rank 24 up; below that an L2 flush — a 256 MiB memset — runs between timed launches), algorithmic
bytes / flops, achieved GB/s and TFLOP/s, fraction of the HBM roofline, and a device-side sanity
check against torch (bit-exact for permutations; 1e-5 / 1e-12 relative for contractions).
The oracle-based parity tests live in tests/; this script measures.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from jet_b200 import ops  # noqa: E402
from bench import measured_peaks  # noqa: E402

dev = torch.device("cuda:0")
TDT = {np.complex64: torch.complex64, np.complex128: torch.complex128}
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _flush.zero_()


def timeit(fn, reps, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def perm_patterns(r, rng):
    c = min(3, r - 1)
    pull = sorted(rng.choice(r, c, replace=False).tolist())
    rest = [i for i in range(r) if i not in pull]
    pats = {"P1_pull_back": rest + pull, "P2_pull_front": pull + rest, "P3_random": rng.permutation(r).tolist(),
            "P4_reverse": list(range(r))[::-1]}
    if r > 6:
        pats["P5_last5_fixed"] = rng.permutation(r - 5).tolist() + list(range(r - 5, r))
    return pats


def sweep_permute(ranks, reps, peak, out):
    for dtype in (np.complex64, np.complex128):
        eb = np.dtype(dtype).itemsize
        for r in ranks:
            n = 2 ** r
            if 2 * n * eb > 100e9:
                continue
            rng = np.random.default_rng(1000 * r)
            x = torch.empty(n, dtype=TDT[dtype], device=dev)
            x.view(torch.float32 if dtype == np.complex64 else torch.float64).normal_()
            y = torch.empty_like(x)
            for name, perm in perm_patterns(r, rng).items():
                ms = timeit(lambda: ops.permute_device(dtype, x.data_ptr(), y.data_ptr(), [2] * r, perm), reps, r < 24)
                ok = None
                if r <= 24:
                    ok = bool(torch.equal(y.view([2] * r), x.view([2] * r).permute(perm)))
                gbs = 2 * n * eb / ms / 1e6
                rec = dict(kind="permute", dtype=np.dtype(dtype).name, rank=r, pattern=name, ms=ms, bytes=2 * n * eb,
                           GBs=gbs, frac_hbm=gbs / peak, bit_exact_vs_torch=ok)
                out.append(rec)
                print(json.dumps(rec), flush=True)
            del x, y
            torch.cuda.empty_cache()


def run_contract(dtype, ra, ia, rb, ib, reps, peak, kind, out, check=True):
    eb = np.dtype(dtype).itemsize
    a = torch.empty(2 ** ra, dtype=TDT[dtype], device=dev)
    b = torch.empty(2 ** rb, dtype=TDT[dtype], device=dev)
    for t in (a, b):
        t.view(torch.float32 if dtype == np.complex64 else torch.float64).uniform_(-1, 1)
    info = ops.contract_info(dtype, [2] * ra, ia, [2] * rb, ib)
    m, n, k = int(info.m), int(info.n), int(info.k)
    c = torch.empty(m * n, dtype=TDT[dtype], device=dev)
    ws = torch.empty(max(int(info.ws_bytes), 16), dtype=torch.uint8, device=dev)
    fn = lambda: ops.contract_device(dtype, [2] * ra, ia, a.data_ptr(), [2] * rb, ib, b.data_ptr(), c.data_ptr(),
                                     ws.data_ptr(), int(info.ws_bytes))
    byt = eb * (2 ** ra + 2 ** rb + m * n)
    ms = timeit(fn, reps, byt < 256e6)
    err = None
    if check and ra <= 24 and rb <= 24 and info.rank_c <= 24 and (2 ** ra + 2 ** rb + m * n) * 16 < 20e9:
        letters = {}
        sym = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
        for i in list(ia) + list(ib):
            letters.setdefault(i, sym[len(letters)])
        sa = "".join(letters[i] for i in ia)
        sb = "".join(letters[i] for i in ib)
        sc = "".join(letters[info.modes_c[i]] for i in range(info.rank_c))
        ref = torch.einsum(f"{sa},{sb}->{sc}", a.view([2] * ra).to(torch.complex128), b.view([2] * rb).to(torch.complex128))
        err = float((c.view(ref.shape).to(torch.complex128) - ref).norm() / ref.norm())
        del ref
    fl = 8.0 * m * n * k
    rec = dict(kind=kind, dtype=np.dtype(dtype).name, rank_a=ra, rank_b=rb, m=m, n=n, k=k, kernel=int(info.kernel), ms=ms,
               bytes=byt, flops=fl, GBs=byt / ms / 1e6, TFLOPs=fl / ms / 1e9, frac_hbm=byt / ms / 1e6 / peak,
               rel_err_vs_torch_c128=err)
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del a, b, c, ws
    torch.cuda.empty_cache()


def sweep_contract(ranks, reps, peak, out, max_square_rank):
    for dtype in (np.complex64, np.complex128):
        eb = np.dtype(dtype).itemsize
        # S1 skinny TN-like: A rank r, B rank 2c with c common + c free at random positions
        for r in ranks:
            if 2 * (2 ** r) * eb * 8 > 150e9:
                continue
            for c in (1, 2, 3, 4):
                if c >= r:
                    continue
                if (2 ** r + 2 ** (r)) * eb > 60e9:
                    continue
                rng = np.random.default_rng(100 * r + c)
                ia = list(range(r))
                common = sorted(rng.choice(r, c, replace=False).tolist())
                ib = common + list(range(100, 100 + c))
                ib = [ib[i] for i in rng.permutation(2 * c)]
                run_contract(dtype, r, ia, 2 * c, ib, reps, peak, "S1_skinny", out, check=r <= 24)
        # S2 square: A, B rank r, c = r/2 common -> M = N = K = 2^(r/2)
        for r in ranks:
            if r > max_square_rank or r % 2:
                continue
            rng = np.random.default_rng(7 * r)
            half = r // 2
            ia = list(range(r))
            common = sorted(rng.choice(r, half, replace=False).tolist())
            ib = common + list(range(100, 100 + half))
            ib = [ib[i] for i in rng.permutation(r)]
            run_contract(dtype, r, ia, r, ib, max(2, reps // 2), peak, "S2_square", out, check=r <= 22)
        # S3 split-K: M = 2^a, N = 2^b small, K = 2^(r - a); includes the m10 step (16, 32, 2^21)
        for (a_, b_, kk) in ((4, 5, 21), (2, 3, 20), (0, 0, 22), (0, 4, 18), (5, 0, 18)):
            ra, rb = a_ + kk, b_ + kk
            rng = np.random.default_rng(kk)
            ia = list(range(ra))
            common = sorted(rng.choice(ra, kk, replace=False).tolist())
            ib = common + list(range(100, 100 + b_))
            ib = [ib[i] for i in rng.permutation(rb)]
            run_contract(dtype, ra, ia, rb, ib, reps, peak, "S3_splitK", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-rank", type=int, default=30)
    ap.add_argument("--max-square-rank", type=int, default=26)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    args = ap.parse_args()
    peak, src = measured_peaks()
    ranks = [r for r in range(10, args.max_rank + 1, 2)]
    if args.quick:
        ranks = [r for r in (12, 20, 26) if r <= args.max_rank]
    out = []
    print(json.dumps(dict(device=ops.device_info(0), version=ops.version(), hbm_peak_GBs=peak, peak_source=src)), flush=True)
    sweep_permute(ranks, args.reps, peak, out)
    sweep_contract(ranks, args.reps, peak, out, args.max_square_rank)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(hbm_peak_GBs=peak, peak_source=src, results=out), open(args.out, "w"), indent=1)
    # compact summary
    big = [r for r in out if r["kind"] == "permute" and r["rank"] >= 26]
    if big:
        print("permute rank>=26: min %.0f / median %.0f / max %.0f GB/s" % (
            min(r["GBs"] for r in big), float(np.median([r["GBs"] for r in big])), max(r["GBs"] for r in big)))


if __name__ == "__main__":
    main()
