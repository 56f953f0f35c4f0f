"""complex64 GEMM with a short K (the one TTGT step of an m=20 slice: M = 2^21, N = 256, K = 32): FMA GemmKernel vs the
tcgen05 3xTF32 kernel (JB_TC_MIN_K=32), time and error against float64.  python tools/gpu/tc_short_k.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from jet_b200 import ops
    rng = np.random.default_rng(3)
    for (m, n, k) in [(1 << 14, 256, 32), (1 << 21, 256, 32), (1 << 21, 256, 48)]:
        a = (rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k))).astype(np.complex64)
        b = (rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))).astype(np.complex64)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        dc = torch.empty((m, n), dtype=torch.complex64, device="cuda")
        wsb = ops.gemm_ws_bytes(np.complex64, m, n, k)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
        run = lambda: ops.gemm_device(np.complex64, m, n, k, da.data_ptr(), db.data_ptr(), dc.data_ptr(), ws.data_ptr(), wsb, 0)
        run()
        times = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        rows = slice(0, 256)
        want = a[rows].astype(np.complex128) @ b.astype(np.complex128)
        got = dc[rows].cpu().numpy()
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        print(f"m=2^{int(np.log2(m))} n={n} k={k}: {min(times):.3f} ms  {8.0 * m * n * k / min(times) / 1e9:.1f} TFLOP/s  rel err {err:.2e}")
else:
    for env in ({}, {"JB_TC_MIN_K": "32"}):
        print("env", env, flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **env), check=False)
