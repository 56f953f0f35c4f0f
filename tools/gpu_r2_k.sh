#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_plan_gpu.py tests/test_jet_api_gpu.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; tail -4 gpurun_out/r2k_pytest.log
python - <<'PY'
import numpy as np, json, sys
sys.path.insert(0,'.')
import bench_micro, bench
peak,_=bench.measured_peaks()
for env in ("packed","scalar"):
    import os
    res=[]
    for c in (2,3,4):
        for r in (26,):
            rng = np.random.default_rng(100 * r + c)
            ia = list(range(r)); common = sorted(rng.choice(r, c, replace=False).tolist())
            ib = common + list(range(100, 100 + c)); ib = [ib[i] for i in rng.permutation(2 * c)]
            bench_micro.run_contract(np.complex64, r, ia, 2 * c, ib, 5, peak, "S1_skinny", res, check=False)
    break
PY
echo "--- scalar kernel"
JB_STREAM_NO_PACKED=1 python - <<'PY'
import numpy as np, json, sys
sys.path.insert(0,'.')
import bench_micro, bench
peak,_=bench.measured_peaks()
res=[]
for c in (3,4):
    r=26
    rng = np.random.default_rng(100 * r + c)
    ia = list(range(r)); common = sorted(rng.choice(r, c, replace=False).tolist())
    ib = common + list(range(100, 100 + c)); ib = [ib[i] for i in rng.permutation(2 * c)]
    bench_micro.run_contract(np.complex64, r, ia, 2 * c, ib, 5, peak, "S1_skinny", res, check=False)
PY
timeout 900 python tools/plan_profile.py sycamore53_m20 0 --top 6 2>&1 | head -14
