#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_plan_gpu.py tests/test_chain_gpu.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -4 gpurun_out/r2n_pytest.log
for c in 2 3 4; do python tools/prof_stream16.py $c 2>&1 | tail -1 | cut -c1-330; done
timeout 900 python tools/plan_profile.py sycamore53_m20 0 --top 4 2>&1 | sed -n 2,6p
