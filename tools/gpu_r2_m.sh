#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_plan_gpu.py tests/test_jet_api_gpu.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -6 gpurun_out/r2m_pytest.log
timeout 900 python tools/plan_profile.py sycamore53_m20 0 --top 4 2>&1 | head -12
timeout 900 python tools/plan_profile.py sycamore53_m12_s9 0 --top 3 2>&1 | head -9
