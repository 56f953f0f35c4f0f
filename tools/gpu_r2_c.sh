#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_plan_gpu.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
timeout 900 python tools/plan_profile.py sycamore53_m20 0 --top 14 > gpurun_out/r2c_m20_profile.txt 2>&1; head -24 gpurun_out/r2c_m20_profile.txt
timeout 900 python tools/plan_profile.py sycamore53_m12_s9 0 --top 8 > gpurun_out/r2c_m12_profile.txt 2>&1; head -16 gpurun_out/r2c_m12_profile.txt
JB_CHAIN_NO_PAD=1 timeout 900 python tools/plan_profile.py sycamore53_m12_s9 0 --top 3 > gpurun_out/r2c_m12_profile_nopad.txt 2>&1; head -6 gpurun_out/r2c_m12_profile_nopad.txt
