#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_m20_synth.py -m gpu -x -q -s -k "round2" > gpurun_out/r2o_pytest.log 2>&1; tail -12 gpurun_out/r2o_pytest.log
