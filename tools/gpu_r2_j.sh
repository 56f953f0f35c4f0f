#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_m20_synth.py::test_m20_per_step_normwise_vs_reference > gpurun_out/r2j_pytest.log 2>&1; tail -6 gpurun_out/r2j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
