#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -4 gpurun_out/r2t_pytest.log
B=jet_b200/cpp/tbc_bench; D=data/_ref
for api in tbc sliced; do
  timeout 300 $B $D/m10.json p7,s7,h4,m1,m2,I2 --api $api --reps 5 | cut -c1-330
  timeout 300 $B $D/m10.json p7,s7,h4,m1,m2,I2,V4,z2,t4,C1 --api $api --reps 5 | cut -c1-330
done
