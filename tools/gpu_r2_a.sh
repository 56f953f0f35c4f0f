#!/bin/bash
# round-2 GPU check A: full gpu test-suite, the C++ drop-in binary, TaskBasedContractor-vs-SlicedContractor timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
B=jet_b200/cpp/tbc_bench
D=data/_ref
{
for api in tbc sliced; do
  timeout 300 $B $D/m10.json p7,s7,h4,m1,m2,I2 --api $api --reps 3
  timeout 300 $B $D/m10.json p7,s7,h4,m1,m2,I2,V4,z2,t4,C1 --api $api --reps 3
  timeout 600 $B $D/m12.json h5,m,H10,w,y,J,S,G10,P0 --api $api --reps 2
done
JET_B200_TBC=stepwise timeout 300 $B $D/m10.json p7,s7,h4,m1,m2,I2 --api tbc --reps 1
} > gpurun_out/r2a_tbc_bench.jsonl 2> gpurun_out/r2a_tbc_bench.err
cat gpurun_out/r2a_tbc_bench.jsonl; tail -5 gpurun_out/r2a_tbc_bench.err
