#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof_chain.py 27 6 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ChainKernel -s 1 -c 1 -o gpurun_out/r2f_chain -f python tools/prof_chain.py 27 6 > gpurun_out/r2f_ncu.log 2>&1; tail -3 gpurun_out/r2f_ncu.log
ls -la gpurun_out/r2f_chain.ncu-rep
