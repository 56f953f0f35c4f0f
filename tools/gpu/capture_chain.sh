#!/bin/bash
# ncu --set full captures of ChainKernel, exported to CSV on the box (the .ncu-rep files exceed what gpurun copies back):
#   1. the dominant launch shape in isolation (tools/prof_chain.py 27 6: rank-27 tensor, 6 fused K = N = 4 steps);
#   2. the heaviest chain launch of one m=20 slice (chain launch 186 of `bench.py --no-graph`).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ChainKernel -s 2 -c 1 -o /tmp/cap_chain -f python tools/prof_chain.py 27 6 > gpurun_out/cap_chain.log 2>&1
ncu -i /tmp/cap_chain.ncu-rep --page raw --csv > gpurun_out/cap_chain_raw.csv 2>/dev/null
ncu -i /tmp/cap_chain.ncu-rep --page source --csv --print-source sass > gpurun_out/cap_chain_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:ChainKernel -s 186 -c 1 -o /tmp/cap_chain_m20 -f python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 0 --no-cpu --no-others --strong-slices 0 > gpurun_out/cap_chain_m20.log 2>&1
ncu -i /tmp/cap_chain_m20.ncu-rep --page raw --csv > gpurun_out/cap_chain_m20_raw.csv 2>/dev/null
ls -la gpurun_out/
